// Standalone hardware probe for the tcgen05 gradient kernels (conv3d data gradient / weight gradient), driven through
// the public C ABI.  Inputs are small dyadic rationals, so every fp32 sum is exact and order-independent: the GPU
// result must equal a plain CPU loop BIT FOR BIT (after the one rounding to the bf16 storage type for dgrad).
//   probe_grad dgrad <case>                    -> PASS/FAIL
//   probe_grad wgrad <case> <x_fmt> <dy_fmt>   -> PASS/FAIL   (formats: 0 = fp16, 1 = bf16)
//   probe_grad time  <case> <iters>            -> TFLOP/s of dgrad and wgrad for one layer shape
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/neuroclear_b200.h"

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)
#define NCK(x)                                                                \
  do {                                                                        \
    if ((x) != 0) {                                                           \
      printf("nc error: %s at %s:%d\n", nc_last_error(), __FILE__, __LINE__); \
      exit(3);                                                                \
    }                                                                         \
  } while (0)

static uint32_t rng_state = 2024u;
static inline uint32_t rnd() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 8;
}

struct Case {
  const char* name;
  int ks, NB, D, H, W, Cin, Cout;
};
static const Case cases[] = {
    {"k3_64_64", 3, 1, 5, 19, 11, 64, 64},      {"k3_128_64", 3, 2, 4, 17, 9, 128, 64},
    {"k3_64_128", 3, 1, 6, 16, 8, 64, 128},     {"k3_256_256", 3, 1, 3, 18, 10, 256, 256},
    {"k5_64_64", 5, 1, 6, 21, 13, 64, 64},      {"k3_64_64_big", 3, 1, 24, 40, 40, 64, 64},
    {"k3_128_128_t", 3, 1, 54, 54, 54, 128, 128}, {"k3_64_64_t", 3, 1, 108, 108, 108, 64, 64},
    {"k5_64_64_t", 5, 1, 108, 108, 108, 64, 64},  {"k3_256_256_t", 3, 1, 27, 27, 27, 256, 256},
    {"k3_128_128_rp", 3, 2, 7, 22, 20, 128, 128}, {"k3_256_256_l2", 3, 1, 9, 37, 37, 256, 256},
};
static const int ncases = sizeof(cases) / sizeof(cases[0]);

static uint16_t to16(float v, int fmt) {
  if (fmt) {
    __nv_bfloat16 b = __float2bfloat16(v);
    return *reinterpret_cast<uint16_t*>(&b);
  }
  __half h = __float2half(v);
  return *reinterpret_cast<uint16_t*>(&h);
}
static float bf16_round(float v) { return __bfloat162float(__float2bfloat16(v)); }
static float bf16_bits_to_float(uint16_t b) {
  uint32_t u = static_cast<uint32_t>(b) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static void fill(std::vector<float>& v, float scale) {
  for (auto& x : v) x = ((int)(rnd() % 9) - 4) * scale;
}

static int check_dgrad(const Case& c) {
  if (c.ks != 3) {
    printf("dgrad %s: skipped (k3 only)\n", c.name);
    return 0;
  }
  const size_t vox = (size_t)c.NB * c.D * c.H * c.W;
  std::vector<float> dy(vox * c.Cout), w((size_t)c.Cout * c.Cin * 27);
  fill(dy, 0.25f);
  fill(w, 0.125f);
  std::vector<uint16_t> dyb(dy.size());
  for (size_t i = 0; i < dy.size(); ++i) dyb[i] = to16(dy[i], 1);
  void *d_dy, *d_w, *d_p, *d_dx;
  CK(cudaMalloc(&d_dy, dyb.size() * 2));
  CK(cudaMalloc(&d_w, w.size() * 4));
  CK(cudaMalloc(&d_p, nc_packed_weight_bytes(c.Cout, c.Cin, 0)));
  CK(cudaMalloc(&d_dx, vox * c.Cin * 2));
  CK(cudaMemcpy(d_dy, dyb.data(), dyb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_dx, 0xff, vox * c.Cin * 2));
  NCK(nc_pack_weights_conv3d_k3_dgrad((const float*)d_w, c.Cout, c.Cin, d_p, nullptr));
  NCK(nc_conv3d_k3_dgrad(d_dy, c.NB, c.D, c.H, c.W, c.Cout, d_p, c.Cin, d_dx, nullptr));
  CK(cudaDeviceSynchronize());
  std::vector<uint16_t> got(vox * c.Cin);
  CK(cudaMemcpy(got.data(), d_dx, got.size() * 2, cudaMemcpyDeviceToHost));
  long long bad = 0;
#pragma omp parallel for reduction(+ : bad)
  for (long long v = 0; v < (long long)vox; ++v) {
    int wv = v % c.W, hv = (v / c.W) % c.H, dv = (v / ((long long)c.W * c.H)) % c.D, nb = v / ((long long)c.W * c.H * c.D);
    std::vector<float> acc(c.Cin, 0.f);
    for (int kd = 0; kd < 3; ++kd)
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          // y[o] += w[tap] x[o + tap - 1]  =>  dx[v] += w[tap] dy[v - tap + 1]
          const int d = dv - kd + 1, h = hv - kh + 1, ww = wv - kw + 1;
          if (d < 0 || d >= c.D || h < 0 || h >= c.H || ww < 0 || ww >= c.W) continue;
          const float* dyr = &dy[((((size_t)nb * c.D + d) * c.H + h) * c.W + ww) * c.Cout];
          const int tap = (kd * 3 + kh) * 3 + kw;
          for (int co = 0; co < c.Cout; ++co) {
            const float g = dyr[co];
            if (g == 0.f) continue;
            const float* wr = &w[((size_t)co * c.Cin) * 27 + tap];
            for (int ci = 0; ci < c.Cin; ++ci) acc[ci] += g * wr[(size_t)ci * 27];
          }
        }
    for (int ci = 0; ci < c.Cin; ++ci)
      if (bf16_bits_to_float(got[v * c.Cin + ci]) != bf16_round(acc[ci])) ++bad;
  }
  printf("dgrad %s (%dx%dx%dx%d, %d->%d): %s (%lld mismatches of %zu)\n", c.name, c.NB, c.D, c.H, c.W, c.Cin, c.Cout,
         bad ? "FAIL" : "PASS", bad, got.size());
  return bad != 0;
}

static int check_wgrad(const Case& c, int xf, int dyf) {
  const size_t vox = (size_t)c.NB * c.D * c.H * c.W;
  const int taps = c.ks * c.ks * c.ks, pad = c.ks / 2;
  std::vector<float> x(vox * c.Cin), dy(vox * c.Cout);
  fill(x, 0.25f);
  fill(dy, 0.25f);
  std::vector<uint16_t> xb(x.size()), dyb(dy.size());
  for (size_t i = 0; i < x.size(); ++i) xb[i] = to16(x[i], xf);
  for (size_t i = 0; i < dy.size(); ++i) dyb[i] = to16(dy[i], dyf);
  void *d_x, *d_dy, *d_s;
  float* d_dw;
  const size_t nw = (size_t)c.Cout * c.Cin * taps;
  const long long sb = nc_conv3d_wgrad_scratch_bytes(c.ks, c.NB, c.D, c.H, c.W, c.Cin, c.Cout);
  CK(cudaMalloc(&d_x, xb.size() * 2));
  CK(cudaMalloc(&d_dy, dyb.size() * 2));
  CK(cudaMalloc(&d_s, sb));
  CK(cudaMalloc(&d_dw, nw * 4));
  CK(cudaMemcpy(d_x, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_dy, dyb.data(), dyb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_s, 0xff, sb));
  CK(cudaMemset(d_dw, 0xff, nw * 4));
  NCK(nc_conv3d_wgrad(d_x, xf, d_dy, dyf, c.NB, c.D, c.H, c.W, c.Cin, c.Cout, c.ks, d_s, d_dw, nullptr));
  CK(cudaDeviceSynchronize());
  std::vector<float> got(nw);
  CK(cudaMemcpy(got.data(), d_dw, nw * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  double worst = 0;
#pragma omp parallel for reduction(+ : bad) reduction(max : worst)
  for (int cot = 0; cot < c.Cout * taps; ++cot) {
    const int co = cot / taps, tap = cot % taps;
    const int kd = tap / (c.ks * c.ks), kh = (tap / c.ks) % c.ks, kw = tap % c.ks;
    std::vector<float> acc(c.Cin, 0.f);
    for (int nb = 0; nb < c.NB; ++nb)
      for (int d = 0; d < c.D; ++d)
        for (int h = 0; h < c.H; ++h)
          for (int w = 0; w < c.W; ++w) {
            const int xd = d + kd - pad, xh = h + kh - pad, xw = w + kw - pad;
            if (xd < 0 || xd >= c.D || xh < 0 || xh >= c.H || xw < 0 || xw >= c.W) continue;
            const float g = dy[((((size_t)nb * c.D + d) * c.H + h) * c.W + w) * c.Cout + co];
            if (g == 0.f) continue;
            const float* xr = &x[((((size_t)nb * c.D + xd) * c.H + xh) * c.W + xw) * c.Cin];
            for (int ci = 0; ci < c.Cin; ++ci) acc[ci] += g * xr[ci];
          }
    for (int ci = 0; ci < c.Cin; ++ci) {
      const float g = got[((size_t)co * c.Cin + ci) * taps + tap];
      if (g != acc[ci]) {
        ++bad;
        const double e = fabs((double)g - acc[ci]);
        if (e > worst || std::isnan(g)) worst = std::isnan(g) ? 1e30 : e;
      }
    }
  }
  printf("wgrad %s k%d (%dx%dx%dx%d, %d->%d) x_fmt %d dy_fmt %d: %s (%lld mismatches of %zu, worst %.4g)\n", c.name,
         c.ks, c.NB, c.D, c.H, c.W, c.Cin, c.Cout, xf, dyf, bad ? "FAIL" : "PASS", bad, nw, worst);
  return bad != 0;
}

static int run_time(const Case& c, int iters) {
  const size_t vox = (size_t)c.NB * c.D * c.H * c.W;
  const int taps = c.ks * c.ks * c.ks;
  void *d_x, *d_dy, *d_s, *d_p, *d_dx, *d_w;
  float* d_dw;
  const size_t nw = (size_t)c.Cout * c.Cin * taps;
  const long long sb = nc_conv3d_wgrad_scratch_bytes(c.ks, c.NB, c.D, c.H, c.W, c.Cin, c.Cout);
  CK(cudaMalloc(&d_x, vox * c.Cin * 2));
  CK(cudaMalloc(&d_dy, vox * c.Cout * 2));
  CK(cudaMalloc(&d_dx, vox * c.Cin * 2));
  CK(cudaMalloc(&d_s, sb));
  CK(cudaMalloc(&d_dw, nw * 4));
  CK(cudaMalloc(&d_w, nw * 4));
  CK(cudaMemset(d_x, 0x11, vox * c.Cin * 2));
  CK(cudaMemset(d_dy, 0x11, vox * c.Cout * 2));
  CK(cudaMemset(d_w, 0, nw * 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const double flop = 2.0 * vox * c.Cin * c.Cout * taps;
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1 && c.ks != 3) break;
    if (pass == 1) {
      CK(cudaMalloc(&d_p, nc_packed_weight_bytes(c.Cout, c.Cin, 0)));
      NCK(nc_pack_weights_conv3d_k3_dgrad((const float*)d_w, c.Cout, c.Cin, d_p, nullptr));
    }
    for (int i = -2; i < iters; ++i) {
      if (i == 0) CK(cudaEventRecord(e0));
      if (pass == 0)
        NCK(nc_conv3d_wgrad(d_x, 1, d_dy, 1, c.NB, c.D, c.H, c.W, c.Cin, c.Cout, c.ks, d_s, d_dw, nullptr));
      else
        NCK(nc_conv3d_k3_dgrad(d_dy, c.NB, c.D, c.H, c.W, c.Cout, d_p, c.Cin, d_dx, nullptr));
    }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("time %s %s: %.3f ms  %.1f TFLOP/s\n", c.name, pass ? "dgrad" : "wgrad", ms / iters,
           flop / (ms / iters * 1e-3) / 1e12);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (getenv("NC_RP")) nc_debug_set_remainder_pairs(atoi(getenv("NC_RP")));   // 0: regular tiles only
  if (argc < 3) {
    printf("usage: probe_grad dgrad|wgrad|time <case> ...\n");
    return 1;
  }
  const int ci = atoi(argv[2]);
  if (ci < 0 || ci >= ncases) return 1;
  if (!strcmp(argv[1], "dgrad")) return check_dgrad(cases[ci]);
  if (!strcmp(argv[1], "wgrad")) return check_wgrad(cases[ci], argc > 3 ? atoi(argv[3]) : 0, argc > 4 ? atoi(argv[4]) : 0);
  if (!strcmp(argv[1], "time")) return run_time(cases[ci], argc > 3 ? atoi(argv[3]) : 5);
  return 1;
}

// Standalone hardware probe for the tcgen05 convolution kernels, driven through the public C ABI.
// Exact-arithmetic inputs (small dyadic rationals) make the fp32 accumulation order-independent, so the GPU
// result must equal a plain CPU direct convolution (rounded once to the fp16 storage type) BIT FOR BIT.  Usage:
//   probe_conv check <case> [max_ctas]           -> prints PASS/FAIL (max_ctas > 0 caps the persistent grid)
//   probe_conv time  <case> <nb> <iters>         -> prints TFLOP/s of one layer shape
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <chrono>
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
#define NCK(x)                                                       \
  do {                                                               \
    if ((x) != 0) {                                                  \
      printf("nc error: %s at %s:%d\n", nc_last_error(), __FILE__, __LINE__); \
      exit(3);                                                       \
    }                                                                \
  } while (0)

static uint32_t rng_state = 12345u;
static inline uint32_t rnd() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 8;
}

struct Case {
  const char* name;
  int transposed, NB, D, H, W, Cin, Cout;
};
static const Case cases[] = {
    {"k3_64_64_small", 0, 1, 5, 20, 12, 64, 64},     {"k3_128_128", 0, 2, 5, 18, 20, 128, 128},
    {"k3_256_256", 0, 1, 9, 18, 10, 256, 256},       {"k3_128_64", 0, 1, 7, 33, 17, 128, 64},
    {"k3_64_128_odd", 0, 1, 3, 5, 3, 64, 128},       {"ct_128_64", 1, 1, 4, 18, 10, 128, 64},
    {"ct_256_128", 1, 2, 3, 17, 9, 256, 128},        {"k3_64_64_deep", 0, 2, 11, 17, 9, 64, 64},
    {"k3_128_64_deep", 0, 1, 9, 20, 20, 128, 64},
    {"k3_128_128_rp6", 0, 2, 7, 22, 20, 128, 128},   {"k3_256_256_rp3", 0, 1, 9, 35, 35, 256, 256},
    {"k3_64_128_rp_small", 0, 1, 6, 5, 9, 64, 128},  {"k3_128_128_l1", 0, 1, 8, 70, 70, 128, 128},
    {"k3_256_256_l2", 0, 1, 35, 35, 35, 256, 256},
};
static const int ncases = sizeof(cases) / sizeof(cases[0]);

static int run_check(const Case& c, int cap) {
  const size_t vox = (size_t)c.NB * c.D * c.H * c.W;
  std::vector<float> x(vox * c.Cin);
  std::vector<__half> xb(x.size());
  for (size_t i = 0; i < x.size(); ++i) {
    x[i] = ((int)(rnd() % 9) - 4) / 4.0f;
    xb[i] = __float2half(x[i]);
  }
  const int taps = c.transposed ? 8 : 27;
  std::vector<float> w((size_t)c.Cout * c.Cin * taps);
  for (auto& v : w) v = ((int)(rnd() % 9) - 4) / 8.0f;
  std::vector<float> bias(c.Cout);
  for (auto& v : bias) v = ((int)(rnd() % 9) - 4) / 2.0f;

  void *dx, *dw, *dp;
  float* dbias;
  CK(cudaMalloc(&dx, xb.size() * 2));
  CK(cudaMalloc(&dw, w.size() * 4));
  const int64_t pbytes = nc_packed_weight_bytes(c.Cout, c.Cin, c.transposed);
  CK(cudaMalloc(&dp, pbytes));
  CK(cudaMalloc(&dbias, c.Cout * 4));
  CK(cudaMemcpy(dx, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, bias.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  long long bad = 0;
  double maxerr = 0;
  if (!c.transposed) {
    NCK(nc_pack_weights_conv3d_k3((const float*)dw, c.Cout, c.Cin, dp, nullptr));
    const int64_t rows = nc_conv3d_k3_stats_rows(c.Cin, c.NB, c.D, c.H, c.W, c.Cout);
    __half* dy;
    float* dst;
    CK(cudaMalloc(&dy, vox * c.Cout * 2));
    CK(cudaMemset(dy, 0xFF, vox * c.Cout * 2));
    CK(cudaMalloc(&dst, rows * 2 * c.Cout * 4));
    NCK(nc_conv3d_k3_fwd(dx, nullptr, c.NB, c.D, c.H, c.W, c.Cin, dp, c.Cout, dy, dst, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<float> y(vox * c.Cout), st(rows * 2 * c.Cout);
    {
      std::vector<__half> yh(vox * c.Cout);
      CK(cudaMemcpy(yh.data(), dy, yh.size() * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < yh.size(); ++i) y[i] = __half2float(yh[i]);
    }
    CK(cudaMemcpy(st.data(), dst, st.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<double> rsum((size_t)c.NB * c.Cout, 0.0), rsq((size_t)c.NB * c.Cout, 0.0);
#pragma omp parallel for collapse(2) reduction(+ : bad) reduction(max : maxerr)
    for (int n = 0; n < c.NB; ++n)
      for (int d = 0; d < c.D; ++d)
        for (int h = 0; h < c.H; ++h)
          for (int ww = 0; ww < c.W; ++ww)
            for (int co = 0; co < c.Cout; ++co) {
              float acc = 0.f;
              for (int kd = 0; kd < 3; ++kd) {
                const int zd = d + kd - 1;
                if (zd < 0 || zd >= c.D) continue;
                for (int kh = 0; kh < 3; ++kh) {
                  const int zh = h + kh - 1;
                  if (zh < 0 || zh >= c.H) continue;
                  for (int kw = 0; kw < 3; ++kw) {
                    const int zw = ww + kw - 1;
                    if (zw < 0 || zw >= c.W) continue;
                    const float* xp = &x[((((size_t)n * c.D + zd) * c.H + zh) * c.W + zw) * c.Cin];
                    const float* wp = &w[(size_t)co * c.Cin * 27 + (kd * 3 + kh) * 3 + kw];
                    for (int ci = 0; ci < c.Cin; ++ci) acc += xp[ci] * wp[(size_t)ci * 27];
                  }
                }
              }
              const float g = y[((((size_t)n * c.D + d) * c.H + h) * c.W + ww) * c.Cout + co];
              const float accr = __half2float(__float2half(acc));   // the kernel stores its exact fp32 sum as fp16
              const double e = fabs((double)g - (double)accr);
              if (!(e == 0.0)) ++bad;
              if (e > maxerr || e != e) maxerr = (e != e) ? 1e30 : e;
            }
    // statistics: compare per-(n, channel) totals
    for (size_t i = 0; i < vox * c.Cout; ++i) {
      const size_t v = i / c.Cout;
      const int n = (int)(v / ((size_t)c.D * c.H * c.W)), co = (int)(i % c.Cout);
      rsum[(size_t)n * c.Cout + co] += y[i];
      rsq[(size_t)n * c.Cout + co] += (double)y[i] * y[i];
    }
    const int64_t rps = rows / c.NB;
    long long sbad = 0;
    for (int n = 0; n < c.NB; ++n)
      for (int co = 0; co < c.Cout; ++co) {
        double s = 0, q = 0;
        for (int64_t r = 0; r < rps; ++r) {
          s += st[((n * rps + r) * 2) * c.Cout + co];
          q += st[((n * rps + r) * 2 + 1) * c.Cout + co];
        }
        // the reference sums are taken over the fp16-rounded outputs, the kernel's over its fp32 accumulators
        if (fabs(s - rsum[(size_t)n * c.Cout + co]) > 3e-3 * (1 + fabs(s)) + 0.5 ||
            fabs(q - rsq[(size_t)n * c.Cout + co]) > 3e-3 * (1 + fabs(q)))
          ++sbad;
      }
    printf("%s cap=%d: %lld/%zu mismatching outputs, max |err| %.4g, stats mismatches %lld -> %s\n", c.name, cap,
           bad, vox * c.Cout, maxerr, sbad, (bad == 0 && sbad == 0) ? "PASS" : "FAIL");
    return (bad == 0 && sbad == 0) ? 0 : 1;
  } else {
    NCK(nc_pack_weights_convT3d_k2s2((const float*)dw, c.Cin, c.Cout, dp, nullptr));
    const int ld = c.Cout * 2, coff = c.Cout;  // write into the upper half of a concat buffer
    const size_t ovox = vox * 8;
    void* dy;
    CK(cudaMalloc(&dy, ovox * ld * 2));
    CK(cudaMemset(dy, 0, ovox * ld * 2));
    NCK(nc_convT3d_k2s2_fwd(dx, nullptr, c.NB, c.D, c.H, c.W, c.Cin, dp, dbias, c.Cout, dy, ld, coff, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<__half> y(ovox * ld);
    CK(cudaMemcpy(y.data(), dy, y.size() * 2, cudaMemcpyDeviceToHost));
#pragma omp parallel for collapse(2) reduction(+ : bad) reduction(max : maxerr)
    for (int n = 0; n < c.NB; ++n)
      for (int od = 0; od < 2 * c.D; ++od)
        for (int oh = 0; oh < 2 * c.H; ++oh)
          for (int ow = 0; ow < 2 * c.W; ++ow) {
            const int d = od / 2, h = oh / 2, ww = ow / 2, tap = ((od & 1) * 2 + (oh & 1)) * 2 + (ow & 1);
            const float* xp = &x[((((size_t)n * c.D + d) * c.H + h) * c.W + ww) * c.Cin];
            const size_t ov = (((size_t)n * 2 * c.D + od) * 2 * c.H + oh) * 2 * c.W + ow;
            for (int co = 0; co < c.Cout; ++co) {
              float acc = 0.f;
              for (int ci = 0; ci < c.Cin; ++ci) acc += xp[ci] * w[((size_t)ci * c.Cout + co) * 8 + tap];
              acc += bias[co];
              const float ref = __half2float(__float2half(acc));
              const float g = __half2float(y[ov * ld + coff + co]);
              const double e = fabs((double)g - (double)ref);
              if (!(e == 0.0)) ++bad;
              if (e > maxerr || e != e) maxerr = (e != e) ? 1e30 : e;
              if (__half2float(y[ov * ld + co]) != 0.f) ++bad;  // lower half must stay untouched
            }
          }
    printf("%s: %lld mismatching outputs, max |err| %.4g -> %s\n", c.name, bad, maxerr, bad == 0 ? "PASS" : "FAIL");
    return bad == 0 ? 0 : 1;
  }
}

struct Shape {
  const char* name;
  int transposed, D, Cin, Cout;
};
static const Shape shapes[] = {
    {"U2  64->64  @140", 0, 140, 64, 64},    {"U3  64->128 @70", 0, 70, 64, 128},   {"U4 128->128 @70", 0, 70, 128, 128},
    {"U5 128->256 @35", 0, 35, 128, 256},    {"U6 256->256 @35", 0, 35, 256, 256},  {"U9 256->128 @70", 0, 70, 256, 128},
    {"U12 128->64 @140", 0, 140, 128, 64},   {"U8 T256->128 @35", 1, 35, 256, 128}, {"U11 T128->64 @70", 1, 70, 128, 64},
};
static const int nshapes = sizeof(shapes) / sizeof(shapes[0]);

static void run_time(const Shape& s, int NB, int iters) {
  const size_t vox = (size_t)NB * s.D * s.D * s.D;
  void *dx, *dp;
  CK(cudaMalloc(&dx, vox * s.Cin * 2));
  CK(cudaMemset(dx, 0, vox * s.Cin * 2));
  const int64_t pbytes = nc_packed_weight_bytes(s.Cout, s.Cin, s.transposed);
  CK(cudaMalloc(&dp, pbytes));
  CK(cudaMemset(dp, 0, pbytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms = 0;
  double flop;
  if (!s.transposed) {
    const int64_t rows = nc_conv3d_k3_stats_rows(s.Cin, NB, s.D, s.D, s.D, s.Cout);
    __half* dy;
    float* dst;
    CK(cudaMalloc(&dy, vox * s.Cout * 2));
    CK(cudaMalloc(&dst, rows * 2 * s.Cout * 4));
    for (int i = 0; i < 2; ++i) NCK(nc_conv3d_k3_fwd(dx, nullptr, NB, s.D, s.D, s.D, s.Cin, dp, s.Cout, dy, dst, nullptr));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) NCK(nc_conv3d_k3_fwd(dx, nullptr, NB, s.D, s.D, s.D, s.Cin, dp, s.Cout, dy, dst, nullptr));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    flop = 2.0 * vox * s.Cout * s.Cin * 27;
    cudaFree(dy), cudaFree(dst);
  } else {
    void* dy;
    float* db;
    CK(cudaMalloc(&dy, vox * 8 * s.Cout * 2 * 2));
    CK(cudaMalloc(&db, s.Cout * 4));
    CK(cudaMemset(db, 0, s.Cout * 4));
    for (int i = 0; i < 2; ++i)
      NCK(nc_convT3d_k2s2_fwd(dx, nullptr, NB, s.D, s.D, s.D, s.Cin, dp, db, s.Cout, dy, 2 * s.Cout, s.Cout, nullptr));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i)
      NCK(nc_convT3d_k2s2_fwd(dx, nullptr, NB, s.D, s.D, s.D, s.Cin, dp, db, s.Cout, dy, 2 * s.Cout, s.Cout, nullptr));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    flop = 2.0 * vox * s.Cout * s.Cin * 8;
    cudaFree(dy), cudaFree(db);
  }
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  printf("TIME %-18s NB=%d  %.3f ms  %.1f TFLOP/s\n", s.name, NB, ms, flop / ms * 1e-9);
  cudaFree(dx), cudaFree(dp);
}

int main(int argc, char** argv) {
  if (getenv("NC_RP")) nc_debug_set_remainder_pairs(atoi(getenv("NC_RP")));   // 0: regular tiles only
  if (argc >= 3 && !strcmp(argv[1], "check")) {
    const int ci = atoi(argv[2]);
    const int cap = argc > 3 ? atoi(argv[3]) : 0;   // persistent-grid cap: > 0 forces several tiles per CTA
    if (ci < 0 || ci >= ncases) return 4;
    nc_debug_set_max_ctas(cap);
    return run_check(cases[ci], cap);
  }
  if (argc >= 3 && !strcmp(argv[1], "time")) {
    const int si = atoi(argv[2]);
    const int nb = argc > 3 ? atoi(argv[3]) : 1;
    const int iters = argc > 4 ? atoi(argv[4]) : 5;
    if (si < 0 || si >= nshapes) return 4;
    run_time(shapes[si], nb, iters);
    return 0;
  }
  printf("cases: %d, shapes: %d\n", ncases, nshapes);
  return 0;
}

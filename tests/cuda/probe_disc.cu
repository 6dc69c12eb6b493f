// Timing probe for the 2-D PatchGAN convolution kernels (csrc/disc2d.cu) on the shapes of one apollo iteration at a
// 108^3 crop.  Built twice by tests/cuda/run_probe_disc.sh: against the tree's disc2d.cu and (-DOLD) against the
// previous revision, so a change is measured A/B in one GPU call.  Per layer and direction: warm time (50 launches
// back to back between two events) and cold time (L2 flushed before every launch, single-launch events, min of 5).
//   usage: probe_disc [cluster override 0|1|2|4|8] [quick]   (quick: one launch per case, for compute-sanitizer)
#include <cstdio>
#include <cstdlib>
#include <vector>
#ifdef OLD
#include "../../build/old/disc2d_old.cu"
#else
#include "../../neuroclear_b200/csrc/disc2d.cu"
#endif
#include "../../neuroclear_b200/csrc/runtime.cu"

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      printf("%s: %s\n", #x, cudaGetErrorString(e_));                             \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

struct Layer { int cin, cout, h, stride; };

static bool g_quick = false;  // one launch per case, no timing (compute-sanitizer runs)

template <class F>
static int time_it(const char* tag, F f, float* flush, size_t flush_bytes) {
  if (g_quick) {
    if (f()) { printf("%s: %s\n", tag, nc::last_error()); return 1; }
    CK(cudaDeviceSynchronize());
    printf("%-28s ran\n", tag);
    return 0;
  }
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i)
    if (f()) { printf("%s: %s\n", tag, nc::last_error()); return 1; }
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < 50; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float warm;
  CK(cudaEventElapsedTime(&warm, a, b));
  float cold = 1e9f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaMemsetAsync(flush, r, flush_bytes));
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float t;
    CK(cudaEventElapsedTime(&t, a, b));
    if (t < cold) cold = t;
  }
  printf("%-28s warm %7.1f us   cold %7.1f us\n", tag, warm * 1000.f / 50, cold * 1000.f);
  return 0;
}

int main(int argc, char** argv) {
#ifndef OLD
  if (argc > 1) nc::debug_set_disc_cluster(atoi(argv[1]));
#endif
  g_quick = argc > 2;
  const Layer L[5] = {{1, 64, 108, 2}, {64, 128, 54, 2}, {128, 256, 27, 2}, {256, 512, 13, 1}, {512, 1, 12, 1}};
  const size_t flush_bytes = 512u << 20;
  float* flush;
  CK(cudaMalloc(&flush, flush_bytes));
  for (int N : {1, 2, 4}) {
    if (g_quick && N == 4) break;
    for (int l = 0; l < 5; ++l) {
      const Layer& y = L[l];
      const int ho = (y.h - 2) / y.stride + 1;
      const size_t nx = (size_t)N * y.cin * y.h * y.h, ny = (size_t)N * y.cout * ho * ho, nw = (size_t)y.cout * y.cin * 16;
      float *x, *w, *b, *o, *dx, *dw, *db;
      CK(cudaMalloc(&x, nx * 4)); CK(cudaMalloc(&w, nw * 4)); CK(cudaMalloc(&b, y.cout * 4)); CK(cudaMalloc(&o, ny * 4));
      CK(cudaMalloc(&dx, nx * 4)); CK(cudaMalloc(&dw, nw * 4)); CK(cudaMalloc(&db, y.cout * 4));
      std::vector<float> hx(nx), hw(nw), hy(ny);
      for (size_t i = 0; i < nx; ++i) hx[i] = (float)((i * 2654435761u) % 1024) / 1024.f - 0.5f;
      for (size_t i = 0; i < nw; ++i) hw[i] = (float)((i * 40503u) % 512) / 4096.f - 0.06f;
      for (size_t i = 0; i < ny; ++i) hy[i] = (float)((i * 69069u) % 256) / 256.f - 0.5f;
      CK(cudaMemcpy(x, hx.data(), nx * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(w, hw.data(), nw * 4, cudaMemcpyHostToDevice));
      CK(cudaMemset(b, 0, y.cout * 4));
      char tag[64];
      snprintf(tag, sizeof tag, "N%d L%d %d->%d %d^2 fwd", N, l + 1, y.cin, y.cout, y.h);
      if (time_it(tag, [&] { return nc::conv2d_k4_fwd(x, w, b, N, y.cin, y.h, y.h, y.cout, y.stride, 1.f, o, 0); }, flush, flush_bytes)) return 1;
      CK(cudaMemcpy(o, hy.data(), ny * 4, cudaMemcpyHostToDevice));
      snprintf(tag, sizeof tag, "N%d L%d %d->%d %d^2 dgrad", N, l + 1, y.cin, y.cout, y.h);
      if (time_it(tag, [&] { return nc::conv2d_k4_dgrad(o, w, N, y.cin, y.h, y.h, y.cout, y.stride, dx, 0); }, flush, flush_bytes)) return 1;
      snprintf(tag, sizeof tag, "N%d L%d %d->%d %d^2 wgrad", N, l + 1, y.cin, y.cout, y.h);
      if (time_it(tag, [&] { return nc::conv2d_k4_wgrad(x, o, N, y.cin, y.h, y.h, y.cout, y.stride, dw, db, 0); }, flush, flush_bytes)) return 1;
      // checksums so that the A/B builds can be compared for equality
      std::vector<float> r(nw);
      CK(cudaMemcpy(r.data(), dw, nw * 4, cudaMemcpyDeviceToHost));
      double cs = 0;
      for (size_t i = 0; i < nw; ++i) cs += r[i] * (double)((i % 7) + 1);
      std::vector<float> rx(nx);
      CK(cudaMemcpy(rx.data(), dx, nx * 4, cudaMemcpyDeviceToHost));
      double cx = 0;
      for (size_t i = 0; i < nx; ++i) cx += rx[i] * (double)((i % 5) + 1);
      printf("   checksum dw %.9e dx %.9e\n", cs, cx);
      cudaFree(x); cudaFree(w); cudaFree(b); cudaFree(o); cudaFree(dx); cudaFree(dw); cudaFree(db);
    }
  }
  return 0;
}

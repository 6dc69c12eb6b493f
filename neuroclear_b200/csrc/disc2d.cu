// 2-D PatchGAN discriminator path of the apollo model (reference models/networks.py:1009-1067 NLayerDiscriminator
// with dimension=2, built at models/axial_to_lateral_gan_apollo_model.py:99-123) and the LSGAN loss
// (networks.py:252-319): forward AND backward kernels.
//
// One discriminator pass sees ONE 2-D image of ~108^2..148^2 pixels (a slice or a max-intensity projection of the
// cube): ~1 GFLOP, 18 passes + backward per training step = ~0.1 % of the step's FLOPs.  The path is bound by launch
// latency and by the small problem size, not by peak throughput, so these are fp32 CUDA-core kernels
// on NCHW tensors (register-tiled implicit GEMMs, no tensor-core staging, no layout changes), one launch per layer and
// direction; patchgan.cu strings them into one call per pass.
#include "internal.h"

namespace nc {

// ------------------------------------------------------------------------------------------------ Conv2d k4 p1
// All three directions are fp32 implicit GEMMs on CUDA cores with the classic register tiling: a block of 256 threads
// owns a 64 x 64 output tile, every thread a 4 x 4 patch; the K dimension is walked in chunks of 16 staged in shared
// memory ([k][m] / [k][n], so the inner product reads float4s) with the next chunk prefetched into registers.
// The images are tiny (M = 121..2916 pixels) while K reaches 8192, so a layer has only 2..46 output tiles: the K
// range of a tile is therefore SPLIT ACROSS A THREAD-BLOCK CLUSTER (up to 8 CTAs) and the partial tiles are summed
// by the cluster's rank 0 through distributed shared memory, in rank order -> bitwise repeatable, no scratch buffer.
// (One pass of the apollo discriminators is ~1 GFLOP forward; there is not enough work for tensor-core tiles to
// pay off, what matters is parallelism and latency.)
constexpr int TB = 64;  // tile edge (M and N)
constexpr int KB = 16;  // K chunk

__device__ __forceinline__ void tile_fma(const float (*As)[TB + 4], const float (*Bs)[TB + 4], int tm, int tn,
                                         float (&acc)[4][4]) {
#pragma unroll
  for (int k = 0; k < KB; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
    const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) acc[i][jn] = fmaf(av[i], bv[jn], acc[i][jn]);
  }
}

__device__ __forceinline__ unsigned cluster_rank_x() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_size_x() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Sum the 4 x 4 accumulators of all CTAs of the cluster into rank 0's registers (ranks added in order 1, 2, ...).
// `slots` = 256 x 16 floats of this CTA's shared memory.  Every CTA of the cluster must call it.
__device__ __forceinline__ void cluster_reduce(float (&acc)[4][4], float* slots, unsigned rank, unsigned size) {
  if (size == 1) return;
  float* mine = slots + threadIdx.x * 16;
  if (rank != 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) mine[i] = acc[i >> 2][i & 3];
  }
  cluster_sync_all();
  if (rank == 0) {
    const uint32_t local = static_cast<uint32_t>(__cvta_generic_to_shared(mine));
    for (unsigned r = 1; r < size; ++r) {
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 v;
        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(remote + q * 16));
        acc[q][0] += v.x, acc[q][1] += v.y, acc[q][2] += v.z, acc[q][3] += v.w;
      }
    }
  }
  cluster_sync_all();  // remote shared memory must stay alive until rank 0 has read it
}

// y[n,co,ho,wo] = b[co] + sum_{ci,kh,kw} w[co,ci,kh,kw] * x[n,ci,ho*s-1+kh,wo*s-1+kw]   (+ optional LeakyReLU)
// GEMM: M = output pixel, N = co, K = (ci, tap); one K chunk = the 16 taps of one input channel.
// grid.x = M tiles x cluster size (the cluster splits the input channels).
__global__ void __launch_bounds__(256)
conv2d_k4_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int Cin,
                     int H, int W, int Cout, int Ho, int Wo, int stride, float slope, float* __restrict__ y) {
  __shared__ __align__(16) float As[KB][TB + 4], Bs[KB][TB + 4];
  __shared__ __align__(16) float slots[256 * 16];
  const unsigned rank = cluster_rank_x(), csize = cluster_size_x();
  const int t = threadIdx.x, tm = t & 15, tn = t >> 4;
  const int p0 = (blockIdx.x / csize) * TB, co0 = blockIdx.y * TB, n = blockIdx.z;
  const int P = Ho * Wo;
  // loader roles: A: pixel t % 64, filter row kh = t / 64 (4 kw each); B: channel t / 4, taps (t % 4) * 4 .. + 3
  const int ap = p0 + (t & 63), akh = t >> 6;
  const int aho = ap / Wo, awo = ap - aho * Wo;
  const int ah = aho * stride - 1 + akh, aw0 = awo * stride - 1;
  const bool a_row_ok = ap < P && ah >= 0 && ah < H;
  const int bco = co0 + (t >> 2), bt = (t & 3) * 4;
  const float* xn = x + static_cast<size_t>(n) * Cin * H * W;
  const int c_begin = static_cast<int>(static_cast<long long>(Cin) * rank / csize);
  const int c_end = static_cast<int>(static_cast<long long>(Cin) * (rank + 1) / csize);
  float acc[4][4] = {};
  float ra[4];
  float4 rb;
  auto fetch = [&](int ci) {
    const float* xr = xn + (static_cast<size_t>(ci) * H + ah) * W;
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) {
      const int ww = aw0 + kw;
      ra[kw] = (a_row_ok && ww >= 0 && ww < W) ? __ldg(xr + ww) : 0.f;
    }
    rb = bco < Cout ? __ldg(reinterpret_cast<const float4*>(w + (static_cast<size_t>(bco) * Cin + ci) * 16 + bt))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (c_begin < c_end) fetch(c_begin);
  for (int ci = c_begin; ci < c_end; ++ci) {
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) As[akh * 4 + kw][t & 63] = ra[kw];
    Bs[bt][t >> 2] = rb.x, Bs[bt + 1][t >> 2] = rb.y, Bs[bt + 2][t >> 2] = rb.z, Bs[bt + 3][t >> 2] = rb.w;
    __syncthreads();
    if (ci + 1 < c_end) fetch(ci + 1);
    tile_fma(As, Bs, tm, tn, acc);
    __syncthreads();
  }
  cluster_reduce(acc, slots, rank, csize);
  if (rank != 0) return;
#pragma unroll
  for (int jn = 0; jn < 4; ++jn) {
    const int co = co0 + tn * 4 + jn;
    if (co >= Cout) continue;
    const float bias = b ? __ldg(b + co) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + tm * 4 + i;
      if (p >= P) continue;
      float v = acc[i][jn] + bias;
      if (slope != 1.f) v = v > 0.f ? v : v * slope;
      y[(static_cast<size_t>(n) * Cout + co) * P + p] = v;
    }
  }
}

// dx[n,ci,h,w] = sum_{co,kh,kw : (h+1-kh) = ho*s, (w+1-kw) = wo*s} w[co,ci,kh,kw] * dy[n,co,ho,wo]
// GEMM: M = input pixel, N = ci, K = (co, tap); one K chunk = the 16 taps of one output channel (split over the
// cluster).
__global__ void __launch_bounds__(256)
conv2d_k4_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, int Cin, int H, int W, int Cout,
                       int Ho, int Wo, int stride, float* __restrict__ dx) {
  __shared__ __align__(16) float As[KB][TB + 4], Bs[KB][TB + 4];
  __shared__ __align__(16) float slots[256 * 16];
  const unsigned rank = cluster_rank_x(), csize = cluster_size_x();
  const int t = threadIdx.x, tm = t & 15, tn = t >> 4;
  const int p0 = (blockIdx.x / csize) * TB, ci0 = blockIdx.y * TB, n = blockIdx.z;
  const int P = H * W;
  const int ap = p0 + (t & 63), akh = t >> 6;
  const int ah = ap / W, aw = ap - ah * W;
  const int th = ah + 1 - akh;
  const bool h_ok = ap < P && th >= 0 && th % stride == 0 && th / stride < Ho;
  const int ho = h_ok ? th / stride : 0;
  int wos[4];
#pragma unroll
  for (int kw = 0; kw < 4; ++kw) {
    const int tw = aw + 1 - kw;
    wos[kw] = (h_ok && tw >= 0 && tw % stride == 0 && tw / stride < Wo) ? tw / stride : -1;
  }
  const int bci = ci0 + (t >> 2), bt = (t & 3) * 4;
  const float* dyn = dy + static_cast<size_t>(n) * Cout * Ho * Wo;
  const int c_begin = static_cast<int>(static_cast<long long>(Cout) * rank / csize);
  const int c_end = static_cast<int>(static_cast<long long>(Cout) * (rank + 1) / csize);
  float acc[4][4] = {};
  float ra[4];
  float4 rb;
  auto fetch = [&](int co) {
    const float* dr = dyn + (static_cast<size_t>(co) * Ho + ho) * Wo;
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) ra[kw] = wos[kw] >= 0 ? __ldg(dr + wos[kw]) : 0.f;
    rb = bci < Cin ? __ldg(reinterpret_cast<const float4*>(w + (static_cast<size_t>(co) * Cin + bci) * 16 + bt))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (c_begin < c_end) fetch(c_begin);
  for (int co = c_begin; co < c_end; ++co) {
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) As[akh * 4 + kw][t & 63] = ra[kw];
    Bs[bt][t >> 2] = rb.x, Bs[bt + 1][t >> 2] = rb.y, Bs[bt + 2][t >> 2] = rb.z, Bs[bt + 3][t >> 2] = rb.w;
    __syncthreads();
    if (co + 1 < c_end) fetch(co + 1);
    tile_fma(As, Bs, tm, tn, acc);
    __syncthreads();
  }
  cluster_reduce(acc, slots, rank, csize);
  if (rank != 0) return;
#pragma unroll
  for (int jn = 0; jn < 4; ++jn) {
    const int ci = ci0 + tn * 4 + jn;
    if (ci >= Cin) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + tm * 4 + i;
      if (p < P) dx[(static_cast<size_t>(n) * Cin + ci) * P + p] = acc[i][jn];
    }
  }
}

// dw[co,ci,kh,kw] = sum_{n,ho,wo} dy[n,co,ho,wo] * x[n,ci,ho*s-1+kh,wo*s-1+kw]
// GEMM: M = co, N = (ci, tap) (a 64-wide tile = 4 input channels x 16 taps), K = (n, output pixel) in chunks of 16
// (the pixel chunks of an image are split over the cluster).  grid.x = M tiles x cluster size.
__global__ void __launch_bounds__(256)
conv2d_k4_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int N, int Cin, int H, int W,
                       int Cout, int Ho, int Wo, int stride, float* __restrict__ dw) {
  __shared__ __align__(16) float As[KB][TB + 4], Bs[KB][TB + 4];
  __shared__ __align__(16) float slots[256 * 16];
  const unsigned rank = cluster_rank_x(), csize = cluster_size_x();
  const int t = threadIdx.x, tm = t & 15, tn = t >> 4;
  const int co0 = (blockIdx.x / csize) * TB, j0 = blockIdx.y * TB;  // j = ci * 16 + tap
  const int P = Ho * Wo;
  // loader roles: A: channel t / 4, pixels (t % 4) * 4 .. + 3 of the chunk; B: pixel t % 16, columns (t / 16) * 4 .. + 3
  const int aco = co0 + (t >> 2), apk = (t & 3) * 4;
  const int bpk = t & 15, bj = (t >> 4) * 4;
  const int chunks = (P + KB - 1) / KB;
  const int k_begin = static_cast<int>(static_cast<long long>(chunks) * rank / csize);
  const int k_end = static_cast<int>(static_cast<long long>(chunks) * (rank + 1) / csize);
  float acc[4][4] = {};
  float ra[4], rb[4];
  for (int n = 0; n < N; ++n) {
    const float* dyc = dy + (static_cast<size_t>(n) * Cout + aco) * P;
    const float* xn = x + static_cast<size_t>(n) * Cin * H * W;
    auto fetch = [&](int kc) {
      const int pc = kc * KB;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int p = pc + apk + i;
        ra[i] = (aco < Cout && p < P) ? __ldg(dyc + p) : 0.f;
      }
      const int p = pc + bpk;
      const int ho = p / Wo, wo = p - ho * Wo;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = j0 + bj + i;
        const int ci = j >> 4, tap = j & 15;
        const int hh = ho * stride - 1 + (tap >> 2), ww = wo * stride - 1 + (tap & 3);
        const bool ok = p < P && ci < Cin && hh >= 0 && hh < H && ww >= 0 && ww < W;
        rb[i] = ok ? __ldg(xn + (static_cast<size_t>(ci) * H + hh) * W + ww) : 0.f;
      }
    };
    if (k_begin < k_end) fetch(k_begin);
    for (int kc = k_begin; kc < k_end; ++kc) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        As[apk + i][t >> 2] = ra[i];
        Bs[bpk][bj + i] = rb[i];
      }
      __syncthreads();
      if (kc + 1 < k_end) fetch(kc + 1);
      tile_fma(As, Bs, tm, tn, acc);
      __syncthreads();
    }
  }
  cluster_reduce(acc, slots, rank, csize);
  if (rank != 0) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + tm * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
      const int j = j0 + tn * 4 + jn;
      if (j < Cin * 16) dw[static_cast<size_t>(co) * Cin * 16 + j] = acc[i][jn];
    }
  }
}

// Cluster size (1, 2, 4 or 8) for a launch of `tiles` output tiles whose K loop has `k_chunks` iterations: split K
// until the grid fills the GPU or a CTA would be left with fewer than 8 chunks.
static int pick_cluster(long long tiles, int k_chunks) {
  int c = 1;
  while (c < 8 && tiles * c < 2 * num_sms() && k_chunks / (2 * c) >= 8) c *= 2;
  return c;
}

template <class Kern, class... Args>
static int launch_clustered(Kern kern, dim3 grid, int cluster, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NC_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
  return 0;
}

// db[co] = sum_{n,p} dy[n,co,p]; one block per channel, fixed-order tree
__global__ void __launch_bounds__(128)
channel_sum_kernel(const float* __restrict__ dy, int N, int C, int P, float* __restrict__ db) {
  __shared__ float red[128];
  const int c = blockIdx.x;
  float s = 0.f;
  for (int n = 0; n < N; ++n)
    for (int p = threadIdx.x; p < P; p += 128) s += dy[(static_cast<size_t>(n) * C + c) * P + p];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 64; o >= 1; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) db[c] = red[0];
}

static int out_extent(int in, int stride) { return (in + 2 - 4) / stride + 1; }

int conv2d_k4_fwd(const float* x, const float* w, const float* b, int N, int Cin, int H, int W, int Cout, int stride,
                  float slope, float* y, cudaStream_t stream) {
  if (stride != 1 && stride != 2) return set_error("conv2d_k4: stride must be 1 or 2");
  const int Ho = out_extent(H, stride), Wo = out_extent(W, stride);
  if (Ho <= 0 || Wo <= 0) return set_error("conv2d_k4: input too small");
  if (Cout > 65535 || N > 65535) return set_error("conv2d_k4: too many channels / images");
  const int mt = (Ho * Wo + TB - 1) / TB, nt = (Cout + TB - 1) / TB;
  const int cl = pick_cluster(static_cast<long long>(mt) * nt * N, Cin);
  return launch_clustered(conv2d_k4_fwd_kernel, dim3(mt * cl, nt, N), cl, stream, x, w, b, Cin, H, W, Cout, Ho, Wo,
                          stride, slope, y);
}
int conv2d_k4_dgrad(const float* dy, const float* w, int N, int Cin, int H, int W, int Cout, int stride, float* dx,
                    cudaStream_t stream) {
  if (stride != 1 && stride != 2) return set_error("conv2d_k4: stride must be 1 or 2");
  const int Ho = out_extent(H, stride), Wo = out_extent(W, stride);
  if (Cin > 65535 || N > 65535) return set_error("conv2d_k4: too many channels / images");
  const int mt = (H * W + TB - 1) / TB, nt = (Cin + TB - 1) / TB;
  const int cl = pick_cluster(static_cast<long long>(mt) * nt * N, Cout);
  return launch_clustered(conv2d_k4_dgrad_kernel, dim3(mt * cl, nt, N), cl, stream, dy, w, Cin, H, W, Cout, Ho, Wo,
                          stride, dx);
}
int conv2d_k4_wgrad(const float* x, const float* dy, int N, int Cin, int H, int W, int Cout, int stride, float* dw,
                    float* db, cudaStream_t stream) {
  if (stride != 1 && stride != 2) return set_error("conv2d_k4: stride must be 1 or 2");
  const int Ho = out_extent(H, stride), Wo = out_extent(W, stride);
  const int mt = (Cout + TB - 1) / TB, nt = (Cin * 16 + TB - 1) / TB;
  const int cl = pick_cluster(static_cast<long long>(mt) * nt, (Ho * Wo + KB - 1) / KB);
  if (int rc = launch_clustered(conv2d_k4_wgrad_kernel, dim3(mt * cl, nt), cl, stream, x, dy, N, Cin, H, W, Cout, Ho,
                                Wo, stride, dw))
    return rc;
  if (db) {
    channel_sum_kernel<<<Cout, 128, 0, stream>>>(dy, N, Cout, Ho * Wo, db);
    NC_CUDA(cudaGetLastError());
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ IN2d + LeakyReLU
// One block per (n, c) plane: y = lrelu((x - mean) * rstd), statistics saved for the backward pass.
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  // fixed-order tree over 256 threads; result broadcast to all threads
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float r = red[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256)
in2d_lrelu_fwd_kernel(const float* __restrict__ x, int P, float eps, float slope, float* __restrict__ y,
                      float* __restrict__ mean_rstd) {
  __shared__ float red[256];
  const size_t base = static_cast<size_t>(blockIdx.x) * P;
  float s = 0.f;
  for (int p = threadIdx.x; p < P; p += 256) s += x[base + p];
  const float mean = block_sum_256(s, red) / P;
  float q = 0.f;
  for (int p = threadIdx.x; p < P; p += 256) {
    const float d = x[base + p] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = 1.0f / sqrtf(block_sum_256(q, red) / P + eps);
  for (int p = threadIdx.x; p < P; p += 256) {
    const float v = (x[base + p] - mean) * rstd;
    y[base + p] = v > 0.f ? v : v * slope;
  }
  if (threadIdx.x == 0) {
    mean_rstd[2 * blockIdx.x] = mean;
    mean_rstd[2 * blockIdx.x + 1] = rstd;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * lrelu'(xhat)
__global__ void __launch_bounds__(256)
in2d_lrelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean_rstd,
                      int P, float slope, float* __restrict__ dx) {
  __shared__ float red[256];
  const size_t base = static_cast<size_t>(blockIdx.x) * P;
  const float mean = mean_rstd[2 * blockIdx.x], rstd = mean_rstd[2 * blockIdx.x + 1];
  float s1 = 0.f, s2 = 0.f;
  for (int p = threadIdx.x; p < P; p += 256) {
    const float xh = (x[base + p] - mean) * rstd;
    const float g = dy[base + p] * (xh > 0.f ? 1.f : slope);
    s1 += g;
    s2 = fmaf(g, xh, s2);
  }
  const float m1 = block_sum_256(s1, red) / P;
  const float m2 = block_sum_256(s2, red) / P;
  for (int p = threadIdx.x; p < P; p += 256) {
    const float xh = (x[base + p] - mean) * rstd;
    const float g = dy[base + p] * (xh > 0.f ? 1.f : slope);
    dx[base + p] = rstd * (g - m1 - xh * m2);
  }
}

int in2d_lrelu_fwd(const float* x, int NC, int P, float eps, float slope, float* y, float* mean_rstd,
                   cudaStream_t stream) {
  in2d_lrelu_fwd_kernel<<<NC, 256, 0, stream>>>(x, P, eps, slope, y, mean_rstd);
  NC_CUDA(cudaGetLastError());
  return 0;
}
int in2d_lrelu_bwd(const float* dy, const float* x, const float* mean_rstd, int NC, int P, float slope, float* dx,
                   cudaStream_t stream) {
  in2d_lrelu_bwd_kernel<<<NC, 256, 0, stream>>>(dy, x, mean_rstd, P, slope, dx);
  NC_CUDA(cudaGetLastError());
  return 0;
}

// LeakyReLU backward from the OUTPUT sign (the forward is fused into the conv): dx = dy * (y > 0 ? 1 : slope)
__global__ void lrelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n, float slope,
                                 float* __restrict__ dx) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll)
    dx[i] = dy[i] * (y[i] > 0.f ? 1.f : slope);
}
int lrelu_bwd(const float* dy, const float* y, long long n, float slope, float* dx, cudaStream_t stream) {
  const int blocks = static_cast<int>((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  lrelu_bwd_kernel<<<blocks, 256, 0, stream>>>(dy, y, n, slope, dx);
  NC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ losses
// mode 0: MSE against a constant target (GANLoss lsgan, networks.py:275-276,311-313): mean((p - t)^2)
// mode 1: L1 against a tensor (torch.nn.L1Loss, apollo_model.py:128,279):               mean(|p - q|)
// One thread-block cluster (1 CTA for the small prediction maps, 8 CTAs for volumes): every thread strides over the
// cluster's elements, block sums are tree-reduced in shared memory, and rank 0 adds the CTAs' sums through
// distributed shared memory in rank order -> bitwise repeatable, no scratch buffer, one launch.  (A single 256-thread
// block took ~0.4 ms for the L1 cycle term of a 108^3 crop: it sits in the critical path of every iteration.)
__global__ void __launch_bounds__(256)
loss_fwd_kernel(const float* __restrict__ p, const float* __restrict__ q, float target, long long n, int mode,
                float* __restrict__ loss) {
  __shared__ double red[256];
  __shared__ double block_sum;
  const unsigned rank = cluster_rank_x(), csize = cluster_size_x();
  double s = 0.0;
  auto term = [&](float pv, float qv) {
    const float d = pv - (mode == 0 ? target : qv);
    return mode == 0 ? static_cast<double>(d) * d : fabs(static_cast<double>(d));
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(q)) & 15) == 0;
  const long long n4 = vec ? (n >> 2) : 0;
  const float4* p4 = reinterpret_cast<const float4*>(p);
  const float4* q4 = reinterpret_cast<const float4*>(q);
#pragma unroll 4
  for (long long i = static_cast<long long>(rank) * 256 + threadIdx.x; i < n4; i += 256ll * csize) {
    const float4 a = __ldg(p4 + i);
    const float4 b = mode == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(q4 + i);
    s += (term(a.x, b.x) + term(a.y, b.y)) + (term(a.z, b.z) + term(a.w, b.w));
  }
  for (long long i = (n4 << 2) + static_cast<long long>(rank) * 256 + threadIdx.x; i < n; i += 256ll * csize)
    s += term(p[i], mode == 0 ? 0.f : q[i]);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sum = red[0];
  if (csize == 1) {
    if (threadIdx.x == 0) *loss = static_cast<float>(red[0] / static_cast<double>(n));
    return;
  }
  cluster_sync_all();
  if (rank == 0 && threadIdx.x == 0) {
    double total = block_sum;
    const unsigned local = static_cast<unsigned>(__cvta_generic_to_shared(&block_sum));
    for (unsigned r = 1; r < csize; ++r) {
      unsigned remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
      double v;
      asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote));
      total += v;
    }
    *loss = static_cast<float>(total / static_cast<double>(n));
  }
  cluster_sync_all();  // remote shared memory must stay alive until rank 0 has read it
}
// dp = upstream * d loss / dp
__global__ void loss_bwd_kernel(const float* __restrict__ p, const float* __restrict__ q, float target, long long n,
                                int mode, const float* __restrict__ upstream, float* __restrict__ dp) {
  const float g = *upstream / static_cast<float>(n);
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    const float d = p[i] - (mode == 0 ? target : q[i]);
    dp[i] = mode == 0 ? 2.f * d * g : (d > 0.f ? g : (d < 0.f ? -g : 0.f));
  }
}
int loss_fwd(const float* p, const float* q, float target, long long n, int mode, float* loss, cudaStream_t stream) {
  if (mode != 0 && mode != 1) return set_error("loss: mode must be 0 (mse vs constant) or 1 (l1 vs tensor)");
  // with 8 CTAs of 256 threads a 108^3 volume is ~600 strided iterations per thread instead of ~4900
  return launch_clustered(loss_fwd_kernel, dim3(n >= (1 << 16) ? 8 : 1), n >= (1 << 16) ? 8 : 1, stream, p, q, target,
                          n, mode, loss);
}
int loss_bwd(const float* p, const float* q, float target, long long n, int mode, const float* upstream, float* dp,
             cudaStream_t stream) {
  if (mode != 0 && mode != 1) return set_error("loss: mode must be 0 (mse vs constant) or 1 (l1 vs tensor)");
  const int blocks = static_cast<int>((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  loss_bwd_kernel<<<blocks, 256, 0, stream>>>(p, q, target, n, mode, upstream, dp);
  NC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam (apollo_model.py:131-136; no weight decay, no amsgrad), one launch per parameter tensor:
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, float beta1, float beta2, float step_size,
                                 float inv_sqrt_bc2, float eps) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    const float gi = g[i];
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}
int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
              int step, cudaStream_t stream) {
  if (step < 1) return set_error("adam_step: step counts from 1");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step), bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  const int blocks = static_cast<int>((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  adam_step_kernel<<<blocks, 256, 0, stream>>>(p, g, m, v, n, beta1, beta2, static_cast<float>(lr / bc1),
                                               static_cast<float>(1.0 / sqrt(bc2)), eps);
  NC_CUDA(cudaGetLastError());
  return 0;
}

// All parameter tensors of an optimiser in ONE launch: `table` is a device array of `count` entries
// {p, g, m, v, n} (5 x 8 bytes); blockIdx.y walks the tensors.  Same arithmetic as adam_step_kernel.
struct AdamEntry {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
};
__global__ void adam_step_multi_kernel(const AdamEntry* __restrict__ table, float beta1, float beta2, float step_size,
                                       float inv_sqrt_bc2, float eps) {
  const AdamEntry e = table[blockIdx.y];
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < e.n; i += gridDim.x * 256ll) {
    const float gi = e.g[i];
    const float mi = fmaf(beta1, e.m[i], (1.f - beta1) * gi);
    const float vi = fmaf(beta2, e.v[i], (1.f - beta2) * gi * gi);
    e.m[i] = mi;
    e.v[i] = vi;
    e.p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}
int adam_step_multi(const void* table, int count, float lr, float beta1, float beta2, float eps, int step,
                    cudaStream_t stream) {
  if (step < 1) return set_error("adam_step_multi: step counts from 1");
  if (count < 1 || count > 65535) return set_error("adam_step_multi: 1..65535 tensors");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step), bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  adam_step_multi_kernel<<<dim3(64, count), 256, 0, stream>>>(static_cast<const AdamEntry*>(table), beta1, beta2,
                                                              static_cast<float>(lr / bc1),
                                                              static_cast<float>(1.0 / sqrt(bc2)), eps);
  NC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nc

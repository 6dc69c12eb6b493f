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
// All three directions are implicit GEMMs with fp32 semantics on the TENSOR cores (mma.sync TF32, "3xTF32").
//   * a CTA owns a 64 x 64 output tile; a warp group of 256 threads (8 warps, 2 x 4) covers it with 32 x 16 warp
//     patches = 2 x 2 mma.sync.m16n8k8 tiles per warp;
//   * the K dimension is walked in chunks of 16 staged in shared memory ([k][m] / [k][n], row pitch 72 words, columns
//     XOR-swizzled by ((k / 4) % 4) * 8: the fragment loads `[k0 + lane % 4][m0 + lane / 4]` AND the producers'
//     stores hit 32 distinct banks);
//   * operands are fp32 in memory; every element is split ONCE, when its chunk is written, into two TF32 numbers
//     hi = tf32(x), lo = tf32(x - hi), and a product is evaluated as lo*hi + hi*lo + hi*hi with fp32 accumulation
//     (~2^-21 relative per product; the fp32 fixtures of tests/test_gpu_discriminator.py hold with unchanged
//     tolerances);
//   * the images are tiny (M = 121..2916 pixels) while K reaches 8192, so a layer has only 2..46 output tiles: the K
//     range of a tile is SPLIT ACROSS A THREAD-BLOCK CLUSTER (up to 8 CTAs).  Partial tiles are summed by the
//     cluster's rank 0 through distributed shared memory in rank order -> bitwise repeatable, no scratch buffer.
//     (NC_DISC_GROUPS = 2 additionally splits a CTA's range over two warp groups with independent pipelines — own
//     stages, own named barrier, summed group 0 + group 1.  Measured: no gain, the SM's legacy-MMA rate is the bound;
//     off by default.)
//   * a pipeline keeps PF chunks in flight in registers (the loads are volatile asm: ptxas otherwise sinks them to
//     the end of the loop body, next to their use) and double buffers its stage (one barrier per chunk).
// History (tests/cuda/probe_disc.cu, profiles/r02_probe_disc_final.log): the round-1 kernels were
// 4 x 4 register-tiled FFMA loops whose time was proportional to the chunks per CTA at ~1 us per chunk, warm or cold
// cache — 32 LDS.128 per 256 FFMA per thread saturate the shared-memory pipe.  The mma form is NOT faster on a B200:
// HMMA.1688.F32.TF32 sustains one MMA per ~5.6 clk per SM, 192 MMAs per chunk = 0.58 us whatever the occupancy.
constexpr int TB = 64;  // tile edge (M and N)
constexpr int KB = 16;  // K chunk
constexpr int PF = 4;   // chunks in flight (register prefetch depth)
constexpr int LDS_PITCH = TB + 8;
#ifndef NC_DISC_GROUPS
#define NC_DISC_GROUPS 1
#endif
constexpr int GROUPS = NC_DISC_GROUPS;  // independent K pipelines (warp groups) per CTA: 1 or 2
constexpr int GT = 256;     // threads per pipeline
constexpr int CONV_THREADS = GROUPS * GT;

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  const float r = v - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

// One K chunk of both operands as TF32 (hi, lo) pairs.
struct Stage {
  uint32_t ah[KB][LDS_PITCH], al[KB][LDS_PITCH];
  uint32_t bh[KB][LDS_PITCH], bl[KB][LDS_PITCH];
  static __device__ __forceinline__ int swz(int k, int c) { return c ^ (((k >> 2) & 3) << 3); }
  __device__ __forceinline__ void put_a(int k, int m, float v) {
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    ah[k][swz(k, m)] = hi, al[k][swz(k, m)] = lo;
  }
  __device__ __forceinline__ void put_b(int k, int n, float v) {
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    bh[k][swz(k, n)] = hi, bl[k][swz(k, n)] = lo;
  }
};
constexpr int CONV_SMEM_BYTES = GROUPS * 2 * static_cast<int>(sizeof(Stage));
static_assert(CONV_SMEM_BYTES >= GT * 16 * static_cast<int>(sizeof(float)), "the reduction slots alias the stages");

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Predicated read-only loads that stay where they are written (volatile): zero when !ok, the address is not touched.
__device__ __forceinline__ float ldg_or_zero(const float* p, bool ok) {
  float v;
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
      : "=f"(v)
      : "l"(p), "r"(static_cast<int>(ok)));
  return v;
}
__device__ __forceinline__ float4 ldg4_or_zero(const float* p, bool ok) {
  float4 v;
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
      "mov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p), "r"(static_cast<int>(ok)));
  return v;
}

// Position of thread t (0..255) of a warp group inside the 64 x 64 tile: warp (wm, wn) owns rows [32 wm, +32) x
// columns [16 wn, +16); accumulator acc[mi][ni][r] is element (row, col) = (32 wm + 16 mi + gid + 8 (r >> 1),
// 16 wn + 8 ni + 2 tig + (r & 1)) (the m16n8 C-fragment layout: gid = lane / 4, tig = lane % 4).
struct TilePos {
  int wm, wn, gid, tig;
  __device__ __forceinline__ explicit TilePos(int t) {
    const int warp = t >> 5, lane = t & 31;
    wm = warp & 1, wn = warp >> 1, gid = lane >> 2, tig = lane & 3;
  }
  __device__ __forceinline__ int row(int mi, int r) const { return 32 * wm + 16 * mi + gid + 8 * (r >> 1); }
  __device__ __forceinline__ int col(int ni, int r) const { return 16 * wn + 8 * ni + 2 * tig + (r & 1); }
};

using Acc = float[2][2][4];

__device__ __forceinline__ void tile_mma(const Stage& s, const TilePos& tp, Acc& acc) {
  // The chunk is accumulated from zero inside the tensor core (6 MMAs per accumulator) and then added to the running
  // sum with a rounded fp32 add: the MMA's own accumulation is not a rounded fp32 add, and weight gradients are small
  // differences of large sums (InstanceNorm makes sum_p dy[p] ~ 0), so chains of 50-1500 MMAs showed errors of 1e-3..
  // 3e-2 against the fp32 reference (tools/debug_disc_grad.py) where 6-MMA chains stay at 1e-5.
  Acc part = {};
#pragma unroll
  for (int k0 = 0; k0 < KB; k0 += 8) {
    uint32_t ah[2][4], al[2][4], bh[2][2], bl[2][2];
    const int ka = k0 + tp.tig, kb = k0 + tp.tig + 4;
    const int sa = Stage::swz(ka, 0), sb = Stage::swz(kb, 0);  // warp-uniform (k0 is a multiple of 8, tig < 4)
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int m = 32 * tp.wm + 16 * mi + tp.gid;
      ah[mi][0] = s.ah[ka][m ^ sa], ah[mi][1] = s.ah[ka][(m + 8) ^ sa];
      ah[mi][2] = s.ah[kb][m ^ sb], ah[mi][3] = s.ah[kb][(m + 8) ^ sb];
      al[mi][0] = s.al[ka][m ^ sa], al[mi][1] = s.al[ka][(m + 8) ^ sa];
      al[mi][2] = s.al[kb][m ^ sb], al[mi][3] = s.al[kb][(m + 8) ^ sb];
    }
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) {
      const int n = 16 * tp.wn + 8 * ni + tp.gid;
      bh[ni][0] = s.bh[ka][n ^ sa], bh[ni][1] = s.bh[kb][n ^ sb];
      bl[ni][0] = s.bl[ka][n ^ sa], bl[ni][1] = s.bl[kb][n ^ sb];
    }
    // small terms first; the four accumulators are interleaved so that consecutive MMAs are independent
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) mma_tf32(part[mi][ni], al[mi], bh[ni]);
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) mma_tf32(part[mi][ni], ah[mi], bl[ni]);
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) mma_tf32(part[mi][ni], ah[mi], bh[ni]);
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[mi][ni][r] += part[mi][ni][r];
}

struct Regs {
  float a[4], b[4];
};

__device__ __forceinline__ void group_barrier(int group) {
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(GT) : "memory");
}

// K loop of ONE warp group over chunks [k_begin, k_end): fetch(k, regs) loads this thread's share of chunk k from
// global memory, store(stage, regs) writes it into a shared-memory stage.  PF chunks are in flight in registers; the
// group's two stages alternate.  Chunk i + 1 is written to its stage AFTER the barrier of chunk i, behind the MMAs
// of chunk i, so that its conversion / store instructions run while the tensor pipe drains (with the store in front
// of the barrier all warps ran store, fragment-load and MMA phases in lock step: 0.66 us per chunk).  The barrier of
// chunk i orders both "stage i is complete" and "everyone has finished reading stage i - 1".
template <class Fetch, class Store>
__device__ __forceinline__ void k_loop(int k_begin, int k_end, Stage* st, int group, const TilePos& tp, Acc& acc,
                                       Fetch fetch, Store store) {
  Regs r[PF];
#pragma unroll
  for (int u = 0; u < PF; ++u)
    if (k_begin + u < k_end) fetch(k_begin + u, r[u]);
#ifdef NC_DISC_STORE_FIRST
  int buf = 0;
  for (int k = k_begin; k < k_end; k += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      if (k + u < k_end) {  // group-uniform
        store(st[buf], r[u]);
        group_barrier(group);
        if (k + u + PF < k_end) fetch(k + u + PF, r[u]);
        tile_mma(st[buf], tp, acc);
        buf ^= 1;
      }
    }
  }
#else
  if (k_begin < k_end) store(st[0], r[0]);
  int buf = 0;
  for (int k = k_begin; k < k_end; k += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      if (k + u < k_end) {  // group-uniform
        group_barrier(group);
        if (k + u + PF < k_end) fetch(k + u + PF, r[u]);                  // r[u] (chunk k + u) is in its stage
        tile_mma(st[buf], tp, acc);
        if (k + u + 1 < k_end) store(st[buf ^ 1], r[(u + 1) % PF]);       // chunk k + u + 1, behind the MMAs
        buf ^= 1;
      }
    }
  }
#endif
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

// acc(group 0) += acc(group 1) through shared memory.  Called by every thread of the CTA after both K loops; `slots`
// (GT x 16 floats) aliases the stages, hence the leading barrier.
__device__ __forceinline__ void combine_groups(Acc& acc, float* slots, int group, int t) {
  float* mine = slots + t * 16;
  __syncthreads();
  if (GROUPS == 1) return;
  if (group == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) mine[i] = acc[i >> 3][(i >> 2) & 1][i & 3];
  }
  __syncthreads();
  if (group == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i >> 3][(i >> 2) & 1][i & 3] += mine[i];
  }
}

// Sum the 16 accumulators per thread of group 0 of all CTAs of the cluster into rank 0's registers (ranks added in
// order 1, 2, ...).  Every thread of every CTA of the cluster must call it (the cluster barrier is CTA-wide).
__device__ __forceinline__ void cluster_reduce(Acc& acc, float* slots, unsigned rank, unsigned size, int group, int t) {
  if (size == 1) return;
  float* mine = slots + t * 16;
  if (group == 0 && rank != 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) mine[i] = acc[i >> 3][(i >> 2) & 1][i & 3];
  }
  cluster_sync_all();
  if (group == 0 && rank == 0) {
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
        float(&a)[4] = acc[q >> 1][q & 1];
        a[0] += v.x, a[1] += v.y, a[2] += v.z, a[3] += v.w;
      }
    }
  }
  cluster_sync_all();  // remote shared memory must stay alive until rank 0 has read it
}

// This CTA's share [begin, end) of `total` chunks (cluster rank) and, inside it, the calling warp group's half.
__device__ __forceinline__ void k_range(int total, unsigned rank, unsigned csize, int group, int& kb, int& ke) {
  const int begin = static_cast<int>(static_cast<long long>(total) * rank / csize);
  const int end = static_cast<int>(static_cast<long long>(total) * (rank + 1) / csize);
  const int mid = GROUPS == 1 ? end : begin + (end - begin + 1) / 2;
  kb = group == 0 ? begin : mid;
  ke = group == 0 ? mid : end;
}

// y[n,co,ho,wo] = b[co] + sum_{ci,kh,kw} w[co,ci,kh,kw] * x[n,ci,ho*s-1+kh,wo*s-1+kw]   (+ optional LeakyReLU)
// GEMM: M = output pixel, N = co, K = (ci, tap); one K chunk = the 16 taps of one input channel.
// grid.x = M tiles x cluster size (the cluster splits the input channels).
__global__ void __launch_bounds__(CONV_THREADS)
conv2d_k4_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int Cin,
                     int H, int W, int Cout, int Ho, int Wo, int stride, float slope, float* __restrict__ y) {
  extern __shared__ __align__(16) uint8_t conv_smem[];
  Stage* stages = reinterpret_cast<Stage*>(conv_smem);  // [GROUPS][2]
  float* slots = reinterpret_cast<float*>(conv_smem);   // reused after the K loops
  const unsigned rank = cluster_rank_x(), csize = cluster_size_x();
  const int group = threadIdx.x / GT, t = threadIdx.x % GT;
  const TilePos tp(t);
  const int p0 = (blockIdx.x / csize) * TB, co0 = blockIdx.y * TB, n = blockIdx.z;
  const int P = Ho * Wo;
  // loader roles: A: pixel t % 64, filter row kh = t / 64 (4 kw each); B: channel t / 4, taps (t % 4) * 4 .. + 3
  const int ap = p0 + (t & 63), akh = t >> 6;
  const int aho = ap / Wo, awo = ap - aho * Wo;
  const int ah = aho * stride - 1 + akh, aw0 = awo * stride - 1;
  const bool a_row_ok = ap < P && ah >= 0 && ah < H;
  const int bco = co0 + (t >> 2), bt = (t & 3) * 4;
  const float* xn = x + static_cast<size_t>(n) * Cin * H * W;
  int kb, ke;
  k_range(Cin, rank, csize, group, kb, ke);
  Acc acc = {};
  k_loop(
      kb, ke, stages + 2 * group, group, tp, acc,
      [&](int ci, Regs& r) {
        const float* xr = xn + (static_cast<size_t>(ci) * H + ah) * W;
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          const int ww = aw0 + kw;
          r.a[kw] = ldg_or_zero(xr + ww, a_row_ok && ww >= 0 && ww < W);
        }
        const float4 v = ldg4_or_zero(w + (static_cast<size_t>(bco) * Cin + ci) * 16 + bt, bco < Cout);
        r.b[0] = v.x, r.b[1] = v.y, r.b[2] = v.z, r.b[3] = v.w;
      },
      [&](Stage& s, const Regs& r) {
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          s.put_a(akh * 4 + kw, t & 63, r.a[kw]);
          s.put_b(bt + kw, t >> 2, r.b[kw]);
        }
      });
  combine_groups(acc, slots, group, t);
  cluster_reduce(acc, slots, rank, csize, group, t);
  if (rank != 0 || group != 0) return;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int p = p0 + tp.row(mi, r), co = co0 + tp.col(ni, r);
        if (p >= P || co >= Cout) continue;
        float v = acc[mi][ni][r] + (b ? __ldg(b + co) : 0.f);
        if (slope != 1.f) v = v > 0.f ? v : v * slope;
        y[(static_cast<size_t>(n) * Cout + co) * P + p] = v;
      }
}

// Cout = 1 (the PatchGAN's last layer, 512 -> 1 on a 12 x 12 map: 121 dot products of length 8192).  As a 64 x 64
// tile problem it is 2 tiles with 63 of 64 columns idle (60-76 us per pass).  Here: one block per output pixel, the
// 256 threads stride over the input channels (16 taps each), fixed-order shuffle + shared-memory tree -> deterministic.
__global__ void __launch_bounds__(256)
conv2d_k4_fwd_cout1_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                           int Cin, int H, int W, int Ho, int Wo, int stride, float slope, float* __restrict__ y) {
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = Ho * Wo;
  const int p = blockIdx.x, n = blockIdx.y;
  const int ho = p / Wo, wo = p - ho * Wo;
  const int h0 = ho * stride - 1, w0 = wo * stride - 1;
  const float* xn = x + static_cast<size_t>(n) * Cin * H * W;
  float s = 0.f;
  for (int ci = threadIdx.x; ci < Cin; ci += 256) {
    const float4* wr = reinterpret_cast<const float4*>(w + static_cast<size_t>(ci) * 16);
    const float* xc = xn + static_cast<size_t>(ci) * H * W;
    float xv[16];
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const int hh = h0 + kh;
      const bool h_ok = hh >= 0 && hh < H;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const int ww = w0 + kw;
        xv[kh * 4 + kw] = (h_ok && ww >= 0 && ww < W) ? __ldg(xc + hh * W + ww) : 0.f;
      }
    }
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const float4 wv = __ldg(wr + kh);
      s = fmaf(wv.x, xv[kh * 4], s);
      s = fmaf(wv.y, xv[kh * 4 + 1], s);
      s = fmaf(wv.z, xv[kh * 4 + 2], s);
      s = fmaf(wv.w, xv[kh * 4 + 3], s);
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    v += b ? __ldg(b) : 0.f;
    if (slope != 1.f) v = v > 0.f ? v : v * slope;
    y[static_cast<size_t>(n) * P + p] = v;
  }
}

// dx[n,ci,h,w] = sum_{co,kh,kw : (h+1-kh) = ho*s, (w+1-kw) = wo*s} w[co,ci,kh,kw] * dy[n,co,ho,wo]
// GEMM: M = input pixel, N = ci, K = (co, tap); one K chunk = the 16 taps of one output channel (split over the
// cluster and the two warp groups).
__global__ void __launch_bounds__(CONV_THREADS)
conv2d_k4_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, int Cin, int H, int W, int Cout,
                       int Ho, int Wo, int stride, float* __restrict__ dx) {
  extern __shared__ __align__(16) uint8_t conv_smem[];
  Stage* stages = reinterpret_cast<Stage*>(conv_smem);
  float* slots = reinterpret_cast<float*>(conv_smem);
  const unsigned rank = cluster_rank_x(), csize = cluster_size_x();
  const int group = threadIdx.x / GT, t = threadIdx.x % GT;
  const TilePos tp(t);
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
  int kb, ke;
  k_range(Cout, rank, csize, group, kb, ke);
  Acc acc = {};
  k_loop(
      kb, ke, stages + 2 * group, group, tp, acc,
      [&](int co, Regs& r) {
        const float* dr = dyn + (static_cast<size_t>(co) * Ho + ho) * Wo;
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) r.a[kw] = ldg_or_zero(dr + wos[kw], wos[kw] >= 0);
        const float4 v = ldg4_or_zero(w + (static_cast<size_t>(co) * Cin + bci) * 16 + bt, bci < Cin);
        r.b[0] = v.x, r.b[1] = v.y, r.b[2] = v.z, r.b[3] = v.w;
      },
      [&](Stage& s, const Regs& r) {
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          s.put_a(akh * 4 + kw, t & 63, r.a[kw]);
          s.put_b(bt + kw, t >> 2, r.b[kw]);
        }
      });
  combine_groups(acc, slots, group, t);
  cluster_reduce(acc, slots, rank, csize, group, t);
  if (rank != 0 || group != 0) return;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int p = p0 + tp.row(mi, r), ci = ci0 + tp.col(ni, r);
        if (p < P && ci < Cin) dx[(static_cast<size_t>(n) * Cin + ci) * P + p] = acc[mi][ni][r];
      }
}

// Stride 2: an input pixel (h, w) is reached only through the taps kh = (h + 1) % 2 (+ 2), kw = (w + 1) % 2 (+ 2) —
// 4 of 16 — so the generic kernel above multiplies 75 % zeros.  Here the image is split into its four parity classes
// (h % 2, w % 2); per class: M = the class's pixels, K = (co, the 4 valid taps) = Cout / 4 chunks instead of Cout.
// Same loader roles, stage layout and reduction as above; grid.x = 4 classes x M tiles x cluster size.
__global__ void __launch_bounds__(CONV_THREADS)
conv2d_k4_dgrad_s2_kernel(const float* __restrict__ dy, const float* __restrict__ w, int Cin, int H, int W, int Cout,
                          int Ho, int Wo, int mt_class, float* __restrict__ dx) {
  extern __shared__ __align__(16) uint8_t conv_smem[];
  Stage* stages = reinterpret_cast<Stage*>(conv_smem);
  float* slots = reinterpret_cast<float*>(conv_smem);
  const unsigned rank = cluster_rank_x(), csize = cluster_size_x();
  const int group = threadIdx.x / GT, t = threadIdx.x % GT;
  const TilePos tp(t);
  const int tile = blockIdx.x / csize;
  const int cls = tile / mt_class, p0 = (tile - cls * mt_class) * TB;
  const int ph = cls >> 1, pw = cls & 1;                  // parity of (h, w)
  const int kh0 = (ph + 1) & 1, kw0 = (pw + 1) & 1;       // first valid tap; the other one is + 2
  const int Hq = (H - ph + 1) / 2, Wq = (W - pw + 1) / 2;  // pixels of this class
  const int Pq = Hq * Wq;
  const int ci0 = blockIdx.y * TB, n = blockIdx.z;
  // loader roles: A: class pixel t % 64, output channel (t / 64) of the chunk, its 4 valid taps;
  //               B: input channel t / 4, output channel (t % 4) of the chunk, the same 4 taps
  const int ap = p0 + (t & 63), aco = t >> 6;
  const int ai = ap / Wq, aj = ap - ai * Wq;
  const int ah = 2 * ai + ph, aw = 2 * aj + pw;
  int offA[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int th = ah + 1 - (kh0 + 2 * (j >> 1)), tw = aw + 1 - (kw0 + 2 * (j & 1));  // even by construction
    const bool ok = ap < Pq && th >= 0 && (th >> 1) < Ho && tw >= 0 && (tw >> 1) < Wo;
    offA[j] = ok ? (th >> 1) * Wo + (tw >> 1) : -1;
  }
  const int bci = ci0 + (t >> 2), bco = t & 3;
  int offB[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) offB[j] = (kh0 + 2 * (j >> 1)) * 4 + kw0 + 2 * (j & 1);
  const float* dyn = dy + static_cast<size_t>(n) * Cout * Ho * Wo;
  int kb, ke;
  k_range((Cout + 3) / 4, rank, csize, group, kb, ke);
  Acc acc = {};
  k_loop(
      kb, ke, stages + 2 * group, group, tp, acc,
      [&](int c, Regs& r) {
        const int coa = 4 * c + aco, cob = 4 * c + bco;
        const float* dr = dyn + static_cast<size_t>(coa) * Ho * Wo;
        const float* wr = w + (static_cast<size_t>(cob) * Cin + bci) * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          r.a[j] = ldg_or_zero(dr + offA[j], offA[j] >= 0 && coa < Cout);
          r.b[j] = ldg_or_zero(wr + offB[j], bci < Cin && cob < Cout);
        }
      },
      [&](Stage& s, const Regs& r) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s.put_a(aco * 4 + j, t & 63, r.a[j]);
          s.put_b(bco * 4 + j, t >> 2, r.b[j]);
        }
      });
  combine_groups(acc, slots, group, t);
  cluster_reduce(acc, slots, rank, csize, group, t);
  if (rank != 0 || group != 0) return;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int r = 0; r < 4; r += 2) {  // rows of this thread: r = 0, 1 share a row, r = 2, 3 the row + 8
      const int p = p0 + tp.row(mi, r);
      if (p >= Pq) continue;
      const int i = p / Wq, j = p - i * Wq;
      const size_t pix = static_cast<size_t>(2 * i + ph) * W + 2 * j + pw;
#pragma unroll
      for (int ni = 0; ni < 2; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ci = ci0 + tp.col(ni, r + e);
          if (ci < Cin) dx[(static_cast<size_t>(n) * Cin + ci) * H * W + pix] = acc[mi][ni][r + e];
        }
    }
}

// Cin = 1 (gradient of the PatchGAN's first layer w.r.t. the image, needed by the generator losses): as a tile
// problem 63 of 64 columns idle (57-166 us on 108 x 108 images).  Here: a block owns 64 image pixels, its four
// quarters split the output channels (ascending inside a quarter) and are added in quarter order -> deterministic;
// the Cout x 16 filter sits in shared memory.
__global__ void __launch_bounds__(256)
conv2d_k4_dgrad_cin1_kernel(const float* __restrict__ dy, const float* __restrict__ w, int H, int W, int Cout, int Ho,
                            int Wo, int stride, float* __restrict__ dx) {
  extern __shared__ float wsm[];  // [Cout][16], then [4][64] partial sums
  float* part = wsm + Cout * 16;
  for (int i = threadIdx.x; i < Cout * 16; i += 256) wsm[i] = __ldg(w + i);
  __syncthreads();
  const int P = H * W;
  const int px = threadIdx.x & 63, quarter = threadIdx.x >> 6;
  const int p = blockIdx.x * 64 + px, n = blockIdx.y;
  float s = 0.f;
  if (p < P) {
    const int h = p / W, wq = p - h * W;
    int off[16];  // offset of the dy element behind tap (kh, kw), -1 when the tap does not reach an output pixel
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const int th = h + 1 - kh;
      const bool h_ok = th >= 0 && th % stride == 0 && th / stride < Ho;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const int tw = wq + 1 - kw;
        const bool ok = h_ok && tw >= 0 && tw % stride == 0 && tw / stride < Wo;
        off[kh * 4 + kw] = ok ? (th / stride) * Wo + tw / stride : -1;
      }
    }
    const int c_begin = Cout * quarter / 4, c_end = Cout * (quarter + 1) / 4;
    const float* dyn = dy + static_cast<size_t>(n) * Cout * Ho * Wo;
#pragma unroll 2
    for (int co = c_begin; co < c_end; ++co) {
      const float* dc = dyn + static_cast<size_t>(co) * Ho * Wo;
      const float* wc = wsm + co * 16;
      float g[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) g[k] = off[k] >= 0 ? __ldg(dc + off[k]) : 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) s = fmaf(wc[k], g[k], s);
    }
  }
  part[quarter * 64 + px] = s;
  __syncthreads();
  if (quarter == 0 && p < P)
    dx[static_cast<size_t>(n) * P + p] = ((part[px] + part[64 + px]) + part[128 + px]) + part[192 + px];
}

// dw[co,ci,kh,kw] = sum_{n,ho,wo} dy[n,co,ho,wo] * x[n,ci,ho*s-1+kh,wo*s-1+kw]
// GEMM: M = co, N = (ci, tap) (a 64-wide tile = 4 input channels x 16 taps), K = (n, output pixel) in chunks of 16
// (the pixel chunks of an image are split over the cluster; the (image, chunk) sequence of a CTA over its two warp
// groups).  grid.x = M tiles x cluster size.
__global__ void __launch_bounds__(CONV_THREADS)
conv2d_k4_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int N, int Cin, int H, int W,
                       int Cout, int Ho, int Wo, int stride, float* __restrict__ dw) {
  extern __shared__ __align__(16) uint8_t conv_smem[];
  Stage* stages = reinterpret_cast<Stage*>(conv_smem);
  float* slots = reinterpret_cast<float*>(conv_smem);
  const unsigned rank = cluster_rank_x(), csize = cluster_size_x();
  const int group = threadIdx.x / GT, t = threadIdx.x % GT;
  const TilePos tp(t);
  const int co0 = (blockIdx.x / csize) * TB, j0 = blockIdx.y * TB;  // j = ci * 16 + tap
  const int P = Ho * Wo;
  // loader roles: A: channel t / 4, pixels (t % 4) * 4 .. + 3 of the chunk; B: pixel t % 16, columns (t / 16) * 4 .. + 3
  const int aco = co0 + (t >> 2), apk = (t & 3) * 4;
  const int bpk = t & 15, bj = (t >> 4) * 4;
  const int chunks = (P + KB - 1) / KB;
  const int k_begin = static_cast<int>(static_cast<long long>(chunks) * rank / csize);
  const int k_end = static_cast<int>(static_cast<long long>(chunks) * (rank + 1) / csize);
  const int per_image = k_end - k_begin;
  // this CTA's sequence of (image, pixel chunk) pairs — images ascending, chunks ascending — halved over the groups
  const int total = per_image * N, mid = GROUPS == 1 ? total : (total + 1) / 2;
  const int kb = group == 0 ? 0 : mid, ke = group == 0 ? mid : total;
  Acc acc = {};
  k_loop(
      kb, ke, stages + 2 * group, group, tp, acc,
      [&](int kk, Regs& r) {
        const int n = kk / per_image, kc = k_begin + (kk - n * per_image);
        const float* dyc = dy + (static_cast<size_t>(n) * Cout + aco) * P;
        const float* xn = x + static_cast<size_t>(n) * Cin * H * W;
        const int pc = kc * KB;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int p = pc + apk + i;
          r.a[i] = ldg_or_zero(dyc + p, aco < Cout && p < P);
        }
        const int p = pc + bpk;
        const int ho = p / Wo, wo = p - ho * Wo;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = j0 + bj + i;
          const int ci = j >> 4, tap = j & 15;
          const int hh = ho * stride - 1 + (tap >> 2), ww = wo * stride - 1 + (tap & 3);
          const bool ok = p < P && ci < Cin && hh >= 0 && hh < H && ww >= 0 && ww < W;
          r.b[i] = ldg_or_zero(xn + (static_cast<size_t>(ci) * H + hh) * W + ww, ok);
        }
      },
      [&](Stage& s, const Regs& r) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s.put_a(apk + i, t >> 2, r.a[i]);
          s.put_b(bpk, bj + i, r.b[i]);
        }
      });
  combine_groups(acc, slots, group, t);
  cluster_reduce(acc, slots, rank, csize, group, t);
  if (rank != 0 || group != 0) return;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int co = co0 + tp.row(mi, r), j = j0 + tp.col(ni, r);
        if (co < Cout && j < Cin * 16) dw[static_cast<size_t>(co) * Cin * 16 + j] = acc[mi][ni][r];
      }
}

// Cin = 1 (weight gradient of the PatchGAN's first layer, 64 x 16 sums over all N * Ho * Wo output pixels): as a tile
// problem one 64 x 16 tile whose K = 2916 N pixels are split over at most 8 CTAs (68-128 us).  Here: one block per
// output channel, the threads stride over (image, pixel) with 16 + 1 accumulators (the bias gradient comes for free),
// fixed-order tree over the block -> deterministic.  x (46 KB per image) stays in L1 / L2.
__global__ void __launch_bounds__(256)
conv2d_k4_wgrad_cin1_kernel(const float* __restrict__ x, const float* __restrict__ dy, int N, int H, int W, int Cout,
                            int Ho, int Wo, int stride, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float red[17][256];
  const int co = blockIdx.x, P = Ho * Wo;
  float s[17];
#pragma unroll
  for (int i = 0; i < 17; ++i) s[i] = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* dyc = dy + (static_cast<size_t>(n) * Cout + co) * P;
    const float* xn = x + static_cast<size_t>(n) * H * W;
    for (int p = threadIdx.x; p < P; p += 256) {
      const float g = __ldg(dyc + p);
      const int ho = p / Wo, wo = p - ho * Wo;
      const int h0 = ho * stride - 1, w0 = wo * stride - 1;
      s[16] += g;
#pragma unroll
      for (int kh = 0; kh < 4; ++kh) {
        const int hh = h0 + kh;
        const bool h_ok = hh >= 0 && hh < H;
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          const int ww = w0 + kw;
          const float xv = (h_ok && ww >= 0 && ww < W) ? __ldg(xn + hh * W + ww) : 0.f;
          s[kh * 4 + kw] = fmaf(g, xv, s[kh * 4 + kw]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 17; ++i) red[i][threadIdx.x] = s[i];
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int i = 0; i < 17; ++i) red[i][threadIdx.x] += red[i][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 16) dw[co * 16 + threadIdx.x] = red[threadIdx.x][0];
  if (threadIdx.x == 16 && db) db[co] = red[16][0];
}

// Cluster size (1, 2, 4 or 8) for a launch of `tiles` output tiles whose K loop has `k_chunks` iterations, from a
// cost model fitted to tests/cuda/probe_disc.cu on a B200: a chunk costs ~0.6 us of an SM (two co-resident CTAs take
// turns rather than overlap), a cluster costs ~2.5 us per rank (launch, two cluster barriers, the serial reduction).
static int g_cluster_override = 0;  // debug hook (tests/cuda/probe_disc.cu): 1, 2, 4, 8 force the cluster size
void debug_set_disc_cluster(int c) { g_cluster_override = c; }
static int pick_cluster(long long tiles, int k_chunks) {
  if (g_cluster_override > 0) return g_cluster_override;
  int best = 1;
  double best_us = 1e30;
  for (int c = 1; c <= 8; c *= 2) {
    if (c > 1 && k_chunks / c < 4) break;
    const long long per_sm = (tiles * c + num_sms() - 1) / num_sms();
    const double us = static_cast<double>(per_sm) * ((k_chunks + c - 1) / c) * 0.6 + 2.5 * c;
    if (us < best_us) best_us = us, best = c;
  }
  return best;
}

template <class Kern, class... Args>
static int launch_clustered_threads(Kern kern, dim3 grid, int threads, int smem_bytes, int cluster,
                                    cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem_bytes;
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

template <class Kern, class... Args>
static int launch_conv(Kern kern, bool* attr_flags, dim3 grid, int cluster, cudaStream_t stream, Args... args) {
  if (first_use_on_device(attr_flags))
    NC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM_BYTES));
  return launch_clustered_threads(kern, grid, CONV_THREADS, CONV_SMEM_BYTES, cluster, stream, args...);
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
  if (Cout == 1) {
    conv2d_k4_fwd_cout1_kernel<<<dim3(Ho * Wo, N), 256, 0, stream>>>(x, w, b, Cin, H, W, Ho, Wo, stride, slope, y);
    NC_CUDA(cudaGetLastError());
    return 0;
  }
  const int mt = (Ho * Wo + TB - 1) / TB, nt = (Cout + TB - 1) / TB;
  const int cl = pick_cluster(static_cast<long long>(mt) * nt * N, Cin);
  static bool attr[64] = {false};
  return launch_conv(conv2d_k4_fwd_kernel, attr, dim3(mt * cl, nt, N), cl, stream, x, w, b, Cin, H, W, Cout, Ho, Wo,
                     stride, slope, y);
}
int conv2d_k4_dgrad(const float* dy, const float* w, int N, int Cin, int H, int W, int Cout, int stride, float* dx,
                    cudaStream_t stream) {
  if (stride != 1 && stride != 2) return set_error("conv2d_k4: stride must be 1 or 2");
  const int Ho = out_extent(H, stride), Wo = out_extent(W, stride);
  if (Cin > 65535 || N > 65535) return set_error("conv2d_k4: too many channels / images");
  if (Cin == 1 && (Cout * 16 + 256) * sizeof(float) <= 48 * 1024) {
    conv2d_k4_dgrad_cin1_kernel<<<dim3((H * W + 63) / 64, N), 256, (Cout * 16 + 256) * sizeof(float), stream>>>(
        dy, w, H, W, Cout, Ho, Wo, stride, dx);
    NC_CUDA(cudaGetLastError());
    return 0;
  }
  const int nt = (Cin + TB - 1) / TB;
  if (stride == 2) {
    const int mtq = (((H + 1) / 2) * ((W + 1) / 2) + TB - 1) / TB;  // M tiles of the largest parity class
    const int cl = pick_cluster(4ll * mtq * nt * N, (Cout + 3) / 4);
    static bool attr2[64] = {false};
    return launch_conv(conv2d_k4_dgrad_s2_kernel, attr2, dim3(4 * mtq * cl, nt, N), cl, stream, dy, w, Cin, H, W, Cout,
                       Ho, Wo, mtq, dx);
  }
  const int mt = (H * W + TB - 1) / TB;
  const int cl = pick_cluster(static_cast<long long>(mt) * nt * N, Cout);
  static bool attr[64] = {false};
  return launch_conv(conv2d_k4_dgrad_kernel, attr, dim3(mt * cl, nt, N), cl, stream, dy, w, Cin, H, W, Cout, Ho, Wo,
                     stride, dx);
}
int conv2d_k4_wgrad(const float* x, const float* dy, int N, int Cin, int H, int W, int Cout, int stride, float* dw,
                    float* db, cudaStream_t stream) {
  if (stride != 1 && stride != 2) return set_error("conv2d_k4: stride must be 1 or 2");
  const int Ho = out_extent(H, stride), Wo = out_extent(W, stride);
  if (Cin == 1) {
    conv2d_k4_wgrad_cin1_kernel<<<Cout, 256, 0, stream>>>(x, dy, N, H, W, Cout, Ho, Wo, stride, dw, db);
    NC_CUDA(cudaGetLastError());
    return 0;
  }
  const int mt = (Cout + TB - 1) / TB, nt = (Cin * 16 + TB - 1) / TB;
  const int cl = pick_cluster(static_cast<long long>(mt) * nt, (Ho * Wo + KB - 1) / KB);
  static bool attr[64] = {false};
  if (int rc = launch_conv(conv2d_k4_wgrad_kernel, attr, dim3(mt * cl, nt), cl, stream, x, dy, N, Cin, H, W, Cout, Ho,
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
// One thread-block cluster (1 CTA of 256 threads for the small prediction maps, 8 CTAs of 1024 threads for volumes):
// every thread strides over the cluster's elements, block sums are tree-reduced in shared memory, and rank 0 adds the
// CTAs' sums through distributed shared memory in rank order -> bitwise repeatable, no scratch buffer, one launch.
// (The L1 cycle term of a 108^3 crop sits in the critical path of every iteration: a single 256-thread block took
// ~0.4 ms, 8 x 256 threads 0.1 ms; 8 x 1024 threads leave ~40 float4 pairs per thread.)
__global__ void __launch_bounds__(1024)
loss_fwd_kernel(const float* __restrict__ p, const float* __restrict__ q, float target, long long n, int mode,
                float* __restrict__ loss) {
  __shared__ double red[1024];
  __shared__ double block_sum;
  const unsigned rank = cluster_rank_x(), csize = cluster_size_x();
  const unsigned nt = blockDim.x;
  double s = 0.0;
  auto term = [&](float pv, float qv) {
    const float d = pv - (mode == 0 ? target : qv);
    return mode == 0 ? static_cast<double>(d) * d : fabs(static_cast<double>(d));
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(q)) & 15) == 0;
  const long long n4 = vec ? (n >> 2) : 0;
  const float4* p4 = reinterpret_cast<const float4*>(p);
  const float4* q4 = reinterpret_cast<const float4*>(q);
  const long long stride = static_cast<long long>(nt) * csize;
#pragma unroll 4
  for (long long i = static_cast<long long>(rank) * nt + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(p4 + i);
    const float4 b = mode == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(q4 + i);
    s += (term(a.x, b.x) + term(a.y, b.y)) + (term(a.z, b.z) + term(a.w, b.w));
  }
  for (long long i = (n4 << 2) + static_cast<long long>(rank) * nt + threadIdx.x; i < n; i += stride)
    s += term(p[i], mode == 0 ? 0.f : q[i]);
  red[threadIdx.x] = s;
  __syncthreads();
  for (unsigned o = nt >> 1; o >= 1; o >>= 1) {
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
  const bool big = n >= (1 << 16);
  return launch_clustered_threads(loss_fwd_kernel, dim3(big ? 8 : 1), big ? 1024 : 256, 0, big ? 8 : 1, stream, p, q,
                                  target, n, mode, loss);
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

// First layer of Unet_deconv — Conv3d(1 -> 64, k3 s1 p1), reference models/networks.py:420,490 — on tcgen05.
//
// With one input channel the GEMM K is only 27, so there is nothing for TMA to stage: the A operand is built in the
// kernel (im2col).  To keep fp32 fidelity of the 16-bit input on fp16 tensor cores every input value x is split
// into hi = fp16(x) and lo = fp16(x - hi) (22 significant bits together); K index = 2*tap + {0: hi, 1: lo}, i.e.
// 27 (hi,lo) words + 5 zero words = 64 halves = ONE 128-byte swizzled row per voxel, and the weight matrix simply
// holds every tap twice.  The split is done ONCE per halo element (when the halo is staged in shared memory as
// packed half2 words), so building a voxel's K-row is 27 word loads and 8 vector stores.
//
//   warps 4-7  producers: per 8(w) x 16(h) x 1(d) tile load the 10 x 18 x 3 fp32 halo into shared memory, then each
//              thread gathers its voxel's 27 neighbours, splits, and writes its 128-byte K-row (SWIZZLE_128B image);
//              fence.proxy.async + mbarrier hand the stage to the tensor core
//   warp  4    additionally: TMEM allocation, weight load (8 KB, resident for the whole kernel) and, once a stage is
//              complete, the MMA issue: 4 x (128x64x16)
//   warps 0-3  epilogue: TMEM -> registers; per-channel sum / sum-of-squares kept in registers across ALL tiles of a
//              cube (one butterfly reduction per cube per CTA instead of per tile); fp16 tile staged in shared memory
//              and written with one TMA store (volume overhang clipped by the TMA unit)
// Persistent grid; A stages (4) and TMEM accumulators (8 x 64 columns) are rings, so producer, tensor core and
// epilogue overlap across tiles.  Statistics partial rows: one per (cube, CTA) -> deterministic.
#include <cuda_fp16.h>

#include "internal.h"
#include "ptx.cuh"

namespace nc {

namespace c1 {
constexpr int TW = 8, TH = 16;
constexpr int HW = TW + 2, HH = TH + 2;          // halo extents
constexpr int HALO_FLOATS = 3 * HH * HW;         // 540
constexpr int HALO_PITCH = 576;                  // floats per stage (2304 B)
constexpr int NA = 4;                            // A-stage ring
constexpr int NACC = 8;                          // accumulator ring (8 x 64 TMEM columns)
constexpr int A_BYTES = 128 * 128;               // 128 voxels x 64 halves
constexpr int B_BYTES = 64 * 128;
constexpr int OUT_BYTES = 128 * 128;
constexpr int THREADS = 256;
constexpr int SMEM_BYTES = NA * A_BYTES + B_BYTES + 2 * OUT_BYTES + NA * HALO_PITCH * 4 + 4 * 2 * 64 * 4 + 1024;
}  // namespace c1

struct Conv1Args {
  const float* x;          // [NB][D][H][W]
  const uint8_t* wpacked;  // 64 rows x 128 B swizzled image
  float* stats_partial;    // [NB][gridDim.x][2][64]
  int NB, D, H, W, tiles_w, tiles_h, tiles_per_cube;
};

// Column sums over the 32 lanes of a warp (recursive halving, 31 shuffles): lane l ends with sum_lanes v[l].
__device__ __forceinline__ float colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float keep = upper ? v[i + off] : v[i];
      const float send = upper ? v[i] : v[i + off];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  __half2 t = __halves2half2(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(c1::THREADS, 1)
conv_cin1_tc_kernel(const __grid_constant__ CUtensorMap tmapOut, const Conv1Args args) {
  using namespace c1;
  // __align__(1024) instead of rounding the pointer up by hand: an integer round trip makes the compiler forget the
  // address space and emit generic LD/ST for every shared-memory access
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smA = smem;
  uint8_t* smB = smA + NA * A_BYTES;
  uint8_t* smOut = smB + B_BYTES;
  float* halo = reinterpret_cast<float*>(smOut + 2 * OUT_BYTES);
  float* statScratch = halo + NA * HALO_PITCH;  // [4 warps][2][64]
  uint64_t* aFull = reinterpret_cast<uint64_t*>(statScratch + 4 * 2 * 64);
  uint64_t* aEmpty = aFull + NA;
  uint64_t* accFull = aEmpty + NA;
  uint64_t* accEmpty = accFull + NACC;
  uint64_t* bFull = accEmpty + NACC;
  uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(bFull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmapOut);
    for (int i = 0; i < NA; ++i) {
      ptx::mbar_init(&aEmpty[i], 1);
    }
    for (int i = 0; i < NACC; ++i) {
      ptx::mbar_init(&accFull[i], 1);
      ptx::mbar_init(&accEmpty[i], 4);
    }
    ptx::mbar_init(bFull, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc<512>(tmemPtr);
    ptx::tmem_relinquish();
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bFull, B_BYTES);
      ptx::bulk_load(smB, args.wpacked, B_BYTES, bFull);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmemPtr;

  if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ im2col producers
    const int pt = threadIdx.x - 128;  // 0..127 = GEMM row = voxel of the tile
    const int mw = pt & 7, mh = pt >> 3;
    constexpr uint32_t idesc = ptx::make_idesc_f16(128, 64);
    constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_lo = ((ptx::smem_u32(smB) >> 4) & 0x3FFF) | (1u << 16);
    const uint32_t smA_u32 = ptx::smem_u32(smA);
    constexpr int NPRE = (HALO_FLOATS + 127) / 128;  // halo elements per thread (5)
    // halo element e of this thread: index i = pt + 128 e -> (dd, hh, ww), fixed for the whole kernel
    int off_d[NPRE], off_h[NPRE], off_w[NPRE];
#pragma unroll
    for (int e = 0; e < NPRE; ++e) {
      const int i = pt + 128 * e;
      off_w[e] = i % HW - 1;
      off_h[e] = (i / HW) % HH - 1;
      off_d[e] = i < HALO_FLOATS ? i / (HW * HH) - 1 : 1 << 20;  // out of range -> always zero
    }
    // software pipeline: the NEXT tile's halo is fetched into registers while the current tile is converted
    auto fetch = [&](int nb, int tile, float (&pre)[NPRE]) {
      const int wt = tile % args.tiles_w;
      const int r = tile / args.tiles_w;
      const int ht = r % args.tiles_h;
      const int d = r / args.tiles_h;
      const float* xin = args.x + static_cast<size_t>(nb) * args.D * args.H * args.W;
#pragma unroll
      for (int e = 0; e < NPRE; ++e) {
        const int gz = d + off_d[e], gy = ht * TH + off_h[e], gx = wt * TW + off_w[e];
        const bool in = gz >= 0 && gz < args.D && gy >= 0 && gy < args.H && gx >= 0 && gx < args.W;
        pre[e] = in ? __ldg(xin + (static_cast<size_t>(gz) * args.H + gy) * args.W + gx) : 0.f;
      }
    };
    // software pipeline: the halos of the NEXT THREE tiles are in flight in registers while the current tile is
    // converted.  (Measured: no faster than one tile of look-ahead — ncu's source view shows the four epilogue warps
    // spinning on accFull 23 % of all samples while these producer warps retire ~325 mostly dependent integer /
    // shared-memory instructions per tile at 0.34 IPC per scheduler: the kernel is bound by the producers' instruction
    // chain, not by the input latency.  The deeper prefetch is kept because it costs nothing.)
    constexpr int PF = 3;
    float pre[PF][NPRE];
    int fnb = 0, ftile = blockIdx.x;                       // fetch cursor
    bool fhave = ftile < args.tiles_per_cube && args.NB > 0;
    auto fetch_next = [&](float (&dst)[NPRE]) {
      if (!fhave) return;
      fetch(fnb, ftile, dst);
      ftile += gridDim.x;
      if (ftile >= args.tiles_per_cube) {
        ftile = blockIdx.x;
        ++fnb;
      }
      fhave = fnb < args.NB;
    };
    const int per_cube = fhave ? (args.tiles_per_cube - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                     static_cast<int>(gridDim.x)
                               : 0;
    const int total = per_cube * args.NB;
#pragma unroll
    for (int u = 0; u < PF; ++u) fetch_next(pre[u]);
    auto step = [&](int it, float (&cur)[NPRE]) {
      const int st = it & (NA - 1);
      ptx::mbar_wait(&aEmpty[st], ((it / NA) & 1) ^ 1);
      uint32_t* hl = reinterpret_cast<uint32_t*>(halo) + st * HALO_PITCH;
#pragma unroll
      for (int e = 0; e < NPRE; ++e)
        if (pt + 128 * e < HALO_FLOATS) {
          const __half h = __float2half_rn(cur[e]);
          hl[pt + 128 * e] = pack_h2(h, __float2half_rn(cur[e] - __half2float(h)));   // (hi, lo)
        }
      fetch_next(cur);                                     // the registers are free again: start tile it + PF
      ptx::named_bar_sync(3, 128);
      uint32_t kw32[32];  // the voxel's K-row: word t = (hi, lo) of tap t, words 27..31 zero
#pragma unroll
      for (int kd = 0; kd < 3; ++kd)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) kw32[(kd * 3 + kh) * 3 + kw] = hl[(kd * HH + mh + kh) * HW + mw + kw];
#pragma unroll
      for (int t = 27; t < 32; ++t) kw32[t] = 0u;
      uint8_t* row = smA + st * A_BYTES + pt * 128;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        *reinterpret_cast<uint4*>(row + ((u ^ (pt & 7)) << 4)) =
            make_uint4(kw32[4 * u], kw32[4 * u + 1], kw32[4 * u + 2], kw32[4 * u + 3]);
      ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
      ptx::named_bar_sync(3, 128);  // the whole stage is written and fenced
      if (warp == 4) {
        // producer warp 4 doubles as the MMA issuer (8 warps keep the register cap at 255: the epilogue holds
        // 128 running statistics per thread)
        const int slot = it & (NACC - 1);
        if (it == 0) ptx::mbar_wait(bFull, 0);
        ptx::mbar_wait(&accEmpty[slot], ((it / NACC) & 1) ^ 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t a_lo = (((smA_u32 + st * A_BYTES) >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_f16(tmem_base + slot * 64, (static_cast<uint64_t>(DESC_HI) << 32) | (a_lo + 2 * k),
                          (static_cast<uint64_t>(DESC_HI) << 32) | (b_lo + 2 * k), idesc, k == 0 ? 0u : 1u);
          ptx::umma_commit(&aEmpty[st]);
          ptx::umma_commit(&accFull[slot]);
        }
        __syncwarp();
      }
    };
    for (int it = 0; it < total; it += PF) {
#pragma unroll
      for (int u = 0; u < PF; ++u)
        if (it + u < total) step(it + u, pre[u]);
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp;  // TMEM lane quarter
    const int m = q * 32 + lane;
    const int mw = m & 7, mh = m >> 3;
    const bool leader = threadIdx.x == 0;
    int it = 0;
    for (int nb = 0; nb < args.NB; ++nb) {
      float csum[2][32], csq[2][32];  // per-lane (= per-row) running column sums for this cube
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) csum[c][i] = csq[c][i] = 0.f;
      for (int tile = blockIdx.x; tile < args.tiles_per_cube; tile += gridDim.x, ++it) {
        const int wt = tile % args.tiles_w;
        const int r = tile / args.tiles_w;
        const int ht = r % args.tiles_h;
        const int d = r / args.tiles_h;
        const int w0 = wt * TW, h0 = ht * TH;
        const bool valid = (w0 + mw < args.W) && (h0 + mh < args.H);
        const int slot = it & (NACC - 1), sb = it & 1;
        ptx::mbar_wait(&accFull[slot], (it / NACC) & 1);
        ptx::tc_fence_after();
        if (leader) ptx::bulk_wait_group_read<1>();
        ptx::named_bar_sync(2, 128);
        uint8_t* stg = smOut + sb * OUT_BYTES + m * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t raw[32];
          ptx::tmem_ld32(tmem_base + slot * 64 + c * 32 + (static_cast<uint32_t>(q * 32) << 16), raw);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = __uint_as_float(raw[8 * i + 2 * e]), b = __uint_as_float(raw[8 * i + 2 * e + 1]);
              __half2 t = __floats2half2_rn(a, b);  // |sum_27 w*x|, x in [0,1]: far inside the fp16 range
              pk[e] = *reinterpret_cast<uint32_t*>(&t);
            }
            *reinterpret_cast<uint4*>(stg + (((c * 4 + i) ^ (m & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          if (valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float x = __uint_as_float(raw[i]);
              csum[c][i] += x;
              csq[c][i] = fmaf(x, x, csq[c][i]);
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&accEmpty[slot]);
        ptx::fence_proxy_async();
        ptx::named_bar_sync(2, 128);
        if (leader) {
          ptx::tma_store_5d(&tmapOut, smOut + sb * OUT_BYTES, 0, w0, h0, d, nb);
          ptx::bulk_commit_group();
        }
      }
      // one deterministic reduction per cube: lanes -> warp (butterfly), 4 warps -> CTA (fixed order)
      float* mine = statScratch + q * 2 * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        mine[c * 32 + lane] = colsum32(csum[c], lane);
        mine[64 + c * 32 + lane] = colsum32(csq[c], lane);
      }
      ptx::named_bar_sync(1, 128);
      {
        const int i = threadIdx.x;  // 0..127 -> (which, column)
        const float s = (statScratch[i] + statScratch[128 + i]) + (statScratch[256 + i] + statScratch[384 + i]);
        args.stats_partial[((static_cast<size_t>(nb) * gridDim.x + blockIdx.x) * 2) * 64 + i] = s;
      }
      ptx::named_bar_sync(1, 128);
    }
    if (leader) ptx::bulk_wait_group_read<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// weights: OIDHW fp32 (64,1,3,3,3) -> 64 rows x 64 halves, k = 2*tap + {0,1}: w[co][tap] (k >= 54: 0); swizzled
__global__ void pack_conv1_weights_kernel(const float* __restrict__ w, __half* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * 64) return;
  const int r = idx >> 6, k = idx & 63;
  const float v = k < 54 ? w[r * 27 + (k >> 1)] : 0.f;
  out[r * 64 + ((((k >> 3) ^ (r & 7)) << 3) | (k & 7))] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
}

int pack_conv1_weights(const float* w, void* packed, cudaStream_t stream) {
  pack_conv1_weights_kernel<<<16, 256, 0, stream>>>(w, static_cast<__half*>(packed));
  NC_CUDA(cudaGetLastError());
  return 0;
}

static int conv1_grid(int tiles_per_cube) { return tiles_per_cube < num_sms() ? tiles_per_cube : num_sms(); }

size_t conv_cin1_stats_tiles(int NB, int D, int H, int W) {
  const int tiles = D * ((H + c1::TH - 1) / c1::TH) * ((W + c1::TW - 1) / c1::TW);
  return static_cast<size_t>(NB) * conv1_grid(tiles);
}

int conv3d_cin1_k3_fwd(const float* x, const void* wpacked, int NB, int D, int H, int W, int Cout, void* y_raw,
                       float* stats_partial, cudaStream_t stream) {
  if (Cout != 64) return set_error("conv3d_cin1_k3_fwd: Cout must be 64");
  if (reinterpret_cast<uintptr_t>(y_raw) & 15) return set_error("conv3d_cin1_k3_fwd: y_raw must be 16-byte aligned");
  auto encode = get_tensor_map_encoder();
  if (!encode) return set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  CUtensorMap tm;
  cuuint64_t dims[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)NB};
  cuuint64_t strides[4] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128, (cuuint64_t)D * H * W * 128};
  cuuint32_t box[5] = {64, c1::TW, c1::TH, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, y_raw, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled (conv1 output) failed (%d)", (int)r);
  Conv1Args a{};
  a.x = x, a.wpacked = static_cast<const uint8_t*>(wpacked), a.stats_partial = stats_partial;
  a.NB = NB, a.D = D, a.H = H, a.W = W;
  a.tiles_w = (W + c1::TW - 1) / c1::TW, a.tiles_h = (H + c1::TH - 1) / c1::TH;
  a.tiles_per_cube = D * a.tiles_h * a.tiles_w;
  static bool attr[64] = {false};
  if (first_use_on_device(attr))
    NC_CUDA(cudaFuncSetAttribute(conv_cin1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c1::SMEM_BYTES));
  conv_cin1_tc_kernel<<<conv1_grid(a.tiles_per_cube), c1::THREADS, c1::SMEM_BYTES, stream>>>(tm, a);
  NC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nc

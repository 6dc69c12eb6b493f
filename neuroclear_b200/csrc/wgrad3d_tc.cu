// Weight gradient of Conv3d (k in {1,3,5}, stride 1, pad k/2) on tcgen05 tensor cores:
//     dW[co][ci][kd][kh][kw] = sum over voxels v of dy[v][co] * x[v + (kd,kh,kw) - pad][ci]
// (what torch autograd computes for the Conv3d layers of Unet_deconv / DeepLinearGenerator, reference
// models/networks.py:413-538, 893-917, when train_onecube.py calls backward).
//
// As a GEMM the reduction dimension is the VOXEL, which is the row index of both NDHWC operands, so both are read
// "MN-major": a 16-voxel K step is 16 consecutive 128-byte rows (two 8-row swizzle groups) and the 64 channels of a
// row are the M / N extent.  The same halo plane the forward kernel stages (one 5-D TMA box, zero-filled outside the
// volume) serves every (kh,kw) tap of one kd as a shifted window:
//   A (M = 128) = two taps of x stacked along M: the second 64-channel atom starts `LBO` bytes after the first,
//                 and LBO is simply the row shift between the two taps;
//   B (N = 64)  = the dy tile (8 x 16 voxels x 64 output channels);
//   D[tap-pair][ci][co] accumulates in TMEM over ALL voxel tiles of the CTA (split-K across CTAs).
// A CTA owns (64-channel chunk of Cin, 64-channel block of Cout, kd, group of <= 16 (kh,kw) taps, split s) and
// streams (x halo plane d+kd-pad, dy tile d) pairs through a TMA ring.  Partial sums go to
// scratch[s][tap][co][ci] and are reduced over s in a fixed order (deterministic) into the OIDHW fp32 gradient.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "internal.h"
#include "ptx.cuh"

namespace nc {

namespace {

constexpr int TW = 8;
constexpr int TH = 16;
constexpr int MAX_PAIRS = 8;  // 8 x 64 fp32 columns = all 512 TMEM columns

// KS = in-plane filter extent, KSD = extent along d (see ConvCfg in conv3d_tc.cu)
template <int KS, int KSD = KS>
struct WgCfg {
  static constexpr int PAD = KS / 2;
  static constexpr int PAD_D = KSD / 2;
  static constexpr int HALO_W = TW + KS - 1;
  static constexpr int HALO_H = TH + KS - 1;
  static constexpr int PLANE_ROWS = HALO_W * HALO_H;
  static constexpr int PLANE_BOX_BYTES = PLANE_ROWS * 128;
  // + one spare row: the dummy second half of an odd tap pair reads one row past the last tap's window
  static constexpr int PLANE_BYTES = (PLANE_BOX_BYTES + 128 + 1023) / 1024 * 1024;
  static constexpr int DY_BYTES = TW * TH * 128;
  static constexpr int STAGE_BYTES = PLANE_BYTES + DY_BYTES;
  static constexpr int NSTAGE_FIT = (232448 - 1024) / STAGE_BYTES;
  static constexpr int NSTAGE = NSTAGE_FIT > 6 ? 6 : NSTAGE_FIT;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024;
  static constexpr int TAPS2 = KS * KS;
  static constexpr int TAPS_PER_GROUP = TAPS2 <= 2 * MAX_PAIRS ? TAPS2 : (TAPS2 + 1) / 2;  // 9 | 13
  static constexpr int GROUPS = (TAPS2 + TAPS_PER_GROUP - 1) / TAPS_PER_GROUP;
  static_assert(TAPS_PER_GROUP <= 2 * MAX_PAIRS, "tap group exceeds TMEM");
  static_assert(NSTAGE >= 3, "ring too shallow");
};

struct WgArgs {
  int W, H, D, NB;
  int Cin, Cout;
  int tiles_w, tiles_h;
  long long plane_tiles;  // NB * tiles_h * tiles_w * D
  int splits;
  uint32_t idesc;
  float* partial;  // [splits][KS^3][Cout][Cin]
};

template <int KS, int KSD>
__global__ void __launch_bounds__(256, 1)
wgrad3d_tc_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapDy,
                  const WgArgs args) {
  using C = WgCfg<KS, KSD>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* aux = smem + C::NSTAGE * C::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty = full + C::NSTAGE;
  uint64_t* accFull = empty + C::NSTAGE;
  uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(accFull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- work item of this CTA
  int item = blockIdx.x / args.splits;
  const int split = blockIdx.x - item * args.splits;
  const int grp = item % C::GROUPS;
  item /= C::GROUPS;
  const int kd = item % KSD;
  item /= KSD;
  const int cout_blocks = args.Cout / 64;
  const int cb = item % cout_blocks;
  const int cc = item / cout_blocks;
  const int t0 = grp * C::TAPS_PER_GROUP;
  const int nt = (C::TAPS2 - t0) < C::TAPS_PER_GROUP ? (C::TAPS2 - t0) : C::TAPS_PER_GROUP;
  const int pairs = (nt + 1) >> 1;
  const long long pt_begin = args.plane_tiles * split / args.splits;
  const long long pt_end = args.plane_tiles * (split + 1) / args.splits;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmapX);
    ptx::prefetch_tmap(&tmapDy);
    for (int i = 0; i < C::NSTAGE; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(accFull, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<512>(tmemPtr);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmemPtr;

  if (warp == 0) {
    // ------------------------------------------------------------ producer: (x halo plane, dy tile) per stage
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (long long pt = pt_begin; pt < pt_end; ++pt) {
        const int d = static_cast<int>(pt % args.D);
        long long r = pt / args.D;
        const int wt = static_cast<int>(r % args.tiles_w);
        r /= args.tiles_w;
        const int ht = static_cast<int>(r % args.tiles_h);
        const int nb = static_cast<int>(r / args.tiles_h);
        uint8_t* stage = smem + st * C::STAGE_BYTES;
        ptx::mbar_wait(&empty[st], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&full[st], C::PLANE_BOX_BYTES + C::DY_BYTES);
        ptx::tma_load_5d(stage, &tmapX, &full[st], cc * 64, wt * TW - C::PAD, ht * TH - C::PAD, d + kd - C::PAD_D,
                         nb);
        ptx::tma_load_5d(stage + C::PLANE_BYTES, &tmapDy, &full[st], cb * 64, wt * TW, ht * TH, d, nb);
        if (++st == C::NSTAGE) {
          st = 0;
          ph ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform control flow)
    // descriptor high words: SBO = distance between the two 8-row K groups, version 1, SWIZZLE_128B
    constexpr uint32_t X_HI = ((C::HALO_W * 128u) >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t DY_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
    uint32_t x_off[MAX_PAIRS];  // (start offset >> 4) | (LBO >> 4) << 16 of each tap pair
#pragma unroll
    for (int p = 0; p < MAX_PAIRS; ++p) {
      const int ta = t0 + 2 * p, tb = (2 * p + 1 < nt) ? ta + 1 : -1;
      const int sa = (ta / KS) * C::HALO_W + (ta % KS);
      const int sb = tb >= 0 ? (tb / KS) * C::HALO_W + (tb % KS) : sa + 1;  // odd count: dummy half, ignored later
      x_off[p] = static_cast<uint32_t>(sa * 8) | (static_cast<uint32_t>((sb - sa) * 8) << 16);
    }
    const uint32_t sm_u32 = ptx::smem_u32(smem);
    int st = 0;
    uint32_t ph = 0;
    bool first = true;
    for (long long pt = pt_begin; pt < pt_end; ++pt) {
      ptx::mbar_wait(&full[st], ph);
      ptx::tc_fence_after();
      const uint32_t xs = ((sm_u32 + st * C::STAGE_BYTES) >> 4);
      const uint32_t dys = ((sm_u32 + st * C::STAGE_BYTES + C::PLANE_BYTES) >> 4);
      if (ptx::elect_one()) {
#pragma unroll 1
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t bdesc = (static_cast<uint64_t>(DY_HI) << 32) | ((dys + ks * 128) & 0x3FFF) | (1u << 16);
          const uint32_t xk = xs + ks * 2 * C::HALO_W * 8;
#pragma unroll
          for (int p = 0; p < MAX_PAIRS; ++p) {
            if (p < pairs) {
              const uint32_t lo = ((xk + (x_off[p] & 0xFFFF)) & 0x3FFF) | (x_off[p] & 0xFFFF0000u);
              ptx::umma_f16(tmem_base + p * 64, (static_cast<uint64_t>(X_HI) << 32) | lo, bdesc, args.idesc,
                            (first && ks == 0) ? 0u : 1u);
            }
          }
        }
        ptx::umma_commit(&empty[st]);
      }
      __syncwarp();
      first = false;
      if (++st == C::NSTAGE) {
        st = 0;
        ph ^= 1;
      }
    }
    if (ptx::elect_one()) ptx::umma_commit(accFull);
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: TMEM -> scratch[split][tap][co][ci]
    const int q = warp & 3;
    const int m = q * 32 + lane;  // accumulator lane: (tap of the pair, input channel)
    ptx::mbar_wait(accFull, 0);
    ptx::tc_fence_after();
    const bool any = pt_end > pt_begin;  // an empty split never issued an MMA: its accumulators are undefined
    const size_t per_split = static_cast<size_t>(KSD * KS * KS) * args.Cout * args.Cin;
    float* base = args.partial + per_split * split;
    for (int p = 0; p < pairs; ++p) {
      const int t = 2 * p + (m >> 6);
      const bool live = t < nt;
      const int tap = kd * C::TAPS2 + t0 + t;
      float* dst = base + (static_cast<size_t>(tap) * args.Cout + cb * 64) * args.Cin + cc * 64 + (m & 63);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t raw[32];
        ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + p * 64 + half * 32, raw);
        ptx::tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            dst[static_cast<size_t>(half * 32 + i) * args.Cin] = any ? __uint_as_float(raw[i]) : 0.f;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// out[co][ci][tap] (OIDHW fp32) = sum_s partial[s][tap][co][ci], s ascending
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int taps, int Cout, int Cin,
                                    float* __restrict__ out) {
  const size_t n = static_cast<size_t>(taps) * Cout * Cin;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // eight splits' loads in flight, added in split order (a rolled loop was a chain of up to 49 L2 round trips)
    float s = 0.f;
    int k = 0;
    for (; k + 8 <= splits; k += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(partial + (k + u) * n + i);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; k < splits; ++k) s += __ldg(partial + k * n + i);
    const int ci = static_cast<int>(i % Cin);
    size_t r = i / Cin;
    const int co = static_cast<int>(r % Cout);
    const int tap = static_cast<int>(r / Cout);
    out[(static_cast<size_t>(co) * Cin + ci) * taps + tap] = s;
  }
}

int make_tmap(CUtensorMap* m, const void* base, int fmt, int C, int W, int H, int D, int NB, int boxW, int boxH) {
  auto encode = get_tensor_map_encoder();
  if (!encode) return set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)NB};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)boxW, (cuuint32_t)boxH, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(m, fmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5,
                      const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

template <int KS, int KSD = KS>
int wgrad_items(int Cin, int Cout) {
  return (Cin / 64) * (Cout / 64) * KSD * WgCfg<KS, KSD>::GROUPS;
}

// ks = 71 selects the (ksd = 7, ksp = 1) filter of the im2col'ed k7 Cin = 1 layer
int items_for(int ks, int Cin, int Cout) {
  return ks == 1    ? wgrad_items<1>(Cin, Cout)
         : ks == 3  ? wgrad_items<3>(Cin, Cout)
         : ks == 5  ? wgrad_items<5>(Cin, Cout)
                    : wgrad_items<1, 7>(Cin, Cout);
}
int taps_for(int ks) { return ks == 71 ? 7 : ks * ks * ks; }

template <int KS, int KSD = KS>
int launch(const void* x, int x_fmt, const void* dy, int dy_fmt, WgArgs a, cudaStream_t stream) {
  using C = WgCfg<KS, KSD>;
  CUtensorMap tx, tdy;
  if (int rc = make_tmap(&tx, x, x_fmt, a.Cin, a.W, a.H, a.D, a.NB, C::HALO_W, C::HALO_H)) return rc;
  if (int rc = make_tmap(&tdy, dy, dy_fmt, a.Cout, a.W, a.H, a.D, a.NB, TW, TH)) return rc;
  auto kern = wgrad3d_tc_kernel<KS, KSD>;
  static bool attr_set[64] = {false};
  if (first_use_on_device(attr_set))
    NC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  kern<<<wgrad_items<KS, KSD>(a.Cin, a.Cout) * a.splits, 256, C::SMEM_BYTES, stream>>>(tx, tdy, a);
  NC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int conv3d_wgrad_splits(int ks, int Cin, int Cout, long long plane_tiles) {
  const int items = items_for(ks, Cin, Cout);
  int s = num_sms() / items;
  if (s < 1) s = 1;
  if (s > plane_tiles) s = static_cast<int>(plane_tiles > 0 ? plane_tiles : 1);
  return s;
}

static long long wgrad_plane_tiles(int NB, int D, int H, int W) {
  return static_cast<long long>(NB) * ((H + TH - 1) / TH) * ((W + TW - 1) / TW) * D;
}

size_t conv3d_wgrad_scratch_bytes(int ks, int NB, int D, int H, int W, int Cin, int Cout) {
  const int s = conv3d_wgrad_splits(ks, Cin, Cout, wgrad_plane_tiles(NB, D, H, W));
  return static_cast<size_t>(s) * taps_for(ks) * Cin * Cout * sizeof(float);
}

int conv3d_wgrad(const void* x, int x_fmt, const void* dy, int dy_fmt, int NB, int D, int H, int W, int Cin,
                 int Cout, int ks, void* scratch, float* dw, cudaStream_t stream) {
  if (ks != 1 && ks != 3 && ks != 5 && ks != 71)
    return set_error("conv3d_wgrad: kernel size must be 1, 3, 5 or 71 (= 7 x 1 x 1)");
  if (Cin % 64 || Cout % 64) return set_error("conv3d_wgrad: Cin and Cout must be multiples of 64");
  if ((x_fmt | dy_fmt) & ~1) return set_error("conv3d_wgrad: operand format must be 0 (fp16) or 1 (bf16)");
  WgArgs a{};
  a.W = W, a.H = H, a.D = D, a.NB = NB, a.Cin = Cin, a.Cout = Cout;
  a.tiles_w = (W + TW - 1) / TW;
  a.tiles_h = (H + TH - 1) / TH;
  a.plane_tiles = wgrad_plane_tiles(NB, D, H, W);
  a.splits = conv3d_wgrad_splits(ks, Cin, Cout, a.plane_tiles);
  a.partial = static_cast<float*>(scratch);
  // fp32 accumulate, A = x and B = dy both MN-major (bits 15, 16), M = 128, N = 64
  a.idesc = ptx::make_idesc_f16(128, 64) | (static_cast<uint32_t>(x_fmt) << 7) |
            (static_cast<uint32_t>(dy_fmt) << 10) | (1u << 15) | (1u << 16);
  if (int rc = (ks == 1   ? launch<1>(x, x_fmt, dy, dy_fmt, a, stream)
                : ks == 3 ? launch<3>(x, x_fmt, dy, dy_fmt, a, stream)
                : ks == 5 ? launch<5>(x, x_fmt, dy, dy_fmt, a, stream)
                          : launch<1, 7>(x, x_fmt, dy, dy_fmt, a, stream)))
    return rc;
  const int taps = taps_for(ks);
  wgrad_reduce_kernel<<<num_sms() * 2, 256, 0, stream>>>(a.partial, a.splits, taps, Cout, Cin, dw);
  NC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nc

// Implicit-GEMM Conv3d (k3 s1 p1) and ConvTranspose3d (k2 s2) for sm_100a on tcgen05 tensor cores.
//
// Replaces the torch.nn.Conv3d / ConvTranspose3d calls of Unet_deconv (reference models/networks.py:478-538,
// layers U2..U12 of SURVEY.md §2c). Design (not a translation of any library kernel):
//
//   * activations are fp16 NDHWC (fp16, not bf16: see DESIGN.md "Operand precision"); one GEMM row = one voxel,
//     K = taps x Cin walked as (64-channel chunk, tap)
//   * a CTA owns an output tile of 8(w) x 16(h) x TD(d) voxels = TD accumulators of 128 rows in TMEM
//   * the input is staged ONCE per (tile, chunk) as TD+KS-1 halo planes of (8+KS-1) x (16+KS-1) voxels, each
//     voxel one 128-byte row, written by a 5-D TMA box with SWIZZLE_128B (out-of-volume rows arrive as zeros
//     = the conv's zero padding). Every filter tap is then just a shifted window into those planes: the
//     UMMA shared-memory descriptor's start address moves by (kh*(8+KS-1)+kw) rows and the plane index by kd,
//     with the 8-row-group stride set to one halo line. Input traffic drops from 27x to ~2x of the tile.
//   * weights are pre-packed as the exact swizzled shared-memory image and streamed with 1-D bulk copies through
//     a ring; one weight stage feeds the MMAs of all TD output planes.
//   * warp roles: 0 = halo-plane TMA producer, 1 = MMA issuer (warp-uniform loops, one elected lane issues),
//     2 = TMEM allocator + weight producer, 4..7 = epilogue, 8..11 (XF only) = in-place InstanceNorm + ReLU of the
//     landed planes. Two accumulator sets ping-pong so the epilogue of tile i overlaps the MMAs of tile i+1.
//     The grid is persistent (<= #SMs CTAs, static tile stride).
//   * STACK (Cout = 64): the three kd taps of a (kh,kw) are stacked along N (N = 64/128/192 MMAs), see ConvCfg.
//   * training reuses this kernel: the data gradient of a stride-1 "same" conv is the conv of dy with the
//     channel-transposed, tap-reversed filter — template flag BF16 (bf16 operands / output, no statistics; gradients
//     need bf16's exponent range), see conv3d_k3_dgrad.  KSD != KS gives filters that are longer along d than in the
//     plane (the k7 Cin = 1 layer of DeepLinearGenerator as a 7-tap depth conv over an in-plane im2col), KS = 5 the
//     k5 layer, KS = 1 in MODE 0 the data gradient of the transposed convs (conv3d_tc_64, conv3d_k1_bf16).
//   * epilogue MODE 0: raw fp16 NDHWC store + per-tile per-channel (sum, sum of squares) partials taken from
//     the fp32 accumulators (InstanceNorm statistics, reduced deterministically later);
//     MODE 1: transposed conv — + bias, fp16 tile staged in shared memory and written by a TMA store whose tensor
//     map walks the fine grid with element stride 2 (pixel shuffle) into a channel slice of the concat buffer.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "internal.h"
#include "ptx.cuh"

namespace nc {

constexpr int TW = 8;   // tile extent in w  (rows inside one 8-row swizzle group)
constexpr int TH = 16;  // tile extent in h  (number of 8-row groups of a 128-row MMA)

constexpr int pow2_at_least(int v) { return v <= 32 ? 32 : v <= 64 ? 64 : v <= 128 ? 128 : v <= 256 ? 256 : 512; }

// STACK (Cout = 64 only): the three kd taps of one (kh,kw) are stacked along the MMA N dimension — one read of a
// shifted A window feeds the accumulators of up to three output planes (N = 64/128/192) instead of one, which takes
// the kernel off the shared-memory-bandwidth roof an N = 64 MMA sits on (4 KB of A per 32 math cycles).  A tile is
// TD = 4 output planes processed in two phases of three input planes; each phase walks all nine (kh,kw) weight
// stages (192 rows = [kd2 | kd1 | kd0]), so the 6-slot plane ring always prefetches the next phase's planes.
// KS = in-plane (h, w) filter extent, KSD = extent along d (= KS for the cubic filters of Unet_deconv; the k7 Cin = 1
// layer of DeepLinearGenerator runs as KS = 1, KSD = 7 over a 49-channel in-plane im2col of its input).
template <int KS, int BN, int TD, bool STACK = false, bool OUT_STAGE = false, int KSD = KS>
struct ConvCfg {
  static constexpr int PAD = KS / 2;
  static constexpr int PAD_D = KSD / 2;
  static constexpr int HALO_W = TW + KS - 1;
  static constexpr int HALO_H = TH + KS - 1;
  static constexpr int PLANE_ROWS = HALO_W * HALO_H;
  static constexpr int PLANE_BOX_BYTES = PLANE_ROWS * 128;
  static constexpr int PLANE_BYTES = (PLANE_BOX_BYTES + 1023) / 1024 * 1024;
  static constexpr int PPC = TD + KSD - 1;  // halo planes per (tile, chunk)
  // plane ring depth: two planes (STACK: one phase) of look-ahead; the 30 KB planes of k5 leave room for none —
  // there the ring refills as the kd groups retire their planes
  static constexpr int NSLOT = (STACK || KS >= 5) ? PPC : PPC + 2;
  static constexpr int TAPS = KSD * KS * KS;
  static constexpr int BSTAGE_BYTES = (STACK ? 3 * BN : BN) * 128;
  static constexpr int STAGES_PER_CHUNK = STACK ? 2 * KS * KS : TAPS;
  static constexpr int AUX_BYTES = 1024 + 4 * BN * 2 * 4;  // barriers + per-warp stats scratch
  static constexpr int SMEM_LIMIT = 232448;
  static constexpr int OUT_BYTES = OUT_STAGE ? 2 * 128 * 128 : 0;  // two 128-row x 64-channel fp16 store tiles
  static constexpr int NBST_FIT = (SMEM_LIMIT - 1024 - AUX_BYTES - OUT_BYTES - NSLOT * PLANE_BYTES) / BSTAGE_BYTES;
  static constexpr int NBST = NBST_FIT > 8 ? 8 : NBST_FIT;
  static constexpr int ACC_COLS = TD * BN;
  static constexpr int TMEM_COLS = pow2_at_least(2 * ACC_COLS);
  static constexpr int SMEM_BYTES = 1024 + NSLOT * PLANE_BYTES + NBST * BSTAGE_BYTES + OUT_BYTES + AUX_BYTES;
  static_assert(2 * ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert(NBST >= 2, "weight ring too shallow");
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "bad BN");
  static_assert(!STACK || (KS == 3 && KSD == 3 && BN == 64 && TD == 4),
                "STACK is built for k3, Cout 64, four output planes");
};

struct ConvTcArgs {
  int W, H, D, NB;
  int max_ctas;  // persistent grid size cap (debug hook: forces several tiles per CTA on small problems)
  // XF: the input is a RAW conv output; InstanceNorm + ReLU of the producer layer are applied on the fly
  const float* in_mr;  // [NB][2][cin_total]: mean, rstd of the input's channels
  int cin_total;
  int chunks;  // Cin / 64
  int n_tiles;  // GEMM N / BN
  int tiles_w, tiles_h, tiles_d;
  int total_tiles;
  int tiles_h_cap;   // > 0: the regular kernel covers only this many 16-line h-tiles (the rest is the RP kernel's)
  int rp_h0, rp_rem;  // remainder-pair kernel: first line and number of lines (1..6) of the remainder strip
  // stats_partial rows are grouped per cube (in_stats_finalize reduces rows [nb * R, (nb + 1) * R)): with the
  // remainder-pair kernel a cube's rows are [regular tiles | RP tiles]
  int stats_nb_extra;  // regular kernel: RP tiles per cube (0 without RP): row = spatial + nb * stats_nb_extra
  int stats_tile0;     // remainder-pair kernel: regular tiles per cube: row = spatial + (nb + 1) * stats_tile0
  const uint8_t* wpacked;
  // MODE 0
  __half* out_raw;       // [NB][D][H][W][ldo] raw conv output, fp16
  float* stats_partial;  // [spatial tile][2][ldo]
  int ldo;               // total output channels (row pitch of out_raw)
  // MODE 1
  __half* out_f16;  // [NB][2D][2H][2W][ld1]
  const float* bias;
  int ld1, coff1, cout1;
};

struct TileCoord {
  int n_tile, nb, d0, h0, w0, spatial;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvTcArgs& a, int tile, int td) {
  // the N tile is the FASTEST index: the CTAs that share a spatial tile (same input planes, different output
  // channels) run at the same time, so the input is read from DRAM once and from L2 by the others.  (With the N tile
  // outermost the transposed conv, 4-8 N tiles, re-read its whole input from DRAM per N tile: ncu 2.81 GB vs 0.70.)
  TileCoord t;
  t.n_tile = tile % a.n_tiles;
  int r = tile / a.n_tiles;
  t.spatial = r;
  const int wt = r % a.tiles_w;
  r /= a.tiles_w;
  const int ht = r % a.tiles_h;
  r /= a.tiles_h;
  const int dt = r % a.tiles_d;
  t.nb = r / a.tiles_d;
  t.w0 = wt * TW;
  t.h0 = ht * TH;
  t.d0 = dt * td;
  return t;
}

// fp32 pair -> packed fp16x2, saturating at the largest finite half instead of overflowing to inf
__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {
  a = fminf(fmaxf(a, -65504.f), 65504.f);
  b = fminf(fmaxf(b, -65504.f), 65504.f);
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// fp32 pair -> packed bf16x2 (bf16 shares fp32's exponent range: no saturation needed)
__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack_pair(float a, float b) {
  if constexpr (BF16) return pack_bf162(a, b);
  return pack_half2_sat(a, b);
}

// Column sums over the 32 lanes of a warp for 32 per-lane values: afterwards lane l holds sum_lanes v[l].
// Recursive halving: 16+8+4+2+1 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
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

// XF (fused InstanceNorm-apply + ReLU of the PREVIOUS layer): the TMA producer stages the previous layer's RAW fp16
// output; four extra warps (8..11) rewrite every landed halo plane in place — relu((x - mean) * rstd) in fp32, back
// to fp16 — before handing it to the MMA warp.  Rows outside the volume were zero-filled by TMA and are left
// untouched, which is exactly the conv's zero padding of the NORMALISED tensor.  This removes the separate
// read-raw / write-normalised pass between two convolutions.
// BF16: operands and output are bf16 and no statistics are taken — the data-gradient pass of the same convolution
// (gradients need bf16's exponent range; see conv3d_k3_dgrad).
template <int KS, int BN, int TD, int MODE, bool STACK, bool XF, bool BF16 = false, int KSD = KS>
__global__ void __launch_bounds__(XF ? 384 : 256, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapOut,
                 const ConvTcArgs args) {
  using C = ConvCfg<KS, BN, TD, STACK, MODE == 1, KSD>;
  // __align__(1024), not a hand-rounded pointer: an integer round trip makes the compiler forget the address space
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smA = smem;
  uint8_t* smB = smem + C::NSLOT * C::PLANE_BYTES;
  uint8_t* smOut = smB + C::NBST * C::BSTAGE_BYTES;  // 1024-aligned: every region is a multiple of 1 KB
  uint8_t* aux = smOut + C::OUT_BYTES;
  uint64_t* planeFull = reinterpret_cast<uint64_t*>(aux);
  uint64_t* planeEmpty = planeFull + C::NSLOT;
  uint64_t* bFull = planeEmpty + C::NSLOT;
  uint64_t* bEmpty = bFull + C::NBST;
  uint64_t* accFull = bEmpty + C::NBST;
  uint64_t* accEmpty = accFull + 2;
  uint64_t* planeReady = accEmpty + 2;  // XF only: plane transformed, MMA may read it
  uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(planeReady + C::NSLOT);
  float* statScratch = reinterpret_cast<float*>(aux + 1024);  // [4 warps][2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmapA);
    if constexpr (MODE == 1) ptx::prefetch_tmap(&tmapOut);
    for (int i = 0; i < C::NSLOT; ++i) {
      ptx::mbar_init(&planeFull[i], 1);
      ptx::mbar_init(&planeEmpty[i], 1);
      ptx::mbar_init(&planeReady[i], 1);
    }
    for (int i = 0; i < C::NBST; ++i) {
      ptx::mbar_init(&bFull[i], 1);
      ptx::mbar_init(&bEmpty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&accFull[i], 1);
      ptx::mbar_init(&accEmpty[i], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<C::TMEM_COLS>(tmemPtr);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmemPtr;

  const int first_tile = blockIdx.x;
  const int tile_stride = gridDim.x;

  if (warp == 0) {
    // ------------------------------------------------------------ halo-plane producer
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      for (int tile = first_tile; tile < args.total_tiles; tile += tile_stride) {
        const TileCoord t = decode_tile(args, tile, TD);
        for (int c = 0; c < args.chunks; ++c) {
          for (int i = 0; i < C::PPC; ++i) {
            ptx::mbar_wait(&planeEmpty[slot], ph ^ 1);
            ptx::mbar_arrive_expect_tx(&planeFull[slot], C::PLANE_BOX_BYTES);
            ptx::tma_load_5d(smA + slot * C::PLANE_BYTES, &tmapA, &planeFull[slot], c * 64, t.w0 - C::PAD,
                             t.h0 - C::PAD, t.d0 - C::PAD_D + i, t.nb);
            if (++slot == C::NSLOT) {
              slot = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------ weight-stage producer
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int tile = first_tile; tile < args.total_tiles; tile += tile_stride) {
        const TileCoord t = decode_tile(args, tile, TD);
        // packed stages per chunk: 27 taps, or (STACK) 9 (kh,kw) blocks of 192 rows walked once per phase
        constexpr int PACKED_PER_CHUNK = STACK ? KS * KS : C::TAPS;
        const uint8_t* wsrc =
            args.wpacked + static_cast<size_t>(t.n_tile) * args.chunks * PACKED_PER_CHUNK * C::BSTAGE_BYTES;
        const int nst = args.chunks * C::STAGES_PER_CHUNK;
        for (int s = 0; s < nst; ++s) {
          const int cc = s / C::STAGES_PER_CHUNK;
          const int src = cc * PACKED_PER_CHUNK + (s - cc * C::STAGES_PER_CHUNK) % PACKED_PER_CHUNK;
          ptx::mbar_wait(&bEmpty[st], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&bFull[st], C::BSTAGE_BYTES);
          ptx::bulk_load(smB + st * C::BSTAGE_BYTES, wsrc + static_cast<size_t>(src) * C::BSTAGE_BYTES,
                         C::BSTAGE_BYTES, &bFull[st]);
          if (++st == C::NBST) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The whole warp walks the (warp-uniform) loops so that every address / descriptor lives in uniform
    // registers; one elected lane issues the tcgen05 instructions.  (Issuing from a divergent single-lane
    // region makes the compiler broadcast each operand through R2UR per MMA, which costs more than the MMA.)
    constexpr uint32_t FMT = BF16 ? ((1u << 7) | (1u << 10)) : 0u;  // A / B operand format: 0 = fp16, 1 = bf16
    constexpr uint32_t idesc = ptx::make_idesc_f16(128, BN) | FMT;
    // high descriptor words are constant: SBO [32,46), version 1 at bit 46, SWIZZLE_128B (2) at [61,64)
    constexpr uint32_t A_HI = ((C::HALO_W * 128u) >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t B_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t LO_FLAGS = 1u << 16;  // LBO field = 1 (ignored for swizzled K-major)
    const uint32_t smA_u32 = ptx::smem_u32(smA);
    const uint32_t smB_u32 = ptx::smem_u32(smB);
    int pslot = 0;  // ring position of plane 0 of the current (tile, chunk)
    uint32_t pph = 0;
    int bst = 0;
    uint32_t bph = 0;
    int it = 0;
    for (int tile = first_tile; tile < args.total_tiles; tile += tile_stride, ++it) {
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      ptx::mbar_wait(&accEmpty[buf], (use & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t acc0 = tmem_base + buf * C::ACC_COLS;
      for (int c = 0; c < args.chunks; ++c) {
        if constexpr (STACK) {
          constexpr uint32_t idesc64 = ptx::make_idesc_f16(128, 64) | FMT,
                             idesc128 = ptx::make_idesc_f16(128, 128) | FMT,
                             idesc192 = ptx::make_idesc_f16(128, 192) | FMT;
          for (int phase = 0; phase < 2; ++phase) {
            for (int i = 0; i < 3; ++i) {
              int s = pslot + 3 * phase + i;
              uint32_t p = pph;
              if (s >= C::NSLOT) {
                s -= C::NSLOT;
                p ^= 1;
              }
              ptx::mbar_wait(XF ? &planeReady[s] : &planeFull[s], p);
            }
            ptx::tc_fence_after();
            for (int khw = 0; khw < 9; ++khw) {
              const int kh = khw / 3, kw = khw - kh * 3;
              ptx::mbar_wait(&bFull[bst], bph);
              ptx::tc_fence_after();
              const uint32_t b_lo = (((smB_u32 + bst * C::BSTAGE_BYTES) >> 4) & 0x3FFF) | LO_FLAGS;
              const uint32_t tap_off = (kh * C::HALO_W + kw) * 128;
              const bool first_tap = (c == 0 && khw == 0);
              if (ptx::elect_one()) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                  const int pi = 3 * phase + i;                    // input plane 0..5 of this (tile, chunk)
                  const int j_lo = pi - 2 > 0 ? pi - 2 : 0;        // output planes fed by this input plane
                  const int j_hi = pi < 3 ? pi : 3;
                  int s = pslot + pi;
                  if (s >= C::NSLOT) s -= C::NSLOT;
                  const uint32_t a_lo = (((smA_u32 + s * C::PLANE_BYTES + tap_off) >> 4) & 0x3FFF) | LO_FLAGS;
                  if (first_tap) {
                    // the first MMA into an accumulator must overwrite it: issue this tap unstacked, kd = 0 first
                    for (int j = j_hi; j >= j_lo; --j) {
                      const int kd = pi - j;
                      const uint32_t bj = b_lo + (((2 - kd) * 64 * 128) >> 4);
#pragma unroll
                      for (int k = 0; k < 4; ++k)
                        ptx::umma_f16(acc0 + j * 64, (static_cast<uint64_t>(A_HI) << 32) | (a_lo + 2 * k),
                                      (static_cast<uint64_t>(B_HI) << 32) | (bj + 2 * k), idesc64,
                                      (kd == 0 && k == 0) ? 0u : 1u);
                    }
                  } else {
                    // rows [ (2-kd_max)*64, ... ) of the [kd2|kd1|kd0] stage line up with columns j_lo..j_hi
                    const int nj = j_hi - j_lo + 1;
                    const uint32_t bj = b_lo + (((2 - (pi - j_lo)) * 64 * 128) >> 4);
                    const uint32_t idn = nj == 1 ? idesc64 : nj == 2 ? idesc128 : idesc192;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      ptx::umma_f16(acc0 + j_lo * 64, (static_cast<uint64_t>(A_HI) << 32) | (a_lo + 2 * k),
                                    (static_cast<uint64_t>(B_HI) << 32) | (bj + 2 * k), idn, 1u);
                  }
                }
                ptx::umma_commit(&bEmpty[bst]);
                if (khw == 8) {
                  for (int i = 0; i < 3; ++i) {
                    int s = pslot + 3 * phase + i;
                    if (s >= C::NSLOT) s -= C::NSLOT;
                    ptx::umma_commit(&planeEmpty[s]);
                  }
                  if (phase == 1 && c == args.chunks - 1) ptx::umma_commit(&accFull[buf]);
                }
              }
              __syncwarp();
              if (++bst == C::NBST) {
                bst = 0;
                bph ^= 1;
              }
            }
          }
        } else {
          int waited = 0;
          for (int kd = 0; kd < KSD; ++kd) {
            for (; waited <= TD - 1 + kd; ++waited) {
              int s = pslot + waited;
              uint32_t p = pph;
              if (s >= C::NSLOT) {
                s -= C::NSLOT;
                p ^= 1;
              }
              ptx::mbar_wait(XF ? &planeReady[s] : &planeFull[s], p);
            }
            ptx::tc_fence_after();
            for (int khw = 0; khw < KS * KS; ++khw) {
              const int kh = khw / KS, kw = khw - kh * KS;
              ptx::mbar_wait(&bFull[bst], bph);
              ptx::tc_fence_after();
              const uint32_t b_lo = (((smB_u32 + bst * C::BSTAGE_BYTES) >> 4) & 0x3FFF) | LO_FLAGS;
              const uint32_t first = (c == 0 && kd == 0 && khw == 0) ? 0u : 1u;
              const uint32_t tap_off = (kh * C::HALO_W + kw) * 128;
              if (ptx::elect_one()) {
  #pragma unroll
                for (int j = 0; j < TD; ++j) {
                  int s = pslot + j + kd;
                  if (s >= C::NSLOT) s -= C::NSLOT;
                  const uint32_t a_lo = (((smA_u32 + s * C::PLANE_BYTES + tap_off) >> 4) & 0x3FFF) | LO_FLAGS;
  #pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint64_t adesc = (static_cast<uint64_t>(A_HI) << 32) | (a_lo + 2 * k);
                    const uint64_t bdesc = (static_cast<uint64_t>(B_HI) << 32) | (b_lo + 2 * k);
                    ptx::umma_f16(acc0 + j * BN, adesc, bdesc, idesc, (k == 0) ? first : 1u);
                  }
                }
                ptx::umma_commit(&bEmpty[bst]);
                // planes whose last reader was this kd group go back to the producer
                if (khw == KS * KS - 1) {
                  if (kd < KSD - 1) {
                    int s = pslot + kd;
                    if (s >= C::NSLOT) s -= C::NSLOT;
                    ptx::umma_commit(&planeEmpty[s]);
                  } else {
                    for (int i = KSD - 1; i < C::PPC; ++i) {
                      int s = pslot + i;
                      if (s >= C::NSLOT) s -= C::NSLOT;
                      ptx::umma_commit(&planeEmpty[s]);
                    }
                    if (c == args.chunks - 1) ptx::umma_commit(&accFull[buf]);
                  }
                }
              }
              __syncwarp();
              if (++bst == C::NBST) {
                bst = 0;
                bph ^= 1;
              }
            }
          }
        }
        pslot += C::PPC;
        if (pslot >= C::NSLOT) {
          pslot -= C::NSLOT;
          pph ^= 1;
        }
      }
    }
  } else if (XF && warp >= 8) {
    // ------------------------------------------------------------ in-place InstanceNorm + ReLU of landed planes
    const int tt = threadIdx.x - 256;  // 0..127
    const int g = tt & 7;              // logical 16-byte unit = channels [8g, 8g+8) of the chunk
    const int r0 = tt >> 3;            // rows r0, r0+16, ...
    int slot = 0;
    uint32_t ph = 0;
    for (int tile = first_tile; tile < args.total_tiles; tile += tile_stride) {
      const TileCoord t = decode_tile(args, tile, TD);
      for (int c = 0; c < args.chunks; ++c) {
        float mu[8], rs[8];
        const float* mr = args.in_mr + static_cast<size_t>(t.nb) * 2 * args.cin_total + c * 64 + g * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          mu[i] = __ldg(mr + i);
          rs[i] = __ldg(mr + args.cin_total + i);
        }
        for (int i = 0; i < C::PPC; ++i) {
          const int d = t.d0 - C::PAD_D + i;
          const bool plane_valid = d >= 0 && d < args.D;
          ptx::mbar_wait(&planeFull[slot], ph);
          uint8_t* pl = smA + slot * C::PLANE_BYTES;
          if (plane_valid) {
#pragma unroll 4
            for (int row = r0; row < C::PLANE_ROWS; row += 16) {
              const int hh = row / C::HALO_W, ww = row - hh * C::HALO_W;
              const int h = t.h0 - C::PAD + hh, w = t.w0 - C::PAD + ww;
              if (h >= 0 && h < args.H && w >= 0 && w < args.W) {
                uint4* p = reinterpret_cast<uint4*>(pl + row * 128 + ((g ^ (row & 7)) << 4));
                uint4 v = *p;
                __half2* h2 = reinterpret_cast<__half2*>(&v);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __half22float2(h2[e]);
                  h2[e] = __floats2half2_rn(fmaxf((f.x - mu[2 * e]) * rs[2 * e], 0.f),
                                            fmaxf((f.y - mu[2 * e + 1]) * rs[2 * e + 1], 0.f));
                }
                *p = v;
              }
            }
          }
          ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
          ptx::named_bar_sync(4, 128);
          if (tt == 0) ptx::mbar_arrive(&planeReady[slot]);
          if (++slot == C::NSLOT) {
            slot = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int q = warp & 3;    // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;  // accumulator row = voxel inside the tile plane
    const int mw = m & 7, mh = m >> 3;
    int it = 0;
    [[maybe_unused]] int ost = 0;  // running index of output staging tiles (MODE 1)
    for (int tile = first_tile; tile < args.total_tiles; tile += tile_stride, ++it) {
      const TileCoord t = decode_tile(args, tile, TD);
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      ptx::mbar_wait(&accFull[buf], use & 1);
      ptx::tc_fence_after();
      const uint32_t acc0 = tmem_base + buf * C::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
      const int w = t.w0 + mw, h = t.h0 + mh;
      const bool valid_hw = (w < args.W) && (h < args.H);

      float csum[BN / 32], csq[BN / 32];
#pragma unroll
      for (int cc = 0; cc < BN / 32; ++cc) csum[cc] = csq[cc] = 0.f;

      if constexpr (MODE == 1) {
        // Transposed conv: every (plane, 64-column group) = 128 coarse voxels x 64 channels of one tap is staged in
        // shared memory as a swizzled 128 x 128 B tile and written by ONE TMA store whose tensor map walks the fine
        // grid with element stride 2 in w and h (pixel shuffle); volume overhang is clipped by the TMA unit.
        // (Per-thread 16-byte scatter stores cost one LSU wavefront each and made this kernel LSU-bound.)
        const bool leader = threadIdx.x == 128;
#pragma unroll 1
        for (int j = 0; j < TD; ++j) {
          const int d = t.d0 + j;
          if (d >= args.D) break;  // warp-uniform
#pragma unroll 1
          for (int g = 0; g < BN / 64; ++g) {
            const int sb = ost & 1;
            if (leader) ptx::bulk_wait_group_read<1>();  // the store that last used this staging tile has read it
            ptx::named_bar_sync(2, 128);
            uint8_t* stg = smOut + sb * (128 * 128) + m * 128;
            const int n0 = t.n_tile * BN + g * 64;
            const int tap = n0 / args.cout1;
            const int co0 = n0 - tap * args.cout1;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t raw[32];
              ptx::tmem_ld32(acc0 + j * BN + g * 64 + half * 32, raw);
              ptx::tmem_ld_wait();
              const float4* bp = reinterpret_cast<const float4*>(args.bias + co0 + half * 32);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 b0 = __ldg(bp + 2 * i), b1 = __ldg(bp + 2 * i + 1);
                const uint4 pk = make_uint4(
                    pack_half2_sat(__uint_as_float(raw[8 * i]) + b0.x, __uint_as_float(raw[8 * i + 1]) + b0.y),
                    pack_half2_sat(__uint_as_float(raw[8 * i + 2]) + b0.z, __uint_as_float(raw[8 * i + 3]) + b0.w),
                    pack_half2_sat(__uint_as_float(raw[8 * i + 4]) + b1.x, __uint_as_float(raw[8 * i + 5]) + b1.y),
                    pack_half2_sat(__uint_as_float(raw[8 * i + 6]) + b1.z, __uint_as_float(raw[8 * i + 7]) + b1.w));
                const int unit = half * 4 + i;
                *reinterpret_cast<uint4*>(stg + ((unit ^ (m & 7)) << 4)) = pk;
              }
            }
            ptx::fence_proxy_async();
            ptx::named_bar_sync(2, 128);
            if (leader) {
              ptx::tma_store_5d(&tmapOut, smOut + sb * (128 * 128), args.coff1 + co0, 2 * t.w0 + (tap & 1),
                                2 * t.h0 + ((tap >> 1) & 1), 2 * d + (tap >> 2), t.nb);
              ptx::bulk_commit_group();
            }
            ++ost;
          }
        }
      }
#pragma unroll 1
      for (int j = 0; MODE == 0 && j < TD; ++j) {
        const int d = t.d0 + j;
        if (d >= args.D) break;  // warp-uniform
#pragma unroll
        for (int cc = 0; cc < BN / 32; ++cc) {
          uint32_t raw[32];
          ptx::tmem_ld32(acc0 + j * BN + cc * 32, raw);
          ptx::tmem_ld_wait();
          if constexpr (MODE == 0) {
            const int co = t.n_tile * BN + cc * 32;
            if (valid_hw) {
              uint4* dst = reinterpret_cast<uint4*>(
                  args.out_raw + (((static_cast<size_t>(t.nb) * args.D + d) * args.H + h) * args.W + w) * args.ldo +
                  co);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                dst[i] = make_uint4(pack_pair<BF16>(__uint_as_float(raw[8 * i]), __uint_as_float(raw[8 * i + 1])),
                                    pack_pair<BF16>(__uint_as_float(raw[8 * i + 2]), __uint_as_float(raw[8 * i + 3])),
                                    pack_pair<BF16>(__uint_as_float(raw[8 * i + 4]), __uint_as_float(raw[8 * i + 5])),
                                    pack_pair<BF16>(__uint_as_float(raw[8 * i + 6]), __uint_as_float(raw[8 * i + 7])));
            }
            if constexpr (!BF16) {
              float v[32], v2[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x = valid_hw ? __uint_as_float(raw[i]) : 0.f;
                v[i] = x;
                v2[i] = x * x;
              }
              csum[cc] += warp_colsum32(v, lane);
              csq[cc] += warp_colsum32(v2, lane);
            }
          }
        }
      }
      // accumulators are drained: hand the TMEM set back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&accEmpty[buf]);

      if constexpr (MODE == 0 && !BF16) {
        // lane l of warp q holds the column-l sums of its 32 rows; combine the 4 warps in fixed order
        float* mine = statScratch + q * 2 * BN;
#pragma unroll
        for (int cc = 0; cc < BN / 32; ++cc) {
          mine[cc * 32 + lane] = csum[cc];
          mine[BN + cc * 32 + lane] = csq[cc];
        }
        ptx::named_bar_sync(1, 128);
        const int e = threadIdx.x - 128;  // 0..127
        for (int i = e; args.stats_partial != nullptr && i < 2 * BN; i += 128) {
          const float s = (statScratch[i] + statScratch[2 * BN + i]) + (statScratch[4 * BN + i] + statScratch[6 * BN + i]);
          const int which = i / BN, col = i - which * BN;
          args.stats_partial[(static_cast<size_t>(t.spatial + t.nb * args.stats_nb_extra) * 2 + which) * args.ldo +
                             t.n_tile * BN + col] = s;
        }
        ptx::named_bar_sync(1, 128);
      }
    }
    if constexpr (MODE == 1) {
      if (threadIdx.x == 128) ptx::bulk_wait_group_read<0>();  // shared memory must outlive the last TMA stores
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ remainder pairs
// Tile quantisation in h.  The MMA's 128 rows are 16 eight-row groups one halo LINE apart, i.e. 16 h-lines of 8 voxels;
// a plane of H = 16 q + R lines costs q + 1 h-tiles, the last one with only R useful lines (70^3: 6 of 16, 35^3: 3 of
// 16 — 17.6 % / 61 % more MMA rows than voxels over those levels).  For R <= 6 the remainder strip of TWO consecutive
// d-planes fits ONE accumulator: the strip of plane d is staged as a box of 8 lines (R <= 6 output lines + 2 halo
// lines), the boxes of planes d-1, d, d+1, ... are laid back to back in shared memory, and because a box is exactly 8
// lines = half of the 16 groups, a 16-line window starting at line (kd * 8 + kh) of that chain reads
//     groups 0..7  -> plane d   + kd - 1, lines h0 - 1 + kh ..      groups 8..15 -> plane d+1 + kd - 1, same lines
// with the SAME uniform group stride the regular kernel uses.  Groups 6, 7 / 14, 15 compute rows nobody stores.  One
// tile = 2 accumulators = 4 output planes fed by a chain of 6 boxes (60 KB per 64-channel chunk, double buffered);
// weights, K order (chunk, kd, kh, kw), epilogue arithmetic and statistics are those of the regular kernel, so every
// stored value is bit-identical to what the 16-line tile would have produced.  70^3: 20 -> 18 MMA tiles per four
// planes, 35^3: 12 -> 10.
template <int BN>
struct RpCfg {
  static constexpr int LINES = 8;
  static constexpr int HALO_W = TW + 2;
  static constexpr int BOX_ROWS = LINES * HALO_W;
  static constexpr int BOX_BYTES = BOX_ROWS * 128;  // 10240: a multiple of the 1 KB swizzle atom
  static constexpr int NACC = 2;                    // accumulators per tile
  static constexpr int PLANES = 2 * NACC;           // output planes per tile
  static constexpr int NBOX = PLANES + 2;
  static constexpr int CHAIN_BYTES = NBOX * BOX_BYTES;
  static constexpr int NCHAIN = 2;
  static constexpr int BSTAGE_BYTES = BN * 128;
  static constexpr int AUX_BYTES = 1024 + 4 * BN * 2 * 4;
  static constexpr int SMEM_LIMIT = 232448;
  static constexpr int NBST_FIT = (SMEM_LIMIT - 1024 - AUX_BYTES - NCHAIN * CHAIN_BYTES) / BSTAGE_BYTES;
  static constexpr int NBST = NBST_FIT > 8 ? 8 : NBST_FIT;
  static constexpr int ACC_COLS = NACC * BN;
  static constexpr int TMEM_COLS = pow2_at_least(2 * ACC_COLS);
  // the last window over-reads 2 lines past its chain (rows nobody stores): chain 1 is followed by the weight ring
  static constexpr int SMEM_BYTES = 1024 + NCHAIN * CHAIN_BYTES + NBST * BSTAGE_BYTES + AUX_BYTES;
  static_assert(BOX_BYTES % 1024 == 0, "boxes must keep the swizzle phase");
  static_assert(NBST >= 2 && 2 * ACC_COLS <= 512, "RP configuration does not fit");
};

struct RpCoord {
  int n_tile, nb, d0, w0, spatial;
};
__device__ __forceinline__ RpCoord decode_rp(const ConvTcArgs& a, int tile) {
  RpCoord t;
  t.n_tile = tile % a.n_tiles;
  int r = tile / a.n_tiles;
  t.spatial = r;
  const int wt = r % a.tiles_w;
  r /= a.tiles_w;
  const int dt = r % a.tiles_d;
  t.nb = r / a.tiles_d;
  t.w0 = wt * TW;
  t.d0 = dt * 4;
  return t;
}

template <int BN, bool XF, bool BF16 = false>
__global__ void __launch_bounds__(XF ? 384 : 256, 1)
conv3d_rp_kernel(const __grid_constant__ CUtensorMap tmapA, const ConvTcArgs args) {
  using C = RpCfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smA = smem;
  uint8_t* smB = smem + C::NCHAIN * C::CHAIN_BYTES;
  uint8_t* aux = smB + C::NBST * C::BSTAGE_BYTES;
  uint64_t* chainFull = reinterpret_cast<uint64_t*>(aux);
  uint64_t* chainEmpty = chainFull + C::NCHAIN;
  uint64_t* chainReady = chainEmpty + C::NCHAIN;  // XF only
  uint64_t* bFull = chainReady + C::NCHAIN;
  uint64_t* bEmpty = bFull + C::NBST;
  uint64_t* accFull = bEmpty + C::NBST;
  uint64_t* accEmpty = accFull + 2;
  uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(accEmpty + 2);
  float* statScratch = reinterpret_cast<float*>(aux + 1024);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmapA);
    for (int i = 0; i < C::NCHAIN; ++i) {
      ptx::mbar_init(&chainFull[i], 1);
      ptx::mbar_init(&chainEmpty[i], 1);
      ptx::mbar_init(&chainReady[i], 1);
    }
    for (int i = 0; i < C::NBST; ++i) {
      ptx::mbar_init(&bFull[i], 1);
      ptx::mbar_init(&bEmpty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&accFull[i], 1);
      ptx::mbar_init(&accEmpty[i], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<C::TMEM_COLS>(tmemPtr);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmemPtr;
  const int first_tile = blockIdx.x, tile_stride = gridDim.x;

  if (warp == 0) {
    // ------------------------------------------------------------ chain producer: 6 boxes of 8 lines per chunk
    if (lane == 0) {
      int cb = 0;
      uint32_t ph = 0;
      for (int tile = first_tile; tile < args.total_tiles; tile += tile_stride) {
        const RpCoord t = decode_rp(args, tile);
        for (int c = 0; c < args.chunks; ++c) {
          ptx::mbar_wait(&chainEmpty[cb], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&chainFull[cb], C::CHAIN_BYTES);
          for (int b = 0; b < C::NBOX; ++b)
            ptx::tma_load_5d(smA + cb * C::CHAIN_BYTES + b * C::BOX_BYTES, &tmapA, &chainFull[cb], c * 64, t.w0 - 1,
                             args.rp_h0 - 1, t.d0 - 1 + b, t.nb);
          if (++cb == C::NCHAIN) {
            cb = 0;
            ph ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------ weight-stage producer (the regular packed image)
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int tile = first_tile; tile < args.total_tiles; tile += tile_stride) {
        const RpCoord t = decode_rp(args, tile);
        const uint8_t* wsrc = args.wpacked + static_cast<size_t>(t.n_tile) * args.chunks * 27 * C::BSTAGE_BYTES;
        const int nst = args.chunks * 27;
        for (int s = 0; s < nst; ++s) {
          ptx::mbar_wait(&bEmpty[st], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&bFull[st], C::BSTAGE_BYTES);
          ptx::bulk_load(smB + st * C::BSTAGE_BYTES, wsrc + static_cast<size_t>(s) * C::BSTAGE_BYTES, C::BSTAGE_BYTES,
                         &bFull[st]);
          if (++st == C::NBST) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform loops, one elected lane)
    constexpr uint32_t FMT = BF16 ? ((1u << 7) | (1u << 10)) : 0u;  // A / B operand format: 0 = fp16, 1 = bf16
    constexpr uint32_t idesc = ptx::make_idesc_f16(128, BN) | FMT;
    constexpr uint32_t A_HI = ((C::HALO_W * 128u) >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t B_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t LO_FLAGS = 1u << 16;
    const uint32_t smA_u32 = ptx::smem_u32(smA);
    const uint32_t smB_u32 = ptx::smem_u32(smB);
    int cb = 0, bst = 0, it = 0;
    uint32_t cph = 0, bph = 0;
    for (int tile = first_tile; tile < args.total_tiles; tile += tile_stride, ++it) {
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      ptx::mbar_wait(&accEmpty[buf], (use & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t acc0 = tmem_base + buf * C::ACC_COLS;
      for (int c = 0; c < args.chunks; ++c) {
        ptx::mbar_wait(XF ? &chainReady[cb] : &chainFull[cb], cph);
        ptx::tc_fence_after();
        const uint32_t chain = smA_u32 + cb * C::CHAIN_BYTES;
        for (int kd = 0; kd < 3; ++kd) {
          for (int khw = 0; khw < 9; ++khw) {
            const int kh = khw / 3, kw = khw - kh * 3;
            ptx::mbar_wait(&bFull[bst], bph);
            ptx::tc_fence_after();
            const uint32_t b_lo = (((smB_u32 + bst * C::BSTAGE_BYTES) >> 4) & 0x3FFF) | LO_FLAGS;
            const uint32_t first = (c == 0 && kd == 0 && khw == 0) ? 0u : 1u;
            if (ptx::elect_one()) {
#pragma unroll
              for (int j = 0; j < C::NACC; ++j) {
                const uint32_t a_addr = chain + (((2 * j + kd) * C::LINES + kh) * C::HALO_W + kw) * 128;
                const uint32_t a_lo = ((a_addr >> 4) & 0x3FFF) | LO_FLAGS;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  ptx::umma_f16(acc0 + j * BN, (static_cast<uint64_t>(A_HI) << 32) | (a_lo + 2 * k),
                                (static_cast<uint64_t>(B_HI) << 32) | (b_lo + 2 * k), idesc, (k == 0) ? first : 1u);
              }
              ptx::umma_commit(&bEmpty[bst]);
              if (kd == 2 && khw == 8) {
                ptx::umma_commit(&chainEmpty[cb]);
                if (c == args.chunks - 1) ptx::umma_commit(&accFull[buf]);
              }
            }
            __syncwarp();
            if (++bst == C::NBST) {
              bst = 0;
              bph ^= 1;
            }
          }
        }
        if (++cb == C::NCHAIN) {
          cb = 0;
          cph ^= 1;
        }
      }
    }
  } else if (XF && warp >= 8) {
    // ------------------------------------------------------------ in-place InstanceNorm + ReLU of the landed chain
    const int tt = threadIdx.x - 256;
    const int g = tt & 7;
    const int r0 = tt >> 3;
    int cb = 0;
    uint32_t ph = 0;
    for (int tile = first_tile; tile < args.total_tiles; tile += tile_stride) {
      const RpCoord t = decode_rp(args, tile);
      for (int c = 0; c < args.chunks; ++c) {
        float mu[8], rs[8];
        const float* mr = args.in_mr + static_cast<size_t>(t.nb) * 2 * args.cin_total + c * 64 + g * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          mu[i] = __ldg(mr + i);
          rs[i] = __ldg(mr + args.cin_total + i);
        }
        ptx::mbar_wait(&chainFull[cb], ph);
        uint8_t* ch = smA + cb * C::CHAIN_BYTES;
#pragma unroll 2
        for (int row = r0; row < C::NBOX * C::BOX_ROWS; row += 16) {
          const int b = row / C::BOX_ROWS, rb = row - b * C::BOX_ROWS;
          const int line = rb / C::HALO_W, ww = rb - line * C::HALO_W;
          const int d = t.d0 - 1 + b, h = args.rp_h0 - 1 + line, w = t.w0 - 1 + ww;
          if (d >= 0 && d < args.D && h >= 0 && h < args.H && w >= 0 && w < args.W) {
            uint4* p = reinterpret_cast<uint4*>(ch + row * 128 + ((g ^ (row & 7)) << 4));
            uint4 v = *p;
            __half2* h2 = reinterpret_cast<__half2*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(h2[e]);
              h2[e] = __floats2half2_rn(fmaxf((f.x - mu[2 * e]) * rs[2 * e], 0.f),
                                        fmaxf((f.y - mu[2 * e + 1]) * rs[2 * e + 1], 0.f));
            }
            *p = v;
          }
        }
        ptx::fence_proxy_async();
        ptx::named_bar_sync(4, 128);
        if (tt == 0) ptx::mbar_arrive(&chainReady[cb]);
        if (++cb == C::NCHAIN) {
          cb = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------ epilogue: raw fp16 store + statistics partials
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int mw = m & 7, grp = m >> 3;
    const int pl = grp >> 3, line = grp & 7;  // pl is warp-uniform (= q >> 1)
    int it = 0;
    for (int tile = first_tile; tile < args.total_tiles; tile += tile_stride, ++it) {
      const RpCoord t = decode_rp(args, tile);
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      ptx::mbar_wait(&accFull[buf], use & 1);
      ptx::tc_fence_after();
      const uint32_t acc0 = tmem_base + buf * C::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
      const int w = t.w0 + mw, h = args.rp_h0 + line;
      const bool valid_hw = (w < args.W) && (line < args.rp_rem);
      float csum[BN / 32], csq[BN / 32];
#pragma unroll
      for (int cc = 0; cc < BN / 32; ++cc) csum[cc] = csq[cc] = 0.f;
#pragma unroll 1
      for (int j = 0; j < C::NACC; ++j) {
        if (t.d0 + 2 * j >= args.D) break;  // CTA-uniform
        const int d = t.d0 + 2 * j + pl;
        const bool valid = valid_hw && d < args.D;
#pragma unroll
        for (int cc = 0; cc < BN / 32; ++cc) {
          uint32_t raw[32];
          ptx::tmem_ld32(acc0 + j * BN + cc * 32, raw);
          ptx::tmem_ld_wait();
          const int co = t.n_tile * BN + cc * 32;
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(
                args.out_raw + (((static_cast<size_t>(t.nb) * args.D + d) * args.H + h) * args.W + w) * args.ldo + co);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_uint4(pack_pair<BF16>(__uint_as_float(raw[8 * i]), __uint_as_float(raw[8 * i + 1])),
                                  pack_pair<BF16>(__uint_as_float(raw[8 * i + 2]), __uint_as_float(raw[8 * i + 3])),
                                  pack_pair<BF16>(__uint_as_float(raw[8 * i + 4]), __uint_as_float(raw[8 * i + 5])),
                                  pack_pair<BF16>(__uint_as_float(raw[8 * i + 6]), __uint_as_float(raw[8 * i + 7])));
          }
          if constexpr (!BF16) {
            float v[32], v2[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float x = valid ? __uint_as_float(raw[i]) : 0.f;
              v[i] = x;
              v2[i] = x * x;
            }
            csum[cc] += warp_colsum32(v, lane);
            csq[cc] += warp_colsum32(v2, lane);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&accEmpty[buf]);

      if constexpr (!BF16) {
        float* mine = statScratch + q * 2 * BN;
#pragma unroll
        for (int cc = 0; cc < BN / 32; ++cc) {
          mine[cc * 32 + lane] = csum[cc];
          mine[BN + cc * 32 + lane] = csq[cc];
        }
        ptx::named_bar_sync(1, 128);
        const int e = threadIdx.x - 128;
        for (int i = e; args.stats_partial != nullptr && i < 2 * BN; i += 128) {
          const float s =
              (statScratch[i] + statScratch[2 * BN + i]) + (statScratch[4 * BN + i] + statScratch[6 * BN + i]);
          const int which = i / BN, col = i - which * BN;
          args.stats_partial[(static_cast<size_t>(t.spatial + (t.nb + 1) * args.stats_tile0) * 2 + which) * args.ldo +
                             t.n_tile * BN + col] = s;
        }
        ptx::named_bar_sync(1, 128);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ weights
// Packed image: [n_tile][chunk][tap][row r < BN][128 B], 16-byte unit j of row r stored at unit j ^ (r & 7).
// conv:  w is OIDHW fp32 (Cout, Cin, k, k, k); GEMM column n = output channel.
// convT: w is IODHW fp32 (Cin, Cout, 2, 2, 2); GEMM column n = tap * Cout + co with tap = (a*2+b)*2+c.
// stacked (conv, Cout = 64): [chunk][kh*3+kw][row = (2-kd)*64 + co][128 B], i.e. stages of 192 rows [kd2|kd1|kd0].
// dgrad: the data gradient of a stride-1 "same" conv is the conv of dy with the channel-transposed, spatially
//        flipped filter: here Cout / Cin are the GEMM's N / K = the layer's Cin / Cout, w is still the layer's OIDHW
//        tensor, and the image is written as bf16.
__global__ void pack_weights_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int Cout, int Cin,
                                    int taps, int BN, int transposed, int stacked, int dgrad) {
  const int chunks = Cin / 64;
  const int ngemm = transposed ? 8 * Cout : Cout;
  const int gtaps = transposed ? 1 : taps;
  const size_t total = static_cast<size_t>(ngemm) * Cin * gtaps;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = idx % 64;
    size_t r0 = idx / 64;
    const int r = r0 % BN;
    r0 /= BN;
    const int tap = r0 % gtaps;
    r0 /= gtaps;
    const int chunk = r0 % chunks;
    const int n_tile = r0 / chunks;
    const int n = n_tile * BN + r;
    const int ci = chunk * 64 + k;
    float v;
    if (dgrad) {
      v = w[(static_cast<size_t>(ci) * Cout + n) * taps + (taps - 1 - tap)];
    } else if (!transposed) {
      v = w[(static_cast<size_t>(n) * Cin + ci) * taps + tap];
    } else {
      const int t8 = n / Cout, co = n - t8 * Cout;
      v = w[(static_cast<size_t>(ci) * Cout + co) * 8 + t8];
    }
    size_t off;
    if (stacked) {
      const int kd = tap / 9, khw = tap - kd * 9;
      const int row = (2 - kd) * 64 + r;
      off = (static_cast<size_t>(chunk) * 9 + khw) * 192 * 64 + static_cast<size_t>(row) * 64 +
            ((((k >> 3) ^ (row & 7)) << 3) | (k & 7));
    } else {
      const size_t stage = (static_cast<size_t>(n_tile) * chunks + chunk) * gtaps + tap;
      off = stage * BN * 64 + static_cast<size_t>(r) * 64 + ((((k >> 3) ^ (r & 7)) << 3) | (k & 7));
    }
    if (dgrad) {
      const __nv_bfloat16 b = __float2bfloat16_rn(v);
      out[off] = *reinterpret_cast<const uint16_t*>(&b);
    } else {
      const __half hv = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
      out[off] = *reinterpret_cast<const uint16_t*>(&hv);
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int make_act_tmap(CUtensorMap* m, const void* base, int C, int W, int H, int D, int NB, int boxW, int boxH,
                         bool bf16 = false) {
  auto encode = get_tensor_map_encoder();
  if (!encode) return set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)NB};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)boxW, (cuuint32_t)boxH, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5,
                      const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

static int g_max_ctas = 0;  // 0 = one CTA per SM
void debug_set_max_ctas(int n) { g_max_ctas = n; }

// Fine-grid output of the transposed conv: (ld, 2W, 2H, 2D, NB) fp16, box = 64 channels x 8 x 16 voxels taken with
// element stride 2 in w and h (the bounding box is 16 x 32), SWIZZLE_128B shared-memory image.
static int make_convT_out_tmap(CUtensorMap* m, const void* base, int ld, int W2, int H2, int D2, int NB) {
  auto encode = get_tensor_map_encoder();
  if (!encode) return set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[5] = {(cuuint64_t)ld, (cuuint64_t)W2, (cuuint64_t)H2, (cuuint64_t)D2, (cuuint64_t)NB};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 2, (cuuint64_t)W2 * ld * 2, (cuuint64_t)H2 * W2 * ld * 2,
                           (cuuint64_t)D2 * H2 * W2 * ld * 2};
  cuuint32_t box[5] = {64, 2 * TW, 2 * TH, 1, 1};
  cuuint32_t estr[5] = {1, 2, 2, 1, 1};
  CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled (convT output) failed (%d)", (int)r);
  return 0;
}

template <int KS, int BN, int TD, int MODE, bool STACK, bool XF, bool BF16 = false, int KSD = KS>
static int launch_one(const CUtensorMap& tm, const CUtensorMap& tmo, const ConvTcArgs& a, int smem_bytes,
                      cudaStream_t stream) {
  auto kern = conv3d_tc_kernel<KS, BN, TD, MODE, STACK, XF, BF16, KSD>;
  static bool attr_set[64] = {false};
  if (first_use_on_device(attr_set))
    NC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int cap = (g_max_ctas > 0 && g_max_ctas < num_sms()) ? g_max_ctas : num_sms();
  const int grid = a.total_tiles < cap ? a.total_tiles : cap;
  kern<<<grid, XF ? 384 : 256, smem_bytes, stream>>>(tm, tmo, a);
  NC_CUDA(cudaGetLastError());
  return 0;
}

template <int KS, int BN, int TD, int MODE, bool STACK = false, bool BF16 = false, int KSD = KS>
static int launch_cfg(const void* x, ConvTcArgs a, int Cin, cudaStream_t stream) {
  using C = ConvCfg<KS, BN, TD, STACK, MODE == 1, KSD>;
  CUtensorMap tm, tmo;
  if (int rc = make_act_tmap(&tm, x, Cin, a.W, a.H, a.D, a.NB, C::HALO_W, C::HALO_H, BF16)) return rc;
  if constexpr (MODE == 1) {
    if (int rc = make_convT_out_tmap(&tmo, a.out_f16, a.ld1, 2 * a.W, 2 * a.H, 2 * a.D, a.NB)) return rc;
  } else {
    tmo = tm;
  }
  a.tiles_w = (a.W + TW - 1) / TW;
  a.tiles_h = a.tiles_h_cap > 0 ? a.tiles_h_cap : (a.H + TH - 1) / TH;
  a.tiles_d = (a.D + TD - 1) / TD;
  a.total_tiles = a.n_tiles * a.NB * a.tiles_d * a.tiles_h * a.tiles_w;
  a.cin_total = Cin;
  if constexpr (BF16 || KSD != KS || KS == 5) {
    return launch_one<KS, BN, TD, MODE, STACK, false, BF16, KSD>(tm, tmo, a, C::SMEM_BYTES, stream);
  } else {
    if (a.in_mr) return launch_one<KS, BN, TD, MODE, STACK, true>(tm, tmo, a, C::SMEM_BYTES, stream);
    return launch_one<KS, BN, TD, MODE, STACK, false>(tm, tmo, a, C::SMEM_BYTES, stream);
  }
}

int conv3d_k3_bn(int Cout) { return Cout == 64 ? 64 : 128; }
int conv3d_k3_td(int Cout) { return Cout == 64 ? 4 : 2; }

// Remainder-pair kernel (see RpCfg): used by the N-tile-128 forward path when the last h-tile would hold <= 6 lines.
static bool g_rp_enabled = true;
void debug_set_remainder_pairs(int on) { g_rp_enabled = on != 0; }
static inline int rp_rem(int H, int Cout) {
  const int r = H % TH;
  return (g_rp_enabled && Cout != 64 && r >= 1 && r <= 6) ? r : 0;
}

size_t conv3d_k3_stats_tiles(int NB, int D, int H, int W, int Cout) {
  const int td = conv3d_k3_td(Cout);
  const size_t tw = (W + TW - 1) / TW;
  if (rp_rem(H, Cout))
    return static_cast<size_t>(NB) * ((D + td - 1) / td) * (H / TH) * tw + static_cast<size_t>(NB) * ((D + 3) / 4) * tw;
  return static_cast<size_t>(NB) * ((D + td - 1) / td) * ((H + TH - 1) / TH) * tw;
}

template <int BN, bool BF16 = false>
static int launch_rp(const void* x, ConvTcArgs a, int Cin, cudaStream_t stream) {
  using C = RpCfg<BN>;
  CUtensorMap tm;
  if (int rc = make_act_tmap(&tm, x, Cin, a.W, a.H, a.D, a.NB, C::HALO_W, C::LINES, BF16)) return rc;
  a.tiles_w = (a.W + TW - 1) / TW;
  a.tiles_h = 1;
  a.tiles_d = (a.D + 3) / 4;
  a.total_tiles = a.n_tiles * a.NB * a.tiles_d * a.tiles_w;
  a.cin_total = Cin;
  const int cap = (g_max_ctas > 0 && g_max_ctas < num_sms()) ? g_max_ctas : num_sms();
  const int grid = a.total_tiles < cap ? a.total_tiles : cap;
  static bool attr_set[2][64] = {{false}};
  if constexpr (BF16) {
    auto kern = conv3d_rp_kernel<BN, false, true>;
    static bool attr_bf16[64] = {false};
    if (first_use_on_device(attr_bf16))
      NC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    kern<<<grid, 256, C::SMEM_BYTES, stream>>>(tm, a);
  } else if (a.in_mr) {
    auto kern = conv3d_rp_kernel<BN, true>;
    if (first_use_on_device(attr_set[1]))
      NC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    kern<<<grid, 384, C::SMEM_BYTES, stream>>>(tm, a);
  } else {
    auto kern = conv3d_rp_kernel<BN, false>;
    if (first_use_on_device(attr_set[0]))
      NC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    kern<<<grid, 256, C::SMEM_BYTES, stream>>>(tm, a);
  }
  NC_CUDA(cudaGetLastError());
  return 0;
}

int conv3d_k3_fwd(const void* x, const float* in_mean_rstd, int NB, int D, int H, int W, int Cin, const void* wpacked,
                  int Cout, void* y_raw, float* stats_partial, cudaStream_t stream) {
  if (Cin % 64 || Cout % 64) return set_error("conv3d_k3_fwd: Cin and Cout must be multiples of 64");
  ConvTcArgs a{};
  a.W = W, a.H = H, a.D = D, a.NB = NB;
  a.chunks = Cin / 64;
  a.in_mr = in_mean_rstd;
  a.wpacked = static_cast<const uint8_t*>(wpacked);
  a.out_raw = static_cast<__half*>(y_raw);
  a.stats_partial = stats_partial;
  a.ldo = Cout;
  if (Cout == 64) {
    a.n_tiles = 1;
    return launch_cfg<3, 64, 4, 0, true>(x, a, Cin, stream);
  }
  if (Cout % 128) return set_error("conv3d_k3_fwd: Cout must be 64 or a multiple of 128");
  a.n_tiles = Cout / 128;
  if (const int rem = rp_rem(H, Cout)) {
    // full 16-line h-tiles by the regular kernel, the remainder strip (rem <= 6 lines) by the remainder-pair kernel
    const int full = H / TH, tw = (W + TW - 1) / TW;
    if (full > 0) {
      a.tiles_h_cap = full;
      a.stats_nb_extra = ((D + 3) / 4) * tw;
      if (int rc = launch_cfg<3, 128, 2, 0>(x, a, Cin, stream)) return rc;
    }
    a.tiles_h_cap = 0, a.stats_nb_extra = 0;
    a.rp_h0 = full * TH, a.rp_rem = rem;
    a.stats_tile0 = ((D + 1) / 2) * full * tw;
    return launch_rp<128>(x, a, Cin, stream);
  }
  return launch_cfg<3, 128, 2, 0>(x, a, Cin, stream);
}

// Data gradient of the k3 s1 p1 conv: dx = conv(dy, flipped / transposed filter) — the forward kernel with bf16
// operands, bf16 output and no statistics.  `wpacked` comes from pack_weights_dgrad (Cout, Cin are the LAYER's).
int conv3d_k3_dgrad(const void* dy, int NB, int D, int H, int W, int Cout, const void* wpacked, int Cin, void* dx,
                    cudaStream_t stream) {
  if (Cin % 64 || Cout % 64) return set_error("conv3d_k3_dgrad: Cin and Cout must be multiples of 64");
  ConvTcArgs a{};
  a.W = W, a.H = H, a.D = D, a.NB = NB;
  a.chunks = Cout / 64;  // the GEMM's K runs over the layer's output channels
  a.wpacked = static_cast<const uint8_t*>(wpacked);
  a.out_raw = static_cast<__half*>(dx);  // 16-bit elements; written as bf16
  a.ldo = Cin;
  if (Cin == 64) {
    a.n_tiles = 1;
    return launch_cfg<3, 64, 4, 0, true, true>(dy, a, Cout, stream);
  }
  if (Cin % 128) return set_error("conv3d_k3_dgrad: Cin must be 64 or a multiple of 128");
  a.n_tiles = Cin / 128;
  if (const int rem = rp_rem(H, Cin)) {   // remainder pairs, as in conv3d_k3_fwd (no statistics here)
    const int full = H / TH;
    if (full > 0) {
      a.tiles_h_cap = full;
      if (int rc = launch_cfg<3, 128, 2, 0, false, true>(dy, a, Cout, stream)) return rc;
    }
    a.tiles_h_cap = 0;
    a.rp_h0 = full * TH, a.rp_rem = rem;
    return launch_rp<128, true>(dy, a, Cout, stream);
  }
  return launch_cfg<3, 128, 2, 0, false, true>(dy, a, Cout, stream);
}

// Generic 64 -> 64 stride-1 "same" convolution for DeepLinearGenerator (reference models/networks.py:893-917): the
// k5 layer (ksd = ksp = 5) and the k7 Cin = 1 layer as a 7-tap depth conv over a 49(+15)-channel in-plane im2col
// (ksd = 7, ksp = 1).  fmt 0: fp16 in / out (forward), 1: bf16 (data gradient, packed with flip).  No statistics.
int conv3d_tc_64(const void* x, int fmt, int NB, int D, int H, int W, const void* wpacked, int ksd, int ksp, void* y,
                 cudaStream_t stream) {
  ConvTcArgs a{};
  a.W = W, a.H = H, a.D = D, a.NB = NB;
  a.chunks = 1;
  a.wpacked = static_cast<const uint8_t*>(wpacked);
  a.out_raw = static_cast<__half*>(y);
  a.ldo = 64;
  a.n_tiles = 1;
  if (ksd == 5 && ksp == 5)
    return fmt ? launch_cfg<5, 64, 2, 0, false, true>(x, a, 64, stream) : launch_cfg<5, 64, 2, 0>(x, a, 64, stream);
  if (ksd == 7 && ksp == 1)
    return fmt ? launch_cfg<1, 64, 4, 0, false, true, 7>(x, a, 64, stream)
               : launch_cfg<1, 64, 4, 0, false, false, 7>(x, a, 64, stream);
  return set_error("conv3d_tc_64: unsupported filter %d x %d x %d", ksd, ksp, ksp);
}

// Generic (unstacked, N tile 64) packed image of w = (Cout = 64, Cin = 64, taps) fp32; dgrad: channel-transposed,
// tap-reversed, bf16.
int pack_weights_64(const float* w, void* out, int taps, int dgrad, cudaStream_t stream) {
  pack_weights_kernel<<<num_sms() * 4, 256, 0, stream>>>(w, static_cast<uint16_t*>(out), 64, 64, taps, 64, 0, 0,
                                                         dgrad);
  NC_CUDA(cudaGetLastError());
  return 0;
}

// Data gradient of ConvTranspose3d(k2, s2): with the output gradient gathered space-to-depth into
// g[coarse voxel][tap * Cout + co] (space_to_depth_bf16), dx[v][ci] = sum_k g[v][k] * W[ci][k] is a 1x1x1 conv with
// K = 8 * Cout and N = Cin: the same kernel, KS = 1, bf16.  `wpacked` from pack_weights_convT_dgrad.
int conv3d_k1_bf16(const void* x, int NB, int D, int H, int W, int K, const void* wpacked, int N, void* y,
                   cudaStream_t stream) {
  if (K % 64 || N % 128) return set_error("conv3d_k1_bf16: K must be a multiple of 64 and N of 128");
  ConvTcArgs a{};
  a.W = W, a.H = H, a.D = D, a.NB = NB;
  a.chunks = K / 64;
  a.wpacked = static_cast<const uint8_t*>(wpacked);
  a.out_raw = static_cast<__half*>(y);
  a.ldo = N;
  a.n_tiles = N / 128;
  return launch_cfg<1, 128, 2, 0, false, true>(x, a, K, stream);
}

// w: IODHW fp32 (Cin, Cout, 2, 2, 2) of the transposed conv -> bf16 image [n_tile][chunk][row ci][64 k],
// k = tap * Cout + co; 8 * Cout * Cin * 2 bytes.
__global__ void pack_convT_dgrad_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int Cin, int Cout) {
  const int K = 8 * Cout, chunks = K / 64;
  const size_t total = static_cast<size_t>(Cin) * K;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int kk = idx % 64;
    size_t r0 = idx / 64;
    const int r = r0 % 128;
    r0 /= 128;
    const int chunk = r0 % chunks;
    const int n_tile = r0 / chunks;
    const int ci = n_tile * 128 + r, k = chunk * 64 + kk;
    const int tap = k / Cout, co = k - tap * Cout;
    const __nv_bfloat16 b = __float2bfloat16_rn(w[(static_cast<size_t>(ci) * Cout + co) * 8 + tap]);
    const size_t stage = static_cast<size_t>(n_tile) * chunks + chunk;
    out[stage * 128 * 64 + static_cast<size_t>(r) * 64 + ((((kk >> 3) ^ (r & 7)) << 3) | (kk & 7))] =
        *reinterpret_cast<const uint16_t*>(&b);
  }
}
int pack_weights_convT_dgrad(const float* w, void* out, int Cin, int Cout, cudaStream_t stream) {
  if (Cin % 128 || Cout % 8) return set_error("pack_weights_convT_dgrad: Cin must be a multiple of 128");
  pack_convT_dgrad_kernel<<<num_sms() * 4, 256, 0, stream>>>(w, static_cast<uint16_t*>(out), Cin, Cout);
  NC_CUDA(cudaGetLastError());
  return 0;
}

int convT3d_k2s2_fwd(const void* x, const float* in_mean_rstd, int NB, int D, int H, int W, int Cin,
                     const void* wpacked, const float* bias, int Cout, void* y, int y_ld, int y_coff,
                     cudaStream_t stream) {
  if (Cin % 64 || (8 * Cout) % 128 || Cout % 32) return set_error("convT3d_k2s2_fwd: unsupported channel counts");
  if (y_ld % 8 || y_coff % 8) return set_error("convT3d_k2s2_fwd: output slice must be 16-byte aligned");
  if (Cout % 64) return set_error("convT3d_k2s2_fwd: Cout must be a multiple of 64");
  if (reinterpret_cast<uintptr_t>(y) & 15) return set_error("convT3d_k2s2_fwd: output must be 16-byte aligned");
  ConvTcArgs a{};
  a.W = W, a.H = H, a.D = D, a.NB = NB;
  a.chunks = Cin / 64;
  a.in_mr = in_mean_rstd;
  a.wpacked = static_cast<const uint8_t*>(wpacked);
  a.out_f16 = static_cast<__half*>(y);
  a.bias = bias;
  a.ld1 = y_ld, a.coff1 = y_coff, a.cout1 = Cout;
  a.n_tiles = 8 * Cout / 128;
  return launch_cfg<1, 128, 2, 1>(x, a, Cin, stream);
}

size_t packed_weight_bytes(int Cout, int Cin, int taps, int transposed) {
  return static_cast<size_t>(transposed ? 8 * Cout : Cout) * Cin * (transposed ? 1 : taps) * 2;
}

int pack_weights(const float* w, void* out, int Cout, int Cin, int taps, int transposed, cudaStream_t stream) {
  if (Cin % 64) return set_error("pack_weights: Cin must be a multiple of 64");
  const int BN = transposed ? 128 : conv3d_k3_bn(Cout);
  const int ngemm = transposed ? 8 * Cout : Cout;
  if (ngemm % BN) return set_error("pack_weights: GEMM N not a multiple of the N tile");
  pack_weights_kernel<<<num_sms() * 4, 256, 0, stream>>>(w, static_cast<uint16_t*>(out), Cout, Cin, taps, BN,
                                                         transposed, (!transposed && Cout == 64) ? 1 : 0, 0);
  NC_CUDA(cudaGetLastError());
  return 0;
}

// Packed bf16 image of the data-gradient filter of a k3 conv layer with weight w = OIDHW (Cout, Cin, 3, 3, 3);
// same size as the forward image (packed_weight_bytes(Cout, Cin, 27, 0)).
int pack_weights_dgrad(const float* w, void* out, int Cout, int Cin, cudaStream_t stream) {
  if (Cin % 64 || Cout % 64) return set_error("pack_weights_dgrad: Cin and Cout must be multiples of 64");
  const int BN = conv3d_k3_bn(Cin);  // GEMM N = the layer's Cin, GEMM K = the layer's Cout
  if (Cin % BN) return set_error("pack_weights_dgrad: Cin must be 64 or a multiple of 128");
  pack_weights_kernel<<<num_sms() * 4, 256, 0, stream>>>(w, static_cast<uint16_t*>(out), Cin, Cout, 27, BN, 0,
                                                         Cin == 64 ? 1 : 0, 1);
  NC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nc

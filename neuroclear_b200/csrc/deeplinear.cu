// Memory-bound kernels of DeepLinearGenerator (reference models/networks.py:893-917: Conv3d k7 1->64, k5 64->64,
// k3 64->64, 1x1 64->32->16->1, no bias, no activation), forward and backward.
//
// Mapping of the network onto this library (neuroclear_b200/deeplinear_engine.py drives it):
//   * k7, Cin = 1: an in-plane im2col writes the 7 x 7 (kh,kw) neighbourhood of every voxel as 49 (+15 zero)
//     channels; the layer is then a 7-tap depth conv 64 -> 64 on the tensor-core kernel (conv3d_tc_64, ksd 7 / ksp 1).
//     Backward: the same GEMMs (wgrad ks = 71, data gradient with the flipped filter) + col2im below.
//   * k5: conv3d_tc_64 (ksd = ksp = 5), nc_conv3d_wgrad(ks = 5).
//   * k3 and the three 1x1 layers are linear and only pointwise ops follow the k3 layer, so they fold EXACTLY
//     (borders included) into one 64 -> 1 k3 stencil with weights K[ci][tap] = sum_co (W6 W5 W4)[co] W3[co][ci][tap]:
//     stencil_fwd / stencil_bwd_data / the first-layer weight-gradient kernel with the roles of x and dy swapped.
//     The fold and its chain rule back to W3..W6 are a few thousand flops on the host side.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "internal.h"

namespace nc {

namespace {

__device__ __forceinline__ uint16_t to_bits(float v, int fmt) {
  if (fmt) {
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    return *reinterpret_cast<const uint16_t*>(&b);
  }
  const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  return *reinterpret_cast<const uint16_t*>(&h);
}

// im2col49 / col2im49 work on in-plane tiles of 8 (h) x 32 (w) voxels whose 3-voxel halo is staged in shared memory.
// (The round-1 kernels went to L1 / L2 for every element — 49 two-byte gathers from 49 different 128-byte rows per
// voxel in col2im, a full index decode per 16 bytes written in im2col — and ran at 0.22 / 0.10 ms per 108^3 crop; the
// HBM bound of the 128 bytes per voxel they move is 0.03 ms.)
namespace ic {
constexpr int TY = 8, TX = 32, HY = TY + 6, HX = TX + 6, ROWS = HY * HX;  // 532 halo voxels
constexpr int PITCH = 33;  // words per staged row in col2im (28 used = 56 channels): odd -> conflict-free gathers
constexpr int COL2IM_SMEM = ROWS * PITCH * 4;
}  // namespace ic

// out[v][kh * 7 + kw] = x[d][h + kh - 3][w + kw - 3] (zero outside the plane), channels 49..63 = 0
// thread = one 8-channel group (16 bytes) of 8 voxels of the tile; 8 consecutive lanes write one voxel's 128 bytes
__global__ void __launch_bounds__(256)
im2col49_kernel(const float* __restrict__ x, int D, int H, int W, int tiles_y, int tiles_x, int fmt,
                uint16_t* __restrict__ out) {
  using namespace ic;
  __shared__ float xs[ROWS];
  const int nb = blockIdx.y;
  const int tx = blockIdx.x % tiles_x;
  const int r = blockIdx.x / tiles_x;
  const int d = r / tiles_y, h0 = (r % tiles_y) * TY, w0 = tx * TX;
  const float* xp = x + (static_cast<size_t>(nb) * D + d) * H * W;
  for (int i = threadIdx.x; i < ROWS; i += 256) {
    const int zh = h0 + i / HX - 3, zw = w0 + i % HX - 3;
    xs[i] = (zh >= 0 && zh < H && zw >= 0 && zw < W) ? __ldg(xp + static_cast<size_t>(zh) * W + zw) : 0.f;
  }
  __syncthreads();
  const int g = threadIdx.x & 7;
  int off[8];  // offset of channel g * 8 + i inside the staged tile relative to the voxel, -1 for the 15 zero channels
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = g * 8 + i;
    off[i] = c < 49 ? (c / 7) * HX + c % 7 : -1;
  }
  uint16_t* op = out + ((static_cast<size_t>(nb) * D + d) * H * W) * 64 + g * 8;
#pragma unroll 2
  for (int j = 0; j < 8; ++j) {
    const int v = (threadIdx.x >> 3) + 32 * j;  // voxel of the tile: ly = j, lx = threadIdx.x / 8
    const int ly = v / TX, lx = v % TX;
    if (h0 + ly >= H || w0 + lx >= W) continue;
    const float* base = xs + ly * HX + lx;
    uint16_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = to_bits(off[i] >= 0 ? base[off[i]] : 0.f, fmt);
    *reinterpret_cast<uint4*>(op + (static_cast<size_t>(h0 + ly) * W + w0 + lx) * 64) =
        *reinterpret_cast<const uint4*>(o);
  }
}

// dx[d][h][w] = sum_{kh,kw} g[d][h - kh + 3][w - kw + 3][kh * 7 + kw]   (adjoint of im2col49; g bf16)
// The 56 first channels of the tile's halo rows are staged as 28 words per row; thread = one voxel, 49 conflict-free
// two-byte gathers, summed kh-major like the plain loop (bitwise the same result).
__global__ void __launch_bounds__(256)
col2im49_kernel(const __nv_bfloat16* __restrict__ g, int D, int H, int W, int tiles_y, int tiles_x,
                float* __restrict__ dx) {
  using namespace ic;
  extern __shared__ uint32_t gs[];  // [ROWS][PITCH]
  const int nb = blockIdx.y;
  const int tx = blockIdx.x % tiles_x;
  const int r = blockIdx.x / tiles_x;
  const int d = r / tiles_y, h0 = (r % tiles_y) * TY, w0 = tx * TX;
  const uint4* gp = reinterpret_cast<const uint4*>(g) + ((static_cast<size_t>(nb) * D + d) * H * W) * 8;
  // 7 x 16 bytes per halo row, several loads in flight per thread (with one 4-byte load per iteration the staging
  // loop was a chain of 58 L2 round trips per thread and the kernel no faster than the gather it replaced)
#pragma unroll 5
  for (int i = threadIdx.x; i < ROWS * 7; i += 256) {
    const int row = i / 7, part = i - row * 7;
    const int zh = h0 + row / HX - 3, zw = w0 + row % HX - 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (zh >= 0 && zh < H && zw >= 0 && zw < W) v = __ldg(gp + (static_cast<size_t>(zh) * W + zw) * 8 + part);
    uint32_t* dst = gs + row * PITCH + part * 4;
    dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;
  }
  __syncthreads();
  const int ly = threadIdx.x / TX, lx = threadIdx.x % TX;
  if (h0 + ly >= H || w0 + lx >= W) return;
  float s = 0.f;
#pragma unroll
  for (int kh = 0; kh < 7; ++kh) {
#pragma unroll
    for (int kw = 0; kw < 7; ++kw) {
      const int c = kh * 7 + kw;
      const uint32_t word = gs[((ly - kh + 6) * HX + lx - kw + 6) * PITCH + (c >> 1)];
      // bf16 -> fp32 is a 16-bit shift; rows outside the plane were staged as zeros (adding +0 changes nothing)
      s += __uint_as_float((c & 1) ? (word & 0xFFFF0000u) : (word << 16));
    }
  }
  dx[(static_cast<size_t>(nb) * D + d) * H * W + static_cast<size_t>(h0 + ly) * W + w0 + lx] = s;
}

// Both stencils below work on tiles of 4 x 8 x 8 voxels staged (with their one-voxel halo) in shared memory; a thread
// owns 8 channels and the 8 outputs of one w-row, so every staged vector / scalar it loads feeds up to 3 x 8 FMAs.
// (The round-1 kernels read every neighbour straight from L1 / L2 — 27 x 128 B per voxel, 4.3 GB per 108^3 crop — and
// ran at 0.64 / 0.52 ms; the FMA bound of 2.2 GMAC is 65 us.)
namespace st {
constexpr int TZ = 4, TY = 8, TX = 8;
constexpr int HZ = TZ + 2, HY = TY + 2, HX = TX + 2, HALO = HZ * HY * HX;  // 600
constexpr int FWD_SMEM = HALO * 128 + 27 * 64 * 4;                        // fp16 halo vectors + K
constexpr int BWD_SMEM = HALO * 4 + 27 * 64 * 4;
}  // namespace st

// out[v] = sum_{tap, ci} h[v + tap - 1][ci] * K[ci][tap]: 64 -> 1 k3 stencil.
__global__ void __launch_bounds__(256)
stencil64to1_kernel(const __half* __restrict__ hin, const float* __restrict__ K, int D, int H, int W, int tiles_y,
                    int tiles_x, int tiles, float* __restrict__ out) {
  using namespace st;
  extern __shared__ __align__(16) uint8_t st_smem[];
  uint4* hs = reinterpret_cast<uint4*>(st_smem);                        // [HALO][8 x 16 B]
  float* Ks = reinterpret_cast<float*>(st_smem + HALO * 128);           // [tap][ci]
  for (int i = threadIdx.x; i < 27 * 64; i += 256) Ks[(i % 27) * 64 + i / 27] = __ldg(K + i);  // K is [ci][tap]
  const int nb = blockIdx.y;
  const int sub = threadIdx.x & 7, row = threadIdx.x >> 3;              // row = lz * TY + ly
  const int lz = row / TY, ly = row % TY;
  const size_t voxels = static_cast<size_t>(D) * H * W;
  const __half* hc = hin + static_cast<size_t>(nb) * voxels * 64;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x;
    const int r = tile / tiles_x;
    const int z0 = (r / tiles_y) * TZ, y0 = (r % tiles_y) * TY, x0 = tx * TX;
    __syncthreads();
    for (int i = threadIdx.x; i < HALO * 8; i += 256) {
      const int hv = i >> 3, part = i & 7;
      const int hx = hv % HX, hy = (hv / HX) % HY, hz = hv / (HX * HY);
      const int gz = z0 + hz - 1, gy = y0 + hy - 1, gx = x0 + hx - 1;
      uint4 g = make_uint4(0u, 0u, 0u, 0u);                             // zero padding
      if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W)
        g = __ldg(reinterpret_cast<const uint4*>(hc + ((static_cast<size_t>(gz) * H + gy) * W + gx) * 64) + part);
      hs[i] = g;
    }
    __syncthreads();
    float acc[TX];
#pragma unroll
    for (int o = 0; o < TX; ++o) acc[o] = 0.f;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        float kk[3][8];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4* kp = reinterpret_cast<const float4*>(Ks + ((kd * 3 + kh) * 3 + kw) * 64 + sub * 8);
          const float4 a = kp[0], b = kp[1];
          kk[kw][0] = a.x, kk[kw][1] = a.y, kk[kw][2] = a.z, kk[kw][3] = a.w;
          kk[kw][4] = b.x, kk[kw][5] = b.y, kk[kw][6] = b.z, kk[kw][7] = b.w;
        }
        const uint4* rp = hs + (((lz + kd) * HY + ly + kh) * HX) * 8 + sub;
#pragma unroll
        for (int j = 0; j < HX; ++j) {                                  // staged voxel j of the row feeds outputs j - kw
          const uint4 raw = rp[j * 8];
          const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
          float f[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 t = __half22float2(h2[i]);
            f[2 * i] = t.x, f[2 * i + 1] = t.y;
          }
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int o = j - kw;
            if (o >= 0 && o < TX) {
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[o] = fmaf(f[i], kk[kw][i], acc[o]);
            }
          }
        }
      }
    const int gz = z0 + lz, gy = y0 + ly;
#pragma unroll
    for (int o = 0; o < TX; ++o) {
      float a = acc[o];
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      const int gx = x0 + o;
      if (sub == (o & 7) && gz < D && gy < H && gx < W)
        out[static_cast<size_t>(nb) * voxels + (static_cast<size_t>(gz) * H + gy) * W + gx] = a;
    }
  }
}

// dh[v][ci] = sum_tap dout[v - tap + 1] * K[ci][tap]   (adjoint of the stencil w.r.t. its input), bf16 out
__global__ void __launch_bounds__(256)
stencil1to64_kernel(const float* __restrict__ dout, const float* __restrict__ K, int D, int H, int W, int tiles_y,
                    int tiles_x, int tiles, __nv_bfloat16* __restrict__ dh) {
  using namespace st;
  extern __shared__ __align__(16) uint8_t st_smem[];
  float* ds = reinterpret_cast<float*>(st_smem);                        // [HALO]
  float* Ks = ds + HALO;                                                // [tap][ci]
  for (int i = threadIdx.x; i < 27 * 64; i += 256) Ks[(i % 27) * 64 + i / 27] = __ldg(K + i);
  const int nb = blockIdx.y;
  const int sub = threadIdx.x & 7, row = threadIdx.x >> 3;
  const int lz = row / TY, ly = row % TY;
  const size_t voxels = static_cast<size_t>(D) * H * W;
  const float* dc = dout + static_cast<size_t>(nb) * voxels;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x;
    const int r = tile / tiles_x;
    const int z0 = (r / tiles_y) * TZ, y0 = (r % tiles_y) * TY, x0 = tx * TX;
    __syncthreads();
    for (int i = threadIdx.x; i < HALO; i += 256) {
      const int hx = i % HX, hy = (i / HX) % HY, hz = i / (HX * HY);
      const int gz = z0 + hz - 1, gy = y0 + hy - 1, gx = x0 + hx - 1;
      const bool in = gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W;
      ds[i] = in ? __ldg(dc + (static_cast<size_t>(gz) * H + gy) * W + gx) : 0.f;
    }
    __syncthreads();
    float acc[TX][8];
#pragma unroll
    for (int o = 0; o < TX; ++o)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[o][i] = 0.f;
    // output voxel (lz, ly, o) reads dout at halo position (lz + 2 - kd, ly + 2 - kh, o + 2 - kw)
#pragma unroll
    for (int kd = 0; kd < 3; ++kd)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        float kk[3][8];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4* kp = reinterpret_cast<const float4*>(Ks + ((kd * 3 + kh) * 3 + kw) * 64 + sub * 8);
          const float4 a = kp[0], b = kp[1];
          kk[kw][0] = a.x, kk[kw][1] = a.y, kk[kw][2] = a.z, kk[kw][3] = a.w;
          kk[kw][4] = b.x, kk[kw][5] = b.y, kk[kw][6] = b.z, kk[kw][7] = b.w;
        }
        const float* rp = ds + ((lz + 2 - kd) * HY + ly + 2 - kh) * HX;
#pragma unroll
        for (int j = 0; j < HX; ++j) {                                  // staged value j feeds outputs o = j + kw - 2
          const float g = rp[j];
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int o = j + kw - 2;
            if (o >= 0 && o < TX) {
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[o][i] = fmaf(g, kk[kw][i], acc[o][i]);
            }
          }
        }
      }
    const int gz = z0 + lz, gy = y0 + ly;
    if (gz < D && gy < H) {
#pragma unroll
      for (int o = 0; o < TX; ++o) {
        const int gx = x0 + o;
        if (gx >= W) break;
        uint4 pk;
        __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int i = 0; i < 4; ++i) p2[i] = __floats2bfloat162_rn(acc[o][2 * i], acc[o][2 * i + 1]);
        *reinterpret_cast<uint4*>(dh + (static_cast<size_t>(nb) * voxels + (static_cast<size_t>(gz) * H + gy) * W + gx) * 64 +
                                  sub * 8) = pk;
      }
    }
  }
}

}  // namespace

int im2col49(const float* x, int NB, int D, int H, int W, int fmt, void* out, cudaStream_t stream) {
  const int ty = (H + ic::TY - 1) / ic::TY, tx = (W + ic::TX - 1) / ic::TX;
  if (static_cast<long long>(D) * ty * tx >= (1ll << 31) || NB > 65535) return set_error("im2col49: volume too large");
  im2col49_kernel<<<dim3(D * ty * tx, NB), 256, 0, stream>>>(x, D, H, W, ty, tx, fmt, static_cast<uint16_t*>(out));
  NC_CUDA(cudaGetLastError());
  return 0;
}
int col2im49(const void* g, int NB, int D, int H, int W, float* dx, cudaStream_t stream) {
  const int ty = (H + ic::TY - 1) / ic::TY, tx = (W + ic::TX - 1) / ic::TX;
  if (static_cast<long long>(D) * ty * tx >= (1ll << 31) || NB > 65535) return set_error("col2im49: volume too large");
  static bool attr[64] = {false};
  if (first_use_on_device(attr))
    NC_CUDA(cudaFuncSetAttribute(col2im49_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ic::COL2IM_SMEM));
  col2im49_kernel<<<dim3(D * ty * tx, NB), 256, ic::COL2IM_SMEM, stream>>>(static_cast<const __nv_bfloat16*>(g), D, H,
                                                                            W, ty, tx, dx);
  NC_CUDA(cudaGetLastError());
  return 0;
}
int stencil64to1_fwd(const void* h, const float* K, int NB, int D, int H, int W, float* out, cudaStream_t stream) {
  if (static_cast<long long>(D) * H * W >= (1ll << 31)) return set_error("stencil64to1_fwd: volume too large");
  const int tz = (D + st::TZ - 1) / st::TZ, ty = (H + st::TY - 1) / st::TY, tx = (W + st::TX - 1) / st::TX;
  static bool attr[64] = {false};
  if (first_use_on_device(attr))
    NC_CUDA(cudaFuncSetAttribute(stencil64to1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, st::FWD_SMEM));
  const int tiles = tz * ty * tx, blocks = tiles < num_sms() * 2 ? tiles : num_sms() * 2;
  stencil64to1_kernel<<<dim3(blocks, NB), 256, st::FWD_SMEM, stream>>>(static_cast<const __half*>(h), K, D, H, W, ty, tx,
                                                                       tiles, out);
  NC_CUDA(cudaGetLastError());
  return 0;
}
int stencil64to1_bwd_data(const float* dout, const float* K, int NB, int D, int H, int W, void* dh,
                          cudaStream_t stream) {
  if (static_cast<long long>(D) * H * W * 8 >= (1ll << 32)) return set_error("stencil64to1_bwd_data: volume too large");
  const int tz = (D + st::TZ - 1) / st::TZ, ty = (H + st::TY - 1) / st::TY, tx = (W + st::TX - 1) / st::TX;
  const int tiles = tz * ty * tx, blocks = tiles < num_sms() * 4 ? tiles : num_sms() * 4;
  stencil1to64_kernel<<<dim3(blocks, NB), 256, st::BWD_SMEM, stream>>>(dout, K, D, H, W, ty, tx, tiles,
                                                                       static_cast<__nv_bfloat16*>(dh));
  NC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nc

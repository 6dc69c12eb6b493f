// Memory-bound backward kernels of Unet_deconv (reference models/networks.py:413-538 under autograd, as driven by
// axial_to_lateral_gan_apollo_model.py:255-295): InstanceNorm3d + ReLU backward (with the max-pool / skip-concat
// gradient routing folded in), the 1x1x1 + 1x1x1 + sigmoid head backward, the weight gradient of the Cin = 1 first
// layer, the space-to-depth gather that turns the transposed conv's backward into plain GEMMs, and small helpers.
//
// Layout: activations / raw conv outputs fp16 NDHWC (as the forward pass left them), gradients bf16 NDHWC, all
// reductions fp32 per block -> fp64 in a fixed order (bitwise repeatable).  One thread owns 8 channels of a voxel
// (one 16-byte access).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "internal.h"

namespace nc {

namespace {

__device__ __forceinline__ void load8_f16(const __half* p, float (&o)[8]) {
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h2[i]);
    o[2 * i] = f.x;
    o[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&o)[8]) {
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h2[i]);
    o[2 * i] = f.x;
    o[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 pk;
  __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
  for (int i = 0; i < 4; ++i) h2[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = pk;
}

// Where the gradient w.r.t. the post-ReLU activation comes from.
//   mode 0: a bf16 tensor (row pitch a_ld, channel offset a_coff): a data-gradient output or a concat-gradient slice
//   mode 1: the head: dA[v][c] = du[v] * w[c]                                 (one_by_one, networks.py:507)
//   mode 2: skip + pool: dA[v][c] = skip[v][c] + (v is the arg-max of its 2x2x2 window ? dpool[window][c] : 0)
//           (torch.cat skip at networks.py:526/531 + MaxPool3d(2) at :491/494; first maximum in d,h,w order wins)
struct GradSrc {
  const __nv_bfloat16* a;
  int a_ld, a_coff;
  const float* du;
  const float* w;
  const __nv_bfloat16* dpool;
};

struct InBwdArgs {
  const __half* raw;     // (NB, D, H, W, C) raw conv output
  const float* mr;       // (NB, 2, C) mean, rstd
  const float* m12;      // (NB, 2, C) mean of g, mean of g * yhat     (apply only)
  int D, H, W, C;
  GradSrc src;
  float* partial;        // reduce: (NB, gridDim.x, 2, C)
  __nv_bfloat16* d_raw;  // apply: (NB, D, H, W, C)
};

// Calls f(voxel index inside the cube, yhat[8], dA[8]) for the voxels of one work unit (a voxel, or the 8 voxels
// of a pooling window in mode 2).
template <int MODE, class F>
__device__ __forceinline__ void for_unit(const InBwdArgs& a, int nb, unsigned unit, int cg, const float (&mu)[8],
                                         const float (&rs)[8], F&& f) {
  const int C = a.C;
  const size_t cube = static_cast<size_t>(a.D) * a.H * a.W;
  if constexpr (MODE != 2) {
    const size_t gv = static_cast<size_t>(nb) * cube + unit;
    float y[8], dA[8];
    load8_f16(a.raw + gv * C + cg * 8, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = (y[i] - mu[i]) * rs[i];
    if constexpr (MODE == 0) {
      load8_bf16(a.src.a + gv * a.src.a_ld + a.src.a_coff + cg * 8, dA);
    } else {
      const float du = __ldg(a.src.du + gv);
#pragma unroll
      for (int i = 0; i < 8; ++i) dA[i] = du * __ldg(a.src.w + cg * 8 + i);
    }
    f(gv, y, dA);
  } else {
    const int PH = a.H / 2, PW = a.W / 2;
    const int pw = unit % PW;
    unsigned r = unit / PW;
    const int ph = r % PH;
    const int pd = r / PH;
    float y[8][8], best[8], dp[8];
    int arg[8];
    load8_bf16(a.src.dpool + (static_cast<size_t>(nb) * (cube / 8) + unit) * C + cg * 8, dp);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const size_t gv =
          ((static_cast<size_t>(nb) * a.D + 2 * pd + (k >> 2)) * a.H + 2 * ph + ((k >> 1) & 1)) * a.W + 2 * pw + (k & 1);
      load8_f16(a.raw + gv * C + cg * 8, y[k]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        y[k][i] = (y[k][i] - mu[i]) * rs[i];
        const float o = fmaxf(y[k][i], 0.f);
        if (k == 0 || o > best[i]) {
          best[i] = o;
          arg[i] = k;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const size_t gv =
          ((static_cast<size_t>(nb) * a.D + 2 * pd + (k >> 2)) * a.H + 2 * ph + ((k >> 1) & 1)) * a.W + 2 * pw + (k & 1);
      float dA[8];
      load8_bf16(a.src.a + gv * a.src.a_ld + a.src.a_coff + cg * 8, dA);
#pragma unroll
      for (int i = 0; i < 8; ++i) dA[i] += (arg[i] == k) ? dp[i] : 0.f;
      f(gv, y[k], dA);
    }
  }
}

// InstanceNorm + ReLU backward, pass 1: per-channel sums of g = dA * [yhat > 0] and g * yhat.
template <int MODE>
__global__ void __launch_bounds__(256)
in_bwd_reduce_kernel(const InBwdArgs a) {
  __shared__ float red[256][17];
  const int nb = blockIdx.y;
  const int cg_per = a.C / 8, lanes = 256 / cg_per;
  const int cg = threadIdx.x % cg_per, lane = threadIdx.x / cg_per;
  const unsigned units = static_cast<unsigned>(a.D) * a.H * a.W / (MODE == 2 ? 8 : 1);
  float mu[8], rs[8], s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = __ldg(a.mr + static_cast<size_t>(nb) * 2 * a.C + cg * 8 + i);
    rs[i] = __ldg(a.mr + static_cast<size_t>(nb) * 2 * a.C + a.C + cg * 8 + i);
    s1[i] = s2[i] = 0.f;
  }
  for (unsigned u = blockIdx.x * lanes + lane; u < units; u += gridDim.x * lanes) {
    for_unit<MODE>(a, nb, u, cg, mu, rs, [&](size_t, const float(&y)[8], const float(&dA)[8]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float g = y[i] > 0.f ? dA[i] : 0.f;
        s1[i] += g;
        s2[i] = fmaf(g, y[i], s2[i]);
      }
    });
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x][i] = s1[i];
    red[threadIdx.x][8 + i] = s2[i];
  }
  __syncthreads();
  // thread t < 2C: which = t / C, channel c = t % C; add the `lanes` voxel lanes in order
  for (int t = threadIdx.x; t < 2 * a.C; t += 256) {
    const int which = t / a.C, c = t - which * a.C;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[l * cg_per + (c >> 3)][which * 8 + (c & 7)];
    a.partial[((static_cast<size_t>(nb) * gridDim.x + blockIdx.x) * 2 + which) * a.C + c] = s;
  }
}

// pass 2: d_raw = rstd * (g - mean(g) - yhat * mean(g * yhat))
template <int MODE>
__global__ void __launch_bounds__(256)
in_bwd_apply_kernel(const InBwdArgs a) {
  const int nb = blockIdx.y;
  const int cg_per = a.C / 8, lanes = 256 / cg_per;
  const int cg = threadIdx.x % cg_per, lane = threadIdx.x / cg_per;
  const unsigned units = static_cast<unsigned>(a.D) * a.H * a.W / (MODE == 2 ? 8 : 1);
  float mu[8], rs[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = __ldg(a.mr + static_cast<size_t>(nb) * 2 * a.C + cg * 8 + i);
    rs[i] = __ldg(a.mr + static_cast<size_t>(nb) * 2 * a.C + a.C + cg * 8 + i);
    m1[i] = __ldg(a.m12 + static_cast<size_t>(nb) * 2 * a.C + cg * 8 + i);
    m2[i] = __ldg(a.m12 + static_cast<size_t>(nb) * 2 * a.C + a.C + cg * 8 + i);
  }
  for (unsigned u = blockIdx.x * lanes + lane; u < units; u += gridDim.x * lanes) {
    for_unit<MODE>(a, nb, u, cg, mu, rs, [&](size_t gv, const float(&y)[8], const float(&dA)[8]) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float g = y[i] > 0.f ? dA[i] : 0.f;
        o[i] = rs[i] * (g - m1[i] - y[i] * m2[i]);
      }
      store8_bf16(a.d_raw + gv * a.C + cg * 8, o);
    });
  }
}

// out[nb][c] = scale * sum over rows of partial[nb][row][c]; fp64, fixed order (32 row lanes, then lanes in order).
// This kernel sits between the two passes of every InstanceNorm backward (15 launches per training iteration): with 8
// row lanes and a rolled loop it was a chain of ~75 dependent L2 round trips (33 us); 32 lanes and eight loads in
// flight per lane make it a few microseconds.
__global__ void __launch_bounds__(1024)
colsum_finalize_kernel(const float* __restrict__ partial, int rows, int C, double scale, float* __restrict__ out) {
  __shared__ double red[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx, nb = blockIdx.y;
  double s = 0.0;
  if (c < C) {
    const float* p = partial + static_cast<size_t>(nb) * rows * C + c;
    int r = ty;
    for (; r + 7 * 32 < rows; r += 8 * 32) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(p + static_cast<size_t>(r + 32 * u) * C);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += static_cast<double>(v[u]);
    }
    for (; r < rows; r += 32) s += static_cast<double>(__ldg(p + static_cast<size_t>(r) * C));
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][tx];
    out[static_cast<size_t>(nb) * C + c] = static_cast<float>(t * scale);
  }
}

// ------------------------------------------------------------------------------------------------ head backward
// out = sigmoid(w2 * (w1 . a + b1) + b2), a = relu(IN(raw)); given dout: du[v] = dout * out * (1 - out) * w2 and the
// per-block partial sums of [d w1 (64) | d b1 | d w2 | d b2].  8 lanes per voxel, 8 channels per lane.
constexpr int HEAD_COLS = 68;  // 64 + 3, padded
__global__ void __launch_bounds__(256)
head_bwd_kernel(const __half* __restrict__ raw, const float* __restrict__ mr, const float* __restrict__ hp,
                const float* __restrict__ dout, unsigned voxels, float* __restrict__ du_out,
                float* __restrict__ partial) {
  constexpr int C = 64;
  __shared__ float red[256][9];
  __shared__ float red3[32][3];
  const int nb = blockIdx.y;
  const int sub = threadIdx.x & 7, lane = threadIdx.x >> 3;  // 32 voxel lanes
  float mu[8], rs[8], w1[8], sw[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = __ldg(mr + static_cast<size_t>(nb) * 2 * C + sub * 8 + i);
    rs[i] = __ldg(mr + static_cast<size_t>(nb) * 2 * C + C + sub * 8 + i);
    w1[i] = __ldg(hp + sub * 8 + i);
    sw[i] = 0.f;
  }
  const float b1 = __ldg(hp + C), w2 = __ldg(hp + C + 1), b2 = __ldg(hp + C + 2);
  float sb1 = 0.f, sw2 = 0.f, sb2 = 0.f;
  // every 8-lane group stays converged: the loop bound is rounded up and out-of-range voxels contribute zero
  const unsigned per_iter = gridDim.x * 32u;
  for (unsigned v0 = blockIdx.x * 32u; v0 < voxels; v0 += per_iter) {
    const unsigned v = v0 + lane;
    const bool ok = v < voxels;
    const size_t gv = static_cast<size_t>(nb) * voxels + (ok ? v : 0u);
    float act[8];
    load8_f16(raw + gv * C + sub * 8, act);
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      act[i] = fmaxf((act[i] - mu[i]) * rs[i], 0.f);
      t = fmaf(act[i], w1[i], t);
    }
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    const float u1 = t + b1;
    const float o = 1.0f / (1.0f + expf(-fmaf(w2, u1, b2)));
    const float dz = ok ? __ldg(dout + gv) * o * (1.0f - o) : 0.f;
    const float du = dz * w2;
#pragma unroll
    for (int i = 0; i < 8; ++i) sw[i] = fmaf(du, act[i], sw[i]);
    if (sub == 0) {
      sb1 += du;
      sw2 = fmaf(dz, u1, sw2);
      sb2 += dz;
      if (ok) du_out[gv] = du;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = sw[i];
  if (sub == 0) {
    red3[lane][0] = sb1;
    red3[lane][1] = sw2;
    red3[lane][2] = sb2;
  }
  __syncthreads();
  float* dst = partial + (static_cast<size_t>(nb) * gridDim.x + blockIdx.x) * HEAD_COLS;
  if (threadIdx.x < 64) {
    const int c = threadIdx.x;
    float s = 0.f;
    for (int l = 0; l < 32; ++l) s += red[l * 8 + (c >> 3)][c & 7];
    dst[c] = s;
  } else if (threadIdx.x < 68) {
    const int k = threadIdx.x - 64;
    float s = 0.f;
    if (k < 3)
      for (int l = 0; l < 32; ++l) s += red3[l][k];
    dst[threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------ first layer wgrad
// dW[co][tap] = sum_v dy[v][co] * x[v + tap - 1] for the Cin = 1 k3 conv (networks.py:420; also dK of the folded
// DeepLinear tail).  2.2 GMAC per 108^3 crop on the CUDA cores: the job is to keep the FMA pipe fed.
// A block owns tiles of 4 x 8 x 16 voxels: the fp32 halo of x (6 x 10 x 18) and the 16-bit dy tile (64 KB) are staged
// in shared memory; thread = (8 output channels) x (the 9 in-plane taps of ONE kd) x (one of 8 voxel lanes): per voxel
// one 16-byte dy load + 9 broadcast x loads feed 72 FMAs (the round-1 kernel issued 27 L1 loads per 108 FMAs from
// every one of the 16 threads that shared a voxel and ran at 155 clk per voxel per SM; the FMA bound is 13.5).
// Voxel lanes are combined in lane order, blocks in block order (colsum_finalize_kernel): bitwise repeatable.
namespace w1 {
constexpr int TZ = 4, TY = 8, TX = 16, VOX = TZ * TY * TX;
constexpr int HZ = TZ + 2, HY = TY + 2, HX = TX + 2, HALO = HZ * HY * HX;
constexpr int THREADS = 192;  // 8 channel groups x 8 voxel lanes x 3 kd
constexpr int SMEM_BYTES = VOX * 8 * 16 + HALO * 4;
}  // namespace w1

template <bool F16>
__global__ void __launch_bounds__(w1::THREADS)
conv1_wgrad_kernel(const float* __restrict__ x, const uint16_t* __restrict__ dy, int D, int H, int W, int tiles_y,
                   int tiles_x, int tiles, float* __restrict__ partial) {
  using namespace w1;
  extern __shared__ __align__(16) uint8_t w1_smem[];
  uint4* dys = reinterpret_cast<uint4*>(w1_smem);                      // [VOX][8]: 64 channels of one voxel
  float* xs = reinterpret_cast<float*>(w1_smem + VOX * 8 * 16);         // [HZ][HY][HX]
  const int nb = blockIdx.y;
  const int cg = threadIdx.x & 7, vl = (threadIdx.x >> 3) & 7, kd = threadIdx.x >> 6;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
  const size_t voxels = static_cast<size_t>(D) * H * W;
  const float* xc = x + static_cast<size_t>(nb) * voxels;
  const uint16_t* dyc = dy + static_cast<size_t>(nb) * voxels * 64;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x;
    const int r = tile / tiles_x;
    const int ty = r % tiles_y, tz = r / tiles_y;
    const int z0 = tz * TZ, y0 = ty * TY, x0 = tx * TX;
    __syncthreads();  // the previous tile has been consumed
    for (int i = threadIdx.x; i < HALO; i += THREADS) {
      const int hx = i % HX, hy = (i / HX) % HY, hz = i / (HX * HY);
      const int gz = z0 + hz - 1, gy = y0 + hy - 1, gx = x0 + hx - 1;
      const bool in = gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W;
      xs[i] = in ? __ldg(xc + (static_cast<size_t>(gz) * H + gy) * W + gx) : 0.f;
    }
    for (int i = threadIdx.x; i < VOX * 8; i += THREADS) {
      const int v = i >> 3, part = i & 7;
      const int lx = v % TX, ly = (v / TX) % TY, lz = v / (TX * TY);
      const int gz = z0 + lz, gy = y0 + ly, gx = x0 + lx;
      uint4 g = make_uint4(0u, 0u, 0u, 0u);  // voxels outside the volume contribute nothing
      if (gz < D && gy < H && gx < W)
        g = __ldg(reinterpret_cast<const uint4*>(dyc + ((static_cast<size_t>(gz) * H + gy) * W + gx) * 64) + part);
      dys[i] = g;
    }
    __syncthreads();
#pragma unroll 2
    for (int v = vl; v < VOX; v += 8) {
      const int lx = v % TX, ly = (v / TX) % TY, lz = v / (TX * TY);
      const uint4 raw = dys[v * 8 + cg];
      float gf[8];
      if constexpr (F16) {
        const __half2* g2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(g2[i]);
          gf[2 * i] = f.x, gf[2 * i + 1] = f.y;
        }
      } else {
        const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(g2[i]);
          gf[2 * i] = f.x, gf[2 * i + 1] = f.y;
        }
      }
      const float* xb = xs + ((lz + kd) * HY + ly) * HX + lx;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float xv = xb[kh * HX + kw];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[kh * 3 + kw][i] = fmaf(gf[i], xv, acc[kh * 3 + kw][i]);
        }
    }
  }
  // voxel lanes -> block, in lane order; red[co][tap] is the weight's own layout
  __syncthreads();
  float* red = reinterpret_cast<float*>(w1_smem);
  for (int l = 0; l < 8; ++l) {
    if (vl == l) {
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = (cg * 8 + i) * 27 + kd * 9 + t;
          red[idx] = (l == 0 ? 0.f : red[idx]) + acc[t][i];
        }
    }
    __syncthreads();
  }
  float* dst = partial + (static_cast<size_t>(nb) * gridDim.x + blockIdx.x) * (27 * 64);
  for (int i = threadIdx.x; i < 27 * 64; i += THREADS) dst[i] = red[i];
}

// ------------------------------------------------------------------------------------------------ helpers
// space-to-depth of a fine-grid bf16 tensor slice: out[coarse voxel][tap * C + c] = src[2 * coarse + tap][coff + c]
__global__ void __launch_bounds__(256)
s2d_kernel(const __nv_bfloat16* __restrict__ src, int ld, int coff, int D, int H, int W, int C,
           __nv_bfloat16* __restrict__ out) {
  const int nb = blockIdx.y;
  const int cg_per = C / 8;
  const unsigned total = static_cast<unsigned>(D) * H * W * 8 * cg_per;  // D, H, W: coarse grid
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const int cg = idx % cg_per;
    unsigned r = idx / cg_per;
    const int tap = r & 7;
    r >>= 3;
    const int w = r % W;
    r /= W;
    const int h = r % H, d = r / H;
    const size_t fv = ((static_cast<size_t>(nb) * 2 * D + 2 * d + (tap >> 2)) * 2 * H + 2 * h + ((tap >> 1) & 1)) * 2 * W +
                      2 * w + (tap & 1);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + fv * ld + coff + cg * 8));
    const size_t cv = ((static_cast<size_t>(nb) * D + d) * H + h) * W + w;
    *reinterpret_cast<uint4*>(out + (cv * 8 + tap) * C + cg * 8) = v;
  }
}

// per-block column sums of a bf16 tensor slice (bias gradients)
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ src, int ld, int coff, unsigned rows, int C,
                   float* __restrict__ partial) {
  __shared__ float red[256][9];
  const int nb = blockIdx.y;
  const int cg_per = C / 8, lanes = 256 / cg_per;
  const int cg = threadIdx.x % cg_per, lane = threadIdx.x / cg_per;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  for (unsigned r = blockIdx.x * lanes + lane; r < rows; r += gridDim.x * lanes) {
    float v[8];
    load8_bf16(src + (static_cast<size_t>(nb) * rows + r) * ld + coff + cg * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += v[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = s[i];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[l * cg_per + (c >> 3)][c & 7];
    partial[(static_cast<size_t>(nb) * gridDim.x + blockIdx.x) * C + c] = t;
  }
}

// fp16 -> bf16 copy of a channel slice
__global__ void __launch_bounds__(256)
cast_f16_bf16_kernel(const __half* __restrict__ src, int src_ld, int src_coff, size_t rows, int C,
                     __nv_bfloat16* __restrict__ dst, int dst_ld, int dst_coff) {
  const int cg_per = C / 8;
  const size_t total = rows * cg_per;
  for (size_t idx = blockIdx.x * static_cast<size_t>(256) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * 256) {
    const int cg = idx % cg_per;
    const size_t r = idx / cg_per;
    float v[8];
    load8_f16(src + r * src_ld + src_coff + cg * 8, v);
    store8_bf16(dst + r * dst_ld + dst_coff + cg * 8, v);
  }
}

template <int MODE>
int launch_in_bwd(const InBwdArgs& a, int NB, bool apply, int blocks, cudaStream_t stream) {
  if (apply)
    in_bwd_apply_kernel<MODE><<<dim3(blocks, NB), 256, 0, stream>>>(a);
  else
    in_bwd_reduce_kernel<MODE><<<dim3(blocks, NB), 256, 0, stream>>>(a);
  NC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int bwd_blocks() { return num_sms() * 8; }

// InstanceNorm3d(affine=False) + ReLU backward of one layer, both passes.  scratch: bwd_blocks() * NB * 2 * C floats.
int in_relu_bwd(const void* raw, const float* mean_rstd, int NB, int D, int H, int W, int C, int mode,
                const void* grad, int grad_ld, int grad_coff, const float* du, const float* w_head, const void* dpool,
                float* scratch, float* m12, void* d_raw, cudaStream_t stream) {
  if (C != 64 && C != 128 && C != 256) return set_error("in_relu_bwd: C must be 64, 128 or 256");
  if (mode < 0 || mode > 2) return set_error("in_relu_bwd: bad gradient source mode");
  if (mode == 2 && ((D | H | W) & 1)) return set_error("in_relu_bwd: pooling needs even D, H, W");
  if (mode != 1 && (grad_ld % 8 || grad_coff % 8)) return set_error("in_relu_bwd: gradient slice must be 16-byte aligned");
  if (mode == 1 && C != 64) return set_error("in_relu_bwd: the head feeds a 64-channel layer");
  if (static_cast<long long>(D) * H * W >= (1ll << 31)) return set_error("in_relu_bwd: cube too large");
  if (NB > 65535) return set_error("in_relu_bwd: NB too large");
  InBwdArgs a{};
  a.raw = static_cast<const __half*>(raw);
  a.mr = mean_rstd;
  a.m12 = m12;
  a.D = D, a.H = H, a.W = W, a.C = C;
  a.src.a = static_cast<const __nv_bfloat16*>(grad);
  a.src.a_ld = grad_ld, a.src.a_coff = grad_coff;
  a.src.du = du, a.src.w = w_head;
  a.src.dpool = static_cast<const __nv_bfloat16*>(dpool);
  a.partial = scratch;
  a.d_raw = static_cast<__nv_bfloat16*>(d_raw);
  const int blocks = bwd_blocks();
  for (int pass = 0; pass < 2; ++pass) {
    int rc = mode == 0   ? launch_in_bwd<0>(a, NB, pass == 1, blocks, stream)
             : mode == 1 ? launch_in_bwd<1>(a, NB, pass == 1, blocks, stream)
                         : launch_in_bwd<2>(a, NB, pass == 1, blocks, stream);
    if (rc) return rc;
    if (pass == 0) {
      colsum_finalize_kernel<<<dim3((2 * C + 31) / 32, NB), 1024, 0, stream>>>(
          scratch, blocks, 2 * C, 1.0 / (static_cast<double>(D) * H * W), m12);
      NC_CUDA(cudaGetLastError());
    }
  }
  return 0;
}

// scratch: bwd_blocks() * NB * 68 floats; grads: (68) = [d w1 (64) | d b1 | d w2 | d b2 | 0], summed over samples
int head_bwd(const void* raw, const float* mean_rstd, const float* hp, const float* dout, int NB, int D, int H, int W,
             float* du, float* scratch, float* grads, cudaStream_t stream) {
  if (static_cast<long long>(D) * H * W >= (1ll << 31)) return set_error("head_bwd: cube too large");
  const int blocks = bwd_blocks();
  head_bwd_kernel<<<dim3(blocks, NB), 256, 0, stream>>>(static_cast<const __half*>(raw), mean_rstd, hp, dout,
                                                        static_cast<unsigned>(D) * H * W, du, scratch);
  NC_CUDA(cudaGetLastError());
  colsum_finalize_kernel<<<dim3((HEAD_COLS + 31) / 32, 1), 1024, 0, stream>>>(scratch, blocks * NB, HEAD_COLS, 1.0,
                                                                             grads);
  NC_CUDA(cudaGetLastError());
  return 0;
}

// scratch: bwd_blocks() / 4 * NB * 1728 floats; dw: (64, 27), summed over samples; dy_fmt 0 = fp16, 1 = bf16
int conv1_wgrad(const float* x, const void* dy, int dy_fmt, int NB, int D, int H, int W, float* scratch, float* dw,
                cudaStream_t stream) {
  if (static_cast<long long>(D) * H * W >= (1ll << 31)) return set_error("conv1_wgrad: cube too large");
  const int blocks = bwd_blocks() / 4;
  const int tz = (D + w1::TZ - 1) / w1::TZ, ty = (H + w1::TY - 1) / w1::TY, tx = (W + w1::TX - 1) / w1::TX;
  static bool attr[2][64] = {{false}};
  if (dy_fmt) {
    if (first_use_on_device(attr[1]))
      NC_CUDA(cudaFuncSetAttribute(conv1_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   w1::SMEM_BYTES));
    conv1_wgrad_kernel<false><<<dim3(blocks, NB), w1::THREADS, w1::SMEM_BYTES, stream>>>(
        x, static_cast<const uint16_t*>(dy), D, H, W, ty, tx, tz * ty * tx, scratch);
  } else {
    if (first_use_on_device(attr[0]))
      NC_CUDA(cudaFuncSetAttribute(conv1_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   w1::SMEM_BYTES));
    conv1_wgrad_kernel<true><<<dim3(blocks, NB), w1::THREADS, w1::SMEM_BYTES, stream>>>(
        x, static_cast<const uint16_t*>(dy), D, H, W, ty, tx, tz * ty * tx, scratch);
  }
  NC_CUDA(cudaGetLastError());
  colsum_finalize_kernel<<<dim3(1728 / 32, 1), 1024, 0, stream>>>(scratch, blocks * NB, 1728, 1.0, dw);
  NC_CUDA(cudaGetLastError());
  return 0;
}

int space_to_depth_bf16(const void* src, int ld, int coff, int NB, int D, int H, int W, int C, void* out,
                        cudaStream_t stream) {
  if (C % 8 || ld % 8 || coff % 8) return set_error("space_to_depth: channel counts must be multiples of 8");
  if (static_cast<long long>(D) * H * W * C >= (1ll << 31)) return set_error("space_to_depth: tensor too large");
  s2d_kernel<<<dim3(bwd_blocks(), NB), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), ld, coff, D, H, W, C,
                                                         static_cast<__nv_bfloat16*>(out));
  NC_CUDA(cudaGetLastError());
  return 0;
}

// scratch: bwd_blocks() * NB * C floats; out: (C), summed over samples
int colsum_bf16(const void* src, int ld, int coff, int NB, long long rows, int C, float* scratch, float* out,
                cudaStream_t stream) {
  if (C % 8 || 256 % (C / 8) || ld % 8 || coff % 8) return set_error("colsum_bf16: unsupported channel count");
  if (rows >= (1ll << 31)) return set_error("colsum_bf16: too many rows");
  const int blocks = bwd_blocks();
  colsum_bf16_kernel<<<dim3(blocks, NB), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), ld, coff,
                                                           static_cast<unsigned>(rows), C, scratch);
  NC_CUDA(cudaGetLastError());
  colsum_finalize_kernel<<<dim3((C + 31) / 32, 1), 1024, 0, stream>>>(scratch, blocks * NB, C, 1.0, out);
  NC_CUDA(cudaGetLastError());
  return 0;
}

int cast_f16_bf16(const void* src, int src_ld, int src_coff, long long rows, int C, void* dst, int dst_ld,
                  int dst_coff, cudaStream_t stream) {
  if (C % 8 || src_ld % 8 || src_coff % 8 || dst_ld % 8 || dst_coff % 8)
    return set_error("cast_f16_bf16: channel counts must be multiples of 8");
  cast_f16_bf16_kernel<<<bwd_blocks(), 256, 0, stream>>>(static_cast<const __half*>(src), src_ld, src_coff,
                                                         static_cast<size_t>(rows), C,
                                                         static_cast<__nv_bfloat16*>(dst), dst_ld, dst_coff);
  NC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nc

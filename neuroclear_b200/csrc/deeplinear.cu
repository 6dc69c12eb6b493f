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

// out[v][kh * 7 + kw] = x[d][h + kh - 3][w + kw - 3] (zero outside the plane), channels 49..63 = 0
__global__ void __launch_bounds__(256)
im2col49_kernel(const float* __restrict__ x, int D, int H, int W, int fmt, uint16_t* __restrict__ out) {
  const int nb = blockIdx.y;
  const unsigned total = static_cast<unsigned>(D) * H * W * 8;
  const float* xc = x + static_cast<size_t>(nb) * D * H * W;
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const int g = idx & 7;
    const unsigned v = idx >> 3;
    const int w = v % W;
    const unsigned r = v / W;
    const int h = r % H, d = r / H;
    uint16_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = g * 8 + i;
      const int kh = c / 7, kw = c - kh * 7;
      const int zh = h + kh - 3, zw = w + kw - 3;
      const bool in = c < 49 && zh >= 0 && zh < H && zw >= 0 && zw < W;
      o[i] = to_bits(in ? __ldg(xc + (static_cast<size_t>(d) * H + zh) * W + zw) : 0.f, fmt);
    }
    *reinterpret_cast<uint4*>(out + (static_cast<size_t>(nb) * D * H * W + v) * 64 + g * 8) =
        *reinterpret_cast<const uint4*>(o);
  }
}

// dx[d][h][w] = sum_{kh,kw} g[d][h - kh + 3][w - kw + 3][kh * 7 + kw]   (adjoint of im2col49; g bf16)
__global__ void __launch_bounds__(256)
col2im49_kernel(const __nv_bfloat16* __restrict__ g, int D, int H, int W, float* __restrict__ dx) {
  const int nb = blockIdx.y;
  const unsigned total = static_cast<unsigned>(D) * H * W;
  const __nv_bfloat16* gc = g + static_cast<size_t>(nb) * total * 64;
  for (unsigned v = blockIdx.x * 256u + threadIdx.x; v < total; v += gridDim.x * 256u) {
    const int w = v % W;
    const unsigned r = v / W;
    const int h = r % H, d = r / H;
    float s = 0.f;
#pragma unroll
    for (int kh = 0; kh < 7; ++kh) {
      const int zh = h - kh + 3;
      if (zh < 0 || zh >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        const int zw = w - kw + 3;
        if (zw >= 0 && zw < W)
          s += __bfloat162float(gc[((static_cast<size_t>(d) * H + zh) * W + zw) * 64 + kh * 7 + kw]);
      }
    }
    dx[static_cast<size_t>(nb) * total + v] = s;
  }
}

// out[v] = sum_{tap, ci} h[v + tap - 1][ci] * K[ci][tap]: 64 -> 1 k3 stencil.  8 lanes per voxel, 8 channels per
// lane; K (64 x 27 floats) sits in shared memory as [tap][ci].
__global__ void __launch_bounds__(256)
stencil64to1_kernel(const __half* __restrict__ hin, const float* __restrict__ K, int D, int H, int W,
                    float* __restrict__ out) {
  __shared__ float Ks[27][64];
  for (int i = threadIdx.x; i < 27 * 64; i += 256) Ks[i % 27][i / 27] = __ldg(K + i);  // K is [ci][tap]
  __syncthreads();
  const int nb = blockIdx.y;
  const int sub = threadIdx.x & 7;
  const unsigned voxels = static_cast<unsigned>(D) * H * W;
  const __half* hc = hin + static_cast<size_t>(nb) * voxels * 64;
  const unsigned per_iter = gridDim.x * 32u;
  for (unsigned v0 = blockIdx.x * 32u; v0 < voxels; v0 += per_iter) {
    const unsigned v = v0 + (threadIdx.x >> 3);
    const bool ok = v < voxels;
    const unsigned vv = ok ? v : 0u;
    const int w = vv % W;
    const unsigned r = vv / W;
    const int h = r % H, d = r / H;
    float acc = 0.f;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const int zd = d + kd - 1;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int zh = h + kh - 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int zw = w + kw - 1;
          if (zd < 0 || zd >= D || zh < 0 || zh >= H || zw < 0 || zw >= W) continue;
          const uint4 raw = __ldg(reinterpret_cast<const uint4*>(
              hc + ((static_cast<size_t>(zd) * H + zh) * W + zw) * 64 + sub * 8));
          const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
          const float* kk = &Ks[(kd * 3 + kh) * 3 + kw][sub * 8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h2[i]);
            acc = fmaf(f.x, kk[2 * i], acc);
            acc = fmaf(f.y, kk[2 * i + 1], acc);
          }
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (ok && sub == 0) out[static_cast<size_t>(nb) * voxels + v] = acc;
  }
}

// dh[v][ci] = sum_tap dout[v - tap + 1] * K[ci][tap]   (adjoint of the stencil w.r.t. its input), bf16 out
__global__ void __launch_bounds__(256)
stencil1to64_kernel(const float* __restrict__ dout, const float* __restrict__ K, int D, int H, int W,
                    __nv_bfloat16* __restrict__ dh) {
  __shared__ float Ks[27][64];
  for (int i = threadIdx.x; i < 27 * 64; i += 256) Ks[i % 27][i / 27] = __ldg(K + i);
  __syncthreads();
  const int nb = blockIdx.y;
  const int sub = threadIdx.x & 7;
  const unsigned voxels = static_cast<unsigned>(D) * H * W;
  const float* dc = dout + static_cast<size_t>(nb) * voxels;
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < voxels * 8; idx += gridDim.x * 256u) {
    const unsigned v = idx >> 3;
    const int w = v % W;
    const unsigned r = v / W;
    const int h = r % H, d = r / H;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const int zd = d - kd + 1;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int zh = h - kh + 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int zw = w - kw + 1;
          if (zd < 0 || zd >= D || zh < 0 || zh >= H || zw < 0 || zw >= W) continue;
          const float g = __ldg(dc + (static_cast<size_t>(zd) * H + zh) * W + zw);
          const float* kk = &Ks[(kd * 3 + kh) * 3 + kw][sub * 8];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(g, kk[i], acc[i]);
        }
      }
    }
    uint4 pk;
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int i = 0; i < 4; ++i) p2[i] = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
    *reinterpret_cast<uint4*>(dh + (static_cast<size_t>(nb) * voxels + v) * 64 + sub * 8) = pk;
  }
}

}  // namespace

int im2col49(const float* x, int NB, int D, int H, int W, int fmt, void* out, cudaStream_t stream) {
  if (static_cast<long long>(D) * H * W * 8 >= (1ll << 32)) return set_error("im2col49: volume too large");
  im2col49_kernel<<<dim3(num_sms() * 8, NB), 256, 0, stream>>>(x, D, H, W, fmt, static_cast<uint16_t*>(out));
  NC_CUDA(cudaGetLastError());
  return 0;
}
int col2im49(const void* g, int NB, int D, int H, int W, float* dx, cudaStream_t stream) {
  if (static_cast<long long>(D) * H * W >= (1ll << 31)) return set_error("col2im49: volume too large");
  col2im49_kernel<<<dim3(num_sms() * 8, NB), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(g), D, H, W, dx);
  NC_CUDA(cudaGetLastError());
  return 0;
}
int stencil64to1_fwd(const void* h, const float* K, int NB, int D, int H, int W, float* out, cudaStream_t stream) {
  if (static_cast<long long>(D) * H * W >= (1ll << 31)) return set_error("stencil64to1_fwd: volume too large");
  stencil64to1_kernel<<<dim3(num_sms() * 8, NB), 256, 0, stream>>>(static_cast<const __half*>(h), K, D, H, W, out);
  NC_CUDA(cudaGetLastError());
  return 0;
}
int stencil64to1_bwd_data(const float* dout, const float* K, int NB, int D, int H, int W, void* dh,
                          cudaStream_t stream) {
  if (static_cast<long long>(D) * H * W * 8 >= (1ll << 32)) return set_error("stencil64to1_bwd_data: volume too large");
  stencil1to64_kernel<<<dim3(num_sms() * 8, NB), 256, 0, stream>>>(dout, K, D, H, W,
                                                                  static_cast<__nv_bfloat16*>(dh));
  NC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nc

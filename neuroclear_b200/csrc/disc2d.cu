// 2-D PatchGAN discriminator path of the apollo model (reference models/networks.py:1009-1067 NLayerDiscriminator
// with dimension=2, built at models/axial_to_lateral_gan_apollo_model.py:99-123) and the LSGAN loss
// (networks.py:252-319): forward AND backward kernels.
//
// One discriminator pass sees ONE 2-D image of ~108^2..148^2 pixels (a slice or a max-intensity projection of the
// cube): ~1 GFLOP, 18 passes + backward per training step = ~0.1 % of the step's FLOPs.  The path is bound by launch
// latency, not by throughput, so these are plain fp32 CUDA-core kernels on NCHW tensors, one launch
// per layer and direction, written for low latency (no tensor-core staging, no layout changes, no workspaces).
#include "internal.h"

namespace nc {

// ------------------------------------------------------------------------------------------------ Conv2d k4 p1
// y[n,co,ho,wo] = b[co] + sum_{ci,kh,kw} w[co,ci,kh,kw] * x[n,ci,ho*s-1+kh,wo*s-1+kw]   (+ optional LeakyReLU)
__global__ void __launch_bounds__(128)
conv2d_k4_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int Cin,
                     int H, int W, int Cout, int Ho, int Wo, int stride, float slope, float* __restrict__ y) {
  const int p = blockIdx.x * 128 + threadIdx.x;
  const int co = blockIdx.y, n = blockIdx.z;
  if (p >= Ho * Wo) return;
  const int ho = p / Wo, wo = p - ho * Wo;
  const int h0 = ho * stride - 1, w0 = wo * stride - 1;
  const float* xn = x + static_cast<size_t>(n) * Cin * H * W;
  const float4* wc = reinterpret_cast<const float4*>(w + static_cast<size_t>(co) * Cin * 16);
  float acc = b ? __ldg(b + co) : 0.f;
  for (int ci = 0; ci < Cin; ++ci) {
    const float* xc = xn + static_cast<size_t>(ci) * H * W;
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const int hh = h0 + kh;
      const float4 wk = __ldg(wc + ci * 4 + kh);
      if (hh < 0 || hh >= H) continue;
      const float* xr = xc + hh * W;
      const float wv[4] = {wk.x, wk.y, wk.z, wk.w};
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const int ww = w0 + kw;
        if (ww >= 0 && ww < W) acc = fmaf(wv[kw], __ldg(xr + ww), acc);
      }
    }
  }
  if (slope != 1.f) acc = acc > 0.f ? acc : acc * slope;
  y[(static_cast<size_t>(n) * Cout + co) * Ho * Wo + p] = acc;
}

// dx[n,ci,h,w] = sum_{co,kh,kw : (h+1-kh) = ho*s, (w+1-kw) = wo*s} w[co,ci,kh,kw] * dy[n,co,ho,wo]
__global__ void __launch_bounds__(128)
conv2d_k4_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, int Cin, int H, int W, int Cout,
                       int Ho, int Wo, int stride, float* __restrict__ dx) {
  const int p = blockIdx.x * 128 + threadIdx.x;
  const int ci = blockIdx.y, n = blockIdx.z;
  if (p >= H * W) return;
  const int h = p / W, ww = p - h * W;
  int hos[4], wos[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int th = h + 1 - k, tw = ww + 1 - k;
    hos[k] = (th >= 0 && th % stride == 0 && th / stride < Ho) ? th / stride : -1;
    wos[k] = (tw >= 0 && tw % stride == 0 && tw / stride < Wo) ? tw / stride : -1;
  }
  const float* dyn = dy + static_cast<size_t>(n) * Cout * Ho * Wo;
  float acc = 0.f;
  for (int co = 0; co < Cout; ++co) {
    const float* wk = w + (static_cast<size_t>(co) * Cin + ci) * 16;
    const float* dyc = dyn + static_cast<size_t>(co) * Ho * Wo;
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      if (hos[kh] < 0) continue;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw)
        if (wos[kw] >= 0) acc = fmaf(__ldg(wk + kh * 4 + kw), __ldg(dyc + hos[kh] * Wo + wos[kw]), acc);
    }
  }
  dx[(static_cast<size_t>(n) * Cin + ci) * H * W + p] = acc;
}

// dw[co,ci,kh,kw] = sum_{n,ho,wo} dy[n,co,ho,wo] * x[n,ci,ho*s-1+kh,wo*s-1+kw]; one warp per (co,ci), lanes over
// the 16 taps x 2 halves of the output positions, fixed-order shuffle reduction -> deterministic.
__global__ void __launch_bounds__(128)
conv2d_k4_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int N, int Cin, int H, int W,
                       int Cout, int Ho, int Wo, int stride, float* __restrict__ dw) {
  const int warp = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Cout * Cin) return;
  const int co = warp / Cin, ci = warp - co * Cin;
  const int tap = lane & 15, half = lane >> 4;
  const int kh = tap >> 2, kw = tap & 3;
  float acc = 0.f;
  const int P = Ho * Wo;
  for (int n = 0; n < N; ++n) {
    const float* xc = x + (static_cast<size_t>(n) * Cin + ci) * H * W;
    const float* dyc = dy + (static_cast<size_t>(n) * Cout + co) * P;
    for (int p = half; p < P; p += 2) {
      const int ho = p / Wo, wo = p - ho * Wo;
      const int hh = ho * stride - 1 + kh, ww = wo * stride - 1 + kw;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) acc = fmaf(__ldg(dyc + p), __ldg(xc + hh * W + ww), acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
  if (half == 0) dw[(static_cast<size_t>(co) * Cin + ci) * 16 + tap] = acc;
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
  dim3 grid((Ho * Wo + 127) / 128, Cout, N);
  conv2d_k4_fwd_kernel<<<grid, 128, 0, stream>>>(x, w, b, Cin, H, W, Cout, Ho, Wo, stride, slope, y);
  NC_CUDA(cudaGetLastError());
  return 0;
}
int conv2d_k4_dgrad(const float* dy, const float* w, int N, int Cin, int H, int W, int Cout, int stride, float* dx,
                    cudaStream_t stream) {
  if (stride != 1 && stride != 2) return set_error("conv2d_k4: stride must be 1 or 2");
  const int Ho = out_extent(H, stride), Wo = out_extent(W, stride);
  if (Cin > 65535 || N > 65535) return set_error("conv2d_k4: too many channels / images");
  dim3 grid((H * W + 127) / 128, Cin, N);
  conv2d_k4_dgrad_kernel<<<grid, 128, 0, stream>>>(dy, w, Cin, H, W, Cout, Ho, Wo, stride, dx);
  NC_CUDA(cudaGetLastError());
  return 0;
}
int conv2d_k4_wgrad(const float* x, const float* dy, int N, int Cin, int H, int W, int Cout, int stride, float* dw,
                    float* db, cudaStream_t stream) {
  if (stride != 1 && stride != 2) return set_error("conv2d_k4: stride must be 1 or 2");
  const int Ho = out_extent(H, stride), Wo = out_extent(W, stride);
  const long long warps = static_cast<long long>(Cout) * Cin;
  conv2d_k4_wgrad_kernel<<<static_cast<unsigned>((warps + 3) / 4), 128, 0, stream>>>(x, dy, N, Cin, H, W, Cout, Ho, Wo,
                                                                                    stride, dw);
  NC_CUDA(cudaGetLastError());
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
__global__ void __launch_bounds__(256)
loss_fwd_kernel(const float* __restrict__ p, const float* __restrict__ q, float target, long long n, int mode,
                float* __restrict__ loss) {
  __shared__ double red[256];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 256) {
    const float d = p[i] - (mode == 0 ? target : q[i]);
    s += mode == 0 ? static_cast<double>(d) * d : fabs(static_cast<double>(d));
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = static_cast<float>(red[0] / static_cast<double>(n));
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
  loss_fwd_kernel<<<1, 256, 0, stream>>>(p, q, target, n, mode, loss);
  NC_CUDA(cudaGetLastError());
  return 0;
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

}  // namespace nc

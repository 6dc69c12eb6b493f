// Whole-network entry points for the 2-D PatchGAN discriminator (reference models/networks.py:1009-1067,
// NLayerDiscriminator with dimension = 2, InstanceNorm, no sigmoid): the layer loop of one forward / one backward
// pass runs here, on the host side of the library, so a pass costs one call from the binding instead of ~25.
// A training iteration of the apollo model runs 18 such passes on ~100 x 100 images (launch-latency bound).
#include "internal.h"

namespace nc {

namespace {

constexpr int MAX_LAYERS = 8;
constexpr float SLOPE = 0.2f;
constexpr float EPS = 1e-5f;

struct Plan {
  int L;                        // number of convolutions
  int c[MAX_LAYERS + 1];        // channels: c[i] -> c[i + 1]
  int stride[MAX_LAYERS];
  bool has_in[MAX_LAYERS];      // InstanceNorm + LeakyReLU after conv i
  int h[MAX_LAYERS + 1], w[MAX_LAYERS + 1];  // spatial size of the input of conv i (h[L] = output)
  size_t y[MAX_LAYERS], z[MAX_LAYERS], mr[MAX_LAYERS];  // workspace offsets (floats)
  size_t tmp0, tmp1, total;
};

int make_plan(Plan& p, int N, int H, int W, int ndf, int n_layers) {
  if (n_layers < 1 || n_layers + 2 > MAX_LAYERS) return set_error("patchgan: unsupported n_layers %d", n_layers);
  p.L = n_layers + 2;
  p.c[0] = 1;
  p.c[1] = ndf;
  for (int i = 2; i <= n_layers; ++i) p.c[i] = ndf * ((1 << (i - 1)) < 8 ? (1 << (i - 1)) : 8);
  p.c[n_layers + 1] = ndf * ((1 << n_layers) < 8 ? (1 << n_layers) : 8);
  p.c[n_layers + 2] = 1;
  p.h[0] = H, p.w[0] = W;
  size_t off = 0, biggest = static_cast<size_t>(N) * H * W;
  for (int i = 0; i < p.L; ++i) {
    p.stride[i] = i < n_layers ? 2 : 1;
    p.has_in[i] = i >= 1 && i <= n_layers;
    p.h[i + 1] = (p.h[i] - 2) / p.stride[i] + 1;
    p.w[i + 1] = (p.w[i] - 2) / p.stride[i] + 1;
    if (p.h[i] < 2 || p.w[i] < 2 || p.h[i + 1] < 1 || p.w[i + 1] < 1)
      return set_error("patchgan: a %d x %d image is too small for %d layers", H, W, n_layers);
    const size_t act = static_cast<size_t>(N) * p.c[i + 1] * p.h[i + 1] * p.w[i + 1];
    if (act > biggest) biggest = act;
    p.y[i] = off, off += act;
    p.z[i] = off, off += p.has_in[i] ? act : 0;
    p.mr[i] = off, off += p.has_in[i] ? static_cast<size_t>(N) * p.c[i + 1] * 2 : 0;
  }
  p.tmp0 = off, off += biggest;
  p.tmp1 = off, off += biggest;
  p.total = off;
  return 0;
}

}  // namespace

long long patchgan_ws_floats(int N, int H, int W, int ndf, int n_layers) {
  Plan p;
  if (make_plan(p, N, H, W, ndf, n_layers)) return -1;
  return static_cast<long long>(p.total);
}

// pred: (N, 1, h_out, w_out).  ws keeps every activation for the backward pass.
int patchgan_fwd(const float* x, int N, int H, int W, int ndf, int n_layers, const float* const* weights,
                 const float* const* biases, float* ws, float* pred, cudaStream_t stream) {
  Plan p;
  if (int rc = make_plan(p, N, H, W, ndf, n_layers)) return rc;
  const float* in = x;
  for (int i = 0; i < p.L; ++i) {
    float* y = i == p.L - 1 ? pred : ws + p.y[i];
    if (int rc = conv2d_k4_fwd(in, weights[i], biases[i], N, p.c[i], p.h[i], p.w[i], p.c[i + 1], p.stride[i],
                               i == 0 ? SLOPE : 1.0f, y, stream))
      return rc;
    if (p.has_in[i]) {
      if (int rc = in2d_lrelu_fwd(y, N * p.c[i + 1], p.h[i + 1] * p.w[i + 1], EPS, SLOPE, ws + p.z[i], ws + p.mr[i],
                                  stream))
        return rc;
      in = ws + p.z[i];
    } else {
      in = y;
    }
  }
  return 0;
}

// dx (nullable): gradient w.r.t. the image; dweights / dbiases (nullable together): parameter gradients, overwritten.
int patchgan_bwd(const float* x, const float* dpred, int N, int H, int W, int ndf, int n_layers,
                 const float* const* weights, float* ws, float* dx, float* const* dweights, float* const* dbiases,
                 cudaStream_t stream) {
  Plan p;
  if (int rc = make_plan(p, N, H, W, ndf, n_layers)) return rc;
  const float* g = dpred;
  float* tmp[2] = {ws + p.tmp0, ws + p.tmp1};
  int t = 0;
  for (int i = p.L - 1; i >= 0; --i) {
    const long long act = static_cast<long long>(N) * p.c[i + 1] * p.h[i + 1] * p.w[i + 1];
    if (p.has_in[i]) {
      if (int rc = in2d_lrelu_bwd(g, ws + p.y[i], ws + p.mr[i], N * p.c[i + 1], p.h[i + 1] * p.w[i + 1], SLOPE,
                                  tmp[t], stream))
        return rc;
      g = tmp[t], t ^= 1;
    } else if (i == 0) {  // LeakyReLU fused into the first conv: y is the activated output
      if (int rc = lrelu_bwd(g, ws + p.y[0], act, SLOPE, tmp[t], stream)) return rc;
      g = tmp[t], t ^= 1;
    }
    const float* in = i == 0 ? x : (p.has_in[i - 1] ? ws + p.z[i - 1] : ws + p.y[i - 1]);
    if (dweights) {
      if (int rc = conv2d_k4_wgrad(in, g, N, p.c[i], p.h[i], p.w[i], p.c[i + 1], p.stride[i], dweights[i],
                                   dbiases ? dbiases[i] : nullptr, stream))
        return rc;
    }
    if (i > 0 || dx) {
      float* gin = i == 0 ? dx : tmp[t];
      if (int rc = conv2d_k4_dgrad(g, weights[i], N, p.c[i], p.h[i], p.w[i], p.c[i + 1], p.stride[i], gin, stream))
        return rc;
      g = gin, t ^= 1;
    }
  }
  return 0;
}

}  // namespace nc

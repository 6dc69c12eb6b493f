// nc_unet_deconv_infer_cube: the whole Unet_deconv forward (reference models/networks.py:512-538) for a batch of cubes
// in ONE library call — the layer sequence of neuroclear_b200/unet_engine.py::UnetDeconvEngine.forward on the host
// side of the library.  Stateless: weights (already packed) and workspace are caller-owned device memory; all 28
// launches go to `stream`, nothing synchronises or allocates, and the TMA descriptors travel as kernel parameters — so
// the call can be captured into a CUDA graph by the caller (fixed x / out / workspace pointers) and replayed.
#include "../../include/neuroclear_b200.h"
#include "internal.h"

using namespace nc;

namespace {
constexpr float IN_EPS = 1e-5f;  // torch.nn.InstanceNorm3d default (networks.get_norm_layer, networks.py:33-34)

struct Layout {
  size_t raw0a, raw0b, a1, cat1, p1, raw1a, raw1b, cat2, p2, raw2a, raw2b, stats, mrA, mrB, fin, total;
};
inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

Layout make_layout(int nb, int d, int h, int w) {
  const size_t l0 = static_cast<size_t>(d) * h * w, l1 = l0 / 8, l2 = l0 / 64, n = nb;
  size_t rows = conv_cin1_stats_tiles(nb, d, h, w) * 64;
  auto upd = [&](size_t r) { if (r > rows) rows = r; };
  upd(conv3d_k3_stats_tiles(nb, d, h, w, 64) * 64);
  upd(conv3d_k3_stats_tiles(nb, d / 2, h / 2, w / 2, 128) * 128);
  upd(conv3d_k3_stats_tiles(nb, d / 4, h / 4, w / 4, 256) * 256);
  Layout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += align256(bytes); return o; };
  L.raw0a = take(n * l0 * 64 * 2), L.raw0b = take(n * l0 * 64 * 2), L.a1 = take(n * l0 * 64 * 2);
  L.cat1 = take(n * l0 * 128 * 2), L.p1 = take(n * l1 * 64 * 2);
  L.raw1a = take(n * l1 * 128 * 2), L.raw1b = take(n * l1 * 128 * 2), L.cat2 = take(n * l1 * 256 * 2);
  L.p2 = take(n * l2 * 128 * 2), L.raw2a = take(n * l2 * 256 * 2), L.raw2b = take(n * l2 * 256 * 2);
  L.stats = take(rows * 2 * 4), L.mrA = take(n * 2 * 256 * 4), L.mrB = take(n * 2 * 256 * 4);
  L.fin = take(in_stats_scratch_bytes(nb, 256));
  L.total = off;
  return L;
}
}  // namespace

extern "C" {

int64_t nc_unet_deconv_workspace_bytes(int32_t nb, int32_t d, int32_t h, int32_t w) {
  if (nb <= 0 || d <= 0 || h <= 0 || w <= 0 || d % 4 || h % 4 || w % 4)
    return set_error("unet_deconv: need nb > 0 and D, H, W divisible by 4");
  return static_cast<int64_t>(make_layout(nb, d, h, w).total);
}

int nc_unet_deconv_workspace_init(void* workspace, int64_t workspace_bytes, int32_t nb, int32_t d, int32_t h, int32_t w,
                                  nc_stream_t stream) {
  const int64_t need = nc_unet_deconv_workspace_bytes(nb, d, h, w);
  if (need < 0) return -1;
  if (workspace_bytes < need) return set_error("unet_deconv: workspace too small");
  const Layout L = make_layout(nb, d, h, w);
  // only the arrival counters of the statistics finalisation carry state between launches (they reset themselves)
  NC_CUDA(cudaMemsetAsync(static_cast<char*>(workspace) + L.fin, 0, in_stats_scratch_bytes(nb, 256),
                          static_cast<cudaStream_t>(stream)));
  return 0;
}

int nc_unet_deconv_infer_cube(const float* x, int32_t nb, int32_t d, int32_t h, int32_t w,
                              const nc_unet_deconv_weights* wt, void* workspace, int64_t workspace_bytes,
                              int32_t crop, float* out, nc_stream_t stream) {
  const int64_t need = nc_unet_deconv_workspace_bytes(nb, d, h, w);
  if (need < 0) return -1;
  if (workspace_bytes < need) return set_error("unet_deconv: workspace too small (%lld < %lld)",
                                               (long long)workspace_bytes, (long long)need);
  if (!x || !wt || !workspace || !out) return set_error("unet_deconv: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Layout L = make_layout(nb, d, h, w);
  char* ws = static_cast<char*>(workspace);
  void *raw0a = ws + L.raw0a, *raw0b = ws + L.raw0b, *a1 = ws + L.a1, *cat1 = ws + L.cat1, *p1 = ws + L.p1;
  void *raw1a = ws + L.raw1a, *raw1b = ws + L.raw1b, *cat2 = ws + L.cat2, *p2 = ws + L.p2;
  void *raw2a = ws + L.raw2a, *raw2b = ws + L.raw2b, *fin = ws + L.fin;
  float* st = reinterpret_cast<float*>(ws + L.stats);
  float* mrA = reinterpret_cast<float*>(ws + L.mrA);
  float* mrB = reinterpret_cast<float*>(ws + L.mrB);
  const int d1 = d / 2, h1 = h / 2, w1 = w / 2, d2 = d / 4, h2 = h / 4, w2 = w / 4;

  auto stats = [&](bool first, int dd, int hh, int ww, int c, float* mr) -> int {
    const size_t tiles = first ? conv_cin1_stats_tiles(nb, dd, hh, ww) : conv3d_k3_stats_tiles(nb, dd, hh, ww, c);
    return in_stats_finalize(st, nb, static_cast<long long>(tiles / nb), c, static_cast<long long>(dd) * hh * ww,
                             IN_EPS, fin, mr, s);
  };
  auto conv = [&](int layer, const void* src, const float* src_mr, int dd, int hh, int ww, int cin, int cout, void* raw,
                  float* raw_mr) -> int {
    if (int rc = conv3d_k3_fwd(src, src_mr, nb, dd, hh, ww, cin, wt->k3[layer], cout, raw, st, s)) return rc;
    return stats(false, dd, hh, ww, cout, raw_mr);
  };
#define NC_TRY(expr) \
  do {               \
    if (int rc_ = (expr)) return rc_; \
  } while (0)
  // ---- level 0 down: U1, U2; conv1 -> cat1[:, :64] and maxpool1
  NC_TRY(conv3d_cin1_k3_fwd(x, wt->first, nb, d, h, w, 64, raw0a, st, s));
  NC_TRY(stats(true, d, h, w, 64, mrA));
  NC_TRY(in_relu_apply(raw0a, mrA, nb, d, h, w, 64, a1, 64, 0, nullptr, s));
  NC_TRY(conv(0, a1, nullptr, d, h, w, 64, 64, raw0b, mrB));
  NC_TRY(in_relu_apply(raw0b, mrB, nb, d, h, w, 64, cat1, 128, 0, p1, s));
  // ---- level 1 down: U3, U4 (IN + ReLU of U3 inside U4); conv2 -> cat2[:, :128] and maxpool2
  NC_TRY(conv(1, p1, nullptr, d1, h1, w1, 64, 128, raw1a, mrA));
  NC_TRY(conv(2, raw1a, mrA, d1, h1, w1, 128, 128, raw1b, mrB));
  NC_TRY(in_relu_apply(raw1b, mrB, nb, d1, h1, w1, 128, cat2, 256, 0, p2, s));
  // ---- bottom: U5, U6, U7
  NC_TRY(conv(3, p2, nullptr, d2, h2, w2, 128, 256, raw2a, mrA));
  NC_TRY(conv(4, raw2a, mrA, d2, h2, w2, 256, 256, raw2b, mrB));
  NC_TRY(conv(5, raw2b, mrB, d2, h2, w2, 256, 256, raw2a, mrA));
  // ---- level 1 up: cat2 = [conv2 | t_conv2(IN+ReLU(U7))], U9, U10
  NC_TRY(in_relu_apply(raw2a, mrA, nb, d2, h2, w2, 256, raw2b, 256, 0, nullptr, s));
  NC_TRY(convT3d_k2s2_fwd(raw2b, nullptr, nb, d2, h2, w2, 256, wt->ct[0], wt->ct_bias[0], 128, cat2, 256, 128, s));
  NC_TRY(conv(6, cat2, nullptr, d1, h1, w1, 256, 128, raw1a, mrA));
  NC_TRY(conv(7, raw1a, mrA, d1, h1, w1, 128, 128, raw1b, mrB));
  // ---- level 0 up: cat1 = [conv1 | t_conv1(IN+ReLU(U10))], U12
  NC_TRY(in_relu_apply(raw1b, mrB, nb, d1, h1, w1, 128, raw1a, 128, 0, nullptr, s));
  NC_TRY(convT3d_k2s2_fwd(raw1a, nullptr, nb, d1, h1, w1, 128, wt->ct[1], wt->ct_bias[1], 64, cat1, 128, 64, s));
  NC_TRY(conv(8, cat1, nullptr, d, h, w, 128, 64, raw0a, mrA));
  // ---- head: IN + ReLU + 1x1x1 + 1x1x1 + sigmoid (+ border cut)
  NC_TRY(head_1x1_sigmoid_fwd(raw0a, mrA, wt->head, nb, d, h, w, 64, crop, out, s));
#undef NC_TRY
  return 0;
}

}  // extern "C"

// extern "C" surface of libneuroclear_b200.so — see include/neuroclear_b200.h for the contract of every symbol.
#include "../../include/neuroclear_b200.h"

#include "internal.h"

using namespace nc;

static inline cudaStream_t S(nc_stream_t s) { return static_cast<cudaStream_t>(s); }

extern "C" {

int nc_abi_version(void) { return NC_ABI_VERSION; }
const char* nc_last_error(void) { return last_error(); }

#ifndef NC_SOURCE_HASH
#define NC_SOURCE_HASH "unstamped"
#endif
const char* nc_build_source_hash(void) { return "NC_SOURCE_HASH=" NC_SOURCE_HASH; }

void nc_debug_set_max_ctas(int32_t n) { debug_set_max_ctas(n); }
void nc_debug_set_remainder_pairs(int32_t on) { debug_set_remainder_pairs(on); }
void nc_debug_set_disc_cluster(int32_t c) { debug_set_disc_cluster(c); }

int nc_memcpy2d_h2d_async(void* dst, const void* src, int64_t pitch_bytes, int64_t width_bytes, int64_t rows,
                          nc_stream_t stream) {
  if (rows <= 0 || width_bytes <= 0) return 0;
  if (width_bytes > pitch_bytes) return set_error("memcpy2d: width exceeds pitch");
  NC_CUDA(cudaMemcpy2DAsync(dst, static_cast<size_t>(pitch_bytes), src, static_cast<size_t>(pitch_bytes),
                            static_cast<size_t>(width_bytes), static_cast<size_t>(rows), cudaMemcpyHostToDevice,
                            S(stream)));
  return 0;
}

int nc_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return set_error("no CUDA device");
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return set_error("no CUDA device");
  return n;
}

int64_t nc_dice_geometry(const int32_t size_zyx[3], int32_t roi, int32_t overlap, int32_t padded_zyx[3],
                         int32_t steps_zyx[3]) {
  const int step = roi - overlap;
  if (roi <= 0 || overlap < 0 || step <= 0) return set_error("dice geometry: need 0 <= overlap < roi");
  int64_t n = 1;
  for (int i = 0; i < 3; ++i) {
    if (size_zyx[i] <= 0) return set_error("dice geometry: empty volume");
    const int counts = (size_zyx[i] + overlap) / step;               // util/util.py:203-205
    const int pad = step * counts + roi - size_zyx[i];               // util/util.py:207-209
    padded_zyx[i] = size_zyx[i] + pad;
    steps_zyx[i] = (padded_zyx[i] - overlap) / step;                 // diceImage_dataset.py:90-92
    n *= steps_zyx[i];
  }
  return n;
}

int nc_dice_extract_u16(const uint16_t* vol, int32_t vol_z0, int32_t vol_nz, const int32_t size_zyx[3],
                        const int32_t padded_zyx[3], const int32_t steps_zyx[3], int32_t roi, int32_t overlap,
                        int32_t border, int64_t cube_begin, int32_t cube_count, float* cubes, nc_stream_t stream) {
  return dice_extract_u16(vol, vol_z0, vol_nz, size_zyx, padded_zyx, steps_zyx, roi, overlap, border, cube_begin,
                          cube_count, cubes, S(stream));
}

int nc_dice_extract_u8(const uint8_t* vol, int32_t vol_z0, int32_t vol_nz, const int32_t size_zyx[3],
                       const int32_t padded_zyx[3], const int32_t steps_zyx[3], int32_t roi, int32_t overlap,
                       int32_t border, int64_t cube_begin, int32_t cube_count, float* cubes, nc_stream_t stream) {
  return dice_extract_u8(vol, vol_z0, vol_nz, size_zyx, padded_zyx, steps_zyx, roi, overlap, border, cube_begin,
                         cube_count, cubes, S(stream));
}

int64_t nc_conv3d_k3_stats_rows(int32_t cin, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t cout) {
  if (cin == 1) return static_cast<int64_t>(conv_cin1_stats_tiles(nb, d, h, w));
  return static_cast<int64_t>(conv3d_k3_stats_tiles(nb, d, h, w, cout));
}

int nc_pack_weights_conv3d_cin1_k3(const float* w, void* packed, nc_stream_t stream) {
  return pack_conv1_weights(w, packed, S(stream));
}
int nc_conv3d_cin1_k3_fwd(const float* x, const void* packed, int32_t nb, int32_t d, int32_t h, int32_t wdt,
                          int32_t cout, void* y_raw, float* stats_partial, nc_stream_t stream) {
  return conv3d_cin1_k3_fwd(x, packed, nb, d, h, wdt, cout, y_raw, stats_partial, S(stream));
}

int64_t nc_packed_weight_bytes(int32_t cout, int32_t cin, int32_t transposed) {
  return static_cast<int64_t>(packed_weight_bytes(cout, cin, 27, transposed));
}
int nc_pack_weights_conv3d_k3(const float* w, int32_t cout, int32_t cin, void* packed, nc_stream_t stream) {
  return pack_weights(w, packed, cout, cin, 27, 0, S(stream));
}
int nc_pack_weights_convT3d_k2s2(const float* w, int32_t cin, int32_t cout, void* packed, nc_stream_t stream) {
  return pack_weights(w, packed, cout, cin, 8, 1, S(stream));
}

int nc_conv3d_k3_fwd(const void* x, const float* in_mean_rstd, int32_t nb, int32_t d, int32_t h, int32_t w,
                     int32_t cin, const void* packed, int32_t cout, void* y_raw, float* stats_partial,
                     nc_stream_t stream) {
  return conv3d_k3_fwd(x, in_mean_rstd, nb, d, h, w, cin, packed, cout, y_raw, stats_partial, S(stream));
}
int nc_pack_weights_conv3d_k3_dgrad(const float* w, int32_t cout, int32_t cin, void* packed, nc_stream_t stream) {
  return pack_weights_dgrad(w, packed, cout, cin, S(stream));
}
int nc_conv3d_k3_dgrad(const void* dy, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t cout, const void* packed,
                       int32_t cin, void* dx, nc_stream_t stream) {
  return conv3d_k3_dgrad(dy, nb, d, h, w, cout, packed, cin, dx, S(stream));
}
int64_t nc_conv3d_wgrad_scratch_bytes(int32_t ks, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t cin,
                                      int32_t cout) {
  return static_cast<int64_t>(conv3d_wgrad_scratch_bytes(ks, nb, d, h, w, cin, cout));
}
int nc_conv3d_wgrad(const void* x, int32_t x_fmt, const void* dy, int32_t dy_fmt, int32_t nb, int32_t d, int32_t h,
                    int32_t w, int32_t cin, int32_t cout, int32_t ks, void* scratch, float* dw, nc_stream_t stream) {
  return conv3d_wgrad(x, x_fmt, dy, dy_fmt, nb, d, h, w, cin, cout, ks, scratch, dw, S(stream));
}
int nc_pack_weights_convT3d_k2s2_dgrad(const float* w, int32_t cin, int32_t cout, void* packed, nc_stream_t stream) {
  return pack_weights_convT_dgrad(w, packed, cin, cout, S(stream));
}
int nc_conv3d_k1_bf16(const void* x, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t k, const void* packed,
                      int32_t n, void* y, nc_stream_t stream) {
  return conv3d_k1_bf16(x, nb, d, h, w, k, packed, n, y, S(stream));
}
int nc_in_relu_apply_bf16(const void* raw, const float* mean_rstd, int32_t nb, int32_t d, int32_t h, int32_t w,
                          int32_t c, void* y, int32_t y_ld, int32_t y_coff, void* pooled, nc_stream_t stream) {
  return in_relu_apply_bf16(raw, mean_rstd, nb, d, h, w, c, y, y_ld, y_coff, pooled, S(stream));
}
int64_t nc_bwd_scratch_bytes(int32_t nb) { return static_cast<int64_t>(bwd_blocks()) * nb * 512 * 4; }
int nc_in_relu_bwd(const void* raw, const float* mean_rstd, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t c,
                   int32_t mode, const void* grad, int32_t grad_ld, int32_t grad_coff, const float* du,
                   const float* w_head, const void* dpool, void* scratch, float* m12, void* d_raw,
                   nc_stream_t stream) {
  return in_relu_bwd(raw, mean_rstd, nb, d, h, w, c, mode, grad, grad_ld, grad_coff, du, w_head, dpool,
                     static_cast<float*>(scratch), m12, d_raw, S(stream));
}
int nc_head_1x1_sigmoid_bwd(const void* raw, const float* mean_rstd, const float* hp, const float* dout, int32_t nb,
                            int32_t d, int32_t h, int32_t w, float* du, void* scratch, float* grads,
                            nc_stream_t stream) {
  return head_bwd(raw, mean_rstd, hp, dout, nb, d, h, w, du, static_cast<float*>(scratch), grads, S(stream));
}
int nc_conv3d_cin1_k3_wgrad(const float* x, const void* dy, int32_t dy_fmt, int32_t nb, int32_t d, int32_t h,
                            int32_t w, void* scratch, float* dw, nc_stream_t stream) {
  return conv1_wgrad(x, dy, dy_fmt, nb, d, h, w, static_cast<float*>(scratch), dw, S(stream));
}
int nc_im2col49(const float* x, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t fmt, void* out,
                nc_stream_t stream) {
  return im2col49(x, nb, d, h, w, fmt, out, S(stream));
}
int nc_col2im49(const void* g, int32_t nb, int32_t d, int32_t h, int32_t w, float* dx, nc_stream_t stream) {
  return col2im49(g, nb, d, h, w, dx, S(stream));
}
int nc_stencil64to1_fwd(const void* hin, const float* k, int32_t nb, int32_t d, int32_t h, int32_t w, float* out,
                        nc_stream_t stream) {
  return stencil64to1_fwd(hin, k, nb, d, h, w, out, S(stream));
}
int nc_stencil64to1_bwd_data(const float* dout, const float* k, int32_t nb, int32_t d, int32_t h, int32_t w, void* dh,
                             nc_stream_t stream) {
  return stencil64to1_bwd_data(dout, k, nb, d, h, w, dh, S(stream));
}
int nc_conv3d_tc_64(const void* x, int32_t fmt, int32_t nb, int32_t d, int32_t h, int32_t w, const void* packed,
                    int32_t ksd, int32_t ksp, void* y, nc_stream_t stream) {
  return conv3d_tc_64(x, fmt, nb, d, h, w, packed, ksd, ksp, y, S(stream));
}
int nc_pack_weights_64(const float* w, int32_t taps, int32_t dgrad, void* packed, nc_stream_t stream) {
  return pack_weights_64(w, packed, taps, dgrad, S(stream));
}
int nc_space_to_depth_bf16(const void* src, int32_t ld, int32_t coff, int32_t nb, int32_t d, int32_t h, int32_t w,
                           int32_t c, void* out, nc_stream_t stream) {
  return space_to_depth_bf16(src, ld, coff, nb, d, h, w, c, out, S(stream));
}
int nc_colsum_bf16(const void* src, int32_t ld, int32_t coff, int32_t nb, int64_t rows, int32_t c, void* scratch,
                   float* out, nc_stream_t stream) {
  return colsum_bf16(src, ld, coff, nb, rows, c, static_cast<float*>(scratch), out, S(stream));
}
int nc_cast_f16_bf16(const void* src, int32_t src_ld, int32_t src_coff, int64_t rows, int32_t c, void* dst,
                     int32_t dst_ld, int32_t dst_coff, nc_stream_t stream) {
  return cast_f16_bf16(src, src_ld, src_coff, rows, c, dst, dst_ld, dst_coff, S(stream));
}
int64_t nc_patchgan_ws_floats(int32_t n, int32_t h, int32_t w, int32_t ndf, int32_t n_layers) {
  return patchgan_ws_floats(n, h, w, ndf, n_layers);
}
int nc_patchgan_fwd(const float* x, int32_t n, int32_t h, int32_t w, int32_t ndf, int32_t n_layers,
                    const float* const* weights, const float* const* biases, float* ws, float* pred,
                    nc_stream_t stream) {
  return patchgan_fwd(x, n, h, w, ndf, n_layers, weights, biases, ws, pred, S(stream));
}
int nc_patchgan_bwd(const float* x, const float* dpred, int32_t n, int32_t h, int32_t w, int32_t ndf, int32_t n_layers,
                    const float* const* weights, float* ws, float* dx, float* const* dweights, float* const* dbiases,
                    nc_stream_t stream) {
  return patchgan_bwd(x, dpred, n, h, w, ndf, n_layers, weights, ws, dx, dweights, dbiases, S(stream));
}
int nc_adam_step_multi(const void* table, int32_t count, float lr, float beta1, float beta2, float eps, int32_t step,
                       nc_stream_t stream) {
  return adam_step_multi(table, count, lr, beta1, beta2, eps, step, S(stream));
}
int nc_augment_crop_u16(const uint16_t* vol, int32_t z, int32_t h, int32_t w, int32_t z0, int32_t cz, int32_t cy,
                        int32_t cx, const int32_t* x0, const int32_t* y0, const int32_t* adelta, const int32_t* bdelta,
                        int32_t flip_mask, float* out, nc_stream_t stream) {
  return augment_crop_u16(vol, z, h, w, z0, cz, cy, cx, x0, y0, adelta, bdelta, flip_mask, out, S(stream));
}
int nc_convT3d_k2s2_fwd(const void* x, const float* in_mean_rstd, int32_t nb, int32_t d, int32_t h, int32_t w,
                        int32_t cin, const void* packed, const float* bias, int32_t cout, void* y, int32_t y_ld,
                        int32_t y_coff, nc_stream_t stream) {
  return convT3d_k2s2_fwd(x, in_mean_rstd, nb, d, h, w, cin, packed, bias, cout, y, y_ld, y_coff, S(stream));
}

int64_t nc_in_stats_scratch_bytes(int32_t nb, int32_t c) { return static_cast<int64_t>(in_stats_scratch_bytes(nb, c)); }
int nc_in_stats_finalize(const float* partial, int32_t nb, int64_t rows, int32_t c, int64_t voxels, float eps,
                         void* scratch, float* mean_rstd, nc_stream_t stream) {
  return in_stats_finalize(partial, nb, rows, c, voxels, eps, scratch, mean_rstd, S(stream));
}

int nc_in_relu_apply(const void* raw, const float* mean_rstd, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t c,
                     void* y, int32_t y_ld, int32_t y_coff, void* pooled, nc_stream_t stream) {
  return in_relu_apply(raw, mean_rstd, nb, d, h, w, c, y, y_ld, y_coff, pooled, S(stream));
}

int nc_head_1x1_sigmoid_fwd(const void* raw, const float* mean_rstd, const float* hp, int32_t nb, int32_t d,
                            int32_t h, int32_t w, int32_t c, int32_t crop, float* y, nc_stream_t stream) {
  return head_1x1_sigmoid_fwd(raw, mean_rstd, hp, nb, d, h, w, c, crop, y, S(stream));
}

int nc_blend_gather_f32(const float* pieces, const int64_t* piece_off, const int32_t* piece_z0,
                        const int32_t padded_zyx[3], const int32_t steps_zyx[3], int32_t roi, int32_t overlap,
                        int32_t out_z0, int32_t out_nz, float* out, nc_stream_t stream) {
  return blend_gather_f32(pieces, reinterpret_cast<const long long*>(piece_off), piece_z0, padded_zyx, steps_zyx, roi,
                          overlap, out_z0, out_nz, out, S(stream));
}

int nc_select_init(const uint64_t ranks[4], void* st, nc_stream_t stream) {
  return select_init(reinterpret_cast<const unsigned long long*>(ranks), st, S(stream));
}
int nc_select_histogram(const float* data, int64_t n, int32_t pass, const void* st, uint64_t* hist,
                        nc_stream_t stream) {
  return select_histogram(data, n, pass, st, reinterpret_cast<unsigned long long*>(hist), S(stream));
}
int nc_select_update(int32_t pass, void* st, uint64_t* hist, nc_stream_t stream) {
  return select_update(pass, st, reinterpret_cast<unsigned long long*>(hist), S(stream));
}
int nc_percentile_lerp(const void* st, double t_lo, double t_hi, double* out64, float* out32, nc_stream_t stream) {
  return percentile_lerp(st, t_lo, t_hi, out64, out32, S(stream));
}

int nc_rescale_u16_crop(const float* vol, int32_t vol_z0, const int32_t padded_zyx[3], const int32_t size_zyx[3],
                        const float* norm3, int32_t z_begin, int32_t z_count, uint16_t* out, nc_stream_t stream) {
  return rescale_u16_crop(vol, vol_z0, padded_zyx, size_zyx, norm3, z_begin, z_count, out, S(stream));
}

int nc_rescale_u8_crop(const float* vol, int32_t vol_z0, const int32_t padded_zyx[3], const int32_t size_zyx[3],
                       const float* norm3, int32_t z_begin, int32_t z_count, uint8_t* out, nc_stream_t stream) {
  return rescale_u8_crop(vol, vol_z0, padded_zyx, size_zyx, norm3, z_begin, z_count, out, S(stream));
}

int nc_mip_fwd(const float* vol, int32_t d, int32_t h, int32_t w, int32_t axis, int32_t start, int32_t depth,
               float* proj, int32_t* argmax, nc_stream_t stream) {
  return mip_fwd(vol, d, h, w, axis, start, depth, proj, argmax, S(stream));
}
int nc_mip_bwd(const float* gproj, const int32_t* argmax, int32_t d, int32_t h, int32_t w, int32_t axis, float* gvol,
               nc_stream_t stream) {
  return mip_bwd(gproj, argmax, d, h, w, axis, gvol, S(stream));
}

int nc_conv2d_k4_fwd(const float* x, const float* w, const float* b, int32_t n, int32_t cin, int32_t h, int32_t wd,
                     int32_t cout, int32_t stride, float lrelu_slope, float* y, nc_stream_t stream) {
  return conv2d_k4_fwd(x, w, b, n, cin, h, wd, cout, stride, lrelu_slope, y, S(stream));
}
int nc_conv2d_k4_dgrad(const float* dy, const float* w, int32_t n, int32_t cin, int32_t h, int32_t wd, int32_t cout,
                       int32_t stride, float* dx, nc_stream_t stream) {
  return conv2d_k4_dgrad(dy, w, n, cin, h, wd, cout, stride, dx, S(stream));
}
int nc_conv2d_k4_wgrad(const float* x, const float* dy, int32_t n, int32_t cin, int32_t h, int32_t wd, int32_t cout,
                       int32_t stride, float* dw, float* db, nc_stream_t stream) {
  return conv2d_k4_wgrad(x, dy, n, cin, h, wd, cout, stride, dw, db, S(stream));
}
int nc_in2d_lrelu_fwd(const float* x, int32_t nc_planes, int32_t plane, float eps, float slope, float* y,
                      float* mean_rstd, nc_stream_t stream) {
  return in2d_lrelu_fwd(x, nc_planes, plane, eps, slope, y, mean_rstd, S(stream));
}
int nc_in2d_lrelu_bwd(const float* dy, const float* x, const float* mean_rstd, int32_t nc_planes, int32_t plane,
                      float slope, float* dx, nc_stream_t stream) {
  return in2d_lrelu_bwd(dy, x, mean_rstd, nc_planes, plane, slope, dx, S(stream));
}
int nc_lrelu_bwd(const float* dy, const float* y, int64_t n, float slope, float* dx, nc_stream_t stream) {
  return lrelu_bwd(dy, y, n, slope, dx, S(stream));
}
int nc_loss_fwd(const float* p, const float* q, float target, int64_t n, int32_t mode, float* loss,
                nc_stream_t stream) {
  return loss_fwd(p, q, target, n, mode, loss, S(stream));
}
int nc_loss_bwd(const float* p, const float* q, float target, int64_t n, int32_t mode, const float* upstream,
                float* dp, nc_stream_t stream) {
  return loss_bwd(p, q, target, n, mode, upstream, dp, S(stream));
}

int nc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 int32_t step, nc_stream_t stream) {
  return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, step, S(stream));
}

}  // extern "C"

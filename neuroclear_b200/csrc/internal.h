// Internal (non-ABI) declarations shared by the translation units of libneuroclear_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace nc {

// Thread-local last-error string behind nc_last_error(); returns -1 so callers can `return set_error(...)`.
int set_error(const char* fmt, ...);

#define NC_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) return ::nc::set_error("%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

int num_sms();
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property: true the first time `flags` (a static
// 64-entry array owned by the call site) is asked about the current device.
bool first_use_on_device(bool* flags);

using TensorMapEncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                            const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                            CUtensorMapFloatOOBfill);
// Resolved through cudaGetDriverEntryPoint so the library has no link-time dependency on libcuda.so.
TensorMapEncodeTiledFn get_tensor_map_encoder();

// conv3d_tc.cu
void debug_set_max_ctas(int n);
void debug_set_remainder_pairs(int on);
int conv3d_k3_bn(int Cout);
int conv3d_k3_td(int Cout);
size_t conv3d_k3_stats_tiles(int NB, int D, int H, int W, int Cout);
int conv3d_k3_fwd(const void* x, const float* in_mean_rstd, int NB, int D, int H, int W, int Cin, const void* wpacked,
                  int Cout, void* y_raw, float* stats_partial, cudaStream_t stream);
int convT3d_k2s2_fwd(const void* x, const float* in_mean_rstd, int NB, int D, int H, int W, int Cin,
                     const void* wpacked, const float* bias, int Cout, void* y, int y_ld, int y_coff,
                     cudaStream_t stream);
int conv3d_k3_dgrad(const void* dy, int NB, int D, int H, int W, int Cout, const void* wpacked, int Cin, void* dx,
                    cudaStream_t stream);
int pack_weights_dgrad(const float* w, void* out, int Cout, int Cin, cudaStream_t stream);
int conv3d_k1_bf16(const void* x, int NB, int D, int H, int W, int K, const void* wpacked, int N, void* y,
                   cudaStream_t stream);
int pack_weights_convT_dgrad(const float* w, void* out, int Cin, int Cout, cudaStream_t stream);
// unet_bwd.cu
int bwd_blocks();
int in_relu_bwd(const void* raw, const float* mean_rstd, int NB, int D, int H, int W, int C, int mode,
                const void* grad, int grad_ld, int grad_coff, const float* du, const float* w_head, const void* dpool,
                float* scratch, float* m12, void* d_raw, cudaStream_t stream);
int head_bwd(const void* raw, const float* mean_rstd, const float* hp, const float* dout, int NB, int D, int H, int W,
             float* du, float* scratch, float* grads, cudaStream_t stream);
int conv1_wgrad(const float* x, const void* dy, int dy_fmt, int NB, int D, int H, int W, float* scratch, float* dw,
                cudaStream_t stream);
// deeplinear.cu
int im2col49(const float* x, int NB, int D, int H, int W, int fmt, void* out, cudaStream_t stream);
int col2im49(const void* g, int NB, int D, int H, int W, float* dx, cudaStream_t stream);
int stencil64to1_fwd(const void* h, const float* K, int NB, int D, int H, int W, float* out, cudaStream_t stream);
int stencil64to1_bwd_data(const float* dout, const float* K, int NB, int D, int H, int W, void* dh,
                          cudaStream_t stream);
int conv3d_tc_64(const void* x, int fmt, int NB, int D, int H, int W, const void* wpacked, int ksd, int ksp, void* y,
                 cudaStream_t stream);
int pack_weights_64(const float* w, void* out, int taps, int dgrad, cudaStream_t stream);
int space_to_depth_bf16(const void* src, int ld, int coff, int NB, int D, int H, int W, int C, void* out,
                        cudaStream_t stream);
int colsum_bf16(const void* src, int ld, int coff, int NB, long long rows, int C, float* scratch, float* out,
                cudaStream_t stream);
int cast_f16_bf16(const void* src, int src_ld, int src_coff, long long rows, int C, void* dst, int dst_ld,
                  int dst_coff, cudaStream_t stream);
int in_relu_apply_bf16(const void* raw, const float* mean_rstd, int NB, int D, int H, int W, int C, void* y, int y_ld,
                       int y_coff, void* pooled, cudaStream_t stream);
// wgrad3d_tc.cu
size_t conv3d_wgrad_scratch_bytes(int ks, int NB, int D, int H, int W, int Cin, int Cout);
int conv3d_wgrad(const void* x, int x_fmt, const void* dy, int dy_fmt, int NB, int D, int H, int W, int Cin,
                 int Cout, int ks, void* scratch, float* dw, cudaStream_t stream);
size_t packed_weight_bytes(int Cout, int Cin, int taps, int transposed);
int pack_weights(const float* w, void* out, int Cout, int Cin, int taps, int transposed, cudaStream_t stream);


// elementwise.cu
int dice_extract_u16(const uint16_t* vol, int vz0, int vnz, const int* size, const int* padded, const int* steps,
                     int roi, int overlap, int border, long long cube_begin, int cube_count, float* cubes,
                     cudaStream_t stream);
int dice_extract_u8(const uint8_t* vol, int vz0, int vnz, const int* size, const int* padded, const int* steps,
                    int roi, int overlap, int border, long long cube_begin, int cube_count, float* cubes,
                    cudaStream_t stream);
int rescale_u8_crop(const float* vol, int vol_z0, const int* padded, const int* size, const float* norm3, int z_begin,
                    int z_count, uint8_t* out, cudaStream_t stream);
size_t conv_cin1_stats_tiles(int NB, int D, int H, int W);
// conv1_tc.cu
int pack_conv1_weights(const float* w, void* packed, cudaStream_t stream);
int conv3d_cin1_k3_fwd(const float* x, const void* wpacked, int NB, int D, int H, int W, int Cout, void* y_raw,
                       float* stats_partial, cudaStream_t stream);
size_t in_stats_scratch_bytes(int NB, int C);
int in_stats_finalize(const float* partial, int NB, long long rows, int C, long long voxels, float eps, void* scratch,
                      float* mean_rstd, cudaStream_t stream);
int in_relu_apply(const void* raw, const float* mean_rstd, int NB, int D, int H, int W, int C, void* y, int y_ld,
                  int y_coff, void* pooled, cudaStream_t stream);
int head_1x1_sigmoid_fwd(const void* raw, const float* mean_rstd, const float* hp, int NB, int D, int H, int W,
                         int C, int crop, float* y, cudaStream_t stream);
int blend_gather_f32(const float* pieces, const long long* piece_off, const int* piece_z0, const int* padded,
                     const int* steps, int roi, int overlap, int out_z0, int out_nz, float* out, cudaStream_t stream);
int select_init(const unsigned long long* ranks, void* st, cudaStream_t stream);
int select_histogram(const float* data, long long n, int pass, const void* st, unsigned long long* hist,
                     cudaStream_t stream);
int select_update(int pass, void* st, unsigned long long* hist, cudaStream_t stream);
int percentile_lerp(const void* st, double t_lo, double t_hi, double* out64, float* out32, cudaStream_t stream);
int rescale_u16_crop(const float* vol, int vol_z0, const int* padded, const int* size, const float* norm3,
                     int z_begin, int z_count, uint16_t* out, cudaStream_t stream);
int mip_fwd(const float* vol, int D, int H, int W, int axis, int start, int depth, float* proj, int* argmax,
            cudaStream_t stream);
int mip_bwd(const float* gproj, const int* argmax, int D, int H, int W, int axis, float* gvol, cudaStream_t stream);

// disc2d.cu
void debug_set_disc_cluster(int c);
int conv2d_k4_fwd(const float* x, const float* w, const float* b, int N, int Cin, int H, int W, int Cout, int stride,
                  float slope, float* y, cudaStream_t stream);
int conv2d_k4_dgrad(const float* dy, const float* w, int N, int Cin, int H, int W, int Cout, int stride, float* dx,
                    cudaStream_t stream);
int conv2d_k4_wgrad(const float* x, const float* dy, int N, int Cin, int H, int W, int Cout, int stride, float* dw,
                    float* db, cudaStream_t stream);
int in2d_lrelu_fwd(const float* x, int NC, int P, float eps, float slope, float* y, float* mean_rstd,
                   cudaStream_t stream);
int in2d_lrelu_bwd(const float* dy, const float* x, const float* mean_rstd, int NC, int P, float slope, float* dx,
                   cudaStream_t stream);
int lrelu_bwd(const float* dy, const float* y, long long n, float slope, float* dx, cudaStream_t stream);
int loss_fwd(const float* p, const float* q, float target, long long n, int mode, float* loss, cudaStream_t stream);
int loss_bwd(const float* p, const float* q, float target, long long n, int mode, const float* upstream, float* dp,
             cudaStream_t stream);

// augment.cu
int augment_crop_u16(const uint16_t* vol, int Z, int H, int W, int z0, int cz, int cy, int cx, const int* x0,
                     const int* y0, const int* adelta, const int* bdelta, int flip_mask, float* out,
                     cudaStream_t stream);
// patchgan.cu
long long patchgan_ws_floats(int N, int H, int W, int ndf, int n_layers);
int patchgan_fwd(const float* x, int N, int H, int W, int ndf, int n_layers, const float* const* weights,
                 const float* const* biases, float* ws, float* pred, cudaStream_t stream);
int patchgan_bwd(const float* x, const float* dpred, int N, int H, int W, int ndf, int n_layers,
                 const float* const* weights, float* ws, float* dx, float* const* dweights, float* const* dbiases,
                 cudaStream_t stream);
int adam_step_multi(const void* table, int count, float lr, float beta1, float beta2, float eps, int step,
                    cudaStream_t stream);

int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
              int step, cudaStream_t stream);

const char* last_error();

}  // namespace nc

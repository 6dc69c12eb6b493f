/* neuroclear_b200 — C ABI of the B200 (sm_100a) hot path of peterhpark/neuroclear.
 *
 * The reference is pure Python (no FFI of its own), so this header IS the drop-in boundary: the Python shims
 * in neuroclear_b200/ (same class / function names as the reference: networks.define_G('unet_deconv'),
 * DiceImageDataSet, Assemble_Dice, Volume.get_projection) bind exactly these symbols through ctypes.
 * Each entry point cites the reference code it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - plain C types only; every pointer marked "device" is caller-owned CUDA device memory (e.g. a torch
 *     tensor's data_ptr()); nothing here allocates persistent device memory.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, one host thread per
 *     device; return value 0 = OK, negative = error with the message in nc_last_error() (thread-local).
 *   - activations between the tensor-core convolutions are fp16, NDHWC ("channels-last-3d"); convolution
 *     outputs that feed InstanceNorm are raw fp16 NDHWC plus per-tile fp32 statistics partials.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with an error.
 */
#ifndef NEUROCLEAR_B200_H_
#define NEUROCLEAR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* nc_stream_t; /* cudaStream_t */

#define NC_ABI_VERSION 1

/* ---- runtime ------------------------------------------------------------------------------------------- */
int nc_abi_version(void);
const char* nc_last_error(void);
/* "NC_SOURCE_HASH=<sha256>": digest of the csrc/ sources, headers, include/neuroclear_b200.h and the nvcc flags this
 * binary was compiled from (neuroclear_b200/build.py passes it as -DNC_SOURCE_HASH); build() recompiles when the
 * digest of the tree differs, so a stale shipped .so cannot pass for the source */
const char* nc_build_source_hash(void);
/* Asynchronous pitched host -> device copy (cudaMemcpy2DAsync; src should be pinned): `rows` segments of width_bytes,
 * pitch_bytes apart in BOTH buffers.  Used to upload a y-range of every z-plane of a volume slab in one call, so the
 * first cube rows start before the rest of the slab has crossed PCIe (pipeline.ChunkedUpload). */
int nc_memcpy2d_h2d_async(void* dst, const void* src, int64_t pitch_bytes, int64_t width_bytes, int64_t rows,
                          nc_stream_t stream);
/* number of SMs of the current device (148 on B200); <0 on error */
int nc_device_sm_count(void);
/* test hook: cap the persistent grid of the tensor-core conv kernels at n CTAs (0 = one per SM) so that small test
 * problems also exercise the several-tiles-per-CTA path (ring wrap-around, accumulator ping-pong) */
void nc_debug_set_max_ctas(int32_t n);
/* test / measurement hook: 0 disables the remainder-pair kernel of the N-tile-128 k3 convolutions (the last, partly
 * filled 16-line h-tile is then computed by the regular kernel); the stored values are identical either way */
void nc_debug_set_remainder_pairs(int32_t on);
/* test / measurement hook: 1, 2, 4 or 8 force the thread-block-cluster size (= split of the reduction dimension) of
 * the PatchGAN convolutions nc_conv2d_k4_*; 0 restores the cost model.  Results agree up to fp32 summation order. */
void nc_debug_set_disc_cluster(int32_t c);

/* ---- dicing geometry (host-only integer math) --------------------------------------------------------------
 * util/util.py:196-215 pad_for_dicing  +  data/diceImage_dataset.py:82-106 DiceCube.__init__/indexToCoordinates
 * +  util/assemble_dice.py:21-25,60-77.  size/padded/steps are (z, y, x).  Returns the number of cubes. */
int64_t nc_dice_geometry(const int32_t size_zyx[3], int32_t roi, int32_t overlap, int32_t padded_zyx[3],
                         int32_t steps_zyx[3]);

/* ---- dice extraction ------------------------------------------------------------------------------------------
 * data/diceImage_dataset.py:95-96,108-120 (reflect pad by border + cube slice) fused with
 * data/base_dataset.py:134-143,291-301 (uint16 / 65535 -> float32, add channel) and the zero far-end padding of
 * util/util.py:212.  `vol` holds original-volume planes [vol_z0, vol_z0+vol_nz) of a (Z,Y,X) uint16 volume.
 * Writes cubes [cube_begin, cube_begin+cube_count) as float32 (count, E, E, E), E = roi + 2*border. */
int nc_dice_extract_u16(const uint16_t* vol /*device*/, int32_t vol_z0, int32_t vol_nz, const int32_t size_zyx[3],
                        const int32_t padded_zyx[3], const int32_t steps_zyx[3], int32_t roi, int32_t overlap,
                        int32_t border, int64_t cube_begin, int32_t cube_count, float* cubes /*device*/,
                        nc_stream_t stream);

/* same for --data_type uint8 volumes (divide by 255, base_dataset.py:135-136) */
int nc_dice_extract_u8(const uint8_t* vol /*device*/, int32_t vol_z0, int32_t vol_nz, const int32_t size_zyx[3],
                       const int32_t padded_zyx[3], const int32_t steps_zyx[3], int32_t roi, int32_t overlap,
                       int32_t border, int64_t cube_begin, int32_t cube_count, float* cubes /*device*/,
                       nc_stream_t stream);

/* ---- Unet_deconv layers (models/networks.py:478-538) ------------------------------------------------------ */

/* Number of statistics-partial rows a conv writes for a (NB,D,H,W) output with Cout channels; the partial
 * buffer is float32 [rows][2][Cout] (sum, sum of squares per tile). `cin` selects the kernel (1 = first layer). */
int64_t nc_conv3d_k3_stats_rows(int32_t cin, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t cout);

/* double_conv1.convolution.0 (networks.py:420,490): Conv3d(1 -> 64, k3 s1 p1) on tcgen05 with an in-kernel im2col:
 * every fp32 input is split into fp16 hi + lo parts (22 significant bits), K = 27 + 27 (+10 zero) = 64.
 * nc_pack_weights_conv3d_cin1_k3: w float32 (64,1,3,3,3) as in the state_dict -> packed (8192 bytes, device).
 * x: float32 (NB,D,H,W); y_raw: fp16 NDHWC (NB,D,H,W,64) WITHOUT bias (a bias in front of
 * InstanceNorm(affine=False) cancels); statistics partials come from the fp32 accumulators. */
int nc_pack_weights_conv3d_cin1_k3(const float* w, void* packed, nc_stream_t stream);
int nc_conv3d_cin1_k3_fwd(const float* x, const void* packed, int32_t nb, int32_t d, int32_t h, int32_t wdt,
                          int32_t cout, void* y_raw, float* stats_partial, nc_stream_t stream);

/* Packed weight sizes / packing for the tensor-core kernels.  conv: w is OIDHW float32 (Cout,Cin,3,3,3);
 * convT: w is IODHW float32 (Cin,Cout,2,2,2) (torch ConvTranspose3d layout).  Output is the fp16 swizzled
 * shared-memory image the kernels stream with bulk copies. */
int64_t nc_packed_weight_bytes(int32_t cout, int32_t cin, int32_t transposed);
int nc_pack_weights_conv3d_k3(const float* w_oidhw, int32_t cout, int32_t cin, void* packed, nc_stream_t stream);
int nc_pack_weights_convT3d_k2s2(const float* w_iodhw, int32_t cin, int32_t cout, void* packed, nc_stream_t stream);

/* nn.Conv3d(Cin -> Cout, k3 s1 p1) of double_conv / triple_conv / last_conv (networks.py:413-476): tcgen05
 * implicit GEMM, fp16 operands, fp32 accumulate.  x: fp16 NDHWC (NB,D,H,W,Cin) (Cin may be a concat buffer).
 * in_mean_rstd == NULL: x is the conv's input as is (already normalised activations).
 * in_mean_rstd != NULL (float32 (NB,2,Cin) from nc_in_stats_finalize): x is the RAW output of the previous conv
 *   and the InstanceNorm + ReLU between the two layers (networks.py:422-423 etc.) is applied inside this kernel
 *   while the input planes sit in shared memory — no separate nc_in_relu_apply pass.
 * y_raw: fp16 NDHWC (NB,D,H,W,Cout) without bias; stats_partial (from the fp32 accumulators) as above.
 * Cin % 64 == 0, Cout in {64,128k}. */
int nc_conv3d_k3_fwd(const void* x_f16, const float* in_mean_rstd, int32_t nb, int32_t d, int32_t h, int32_t w,
                     int32_t cin, const void* packed, int32_t cout, void* y_raw, float* stats_partial,
                     nc_stream_t stream);

/* ---- training: gradients of the same Conv3d layers (what autograd computes for networks.py:413-538 when
 * train_onecube.py / axial_to_lateral_gan_apollo_model.py:255-295 call backward) -------------------------------------
 * Gradient tensors are bf16 NDHWC (fp16 would underflow: d loss / d voxel ~ 1e-6), accumulation is fp32.
 *
 * nc_conv3d_k3_dgrad: dx = conv(dy, filter transposed over channels and flipped in space) on the forward tcgen05
 *   kernel.  dy: bf16 (NB,D,H,W,Cout); packed: nc_pack_weights_conv3d_k3_dgrad image (bf16, same size as the forward
 *   image); dx: bf16 (NB,D,H,W,Cin).  Cout % 64 == 0, Cin in {64,128k}.
 * nc_conv3d_wgrad: dW[co][ci][kd][kh][kw] = sum_v dy[v][co] * x[v + tap - pad][ci], tcgen05 GEMM whose K dimension
 *   is the voxel (both operands read MN-major from TMA-staged halo planes), split-K over the CTAs with a fixed-order
 *   (deterministic) reduction.  x: 16-bit NDHWC (NB,D,H,W,Cin), x_fmt 0 = fp16 / 1 = bf16 (the forward activations
 *   are fp16); dy likewise (dy_fmt).  ks in {3,5} (Unet_deconv k3; DeepLinearGenerator k5 / k3, networks.py:899-905).
 *   ks = 71 selects the 7 x 1 x 1 (depth-only) filter of the im2col'ed k7 layer, ks = 1 the 1x1x1 GEMM.
 *   scratch: nc_conv3d_wgrad_scratch_bytes(...) bytes; dw: float32 OIDHW (Cout,Cin,ks,ks,ks), overwritten. */
int nc_pack_weights_conv3d_k3_dgrad(const float* w_oidhw, int32_t cout, int32_t cin, void* packed,
                                    nc_stream_t stream);
int nc_conv3d_k3_dgrad(const void* dy_bf16, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t cout,
                       const void* packed, int32_t cin, void* dx_bf16, nc_stream_t stream);
int64_t nc_conv3d_wgrad_scratch_bytes(int32_t ks, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t cin,
                                      int32_t cout);
int nc_conv3d_wgrad(const void* x, int32_t x_fmt, const void* dy, int32_t dy_fmt, int32_t nb, int32_t d, int32_t h,
                    int32_t w, int32_t cin, int32_t cout, int32_t ks, void* scratch, float* dw, nc_stream_t stream);

/* Backward of ConvTranspose3d(k2, s2) (networks.py:500,503) as two plain GEMMs over a space-to-depth gather of the
 * output gradient: g[coarse voxel][tap * Cout + co] = dy[2 * voxel + tap][y_coff + co], tap = (a*2+b)*2+c.
 *   data gradient:   dx = nc_conv3d_k1_bf16(g, K = 8*Cout, packed = nc_pack_weights_convT3d_k2s2_dgrad image, N = Cin)
 *   weight gradient: nc_conv3d_wgrad(x, g, ks = 1, cin = Cin, cout = 8*Cout) -> (8*Cout, Cin) = dW[ci][co][tap]^T
 *   bias gradient:   nc_colsum_bf16 over dy. */
int nc_space_to_depth_bf16(const void* src_bf16, int32_t ld, int32_t coff, int32_t nb, int32_t d_coarse,
                           int32_t h_coarse, int32_t w_coarse, int32_t c, void* out_bf16, nc_stream_t stream);
int nc_pack_weights_convT3d_k2s2_dgrad(const float* w_iodhw, int32_t cin, int32_t cout, void* packed,
                                       nc_stream_t stream);
int nc_conv3d_k1_bf16(const void* x_bf16, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t k, const void* packed,
                      int32_t n, void* y_bf16, nc_stream_t stream);
/* out[c] = sum over all nb * rows rows of src[row][coff + c] (bf16 in, fixed-order fp64 reduction). */
int nc_colsum_bf16(const void* src_bf16, int32_t ld, int32_t coff, int32_t nb, int64_t rows, int32_t c, void* scratch,
                   float* out, nc_stream_t stream);
/* fp16 -> bf16 copy of a channel slice of an NDHWC tensor (rows voxels). */
int nc_cast_f16_bf16(const void* src_f16, int32_t src_ld, int32_t src_coff, int64_t rows, int32_t c, void* dst_bf16,
                     int32_t dst_ld, int32_t dst_coff, nc_stream_t stream);

/* Scratch for the reductions of the backward kernels below (one buffer serves them all). */
int64_t nc_bwd_scratch_bytes(int32_t nb);
/* nc_in_relu_apply with bf16 output: the weight-gradient GEMM reads activations in the gradient's format. */
int nc_in_relu_apply_bf16(const void* raw, const float* mean_rstd, int32_t nb, int32_t d, int32_t h, int32_t w,
                          int32_t c, void* y_bf16, int32_t y_ld, int32_t y_coff, void* pooled_bf16,
                          nc_stream_t stream);
/* Backward of InstanceNorm3d(affine=False) + ReLU (networks.py:422-423 etc.) of one layer:
 *   yhat = (raw - mean) * rstd, g = dA * [yhat > 0], d_raw = rstd * (g - mean_v(g) - yhat * mean_v(g * yhat)).
 * dA (gradient w.r.t. the activation) comes from, by `mode`:
 *   0: the bf16 tensor `grad` (row pitch grad_ld, channel offset grad_coff);
 *   1: the head: dA[v][c] = du[v] * w_head[c] (nc_head_1x1_sigmoid_bwd);
 *   2: skip + pool: grad slice as in 0 PLUS the MaxPool3d(2) routing of `dpool` (bf16 (NB,D/2,H/2,W/2,C)) to the
 *      first maximum of every 2x2x2 window — the gradient of `torch.cat([conv, up])` + `maxpool(conv)`.
 * m12: float32 (NB,2,C) work buffer (the two means); d_raw: bf16 (NB,D,H,W,C). */
int nc_in_relu_bwd(const void* raw_f16, const float* mean_rstd, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t c,
                   int32_t mode, const void* grad_bf16, int32_t grad_ld, int32_t grad_coff, const float* du,
                   const float* w_head, const void* dpool_bf16, void* scratch, float* m12, void* d_raw_bf16,
                   nc_stream_t stream);
/* Backward of the head (networks.py:507-510,536): given dout = dL/d sigmoid output (float32 (NB,D,H,W)):
 * du = dout * out * (1 - out) * w2 (float32 per voxel) and grads (68 floats) = [d one_by_one.weight (64) |
 * d one_by_one.bias | d one_by_one_2.weight | d one_by_one_2.bias | 0], summed over the samples. */
int nc_head_1x1_sigmoid_bwd(const void* raw_f16, const float* mean_rstd, const float* head_params, const float* dout,
                            int32_t nb, int32_t d, int32_t h, int32_t w, float* du, void* scratch, float* grads,
                            nc_stream_t stream);
/* Weight gradient of the Cin = 1 first conv (networks.py:420): dw float32 (64, 27). x float32 (NB,D,H,W),
 * dy bf16 (NB,D,H,W,64). */
int nc_conv3d_cin1_k3_wgrad(const float* x, const void* dy, int32_t dy_fmt /* 0 fp16, 1 bf16 */, int32_t nb,
                            int32_t d, int32_t h, int32_t w, void* scratch, float* dw, nc_stream_t stream);

/* ---- DeepLinearGenerator (networks.py:893-917; G_B of axial_to_lateral_gan_apollo): k7 1->64, k5 64->64, k3 64->64,
 * three 1x1 (64->32->16->1), no bias, no activation.
 *   k7 layer  = nc_im2col49 (the 7x7 in-plane neighbourhood as 49 (+15 zero) channels) + nc_conv3d_tc_64(ksd 7, ksp 1);
 *               backward: nc_conv3d_wgrad(ks = 71), nc_conv3d_tc_64(fmt 1, flipped filter) + nc_col2im49.
 *   k5 layer  = nc_conv3d_tc_64(ksd 5, ksp 5); backward: nc_conv3d_wgrad(ks = 5), nc_conv3d_tc_64(fmt 1).
 *   k3 + 1x1s = folded exactly into ONE 64->1 k3 stencil K[ci][tap] = sum_co (W6 W5 W4)[co] W3[co][ci][tap]:
 *               nc_stencil64to1_fwd; backward: nc_stencil64to1_bwd_data (dh, bf16) and, for dK,
 *               nc_conv3d_cin1_k3_wgrad(x = dout, dy = h) with the taps reversed.
 * nc_pack_weights_64: w float32 (64, 64, taps) -> packed image (fp16; dgrad != 0: channel-transposed, taps reversed,
 * bf16), 64*64*taps*2 bytes.  nc_conv3d_tc_64: x, y 16-bit NDHWC (NB,D,H,W,64); fmt 0 = fp16, 1 = bf16. */
int nc_im2col49(const float* x, int32_t nb, int32_t d, int32_t h, int32_t w, int32_t fmt, void* out, nc_stream_t stream);
int nc_col2im49(const void* g_bf16, int32_t nb, int32_t d, int32_t h, int32_t w, float* dx, nc_stream_t stream);
int nc_pack_weights_64(const float* w, int32_t taps, int32_t dgrad, void* packed, nc_stream_t stream);
int nc_conv3d_tc_64(const void* x, int32_t fmt, int32_t nb, int32_t d, int32_t h, int32_t w, const void* packed,
                    int32_t ksd, int32_t ksp, void* y, nc_stream_t stream);
int nc_stencil64to1_fwd(const void* h_f16, const float* k, int32_t nb, int32_t d, int32_t h, int32_t w, float* out,
                        nc_stream_t stream);
int nc_stencil64to1_bwd_data(const float* dout, const float* k, int32_t nb, int32_t d, int32_t h, int32_t w,
                             void* dh_bf16, nc_stream_t stream);

/* nn.ConvTranspose3d(Cin -> Cout, k2 s2) t_conv2 / t_conv1 (networks.py:500,503) fused with the channel concat
 * torch.cat([skip, up], 1) (networks.py:526,531): GEMM + pixel-shuffle scatter + bias, written as fp16 into
 * channels [y_coff, y_coff+Cout) of an NDHWC buffer (NB,2D,2H,2W,y_ld). */
int nc_convT3d_k2s2_fwd(const void* x_f16, const float* in_mean_rstd /* as for nc_conv3d_k3_fwd */, int32_t nb,
                        int32_t d, int32_t h, int32_t w, int32_t cin, const void* packed, const float* bias,
                        int32_t cout, void* y_f16, int32_t y_ld, int32_t y_coff, nc_stream_t stream);

/* InstanceNorm3d(affine=False, eps) statistics (networks.py:33-34): deterministic fixed-order two-level reduction
 * of the per-tile partials in fp64 -> mean_rstd float32 (NB, 2, C): [:,0,:] = mean, [:,1,:] = 1/sqrt(var_biased+eps).
 * scratch: device buffer of nc_in_stats_scratch_bytes(nb, c) bytes, ZERO-FILLED once by the caller before its first
 * use (it holds arrival counters the kernel resets itself) and reusable across calls on one stream. */
int64_t nc_in_stats_scratch_bytes(int32_t nb, int32_t c);
int nc_in_stats_finalize(const float* stats_partial, int32_t nb, int64_t rows_per_sample, int32_t c,
                         int64_t voxels_per_sample, float eps, void* scratch, float* mean_rstd, nc_stream_t stream);

/* InstanceNorm apply + ReLU (+ MaxPool3d(2), networks.py:491,494) + write into a concat slice:
 * y[..., y_coff:y_coff+C] = fp16(relu((raw - mean) * rstd)); if pooled != NULL also the 2x2x2 max as fp16
 * NDHWC (NB,D/2,H/2,W/2,C).  raw: fp16 NDHWC (normalised in fp32). */
int nc_in_relu_apply(const void* raw, const float* mean_rstd, int32_t nb, int32_t d, int32_t h, int32_t w,
                     int32_t c, void* y_f16, int32_t y_ld, int32_t y_coff, void* pooled_f16, nc_stream_t stream);

/* Tail of Unet_deconv.forward (networks.py:504-510,533-536): InstanceNorm+ReLU of ex_conv1_1, one_by_one (C->1),
 * one_by_one_2 (1->1) and sigmoid in one pass; optionally drops `crop` voxels per side (the border cut of
 * util/assemble_dice.py:143).  head_params (device float32): [w1[0..C), b1, w2, b2].
 * raw: fp16 NDHWC (NB,D,H,W,C); y: float32 (NB, D-2crop, H-2crop, W-2crop). */
int nc_head_1x1_sigmoid_fwd(const void* raw, const float* mean_rstd, const float* head_params, int32_t nb,
                            int32_t d, int32_t h, int32_t w, int32_t c, int32_t crop, float* y, nc_stream_t stream);

/* ---- assembly (util/assemble_dice.py:161-213) ---------------------------------------------------------------
 * Overlap-add blend in the reference's exact fp32 order: for every padded-volume voxel p in planes
 * [out_z0, out_z0+out_nz): acc = 0; for cubes covering p in ascending cube index: acc += cube[p-origin] * 0.125f;
 * out = (acc / n(p)) * 8 with n(p) the analytic overlap count (assemble_dice.py:167-184).
 * Cube outputs are given as "pieces": piece_off[i] = float offset into `pieces` of cube i's stored planes
 * (-1: cube absent), piece_z0[i] = cube-local z of the first stored plane; each plane is roi*roi floats. */
int nc_blend_gather_f32(const float* pieces, const int64_t* piece_off, const int32_t* piece_z0,
                        const int32_t padded_zyx[3], const int32_t steps_zyx[3], int32_t roi, int32_t overlap,
                        int32_t out_z0, int32_t out_nz, float* out, nc_stream_t stream);

/* Exact order statistics for np.percentile (assemble_dice.py:191) by 3-pass radix select over the fp32 bit
 * pattern (values must be >= 0).  State lives on the device:  sel_state = uint64 rank[4], uint32 prefix[4].
 *   nc_select_init      : set the (up to 4) zero-based target ranks
 *   nc_select_histogram : pass in {0,1,2}; adds to hist (uint64 [4][4096]) the counts of the pass's digit for
 *                         elements matching each target's current prefix (all-reduce hist across GPUs between
 *                         the two calls when the volume is sharded)
 *   nc_select_update    : consume hist, narrow prefix/rank; after pass 2 prefix[t] is the value's bit pattern
 *   nc_percentile_lerp  : numpy's linear-method lerp in fp64 for two percentiles from the four order
 *                         statistics: out_f64[2] = (p_lo, p_hi), out_f32[3] = (f32(p_lo), f32(p_hi), f32(p_hi - p_lo)) */
int nc_select_init(const uint64_t ranks[4], void* sel_state, nc_stream_t stream);
int nc_select_histogram(const float* data, int64_t n, int32_t pass, const void* sel_state, uint64_t* hist,
                        nc_stream_t stream);
int nc_select_update(int32_t pass, void* sel_state, uint64_t* hist, nc_stream_t stream);
int nc_percentile_lerp(const void* sel_state, double t_lo, double t_hi, double* out_f64, float* out_f32,
                       nc_stream_t stream);

/* skimage.exposure.rescale_intensity(in_range=(p1,p99)) + *65535 + astype(uint16) + un-pad crop
 * (assemble_dice.py:190-213).  vol: float32 padded planes [vol_z0, ...); writes uint16 planes
 * [z_begin, z_begin+z_count) of the ORIGINAL (Z,Y,X) volume to `out` (count, Y, X).
 * norm3 (device float32[3]) = (f32(p1), f32(p99), f32(p99-p1)) as written by nc_percentile_lerp;
 * NULL => no intensity normalisation. */
int nc_rescale_u16_crop(const float* vol, int32_t vol_z0, const int32_t padded_zyx[3], const int32_t size_zyx[3],
                        const float* norm3, int32_t z_begin, int32_t z_count, uint16_t* out, nc_stream_t stream);

/* --data_type uint8: *255 + astype(uint8) (assemble_dice.py:195-200) */
int nc_rescale_u8_crop(const float* vol, int32_t vol_z0, const int32_t padded_zyx[3], const int32_t size_zyx[3],
                       const float* norm3, int32_t z_begin, int32_t z_count, uint8_t* out, nc_stream_t stream);

/* ---- randomised-depth max-intensity projection (models/axial_to_lateral_gan_apollo_model.py:339-351) ------
 * vol: float32 (D,H,W) single-channel cube; projects planes [start, start+depth) along `axis` (0=z,1=y,2=x)
 * with max; fwd also records the arg-max plane (int32) used by bwd to route the gradient
 * (torch.max(dim)[0] autograd semantics: first maximal index). */
int nc_mip_fwd(const float* vol, int32_t d, int32_t h, int32_t w, int32_t axis, int32_t start, int32_t depth,
               float* proj, int32_t* argmax, nc_stream_t stream);
int nc_mip_bwd(const float* grad_proj, const int32_t* argmax, int32_t d, int32_t h, int32_t w, int32_t axis,
               float* grad_vol /* accumulated into */, nc_stream_t stream);

/* ---- 2-D PatchGAN discriminator path of the apollo model (training) ------------------------------------------
 * NLayerDiscriminator(dimension=2) (models/networks.py:1009-1067, built at
 * models/axial_to_lateral_gan_apollo_model.py:99-123): Conv2d(k4, p1, stride 1|2) forward / data gradient / weight
 * (+bias) gradient, InstanceNorm2d(affine=False)+LeakyReLU forward / backward, LeakyReLU backward, and the two
 * losses of the step (GANLoss 'lsgan' networks.py:252-319 = MSE against a constant; torch.nn.L1Loss
 * apollo_model.py:128).  fp32, NCHW, latency-bound by nature (one ~108^2 image per pass).
 * x (N,Cin,H,W), w (Cout,Cin,4,4) and b (Cout) as in the state_dict, y (N,Cout,Ho,Wo), Ho = (H-2)/stride + 1.
 * lrelu_slope = 1 -> no activation; 0.2 fuses the LeakyReLU(0.2) that follows the first conv. */
int nc_conv2d_k4_fwd(const float* x, const float* w, const float* b /*nullable*/, int32_t n, int32_t cin, int32_t h,
                     int32_t wd, int32_t cout, int32_t stride, float lrelu_slope, float* y, nc_stream_t stream);
int nc_conv2d_k4_dgrad(const float* dy, const float* w, int32_t n, int32_t cin, int32_t h, int32_t wd, int32_t cout,
                       int32_t stride, float* dx, nc_stream_t stream);
int nc_conv2d_k4_wgrad(const float* x, const float* dy, int32_t n, int32_t cin, int32_t h, int32_t wd, int32_t cout,
                       int32_t stride, float* dw, float* db /*nullable*/, nc_stream_t stream);
/* x, y, dy, dx: (nc_planes, plane) = (N*C, H*W); mean_rstd: (nc_planes, 2) saved by fwd for bwd; eps = 1e-5 */
int nc_in2d_lrelu_fwd(const float* x, int32_t nc_planes, int32_t plane, float eps, float slope, float* y,
                      float* mean_rstd, nc_stream_t stream);
int nc_in2d_lrelu_bwd(const float* dy, const float* x, const float* mean_rstd, int32_t nc_planes, int32_t plane,
                      float slope, float* dx, nc_stream_t stream);
/* dx = dy * (y > 0 ? 1 : slope) with y the OUTPUT of the (fused) LeakyReLU */
int nc_lrelu_bwd(const float* dy, const float* y, int64_t n, float slope, float* dx, nc_stream_t stream);
/* mode 0: loss = mean((p - target)^2) (q unused); mode 1: loss = mean(|p - q|).  bwd: dp = *upstream * dloss/dp */
int nc_loss_fwd(const float* p, const float* q, float target, int64_t n, int32_t mode, float* loss,
                nc_stream_t stream);
int nc_loss_bwd(const float* p, const float* q, float target, int64_t n, int32_t mode, const float* upstream,
                float* dp, nc_stream_t stream);

/* torch.optim.Adam step (apollo_model.py:131-136: lr, betas=(beta1, 0.999), eps 1e-8, no weight decay) for one
 * parameter tensor: p, exp_avg m, exp_avg_sq v updated in place from gradient g; `step` counts from 1. */
int nc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 int32_t step, nc_stream_t stream);
/* Training-data augmentation on the device (data/base_dataset.py:87-131, README --preprocess
 * random3Drotate_randomcrop_randomflip_...): the rotated (cv2.warpAffine INTER_LINEAR, bit-exact), cropped, normalised
 * and flipped float32 crop (cz,cy,cx), evaluated only at the crop's voxels from the uint16 volume (Z,H,W) in HBM.
 * x0 / y0 (cy entries) and adelta / bdelta (cx entries): the inverse affine map in 1/1024 fixed point for the crop's
 * destination rows / columns, computed by the host (neuroclear_b200/augment.py); flip_mask bits 0/1/2 = z/y/x. */
int nc_augment_crop_u16(const uint16_t* vol, int32_t z, int32_t h, int32_t w, int32_t z0, int32_t cz, int32_t cy,
                        int32_t cx, const int32_t* x0, const int32_t* y0, const int32_t* adelta, const int32_t* bdelta,
                        int32_t flip_mask, float* out, nc_stream_t stream);

/* All parameter tensors of one optimiser in a single launch.  table: DEVICE array of `count` entries
 * { float* p; const float* g; float* m; float* v; int64_t n; } (40 bytes each); same arithmetic as nc_adam_step. */
int nc_adam_step_multi(const void* table, int32_t count, float lr, float beta1, float beta2, float eps, int32_t step,
                       nc_stream_t stream);

/* The whole 2-D PatchGAN (NLayerDiscriminator, networks.py:1030-1057: conv s2 + LReLU, (conv s2, IN, LReLU) x
 * (n_layers-1), (conv s1, IN, LReLU), conv s1 -> 1) in one call per direction: the layer loop runs inside the library.
 * x: float32 (N,1,H,W); weights / biases: HOST arrays of n_layers+2 device pointers (OIHW float32 / float32);
 * ws: nc_patchgan_ws_floats(...) floats, written by fwd and consumed by bwd; pred: (N,1,h_out,w_out).
 * bwd: dx (nullable) = gradient w.r.t. the image; dweights / dbiases (HOST arrays of device pointers, nullable
 * together) = parameter gradients, overwritten. */
int64_t nc_patchgan_ws_floats(int32_t n, int32_t h, int32_t w, int32_t ndf, int32_t n_layers);
int nc_patchgan_fwd(const float* x, int32_t n, int32_t h, int32_t w, int32_t ndf, int32_t n_layers,
                    const float* const* weights, const float* const* biases, float* ws, float* pred,
                    nc_stream_t stream);
int nc_patchgan_bwd(const float* x, const float* dpred, int32_t n, int32_t h, int32_t w, int32_t ndf, int32_t n_layers,
                    const float* const* weights, float* ws, float* dx, float* const* dweights, float* const* dbiases,
                    nc_stream_t stream);

/* ---- whole-network entry (nc_unet_deconv_infer_cube) ------------------------------------------------
 * Unet_deconv.forward (models/networks.py:512-538) for nb cubes of d x h x w voxels (each divisible by 4) in ONE call:
 * x float32 (nb, d, h, w) -> out float32 (nb, d-2c, h-2c, w-2c), sigmoid output with the border `crop` already cut
 * (util/assemble_dice.py:143).  Weights are the PACKED images produced by the nc_pack_weights_* calls, in the order
 * of the reference's state_dict (k3: double_conv1.3, double_conv2.0, double_conv2.3, bottom_layer.0/.3/.6,
 * ex_double_conv2.0/.3, ex_conv1_1.0; ct: t_conv2, t_conv1; head = {one_by_one.weight[64], one_by_one.bias,
 * one_by_one_2.weight, one_by_one_2.bias} float32).  workspace: nc_unet_deconv_workspace_bytes(...) bytes of device
 * memory, initialised ONCE with nc_unet_deconv_workspace_init.  Stateless, asynchronous on `stream`, allocation-free:
 * the call may be captured into a CUDA graph (fixed x / out / workspace) and replayed. */
typedef struct nc_unet_deconv_weights {
  const void* first;        /* nc_pack_weights_conv3d_cin1_k3 */
  const void* k3[9];        /* nc_pack_weights_conv3d_k3 */
  const void* ct[2];        /* nc_pack_weights_convT3d_k2s2 */
  const float* ct_bias[2];  /* t_conv2.bias, t_conv1.bias */
  const float* head;        /* 67 floats */
} nc_unet_deconv_weights;
int64_t nc_unet_deconv_workspace_bytes(int32_t nb, int32_t d, int32_t h, int32_t w);
int nc_unet_deconv_workspace_init(void* workspace, int64_t workspace_bytes, int32_t nb, int32_t d, int32_t h, int32_t w,
                                  nc_stream_t stream);
int nc_unet_deconv_infer_cube(const float* x, int32_t nb, int32_t d, int32_t h, int32_t w,
                              const nc_unet_deconv_weights* weights, void* workspace, int64_t workspace_bytes,
                              int32_t crop, float* out, nc_stream_t stream);

/* ---- remaining Assemble_Dice / test_dice.py options -------------------------------------------
 * --histogram_match (util/assemble_dice.py:150-151): skimage.exposure.match_histograms(fake, real) of one border-cut
 * cube (n = roi^3 float32 voxels each), skimage/exposure/histogram_matching.py::_match_cumulative_cdf restated:
 * np.unique both, quantiles = cumsum(counts) / n, np.interp(src_q, tmpl_q, tmpl_values)[inverse] -> float64 (n).
 * scratch: nc_hist_match_scratch_bytes(n) bytes of device memory (two sorted copies + radix-sort workspace). */
int64_t nc_hist_match_scratch_bytes(int64_t n);
int nc_hist_match_f32(const float* fake, const float* real, int64_t n, void* scratch, int64_t scratch_bytes,
                      double* out, nc_stream_t stream);
/* nc_blend_gather_f32 for a queue of float64 cubes (what match_histograms returns): numpy evaluates
 * `visual_ret[...] += cube / 8` (assemble_dice.py:172) with the float64 loop and stores float32. */
int nc_blend_gather_f64(const double* pieces, const int64_t* piece_off, const int32_t* piece_z0,
                        const int32_t padded_zyx[3], const int32_t steps_zyx[3], int32_t roi, int32_t overlap,
                        int32_t out_z0, int32_t out_nz, float* out, nc_stream_t stream);
/* --save_projections (test_dice.py:159-177): np.amax(volume[a0:a1 along axis], axis) of a (z,y,x) uint16
 * (elem_bytes 2) or uint8 (1) volume; a0/a1 are clamped like a numpy slice. out: the 2-D image, same type. */
int nc_amax_axis(const void* vol, int32_t elem_bytes, int32_t z, int32_t y, int32_t x, int32_t axis, int32_t a0,
                 int32_t a1, void* out, nc_stream_t stream);
/* PSNR report (test_dice.py:229-270, util/util.py:56-71,107-115): exact integer moments {sum, sum of squares, min,
 * max} of a uint8 / uint16 volume; util.normalize(util.standardize(v), np.uint8) per voxel in float64 with
 * params4 = {mean, std, standardized minimum, 255 / (standardized max - min)}; sum of squared differences of two
 * uint8 volumes (an exact integer). */
int nc_volume_moments(const void* vol, int32_t elem_bytes, int64_t n, uint64_t* out4, nc_stream_t stream);
/* sum over voxels of (double(v) - mean)^2 in numpy's pairwise association order (what np.std computes: the last bit
 * of this float64 sum decides the reference's double-normalised uint8 volumes).  scratch:
 * nc_pairwise_sqdev_scratch_doubles(n) doubles; out1: one double. */
int64_t nc_pairwise_sqdev_scratch_doubles(int64_t n);
int nc_pairwise_sqdev_sum(const void* vol, int32_t elem_bytes, int64_t n, double mean, double* scratch, double* out1,
                          nc_stream_t stream);
int nc_standardize_normalize_u8(const void* vol, int32_t elem_bytes, int64_t n, const double* params4, uint8_t* out,
                                nc_stream_t stream);
int nc_sqdiff_u8(const uint8_t* a, const uint8_t* b, int64_t n, uint64_t* out1, nc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NEUROCLEAR_B200_H_ */

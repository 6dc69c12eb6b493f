// Per-voxel arithmetic of nc_augment_crop_u16, shared by the CUDA kernel (augment.cu) and a plain-C host build used
// by the CPU tests (tests/cuda/augment_host.c, compiled with -ffp-contract=off) so that the index and rounding logic
// of the kernel source itself is checked against the reference fixture without a GPU.
//
// cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) on a 16-bit image (OpenCV imgwarp.cpp, remapBilinear<Cast<float,
// ushort>, RemapNoVec, float>): fixed-point source coordinate with 5 fractional bits, weights (1-fy)(1-fx), (1-fy)fx,
// fy(1-fx), fy fx formed in float from the 1/32 table, products and sums in float, left to right, no contraction;
// saturate_cast<ushort> = round half to even + clamp; then the reference's __normalize (/ 65535 in float64) and
// .float().
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define NC_FMUL(a, b) __fmul_rn((a), (b))
#define NC_FADD(a, b) __fadd_rn((a), (b))
#else
#define NC_FMUL(a, b) ((a) * (b))
#define NC_FADD(a, b) ((a) + (b))
#endif
#if defined(__CUDACC__)
#define NC_HD __host__ __device__ __forceinline__
#else
#define NC_HD static inline
#endif

NC_HD float nc_augment_tabw(int f, int which) {
  const float t = (float)f * (1.0f / 32.0f); /* exact */
  return which ? t : 1.0f - t;
}

NC_HD float nc_augment_fetch(const uint16_t* slice, int H, int W, int yy, int xx) {
  return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (float)slice[(size_t)yy * W + xx] : 0.f;
}

/* xfix / yfix: (x0[row] + adelta[col]) and (y0[row] + bdelta[col]) of cv2's inverse map (1/1024 units, round_delta
 * included).  Returns the normalised float32 voxel. */
NC_HD float nc_augment_voxel(const uint16_t* slice, int H, int W, int xfix, int yfix) {
  const int X = xfix >> 5, Y = yfix >> 5; /* AB_BITS - INTER_BITS; arithmetic shift */
  const int sx = X >> 5, sy = Y >> 5, fx = X & 31, fy = Y & 31;
  const float w00 = NC_FMUL(nc_augment_tabw(fy, 0), nc_augment_tabw(fx, 0));
  const float w01 = NC_FMUL(nc_augment_tabw(fy, 0), nc_augment_tabw(fx, 1));
  const float w10 = NC_FMUL(nc_augment_tabw(fy, 1), nc_augment_tabw(fx, 0));
  const float w11 = NC_FMUL(nc_augment_tabw(fy, 1), nc_augment_tabw(fx, 1));
  float acc = NC_FMUL(nc_augment_fetch(slice, H, W, sy, sx), w00);
  acc = NC_FADD(acc, NC_FMUL(nc_augment_fetch(slice, H, W, sy, sx + 1), w01));
  acc = NC_FADD(acc, NC_FMUL(nc_augment_fetch(slice, H, W, sy + 1, sx), w10));
  acc = NC_FADD(acc, NC_FMUL(nc_augment_fetch(slice, H, W, sy + 1, sx + 1), w11));
  float v = rintf(acc); /* round half to even (default rounding mode) */
  v = fminf(fmaxf(v, 0.f), 65535.f);
  return (float)((double)v / 65535.0);
}

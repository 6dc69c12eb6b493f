/* Host build of the per-voxel arithmetic of nc_augment_crop_u16 (neuroclear_b200/csrc/augment_math.h): the same
 * loop as the CUDA kernel, compiled by the CPU tests with `gcc -O2 -ffp-contract=off -shared -fPIC`. */
#include "../../neuroclear_b200/csrc/augment_math.h"

void augment_crop_host(const uint16_t* vol, int H, int W, int z0, int cz, int cy, int cx, const int* x0, const int* y0,
                       const int* adelta, const int* bdelta, int flip_mask, float* out) {
  for (int k = 0; k < cz; ++k)
    for (int i = 0; i < cy; ++i)
      for (int j = 0; j < cx; ++j) {
        const float o = nc_augment_voxel(vol + (size_t)(z0 + k) * H * W, H, W, x0[i] + adelta[j], y0[i] + bdelta[j]);
        const int ko = (flip_mask & 1) ? cz - 1 - k : k;
        const int io = (flip_mask & 2) ? cy - 1 - i : i;
        const int jo = (flip_mask & 4) ? cx - 1 - j : j;
        out[((size_t)ko * cy + io) * cx + jo] = o;
      }
}

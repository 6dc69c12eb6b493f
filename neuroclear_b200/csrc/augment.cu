// Training-data augmentation on the device (reference data/base_dataset.py:87-131 with the README's --preprocess
// random3Drotate_randomcrop_randomflip_addColorChannel_addBatchChannel).  The reference
// rotates EVERY z-slice of the whole volume with cv2.warpAffine on the host for every iteration and then crops; this
// kernel evaluates the same rotation only at the voxels of the requested crop, from the uint16 volume resident in
// HBM, and writes the normalised (and flipped) float32 crop the generator consumes.
//
// Bit-exact restatement of cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) for 16-bit images (imgwarp.cpp): the
// inverse map in 1/1024 fixed point (host side: x0/y0 per destination row, adelta/bdelta per destination column),
// 1/32 sub-pixel weights from a float table, float accumulation in a fixed order without contraction, round-half-even.
#include "augment_math.h"
#include "internal.h"

namespace nc {

namespace {

__global__ void __launch_bounds__(256)
augment_crop_kernel(const uint16_t* __restrict__ vol, int H, int W, int z0, int cz, int cy, int cx,
                    const int* __restrict__ x0, const int* __restrict__ y0, const int* __restrict__ adelta,
                    const int* __restrict__ bdelta, int flip_mask, float* __restrict__ out) {
  const long long total = static_cast<long long>(cz) * cy * cx;
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += gridDim.x * 256ll) {
    const int j = static_cast<int>(idx % cx);
    const long long r = idx / cx;
    const int i = static_cast<int>(r % cy), k = static_cast<int>(r / cy);
    const float o = nc_augment_voxel(vol + static_cast<size_t>(z0 + k) * H * W, H, W, x0[i] + adelta[j],
                                     y0[i] + bdelta[j]);
    const int ko = (flip_mask & 1) ? cz - 1 - k : k;
    const int io = (flip_mask & 2) ? cy - 1 - i : i;
    const int jo = (flip_mask & 4) ? cx - 1 - j : j;
    out[(static_cast<size_t>(ko) * cy + io) * cx + jo] = o;
  }
}

}  // namespace

int augment_crop_u16(const uint16_t* vol, int Z, int H, int W, int z0, int cz, int cy, int cx, const int* x0,
                     const int* y0, const int* adelta, const int* bdelta, int flip_mask, float* out,
                     cudaStream_t stream) {
  if (z0 < 0 || cz < 1 || z0 + cz > Z || cy < 1 || cx < 1) return set_error("augment_crop_u16: crop outside the volume");
  if (flip_mask & ~7) return set_error("augment_crop_u16: flip mask has bits 0 (z), 1 (y), 2 (x)");
  augment_crop_kernel<<<num_sms() * 8, 256, 0, stream>>>(vol, H, W, z0, cz, cy, cx, x0, y0, adelta, bdelta, flip_mask,
                                                         out);
  NC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nc

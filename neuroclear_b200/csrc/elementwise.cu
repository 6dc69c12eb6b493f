// Memory-bound kernels of the hot path: dice extraction, InstanceNorm
// finalize/apply(+pool), the 1x1x1+sigmoid head, overlap blend, radix-select percentile, rescale/cast/crop and
// the max-intensity projection.  All are coalesced, vectorised where the layout allows, and bit-exact where the
// reference is integer / fixed-order fp32 arithmetic.  Reference citations are in include/neuroclear_b200.h.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "internal.h"

namespace nc {

// ------------------------------------------------------------------------------------------------ dice
__device__ __forceinline__ int reflect_index(int j, int n) {
  // numpy 'reflect' (no edge repeat): -i -> i, n-1+i -> n-1-i
  if (j < 0) j = -j;
  if (j >= n) j = 2 * (n - 1) - j;
  return j;
}

template <typename T>
__global__ void dice_extract_kernel(const T* __restrict__ vol, int vz0, int vnz, int Z, int Y, int X, int Pz,
                                    int Py, int Px, int ny, int nx, int step, int bc, int E, long long cube_begin,
                                    float* __restrict__ out) {
  const int local_cube = blockIdx.y;
  const long long cube = cube_begin + local_cube;
  const int cx = static_cast<int>(cube % nx);
  const int cy = static_cast<int>((cube / nx) % ny);
  const int cz = static_cast<int>(cube / (static_cast<long long>(nx) * ny));
  const unsigned E3 = static_cast<unsigned>(E) * E * E;
  float* dst = out + static_cast<size_t>(local_cube) * E3;
  for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < E3; r += gridDim.x * blockDim.x) {
    const int c = r % E;
    const unsigned r2 = r / E;
    const int b = r2 % E;
    const int a = r2 / E;
    const int z = reflect_index(cz * step - bc + a, Pz);
    const int y = reflect_index(cy * step - bc + b, Py);
    const int x = reflect_index(cx * step - bc + c, Px);
    float v = 0.f;  // far-end zero padding of pad_for_dicing
    if (z < Z && y < Y && x < X) {
      const int zl = z - vz0;
      if (zl >= 0 && zl < vnz) v = static_cast<float>(vol[(static_cast<size_t>(zl) * Y + y) * X + x]);
    }
    dst[r] = __fdiv_rn(v, sizeof(T) == 2 ? 65535.0f : 255.0f);   // base_dataset.py:134-143
  }
}

template <typename T>
static int dice_extract_t(const T* vol, int vz0, int vnz, const int* size, const int* padded, const int* steps,
                          int roi, int overlap, int border, long long cube_begin, int cube_count, float* cubes,
                          cudaStream_t stream) {
  if (cube_count <= 0) return 0;
  if (cube_count > 65535) return set_error("dice_extract: at most 65535 cubes per call");
  const int E = roi + 2 * border;
  if (border >= padded[0] || border >= padded[1] || border >= padded[2])
    return set_error("dice_extract: border_cut must be smaller than the padded volume");
  const unsigned E3 = static_cast<unsigned>(E) * E * E;
  dim3 grid((E3 + 255) / 256, cube_count);
  if (grid.x > 4096) grid.x = 4096;
  dice_extract_kernel<T><<<grid, 256, 0, stream>>>(vol, vz0, vnz, size[0], size[1], size[2], padded[0], padded[1],
                                                padded[2], steps[1], steps[2], roi - overlap, border, E, cube_begin,
                                                cubes);
  NC_CUDA(cudaGetLastError());
  return 0;
}

int dice_extract_u16(const uint16_t* vol, int vz0, int vnz, const int* size, const int* padded, const int* steps,
                     int roi, int overlap, int border, long long cube_begin, int cube_count, float* cubes,
                     cudaStream_t stream) {
  return dice_extract_t(vol, vz0, vnz, size, padded, steps, roi, overlap, border, cube_begin, cube_count, cubes, stream);
}
int dice_extract_u8(const uint8_t* vol, int vz0, int vnz, const int* size, const int* padded, const int* steps,
                    int roi, int overlap, int border, long long cube_begin, int cube_count, float* cubes,
                    cudaStream_t stream) {
  return dice_extract_t(vol, vz0, vnz, size, padded, steps, roi, overlap, border, cube_begin, cube_count, cubes, stream);
}

// ------------------------------------------------------------------------------------------------ InstanceNorm
// Two-level fixed-order reduction in fp64: grid (C/32, NB, S). Every block sums one slice of the partial rows
// (8 row-lanes x 32 channels, combined in a fixed order), stores its slice total, and the LAST block to finish for
// a (sample, channel group) — found with an arrival counter — adds the S slice totals in slice order and writes
// mean / rstd.  The summation order never depends on scheduling, so results are bitwise repeatable.
constexpr int FIN_SLICES = 64;
static_assert(FIN_SLICES % 8 == 0, "the final pass loads eight slice totals at a time");

__global__ void __launch_bounds__(256)
in_stats_finalize_kernel(const float* __restrict__ partial, long long rows, int C, double inv_n, float eps,
                         double* __restrict__ slice_tot, unsigned int* __restrict__ counters,
                         float* __restrict__ mean_rstd) {
  __shared__ double ssum[8][32], ssq[8][32];
  __shared__ bool is_last;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int nb = blockIdx.y, slice = blockIdx.z, S = gridDim.z;
  const long long per = (rows + S - 1) / S;
  const long long r0 = slice * per, r1 = min(rows, r0 + per);
  double s = 0.0, q = 0.0;
  if (c < C) {
    const float* p = partial + static_cast<size_t>(nb) * rows * 2 * C;
    // four rows' loads in flight per step, added in the same (ascending) order as a plain loop: this kernel sits
    // between every conv and its consumer, and a rolled loop made it a chain of L2 round trips (26 us per launch)
    long long r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      float vs[4], vq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        vs[u] = __ldg(p + ((r + 8 * u) * 2) * C + c);
        vq[u] = __ldg(p + ((r + 8 * u) * 2 + 1) * C + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s += static_cast<double>(vs[u]);
        q += static_cast<double>(vq[u]);
      }
    }
    for (; r < r1; r += 8) {
      s += static_cast<double>(__ldg(p + (r * 2) * C + c));
      q += static_cast<double>(__ldg(p + (r * 2 + 1) * C + c));
    }
  }
  ssum[ty][tx] = s;
  ssq[ty][tx] = q;
  __syncthreads();
  double* tot = slice_tot + ((static_cast<size_t>(nb) * S + slice) * 2) * C;
  if (ty == 0 && c < C) {
    double S1 = 0.0, Q1 = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      S1 += ssum[i][tx];
      Q1 += ssq[i][tx];
    }
    tot[c] = S1;
    tot[C + c] = Q1;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* cnt = counters + nb * gridDim.x + blockIdx.x;
    const unsigned int prev = atomicAdd(cnt, 1u);
    is_last = (prev == static_cast<unsigned int>(S - 1));
    if (is_last) *cnt = 0;  // ready for the next launch
  }
  __syncthreads();
  if (is_last && ty == 0 && c < C) {
    __threadfence();
    double S1 = 0.0, Q1 = 0.0;
    const double* base = slice_tot + (static_cast<size_t>(nb) * S * 2) * C;
    for (int i0 = 0; i0 < S; i0 += 8) {  // S = FIN_SLICES is a multiple of 8; loads first, adds in slice order
      double vs[8], vq[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        vs[u] = __ldcg(base + (static_cast<size_t>(i0 + u) * 2) * C + c);
        vq[u] = __ldcg(base + (static_cast<size_t>(i0 + u) * 2 + 1) * C + c);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        S1 += vs[u];
        Q1 += vq[u];
      }
    }
    const double mean = S1 * inv_n;
    double var = Q1 * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_rstd[(static_cast<size_t>(nb) * 2) * C + c] = static_cast<float>(mean);
    mean_rstd[(static_cast<size_t>(nb) * 2 + 1) * C + c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

// scratch layout: [arrival counters: FIN_COUNTER_BYTES, at a FIXED offset so that calls with different (NB, C)
// sharing one scratch buffer never overwrite each other's counters][slice totals: NB x S x 2 x C doubles]
constexpr size_t FIN_COUNTER_BYTES = 65536;

size_t in_stats_scratch_bytes(int NB, int C) {
  return FIN_COUNTER_BYTES + static_cast<size_t>(NB) * FIN_SLICES * 2 * C * sizeof(double);
}

int in_stats_finalize(const float* partial, int NB, long long rows, int C, long long voxels, float eps, void* scratch,
                      float* mean_rstd, cudaStream_t stream) {
  if (NB > 65535) return set_error("in_stats_finalize: NB too large");
  if (static_cast<size_t>(NB) * ((C + 31) / 32) * 4 > FIN_COUNTER_BYTES)
    return set_error("in_stats_finalize: NB * C too large for the counter region");
  unsigned int* counters = static_cast<unsigned int*>(scratch);
  double* slice_tot = reinterpret_cast<double*>(static_cast<char*>(scratch) + FIN_COUNTER_BYTES);
  dim3 grid((C + 31) / 32, NB, FIN_SLICES);
  in_stats_finalize_kernel<<<grid, 256, 0, stream>>>(partial, rows, C, 1.0 / static_cast<double>(voxels), eps,
                                                     slice_tot, counters, mean_rstd);
  NC_CUDA(cudaGetLastError());
  return 0;
}

__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  // InstanceNorm outputs are bounded by sqrt(#voxels) << 65504, the clamp only guards degenerate inputs
  __half2 t = __floats2half2_rn(fminf(a, 65504.f), fminf(b, 65504.f));
  return *reinterpret_cast<uint32_t*>(&t);
}

// eight channels of one voxel (one 16-byte fp16 load): relu((x-mean)*rstd) in fp32
__device__ __forceinline__ void norm8(const __half* __restrict__ src, const float (&mu)[8], const float (&rs)[8],
                                      float (&o)[8]) {
  const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(src));
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h2[i]);
    o[2 * i] = fmaxf((f.x - mu[2 * i]) * rs[2 * i], 0.f);
    o[2 * i + 1] = fmaxf((f.y - mu[2 * i + 1]) * rs[2 * i + 1], 0.f);
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// BF16: the output (same values, relu(IN(raw))) is written as bf16 — the weight-gradient kernels read activations in
// the gradient's format (tcgen05 cannot mix fp16 and bf16 operands).
template <bool BF16>
__device__ __forceinline__ uint4 pack8(const float (&o)[8]) {
  if constexpr (BF16)
    return make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                      pack_bf16x2(o[6], o[7]));
  return make_uint4(pack_f16x2(o[0], o[1]), pack_f16x2(o[2], o[3]), pack_f16x2(o[4], o[5]), pack_f16x2(o[6], o[7]));
}

template <bool BF16>
__global__ void __launch_bounds__(256)
in_relu_apply_kernel(const __half* __restrict__ raw, const float* __restrict__ mean_rstd, long long voxels, int C,
                     __half* __restrict__ y, int y_ld, int y_coff) {
  const int nb = blockIdx.y;
  const unsigned cg_per = C / 8;
  const unsigned total = static_cast<unsigned>(voxels) * cg_per;  // per cube: < 2^31 (checked by the launcher)
  const float* mr = mean_rstd + static_cast<size_t>(nb) * 2 * C;
  // the grid stride (gridDim.x * 256) is a multiple of C / 8 (a power of two <= 64): a thread keeps its channel group
  const int cg = static_cast<int>((blockIdx.x * 256u + threadIdx.x) % cg_per);
  float mu[8], rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = __ldg(mr + cg * 8 + i);
    rs[i] = __ldg(mr + C + cg * 8 + i);
  }
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const unsigned vox = idx / cg_per;
    float o[8];
    const long long gv = static_cast<long long>(nb) * voxels + vox;
    norm8(raw + gv * C + cg * 8, mu, rs, o);
    *reinterpret_cast<uint4*>(y + gv * y_ld + y_coff + cg * 8) = pack8<BF16>(o);
  }
}

template <bool BF16>
__global__ void __launch_bounds__(256, 3)
in_relu_pool_apply_kernel(const __half* __restrict__ raw, const float* __restrict__ mean_rstd, int D, int H, int W,
                          int C, __half* __restrict__ y, int y_ld, int y_coff,
                          __half* __restrict__ pooled) {
  const int nb = blockIdx.y;
  const int cg_per = C / 8;
  const int PD = D / 2, PH = H / 2, PW = W / 2;
  const unsigned total = static_cast<unsigned>(PD) * PH * PW * cg_per;  // per cube: < 2^31 (checked by the launcher)
  const float* mr = mean_rstd + static_cast<size_t>(nb) * 2 * C;
  // the grid stride (gridDim.x * 256) is a multiple of C / 8 (a power of two <= 64): a thread keeps its channel group
  const int cg = static_cast<int>((blockIdx.x * 256u + threadIdx.x) % cg_per);
  float mu[8], rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = __ldg(mr + cg * 8 + i);
    rs[i] = __ldg(mr + C + cg * 8 + i);
  }
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    unsigned r = idx / cg_per;
    const int pw = static_cast<int>(r % PW);
    r /= PW;
    const int ph = static_cast<int>(r % PH);
    const int pd = static_cast<int>(r / PH);
    // all eight 16-byte loads of the 2 x 2 x 2 window are issued before the first use
    uint4 rawv[8];
    const long long gv0 = ((static_cast<long long>(nb) * D + 2 * pd) * H + 2 * ph) * W + 2 * pw;
    const long long sH = W, sD = static_cast<long long>(H) * W;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long gv = gv0 + (k >> 2) * sD + ((k >> 1) & 1) * sH + (k & 1);
      rawv[k] = __ldcs(reinterpret_cast<const uint4*>(raw + gv * C + cg * 8));
    }
    float mx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) mx[i] = 0.f;  // post-ReLU values are >= 0
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&rawv[k]);
      float o[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h2[i]);
        o[2 * i] = fmaxf((f.x - mu[2 * i]) * rs[2 * i], 0.f);
        o[2 * i + 1] = fmaxf((f.y - mu[2 * i + 1]) * rs[2 * i + 1], 0.f);
      }
      const long long gv = gv0 + (k >> 2) * sD + ((k >> 1) & 1) * sH + (k & 1);
      *reinterpret_cast<uint4*>(y + gv * y_ld + y_coff + cg * 8) = pack8<BF16>(o);
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = fmaxf(mx[i], o[i]);
    }
    const long long pv = ((static_cast<long long>(nb) * PD + pd) * PH + ph) * PW + pw;
    *reinterpret_cast<uint4*>(pooled + pv * C + cg * 8) = pack8<BF16>(mx);
  }
}

template <bool BF16>
static int in_relu_apply_t(const void* raw_v, const float* mean_rstd, int NB, int D, int H, int W, int C, void* y,
                           int y_ld, int y_coff, void* pooled, cudaStream_t stream) {
  if (C % 8 || y_ld % 8 || y_coff % 8) return set_error("in_relu_apply: channel counts must be multiples of 8");
  if (256 % (C / 8)) return set_error("in_relu_apply: C / 8 must divide 256 (a thread keeps one channel group)");
  if (NB > 65535) return set_error("in_relu_apply: NB too large");
  if (static_cast<long long>(D) * H * W * (C / 8) >= (1ll << 31)) return set_error("in_relu_apply: cube too large");
  const __half* raw = static_cast<const __half*>(raw_v);
  const int blocks = num_sms() * 8;
  if (pooled) {
    if ((D | H | W) & 1) return set_error("in_relu_apply: pooling needs even D, H, W");
    in_relu_pool_apply_kernel<BF16><<<dim3(blocks, NB), 256, 0, stream>>>(raw, mean_rstd, D, H, W, C,
                                                                    static_cast<__half*>(y), y_ld, y_coff,
                                                                    static_cast<__half*>(pooled));
  } else {
    in_relu_apply_kernel<BF16><<<dim3(blocks, NB), 256, 0, stream>>>(raw, mean_rstd, static_cast<long long>(D) * H * W, C,
                                                               static_cast<__half*>(y), y_ld, y_coff);
  }
  NC_CUDA(cudaGetLastError());
  return 0;
}

int in_relu_apply(const void* raw_v, const float* mean_rstd, int NB, int D, int H, int W, int C, void* y, int y_ld,
                  int y_coff, void* pooled, cudaStream_t stream) {
  return in_relu_apply_t<false>(raw_v, mean_rstd, NB, D, H, W, C, y, y_ld, y_coff, pooled, stream);
}
int in_relu_apply_bf16(const void* raw_v, const float* mean_rstd, int NB, int D, int H, int W, int C, void* y,
                       int y_ld, int y_coff, void* pooled, cudaStream_t stream) {
  return in_relu_apply_t<true>(raw_v, mean_rstd, NB, D, H, W, C, y, y_ld, y_coff, pooled, stream);
}

// ------------------------------------------------------------------------------------------------ head
// 8 lanes per voxel, 8 channels per lane (C == 64): IN + ReLU + dot(w1) + b1, * w2 + b2, sigmoid.
__global__ void __launch_bounds__(256)
head_kernel(const __half* __restrict__ raw, const float* __restrict__ mean_rstd, const float* __restrict__ hp, int D,
            int H, int W, int crop, float* __restrict__ y) {
  constexpr int C = 64;
  const int nb = blockIdx.y;
  const int sub = threadIdx.x & 7;
  const float* mr = mean_rstd + static_cast<size_t>(nb) * 2 * C;
  float mu[8], rs[8], w1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = __ldg(mr + sub * 8 + i);
    rs[i] = __ldg(mr + C + sub * 8 + i);
    w1[i] = __ldg(hp + sub * 8 + i);
  }
  const float b1 = __ldg(hp + C), w2 = __ldg(hp + C + 1), b2 = __ldg(hp + C + 2);
  const unsigned voxels = static_cast<unsigned>(D) * H * W;  // per cube: < 2^29 (checked by the launcher)
  const int OD = D - 2 * crop, OH = H - 2 * crop, OW = W - 2 * crop;
  // A warp handles 4 consecutive voxels per step (8 lanes x 8 channels each) and 8 steps per iteration.  The eight
  // partial dot products of a lane (one per step) are reduced across the 8 lanes of its group by recursive halving
  // (7 shuffles), which leaves lane `sub` with the complete sum of step `sub`: every lane then finishes ONE voxel
  // (bias, second 1x1, sigmoid, border test, store) instead of one lane in eight doing all of it.
  const unsigned warps_total = gridDim.x * 8u;
  const unsigned grp = (threadIdx.x & 31) >> 3;
  for (unsigned q = blockIdx.x * 8u + (threadIdx.x >> 5); q * 4 < voxels; q += 8 * warps_total) {
    float part[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const unsigned vox = (q + u * warps_total) * 4 + grp;
      const size_t gv = static_cast<size_t>(nb) * voxels + (vox < voxels ? vox : 0u);
      float o[8];
      norm8(raw + gv * C + sub * 8, mu, rs, o);
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t = fmaf(o[i], w1[i], t);
      part[u] = t;
    }
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) {
      const bool upper = (sub & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float keep = upper ? part[i + off] : part[i];
        const float send = upper ? part[i] : part[i + off];
        part[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    const unsigned vox = (q + sub * warps_total) * 4 + grp;  // the voxel this lane finishes
    if (vox < voxels) {
      const int w = static_cast<int>(vox % W);
      const unsigned r = vox / W;
      const int h = static_cast<int>(r % H);
      const int d = static_cast<int>(r / H);
      const int od = d - crop, oh = h - crop, ow = w - crop;
      if (od >= 0 && od < OD && oh >= 0 && oh < OH && ow >= 0 && ow < OW) {
        const float uu = fmaf(w2, part[0] + b1, b2);
        y[((static_cast<size_t>(nb) * OD + od) * OH + oh) * OW + ow] = 1.0f / (1.0f + expf(-uu));
      }
    }
  }
}

int head_1x1_sigmoid_fwd(const void* raw, const float* mean_rstd, const float* hp, int NB, int D, int H, int W,
                         int C, int crop, float* y, cudaStream_t stream) {
  if (C != 64) return set_error("head_1x1_sigmoid_fwd: C must be 64");
  if (crop < 0 || 2 * crop >= D || 2 * crop >= H || 2 * crop >= W) return set_error("head: bad crop");
  if (static_cast<long long>(D) * H * W >= (1ll << 29)) return set_error("head: cube too large");
  const int blocks = num_sms() * 8;
  head_kernel<<<dim3(blocks, NB), 256, 0, stream>>>(static_cast<const __half*>(raw), mean_rstd, hp, D, H, W, crop, y);
  NC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ blend
struct AxisCover {
  int j0, j1;  // lower covering cube (or -1) and upper covering cube
};
__device__ __forceinline__ AxisCover axis_cover(int q, int step, int roi, int k) {
  AxisCover c;
  c.j1 = min(q / step, k - 1);
  c.j0 = (c.j1 >= 1 && q < (c.j1 - 1) * step + roi) ? c.j1 - 1 : -1;
  return c;
}

__global__ void __launch_bounds__(256)
blend_gather_kernel(const float* __restrict__ pieces, const long long* __restrict__ piece_off,
                    const int* __restrict__ piece_z0, int Py, int Px, int nz, int ny, int nx, int roi, int step,
                    int out_z0, float* __restrict__ out) {
  const int x = blockIdx.x * 256 + threadIdx.x;
  if (x >= Px) return;
  const int y = blockIdx.y;
  const int z = out_z0 + blockIdx.z;
  const AxisCover cz = axis_cover(z, step, roi, nz), cy = axis_cover(y, step, roi, ny),
                  cx = axis_cover(x, step, roi, nx);
  float acc = 0.f;
  int n = 0;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int jz = a ? cz.j1 : cz.j0;
    if (jz < 0) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int jy = b ? cy.j1 : cy.j0;
      if (jy < 0) continue;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int jx = c ? cx.j1 : cx.j0;
        if (jx < 0) continue;
        const long long cube = (static_cast<long long>(jz) * ny + jy) * nx + jx;
        const long long off = piece_off[cube];
        const int lz = z - jz * step - piece_z0[cube];
        const float v = pieces[off + (static_cast<long long>(lz) * roi + (y - jy * step)) * roi + (x - jx * step)];
        acc = acc + v * 0.125f;  // visual_ret += cube / 8 (assemble_dice.py:172), ascending cube index
        ++n;
      }
    }
  }
  // (visual_ret / mask_ret) * 8  (assemble_dice.py:184)
  out[(static_cast<size_t>(blockIdx.z) * Py + y) * Px + x] = __fdiv_rn(acc, static_cast<float>(n)) * 8.0f;
}

int blend_gather_f32(const float* pieces, const long long* piece_off, const int* piece_z0, const int* padded,
                     const int* steps, int roi, int overlap, int out_z0, int out_nz, float* out,
                     cudaStream_t stream) {
  if (overlap <= 0) return set_error("blend_gather: overlap must be > 0 (the reference produces zeros otherwise)");
  if (2 * overlap > roi) return set_error("blend_gather: overlap must be <= roi - overlap");
  if (out_nz <= 0) return 0;
  if (out_nz > 65535 || padded[1] > 65535) return set_error("blend_gather: slab too large for one launch");
  dim3 grid((padded[2] + 255) / 256, padded[1], out_nz);
  blend_gather_kernel<<<grid, 256, 0, stream>>>(pieces, piece_off, piece_z0, padded[1], padded[2], steps[0], steps[1],
                                                steps[2], roi, roi - overlap, out_z0, out);
  NC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ radix select
struct SelectState {
  unsigned long long rank[4];
  unsigned int prefix[4];
};
constexpr int SEL_BINS = 4096;
__host__ __device__ inline int sel_shift(int pass) { return pass == 0 ? 20 : pass == 1 ? 8 : 0; }
__host__ __device__ inline int sel_bits(int pass) { return pass == 2 ? 8 : 12; }

__global__ void select_init_kernel(SelectState* st, unsigned long long r0, unsigned long long r1,
                                   unsigned long long r2, unsigned long long r3) {
  st->rank[0] = r0, st->rank[1] = r1, st->rank[2] = r2, st->rank[3] = r3;
  for (int i = 0; i < 4; ++i) st->prefix[i] = 0;
}

__global__ void __launch_bounds__(512)
select_hist_kernel(const float* __restrict__ data, long long n, int pass, const SelectState* __restrict__ st,
                   unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int sh[];  // [ntab][bins]
  const int shift = sel_shift(pass), bits = sel_bits(pass);
  const int bins = 1 << bits;
  const int ntab = pass == 0 ? 1 : 4;
  for (int i = threadIdx.x; i < ntab * bins; i += blockDim.x) sh[i] = 0;
  unsigned int pre[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) pre[t] = st->prefix[t];
  const unsigned int hi_mask = pass == 0 ? 0u : ~((1u << (shift + bits)) - 1u);
  __syncthreads();
  const long long n4 = n >> 2;
  const float4* d4 = reinterpret_cast<const float4*>(data);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldcs(d4 + i);
    const unsigned int u[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const unsigned int digit = (u[e] >> shift) & (bins - 1);
      if (pass == 0) {
        atomicAdd(&sh[digit], 1u);
      } else {
        const unsigned int hi = u[e] & hi_mask;
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (hi == pre[t]) atomicAdd(&sh[t * bins + digit], 1u);
      }
    }
  }
  if (blockIdx.x == 0) {  // scalar tail
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const unsigned int u = __float_as_uint(data[i]);
      const unsigned int digit = (u >> shift) & (bins - 1);
      if (pass == 0) {
        atomicAdd(&sh[digit], 1u);
      } else {
        const unsigned int hi = u & hi_mask;
        for (int t = 0; t < 4; ++t)
          if (hi == pre[t]) atomicAdd(&sh[t * bins + digit], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ntab * bins; i += blockDim.x) {
    const unsigned int c = sh[i];
    if (c) atomicAdd(&hist[(i / bins) * SEL_BINS + (i % bins)], static_cast<unsigned long long>(c));
  }
}

__global__ void select_update_kernel(int pass, SelectState* st, unsigned long long* hist) {
  const int shift = sel_shift(pass), bins = 1 << sel_bits(pass);
  if (threadIdx.x < 4) {
    const int t = threadIdx.x;
    const unsigned long long* h = hist + (pass == 0 ? 0 : t) * SEL_BINS;
    unsigned long long cum = 0, rank = st->rank[t];
    int b = 0;
    for (; b < bins - 1; ++b) {
      if (cum + h[b] > rank) break;
      cum += h[b];
    }
    st->rank[t] = rank - cum;
    st->prefix[t] |= static_cast<unsigned int>(b) << shift;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * SEL_BINS; i += blockDim.x) hist[i] = 0;
}

// numpy _lerp on float32 order statistics with float64 gamma (numpy/lib/_function_base_impl.py):
//   diff = f32(b - a);  r = a + diff*t;  if t >= 0.5: r = b - diff*(1-t)      (r, t in float64)
__device__ inline double np_lerp(float a, float b, double t) {
  // separate multiply and add (no FMA contraction): numpy rounds the product before the sum
  const float diff = __fsub_rn(b, a);
  double r = __dadd_rn(static_cast<double>(a), __dmul_rn(static_cast<double>(diff), t));
  if (t >= 0.5) r = __dsub_rn(static_cast<double>(b), __dmul_rn(static_cast<double>(diff), __dsub_rn(1.0, t)));
  return r;
}
__global__ void percentile_lerp_kernel(const SelectState* st, double t_lo, double t_hi, double* out64, float* out32) {
  const double p_lo = np_lerp(__uint_as_float(st->prefix[0]), __uint_as_float(st->prefix[1]), t_lo);
  const double p_hi = np_lerp(__uint_as_float(st->prefix[2]), __uint_as_float(st->prefix[3]), t_hi);
  out64[0] = p_lo;
  out64[1] = p_hi;
  out32[0] = static_cast<float>(p_lo);
  out32[1] = static_cast<float>(p_hi);
  // skimage tests `imin != imax` on the float64 percentiles (exposure.rescale_intensity); the degenerate case is
  // flagged with a NaN span so that the float32 rounding of a tiny non-zero span cannot take the wrong branch
  out32[2] = p_lo != p_hi ? static_cast<float>(p_hi - p_lo) : __int_as_float(0x7fc00000);
}

int select_init(const unsigned long long* ranks, void* st, cudaStream_t stream) {
  select_init_kernel<<<1, 1, 0, stream>>>(static_cast<SelectState*>(st), ranks[0], ranks[1], ranks[2], ranks[3]);
  NC_CUDA(cudaGetLastError());
  return 0;
}
int select_histogram(const float* data, long long n, int pass, const void* st, unsigned long long* hist,
                     cudaStream_t stream) {
  if (pass < 0 || pass > 2) return set_error("select_histogram: pass must be 0, 1 or 2");
  if (reinterpret_cast<uintptr_t>(data) & 15) return set_error("select_histogram: data must be 16-byte aligned");
  const int ntab = pass == 0 ? 1 : 4;
  const size_t smem = static_cast<size_t>(ntab) * (1 << sel_bits(pass)) * sizeof(unsigned int);
  static bool attr[64] = {false};
  if (first_use_on_device(attr))
    NC_CUDA(cudaFuncSetAttribute(select_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * SEL_BINS * 4));
  select_hist_kernel<<<num_sms() * 2, 512, smem, stream>>>(data, n, pass, static_cast<const SelectState*>(st), hist);
  NC_CUDA(cudaGetLastError());
  return 0;
}
int select_update(int pass, void* st, unsigned long long* hist, cudaStream_t stream) {
  if (pass < 0 || pass > 2) return set_error("select_update: pass must be 0, 1 or 2");
  select_update_kernel<<<1, 256, 0, stream>>>(pass, static_cast<SelectState*>(st), hist);
  NC_CUDA(cudaGetLastError());
  return 0;
}
int percentile_lerp(const void* st, double t_lo, double t_hi, double* out64, float* out32, cudaStream_t stream) {
  percentile_lerp_kernel<<<1, 1, 0, stream>>>(static_cast<const SelectState*>(st), t_lo, t_hi, out64, out32);
  NC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ rescale
template <typename T>
__global__ void __launch_bounds__(256)
rescale_crop_kernel(const float* __restrict__ vol, int vol_z0, int Py, int Px, int Y, int X,
                    const float* __restrict__ norm3, int z_begin, T* __restrict__ out) {
  const int x = blockIdx.x * 256 + threadIdx.x;
  if (x >= X) return;
  const int y = blockIdx.y;
  const int z = z_begin + blockIdx.z;
  float v = __ldcs(vol + (static_cast<size_t>(z - vol_z0) * Py + y) * Px + x);
  if (norm3) {
    const float lo = norm3[0], hi = norm3[1], span = norm3[2];
    v = fminf(fmaxf(v, lo), hi);                // np.clip(image, imin, imax) — in both branches
    if (span == span) {
      v = __fdiv_rn(__fsub_rn(v, lo), span);    // (image - imin) / (imax - imin)
    } else {                                    // imin == imax: every voxel is imin, then np.clip(image, omin, omax)
      v = fminf(fmaxf(v, 0.f), 1.f);
    }
  }
  v = __fmul_rn(v, sizeof(T) == 2 ? 65535.0f : 255.0f);   // *= 2**16 - 1 (or 255)
  out[(static_cast<size_t>(blockIdx.z) * Y + y) * X + x] = static_cast<T>(static_cast<int>(v));  // truncation
}

template <typename T>
static int rescale_crop_t(const float* vol, int vol_z0, const int* padded, const int* size, const float* norm3,
                          int z_begin, int z_count, T* out, cudaStream_t stream) {
  if (z_count <= 0) return 0;
  if (z_count > 65535 || size[1] > 65535) return set_error("rescale_u16_crop: slab too large for one launch");
  if (z_begin < vol_z0 || z_begin + z_count > size[0]) return set_error("rescale_u16_crop: plane range out of bounds");
  dim3 grid((size[2] + 255) / 256, size[1], z_count);
  rescale_crop_kernel<T><<<grid, 256, 0, stream>>>(vol, vol_z0, padded[1], padded[2], size[1], size[2], norm3, z_begin,
                                                   out);
  NC_CUDA(cudaGetLastError());
  return 0;
}
int rescale_u16_crop(const float* vol, int vol_z0, const int* padded, const int* size, const float* norm3,
                     int z_begin, int z_count, uint16_t* out, cudaStream_t stream) {
  return rescale_crop_t(vol, vol_z0, padded, size, norm3, z_begin, z_count, out, stream);
}
int rescale_u8_crop(const float* vol, int vol_z0, const int* padded, const int* size, const float* norm3, int z_begin,
                    int z_count, uint8_t* out, cudaStream_t stream) {
  return rescale_crop_t(vol, vol_z0, padded, size, norm3, z_begin, z_count, out, stream);
}

// ------------------------------------------------------------------------------------------------ MIP
__global__ void mip_fwd_kernel(const float* __restrict__ vol, int D, int H, int W, int axis, int start, int depth,
                               float* __restrict__ proj, int* __restrict__ argmax) {
  const int n0 = axis == 0 ? H : D, n1 = axis == 2 ? H : W;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n0 * n1) return;
  const int i = idx / n1, j = idx - i * n1;
  size_t base, stride;
  if (axis == 0) {
    base = static_cast<size_t>(i) * W + j, stride = static_cast<size_t>(H) * W;
  } else if (axis == 1) {
    base = static_cast<size_t>(i) * H * W + j, stride = W;
  } else {
    base = (static_cast<size_t>(i) * H + j) * W, stride = 1;
  }
  float best = vol[base + start * stride];
  int arg = start;
  for (int k = 1; k < depth; ++k) {
    const float v = vol[base + (start + k) * stride];
    if (v > best) {  // first maximal index wins, as torch.max(dim)
      best = v;
      arg = start + k;
    }
  }
  proj[idx] = best;
  if (argmax) argmax[idx] = arg;
}

__global__ void mip_bwd_kernel(const float* __restrict__ gproj, const int* __restrict__ argmax, int D, int H, int W,
                               int axis, float* __restrict__ gvol) {
  const int n0 = axis == 0 ? H : D, n1 = axis == 2 ? H : W;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n0 * n1) return;
  const int i = idx / n1, j = idx - i * n1;
  const int k = argmax[idx];
  size_t off;
  if (axis == 0)
    off = (static_cast<size_t>(k) * H + i) * W + j;
  else if (axis == 1)
    off = (static_cast<size_t>(i) * H + k) * W + j;
  else
    off = (static_cast<size_t>(i) * H + j) * W + k;
  gvol[off] += gproj[idx];
}

int mip_fwd(const float* vol, int D, int H, int W, int axis, int start, int depth, float* proj, int* argmax,
            cudaStream_t stream) {
  if (axis < 0 || axis > 2) return set_error("mip: axis must be 0, 1 or 2");
  const int len = axis == 0 ? D : axis == 1 ? H : W;
  if (depth < 1 || start < 0 || start + depth > len) return set_error("mip: slab [start, start+depth) out of range");
  const int n = (axis == 0 ? H : D) * (axis == 2 ? H : W);
  mip_fwd_kernel<<<(n + 255) / 256, 256, 0, stream>>>(vol, D, H, W, axis, start, depth, proj, argmax);
  NC_CUDA(cudaGetLastError());
  return 0;
}
int mip_bwd(const float* gproj, const int* argmax, int D, int H, int W, int axis, float* gvol, cudaStream_t stream) {
  if (axis < 0 || axis > 2) return set_error("mip: axis must be 0, 1 or 2");
  const int n = (axis == 0 ? H : D) * (axis == 2 ? H : W);
  mip_bwd_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gproj, argmax, D, H, W, axis, gvol);
  NC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nc

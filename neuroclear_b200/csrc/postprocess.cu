// The remaining Assemble_Dice / test_dice.py options: --histogram_match, --save_projections and
// the whole-volume PSNR report.  All memory-bound; see include/neuroclear_b200.h for the contract of each symbol.
#include <cub/device/device_radix_sort.cuh>

#include "../../include/neuroclear_b200.h"
#include "internal.h"

using namespace nc;

namespace {
inline cudaStream_t S(nc_stream_t s) { return static_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------------------------------------ histogram match
// skimage.exposure.match_histograms(image=fake, reference=real) for two arrays of the SAME size n (a border-cut cube
// of roi^3 voxels), skimage/exposure/histogram_matching.py::_match_cumulative_cdf:
//     src_values, src_idx, src_counts = np.unique(source, return_inverse, return_counts)
//     tmpl_values, tmpl_counts        = np.unique(template, return_counts)
//     src_q  = np.cumsum(src_counts)  / source.size          (float64)
//     tmpl_q = np.cumsum(tmpl_counts) / template.size
//     out    = np.interp(src_q, tmpl_q, tmpl_values)[src_idx]           (float64)
// With both arrays sorted (S, T): a voxel of value v has src_q = cnt / n, cnt = #{S <= v} = upper_bound(S, v).
// The template's knots are xp[j] = end_j / n at the END (exclusive) of every run of equal values of T, fp[j] = that
// value.  Same n on both sides, so knot comparisons are integer comparisons: j = the last run with end_j <= cnt.
// np.interp (numpy/core/src/multiarray/compiled_base.c::arr_interp): below the first knot -> fp[0]; on the last knot
// or exactly on a knot -> fp[j]; else slope = (fp[j+1] - fp[j]) / (xp[j+1] - xp[j]);  slope * (x - xp[j]) + fp[j],
// every operation rounded separately in float64 (no FMA).
__device__ __forceinline__ int upper_bound_f(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int lower_bound_f(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
hist_match_map_kernel(const float* __restrict__ src, const float* __restrict__ S, const float* __restrict__ T, int n,
                      double* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float v = src[i];
  const int cnt = upper_bound_f(S, n, v);                 // >= 1
  const double dn = static_cast<double>(n);
  const double x = __ddiv_rn(static_cast<double>(cnt), dn);
  // run of T that contains position cnt-1
  const float tv = T[cnt - 1];
  const int end_r = upper_bound_f(T, n, tv);
  int end_j;                                              // exclusive end of knot j = last run with end <= cnt
  if (end_r == cnt) {
    end_j = end_r;
  } else {
    end_j = lower_bound_f(T, n, tv);                      // = end of the previous run (0 when there is none)
    if (end_j == 0) {                                     // x < xp[0]: np.interp's default left = fp[0]
      out[i] = static_cast<double>(T[0]);
      return;
    }
  }
  const double fj = static_cast<double>(T[end_j - 1]);
  if (end_j == n || end_j == cnt) {                       // last knot, or exactly on a knot
    out[i] = fj;
    return;
  }
  const float tn = T[end_j];
  const int end_n = upper_bound_f(T, n, tn);
  const double xj = __ddiv_rn(static_cast<double>(end_j), dn), xn = __ddiv_rn(static_cast<double>(end_n), dn);
  const double slope = __ddiv_rn(__dsub_rn(static_cast<double>(tn), fj), __dsub_rn(xn, xj));
  out[i] = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xj)), fj);
}

// ------------------------------------------------------------------------------------------------ blend (f64 pieces)
struct Cover {
  int j0, j1;
};
__device__ __forceinline__ Cover cover(int q, int step, int roi, int n) {
  Cover c;
  c.j1 = min(q / step, n - 1);
  c.j0 = (c.j1 >= 1 && q < (c.j1 - 1) * step + roi) ? c.j1 - 1 : -1;
  return c;
}

// assemble_dice.py:167-184 when the queue holds float64 cubes (match_histograms returns float64): numpy evaluates
// `visual_ret[...] += cube / 8` with the float64 loop and stores float32, i.e. acc = f32(f64(acc) + v / 8).
__global__ void __launch_bounds__(256)
blend_gather_f64_kernel(const double* __restrict__ pieces, const long long* __restrict__ piece_off,
                        const int* __restrict__ piece_z0, int Py, int Px, int nz, int ny, int nx, int roi, int step,
                        int out_z0, float* __restrict__ out) {
  const int x = blockIdx.x * 256 + threadIdx.x;
  if (x >= Px) return;
  const int y = blockIdx.y;
  const int z = out_z0 + blockIdx.z;
  const Cover cz = cover(z, step, roi, nz), cy = cover(y, step, roi, ny), cx = cover(x, step, roi, nx);
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
        const double v = pieces[off + (static_cast<long long>(lz) * roi + (y - jy * step)) * roi + (x - jx * step)];
        acc = __double2float_rn(__dadd_rn(static_cast<double>(acc), __dmul_rn(v, 0.125)));
        ++n;
      }
    }
  }
  out[(static_cast<size_t>(blockIdx.z) * Py + y) * Px + x] = __fdiv_rn(acc, static_cast<float>(n)) * 8.0f;
}

// ------------------------------------------------------------------------------------------------ projections
// np.amax(volume[a0:a1 along `axis`], axis) of a (Z,Y,X) uint16 / uint8 volume (test_dice.py:159-177).
template <typename T>
__global__ void __launch_bounds__(256)
amax_axis_kernel(const T* __restrict__ vol, int Z, int Y, int X, int axis, int a0, int a1, T* __restrict__ out) {
  const int n1 = axis == 2 ? Y : X;                       // inner extent of the projection image
  const int n0 = axis == 0 ? Y : Z;
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= static_cast<long long>(n0) * n1) return;
  const int i = static_cast<int>(idx / n1), j = static_cast<int>(idx - static_cast<long long>(i) * n1);
  size_t base, stride;
  if (axis == 0) {
    base = static_cast<size_t>(i) * X + j, stride = static_cast<size_t>(Y) * X;
  } else if (axis == 1) {
    base = static_cast<size_t>(i) * Y * X + j, stride = X;
  } else {
    base = (static_cast<size_t>(i) * Y + j) * X, stride = 1;
  }
  T best = 0;
  for (int k = a0; k < a1; ++k) best = max(best, vol[base + k * stride]);
  out[idx] = best;
}

// ------------------------------------------------------------------------------------------------ PSNR report
// Exact integer moments of a uint16 / uint8 volume: {sum, sum of squares, min, max} as uint64 (n < 2^31 * 2 voxels
// of 16 bits: sum of squares < 2^64).  Deterministic: integer atomics commute.
template <typename T>
__global__ void __launch_bounds__(256)
moments_kernel(const T* __restrict__ v, long long n, unsigned long long* __restrict__ out) {
  unsigned long long s = 0, q = 0;
  unsigned int mn = 0xffffffffu, mx = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * 256) {
    const unsigned int x = v[i];
    s += x, q += static_cast<unsigned long long>(x) * x;
    mn = min(mn, x), mx = max(mx, x);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(~0u, s, o), q += __shfl_xor_sync(~0u, q, o);
    mn = min(mn, __shfl_xor_sync(~0u, mn, o)), mx = max(mx, __shfl_xor_sync(~0u, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 0, s), atomicAdd(out + 1, q);
    atomicMin(out + 2, static_cast<unsigned long long>(mn)), atomicMax(out + 3, static_cast<unsigned long long>(mx));
  }
}

// np.std's sum: numpy/core/src/umath/loops_utils.h.src::pairwise_sum over x_i = (double(v_i) - mean)^2, reproduced in
// numpy's exact association order (np.std of the whole volume = one contiguous float64 reduction of n elements):
//   n < 8: sequential;  n <= 128: eight interleaved accumulators, ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail;
//   else split at n2 = n/2 - (n/2) % 8 and add the two halves.
// The reference's double normalisation (test_dice.py:243-249) puts every voxel of its second pass exactly on an
// integer boundary of the truncating uint8 cast, so the LAST BIT of this sum decides the volume: any other summation
// order changes ~half of the voxels by one grey level.  Thread t owns the depth-`depth` node of that tree (>= 2048
// elements), evaluates it serially in numpy's order, and one block then folds the 2^depth partial sums pairwise.
template <typename T>
__device__ double pw_leaf(const T* __restrict__ v, long long off, int n, double mean) {
  auto X = [&](long long i) {
    const double d = __dsub_rn(static_cast<double>(v[off + i]), mean);
    return __dmul_rn(d, d);
  };
  if (n < 8) {
    double res = -0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, X(i));
    return res;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = X(j);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], X(i + j));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, X(i));
  return res;
}
template <typename T>
__device__ double pw_node(const T* __restrict__ v, long long off, long long n, double mean) {
  if (n <= 128) return pw_leaf(v, off, static_cast<int>(n), mean);
  long long n2 = n / 2;
  n2 -= n2 % 8;
  const double a = pw_node(v, off, n2, mean);
  const double b = pw_node(v, off + n2, n - n2, mean);
  return __dadd_rn(a, b);
}
template <typename T>
__global__ void __launch_bounds__(128)
pairwise_sqdev_kernel(const T* __restrict__ v, long long n, double mean, int depth, double* __restrict__ partial) {
  const long long t = static_cast<long long>(blockIdx.x) * 128 + threadIdx.x;
  if (t >= (1ll << depth)) return;
  long long off = 0, len = n;
  for (int b = depth - 1; b >= 0; --b) {
    long long n2 = len / 2;
    n2 -= n2 % 8;
    if ((t >> b) & 1) off += n2, len -= n2; else len = n2;
  }
  partial[t] = pw_node(v, off, len, mean);
}
__global__ void __launch_bounds__(1024) pairwise_fold_kernel(double* __restrict__ partial, int depth, double* out) {
  for (int lv = depth - 1; lv >= 0; --lv) {       // in place: node j of level lv = children 2j, 2j+1 of level lv+1
    const long long cnt = 1ll << lv;
    // read both children before anyone overwrites them: two phases per level
    for (long long base = 0; base < cnt; base += 1024) {
      const long long j = base + threadIdx.x;
      double s = 0.0;
      if (j < cnt) s = __dadd_rn(partial[2 * j], partial[2 * j + 1]);
      __syncthreads();
      if (j < cnt) partial[j] = s;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) out[0] = partial[0];
}

// util.normalize(util.standardize(v), np.uint8) (util/util.py:56-71,107-108) per voxel, float64 as numpy evaluates it:
//   s = (v - mean) / std;  out = uint8((s - s_min) * (255 / (s_max - s_min)) + 0)       (truncation)
// p = {mean, std, s_min, scale}; the four scalars come from the exact moments (host side, float64).
template <typename T>
__global__ void __launch_bounds__(256)
standardize_normalize_u8_kernel(const T* __restrict__ v, long long n, const double* __restrict__ p,
                                uint8_t* __restrict__ out) {
  const double mean = p[0], sd = p[1], smin = p[2], scale = p[3];
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * 256) {
    const double s = __ddiv_rn(__dsub_rn(static_cast<double>(v[i]), mean), sd);
    const double r = __dadd_rn(__dmul_rn(__dsub_rn(s, smin), scale), 0.0);
    out[i] = static_cast<uint8_t>(static_cast<int>(r));
  }
}

// sum over voxels of (target - source)^2 for two uint8 volumes: an exact integer.
__global__ void __launch_bounds__(256)
sqdiff_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, long long n,
                 unsigned long long* __restrict__ out) {
  unsigned long long s = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * 256) {
    const int d = static_cast<int>(a[i]) - static_cast<int>(b[i]);
    s += static_cast<unsigned long long>(d * d);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(~0u, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}
}  // namespace

extern "C" {

int64_t nc_hist_match_scratch_bytes(int64_t n) {
  if (n <= 0 || n > (1ll << 30)) return set_error("hist_match: n out of range");
  size_t tmp = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp, static_cast<const float*>(nullptr), static_cast<float*>(nullptr),
                                 static_cast<int>(n));
  return static_cast<int64_t>(2 * ((n * 4 + 255) / 256 * 256) + tmp + 256);
}

int nc_hist_match_f32(const float* fake, const float* real, int64_t n, void* scratch, int64_t scratch_bytes,
                      double* out, nc_stream_t stream) {
  const int64_t need = nc_hist_match_scratch_bytes(n);
  if (need < 0) return -1;
  if (scratch_bytes < need) return set_error("hist_match: scratch too small (%lld < %lld)", (long long)scratch_bytes,
                                             (long long)need);
  const size_t arr = static_cast<size_t>((n * 4 + 255) / 256 * 256);
  float* Ssorted = static_cast<float*>(scratch);
  float* Tsorted = reinterpret_cast<float*>(static_cast<char*>(scratch) + arr);
  void* tmp = static_cast<char*>(scratch) + 2 * arr;
  size_t tmp_bytes = static_cast<size_t>(scratch_bytes) - 2 * arr;
  NC_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, fake, Ssorted, static_cast<int>(n), 0, 32, S(stream)));
  NC_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, real, Tsorted, static_cast<int>(n), 0, 32, S(stream)));
  hist_match_map_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, S(stream)>>>(fake, Ssorted, Tsorted,
                                                                                      static_cast<int>(n), out);
  NC_CUDA(cudaGetLastError());
  return 0;
}

int nc_blend_gather_f64(const double* pieces, const int64_t* piece_off, const int32_t* piece_z0,
                        const int32_t padded_zyx[3], const int32_t steps_zyx[3], int32_t roi, int32_t overlap,
                        int32_t out_z0, int32_t out_nz, float* out, nc_stream_t stream) {
  if (overlap <= 0) return set_error("blend_gather: overlap must be > 0 (the reference produces zeros otherwise)");
  if (2 * overlap > roi) return set_error("blend_gather: overlap must be <= roi - overlap");
  if (out_nz <= 0) return 0;
  if (out_nz > 65535 || padded_zyx[1] > 65535) return set_error("blend_gather: slab too large for one launch");
  dim3 grid((padded_zyx[2] + 255) / 256, padded_zyx[1], out_nz);
  blend_gather_f64_kernel<<<grid, 256, 0, S(stream)>>>(pieces, reinterpret_cast<const long long*>(piece_off), piece_z0,
                                                       padded_zyx[1], padded_zyx[2], steps_zyx[0], steps_zyx[1],
                                                       steps_zyx[2], roi, roi - overlap, out_z0, out);
  NC_CUDA(cudaGetLastError());
  return 0;
}

int nc_amax_axis(const void* vol, int32_t elem_bytes, int32_t z, int32_t y, int32_t x, int32_t axis, int32_t a0,
                 int32_t a1, void* out, nc_stream_t stream) {
  if (axis < 0 || axis > 2 || (elem_bytes != 1 && elem_bytes != 2)) return set_error("amax_axis: bad axis / type");
  const int ext = axis == 0 ? z : axis == 1 ? y : x;
  a0 = a0 < 0 ? 0 : a0;                                   // numpy slice clamping
  a1 = a1 > ext ? ext : a1;
  if (a1 <= a0) return set_error("amax_axis: empty range (np.amax of an empty slice raises)");
  const long long n = static_cast<long long>(axis == 0 ? y : z) * (axis == 2 ? y : x);
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  if (elem_bytes == 2)
    amax_axis_kernel<uint16_t><<<blocks, 256, 0, S(stream)>>>(static_cast<const uint16_t*>(vol), z, y, x, axis, a0, a1,
                                                              static_cast<uint16_t*>(out));
  else
    amax_axis_kernel<uint8_t><<<blocks, 256, 0, S(stream)>>>(static_cast<const uint8_t*>(vol), z, y, x, axis, a0, a1,
                                                             static_cast<uint8_t*>(out));
  NC_CUDA(cudaGetLastError());
  return 0;
}

int nc_volume_moments(const void* vol, int32_t elem_bytes, int64_t n, uint64_t* out4, nc_stream_t stream) {
  if (elem_bytes != 1 && elem_bytes != 2) return set_error("volume_moments: uint8 / uint16 volumes only");
  const unsigned long long init[4] = {0, 0, ~0ull, 0};
  NC_CUDA(cudaMemcpyAsync(out4, init, sizeof(init), cudaMemcpyHostToDevice, S(stream)));
  const int blocks = num_sms() * 8;
  if (elem_bytes == 2)
    moments_kernel<uint16_t><<<blocks, 256, 0, S(stream)>>>(static_cast<const uint16_t*>(vol), n,
                                                            reinterpret_cast<unsigned long long*>(out4));
  else
    moments_kernel<uint8_t><<<blocks, 256, 0, S(stream)>>>(static_cast<const uint8_t*>(vol), n,
                                                           reinterpret_cast<unsigned long long*>(out4));
  NC_CUDA(cudaGetLastError());
  return 0;
}

int64_t nc_pairwise_sqdev_scratch_doubles(int64_t n) {
  int depth = 0;
  while ((n >> (depth + 1)) >= 2048) ++depth;
  return 1ll << depth;
}

int nc_pairwise_sqdev_sum(const void* vol, int32_t elem_bytes, int64_t n, double mean, double* scratch, double* out1,
                          nc_stream_t stream) {
  if (elem_bytes != 1 && elem_bytes != 2) return set_error("pairwise_sqdev_sum: uint8 / uint16 volumes only");
  if (n <= 0) return set_error("pairwise_sqdev_sum: empty volume");
  int depth = 0;
  while ((n >> (depth + 1)) >= 2048) ++depth;      // every node above `depth` has > 128 elements: internal in numpy
  static bool once[64] = {false};
  if (first_use_on_device(once)) NC_CUDA(cudaDeviceSetLimit(cudaLimitStackSize, 4096));
  const long long threads = 1ll << depth;
  const unsigned blocks = static_cast<unsigned>((threads + 127) / 128);
  if (elem_bytes == 2)
    pairwise_sqdev_kernel<uint16_t><<<blocks, 128, 0, S(stream)>>>(static_cast<const uint16_t*>(vol), n, mean, depth,
                                                                   scratch);
  else
    pairwise_sqdev_kernel<uint8_t><<<blocks, 128, 0, S(stream)>>>(static_cast<const uint8_t*>(vol), n, mean, depth,
                                                                  scratch);
  NC_CUDA(cudaGetLastError());
  pairwise_fold_kernel<<<1, 1024, 0, S(stream)>>>(scratch, depth, out1);
  NC_CUDA(cudaGetLastError());
  return 0;
}

int nc_standardize_normalize_u8(const void* vol, int32_t elem_bytes, int64_t n, const double* params4, uint8_t* out,
                                nc_stream_t stream) {
  if (elem_bytes != 1 && elem_bytes != 2) return set_error("standardize_normalize: uint8 / uint16 volumes only");
  const int blocks = num_sms() * 8;
  if (elem_bytes == 2)
    standardize_normalize_u8_kernel<uint16_t><<<blocks, 256, 0, S(stream)>>>(static_cast<const uint16_t*>(vol), n,
                                                                             params4, out);
  else
    standardize_normalize_u8_kernel<uint8_t><<<blocks, 256, 0, S(stream)>>>(static_cast<const uint8_t*>(vol), n,
                                                                            params4, out);
  NC_CUDA(cudaGetLastError());
  return 0;
}

int nc_sqdiff_u8(const uint8_t* a, const uint8_t* b, int64_t n, uint64_t* out1, nc_stream_t stream) {
  NC_CUDA(cudaMemsetAsync(out1, 0, 8, S(stream)));
  sqdiff_u8_kernel<<<num_sms() * 8, 256, 0, S(stream)>>>(a, b, n, reinterpret_cast<unsigned long long*>(out1));
  NC_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"

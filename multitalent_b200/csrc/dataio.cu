// The steps either side of the hot path, as HBM-streaming kernels on device-resident
// volumes: patch crop + pad out of a preprocessed case (dataset_loading.py:340-378), nearest-neighbour down-sampling of
// the label map to the deep-supervision scales (data_augmentation/downsampling.py:87-104), and the export-side
// resampling of the probability volume to the original grid with the in-order threshold fused
// (inference/segmentation_export.py:77-123, preprocessing/preprocessing.py:109-197).
//
// Coordinate convention of the resizing the reference delegates to skimage.transform.resize (un-vendored; order 0 / 1,
// mode 'edge', anti_aliasing False = scipy.ndimage.zoom(grid_mode=True, mode='nearest')): output index o samples the
// input at x = (o + 0.5) * in / out - 0.5 (pixel centres), clamped to the volume; order 0 takes floor(x + 0.5),
// order 1 interpolates linearly between floor(x) and floor(x) + 1.
#include "common.cuh"

namespace mtb {

// ---- crop + pad ------------------------------------------------------------------------------------------------------
// dst[c][i][j][k] = src[c][lb + (i,j,k)] inside the case, else pad (constant per channel, or the nearest edge voxel)
__global__ void crop_pad_kernel(const float* __restrict__ src, int C, int X, int Y, int Z, int lbx, int lby, int lbz,
                                float* __restrict__ dst, int pd, int ph, int pw, int edge_mode,
                                const float* __restrict__ pad_values) {
  const long long n = (long long)C * pd * ph * pw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % pw);
    long long t = i / pw;
    const int j = (int)(t % ph); t /= ph;
    const int ii = (int)(t % pd);
    const int c = (int)(t / pd);
    int x = lbx + ii, y = lby + j, z = lbz + k;
    const bool inside = x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z;
    float v;
    if (inside || edge_mode) {
      x = min(max(x, 0), X - 1); y = min(max(y, 0), Y - 1); z = min(max(z, 0), Z - 1);
      v = src[(((long long)c * X + x) * Y + y) * Z + z];
    } else {
      v = pad_values ? pad_values[c] : 0.f;
    }
    dst[i] = v;
  }
}

int crop_pad(const float* src, int C, int X, int Y, int Z, int lbx, int lby, int lbz, float* dst, int pd, int ph, int pw,
             int edge_mode, const float* pad_values, cudaStream_t s) {
  const long long n = (long long)C * pd * ph * pw;
  if (n == 0) return MTB200_OK;
  const int blocks = (int)min((long long)num_sms() * 16, (n + 255) / 256);
  crop_pad_kernel<<<blocks, 256, 0, s>>>(src, C, X, Y, Z, lbx, lby, lbz, dst, pd, ph, pw, edge_mode, pad_values);
  return check_launch("crop_pad");
}

// ---- nearest-neighbour resize of label maps ----------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int o, int n_in, int n_out) {
  // floor((o + 0.5) * in / out - 0.5 + 0.5), in exact integer arithmetic: floor((2 o + 1) * in / (2 out))
  const long long num = (long long)(2 * o + 1) * n_in;
  const int v = (int)(num / (2LL * n_out));
  return min(max(v, 0), n_in - 1);
}

__global__ void resize_nearest_kernel(const float* __restrict__ src, long long NC, int X, int Y, int Z,
                                      float* __restrict__ dst, int X2, int Y2, int Z2) {
  const long long n = NC * X2 * Y2 * Z2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Z2);
    long long t = i / Z2;
    const int j = (int)(t % Y2); t /= Y2;
    const int ii = (int)(t % X2);
    const long long c = t / X2;
    dst[i] = src[((c * X + nearest_src(ii, X, X2)) * Y + nearest_src(j, Y, Y2)) * Z + nearest_src(k, Z, Z2)];
  }
}

int resize_nearest(const float* src, long long NC, int X, int Y, int Z, float* dst, int X2, int Y2, int Z2, cudaStream_t s) {
  const long long n = NC * X2 * Y2 * Z2;
  if (n == 0) return MTB200_OK;
  const int blocks = (int)min((long long)num_sms() * 16, (n + 255) / 256);
  resize_nearest_kernel<<<blocks, 256, 0, s>>>(src, NC, X, Y, Z, dst, X2, Y2, Z2);
  return check_launch("resize_nearest");
}

// ---- probability resampling (per-axis order 0 / 1) + in-order threshold ------------------------------------------------
struct Axis { int i0, i1; float w1; };
__device__ __forceinline__ Axis axis_sample(int o, int n_in, int n_out, int order) {
  Axis a;
  if (order == 0 || n_in == n_out) {
    a.i0 = a.i1 = n_in == n_out ? o : nearest_src(o, n_in, n_out);
    a.w1 = 0.f;
    return a;
  }
  // double precision for the coordinate, as scipy's zoom computes it
  const double x = ((double)o + 0.5) * (double)n_in / (double)n_out - 0.5;
  const double fl = floor(x);
  a.w1 = (float)(x - fl);
  const int i = (int)fl;
  a.i0 = min(max(i, 0), n_in - 1);
  a.i1 = min(max(i + 1, 0), n_in - 1);
  return a;
}

// one thread per output voxel, loop over the channels in class order: prob (optional, fp32 or fp16) and
// seg[v] = class_order[last i with prob_i > 0.5] (segmentation_export.py:118-123) or, class_order == NULL, argmax
template <typename PT>
__global__ void resample_probs_kernel(const float* __restrict__ src, int C, int X, int Y, int Z, int X2, int Y2, int Z2,
                                      int ox, int oy, int oz, PT* __restrict__ prob, const float* __restrict__ class_order,
                                      unsigned char* __restrict__ seg) {
  const long long n = (long long)X2 * Y2 * Z2;
  const long long plane_in = (long long)Y * Z, vol_in = (long long)X * plane_in;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Z2);
    long long t = i / Z2;
    const int j = (int)(t % Y2);
    const int ii = (int)(t / Y2);
    const Axis ax = axis_sample(ii, X, X2, ox), ay = axis_sample(j, Y, Y2, oy), az = axis_sample(k, Z, Z2, oz);
    float sv = 0.f, best = -INFINITY;
    for (int c = 0; c < C; ++c) {
      const float* p = src + c * vol_in;
      auto at = [&](int x, int y, int z) { return p[x * plane_in + (long long)y * Z + z]; };
      // interpolate along z, then y, then x (the order skimage's separable spline filter applies is immaterial for
      // linear weights up to rounding)
      const float c00 = fmaf(az.w1, at(ax.i0, ay.i0, az.i1) - at(ax.i0, ay.i0, az.i0), at(ax.i0, ay.i0, az.i0));
      const float c01 = fmaf(az.w1, at(ax.i0, ay.i1, az.i1) - at(ax.i0, ay.i1, az.i0), at(ax.i0, ay.i1, az.i0));
      const float c10 = fmaf(az.w1, at(ax.i1, ay.i0, az.i1) - at(ax.i1, ay.i0, az.i0), at(ax.i1, ay.i0, az.i0));
      const float c11 = fmaf(az.w1, at(ax.i1, ay.i1, az.i1) - at(ax.i1, ay.i1, az.i0), at(ax.i1, ay.i1, az.i0));
      const float c0 = fmaf(ay.w1, c01 - c00, c00), c1 = fmaf(ay.w1, c11 - c10, c10);
      const float v = fmaf(ax.w1, c1 - c0, c0);
      if (prob) prob[(long long)c * n + i] = (PT)v;
      if (class_order) {
        if (v > 0.5f) sv = class_order[c];
      } else if (v > best) {
        best = v;
        sv = (float)c;
      }
    }
    if (seg) seg[i] = (unsigned char)sv;
  }
}

int resample_probs(const float* src, int C, int X, int Y, int Z, int X2, int Y2, int Z2, int ox, int oy, int oz,
                   void* prob, int prob_is_f16, const float* class_order, unsigned char* seg, cudaStream_t s) {
  const long long n = (long long)X2 * Y2 * Z2;
  if (n == 0) return MTB200_OK;
  const int blocks = (int)min((long long)num_sms() * 16, (n + 255) / 256);
  if (prob_is_f16)
    resample_probs_kernel<__half><<<blocks, 256, 0, s>>>(src, C, X, Y, Z, X2, Y2, Z2, ox, oy, oz,
                                                         reinterpret_cast<__half*>(prob), class_order, seg);
  else
    resample_probs_kernel<float><<<blocks, 256, 0, s>>>(src, C, X, Y, Z, X2, Y2, Z2, ox, oy, oz,
                                                        reinterpret_cast<float*>(prob), class_order, seg);
  return check_launch("resample_probs");
}

}  // namespace mtb

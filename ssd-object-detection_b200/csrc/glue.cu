// Input glue of the training path: what the reference's loaders do to every
// image's annotations and pixels before the matcher sees them, batched on the device.
//
//   gt_prepare_kernel   COCO [x, y, w, h] pixel boxes -> relative [cx, cy, w, h] float32:
//                         centre = corner + size / 2 in the annotations' own dtype
//                                                            data_loaders/coco/make_dataset.py:132
//                         stored as float32 (the generator's TensorSpec, :140-142), then divided IN PLACE
//                         by the integer [w, h, w, h] of its image: NumPy evaluates float32 /= int64 in
//                         float64 and rounds once to float32        data_loaders/ssd/make_dataset.py:43-44
//   image_norm_kernel   (x - 0.5) * 2 in float32             models/ssd_model.py:214
#include "common.cuh"

namespace ssdg {

template <typename T>
__global__ void __launch_bounds__(256) gt_prepare_kernel(const T* __restrict__ xywh, const int* __restrict__ img_wh,
                                                         const int* __restrict__ offsets, int batch, long long rows,
                                                         float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  // image of this row: last b with offsets[b] <= r
  int lo = 0, hi = batch;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((long long)offsets[mid] <= r) lo = mid; else hi = mid;
  }
  const double sw = (double)img_wh[2 * lo], sh = (double)img_wh[2 * lo + 1];
  T x = xywh[4 * r], y = xywh[4 * r + 1];
  const T w = xywh[4 * r + 2], h = xywh[4 * r + 3];
  float cx, cy;
  if (sizeof(T) == 8) {
    cx = __double2float_rn(__dadd_rn((double)x, __dmul_rn((double)w, 0.5)));
    cy = __double2float_rn(__dadd_rn((double)y, __dmul_rn((double)h, 0.5)));
  } else {
    cx = __fadd_rn((float)x, __fmul_rn((float)w, 0.5f));
    cy = __fadd_rn((float)y, __fmul_rn((float)h, 0.5f));
  }
  float4 o;
  o.x = __double2float_rn(__ddiv_rn((double)cx, sw));
  o.y = __double2float_rn(__ddiv_rn((double)cy, sh));
  o.z = __double2float_rn(__ddiv_rn((double)(float)w, sw));
  o.w = __double2float_rn(__ddiv_rn((double)(float)h, sh));
  reinterpret_cast<float4*>(out)[r] = o;
}

__global__ void __launch_bounds__(256) image_norm_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         long long n) {
  const long long n4 = n >> 2;
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gstride = (long long)gridDim.x * blockDim.x;
  const bool vec = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
  long long done = 0;
  if (vec) {
    for (long long i = gtid; i < n4; i += gstride) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(in) + i);
      float4 o;
      o.x = __fmul_rn(__fsub_rn(v.x, 0.5f), 2.f); o.y = __fmul_rn(__fsub_rn(v.y, 0.5f), 2.f);
      o.z = __fmul_rn(__fsub_rn(v.z, 0.5f), 2.f); o.w = __fmul_rn(__fsub_rn(v.w, 0.5f), 2.f);
      __stcs(reinterpret_cast<float4*>(out) + i, o);
    }
    done = n4 << 2;
  }
  for (long long i = done + gtid; i < n; i += gstride) out[i] = __fmul_rn(__fsub_rn(in[i], 0.5f), 2.f);
}

// Generalised anchor tables: the two options SSD variants add to the reference's rule.
//   clip_kernel        priors clamped to [0, 1] component-wise (cx, cy, w, h), in place
//   loc_scale_kernel   encoded offsets times (sxy, sxy, swh, swh): 1/variance after encoding, variance before decoding
template <typename T>
__global__ void __launch_bounds__(256) clip_kernel(T* __restrict__ v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = v[i] < (T)0 ? (T)0 : (v[i] > (T)1 ? (T)1 : v[i]);
}

__global__ void __launch_bounds__(256) loc_scale_kernel(const float4* __restrict__ in, float4* __restrict__ out, long long rows,
                                                        float sxy, float swh) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float4 v = in[i];
  out[i] = make_float4(__fmul_rn(v.x, sxy), __fmul_rn(v.y, sxy), __fmul_rn(v.z, swh), __fmul_rn(v.w, swh));
}

}  // namespace ssdg

using namespace ssdg;

extern "C" int ssdg_priors_clip(void* priors, int32_t dtype, int64_t n_priors, void* stream) {
  if (n_priors < 0 || (dtype != SSDG_F32 && dtype != SSDG_F64)) return SSDG_ERR_ARG;
  if (n_priors == 0) return SSDG_OK;
  if (!priors) return SSDG_ERR_ARG;
  const long long n = n_priors * 4;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == SSDG_F64) clip_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((double*)priors, n);
  else clip_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)priors, n);
  SSDG_LAUNCH_CHECK();
  return SSDG_OK;
}

extern "C" int ssdg_loc_scale(const float* in, float* out, int64_t rows, float scale_xy, float scale_wh, void* stream) {
  if (rows < 0) return SSDG_ERR_ARG;
  if (rows == 0) return SSDG_OK;
  if (!in || !out) return SSDG_ERR_ARG;
  if (((uintptr_t)in | (uintptr_t)out) & 15) return SSDG_ERR_ALIGN;
  loc_scale_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)in, (float4*)out, rows,
                                                                                     scale_xy, scale_wh);
  SSDG_LAUNCH_CHECK();
  return SSDG_OK;
}

extern "C" int ssdg_gt_prepare(const void* xywh, int32_t dtype, const int32_t* img_wh, const int32_t* gt_offsets,
                               int64_t batch, int64_t rows, float* out_boxes, void* stream) {
  if (!img_wh || !gt_offsets || batch <= 0 || rows < 0 || batch > 0x7fffffff) return SSDG_ERR_ARG;
  if (dtype != SSDG_F32 && dtype != SSDG_F64) return SSDG_ERR_ARG;
  if (rows == 0) return SSDG_OK;
  if (!xywh || !out_boxes) return SSDG_ERR_ARG;
  if (((uintptr_t)out_boxes & 15) || ((uintptr_t)xywh & (dtype == SSDG_F64 ? 7 : 3))) return SSDG_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((rows + 255) / 256);
  if (dtype == SSDG_F64)
    gt_prepare_kernel<double><<<grid, 256, 0, st>>>((const double*)xywh, img_wh, gt_offsets, (int)batch, rows, out_boxes);
  else
    gt_prepare_kernel<float><<<grid, 256, 0, st>>>((const float*)xywh, img_wh, gt_offsets, (int)batch, rows, out_boxes);
  SSDG_LAUNCH_CHECK();
  return SSDG_OK;
}

extern "C" int ssdg_image_normalize(const float* in, float* out, int64_t n, void* stream) {
  if (n < 0) return SSDG_ERR_ARG;
  if (n == 0) return SSDG_OK;
  if (!in || !out) return SSDG_ERR_ARG;
  if (((uintptr_t)in | (uintptr_t)out) & 3) return SSDG_ERR_ALIGN;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  image_norm_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n);
  SSDG_LAUNCH_CHECK();
  return SSDG_OK;
}

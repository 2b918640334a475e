// Target rasteriser for sm_100a: the dense training targets DetectionLoss consumes,
// built on the GPU from per-sample object lists instead of being rasterised with numpy in the dataset
// worker and copied host->device every step (datasets/coco.py:168-215, utils/image.py:8-57).  For the
// default experiment shape this replaces a 6.3 MB H2D copy of batch['hm'] by a 58 KB copy of boxes.
//
// One CTA per object slot.  Per object, in the reference's arithmetic (float64, same operation order,
// no fused multiply-add):
//   clip the box to the map, h = y2-y1, w = x2-x1, skip unless h > 0 and w > 0      (coco.py:199-202)
//   radius = max(0, int(gaussian_radius(ceil(h), ceil(w))))                           (coco.py:203-204, image.py:8-28)
//   ct = float32((x1+x2)/2, (y1+y2)/2), ct_int = trunc(ct)                            (coco.py:205-208)
//   hm[cls] = max(hm[cls], exp(-(dx^2+dy^2)/(2 sigma^2))), sigma = (2 radius+1)/6,
//             over the (2 radius+1)^2 window clipped at the borders                  (image.py:31-57)
//   wh = (w, h), ind = cy*W + cx, reg = ct - ct_int, reg_mask = 1                     (coco.py:210-213)
// The max-blend is an integer atomicMax on the float bits (all values are >= 0, so the orders agree):
// commutative, hence identical to the reference's sequential np.maximum for any scheduling.
#include "common.cuh"

namespace cnh {

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// utils/image.py:8-28, operation for operation (the reference divides every root by 2, not by 2a)
__device__ double gaussian_radius_ref(double height, double width, double mo) {
  const double hw = add_rn(height, width), area = mul_rn(width, height);
  const double b1 = hw;
  const double c1 = div_rn(mul_rn(area, sub_rn(1.0, mo)), add_rn(1.0, mo));
  const double sq1 = __dsqrt_rn(sub_rn(mul_rn(b1, b1), mul_rn(mul_rn(4.0, 1.0), c1)));
  const double r1 = div_rn(add_rn(b1, sq1), 2.0);
  const double b2 = mul_rn(2.0, hw);
  const double c2 = mul_rn(mul_rn(sub_rn(1.0, mo), width), height);
  const double sq2 = __dsqrt_rn(sub_rn(mul_rn(b2, b2), mul_rn(mul_rn(4.0, 4.0), c2)));
  const double r2 = div_rn(add_rn(b2, sq2), 2.0);
  const double a3 = mul_rn(4.0, mo);
  const double b3 = mul_rn(mul_rn(-2.0, mo), hw);
  const double c3 = mul_rn(mul_rn(sub_rn(mo, 1.0), width), height);
  const double sq3 = __dsqrt_rn(sub_rn(mul_rn(b3, b3), mul_rn(mul_rn(4.0, a3), c3)));
  const double r3 = div_rn(add_rn(b3, sq3), 2.0);
  return fmin(fmin(r1, r2), r3);
}

__global__ void __launch_bounds__(kThreads)
raster_kernel(const cnh_raster_args a) {
  const int slot = blockIdx.x, b = slot / a.M, k = slot - b * a.M;
  const int tid = threadIdx.x;
  __shared__ int sh[5];                                      // valid, cx, cy, radius, cls
  if (tid == 0) {
    int valid = 0, cxi = 0, cyi = 0, radius = 0, cls = 0;
    float whx = 0.f, why = 0.f, rx = 0.f, ry = 0.f;
    if (k < a.n_obj[b]) {
      const float* bx = a.boxes + (long long)slot * 4;
      const double xmax = (double)(a.W - 1), ymax = (double)(a.H - 1);
      const double x1 = fmin(fmax((double)bx[0], 0.0), xmax), x2 = fmin(fmax((double)bx[2], 0.0), xmax);
      const double y1 = fmin(fmax((double)bx[1], 0.0), ymax), y2 = fmin(fmax((double)bx[3], 0.0), ymax);
      const double h = sub_rn(y2, y1), w = sub_rn(x2, x1);
      cls = a.classes[slot];
      if (h > 0.0 && w > 0.0 && cls >= 0 && cls < a.C) {
        const double r = gaussian_radius_ref(ceil(h), ceil(w), (double)a.min_overlap_num / (double)a.min_overlap_den);
        radius = r > 0.0 ? (int)r : 0;
        const float cx = (float)div_rn(add_rn(x1, x2), 2.0), cy = (float)div_rn(add_rn(y1, y2), 2.0);
        cxi = (int)cx;
        cyi = (int)cy;
        whx = (float)w;
        why = (float)h;
        rx = cx - (float)cxi;
        ry = cy - (float)cyi;
        valid = 1;
      }
    }
    sh[0] = valid; sh[1] = cxi; sh[2] = cyi; sh[3] = radius; sh[4] = cls;
    a.wh[(long long)slot * 2 + 0] = whx;
    a.wh[(long long)slot * 2 + 1] = why;
    a.reg[(long long)slot * 2 + 0] = rx;
    a.reg[(long long)slot * 2 + 1] = ry;
    a.ind[slot] = valid ? (long long)cyi * a.W + cxi : 0ll;
    a.reg_mask[slot] = (uint8_t)valid;
  }
  __syncthreads();
  if (!sh[0]) return;
  const int cx = sh[1], cy = sh[2], radius = sh[3];
  const int left = min(cx, radius), right = min(a.W - cx, radius + 1);
  const int top = min(cy, radius), bottom = min(a.H - cy, radius + 1);
  const int ww = left + right, hh = top + bottom;
  if (ww <= 0 || hh <= 0) return;
  const double sigma = div_rn((double)(2 * radius + 1), 6.0);
  const double denom = mul_rn(mul_rn(2.0, sigma), sigma);
  int* plane = reinterpret_cast<int*>(a.hm + ((long long)b * a.C + sh[4]) * a.H * a.W);
  for (int i = tid; i < ww * hh; i += kThreads) {
    const int yy = i / ww, xx = i - yy * ww;
    const int dx = xx - left, dy = yy - top;
    const double g = exp(div_rn(-(double)(dx * dx + dy * dy), denom));
    // (h < eps * h.max() -> 0, utils/image.py:36, never fires: the window ends at 3 sigma, g >= e^-9)
    atomicMax(plane + (long long)(cy + dy) * a.W + (cx + dx), __float_as_int((float)g));
  }
}

}  // namespace cnh

using namespace cnh;

extern "C" int cnh_raster_targets(const cnh_raster_args* a, cnh_stream_t stream) {
  CNH_REQUIRE(a != nullptr, CNH_E_NULL, "raster_targets: args is NULL");
  CNH_REQUIRE(a->B > 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->M > 0, CNH_E_SHAPE,
              "raster_targets: bad dims B=%d C=%d H=%d W=%d M=%d", a->B, a->C, a->H, a->W, a->M);
  CNH_REQUIRE((long long)a->B * a->M < (1ll << 31), CNH_E_SHAPE, "raster_targets: B*M too large");
  CNH_REQUIRE(a->boxes && a->classes && a->n_obj && a->hm && a->wh && a->reg && a->ind && a->reg_mask, CNH_E_NULL,
              "raster_targets: a required pointer is NULL");
  CNH_REQUIRE(a->min_overlap_den > 0 && a->min_overlap_num > 0 && a->min_overlap_num < a->min_overlap_den,
              CNH_E_UNSUPPORTED, "raster_targets: min_overlap %d/%d outside (0,1)", a->min_overlap_num, a->min_overlap_den);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CNH_CUDA(cudaMemsetAsync(a->hm, 0, sizeof(float) * (size_t)a->B * a->C * a->H * a->W, st));
  raster_kernel<<<(unsigned)(a->B * a->M), kThreads, 0, st>>>(*a);
  CNH_CUDA(cudaGetLastError());
  return CNH_OK;
}

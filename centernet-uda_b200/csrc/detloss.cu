// DetectionLoss forward+backward for sm_100a: sigmoid+clamp, penalty-reduced focal loss,
// masked gather-L1 heads (plain / sigmoid-angle / RAPiD-periodic), all gradients, in ONE
// persistent launch.  Replaces losses/centernet.py:7-95,98-133,192-223 and
// utils/tensor.py:5-25 of the reference (chains of ~100 eager ATen kernels).
//
// HBM-bound streaming work: every heat-map element is read once (logit + target, float4,
// ld.global.cs) and written once (clamped probability + gradient).  The gradient needs
// the batch-wide num_pos, which is only known after everything has been read:
//   * STASH    (small problems): the raw gradient of <= kStash chunks per CTA stays in
//               registers across a cooperative grid barrier, then is scaled and stored:
//               16 B/element, the algorithmic floor.
//   * PRECOUNT (large problems): phase 0 counts num_pos over the target only (forward
//               order), grid barrier, phase 1 does the full pass in REVERSE order so the
//               tail of the target is still in the 126 MB L2: <= 20 B/element.
//   * COUNT + MAIN: the same two phases as separate launches with the batch-wide
//               normalisers supplied by the caller (the all-reduce of a sharded run sits
//               between them).
//   * FWD: no gradients (validation under no_grad): 12 B/element.
// Reductions are deterministic: fixed-shape block trees -> per-chunk partials -> per-sample
// double sums in chunk order -> batch sum in sample order.  The per-sample partials are
// independent of grid size and of how samples are sharded over GPUs, so a sharded run
// reproduces the single-device loss bit for bit.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cnh {

constexpr int kVec = 4;                       // float4 per thread per chunk
constexpr int kElemsPerThread = kVec * 4;
constexpr int kChunk = kThreads * kElemsPerThread;   // 4096 heat-map elements
constexpr int kStash = 2;                     // chunks a CTA may keep in registers
constexpr int kPiece = kThreads * 16;         // 4096 floats of one regression plane

enum Mode { M_STASH = 0, M_PRECOUNT = 1, M_MAIN = 2, M_COUNT = 3, M_FWD = 4 };

struct WsHeader {                             // 64 bytes, zero between launches
  int npos_total;
  int cnt_total[CNH_MAX_HEADS];
  unsigned done;
  int pad[11];
};

struct Geo {
  int HW;
  long long CHW;
  int cps;                 // chunks per sample
  int n_chunks;            // B * cps
  int ppp;                 // pieces per regression plane
  int item0[CNH_MAX_HEADS + 1];   // first item of each head; head h owns B*D_h*ppp items
  int n_items;
  int n_count;             // B * n_heads
  int vec_planes;          // regression planes can be zero-filled with float4 stores
  float* chunk_sum;        // [n_chunks]
  int* chunk_npos;         // [n_chunks]
  float* item_l1;          // [n_items]
  float* item_ang;         // [n_items]
  WsHeader* hdr;
};

// ---- focal element ----------------------------------------------------------------------
// term = log(p)(1-p)^2 [gt==1]  or  log(1-p) p^2 (1-gt)^4 [gt<1]   (losses/centernet.py:82-84)
// graw = d(term)/dx with p = clamp(s): in*s(1-s)*d(term)/dp; for in-range s (p == s) this is
//        +((1-s)^3 - 2 s (1-s)^2 log s)            for gt == 1
//        -(1-gt)^4 (s^3 - 2 s^2 (1-s) log(1-s))    for gt <  1
// so one log and no second reciprocal per element.
template <bool FAST>
__device__ __forceinline__ void focal_elem(float x, float gt, float& p, float& term, float& graw,
                                           int& npos) {
  const float s = sigmoidf_<FAST>(x);
  p = clamp_prob(s);
  const bool pos = (gt == 1.0f);
  const bool neg = (gt < 1.0f);
  // explicit _rn intrinsics: no FMA contraction, so every schedule (STASH / PRECOUNT / MAIN / FWD)
  // produces bit-identical terms and gradients
  const float q = __fsub_rn(1.0f, p);
  const float arg = pos ? p : q;              // argument of the log
  const float a = pos ? q : p;                // the squared factor
  const float L = logf_<FAST>(arg);
  const float omg = __fsub_rn(1.0f, gt);
  float w4 = __fmul_rn(omg, omg);
  w4 = __fmul_rn(w4, w4);
  const float wgt = pos ? 1.0f : (neg ? w4 : 0.0f);
  const float a2 = __fmul_rn(a, a);
  term = __fmul_rn(__fmul_rn(L, a2), wgt);
  const float inner = __fmaf_rn(-__fmul_rn(__fmul_rn(2.0f, a2), arg), L, __fmul_rn(a2, a));
  graw = (p == s) ? (pos ? inner : -__fmul_rn(wgt, inner)) : 0.0f;   // clamp passes gradient inclusively
  npos += pos ? 1 : 0;
}

__device__ __forceinline__ void block_reduce2(float& s, int& n, float* red_f, int* red_i) {
  s = warp_sum(s);
  n = warp_sum(n);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) {
    red_f[warp] = s;
    red_i[warp] = n;
  }
  __syncthreads();
  float ts = red_f[0];
  int tn = red_i[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) {
    ts += red_f[w];
    tn += red_i[w];
  }
  s = ts;
  n = tn;
}

// One chunk of the heat map.  WRITE_GRAD: scale known, store the gradient now.
// Otherwise (STASH) the raw gradient is returned in `graw`.
template <bool NEED_GRAD, bool WRITE_GRAD, bool FAST, bool VEC>
__device__ __forceinline__ void focal_chunk(const cnh_detloss_args& a, const Geo& g, int chunk,
                                            float scale, float (&graw)[kElemsPerThread],
                                            float* red_f, int* red_i) {
  const int b = chunk / g.cps, j = chunk - b * g.cps;
  const long long in_sample = (long long)j * kChunk;
  const long long base = (long long)b * g.CHW + in_sample;
  const long long left = g.CHW - in_sample;
  const int n = left < kChunk ? (int)left : kChunk;
  const float* __restrict__ xp = a.hm_logits + base;
  const float* __restrict__ gp = a.hm_gt + base;
  float* __restrict__ pp = a.prob + base;

  float xs[kElemsPerThread], gs[kElemsPerThread];
#pragma unroll
  for (int v = 0; v < kVec; ++v) {
    const int off = v * kThreads * 4 + threadIdx.x * 4;
    if (VEC) {
      float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(2.f, 2.f, 2.f, 2.f);
      if (off < n) {
        x4 = ldg_stream(reinterpret_cast<const float4*>(xp + off));
        g4 = ldg_stream(reinterpret_cast<const float4*>(gp + off));
      }
      xs[4 * v + 0] = x4.x; xs[4 * v + 1] = x4.y; xs[4 * v + 2] = x4.z; xs[4 * v + 3] = x4.w;
      gs[4 * v + 0] = g4.x; gs[4 * v + 1] = g4.y; gs[4 * v + 2] = g4.z; gs[4 * v + 3] = g4.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = off + e < n;
        xs[4 * v + e] = ok ? __ldcs(xp + off + e) : 0.f;
        gs[4 * v + e] = ok ? __ldcs(gp + off + e) : 2.f;   // gt = 2: neither pos nor neg
      }
    }
  }
  float sum = 0.f;
  int npos = 0;
#pragma unroll
  for (int v = 0; v < kVec; ++v) {
    const int off = v * kThreads * 4 + threadIdx.x * 4;
    float ps[4], gr[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float term;
      focal_elem<FAST>(xs[4 * v + e], gs[4 * v + e], ps[e], term, gr[e], npos);
      sum = __fadd_rn(sum, term);
      if (NEED_GRAD) graw[4 * v + e] = WRITE_GRAD ? __fmul_rn(gr[e], scale) : gr[e];
    }
    if (VEC) {
      if (off < n) {
        *reinterpret_cast<float4*>(pp + off) = make_float4(ps[0], ps[1], ps[2], ps[3]);
        if (WRITE_GRAD)
          stg_stream(reinterpret_cast<float4*>(a.grad_hm + base + off),
                     make_float4(graw[4 * v], graw[4 * v + 1], graw[4 * v + 2], graw[4 * v + 3]));
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (off + e < n) {
          pp[off + e] = ps[e];
          if (WRITE_GRAD) a.grad_hm[base + off + e] = graw[4 * v + e];
        }
    }
  }
  block_reduce2(sum, npos, red_f, red_i);
  if (threadIdx.x == 0) {
    g.chunk_sum[chunk] = sum;
    g.chunk_npos[chunk] = npos;
  }
}

template <bool VEC>
__device__ __forceinline__ void focal_store_stash(const cnh_detloss_args& a, const Geo& g, int chunk,
                                                  float scale, const float (&graw)[kElemsPerThread]) {
  const int b = chunk / g.cps, j = chunk - b * g.cps;
  const long long in_sample = (long long)j * kChunk;
  const long long base = (long long)b * g.CHW + in_sample;
  const long long left = g.CHW - in_sample;
  const int n = left < kChunk ? (int)left : kChunk;
#pragma unroll
  for (int v = 0; v < kVec; ++v) {
    const int off = v * kThreads * 4 + threadIdx.x * 4;
    if (VEC) {
      if (off < n)
        stg_stream(reinterpret_cast<float4*>(a.grad_hm + base + off),
                   make_float4(__fmul_rn(graw[4 * v], scale), __fmul_rn(graw[4 * v + 1], scale),
                               __fmul_rn(graw[4 * v + 2], scale), __fmul_rn(graw[4 * v + 3], scale)));
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (off + e < n) a.grad_hm[base + off + e] = __fmul_rn(graw[4 * v + e], scale);
    }
  }
}

// num_pos of one chunk from the target only (phase 0 of PRECOUNT / COUNT).
template <bool VEC>
__device__ __forceinline__ int count_chunk(const cnh_detloss_args& a, const Geo& g, int chunk) {
  const int b = chunk / g.cps, j = chunk - b * g.cps;
  const long long in_sample = (long long)j * kChunk;
  const long long left = g.CHW - in_sample;
  const int n = left < kChunk ? (int)left : kChunk;
  const float* __restrict__ gp = a.hm_gt + (long long)b * g.CHW + in_sample;
  int npos = 0;
#pragma unroll
  for (int v = 0; v < kVec; ++v) {
    const int off = v * kThreads * 4 + threadIdx.x * 4;
    if (VEC) {
      if (off < n) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gp + off));   // keep in L2
        npos += (g4.x == 1.f) + (g4.y == 1.f) + (g4.z == 1.f) + (g4.w == 1.f);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (off + e < n) npos += (__ldg(gp + off + e) == 1.f);
    }
  }
  return npos;
}

// ---- masked gather-L1 heads -----------------------------------------------------------------
struct ItemRef {
  int h, b, d, p0, p1;
};
__device__ __forceinline__ ItemRef decode_item(const cnh_detloss_args& a, const Geo& g, int item) {
  ItemRef r;
  r.h = 0;
#pragma unroll
  for (int h = 1; h < CNH_MAX_HEADS; ++h)
    if (h < a.n_heads && item >= g.item0[h]) r.h = h;
  const int local = item - g.item0[r.h];
  const int piece = local % g.ppp;
  const int plane = local / g.ppp;
  const int D = a.heads[r.h].D;
  r.d = plane % D;
  r.b = plane / D;
  r.p0 = piece * kPiece;
  r.p1 = min(g.HW, r.p0 + kPiece);
  return r;
}

__device__ __forceinline__ void l1_zero_fill(const cnh_detloss_args& a, const Geo& g, const ItemRef& r) {
  const cnh_head& hd = a.heads[r.h];
  float* __restrict__ dst = hd.grad + ((long long)r.b * hd.D + r.d) * g.HW;
  if (g.vec_planes) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int off = r.p0 + v * kThreads * 4 + threadIdx.x * 4;
      if (off < r.p1) *reinterpret_cast<float4*>(dst + off) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    for (int off = r.p0 + threadIdx.x; off < r.p1; off += kThreads) dst[off] = 0.f;
  }
}

// Forward terms (and, when SCATTER, the gradient scatter) of the object slots whose centre
// falls inside this item's piece of the plane.  inv_denom = 1 / (sum(mask_expanded) + 1e-4).
template <bool FORWARD, bool SCATTER, bool FAST>
__device__ __forceinline__ void l1_slots(const cnh_detloss_args& a, const Geo& g, const ItemRef& r,
                                         float inv_denom, float& l1_acc, float& ang_acc) {
  const cnh_head& hd = a.heads[r.h];
  const int D = hd.D;
  const bool is_angle = (D == 3 && r.d == 2 && hd.angle_mode != CNH_ANGLE_NONE);
  const float* __restrict__ plane = hd.map + ((long long)r.b * D + r.d) * g.HW;
  float* __restrict__ gplane = SCATTER ? hd.grad + ((long long)r.b * D + r.d) * g.HW : nullptr;
  constexpr float kPi = 3.14159265358979323846f;          // float(np.pi)
  constexpr float kHalfPi = 1.57079632679489661923f;      // float(np.pi / 2)
  constexpr float kDeg = 0.017453292519943295f;           // torch.deg2rad constant
  for (int k = threadIdx.x; k < a.M; k += kThreads) {
    const long long slot = (long long)r.b * a.M + k;
    const long long i = a.ind[slot];
    if (i < r.p0 || i >= r.p1) continue;
    const float m = (float)(hd.elementwise_mask ? hd.mask[slot * D + r.d] : hd.mask[slot]);
    const float pv = plane[i] * m;                          // pred *= mask   (centernet.py:108)
    const float tv = hd.target[slot * D + r.d] * m;         // target *= mask (centernet.py:109)
    float val, gcoef;                                       // |.| term and d(term)/d(pred*m)
    if (!is_angle) {
      const float diff = pv - tv;
      val = fabsf(diff);
      gcoef = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
    } else {
      const float s = sigmoidf_<FAST>(pv);
      const float sc = clamp_prob(s);
      const float ds = (sc == s) ? s * (1.f - s) : 0.f;
      if (hd.angle_mode == CNH_ANGLE_SIGMOID) {             // centernet.py:112-126
        const float diff = sc - clamp_prob(sigmoidf_<FAST>(tv));
        val = fabsf(diff);
        gcoef = ((diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f)) * ds;
      } else {                                              // centernet.py:203-220
        const float pa = sc * 2.f * kPi - kPi;
        const float ta = tv * kDeg;
        float rem = fmodf((pa - ta) - kHalfPi, kPi);
        if (rem != 0.f && rem < 0.f) rem += kPi;
        const float diff = rem - kHalfPi;
        val = fabsf(diff);
        gcoef = ((diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f)) * 2.f * kPi * ds;
      }
    }
    if (FORWARD) {
      if (is_angle) ang_acc += val; else l1_acc += val;
    }
    if (SCATTER) {
      const float w = is_angle ? hd.angle_weight : hd.weight;
      const float gv = gcoef * m * w * inv_denom;
      if (gv != 0.f) atomicAdd(gplane + i, gv);             // duplicates of `ind` accumulate
    }
  }
}

__device__ __forceinline__ void l1_store_partials(const Geo& g, int item, float l1, float ang,
                                                  float* red_f) {
  l1 = block_sum(l1, red_f);
  ang = block_sum(ang, red_f);
  if (threadIdx.x == 0) {
    g.item_l1[item] = l1;
    g.item_ang[item] = ang;
  }
}

// sum(mask_expanded) of one (head, sample): D * sum(mask[b,:]) or sum(mask[b,:,:]).
__device__ __forceinline__ void count_unit(const cnh_detloss_args& a, const Geo& g, int unit, int* red_i,
                                           bool add_total) {
  const int h = unit / a.B, b = unit - h * a.B;
  const cnh_head& hd = a.heads[h];
  const int n = hd.elementwise_mask ? a.M * hd.D : a.M;
  const uint8_t* __restrict__ mp = hd.mask + (long long)b * n;
  int c = 0;
  for (int k = threadIdx.x; k < n; k += kThreads) c += mp[k];
  c = block_sum(c, red_i);
  if (!hd.elementwise_mask) c *= hd.D;
  if (threadIdx.x == 0) {
    a.partials[(long long)b * CNH_PARTIALS + 4 + 3 * h] = (double)c;
    if (add_total) atomicAdd(&g.hdr->cnt_total[h], c);
  }
}

// ---- finalisation -------------------------------------------------------------------------
// partials rows -> scalars, mirroring the reference's fp32 arithmetic
// (centernet.py:91-95,119-131,213-222,42).  Run by one CTA; columns are summed in row order.
__device__ void combine_partials(const cnh_detloss_args& a, const double* part, int Btot, float* out,
                                 double* sh /* [CNH_PARTIALS] */) {
  if (threadIdx.x < CNH_PARTIALS) {
    double s = 0.0;   // rows may have been written by this very launch: read through L2
    for (int b = 0; b < Btot; ++b) s += __ldcg(part + (long long)b * CNH_PARTIALS + threadIdx.x);
    sh[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float fsum = (float)sh[0], npos = (float)sh[1];
    float hm = (npos == 0.f) ? (0.f - fsum) : (0.f - fsum / npos);
    hm *= a.hm_weight;
    float total = hm;
    out[1] = hm;
    for (int h = 0; h < CNH_MAX_HEADS; ++h) {
      float l = 0.f;
      if (h < a.n_heads) {
        const cnh_head& hd = a.heads[h];
        const float denom = (float)sh[4 + 3 * h] + 1e-4f;
        l = (float)sh[2 + 3 * h] / denom * hd.weight;
        if (hd.D == 3 && hd.angle_mode != CNH_ANGLE_NONE)
          l += (float)sh[3 + 3 * h] / denom * hd.angle_weight;
        total += l;
      }
      out[2 + h] = l;
    }
    out[0] = total;
    out[5] = npos;
    out[6] = 0.f;
    out[7] = 0.f;
  }
}

// chunk / item partials -> per-sample rows (double), one warp per sample.
__device__ void build_partials(const cnh_detloss_args& a, const Geo& g, bool with_items) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = warp; b < a.B; b += kWarps) {
    double s = 0.0;
    int n = 0;
    for (int j = lane; j < g.cps; j += 32) {
      s += (double)__ldcg(g.chunk_sum + (long long)b * g.cps + j);
      n += __ldcg(g.chunk_npos + (long long)b * g.cps + j);
    }
    s = warp_sum(s);
    n = warp_sum(n);
    double* row = a.partials + (long long)b * CNH_PARTIALS;
    if (lane == 0) {
      row[0] = s;
      row[1] = (double)n;
      row[11] = 0.0;
    }
    for (int h = 0; h < CNH_MAX_HEADS; ++h) {
      double l1 = 0.0, ang = 0.0;
      if (with_items && h < a.n_heads) {
        const int per = a.heads[h].D * g.ppp;
        const int first = g.item0[h] + b * per;
        for (int t = lane; t < per; t += 32) {
          l1 += (double)__ldcg(g.item_l1 + first + t);
          ang += (double)__ldcg(g.item_ang + first + t);
        }
        l1 = warp_sum(l1);
        ang = warp_sum(ang);
      }
      if (lane == 0) {
        row[2 + 3 * h] = l1;
        row[3 + 3 * h] = ang;
        if (h >= a.n_heads) row[4 + 3 * h] = 0.0;
      }
    }
  }
}

template <int MODE, bool FAST, bool VEC>
__global__ void __launch_bounds__(kThreads, 3)
detloss_kernel(const cnh_detloss_args a, const Geo g) {
  __shared__ float red_f[kWarps];
  __shared__ int red_i[kWarps];
  __shared__ double sh_cols[CNH_PARTIALS];
  __shared__ unsigned sh_ticket;
  const int bid = blockIdx.x, grid = gridDim.x;
  constexpr bool kGrad = (MODE == M_STASH || MODE == M_PRECOUNT || MODE == M_MAIN);

  if (MODE == M_STASH) {
    // ---- phase 1: everything that does not need the normalisers -----------------------
    float stash[kStash][kElemsPerThread];
#pragma unroll
    for (int r = 0; r < kStash; ++r) {
      const int chunk = bid + r * grid;
      if (chunk < g.n_chunks) {
        focal_chunk<true, false, FAST, VEC>(a, g, chunk, 0.f, stash[r], red_f, red_i);
        if (threadIdx.x == 0) atomicAdd(&g.hdr->npos_total, g.chunk_npos[chunk]);
      }
    }
    const int n_other = g.n_items + g.n_count;
    // regression work is dealt from the LAST CTA backwards: those hold the fewest chunks
    for (int o = grid - 1 - bid; o < n_other; o += grid) {
      if (o < g.n_items) {
        const ItemRef r = decode_item(a, g, o);
        l1_zero_fill(a, g, r);
        float l1 = 0.f, ang = 0.f;
        l1_slots<true, false, FAST>(a, g, r, 0.f, l1, ang);
        l1_store_partials(g, o, l1, ang, red_f);
      } else {
        count_unit(a, g, o - g.n_items, red_i, true);
      }
    }
    __threadfence();
    cg::this_grid().sync();
    // ---- phase 2: normalisers are final ------------------------------------------------
    const int npos = __ldcg(&g.hdr->npos_total);
    const float scale = (npos == 0) ? -a.hm_weight : -a.hm_weight / (float)npos;
#pragma unroll
    for (int r = 0; r < kStash; ++r) {
      const int chunk = bid + r * grid;
      if (chunk < g.n_chunks) focal_store_stash<VEC>(a, g, chunk, scale, stash[r]);
    }
    for (int o = grid - 1 - bid; o < g.n_items; o += grid) {
      const ItemRef r = decode_item(a, g, o);
      const float inv = 1.f / ((float)__ldcg(&g.hdr->cnt_total[r.h]) + 1e-4f);
      float l1 = 0.f, ang = 0.f;
      l1_slots<false, true, FAST>(a, g, r, inv, l1, ang);
    }
  } else {
    float scale = 0.f;
    float inv_denom[CNH_MAX_HEADS] = {0.f, 0.f, 0.f};
    if (MODE == M_PRECOUNT || MODE == M_COUNT) {
      // ---- phase 0: normalisers from the targets only ----------------------------------
      int cta_npos = 0;
      for (int chunk = bid; chunk < g.n_chunks; chunk += grid) {
        int c = count_chunk<VEC>(a, g, chunk);
        c = block_sum(c, red_i);
        if (MODE == M_COUNT && threadIdx.x == 0) g.chunk_npos[chunk] = c;
        cta_npos += c;
      }
      if (threadIdx.x == 0 && cta_npos) atomicAdd(&g.hdr->npos_total, cta_npos);
      for (int u = grid - 1 - bid; u < g.n_count; u += grid) count_unit(a, g, u, red_i, true);
    }
    if (MODE == M_PRECOUNT) {
      __threadfence();
      cg::this_grid().sync();
      const int npos = __ldcg(&g.hdr->npos_total);
      scale = (npos == 0) ? -a.hm_weight : -a.hm_weight / (float)npos;
      for (int h = 0; h < a.n_heads; ++h)
        inv_denom[h] = 1.f / ((float)__ldcg(&g.hdr->cnt_total[h]) + 1e-4f);
    }
    if (MODE == M_MAIN) {
      const float npos = (float)a.norm[0];
      scale = (npos == 0.f) ? -a.hm_weight : -a.hm_weight / npos;
      for (int h = 0; h < a.n_heads; ++h) inv_denom[h] = 1.f / ((float)a.norm[1 + h] + 1e-4f);
    }
    if (MODE != M_COUNT) {
      // ---- the streaming pass (reverse chunk order after a pre-count: L2 reuse) ---------
      float unused[kElemsPerThread];
      for (int u = bid; u < g.n_chunks; u += grid) {
        const int chunk = (MODE == M_PRECOUNT) ? g.n_chunks - 1 - u : u;
        focal_chunk<kGrad, kGrad, FAST, VEC>(a, g, chunk, scale, unused, red_f, red_i);
      }
      const int n_other = g.n_items + (MODE == M_FWD ? g.n_count : 0);
      for (int o = grid - 1 - bid; o < n_other; o += grid) {
        if (o < g.n_items) {
          const ItemRef r = decode_item(a, g, o);
          float l1 = 0.f, ang = 0.f;
          if (kGrad) {
            l1_zero_fill(a, g, r);
            __syncthreads();
            l1_slots<true, true, FAST>(a, g, r, inv_denom[r.h], l1, ang);
          } else {
            l1_slots<true, false, FAST>(a, g, r, 0.f, l1, ang);
          }
          l1_store_partials(g, o, l1, ang, red_f);
        } else {
          count_unit(a, g, o - g.n_items, red_i, false);
        }
      }
    }
  }

  // ---- last CTA: per-sample partials, scalars, reset the workspace counters -------------
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) sh_ticket = atomicAdd(&g.hdr->done, 1u);
  __syncthreads();
  if (sh_ticket != (unsigned)(grid - 1)) return;
  __threadfence();
  if (MODE == M_COUNT) {
    // per-sample num_pos rows + this shard's totals
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = warp; b < a.B; b += kWarps) {
      int n = 0;
      for (int j = lane; j < g.cps; j += 32) n += __ldcg(g.chunk_npos + (long long)b * g.cps + j);
      n = warp_sum(n);
      if (lane == 0) a.partials[(long long)b * CNH_PARTIALS + 1] = (double)n;
    }
    if (threadIdx.x == 0) {
      a.norm_out[0] = (double)__ldcg(&g.hdr->npos_total);
      for (int h = 0; h < CNH_MAX_HEADS; ++h)
        a.norm_out[1 + h] = (h < a.n_heads) ? (double)__ldcg(&g.hdr->cnt_total[h]) : 0.0;
    }
  } else {
    build_partials(a, g, true);
    __threadfence();
    __syncthreads();
    if (a.scalars != nullptr) combine_partials(a, a.partials, a.B, a.scalars, sh_cols);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    g.hdr->npos_total = 0;
    for (int h = 0; h < CNH_MAX_HEADS; ++h) g.hdr->cnt_total[h] = 0;
    g.hdr->done = 0;
  }
}

__global__ void __launch_bounds__(kThreads)
detloss_finalize_kernel(const cnh_detloss_args a, const double* __restrict__ partials, int Btot) {
  __shared__ double sh_cols[CNH_PARTIALS];
  combine_partials(a, partials, Btot, a.scalars, sh_cols);
}

// g *= factor (factor read from device scalars); nothing to do when factor == 1.
__global__ void __launch_bounds__(kThreads)
scale_inplace_kernel(const cnh_scale_args s) {
  for (int t = 0; t < s.n_tensors; ++t) {
    const float f = (s.fa[t] ? __ldg(s.fa[t]) : 0.f) + (s.fb[t] ? __ldg(s.fb[t]) : 0.f);
    if (f == 1.0f) continue;
    float* __restrict__ p = s.data[t];
    const long long n = s.count[t];
    const long long stride = (long long)gridDim.x * kThreads;
    long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
      const long long n4 = n >> 2;
      float4* p4 = reinterpret_cast<float4*>(p);
      for (long long k = i; k < n4; k += stride) {
        float4 v = p4[k];
        v.x *= f; v.y *= f; v.z *= f; v.w *= f;
        p4[k] = v;
      }
      for (long long k = (n4 << 2) + i; k < n; k += stride) p[k] *= f;
    } else {
      for (long long k = i; k < n; k += stride) p[k] *= f;
    }
  }
}

// ---- host side -----------------------------------------------------------------------------
static int validate(const cnh_detloss_args* a, bool need_grad_ptrs) {
  CNH_REQUIRE(a != nullptr, CNH_E_NULL, "detloss: args is NULL");
  CNH_REQUIRE(a->B > 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->M >= 0, CNH_E_SHAPE,
              "detloss: bad dims B=%d C=%d H=%d W=%d M=%d", a->B, a->C, a->H, a->W, a->M);
  CNH_REQUIRE((long long)a->H * a->W < (1ll << 30), CNH_E_SHAPE, "detloss: H*W too large");
  CNH_REQUIRE((long long)a->B * a->C * a->H * a->W / kChunk + a->B < (1ll << 30), CNH_E_SHAPE,
              "detloss: heat map too large");
  CNH_REQUIRE(a->n_heads >= 0 && a->n_heads <= CNH_MAX_HEADS, CNH_E_SHAPE, "detloss: n_heads=%d",
              a->n_heads);
  CNH_REQUIRE(a->hm_logits && a->hm_gt && a->prob, CNH_E_NULL, "detloss: hm_logits/hm_gt/prob is NULL");
  CNH_REQUIRE(a->partials != nullptr, CNH_E_NULL, "detloss: partials is NULL");
  CNH_REQUIRE(a->n_heads == 0 || a->M == 0 || a->ind != nullptr, CNH_E_NULL, "detloss: ind is NULL");
  for (int h = 0; h < a->n_heads; ++h) {
    const cnh_head& hd = a->heads[h];
    CNH_REQUIRE(hd.map != nullptr, CNH_E_NULL, "detloss: head %d map is NULL", h);
    CNH_REQUIRE(a->M == 0 || (hd.target && hd.mask), CNH_E_NULL, "detloss: head %d target/mask is NULL", h);
    CNH_REQUIRE(hd.D > 0 && hd.D <= 1024, CNH_E_SHAPE, "detloss: head %d D=%d", h, hd.D);
    CNH_REQUIRE(hd.angle_mode >= CNH_ANGLE_NONE && hd.angle_mode <= CNH_ANGLE_PERIODIC, CNH_E_UNSUPPORTED,
                "detloss: head %d angle_mode=%d", h, hd.angle_mode);
    CNH_REQUIRE(!(hd.angle_mode == CNH_ANGLE_PERIODIC && hd.D != 3), CNH_E_SHAPE,
                "detloss: periodic angle loss needs a 3-channel head (got D=%d)", hd.D);
    if (need_grad_ptrs) {
      const bool any = a->grad_hm != nullptr;
      CNH_REQUIRE((hd.grad != nullptr) == any, CNH_E_NULL,
                  "detloss: grad pointers must be all set or all NULL (head %d)", h);
    }
  }
  return CNH_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static Geo make_geo(const cnh_detloss_args* a, void* ws) {
  Geo g;
  g.HW = a->H * a->W;
  g.CHW = (long long)a->C * g.HW;
  g.cps = (int)((g.CHW + kChunk - 1) / kChunk);
  g.n_chunks = a->B * g.cps;
  g.ppp = (g.HW + kPiece - 1) / kPiece;
  int it = 0;
  bool vec_planes = (g.HW % 4 == 0);
  for (int h = 0; h <= CNH_MAX_HEADS; ++h) {
    g.item0[h] = it;
    if (h < a->n_heads) {
      it += a->B * a->heads[h].D * g.ppp;
      vec_planes = vec_planes && aligned16(a->heads[h].grad);
    }
  }
  g.n_items = it;
  g.n_count = a->B * a->n_heads;
  g.vec_planes = vec_planes ? 1 : 0;
  char* p = static_cast<char*>(ws);
  g.hdr = reinterpret_cast<WsHeader*>(p);
  p += sizeof(WsHeader);
  g.chunk_sum = reinterpret_cast<float*>(p);
  p += align_up((size_t)g.n_chunks * 4, 16);
  g.chunk_npos = reinterpret_cast<int*>(p);
  p += align_up((size_t)g.n_chunks * 4, 16);
  g.item_l1 = reinterpret_cast<float*>(p);
  p += align_up((size_t)it * 4, 16);
  g.item_ang = reinterpret_cast<float*>(p);
  return g;
}

static size_t ws_bytes(const cnh_detloss_args* a) {
  Geo g = make_geo(a, nullptr);
  int items = g.item0[CNH_MAX_HEADS];
  return sizeof(WsHeader) + 2 * align_up((size_t)g.n_chunks * 4, 16) + 2 * align_up((size_t)items * 4, 16);
}

static bool use_vec(const cnh_detloss_args* a, const Geo& g) {
  return (g.CHW % 4 == 0) && aligned16(a->hm_logits) && aligned16(a->hm_gt) && aligned16(a->prob) &&
         (a->grad_hm == nullptr || aligned16(a->grad_hm));
}

template <int MODE>
static const void* pick_kernel(bool fast, bool vec) {
  if (fast) return vec ? (const void*)detloss_kernel<MODE, true, true> : (const void*)detloss_kernel<MODE, true, false>;
  return vec ? (const void*)detloss_kernel<MODE, false, true> : (const void*)detloss_kernel<MODE, false, false>;
}

static int max_resident_ctas(const void* kernel) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  return per_sm * sm_count();
}

static int launch(const void* kernel, bool cooperative, int grid, const cnh_detloss_args* a, const Geo& g,
                  cudaStream_t stream) {
  void* params[2] = {const_cast<cnh_detloss_args*>(a), const_cast<Geo*>(&g)};
  if (cooperative)
    CNH_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kThreads), params, 0, stream));
  else
    CNH_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(kThreads), params, 0, stream));
  return CNH_OK;
}

}  // namespace cnh

using namespace cnh;

extern "C" size_t cnh_detloss_workspace_bytes(const cnh_detloss_args* a) {
  if (validate(a, false) != CNH_OK) return 0;
  return ws_bytes(a);
}

extern "C" int cnh_detloss_fused(const cnh_detloss_args* a, void* workspace, size_t workspace_bytes,
                                 cnh_stream_t stream) {
  if (int rc = validate(a, true)) return rc;
  CNH_REQUIRE(a->scalars != nullptr, CNH_E_NULL, "detloss_fused: scalars is NULL");
  CNH_REQUIRE(workspace != nullptr && workspace_bytes >= ws_bytes(a), CNH_E_WORKSPACE,
              "detloss_fused: workspace %zu < %zu bytes", workspace_bytes, ws_bytes(a));
  const Geo g = make_geo(a, workspace);
  const bool fast = !(a->flags & CNH_FLAG_ACCURATE_MATH), vec = use_vec(a, g);
  const int units = g.n_chunks + g.n_items + g.n_count;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->grad_hm == nullptr) {
    const void* k = pick_kernel<M_FWD>(fast, vec);
    return launch(k, false, units < max_resident_ctas(k) ? units : max_resident_ctas(k), a, g, st);
  }
  const void* ks = pick_kernel<M_STASH>(fast, vec);
  const int cap = max_resident_ctas(ks);
  if (!(a->flags & CNH_FLAG_NO_STASH) && g.n_chunks <= (long long)cap * kStash) {
    int grid = units < cap ? units : cap;
    const int need = (g.n_chunks + kStash - 1) / kStash;
    if (grid < need) grid = need;
    return launch(ks, true, grid, a, g, st);
  }
  const void* kp = pick_kernel<M_PRECOUNT>(fast, vec);
  const int capp = max_resident_ctas(kp);
  return launch(kp, true, units < capp ? units : capp, a, g, st);
}

extern "C" int cnh_detloss_count(const cnh_detloss_args* a, void* workspace, size_t workspace_bytes,
                                 cnh_stream_t stream) {
  if (int rc = validate(a, false)) return rc;
  CNH_REQUIRE(a->norm_out != nullptr, CNH_E_NULL, "detloss_count: norm_out is NULL");
  CNH_REQUIRE(workspace != nullptr && workspace_bytes >= ws_bytes(a), CNH_E_WORKSPACE,
              "detloss_count: workspace %zu < %zu bytes", workspace_bytes, ws_bytes(a));
  const Geo g = make_geo(a, workspace);
  const void* k = pick_kernel<M_COUNT>(true, use_vec(a, g));
  const int units = g.n_chunks + g.n_count, cap = max_resident_ctas(k);
  return launch(k, false, units < cap ? units : cap, a, g, static_cast<cudaStream_t>(stream));
}

extern "C" int cnh_detloss_main(const cnh_detloss_args* a, void* workspace, size_t workspace_bytes,
                                cnh_stream_t stream) {
  if (int rc = validate(a, true)) return rc;
  CNH_REQUIRE(a->grad_hm != nullptr, CNH_E_NULL, "detloss_main: grad_hm is NULL (use cnh_detloss_fused for forward only)");
  CNH_REQUIRE(a->norm != nullptr, CNH_E_NULL, "detloss_main: norm is NULL");
  CNH_REQUIRE(workspace != nullptr && workspace_bytes >= ws_bytes(a), CNH_E_WORKSPACE,
              "detloss_main: workspace %zu < %zu bytes", workspace_bytes, ws_bytes(a));
  const Geo g = make_geo(a, workspace);
  const void* k = pick_kernel<M_MAIN>(!(a->flags & CNH_FLAG_ACCURATE_MATH), use_vec(a, g));
  const int units = g.n_chunks + g.n_items, cap = max_resident_ctas(k);
  return launch(k, false, units < cap ? units : cap, a, g, static_cast<cudaStream_t>(stream));
}

extern "C" int cnh_detloss_finalize(const cnh_detloss_args* a, const double* partials, int32_t B_total,
                                    cnh_stream_t stream) {
  CNH_REQUIRE(a != nullptr && partials != nullptr && a->scalars != nullptr, CNH_E_NULL,
              "detloss_finalize: args/partials/scalars is NULL");
  CNH_REQUIRE(B_total > 0 && a->n_heads >= 0 && a->n_heads <= CNH_MAX_HEADS, CNH_E_SHAPE,
              "detloss_finalize: B_total=%d n_heads=%d", B_total, a->n_heads);
  detloss_finalize_kernel<<<1, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(*a, partials, B_total);
  CNH_CUDA(cudaGetLastError());
  return CNH_OK;
}

extern "C" int cnh_scale_inplace(const cnh_scale_args* s, cnh_stream_t stream) {
  CNH_REQUIRE(s != nullptr, CNH_E_NULL, "scale_inplace: args is NULL");
  CNH_REQUIRE(s->n_tensors >= 0 && s->n_tensors <= 4, CNH_E_SHAPE, "scale_inplace: n_tensors=%d", s->n_tensors);
  long long most = 0;
  for (int t = 0; t < s->n_tensors; ++t) {
    CNH_REQUIRE(s->data[t] != nullptr && s->count[t] >= 0, CNH_E_NULL, "scale_inplace: tensor %d", t);
    CNH_REQUIRE((reinterpret_cast<uintptr_t>(s->data[t]) & 3u) == 0, CNH_E_ALIGN, "scale_inplace: tensor %d misaligned", t);
    if (s->count[t] > most) most = s->count[t];
  }
  if (s->n_tensors == 0 || most == 0) return CNH_OK;
  long long want = (most / 4 + kThreads - 1) / kThreads;
  const int cap = sm_count() * 8;
  const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  scale_inplace_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(*s);
  CNH_CUDA(cudaGetLastError());
  return CNH_OK;
}

// Peak-candidate lists shared by the streaming decode kernel and the detection-loss kernels (which can emit the
// candidates of their own probability tiles while those are still in shared memory: the decode then never reads the
// heat map again).  Layout of the candidate workspace, helpers of the two sides, and the tile scan.
#pragma once
#include "common.cuh"

namespace cnh {

constexpr int kFineBins = 4096;           // histogram: bin = min(4095, int(score * 4096))
constexpr int kSuperBins = 64;            // kFineBins / 64
constexpr int kSliceCap = 4096;           // keys one CTA may forward
constexpr int kCandRows = 32;             // rows of a candidate tile (one 4096-element chunk of a 128-wide plane)

// Per-sample state, zero between launches (the finish kernel leaves it so).
struct CandState {
  unsigned overflow;                      // 1: a buffer ran over, the cluster kernel must redo the sample; 2: it has
  unsigned pad[3];
};

// Candidate workspace: [B] CandState | [B][64] super bins | [B][4096] fine bins | [B][G] keys per slice |
// [B][G][kSliceCap] keys.  Everything but the keys must be zero between launches.
struct CandGeo {
  int B, G;                               // samples, CTAs (slices) per sample
  CandState* state;
  unsigned* shist;
  unsigned* fhist;
  unsigned* cta_cnt;
  u64* slices;
};
__host__ __device__ inline size_t cand_up128(size_t v) { return (v + 127) / 128 * 128; }
inline size_t cand_ws_bytes(int B, int G) {
  return cand_up128((size_t)B * sizeof(CandState)) + cand_up128((size_t)B * kSuperBins * 4) + cand_up128((size_t)B * kFineBins * 4) +
         cand_up128((size_t)B * G * 4) + (size_t)B * G * kSliceCap * sizeof(u64);
}
inline CandGeo cand_geo(void* ws, int B, int G) {
  CandGeo c;
  c.B = B;
  c.G = G;
  char* p = static_cast<char*>(ws);
  c.state = reinterpret_cast<CandState*>(p);
  p += cand_up128((size_t)B * sizeof(CandState));
  c.shist = reinterpret_cast<unsigned*>(p);
  p += cand_up128((size_t)B * kSuperBins * 4);
  c.fhist = reinterpret_cast<unsigned*>(p);
  p += cand_up128((size_t)B * kFineBins * 4);
  c.cta_cnt = reinterpret_cast<unsigned*>(p);
  p += cand_up128((size_t)B * G * 4);
  c.slices = reinterpret_cast<u64*>(p);
  return c;
}

__device__ __forceinline__ int fine_bin(unsigned score_bits) {
  const int b = (int)(__uint_as_float(score_bits) * (float)kFineBins);   // exact: power-of-two scale
  return b < kFineBins - 1 ? b : kFineBins - 1;
}
__device__ __forceinline__ void red_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One warp forwards candidate keys to its CTA's slice of a sample's list, counts them into the sample's two-level
// histogram (fire-and-forget REDs) and derives the pruning threshold from that histogram: the lower edge of the fine
// bin of the K-th counted key.  Only keys already forwarded are counted, so every threshold is valid (at least K real
// peaks reach it).  All members are called by the full warp.
struct CandEmitter {
  unsigned* shist;
  unsigned* fhist;
  u64* slice;
  unsigned local_cnt, thr, cap;            // keys in the slice, threshold (score bits), keys the slice may hold
  unsigned* shared_cnt;                    // not null: several warps share the slice, push4 draws its positions here
  bool overflow;
  int K, lane;
  int pending, sb_sel, issued_at;          // refresh pipeline: 0 idle, 1 super bins in flight, 2 fine bins in flight
  unsigned h0, h1, above_sb;

  __device__ __forceinline__ void init(const CandGeo& c, int b, int j, int K_) {
    shist = c.shist + (long long)b * kSuperBins;
    fhist = c.fhist + (long long)b * kFineBins;
    slice = c.slices + ((long long)b * c.G + j) * kSliceCap;
    local_cnt = 0;
    thr = 0;
    cap = (unsigned)kSliceCap;
    shared_cnt = nullptr;
    overflow = false;
    K = K_;
    lane = threadIdx.x & 31;
    pending = 0;
    sb_sel = 0;
    issued_at = 0;
    h0 = h1 = above_sb = 0;
  }
  // keys[0..n) (shared memory) -> slice; `counted(key)` says whether the key may enter the histogram (a key whose
  // 3x3 test is incomplete may not: it could turn out not to be a peak)
  template <class Counted>
  __device__ __forceinline__ void forward(const u64* keys, unsigned n, unsigned cap, Counted counted) {
    if (n > cap) { overflow = true; n = cap; }
    if (local_cnt + n > cap) { overflow = true; n = cap - local_cnt; }
    // Keys of one 32-lane step that fall into the same bin are counted with ONE RED (match.any): the top bins of a
    // sample are hit by every CTA that serves it, and same-address atomics serialise in the L2 slice.
    for (unsigned k0 = 0; k0 < n; k0 += 32) {
      const unsigned k = k0 + (unsigned)lane;
      int bin = -1;
      if (k < n) {
        const u64 key = keys[k];
        slice[local_cnt + k] = key;
        if (counted(key)) bin = fine_bin((unsigned)(key >> 32));
      }
      const unsigned same_f = __match_any_sync(0xffffffffu, bin);
      const unsigned same_s = __match_any_sync(0xffffffffu, bin >> 6);      // (-1 >> 6 == -1: the idle lanes pair up)
      if (bin >= 0) {
        if ((int)(__ffs(same_f) - 1) == lane) red_add_u32(fhist + bin, (unsigned)__popc(same_f));
        if ((int)(__ffs(same_s) - 1) == lane) red_add_u32(shist + (bin >> 6), (unsigned)__popc(same_s));
      }
    }
    local_cnt += n;
    __syncwarp();
  }
  // one key per lane (or none) straight from registers; `count`: may it enter the histogram (complete 3x3 test)?
  __device__ __forceinline__ void push(bool has, u64 key, bool count) {
    const unsigned bal = __ballot_sync(0xffffffffu, has);
    if (bal == 0u) return;
    const unsigned n = (unsigned)__popc(bal);
    if (local_cnt + n > cap) {
      overflow = true;
      return;
    }
    int bin = -1;
    if (has) {
      slice[local_cnt + (unsigned)__popc(bal & ((1u << lane) - 1u))] = key;
      if (count) bin = fine_bin((unsigned)(key >> 32));
    }
    const unsigned same_f = __match_any_sync(0xffffffffu, bin);
    const unsigned same_s = __match_any_sync(0xffffffffu, bin >> 6);
    if (bin >= 0) {
      if ((int)(__ffs(same_f) - 1) == lane) red_add_u32(fhist + bin, (unsigned)__popc(same_f));
      if ((int)(__ffs(same_s) - 1) == lane) red_add_u32(shist + (bin >> 6), (unsigned)__popc(same_s));
    }
    local_cnt += n;
  }
  // up to four keys per lane (bit e of `flags` = key[e] is valid): one prefix sum for the warp, plain REDs
  __device__ __forceinline__ void push4(unsigned flags, const u64 (&key)[4], bool count) {
    if (__ballot_sync(0xffffffffu, flags != 0u) == 0u) return;
    const int mine = __popc(flags);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      incl += (lane >= o) ? v : 0;
    }
    const unsigned n = (unsigned)__shfl_sync(0xffffffffu, incl, 31);
    unsigned first = local_cnt;
    if (shared_cnt != nullptr) {                             // (shared-memory counter: one atomic per warp and call)
      if (lane == 31) first = atomicAdd(shared_cnt, n);
      first = __shfl_sync(0xffffffffu, first, 31);
    }
    if (first + n > cap) {
      overflow = true;
      return;
    }
    unsigned pos = first + (unsigned)(incl - mine);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (flags & (1u << e)) {
        slice[pos++] = key[e];
        if (count) {
          const int bin = fine_bin((unsigned)(key[e] >> 32));
          red_add_u32(fhist + bin, 1u);
          red_add_u32(shist + (bin >> 6), 1u);
        }
      }
    local_cnt += n;
  }
  __device__ __forceinline__ void load_super() {
    const uint2 v = __ldcg(reinterpret_cast<const uint2*>(shist) + lane);
    h0 = v.x;
    h1 = v.y;
  }
  // h0/h1 = this lane's two super bins; true: the fine bins of the K-th key's super bin are being loaded
  __device__ __forceinline__ bool super_step() {
    const unsigned mine = h0 + h1;
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += v;
    }
    const unsigned above = incl - mine;                      // keys in the super bins above this lane's pair
    int sel = -1;
    unsigned ab = 0;
    if (above < (unsigned)K && incl >= (unsigned)K) {
      if (above + h1 >= (unsigned)K) { sel = 2 * lane + 1; ab = above; }
      else { sel = 2 * lane; ab = above + h1; }
    }
    const unsigned who = __ballot_sync(0xffffffffu, sel >= 0);
    if (who == 0u) return false;                             // fewer than K keys counted so far: no threshold yet
    const int src = __ffs(who) - 1;
    sb_sel = __shfl_sync(0xffffffffu, sel, src);
    above_sb = __shfl_sync(0xffffffffu, ab, src);
    const uint2 f = __ldcg(reinterpret_cast<const uint2*>(fhist + sb_sel * 64) + lane);
    h0 = f.x;
    h1 = f.y;
    return true;
  }
  // h0/h1 = this lane's two fine bins of super bin sb_sel
  __device__ __forceinline__ void fine_step() {
    const unsigned mine = h0 + h1;
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += v;
    }
    const unsigned above = above_sb + incl - mine;
    int sel = -1;
    if (above < (unsigned)K && above + mine >= (unsigned)K) sel = (above + h1 >= (unsigned)K) ? 2 * lane + 1 : 2 * lane;
    const unsigned who = __ballot_sync(0xffffffffu, sel >= 0);
    if (who != 0u) {
      const int fb = sb_sel * 64 + __shfl_sync(0xffffffffu, sel, __ffs(who) - 1);
      const unsigned t_new = __float_as_uint((float)fb * (1.0f / (float)kFineBins));
      if (t_new > thr) thr = t_new;
    }
  }
  // the threshold NOW (two L2 round trips are waited for)
  __device__ __forceinline__ void refresh_blocking() {
    load_super();
    if (super_step()) fine_step();
    pending = 0;
  }
  // one step of the pipelined refresh at iteration i: no load is consumed before it is kAge iterations old
  template <int kAge = 3>
  __device__ __forceinline__ void refresh_step(int i, bool more) {
    if (pending == 1 && i - issued_at >= kAge) {
      pending = super_step() ? 2 : 0;
      issued_at = i;
    } else if (pending == 2 && i - issued_at >= kAge) {
      fine_step();
      pending = 0;
    }
    if (pending == 0 && more) {
      load_super();
      pending = 1;
      issued_at = i;
    }
  }
  // On the way out: the slice once more against the latest threshold (most of what it holds was scanned before
  // there was one), in place, eight 32-key chunks per batch: a batch is in registers before anything at or below it
  // is written.  `touch(key)` is called for every survivor (prefetch of what the finish kernel will gather).
  template <class Touch>
  __device__ __forceinline__ void reprune(Touch touch) {
    if (overflow || local_cnt == 0u) return;
    if (pending == 2) { fine_step(); pending = 0; }
    unsigned kept = 0;
    constexpr int kBatch = 8;
    for (unsigned k0 = 0; k0 < local_cnt; k0 += 32 * kBatch) {
      u64 key[kBatch];
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const unsigned k = k0 + q * 32 + lane;
        key[q] = k < local_cnt ? __ldcg(slice + k) : 0ull;
      }
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const bool keep = (unsigned)(key[q] >> 32) >= thr && key[q] != 0ull;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          slice[kept + __popc(bal & ((1u << lane) - 1u))] = key[q];
          touch(key[q]);
        }
        kept += __popc(bal);
      }
    }
    local_cnt = kept;
  }
};


// ---- the emitter warp of the detection-loss kernels: peak tests on a 32 x 128 probability tile that lies in GLOBAL
// memory (the chunk the CTA's other warps have just written; read through L2, so no shared-memory stage is held).
// A tile is HALO-LESS: the rows above row 0 and below row 31 belong to other chunks and are taken as 0, so a pixel of
// those two rows is tested against its five neighbours inside the tile only; such keys are forwarded but not counted
// into the threshold histogram, and the finish kernel checks the other three neighbours (`verify_rows`).
// Key = (score bits << 32) | ~(flat index): descending key order = score descending, ties to the lower index.
struct GTile {
  const float* p;                          // tile origin (row 0, column 0), 128 floats per row
  unsigned flat0;                          // flat index (inside the sample) of its first pixel
  int y0, H, HW;                           // first image row, image height (which keys are complete tests)
  __device__ __forceinline__ float at(int r, int x) const {
    return (r >= 0 && r < kCandRows && x >= 0 && x < 128) ? __ldcg(p + r * 128 + x) : 0.f;
  }
  __device__ __forceinline__ bool complete(int r) const {          // was row r tested against real rows on both sides?
    const int y = y0 + r;
    return !((r == 0 && y > 0) || (r == kCandRows - 1 && y < H - 1));
  }
  __device__ __forceinline__ u64 key(float v, int r, int x) const {
    return ((u64)__float_as_uint(v) << 32) | (u64)(0xffffffffu - (flat0 + (unsigned)(r * 128 + x)));
  }
};

// (A) rows [r0, r0 + kR) with NO threshold, row-wise in registers (kR + 2 16-byte loads per lane, one round trip)
template <int kR = 4>
__device__ __forceinline__ void gtile_rows_unpruned(CandEmitter& em, const GTile& t, int r0) {
  const int lane = threadIdx.x & 31;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 row[kR + 2];
#pragma unroll
  for (int q = 0; q < kR + 2; ++q) {
    const int r = r0 - 1 + q;
    row[q] = (r >= 0 && r < kCandRows) ? __ldcg(reinterpret_cast<const float4*>(t.p + r * 128) + lane) : zero;
  }
#pragma unroll
  for (int rr = 0; rr < kR; ++rr) {
    const float4 up = row[rr], mid = row[rr + 1], dn = row[rr + 2];
    const float v0 = fmaxf(fmaxf(up.x, mid.x), dn.x), v1 = fmaxf(fmaxf(up.y, mid.y), dn.y);
    const float v2 = fmaxf(fmaxf(up.z, mid.z), dn.z), v3 = fmaxf(fmaxf(up.w, mid.w), dn.w);
    float left = __shfl_up_sync(0xffffffffu, v3, 1), right = __shfl_down_sync(0xffffffffu, v0, 1);
    if (lane == 0) left = 0.f;
    if (lane == 31) right = 0.f;
    const float h[4] = {fmaxf(fmaxf(left, v0), v1), fmaxf(fmaxf(v0, v1), v2), fmaxf(fmaxf(v1, v2), v3), fmaxf(fmaxf(v2, v3), right)};
    const float m[4] = {mid.x, mid.y, mid.z, mid.w};
    unsigned flags = 0;
    u64 key[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      flags |= (m[e] > 0.f && m[e] == h[e]) ? (1u << e) : 0u;
      key[e] = t.key(m[e], r0 + rr, 4 * lane + e);
    }
    em.push4(flags, key, t.complete(r0 + rr));
  }
}

// one pixel per lane (or none): the 3x3 test with scalar loads, then push
__device__ __forceinline__ void gtile_test_push(CandEmitter& em, const GTile& t, bool has, int off, float thr_eff) {
  bool peak = false;
  float v = 0.f;
  const int r = off >> 7, x = off & 127;
  if (has) {
    float nb[8];                                                       // the pixel and its neighbours: nine independent
    int q = 0;                                                         // loads, ONE round trip through L2
    v = __ldcg(t.p + off);
#pragma unroll
    for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx)
        if (dr != 0 || dx != 0) nb[q++] = t.at(r + dr, x + dx);
    float mx = v;
#pragma unroll
    for (int k = 0; k < 8; ++k) mx = fmaxf(mx, nb[k]);
    peak = (v >= thr_eff && v == mx);
  }
  em.push(peak, t.key(v, r, x), t.complete(r));
}

// (B) the pixels the loss warps noted (offsets inside the tile): lane k takes entry k, k + 32, ...
__device__ __forceinline__ void gtile_pending(CandEmitter& em, const GTile& t, const unsigned short* pend, unsigned n_pend) {
  const int lane = threadIdx.x & 31;
  const float thr_eff = fmaxf(__uint_as_float(em.thr), __uint_as_float(1u));
  for (unsigned k0 = 0; k0 < n_pend; k0 += 32) {
    const unsigned k = k0 + (unsigned)lane;
    gtile_test_push(em, t, k < n_pend, k < n_pend ? (int)pend[k] : 0, thr_eff);
  }
}

// (C) rows [r_begin, 32) against em.thr without a list: pass 1 marks this lane's pixels >= thr (one 16-byte load per
// row, all in flight), pass 2 tests them, every lane one pixel per trip
__device__ __forceinline__ void gtile_scan(CandEmitter& em, const GTile& t, int r_begin) {
  const int lane = threadIdx.x & 31;
  const float thr_eff = fmaxf(__uint_as_float(em.thr), __uint_as_float(1u));
  unsigned m0 = 0, m1 = 0, m2 = 0, m3 = 0;
#pragma unroll
  for (int r = 0; r < kCandRows; ++r) {
    if (r >= r_begin) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(t.p + r * 128) + lane);
      m0 |= (v.x >= thr_eff ? 1u : 0u) << r;
      m1 |= (v.y >= thr_eff ? 1u : 0u) << r;
      m2 |= (v.z >= thr_eff ? 1u : 0u) << r;
      m3 |= (v.w >= thr_eff ? 1u : 0u) << r;
    }
  }
  while (__any_sync(0xffffffffu, (m0 | m1 | m2 | m3) != 0u)) {
    int e = -1, r = 0;
    if (m0) { e = 0; r = __ffs(m0) - 1; m0 &= m0 - 1u; }
    else if (m1) { e = 1; r = __ffs(m1) - 1; m1 &= m1 - 1u; }
    else if (m2) { e = 2; r = __ffs(m2) - 1; m2 &= m2 - 1u; }
    else if (m3) { e = 3; r = __ffs(m3) - 1; m3 &= m3 - 1u; }
    gtile_test_push(em, t, e >= 0, r * 128 + 4 * lane + (e >= 0 ? e : 0), thr_eff);
  }
}

}  // namespace cnh

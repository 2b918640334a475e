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

// Scan of a HALO-LESS candidate tile: 32 rows x 128 columns of clamped probabilities in shared memory (what a
// detection-loss chunk leaves of one class plane), four rows per warp (warp w of 8: rows 4w..4w+3), lane owns
// columns [4*lane, 4*lane+4).  The rows above row 0 and below row 31 belong to other chunks and are taken as 0: a
// pixel of those two rows is tested against its five neighbours inside the tile only (the finish kernel checks the
// other three, see `verify_rows`).  Peaks >= thr are appended to keys[*key_cnt ...] (bounded by cap; the counter may
// run past it).  Key = (score bits << 32) | ~(flat index): descending key order = score descending, ties to the lower
// index.
__device__ __forceinline__ void scan_chunk_rows(u64* keys, unsigned* key_cnt, unsigned cap, const float* tile, unsigned thr,
                                                unsigned flat_tile0, int warp) {
  constexpr int W = 128, RW = 4;
  const int lane = threadIdx.x & 31;
  const int r0 = RW * warp;
  const float* base_row = tile + r0 * W + 4 * lane;                               // the warp's first row
  const float thr_eff = fmaxf(__uint_as_float(thr), __uint_as_float(1u));      // >= thr and > 0
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 up = r0 > 0 ? *reinterpret_cast<const float4*>(base_row - W) : zero;
  float4 mid = *reinterpret_cast<const float4*>(base_row);
  unsigned flags = 0;                                                             // bit 4*rr + e
#pragma unroll
  for (int rr = 0; rr < RW; ++rr) {
    const float4 dn = (r0 + rr + 1 < kCandRows) ? *reinterpret_cast<const float4*>(base_row + (rr + 1) * W) : zero;
    const bool any = fmaxf(fmaxf(mid.x, mid.y), fmaxf(mid.z, mid.w)) >= thr_eff;
    if (__ballot_sync(0xffffffffu, any) != 0u) {
      const unsigned pass = (mid.x >= thr_eff ? 1u : 0u) | (mid.y >= thr_eff ? 2u : 0u) | (mid.z >= thr_eff ? 4u : 0u) |
                            (mid.w >= thr_eff ? 8u : 0u);
      const float v0 = fmaxf(fmaxf(up.x, mid.x), dn.x), v1 = fmaxf(fmaxf(up.y, mid.y), dn.y);
      const float v2 = fmaxf(fmaxf(up.z, mid.z), dn.z), v3 = fmaxf(fmaxf(up.w, mid.w), dn.w);
      float left = __shfl_up_sync(0xffffffffu, v3, 1), right = __shfl_down_sync(0xffffffffu, v0, 1);
      if (lane == 0) left = 0.f;
      if (lane == 31) right = 0.f;
      const float h0 = fmaxf(fmaxf(left, v0), v1), h1 = fmaxf(fmaxf(v0, v1), v2);
      const float h2 = fmaxf(fmaxf(v1, v2), v3), h3 = fmaxf(fmaxf(v2, v3), right);
      const unsigned f = pass & ((mid.x == h0 ? 1u : 0u) | (mid.y == h1 ? 2u : 0u) | (mid.z == h2 ? 4u : 0u) |
                                 (mid.w == h3 ? 8u : 0u));
      flags |= f << (4 * rr);
    }
    up = mid;
    mid = dn;
  }
  if (__ballot_sync(0xffffffffu, flags != 0u) == 0u) return;
  const int mine = __popc(flags);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    incl += (lane >= o) ? v : 0;
  }
  unsigned base = 0;
  if (lane == 31) base = atomicAdd(key_cnt, (unsigned)incl);
  base = __shfl_sync(0xffffffffu, base, 31);
  unsigned pos = base + (unsigned)(incl - mine);
  const unsigned flat_lane0 = flat_tile0 + (unsigned)r0 * (unsigned)W + 4u * (unsigned)lane;
  while (flags) {
    const int bit = __ffs(flags) - 1;
    flags &= flags - 1u;
    const int rr = bit >> 2, e = bit & 3;
    const float v = base_row[rr * W + e];
    if (pos < cap) keys[pos] = ((u64)__float_as_uint(v) << 32) | (u64)(0xffffffffu - (flat_lane0 + (unsigned)(rr * W + e)));
    ++pos;
  }
}
// may a key of such a tile enter the threshold histogram?  Only if its 3x3 test was complete.
__device__ __forceinline__ bool chunk_key_verified(u64 key, int HW, int H) {
  const unsigned flat = 0xffffffffu - (unsigned)(key & 0xffffffffu);
  const int y = (int)((flat % (unsigned)HW) >> 7);           // W == 128
  const int r = y & (kCandRows - 1);
  return !((r == 0 && y > 0) || (r == kCandRows - 1 && y < H - 1));
}

// One warp forwards candidate keys to its CTA's slice of a sample's list, counts them into the sample's two-level
// histogram (fire-and-forget REDs) and derives the pruning threshold from that histogram: the lower edge of the fine
// bin of the K-th counted key.  Only keys already forwarded are counted, so every threshold is valid (at least K real
// peaks reach it).  All members are called by the full warp.
struct CandEmitter {
  unsigned* shist;
  unsigned* fhist;
  u64* slice;
  unsigned local_cnt, thr;
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
    if (local_cnt + n > (unsigned)kSliceCap) { overflow = true; n = (unsigned)kSliceCap - local_cnt; }
    for (unsigned k = lane; k < n; k += 32) {
      const u64 key = keys[k];
      slice[local_cnt + k] = key;
      if (counted(key)) {
        const int bin = fine_bin((unsigned)(key >> 32));
        red_add_u32(fhist + bin, 1u);
        red_add_u32(shist + (bin >> 6), 1u);
      }
    }
    local_cnt += n;
    __syncwarp();
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
  __device__ __forceinline__ void refresh_step(int i, bool more) {
    constexpr int kAge = 3;
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
    if (pending == 2) fine_step();
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

}  // namespace cnh

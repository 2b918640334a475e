// UDA target-domain losses over a per-pixel channel softmax, forward + backward fused:
//   entropy            losses/entropy.py:10-28     (eta == None and the FDA eta variant)
//   max-squares        losses/max_square.py:6-14
//   self-information   utils/image.py:121-124      (entropy_map, fwd and bwd)
//   BCE vs constant    losses/advent.py:10-18
// HBM-bound: logits are read once (float4 along W, one thread owns 4 adjacent pixels and
// walks the C planes), the gradient is written once.  For C <= 8 the C*4 logits stay in
// registers; for larger C an online softmax pass (max / sum / sum e*(x-m) / sum e^2) is
// followed by a second pass that re-reads the logits through L2.
//
// Closed forms (v = softmax_c(x), S = sum_c exp(x_c - m), l_c = log2 v_c):
//   H  = -sum v l = log2 S - log2(e) * sum_c e_c (x_c - m) / S          (bits)
//   dH/dx_j = -v_j (l_j + H)
//   d(sum v^2)/dx_j = 2 v_j (v_j - sum v^2)
// The reference's log2(v + 1e-30) differs from log2 v only where v < ~1e-23, i.e. by
// < 1e-21 in any term; the self-information MAP keeps the epsilon literally.
#include "common.cuh"

namespace cnh {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

template <int PV>
struct Pix {
  float v[PV];
};
template <int PV>
__device__ __forceinline__ Pix<PV> load_pix(const float* p) {
  Pix<PV> r;
  if (PV == 4) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2 % PV] = t.z; r.v[3 % PV] = t.w;
  } else {
    r.v[0] = __ldcs(p);
  }
  return r;
}
template <int PV>
__device__ __forceinline__ void store_pix(float* p, const Pix<PV>& r) {
  if (PV == 4) __stcs(reinterpret_cast<float4*>(p), make_float4(r.v[0], r.v[1 % PV], r.v[2 % PV], r.v[3 % PV]));
  else __stcs(p, r.v[0]);
}

struct SoftStat {           // per pixel
  float m, S, T, Q;         // max, sum e, sum e*(x-m), sum e^2
};
__device__ __forceinline__ void stat_init(SoftStat& s, float x) { s.m = x; s.S = 1.f; s.T = 0.f; s.Q = 1.f; }
__device__ __forceinline__ void stat_push(SoftStat& s, float x) {     // online update
  const float mn = fmaxf(s.m, x);
  const float r = __expf(s.m - mn), e = __expf(x - mn);
  s.T = r * (s.T + s.S * (s.m - mn)) + e * (x - mn);
  s.S = s.S * r + e;
  s.Q = s.Q * r * r + e * e;
  s.m = mn;
}

struct LossParams {
  int N, C, HW;
  float k_ent;        // 1 / (Ntot*HW*log2 C)
  float k_mean;       // 1 / (Ntot*HW)
  float k_msq;        // 1 / (Ntot*C*HW)
  float inv_log2c;
  float eta;
};

// MODE: CNH_SOFTMAX_*.  CS: static channel count (0 = dynamic).  PV: pixels per thread.
template <int MODE, int CS, int PV, bool GRAD>
__global__ void __launch_bounds__(kThreads)
softmax_loss_kernel(const float* __restrict__ x, float* __restrict__ grad, float* __restrict__ cta_part,
                    unsigned* __restrict__ ticket, float* __restrict__ loss_out, const LossParams P) {
  __shared__ float red_f[kWarps];
  __shared__ double red_d[kWarps];
  __shared__ unsigned sh_ticket;
  const int C = CS ? CS : P.C;
  const long long per_n = P.HW / PV;                   // pixel vectors per sample
  const long long total = (long long)P.N * per_n;
  float acc = 0.f;
  for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < total;
       q += (long long)gridDim.x * kThreads) {
    const long long n = q / per_n, pq = q - n * per_n;
    const float* xp = x + (n * C) * P.HW + pq * PV;
    float* gp = GRAD ? grad + (n * C) * P.HW + pq * PV : nullptr;
    Pix<PV> xs[CS ? CS : 1];
    SoftStat st[PV];
    if (CS) {
#pragma unroll
      for (int c = 0; c < CS; ++c) xs[c] = load_pix<PV>(xp + (long long)c * P.HW);
#pragma unroll
      for (int i = 0; i < PV; ++i) {
        float m = xs[0].v[i];
#pragma unroll
        for (int c = 1; c < CS; ++c) m = fmaxf(m, xs[c].v[i]);
        float S = 0.f, T = 0.f, Q = 0.f;
#pragma unroll
        for (int c = 0; c < CS; ++c) {
          const float d = xs[c].v[i] - m, e = __expf(d);
          S += e; T += e * d; Q += e * e;
        }
        st[i].m = m; st[i].S = S; st[i].T = T; st[i].Q = Q;
      }
    } else {
      Pix<PV> t = load_pix<PV>(xp);
#pragma unroll
      for (int i = 0; i < PV; ++i) stat_init(st[i], t.v[i]);
      for (int c = 1; c < C; ++c) {
        t = load_pix<PV>(xp + (long long)c * P.HW);
#pragma unroll
        for (int i = 0; i < PV; ++i) stat_push(st[i], t.v[i]);
      }
    }
    float lgS[PV], invS[PV], stat[PV], coef[PV];
#pragma unroll
    for (int i = 0; i < PV; ++i) {
      invS[i] = __fdividef(1.f, st[i].S);
      lgS[i] = __log2f(st[i].S);
      if (MODE == CNH_SOFTMAX_MAX_SQUARE) {
        stat[i] = st[i].Q * invS[i] * invS[i];               // sum v^2
        acc += stat[i];
        coef[i] = -P.k_msq;
      } else {
        stat[i] = lgS[i] - kLog2e * st[i].T * invS[i];       // H in bits
        if (MODE == CNH_SOFTMAX_ENTROPY) {
          acc += stat[i];
          coef[i] = -P.k_ent;
        } else {
          const float ent = stat[i] * P.inv_log2c;
          const float base = ent * ent + 1e-30f;
          acc += powf(base, P.eta);
          coef[i] = -P.eta * powf(base, P.eta - 1.f) * 2.f * ent * P.inv_log2c * P.k_mean;
        }
      }
    }
    if (GRAD) {
      for (int c = 0; c < C; ++c) {
        Pix<PV> t, o;
        if (CS) {
          t = xs[CS ? c : 0];
        } else {
          t = load_pix<PV>(xp + (long long)c * P.HW);
        }
#pragma unroll
        for (int i = 0; i < PV; ++i) {
          const float d = t.v[i] - st[i].m;
          const float v = __expf(d) * invS[i];
          if (MODE == CNH_SOFTMAX_MAX_SQUARE) o.v[i] = coef[i] * v * (v - stat[i]);
          else o.v[i] = coef[i] * v * (d * kLog2e - lgS[i] + stat[i]);
        }
        store_pix<PV>(gp + (long long)c * P.HW, o);
      }
    }
  }
  // deterministic: per-CTA partial (fixed tree), last CTA sums the partials in order
  acc = block_sum(acc, red_f);
  if (threadIdx.x == 0) {
    cta_part[blockIdx.x] = acc;
    __threadfence();
    sh_ticket = atomicAdd(ticket, 1u);
  }
  __syncthreads();
  if (sh_ticket != gridDim.x - 1) return;
  __threadfence();
  double s = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) s += (double)__ldcg(cta_part + i);
  s = block_sum(s, red_d);
  if (threadIdx.x == 0) {
    float l;
    if (MODE == CNH_SOFTMAX_ENTROPY) l = (float)(s * (double)P.k_ent);
    else if (MODE == CNH_SOFTMAX_ENTROPY_ETA) l = (float)(s * (double)P.k_mean);
    else l = (float)(-0.5 * s * (double)P.k_msq);
    loss_out[0] = l;
    *ticket = 0u;
  }
}

// ---- self-information map ---------------------------------------------------------------
template <int CS, int PV, bool BWD>
__global__ void __launch_bounds__(kThreads)
entropy_map_kernel(const float* __restrict__ x, const float* __restrict__ gout, float* __restrict__ out,
                   int N, int Cdyn, int HW, float inv_log2c) {
  const int C = CS ? CS : Cdyn;
  const long long per_n = HW / PV;
  const long long total = (long long)N * per_n;
  for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < total;
       q += (long long)gridDim.x * kThreads) {
    const long long n = q / per_n, pq = q - n * per_n;
    const long long off = (n * C) * HW + pq * PV;
    const float* xp = x + off;
    Pix<PV> xs[CS ? CS : 1];
    float m[PV], S[PV];
    if (CS) {
#pragma unroll
      for (int c = 0; c < CS; ++c) xs[c] = load_pix<PV>(xp + (long long)c * HW);
#pragma unroll
      for (int i = 0; i < PV; ++i) {
        m[i] = xs[0].v[i];
#pragma unroll
        for (int c = 1; c < CS; ++c) m[i] = fmaxf(m[i], xs[c].v[i]);
        S[i] = 0.f;
#pragma unroll
        for (int c = 0; c < CS; ++c) S[i] += __expf(xs[c].v[i] - m[i]);
      }
    } else {
      Pix<PV> t = load_pix<PV>(xp);
#pragma unroll
      for (int i = 0; i < PV; ++i) { m[i] = t.v[i]; S[i] = 1.f; }
      for (int c = 1; c < C; ++c) {
        t = load_pix<PV>(xp + (long long)c * HW);
#pragma unroll
        for (int i = 0; i < PV; ++i) {
          const float mn = fmaxf(m[i], t.v[i]);
          S[i] = S[i] * __expf(m[i] - mn) + __expf(t.v[i] - mn);
          m[i] = mn;
        }
      }
    }
    float invS[PV];
#pragma unroll
    for (int i = 0; i < PV; ++i) invS[i] = __fdividef(1.f, S[i]);
    if (!BWD) {
      for (int c = 0; c < C; ++c) {
        Pix<PV> t, o;
        if (CS) t = xs[CS ? c : 0]; else t = load_pix<PV>(xp + (long long)c * HW);
#pragma unroll
        for (int i = 0; i < PV; ++i) {
          const float v = __expf(t.v[i] - m[i]) * invS[i];
          o.v[i] = -(v * __log2f(v + 1e-30f)) * inv_log2c;        // utils/image.py:124
        }
        store_pix<PV>(out + off + (long long)c * HW, o);
      }
    } else {
      // d/dx_j = v_j (g_j a_j - sum_c g_c a_c v_c),  a_c = -(log2(v_c+eps) + v_c/((v_c+eps) ln2)) / log2 C
      float dot[PV];
#pragma unroll
      for (int i = 0; i < PV; ++i) dot[i] = 0.f;
      for (int c = 0; c < C; ++c) {
        Pix<PV> t;
        if (CS) t = xs[CS ? c : 0]; else t = load_pix<PV>(xp + (long long)c * HW);
        const Pix<PV> gq = load_pix<PV>(gout + off + (long long)c * HW);
#pragma unroll
        for (int i = 0; i < PV; ++i) {
          const float v = __expf(t.v[i] - m[i]) * invS[i];
          const float ve = v + 1e-30f;
          const float a = -(__log2f(ve) + __fdividef(v, ve * kLn2)) * inv_log2c;
          dot[i] += gq.v[i] * a * v;
        }
      }
      for (int c = 0; c < C; ++c) {
        Pix<PV> t, o;
        if (CS) t = xs[CS ? c : 0]; else t = load_pix<PV>(xp + (long long)c * HW);
        const Pix<PV> gq = load_pix<PV>(gout + off + (long long)c * HW);   // second read: L1/L2 hit
#pragma unroll
        for (int i = 0; i < PV; ++i) {
          const float v = __expf(t.v[i] - m[i]) * invS[i];
          const float ve = v + 1e-30f;
          const float a = -(__log2f(ve) + __fdividef(v, ve * kLn2)) * inv_log2c;
          o.v[i] = v * (gq.v[i] * a - dot[i]);
        }
        store_pix<PV>(out + off + (long long)c * HW, o);
      }
    }
  }
}

// ---- BCE with logits vs a constant label (tiny: one CTA, deterministic) ---------------------
__global__ void __launch_bounds__(kThreads)
bce_const_kernel(const float* __restrict__ y, float* __restrict__ grad, float* __restrict__ loss_out,
                 long long n, float label) {
  __shared__ double red_d[kWarps];
  double acc = 0.0;
  const float inv_n = 1.0f / (float)n;
  for (long long i = threadIdx.x; i < n; i += kThreads) {
    const float v = y[i];
    // max(v,0) - v*t + log(1 + exp(-|v|))   (ATen binary_cross_entropy_with_logits)
    acc += (double)(fmaxf(v, 0.f) - v * label + log1pf(expf(-fabsf(v))));
    if (grad) grad[i] = (1.0f / (1.0f + expf(-v)) - label) * inv_n;
  }
  acc = block_sum(acc, red_d);
  if (threadIdx.x == 0) loss_out[0] = (float)(acc / (double)n);
}

// ---- host -----------------------------------------------------------------------------------
static int stream_grid(long long work_items) {
  long long want = (work_items + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count() * 8;
  if (want < 1) want = 1;
  return (int)(want > cap ? cap : want);
}

template <int MODE, int PV, bool GRAD>
static void launch_loss(int C, int grid, cudaStream_t st, const float* x, float* g, float* part, unsigned* tk,
                        float* out, const LossParams& P) {
#define CNH_CASE(cs) \
  case cs: softmax_loss_kernel<MODE, cs, PV, GRAD><<<grid, kThreads, 0, st>>>(x, g, part, tk, out, P); break;
  switch (C) {
    CNH_CASE(1) CNH_CASE(2) CNH_CASE(3) CNH_CASE(4) CNH_CASE(5) CNH_CASE(6) CNH_CASE(7) CNH_CASE(8)
    default: softmax_loss_kernel<MODE, 0, PV, GRAD><<<grid, kThreads, 0, st>>>(x, g, part, tk, out, P);
  }
#undef CNH_CASE
}

template <int PV, bool BWD>
static void launch_map(int C, int grid, cudaStream_t st, const float* x, const float* go, float* out, int N,
                       int HW, float inv_log2c) {
#define CNH_CASE(cs) \
  case cs: entropy_map_kernel<cs, PV, BWD><<<grid, kThreads, 0, st>>>(x, go, out, N, C, HW, inv_log2c); break;
  switch (C) {
    CNH_CASE(1) CNH_CASE(2) CNH_CASE(3) CNH_CASE(4) CNH_CASE(5) CNH_CASE(6) CNH_CASE(7) CNH_CASE(8)
    default: entropy_map_kernel<0, PV, BWD><<<grid, kThreads, 0, st>>>(x, go, out, N, C, HW, inv_log2c);
  }
#undef CNH_CASE
}

constexpr int kMaxLossGrid = 148 * 16;

}  // namespace cnh

using namespace cnh;

extern "C" size_t cnh_softmax_workspace_bytes(int32_t, int32_t, int32_t, int32_t) {
  return 64 + sizeof(float) * kMaxLossGrid;
}

extern "C" int cnh_softmax_loss(const float* logits, float* grad, float* loss_out, int32_t N, int32_t C,
                                int32_t H, int32_t W, int32_t n_total, int32_t mode, float eta,
                                void* workspace, size_t workspace_bytes, cnh_stream_t stream) {
  CNH_REQUIRE(logits && loss_out, CNH_E_NULL, "softmax_loss: logits/loss_out is NULL");
  CNH_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && n_total >= N, CNH_E_SHAPE,
              "softmax_loss: bad dims N=%d C=%d H=%d W=%d n_total=%d", N, C, H, W, n_total);
  CNH_REQUIRE(mode >= CNH_SOFTMAX_ENTROPY && mode <= CNH_SOFTMAX_MAX_SQUARE, CNH_E_UNSUPPORTED,
              "softmax_loss: mode=%d", mode);
  CNH_REQUIRE(workspace && workspace_bytes >= cnh_softmax_workspace_bytes(N, C, H, W), CNH_E_WORKSPACE,
              "softmax_loss: workspace too small");
  const long long HW = (long long)H * W;
  CNH_REQUIRE(HW < (1ll << 30), CNH_E_SHAPE, "softmax_loss: H*W too large");
  LossParams P;
  P.N = N; P.C = C; P.HW = (int)HW; P.eta = eta;
  const double log2c = (double)log2f((float)C);              // fp32 log2, losses/entropy.py:25
  P.inv_log2c = (float)(1.0 / log2c);
  P.k_ent = (float)(1.0 / ((double)n_total * (double)HW * log2c));
  P.k_mean = (float)(1.0 / ((double)n_total * (double)HW));
  P.k_msq = (float)(1.0 / ((double)n_total * (double)C * (double)HW));
  const bool vec = (HW % 4 == 0) && aligned16(logits) && (grad == nullptr || aligned16(grad));
  const long long items = (long long)N * (vec ? HW / 4 : HW);
  int grid = stream_grid(items);
  if (grid > kMaxLossGrid) grid = kMaxLossGrid;
  unsigned* tk = static_cast<unsigned*>(workspace);
  float* part = reinterpret_cast<float*>(static_cast<char*>(workspace) + 64);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CNH_GO(M)                                                                             \
  do {                                                                                        \
    if (vec) { if (grad) launch_loss<M, 4, true>(C, grid, st, logits, grad, part, tk, loss_out, P);  \
               else launch_loss<M, 4, false>(C, grid, st, logits, grad, part, tk, loss_out, P); }    \
    else     { if (grad) launch_loss<M, 1, true>(C, grid, st, logits, grad, part, tk, loss_out, P);  \
               else launch_loss<M, 1, false>(C, grid, st, logits, grad, part, tk, loss_out, P); }    \
  } while (0)
  if (mode == CNH_SOFTMAX_ENTROPY) CNH_GO(CNH_SOFTMAX_ENTROPY);
  else if (mode == CNH_SOFTMAX_ENTROPY_ETA) CNH_GO(CNH_SOFTMAX_ENTROPY_ETA);
  else CNH_GO(CNH_SOFTMAX_MAX_SQUARE);
#undef CNH_GO
  CNH_CUDA(cudaGetLastError());
  return CNH_OK;
}

static int map_common(const float* logits, const float* gout, float* out, int32_t N, int32_t C, int32_t H,
                      int32_t W, bool bwd, cnh_stream_t stream) {
  CNH_REQUIRE(logits && out && (!bwd || gout), CNH_E_NULL, "entropy_map: NULL pointer");
  CNH_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, CNH_E_SHAPE, "entropy_map: bad dims N=%d C=%d H=%d W=%d", N, C, H, W);
  const long long HW = (long long)H * W;
  CNH_REQUIRE(HW < (1ll << 30), CNH_E_SHAPE, "entropy_map: H*W too large");
  const float inv_log2c = (float)(1.0 / log2((double)C));    // np.log2(c), utils/image.py:124
  const bool vec = (HW % 4 == 0) && aligned16(logits) && aligned16(out) && (!bwd || aligned16(gout));
  const int grid = stream_grid((long long)N * (vec ? HW / 4 : HW));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec) { if (bwd) launch_map<4, true>(C, grid, st, logits, gout, out, N, (int)HW, inv_log2c);
             else launch_map<4, false>(C, grid, st, logits, gout, out, N, (int)HW, inv_log2c); }
  else     { if (bwd) launch_map<1, true>(C, grid, st, logits, gout, out, N, (int)HW, inv_log2c);
             else launch_map<1, false>(C, grid, st, logits, gout, out, N, (int)HW, inv_log2c); }
  CNH_CUDA(cudaGetLastError());
  return CNH_OK;
}

extern "C" int cnh_entropy_map_fwd(const float* logits, float* out, int32_t N, int32_t C, int32_t H, int32_t W,
                                   cnh_stream_t stream) {
  return map_common(logits, nullptr, out, N, C, H, W, false, stream);
}

extern "C" int cnh_entropy_map_bwd(const float* logits, const float* grad_out, float* grad_in, int32_t N,
                                   int32_t C, int32_t H, int32_t W, cnh_stream_t stream) {
  return map_common(logits, grad_out, grad_in, N, C, H, W, true, stream);
}

extern "C" int cnh_bce_const(const float* y, float* grad, float* loss_out, int64_t n, float label,
                             cnh_stream_t stream) {
  CNH_REQUIRE(y && loss_out, CNH_E_NULL, "bce_const: y/loss_out is NULL");
  CNH_REQUIRE(n > 0, CNH_E_SHAPE, "bce_const: n=%lld", (long long)n);
  bce_const_kernel<<<1, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(y, grad, loss_out, n, label);
  CNH_CUDA(cudaGetLastError());
  return CNH_OK;
}

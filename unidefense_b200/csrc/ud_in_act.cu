// a2 -- decoder epilogues: InstanceNorm2d(affine) + Swish/ReLU, and Tanh      
//
// Reference: nn.InstanceNorm2d(affine=True) + MemoryEfficientSwish / nn.ReLU(inplace) after every
// decoder conv (model/unidefense.py:54-56,:61-98; :276-277,:286-305; :456-457,:466-497) and the
// final nn.Tanh (:101,:307,:499).  Per (n,c) plane of E = H*W elements:
//     mu = mean(x), var = mean((x-mu)^2) (biased), xh = (x-mu)*rsqrt(var+eps),
//     z = xh*gamma[c] + beta[c],  y = act(z)
// Backward (closed form; swish' from efficientnet/utils.py:73-77):
//     gz = gy*act'(z); S1 = sum gz; S2 = sum gz*xh
//     gx = gamma*rstd*(gz - S1/E - xh*S2/E); ggamma[c] = sum_n S2; gbeta[c] = sum_n S1
//
// Single-pass design: a plane lives in REGISTERS, split over `nw` warps that each own a contiguous slab
// (<= 9 float4 per lane forward, 5 + 5 backward), so x is read from HBM exactly once and y written once (the
// algorithmic minimum, 2E*4 B fwd / 3E*4 B bwd).  nw <= 8: G = nw warps per plane, 8/G planes per CTA;
// nw > 8: a thread-block cluster of nw/8 CTAs per plane exchanging partials through distributed shared memory.
// Statistics are thread-local two-pass (mean, centred M2) merged with Chan's formula, so there is no
// E[x^2]-E[x]^2 cancellation and only ONE cross-warp exchange per reduction.
// The forward can also emit the per-plane mean of y (the triplet features dec_out.mean([-2,-1]),
// model/unidefense.py:232-236) for free; the backward accepts its gradient.
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "../../include/unidefense_b200.h"
#include "ud_common.cuh"

namespace cg = cooperative_groups;

#define IA_THREADS 256
#define IA_WARPS (IA_THREADS / 32)
#define IA_VMAX_FWD 9
#define IA_VMAX_BWD 5
#define IA_MAX_CS 8

template <int ACT>
__device__ __forceinline__ float ia_sigmoid(float z) { return ud_sigmoid_fast(z); }
template <int ACT>
__device__ __forceinline__ float ia_act(float z) {
  if (ACT == UD_ACT_RELU) return fmaxf(z, 0.f);
  if (ACT == UD_ACT_SWISH) return z * ia_sigmoid<ACT>(z);
  return z;
}
// d act(z)/dz; swish: s*(1+z*(1-s))  (SwishImplementation.backward, model/efficientnet/utils.py:73-77)
template <int ACT>
__device__ __forceinline__ float ia_act_grad(float z) {
  if (ACT == UD_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (ACT == UD_ACT_SWISH) {
    const float s = ia_sigmoid<ACT>(z);
    return s * fmaf(z, 1.f - s, 1.f);
  }
  return 1.f;
}

// Chan et al. merge of (count, mean, M2) partial statistics: exact two-pass quality without a second
// group-wide reduction.
struct IaStat {
  float n, mean, m2;
};
// Branch-free: counts are 0 or >= 4, so max(n, 1) only changes the empty-with-empty case (f = 0, result = a).
__device__ __forceinline__ IaStat ia_merge(IaStat a, IaStat b) {
  const float n = a.n + b.n;
  const float d = b.mean - a.mean;
  const float f = b.n * ud_rcp_ftz(fmaxf(n, 1.f));
  IaStat r;
  r.n = n;
  r.mean = fmaf(d, f, a.mean);
  r.m2 = a.m2 + b.m2 + d * d * a.n * f;
  return r;
}
__device__ __forceinline__ IaStat ia_warp_merge(IaStat s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    IaStat t;
    t.n = __shfl_xor_sync(0xffffffffu, s.n, o);
    t.mean = __shfl_xor_sync(0xffffffffu, s.mean, o);
    t.m2 = __shfl_xor_sync(0xffffffffu, s.m2, o);
    // order the pair by lane so that both partners compute bit-identical results (one merge, operands selected)
    const bool lo = (threadIdx.x & o) == 0;
    IaStat a, b;
    a.n = lo ? s.n : t.n;          b.n = lo ? t.n : s.n;
    a.mean = lo ? s.mean : t.mean; b.mean = lo ? t.mean : s.mean;
    a.m2 = lo ? s.m2 : t.m2;       b.m2 = lo ? t.m2 : s.m2;
    s = ia_merge(a, b);
  }
  return s;
}

// Work decomposition: a plane is split over `nw` WARPS, each owning a contiguous slab of ceil(E4/nw)
// float4 that it keeps in registers (lane-strided, 512 contiguous bytes per warp access).
//   nw <= 8  : G = nw warps per plane, 8/G planes per 256-thread CTA (no cluster) -- small planes still put
//              ~150 KB per SM in flight;
//   nw  > 8  : a thread-block cluster of cs = nw/8 CTAs per plane.
// Reductions: per-warp shuffles, then ONE exchange: each warp publishes its partial to a shared-memory slot
// (for clusters: PUSHED into the slot array of every CTA of the cluster through DSMEM), one barrier, and
// every thread folds the slots of its plane in fixed order (deterministic, no trailing barrier).
struct IaGeom {
  int plane, wsub, nw, slot0, rank;
  bool live;
};
template <bool CLUSTER>
__device__ __forceinline__ IaGeom ia_geom(int planes, int G, int cs) {
  IaGeom g;
  const int warp = threadIdx.x >> 5;
  if (CLUSTER) {
    g.plane = blockIdx.x / cs;
    g.rank = blockIdx.x % cs;
    g.wsub = g.rank * IA_WARPS + warp;
    g.nw = cs * IA_WARPS;
    g.slot0 = 0;
  } else {
    g.plane = blockIdx.x * (IA_WARPS / G) + warp / G;
    g.rank = 0;
    g.wsub = warp % G;
    g.nw = G;
    g.slot0 = (warp / G) * G;
  }
  g.live = g.plane < planes;
  return g;
}

template <bool CLUSTER, int K>
__device__ __forceinline__ void ia_exchange(float (*slots)[K], const float (&mine)[K], int rank, int cs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (CLUSTER) {
    cg::cluster_group cluster = cg::this_cluster();
    if (lane < cs) {
      float(*remote)[K] = cluster.map_shared_rank(slots, lane);
#pragma unroll
      for (int k = 0; k < K; ++k) remote[rank * IA_WARPS + warp][k] = mine[k];
    }
    cluster.sync();
  } else {
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) slots[warp][k] = mine[k];
    }
    __syncthreads();
  }
}

// Folding the nw (<= 64) published slots: lane l takes slots l and l+32, then one shuffle tree -- 6 merges per
// thread instead of nw serial ones (which cost as many instructions as the plane's own arithmetic).
__device__ __forceinline__ IaStat ia_fold_stats(const float (*slots)[3], int slot0, int nw) {
  const int lane = threadIdx.x & 31;
  IaStat st = {0.f, 0.f, 0.f};
  if (lane < nw) st = IaStat{slots[slot0 + lane][0], slots[slot0 + lane][1], slots[slot0 + lane][2]};
  if (lane + 32 < nw) st = ia_merge(st, IaStat{slots[slot0 + lane + 32][0], slots[slot0 + lane + 32][1], slots[slot0 + lane + 32][2]});
  return ia_warp_merge(st);
}
template <int K>
__device__ __forceinline__ float ia_fold_sum(const float (*slots)[K], int k, int slot0, int nw) {
  const int lane = threadIdx.x & 31;
  float v = (lane < nw) ? slots[slot0 + lane][k] : 0.f;
  if (lane + 32 < nw) v += slots[slot0 + lane + 32][k];
  return ud_warp_sum(v);
}

// ---- I/O vector types: fp32 planes move as float4 (4 elements), bf16 planes as uint4 (8 elements) -------------
// All arithmetic and the statistics are fp32 either way; a bf16 store rounds to nearest even, exactly what the
// reference-side `.to(torch.bfloat16)` around an fp32 kernel would do (so the bf16 entry points are bit-identical to
// cast -> fp32 kernel -> cast, minus the two casting passes over HBM).
template <class IO>
struct IaVec;
template <>
struct IaVec<float> {
  typedef float4 V;
  static constexpr int E = 4;
  static __device__ __forceinline__ void load(const V* p, float (&f)[4]) {
    const float4 t = __ldcs(p);
    f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
  }
  static __device__ __forceinline__ void store(V* p, const float (&f)[4]) { *p = make_float4(f[0], f[1], f[2], f[3]); }
};
template <>
struct IaVec<__nv_bfloat16> {
  typedef uint4 V;
  static constexpr int E = 8;
  static __device__ __forceinline__ void load(const V* p, float (&f)[8]) {
    const uint4 t = __ldcs(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {                      // bf16 -> fp32 is a 16-bit shift
      f[2 * j] = __uint_as_float(w[j] << 16);
      f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(V* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);   // .x (low half) = first element
      w[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *p = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// Slot validity of a lane.  A warp's slab of `len` vectors is read lane-strided: slot i of lane l is vector
// i*32 + l.  FAST (warp-uniform: the slab fills all but possibly the last of the VPT slots -- every shape of the
// shipped configurations) makes slots 0..VPT-2 unconditional at compile time and leaves ONE runtime predicate for the
// last slot, so the unrolled body carries no per-slot compare/branch/address arithmetic (which, with the range-checked
// special functions, was 2/3 of the instructions of these issue-bound kernels).
template <bool FAST, int VPT>
__device__ __forceinline__ bool ia_ok(int i, int lane, int len, bool last_ok) {
  if (FAST) return (i < VPT - 1) ? true : last_ok;
  return i * 32 + lane < len;
}

template <class IO, bool CLUSTER, int ACT, bool WANT_MEAN, int VPT, bool FAST>
__device__ __forceinline__ void ia_fwd_body(const typename IaVec<IO>::V* __restrict__ xl,
                                            typename IaVec<IO>::V* __restrict__ yl, int lane, int len,
                                            float (*slots)[3], float (*yslots)[1], const IaGeom& ge, int cs, float g,
                                            float b, float eps, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                            float* __restrict__ ymean_out) {
  constexpr int E = IaVec<IO>::E;
  const bool last_ok = (VPT - 1) * 32 + lane < len;
  float v[VPT][E];
  int nv = 0;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    if (ia_ok<FAST, VPT>(i, lane, len, last_ok)) {
      IaVec<IO>::load(xl + i * 32, v[i]);
      ++nv;
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) v[i][e] = 0.f;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    float t = 0.f;
#pragma unroll
    for (int e = 0; e < E; e += 2) t += v[i][e] + v[i][e + 1];
    s += t;
  }
  // thread-local two-pass statistics, then one merge tree
  IaStat st;
  st.n = (float)(E * nv);
  st.mean = nv ? s * __frcp_rn(st.n) : 0.f;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    if (ia_ok<FAST, VPT>(i, lane, len, last_ok)) {
      float t = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float d = v[i][e] - st.mean;
        t = fmaf(d, d, t);
      }
      q += t;
    }
  }
  st.m2 = q;
  st = ia_warp_merge(st);
  const float mine[3] = {st.n, st.mean, st.m2};
  ia_exchange<CLUSTER, 3>(slots, mine, ge.rank, cs);
  const IaStat tot = ia_fold_stats(slots, ge.slot0, ge.nw);
  const float mu = tot.mean;
  const float rstd = rsqrtf(tot.m2 / fmaxf(tot.n, 1.f) + eps);
  const float a_ = g * rstd, b_ = b - mu * g * rstd;
  float ys = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    if (ia_ok<FAST, VPT>(i, lane, len, last_ok)) {
      float o[E];
#pragma unroll
      for (int e = 0; e < E; ++e) o[e] = ia_act<ACT>(fmaf(v[i][e], a_, b_));
      if (WANT_MEAN) {
#pragma unroll
        for (int e = 0; e < E; e += 2) ys += o[e] + o[e + 1];
      }
      IaVec<IO>::store(yl + i * 32, o);
    }
  }
  const bool writer = ge.live && ge.wsub == 0 && lane == 0;
  if (writer) {
    mean_out[ge.plane] = mu;
    rstd_out[ge.plane] = rstd;
  }
  if (WANT_MEAN) {
    ys = ud_warp_sum(ys);
    const float ymine[1] = {ys};
    ia_exchange<CLUSTER, 1>(yslots, ymine, ge.rank, cs);
    const float t = ia_fold_sum<1>(yslots, 0, ge.slot0, ge.nw);
    if (writer) ymean_out[ge.plane] = t / tot.n;
  }
}

// EV = vectors per plane (HW / 4 for fp32, HW / 8 for bf16)
template <class IO, bool CLUSTER, int ACT, bool WANT_MEAN, int VPT>
__global__ void __launch_bounds__(IA_THREADS, 4)
ia_fwd_kernel(const typename IaVec<IO>::V* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              typename IaVec<IO>::V* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
              float* __restrict__ ymean_out, int planes, int C, int EV, int G, int cs, float eps) {
  __shared__ float slots[IA_MAX_CS * IA_WARPS][3];
  __shared__ float yslots[IA_MAX_CS * IA_WARPS][1];
  const IaGeom ge = ia_geom<CLUSTER>(planes, G, cs);
  const int lane = threadIdx.x & 31;
  const int per_w = (EV + ge.nw - 1) / ge.nw;
  const int beg = ge.wsub * per_w;
  const int len = ge.live ? max(min(EV, beg + per_w) - beg, 0) : 0;     // vectors of this warp's slab (<= VPT*32)
  const int pl = ge.live ? ge.plane : 0;
  const typename IaVec<IO>::V* xl = x + (long long)pl * EV + beg + lane;
  typename IaVec<IO>::V* yl = y + (long long)pl * EV + beg + lane;
  const int ch = pl % C;
  const float g = gamma ? __ldg(gamma + ch) : 1.f;
  const float b = beta ? __ldg(beta + ch) : 0.f;
  // both branches run the same barriers / cluster syncs, so a CTA (cluster) may mix them
  if (len > (VPT - 1) * 32)
    ia_fwd_body<IO, CLUSTER, ACT, WANT_MEAN, VPT, true>(xl, yl, lane, len, slots, yslots, ge, cs, g, b, eps, mean_out,
                                                        rstd_out, ymean_out);
  else
    ia_fwd_body<IO, CLUSTER, ACT, WANT_MEAN, VPT, false>(xl, yl, lane, len, slots, yslots, ge, cs, g, b, eps, mean_out,
                                                         rstd_out, ymean_out);
}

template <class IO, bool CLUSTER, int ACT, int VPT, bool FAST>
__device__ __forceinline__ void ia_bwd_body(const typename IaVec<IO>::V* __restrict__ xl,
                                            const typename IaVec<IO>::V* __restrict__ gl,
                                            typename IaVec<IO>::V* __restrict__ ol, int lane, int len, float (*slots)[2],
                                            const IaGeom& ge, int cs, float g, float b, float mu, float rstd, float gadd,
                                            float invE, float* __restrict__ s1_out, float* __restrict__ s2_out) {
  constexpr int E = IaVec<IO>::E;
  const bool last_ok = (VPT - 1) * 32 + lane < len;
  float xh[VPT][E], gz[VPT][E];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {   // all loads in flight before any math
    if (ia_ok<FAST, VPT>(i, lane, len, last_ok)) {
      IaVec<IO>::load(xl + i * 32, xh[i]);
      IaVec<IO>::load(gl + i * 32, gz[i]);
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) xh[i][e] = gz[i][e] = 0.f;
    }
  }
  float s1 = 0.f, s2 = 0.f;
  const float nmr = -mu * rstd;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    if (ia_ok<FAST, VPT>(i, lane, len, last_ok)) {
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float h = fmaf(xh[i][e], rstd, nmr);
        const float z = (gz[i][e] + gadd) * ia_act_grad<ACT>(fmaf(h, g, b));
        t1 += z;
        t2 = fmaf(z, h, t2);
        xh[i][e] = h;
        gz[i][e] = z;
      }
      s1 += t1;
      s2 += t2;
    }
  }
  s1 = ud_warp_sum(s1);
  s2 = ud_warp_sum(s2);
  const float mine[2] = {s1, s2};
  ia_exchange<CLUSTER, 2>(slots, mine, ge.rank, cs);
  // plain broadcast reads: measured faster here than a second shuffle tree (two dependent 5-step chains)
  float S1 = 0.f, S2 = 0.f;
  for (int i = 0; i < ge.nw; ++i) {
    const float2 sv = *reinterpret_cast<const float2*>(&slots[ge.slot0 + i][0]);
    S1 += sv.x;
    S2 += sv.y;
  }
  const float m1 = S1 * invE, m2 = S2 * invE, k = g * rstd;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    if (ia_ok<FAST, VPT>(i, lane, len, last_ok)) {
      float o[E];
#pragma unroll
      for (int e = 0; e < E; ++e) o[e] = k * (gz[i][e] - m1 - xh[i][e] * m2);
      IaVec<IO>::store(ol + i * 32, o);
    }
  }
  if (ge.live && ge.wsub == 0 && lane == 0) {
    s1_out[ge.plane] = S1;
    s2_out[ge.plane] = S2;
  }
}

template <class IO, bool CLUSTER, int ACT, int VPT>
__global__ void __launch_bounds__(IA_THREADS, 4)
ia_bwd_kernel(const typename IaVec<IO>::V* __restrict__ x, const typename IaVec<IO>::V* __restrict__ gy,
              const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
              const float* __restrict__ rstd_in, const float* __restrict__ g_ymean, typename IaVec<IO>::V* __restrict__ gx,
              float* __restrict__ s1_out, float* __restrict__ s2_out, int planes, int C, int EV, int G, int cs) {
  __shared__ __align__(8) float slots[IA_MAX_CS * IA_WARPS][2];
  const IaGeom ge = ia_geom<CLUSTER>(planes, G, cs);
  const int lane = threadIdx.x & 31;
  const int per_w = (EV + ge.nw - 1) / ge.nw;
  const int beg = ge.wsub * per_w;
  const int len = ge.live ? max(min(EV, beg + per_w) - beg, 0) : 0;
  const int pl = ge.live ? ge.plane : 0;
  const long long off = (long long)pl * EV + beg + lane;
  const int ch = pl % C;
  const float g = gamma ? __ldg(gamma + ch) : 1.f;
  const float b = beta ? __ldg(beta + ch) : 0.f;
  const float mu = mean[pl], rstd = rstd_in[pl];
  const float invE = 1.f / ((float)IaVec<IO>::E * (float)EV);
  const float gadd = g_ymean ? g_ymean[pl] * invE : 0.f;  // d mean(y) term
  if (len > (VPT - 1) * 32)
    ia_bwd_body<IO, CLUSTER, ACT, VPT, true>(x + off, gy + off, gx + off, lane, len, slots, ge, cs, g, b, mu, rstd, gadd,
                                             invE, s1_out, s2_out);
  else
    ia_bwd_body<IO, CLUSTER, ACT, VPT, false>(x + off, gy + off, gx + off, lane, len, slots, ge, cs, g, b, mu, rstd, gadd,
                                              invE, s1_out, s2_out);
}

// ---- generic fallback: any plane size (odd E, huge planes): one CTA per plane, re-reads hit L1/L2 ----
__device__ __forceinline__ float ia_ld(const float* p) { return *p; }
__device__ __forceinline__ float ia_ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void ia_st(float* p, float v) { *p = v; }
__device__ __forceinline__ void ia_st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <class IO>
__global__ void __launch_bounds__(IA_THREADS)
ia_fwd_generic_kernel(const IO* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      IO* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                      float* __restrict__ ymean_out, int C, int E, float eps, int act) {
  __shared__ float red[33];
  const int plane = blockIdx.x;
  const IO* xp = x + (long long)plane * E;
  IO* yp = y + (long long)plane * E;
  float s = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) s += ia_ld(xp + i);
  const float mu = ud_block_sum(s, red) / (float)E;
  float q = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    const float d = ia_ld(xp + i) - mu;
    q += d * d;
  }
  const float rstd = rsqrtf(ud_block_sum(q, red) / (float)E + eps);
  const int ch = plane % C;
  const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
  const float a_ = g * rstd, b_ = b - mu * g * rstd;
  float ys = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    const float o = ud_act_fwd(fmaf(ia_ld(xp + i), a_, b_), act);
    ys += o;
    ia_st(yp + i, o);
  }
  if (ymean_out != nullptr) {
    const float t = ud_block_sum(ys, red) / (float)E;
    if (threadIdx.x == 0) ymean_out[plane] = t;
  }
  if (threadIdx.x == 0) {
    mean_out[plane] = mu;
    rstd_out[plane] = rstd;
  }
}

template <class IO>
__global__ void __launch_bounds__(IA_THREADS)
ia_bwd_generic_kernel(const IO* __restrict__ x, const IO* __restrict__ gy, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ mean,
                      const float* __restrict__ rstd_in, const float* __restrict__ g_ymean, IO* __restrict__ gx,
                      float* __restrict__ s1_out, float* __restrict__ s2_out, int C, int E, int act) {
  __shared__ float red[33];
  const int plane = blockIdx.x;
  const long long base = (long long)plane * E;
  const int ch = plane % C;
  const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
  const float mu = mean[plane], rstd = rstd_in[plane];
  const float gadd = g_ymean ? g_ymean[plane] / (float)E : 0.f;
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    const float h = (ia_ld(x + base + i) - mu) * rstd;
    const float z = (ia_ld(gy + base + i) + gadd) * ud_act_grad(fmaf(h, g, b), act);
    s1 += z;
    s2 += z * h;
  }
  const float S1 = ud_block_sum(s1, red);
  const float S2 = ud_block_sum(s2, red);
  const float m1 = S1 / (float)E, m2 = S2 / (float)E, k = g * rstd;
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    const float h = (ia_ld(x + base + i) - mu) * rstd;
    const float z = (ia_ld(gy + base + i) + gadd) * ud_act_grad(fmaf(h, g, b), act);
    ia_st(gx + base + i, k * (z - m1 - h * m2));
  }
  if (threadIdx.x == 0) {
    s1_out[plane] = S1;
    s2_out[plane] = S2;
  }
}

// ggamma[c] = sum_n s2[n,c], gbeta[c] = sum_n s1[n,c]  (fixed order).  A separate 5 us launch on purpose: folding it into
// the backward kernel with the last-CTA-reduces pattern (ticket + __threadfence) was measured SLOWER (bf16 20 x 192^2
// backward 68 -> 75 us): the device-wide fence at the end of every CTA has to wait for that CTA's gx stores.
__global__ void ia_param_grad_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                                     float* __restrict__ ggamma, float* __restrict__ gbeta, int N, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int n = 0; n < N; ++n) {
    a += s2[n * C + c];
    b += s1[n * C + c];
  }
  if (ggamma) ggamma[c] = a;
  if (gbeta) gbeta[c] = b;
}

// smallest power-of-two warp count nw (<= 64) whose per-warp slab fits vmax vectors (of epv elements) per lane
static bool ia_pick(int E, int epv, int vmax, int* G, int* cs, int* vpt) {
  if (E % epv != 0) return false;
  const int EV = E / epv;
  for (int nw = 1; nw <= IA_MAX_CS * IA_WARPS; nw *= 2) {
    const int per = (EV + nw - 1) / nw;
    const int v = (per + 31) / 32;
    if (v <= vmax) {
      *G = nw <= IA_WARPS ? nw : IA_WARPS;
      *cs = nw <= IA_WARPS ? 1 : nw / IA_WARPS;
      *vpt = v;
      return true;
    }
  }
  return false;
}

// expands STMT once per activation code with ACT_ bound to the compile-time constant
#define IA_ACT_SWITCH(act, STMT)                                            \
  do {                                                                      \
    if ((act) == UD_ACT_SWISH) { constexpr int ACT_ = UD_ACT_SWISH; STMT; } \
    else if ((act) == UD_ACT_RELU) { constexpr int ACT_ = UD_ACT_RELU; STMT; } \
    else { constexpr int ACT_ = UD_ACT_NONE; STMT; }                        \
  } while (0)

template <class K, class... Args>
static int ia_launch(K kernel, int blocks, int cs, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(IA_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cs > 1 ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  ud_count_launch();
  if (e != cudaSuccess) {
    ud_set_error("in_act: cudaLaunchKernelEx failed: %s", cudaGetErrorString(e));
    return UD_ERR_CUDA;
  }
  return UD_OK;
}

// registers: a slot holds 4 (fp32) or 8 (bf16) values per lane -> 9/5 slots forward, 5/3 (x and gy) backward
template <class IO> struct IaBudget;
template <> struct IaBudget<float> { static constexpr int FWD = 9, FWD_SMALL = 5, BWD = 5; };
template <> struct IaBudget<__nv_bfloat16> { static constexpr int FWD = 5, FWD_SMALL = 3, BWD = 3; };

template <class IO>
static int ia_fwd_impl(const IO* x, const float* gamma, const float* beta, IO* y, float* mean, float* rstd, float* ymean,
                       int N, int C, int HW, float eps, int act, cudaStream_t stream) {
  UD_REQUIRE(N >= 0 && C >= 1 && HW >= 1, UD_ERR_INVALID, "in_act_fwd: bad shape N=%d C=%d HW=%d", N, C, HW);
  UD_REQUIRE(act >= UD_ACT_NONE && act <= UD_ACT_SWISH, UD_ERR_INVALID, "in_act_fwd: bad activation %d", act);
  if (N == 0) return UD_OK;
  UD_REQUIRE(x && y && mean && rstd, UD_ERR_INVALID, "in_act_fwd: null pointer");
  typedef typename IaVec<IO>::V V;
  constexpr int EPV = IaVec<IO>::E;
  const int planes = N * C;
  int cs = 1, vpt = 1;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  int G = 1;
  if (aligned && ia_pick(HW, EPV, IaBudget<IO>::FWD, &G, &cs, &vpt)) {
    const V* xv = reinterpret_cast<const V*>(x);
    V* yv = reinterpret_cast<V*>(y);
#define IA_FWD_V(CL, WM, BLOCKS, VP)                                                                                     \
  IA_ACT_SWITCH(act, return ia_launch(ia_fwd_kernel<IO, CL, ACT_, WM, VP>, BLOCKS, cs, stream, xv, gamma, beta, yv, mean, \
                                      rstd, ymean, planes, C, HW / EPV, G, cs, eps))
#define IA_FWD(CL, WM, BLOCKS)                                                  \
  do {                                                                          \
    if (vpt <= IaBudget<IO>::FWD_SMALL) IA_FWD_V(CL, WM, BLOCKS, IaBudget<IO>::FWD_SMALL); \
    IA_FWD_V(CL, WM, BLOCKS, IaBudget<IO>::FWD);                                \
  } while (0)
    if (cs == 1 && ymean) IA_FWD(false, true, ud_cdiv(planes, IA_WARPS / G));
    if (cs == 1) IA_FWD(false, false, ud_cdiv(planes, IA_WARPS / G));
    if (ymean) IA_FWD(true, true, planes * cs);
    IA_FWD(true, false, planes * cs);
#undef IA_FWD
#undef IA_FWD_V
  }
  ia_fwd_generic_kernel<IO><<<planes, IA_THREADS, 0, stream>>>(x, gamma, beta, y, mean, rstd, ymean, C, HW, eps, act);
  return ud_check_launch("ia_fwd_generic");
}

extern "C" int ud_in_act_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                             float* rstd, float* ymean, int N, int C, int HW, float eps, int act,
                             cudaStream_t stream) {
  return ia_fwd_impl<float>(x, gamma, beta, y, mean, rstd, ymean, N, C, HW, eps, act, stream);
}
extern "C" int ud_in_act_fwd_bf16(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                  float* rstd, float* ymean, int N, int C, int HW, float eps, int act,
                                  cudaStream_t stream) {
  return ia_fwd_impl<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(x), gamma, beta, static_cast<__nv_bfloat16*>(y), mean,
                                    rstd, ymean, N, C, HW, eps, act, stream);
}

extern "C" size_t ud_in_act_bwd_workspace_bytes(int N, int C) { return sizeof(float) * 2ull * N * C; }

template <class IO>
static int ia_bwd_impl(const IO* x, const IO* gy, const float* gamma, const float* beta, const float* mean,
                       const float* rstd, const float* g_ymean, IO* gx, float* ggamma, float* gbeta, void* ws,
                       size_t ws_bytes, int N, int C, int HW, int act, cudaStream_t stream) {
  UD_REQUIRE(N >= 0 && C >= 1 && HW >= 1, UD_ERR_INVALID, "in_act_bwd: bad shape N=%d C=%d HW=%d", N, C, HW);
  UD_REQUIRE(act >= UD_ACT_NONE && act <= UD_ACT_SWISH, UD_ERR_INVALID, "in_act_bwd: bad activation %d", act);
  if (N == 0) {
    if (ggamma) UD_CUDA(cudaMemsetAsync(ggamma, 0, sizeof(float) * C, stream));
    if (gbeta) UD_CUDA(cudaMemsetAsync(gbeta, 0, sizeof(float) * C, stream));
    return UD_OK;
  }
  UD_REQUIRE(x && gy && mean && rstd && gx && ws, UD_ERR_INVALID, "in_act_bwd: null pointer");
  UD_REQUIRE(ws_bytes >= ud_in_act_bwd_workspace_bytes(N, C), UD_ERR_WORKSPACE, "in_act_bwd: workspace too small");
  typedef typename IaVec<IO>::V V;
  constexpr int EPV = IaVec<IO>::E;
  float* s1 = static_cast<float*>(ws);
  float* s2 = s1 + (size_t)N * C;
  const int planes = N * C;
  int cs = 1, vpt = 1, rc;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy) |
                         reinterpret_cast<uintptr_t>(gx)) & 15) == 0;
  int G = 1;
  if (aligned && ia_pick(HW, EPV, IaBudget<IO>::BWD, &G, &cs, &vpt)) {
    const V* xv = reinterpret_cast<const V*>(x);
    const V* gv = reinterpret_cast<const V*>(gy);
    V* ov = reinterpret_cast<V*>(gx);
    rc = UD_OK;
    if (cs == 1)
      IA_ACT_SWITCH(act, rc = ia_launch(ia_bwd_kernel<IO, false, ACT_, IaBudget<IO>::BWD>, ud_cdiv(planes, IA_WARPS / G), 1,
                                        stream, xv, gv, gamma, beta, mean, rstd, g_ymean, ov, s1, s2, planes, C, HW / EPV, G,
                                        cs));
    else
      IA_ACT_SWITCH(act, rc = ia_launch(ia_bwd_kernel<IO, true, ACT_, IaBudget<IO>::BWD>, planes * cs, cs, stream, xv, gv,
                                        gamma, beta, mean, rstd, g_ymean, ov, s1, s2, planes, C, HW / EPV, G, cs));
    if (rc != UD_OK) return rc;
  } else {
    ia_bwd_generic_kernel<IO><<<planes, IA_THREADS, 0, stream>>>(x, gy, gamma, beta, mean, rstd, g_ymean, gx, s1, s2, C, HW,
                                                                  act);
    if ((rc = ud_check_launch("ia_bwd_generic")) != UD_OK) return rc;
  }
  if (ggamma || gbeta) {
    ia_param_grad_kernel<<<ud_cdiv(C, 128), 128, 0, stream>>>(s1, s2, ggamma, gbeta, N, C);
    return ud_check_launch("ia_param_grad");
  }
  return UD_OK;
}

extern "C" int ud_in_act_bwd(const float* x, const float* gy, const float* gamma, const float* beta,
                             const float* mean, const float* rstd, const float* g_ymean, float* gx,
                             float* ggamma, float* gbeta, void* ws, size_t ws_bytes, int N, int C, int HW, int act,
                             cudaStream_t stream) {
  return ia_bwd_impl<float>(x, gy, gamma, beta, mean, rstd, g_ymean, gx, ggamma, gbeta, ws, ws_bytes, N, C, HW, act, stream);
}
extern "C" int ud_in_act_bwd_bf16(const void* x, const void* gy, const float* gamma, const float* beta, const float* mean,
                                  const float* rstd, const float* g_ymean, void* gx, float* ggamma, float* gbeta, void* ws,
                                  size_t ws_bytes, int N, int C, int HW, int act, cudaStream_t stream) {
  return ia_bwd_impl<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(gy), gamma, beta,
                                    mean, rstd, g_ymean, static_cast<__nv_bfloat16*>(gx), ggamma, gbeta, ws, ws_bytes, N, C,
                                    HW, act, stream);
}

// ---- tanh epilogue (model/unidefense.py:101,:307,:499) ------------------------------------------
__global__ void ia_tanh_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = tanhf(x[i]);
}
__global__ void ia_tanh_bwd_kernel(const float* __restrict__ y, const float* __restrict__ gy, float* __restrict__ gx,
                                   long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float t = y[i];
    gx[i] = gy[i] * (1.f - t * t);
  }
}

extern "C" int ud_tanh_fwd(const float* x, float* y, long long n, cudaStream_t stream) {
  UD_REQUIRE(n >= 0, UD_ERR_INVALID, "tanh_fwd: negative size");
  if (n == 0) return UD_OK;
  UD_REQUIRE(x && y, UD_ERR_INVALID, "tanh_fwd: null pointer");
  const int blocks = (int)min((long long)UD_NUM_SMS * 8, (n + 255) / 256);
  ia_tanh_fwd_kernel<<<blocks, 256, 0, stream>>>(x, y, n);
  return ud_check_launch("tanh_fwd");
}
extern "C" int ud_tanh_bwd(const float* y, const float* gy, float* gx, long long n, cudaStream_t stream) {
  UD_REQUIRE(n >= 0, UD_ERR_INVALID, "tanh_bwd: negative size");
  if (n == 0) return UD_OK;
  UD_REQUIRE(y && gy && gx, UD_ERR_INVALID, "tanh_bwd: null pointer");
  const int blocks = (int)min((long long)UD_NUM_SMS * 8, (n + 255) / 256);
  ia_tanh_bwd_kernel<<<blocks, 256, 0, stream>>>(y, gy, gx, n);
  return ud_check_launch("tanh_bwd");
}

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
// Single-pass design: a plane (or a 1/CS slice of it) lives in REGISTERS -- 256 threads x up to 9
// float4 -- so x is read from HBM exactly once and y written once (the algorithmic minimum,
// 2E*4 B fwd / 3E*4 B bwd).  Planes too large for one CTA are split over a thread-block cluster
// of CS in {2,4,8} CTAs that exchange their partial sums through distributed shared memory.
// Two-pass statistics (mean, then centred second moment) from the registers, like ATen's
// instance_norm, so no E[x^2]-E[x]^2 cancellation.
// The forward can also emit the per-plane mean of y (the triplet features dec_out.mean([-2,-1]),
// model/unidefense.py:232-236) for free; the backward accepts its gradient.
#include <cooperative_groups.h>

#include "../../include/unidefense_b200.h"
#include "ud_common.cuh"

namespace cg = cooperative_groups;

#define IA_THREADS 256
#define IA_VMAX_FWD 9
#define IA_VMAX_BWD 5

// cluster-wide (or block-wide when CS==1) sum, broadcast to all threads of all CTAs.
// `slot` is a per-CTA shared float written by thread 0; `red` is block scratch (33 floats).
template <bool CLUSTER>
__device__ __forceinline__ float ia_group_sum(float v, float* red, float* slot, int cs) {
  float t = ud_block_sum(v, red);
  if (CLUSTER) {
    cg::cluster_group cluster = cg::this_cluster();
    if (threadIdx.x == 0) *slot = t;
    cluster.sync();
    float tot = 0.f;
    for (int r = 0; r < cs; ++r) tot += *cluster.map_shared_rank(slot, r);  // fixed order: deterministic
    t = tot;
  }
  return t;
}

template <bool CLUSTER>
__global__ void __launch_bounds__(IA_THREADS)
ia_fwd_kernel(const float4* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              float4* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
              float* __restrict__ ymean_out, int C, int E4, int cs, int vpt, float eps, int act) {
  __shared__ float red[33];
  __shared__ float slots[3];
  const int plane = CLUSTER ? (blockIdx.x / cs) : blockIdx.x;
  const int rank = CLUSTER ? (blockIdx.x % cs) : 0;
  const int per_cta = (E4 + cs - 1) / cs;
  const int beg = rank * per_cta;
  const int end = min(E4, beg + per_cta);
  const float4* xp = x + (long long)plane * E4;
  float4 v[IA_VMAX_FWD];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < IA_VMAX_FWD; ++i) {
    const int idx = beg + i * IA_THREADS + threadIdx.x;
    if (i < vpt && idx < end) {
      v[i] = __ldcs(xp + idx);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float invE = 1.f / (4.f * (float)E4);
  const float mu = ia_group_sum<CLUSTER>(s, red, &slots[0], cs) * invE;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < IA_VMAX_FWD; ++i) {
    const int idx = beg + i * IA_THREADS + threadIdx.x;
    if (i < vpt && idx < end) {
      const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float var = ia_group_sum<CLUSTER>(q, red, &slots[1], cs) * invE;
  const float rstd = rsqrtf(var + eps);
  const int ch = plane % C;
  const float g = gamma ? __ldg(gamma + ch) : 1.f;
  const float b = beta ? __ldg(beta + ch) : 0.f;
  const float a_ = g * rstd, b_ = b - mu * g * rstd;
  float4* yp = y + (long long)plane * E4;
  float ys = 0.f;
#pragma unroll
  for (int i = 0; i < IA_VMAX_FWD; ++i) {
    const int idx = beg + i * IA_THREADS + threadIdx.x;
    if (i < vpt && idx < end) {
      float4 o;
      o.x = ud_act_fwd(fmaf(v[i].x, a_, b_), act);
      o.y = ud_act_fwd(fmaf(v[i].y, a_, b_), act);
      o.z = ud_act_fwd(fmaf(v[i].z, a_, b_), act);
      o.w = ud_act_fwd(fmaf(v[i].w, a_, b_), act);
      ys += (o.x + o.y) + (o.z + o.w);
      yp[idx] = o;
    }
  }
  if (ymean_out != nullptr) {
    const float t = ia_group_sum<CLUSTER>(ys, red, &slots[2], cs) * invE;
    if (rank == 0 && threadIdx.x == 0) ymean_out[plane] = t;
  }
  if (rank == 0 && threadIdx.x == 0) {
    mean_out[plane] = mu;
    rstd_out[plane] = rstd;
  }
  if (CLUSTER) cg::this_cluster().sync();  // keep our smem alive until every peer has read it
}

template <bool CLUSTER>
__global__ void __launch_bounds__(IA_THREADS)
ia_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ gy, const float* __restrict__ gamma,
              const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd_in,
              const float* __restrict__ g_ymean, float4* __restrict__ gx, float* __restrict__ s1_out,
              float* __restrict__ s2_out, int C, int E4, int cs, int vpt, int act) {
  __shared__ float red[33];
  __shared__ float slots[2];
  const int plane = CLUSTER ? (blockIdx.x / cs) : blockIdx.x;
  const int rank = CLUSTER ? (blockIdx.x % cs) : 0;
  const int per_cta = (E4 + cs - 1) / cs;
  const int beg = rank * per_cta;
  const int end = min(E4, beg + per_cta);
  const long long base = (long long)plane * E4;
  const int ch = plane % C;
  const float g = gamma ? __ldg(gamma + ch) : 1.f;
  const float b = beta ? __ldg(beta + ch) : 0.f;
  const float mu = mean[plane], rstd = rstd_in[plane];
  const float invE = 1.f / (4.f * (float)E4);
  const float gadd = g_ymean ? g_ymean[plane] * invE : 0.f;  // d mean(y) term
  float4 xh[IA_VMAX_BWD], gz[IA_VMAX_BWD];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < IA_VMAX_BWD; ++i) {
    const int idx = beg + i * IA_THREADS + threadIdx.x;
    if (i < vpt && idx < end) {
      const float4 xv = __ldcs(x + base + idx);
      const float4 gv = __ldcs(gy + base + idx);
      float4 h, z;
      h.x = (xv.x - mu) * rstd; h.y = (xv.y - mu) * rstd; h.z = (xv.z - mu) * rstd; h.w = (xv.w - mu) * rstd;
      z.x = (gv.x + gadd) * ud_act_grad(fmaf(h.x, g, b), act);
      z.y = (gv.y + gadd) * ud_act_grad(fmaf(h.y, g, b), act);
      z.z = (gv.z + gadd) * ud_act_grad(fmaf(h.z, g, b), act);
      z.w = (gv.w + gadd) * ud_act_grad(fmaf(h.w, g, b), act);
      s1 += (z.x + z.y) + (z.z + z.w);
      s2 += (z.x * h.x + z.y * h.y) + (z.z * h.z + z.w * h.w);
      xh[i] = h;
      gz[i] = z;
    } else {
      xh[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      gz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float S1 = ia_group_sum<CLUSTER>(s1, red, &slots[0], cs);
  const float S2 = ia_group_sum<CLUSTER>(s2, red, &slots[1], cs);
  const float m1 = S1 * invE, m2 = S2 * invE, k = g * rstd;
#pragma unroll
  for (int i = 0; i < IA_VMAX_BWD; ++i) {
    const int idx = beg + i * IA_THREADS + threadIdx.x;
    if (i < vpt && idx < end) {
      float4 o;
      o.x = k * (gz[i].x - m1 - xh[i].x * m2);
      o.y = k * (gz[i].y - m1 - xh[i].y * m2);
      o.z = k * (gz[i].z - m1 - xh[i].z * m2);
      o.w = k * (gz[i].w - m1 - xh[i].w * m2);
      gx[base + idx] = o;
    }
  }
  if (rank == 0 && threadIdx.x == 0) {
    s1_out[plane] = S1;
    s2_out[plane] = S2;
  }
  if (CLUSTER) cg::this_cluster().sync();
}

// ---- generic fallback: any plane size (odd E, huge planes): one CTA per plane, re-reads hit L1/L2 ----
__global__ void __launch_bounds__(IA_THREADS)
ia_fwd_generic_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                      float* __restrict__ ymean_out, int C, int E, float eps, int act) {
  __shared__ float red[33];
  const int plane = blockIdx.x;
  const float* xp = x + (long long)plane * E;
  float* yp = y + (long long)plane * E;
  float s = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) s += xp[i];
  const float mu = ud_block_sum(s, red) / (float)E;
  float q = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    const float d = xp[i] - mu;
    q += d * d;
  }
  const float rstd = rsqrtf(ud_block_sum(q, red) / (float)E + eps);
  const int ch = plane % C;
  const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
  const float a_ = g * rstd, b_ = b - mu * g * rstd;
  float ys = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    const float o = ud_act_fwd(fmaf(xp[i], a_, b_), act);
    ys += o;
    yp[i] = o;
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

__global__ void __launch_bounds__(IA_THREADS)
ia_bwd_generic_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ mean,
                      const float* __restrict__ rstd_in, const float* __restrict__ g_ymean, float* __restrict__ gx,
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
    const float h = (x[base + i] - mu) * rstd;
    const float z = (gy[base + i] + gadd) * ud_act_grad(fmaf(h, g, b), act);
    s1 += z;
    s2 += z * h;
  }
  const float S1 = ud_block_sum(s1, red);
  const float S2 = ud_block_sum(s2, red);
  const float m1 = S1 / (float)E, m2 = S2 / (float)E, k = g * rstd;
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    const float h = (x[base + i] - mu) * rstd;
    const float z = (gy[base + i] + gadd) * ud_act_grad(fmaf(h, g, b), act);
    gx[base + i] = k * (z - m1 - h * m2);
  }
  if (threadIdx.x == 0) {
    s1_out[plane] = S1;
    s2_out[plane] = S2;
  }
}

// ggamma[c] = sum_n s2[n,c], gbeta[c] = sum_n s1[n,c]  (fixed order)
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

static bool ia_pick(int E, int vmax, int* cs, int* vpt) {
  if (E % 4 != 0) return false;
  const int E4 = E / 4;
  for (int c = 1; c <= 8; c *= 2) {
    const int per = (E4 + c - 1) / c;
    const int v = (per + IA_THREADS - 1) / IA_THREADS;
    if (v <= vmax) {
      *cs = c;
      *vpt = v;
      return true;
    }
  }
  return false;
}

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

extern "C" int ud_in_act_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                             float* rstd, float* ymean, int N, int C, int HW, float eps, int act,
                             cudaStream_t stream) {
  UD_REQUIRE(N >= 0 && C >= 1 && HW >= 1, UD_ERR_INVALID, "in_act_fwd: bad shape N=%d C=%d HW=%d", N, C, HW);
  UD_REQUIRE(act >= UD_ACT_NONE && act <= UD_ACT_SWISH, UD_ERR_INVALID, "in_act_fwd: bad activation %d", act);
  if (N == 0) return UD_OK;
  UD_REQUIRE(x && y && mean && rstd, UD_ERR_INVALID, "in_act_fwd: null pointer");
  const int planes = N * C;
  int cs = 1, vpt = 1;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (aligned && ia_pick(HW, IA_VMAX_FWD, &cs, &vpt)) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    if (cs == 1)
      return ia_launch(ia_fwd_kernel<false>, planes, 1, stream, x4, gamma, beta, y4, mean, rstd, ymean, C, HW / 4,
                       cs, vpt, eps, act);
    return ia_launch(ia_fwd_kernel<true>, planes * cs, cs, stream, x4, gamma, beta, y4, mean, rstd, ymean, C,
                     HW / 4, cs, vpt, eps, act);
  }
  ia_fwd_generic_kernel<<<planes, IA_THREADS, 0, stream>>>(x, gamma, beta, y, mean, rstd, ymean, C, HW, eps, act);
  return ud_check_launch("ia_fwd_generic");
}

extern "C" size_t ud_in_act_bwd_workspace_bytes(int N, int C) { return sizeof(float) * 2ull * N * C; }

extern "C" int ud_in_act_bwd(const float* x, const float* gy, const float* gamma, const float* beta,
                             const float* mean, const float* rstd, const float* g_ymean, float* gx,
                             float* ggamma, float* gbeta, void* ws, size_t ws_bytes, int N, int C, int HW, int act,
                             cudaStream_t stream) {
  UD_REQUIRE(N >= 0 && C >= 1 && HW >= 1, UD_ERR_INVALID, "in_act_bwd: bad shape N=%d C=%d HW=%d", N, C, HW);
  UD_REQUIRE(act >= UD_ACT_NONE && act <= UD_ACT_SWISH, UD_ERR_INVALID, "in_act_bwd: bad activation %d", act);
  if (N == 0) {
    if (ggamma) UD_CUDA(cudaMemsetAsync(ggamma, 0, sizeof(float) * C, stream));
    if (gbeta) UD_CUDA(cudaMemsetAsync(gbeta, 0, sizeof(float) * C, stream));
    return UD_OK;
  }
  UD_REQUIRE(x && gy && mean && rstd && gx && ws, UD_ERR_INVALID, "in_act_bwd: null pointer");
  UD_REQUIRE(ws_bytes >= ud_in_act_bwd_workspace_bytes(N, C), UD_ERR_WORKSPACE, "in_act_bwd: workspace too small");
  float* s1 = static_cast<float*>(ws);
  float* s2 = s1 + (size_t)N * C;
  const int planes = N * C;
  int cs = 1, vpt = 1, rc;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy) |
                         reinterpret_cast<uintptr_t>(gx)) & 15) == 0;
  if (aligned && ia_pick(HW, IA_VMAX_BWD, &cs, &vpt)) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* g4 = reinterpret_cast<const float4*>(gy);
    float4* o4 = reinterpret_cast<float4*>(gx);
    if (cs == 1)
      rc = ia_launch(ia_bwd_kernel<false>, planes, 1, stream, x4, g4, gamma, beta, mean, rstd, g_ymean, o4, s1, s2,
                     C, HW / 4, cs, vpt, act);
    else
      rc = ia_launch(ia_bwd_kernel<true>, planes * cs, cs, stream, x4, g4, gamma, beta, mean, rstd, g_ymean, o4, s1,
                     s2, C, HW / 4, cs, vpt, act);
    if (rc != UD_OK) return rc;
  } else {
    ia_bwd_generic_kernel<<<planes, IA_THREADS, 0, stream>>>(x, gy, gamma, beta, mean, rstd, g_ymean, gx, s1, s2, C,
                                                             HW, act);
    if ((rc = ud_check_launch("ia_bwd_generic")) != UD_OK) return rc;
  }
  if (ggamma || gbeta) {
    ia_param_grad_kernel<<<ud_cdiv(C, 128), 128, 0, stream>>>(s1, s2, ggamma, gbeta, N, C);
    return ud_check_launch("ia_param_grad");
  }
  return UD_OK;
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

// a9 / a10 / a11 -- losses, each fused forward + gradient, no host syncs     
//
//  a9  AsymmetricalWeightedTripletLoss   loss/triplet_loss.py:16-82
//  a10 FactorizationLoss                 loss/calib_loss.py:17-28
//  a11 mask KL alignment                 engine/abstract_engine.py:331-346 (KLDivLoss batchmean, log_target)
//
// Every kernel here also emits d loss / d input for upstream gradient 1, so autograd's backward is
// a scalar multiply; the reference's boolean-mask indexing / torch.where host sync
// (triplet_loss.py:39,51,53) disappears.  All reductions are fixed-order (deterministic).
#include "../../include/unidefense_b200.h"
#include "ud_common.cuh"

#define LS_MAX_N 128

// ---------------------------------------------------------------- a9 triplet
// one CTA of 512 threads; dynamic smem: dist [N][N], H [N][N] (later Hs = H + H^T), sq [N], gu/fp/fn [N] each, then
// (when it fits) a copy of feat with row stride cs = c padded to 4 (mod 32) floats.  The kernel is a chain of short
// dependent phases on one SM, so its time is latency: every phase is laid out so that ALL threads work at once --
// one thread per (anchor, j) pair for the distances (float4 dot products from the padded copy: conflict-free),
// one warp per anchor for the softmax-weighted sums, one thread per gradient element.
#define LS_TRI_THREADS 512
__global__ void __launch_bounds__(LS_TRI_THREADS)
ls_triplet_kernel(const float* __restrict__ feat, const long long* __restrict__ labels, float* __restrict__ loss,
                  float* __restrict__ gfeat, int N, int c, int cs, int staged) {
  extern __shared__ float smf[];
  __shared__ int s_nr;
  __shared__ float red[33];
  __shared__ int s_lab[LS_MAX_N];
  const int tid = threadIdx.x, nth = blockDim.x;
  if (tid == 0) s_nr = 0;
  __syncthreads();
  for (int i = tid; i < N; i += nth) {
    const int l = (int)labels[i];
    s_lab[i] = l;
    if (l == 0) atomicAdd(&s_nr, 1);  // N_real = #(labels == 0)   (triplet_loss.py:39)
  }
  float* dist = smf;
  float* Hm = dist + N * N;
  float* sq = Hm + N * N;
  float* gu = sq + N;
  float* fpv = gu + N;
  float* fnv = fpv + N;
  float* fstage = smf + ((2 * N * N + 4 * N + 3) & ~3);   // 16-byte aligned (host sizes it the same way)
  if (staged) {
    for (int i = tid; i < N * cs; i += nth) {
      const int r = i / cs, k = i - r * cs;
      fstage[i] = (k < c) ? __ldg(feat + (long long)r * c + k) : 0.f;     // zero padding: dot products may run over cs
    }
  }
  const float* f = staged ? fstage : feat;
  const int fs = staged ? cs : c;                          // row stride of f
  __syncthreads();
  const int nr = s_nr;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nth >> 5;
  // squared norms: one warp per row
  for (int i = warp; i < N; i += nwarps) {
    float s = 0.f;
    for (int k = lane; k < c; k += 32) {
      const float v = f[(long long)i * fs + k];
      s = fmaf(v, v, s);
    }
    s = ud_warp_sum(s);
    if (lane == 0) sq[i] = s;
  }
  __syncthreads();
  // distances of the anchor rows: one THREAD per (i, j) pair
  for (int p = tid; p < nr * N; p += nth) {
    const int i = p / N, j = p - i * N;
    float s = 0.f;
    if (staged) {
      const float4* a = reinterpret_cast<const float4*>(fstage + i * cs);
      const float4* b = reinterpret_cast<const float4*>(fstage + j * cs);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int k = 0; k < cs / 4; ++k) {
        const float4 u = a[k], v = b[k];
        s0 = fmaf(u.x, v.x, s0); s1 = fmaf(u.y, v.y, s1); s2 = fmaf(u.z, v.z, s2); s3 = fmaf(u.w, v.w, s3);
      }
      s = (s0 + s1) + (s2 + s3);
    } else {
      for (int k = 0; k < c; ++k) s = fmaf(__ldg(feat + (long long)i * c + k), __ldg(feat + (long long)j * c + k), s);
    }
    const float q = sq[i] + sq[j] - 2.f * s;
    dist[p] = sqrtf(fmaxf(q, 1e-12f));
    Hm[p] = (q >= 1e-12f) ? 1.f : 0.f;  // clamp(min) passes gradient only where q >= min
  }
  __syncthreads();
  // per-anchor weighted sums: one warp per anchor
  const float eps = 1e-12f;
  float lsum = 0.f;
  for (int i = warp; i < nr; i += nwarps) {
    const int li = s_lab[i];
    float sp = 0.f, sn = 0.f, dp = 0.f, dn = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float d = dist[i * N + j];
      if (s_lab[j] == li) {
        if (j != i) {
          const float e = expf(d);
          sp += e;
          dp = fmaf(e, d, dp);
        }
      } else {
        const float e = expf(-d);
        sn += e;
        dn = fmaf(e, d, dn);
      }
    }
    sp = ud_warp_sum(sp); sn = ud_warp_sum(sn); dp = ud_warp_sum(dp); dn = ud_warp_sum(dn);
    const float fp = dp / (sp + eps), fn = dn / (sn + eps);
    const float u = fn - fp;
    // SoftMarginLoss(u, 1) = log(1 + exp(-u)); d/du = -sigmoid(-u)
    const float l = (u > 0.f) ? log1pf(expf(-u)) : (-u + log1pf(expf(u)));
    const float gui = -1.f / (1.f + expf(u)) / (float)nr;
    if (lane == 0) {
      lsum += l;
      fpv[i] = fp;
      fnv[i] = fn;
      gu[i] = gui;
    }
    // dL/dd_ij -> H_ij = (dL/dd_ij) / d_ij (masked by the clamp)
    for (int j = lane; j < N; j += 32) {
      const float d = dist[i * N + j];
      float g = 0.f;
      if (s_lab[j] == li) {
        if (j != i) g = -gui * (expf(d) / (sp + eps)) * (1.f + d - fp);
      } else {
        g = gui * (expf(-d) / (sn + eps)) * (1.f - d + fn);
      }
      Hm[i * N + j] = Hm[i * N + j] * g / d;
    }
  }
  lsum = ud_block_sum(lsum, red);  // also orders the Hm writes before the reads below
  if (tid == 0) *loss = (nr > 0) ? lsum / (float)nr : 0.f;
  if (gfeat == nullptr) return;
  // Hs = H + H^T (rows >= nr of H are zero) into `dist`, then its row sums into sq
  for (int p = tid; p < N * N; p += nth) {
    const int m = p / N, j = p - m * N;
    float h = 0.f;
    if (m < nr) h += Hm[m * N + j];
    if (j < nr) h += Hm[j * N + m];
    dist[p] = h;
  }
  __syncthreads();
  for (int m = warp; m < N; m += nwarps) {
    float rs = 0.f;
    for (int j = lane; j < N; j += 32) rs += dist[m * N + j];
    rs = ud_warp_sum(rs);
    if (lane == 0) sq[m] = rs;
  }
  __syncthreads();
  // gx_m = x_m * rowsum(Hs)_m - sum_j Hs_mj x_j : one thread per element (consecutive k: conflict-free, Hs broadcast)
  for (int e = tid; e < N * c; e += nth) {
    const int m = e / c, k = e - m * c;
    float acc = f[(long long)m * fs + k] * sq[m];
    const float* hs = dist + m * N;
    for (int j = 0; j < N; ++j) acc = fmaf(-hs[j], f[(long long)j * fs + k], acc);
    gfeat[e] = acc;
  }
}

extern "C" int ud_triplet_fwd(const float* feat, const long long* labels, float* loss, float* gfeat, int N, int c,
                              cudaStream_t stream) {
  UD_REQUIRE(N >= 1 && c >= 1, UD_ERR_INVALID, "triplet: bad shape N=%d c=%d", N, c);
  UD_REQUIRE(N <= LS_MAX_N, UD_ERR_UNSUPPORTED, "triplet: per-rank batch %d > %d unsupported", N, LS_MAX_N);
  UD_REQUIRE(feat && labels && loss, UD_ERR_INVALID, "triplet: null pointer");
  const size_t base = sizeof(float) * (2ull * N * N + 4ull * N);
  int cs = (c + 3) & ~3;                                  // multiple of 4 floats (float4 rows) ...
  while ((cs & 31) != 4 && (cs & 31) != 12 && (cs & 31) != 20 && (cs & 31) != 28) cs += 4;   // ... and = 4 (mod 8) banks-wise
  const size_t stage = sizeof(float) * (size_t)N * cs;
  const int staged = (ud_align_up(base, 16) + stage <= (200u << 10)) ? 1 : 0;
  const size_t smem = staged ? ud_align_up(base, 16) + stage : base;
  if (smem > (48u << 10))
    UD_CUDA(cudaFuncSetAttribute(ls_triplet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ls_triplet_kernel<<<1, LS_TRI_THREADS, smem, stream>>>(feat, labels, loss, gfeat, N, c, cs, staged);
  return ud_check_launch("triplet");
}

// ---------------------------------------------------------------- a10 factorization
#define FC_CHUNK 64   // features per CTA

// per-feature statistics, normalised copies, diagonal of c, partial Grams.   grid = ceil(F/FC_CHUNK), block 256
__global__ void __launch_bounds__(256)
ls_fac_stats_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ ah,
                    float* __restrict__ bh, float* __restrict__ ra, float* __restrict__ sa, float* __restrict__ cdiag,
                    float* __restrict__ part_on, float* __restrict__ part_dd, float* __restrict__ gram_part, int N,
                    int F, float eps) {
  extern __shared__ float smf[];
  __shared__ float red[33];
  float* tA = smf;                 // [N][FC_CHUNK]
  float* tB = tA + N * FC_CHUNK;   // [N][FC_CHUNK]
  const int f0 = blockIdx.x * FC_CHUNK;
  const int tid = threadIdx.x;
  float on = 0.f, dd = 0.f;
  if (tid < FC_CHUNK) {
    const int f = f0 + tid;
    if (f < F) {
      float ma = 0.f, mb = 0.f;
      for (int n = 0; n < N; ++n) {
        ma += a[(long long)n * F + f];
        mb += b[(long long)n * F + f];
      }
      ma /= (float)N;
      mb /= (float)N;
      float va = 0.f, vb = 0.f;
      for (int n = 0; n < N; ++n) {
        const float da = a[(long long)n * F + f] - ma, db = b[(long long)n * F + f] - mb;
        va = fmaf(da, da, va);
        vb = fmaf(db, db, vb);
      }
      const float stda = sqrtf(va / (float)(N - 1)), stdb = sqrtf(vb / (float)(N - 1));   // unbiased (calib_loss.py:20-21)
      const float r_a = 1.f / (stda + eps), r_b = 1.f / (stdb + eps);
      float cd = 0.f;
      for (int n = 0; n < N; ++n) {
        const float xa = (a[(long long)n * F + f] - ma) * r_a, xb = (b[(long long)n * F + f] - mb) * r_b;
        tA[n * FC_CHUNK + tid] = xa;
        tB[n * FC_CHUNK + tid] = xb;
        ah[(long long)n * F + f] = xa;
        bh[(long long)n * F + f] = xb;
        cd = fmaf(xa, xb, cd);
      }
      cd /= (float)N;
      ra[f] = r_a;
      sa[f] = stda;
      cdiag[f] = cd;
      on = (cd - 1.f) * (cd - 1.f);
      dd = cd * cd;
    } else {
      for (int n = 0; n < N; ++n) {
        tA[n * FC_CHUNK + tid] = 0.f;
        tB[n * FC_CHUNK + tid] = 0.f;
      }
    }
  }
  on = ud_block_sum(on, red);
  dd = ud_block_sum(dd, red);
  if (tid == 0) {
    part_on[blockIdx.x] = on;
    part_dd[blockIdx.x] = dd;
  }
  // partial Grams over this chunk: GA[m][n] = sum_f Ah[m,f] Ah[n,f], GB likewise
  float* gp = gram_part + (long long)blockIdx.x * 2 * N * N;
  for (int p = tid; p < N * N; p += blockDim.x) {
    const int m = p / N, n = p - m * N;
    float ga = 0.f, gb = 0.f;
#pragma unroll 8
    for (int k = 0; k < FC_CHUNK; ++k) {
      ga = fmaf(tA[m * FC_CHUNK + k], tA[n * FC_CHUNK + k], ga);
      gb = fmaf(tB[m * FC_CHUNK + k], tB[n * FC_CHUNK + k], gb);
    }
    gp[p] = ga;
    gp[N * N + p] = gb;
  }
}

// single CTA: sum the partials in fixed order, loss, and GB for the backward
__global__ void __launch_bounds__(256)
ls_fac_loss_kernel(const float* __restrict__ part_on, const float* __restrict__ part_dd,
                   const float* __restrict__ gram_part, float* __restrict__ GB, float* __restrict__ loss, int N, int F,
                   int nchunks, float off_w) {
  __shared__ float red[33];
  float dot = 0.f;
  for (int p = threadIdx.x; p < N * N; p += blockDim.x) {
    float ga = 0.f, gb = 0.f;
    for (int ch = 0; ch < nchunks; ++ch) {
      ga += gram_part[(long long)ch * 2 * N * N + p];
      gb += gram_part[(long long)ch * 2 * N * N + N * N + p];
    }
    GB[p] = gb;
    dot = fmaf(ga, gb, dot);
  }
  dot = ud_block_sum(dot, red);
  float on = 0.f, dd = 0.f;
  for (int i = threadIdx.x; i < nchunks; i += blockDim.x) {
    on += part_on[i];
    dd += part_dd[i];
  }
  on = ud_block_sum(on, red);
  dd = ud_block_sum(dd, red);
  if (threadIdx.x == 0) {
    const float sumsq = dot / ((float)N * (float)N);      // sum_ij c_ij^2 = <Ah Ah^T, Bh Bh^T> / N^2
    const float off = (F > 1) ? (sumsq - dd) / ((float)F * (float)(F - 1)) : 0.f;
    *loss = on / (float)F + off_w * off;
  }
}

// d loss / d emb_a (upstream 1).   grid = ceil(F/FC_CHUNK), block 256 (thread = (n-group, feature))
__global__ void __launch_bounds__(256)
ls_fac_grad_kernel(const float* __restrict__ a, const float* __restrict__ ah, const float* __restrict__ bh,
                   const float* __restrict__ ra, const float* __restrict__ sa, const float* __restrict__ cdiag,
                   const float* __restrict__ GB, float* __restrict__ ga, int N, int F, float off_w) {
  extern __shared__ float smf[];
  float* dA = smf;  // [N][FC_CHUNK]
  const int f0 = blockIdx.x * FC_CHUNK;
  const int fl = threadIdx.x % FC_CHUNK, ng = threadIdx.x / FC_CHUNK, ngs = blockDim.x / FC_CHUNK;
  const int f = f0 + fl;
  const float k_off = (F > 1) ? 2.f * off_w / ((float)F * (float)(F - 1)) : 0.f;
  const float k_on = 2.f / (float)F;
  const float invN = 1.f / (float)N;
  if (f < F) {
    const float cd = cdiag[f];
    for (int n = ng; n < N; n += ngs) {
      float t = 0.f;  // (GB Ah)[n, f]
      for (int m = 0; m < N; ++m) t = fmaf(GB[n * N + m], ah[(long long)m * F + f], t);
      const float bv = bh[(long long)n * F + f];
      dA[n * FC_CHUNK + fl] = invN * (k_off * (invN * t - cd * bv) + k_on * (cd - 1.f) * bv);
    }
  }
  __syncthreads();
  if (f < F && ng == 0) {
    // through Ah = (a - mean) * r, r = 1/(std + eps), std unbiased
    float m1 = 0.f, m2 = 0.f;
    for (int n = 0; n < N; ++n) {
      const float d = dA[n * FC_CHUNK + fl];
      m1 += d;
      m2 = fmaf(d, ah[(long long)n * F + f], m2);  // sum dA * t * r
    }
    m1 *= invN;
    const float r = ra[f], s = sa[f];
    // dL/da_n = r (dA_n - mean dA) - r^2 (sum_m dA_m t_m) t_n / ((N-1) s),  t = Ah / r
    const float k2 = (s > 0.f) ? (m2 / r) * r * r / ((float)(N - 1) * s) : 0.f;
    for (int n = 0; n < N; ++n) {
      const float t_n = ah[(long long)n * F + f] / r;
      ga[(long long)n * F + f] = r * (dA[n * FC_CHUNK + fl] - m1) - k2 * t_n;
    }
  }
}

extern "C" size_t ud_factorization_workspace_bytes(int N, int F) {
  const size_t nch = ud_cdiv(F, FC_CHUNK);
  return sizeof(float) * (2ull * N * F + 3ull * F + 2ull * nch + nch * 2ull * N * N + (size_t)N * N) + 256;
}

extern "C" int ud_factorization_fwd(const float* emb_a, const float* emb_b, float* loss, float* g_a, void* ws,
                                    size_t ws_bytes, int N, int F, float off_diag_weight, float eps,
                                    cudaStream_t stream) {
  UD_REQUIRE(N >= 2 && F >= 1, UD_ERR_INVALID, "factorization: needs N >= 2 (unbiased std), got N=%d F=%d", N, F);
  UD_REQUIRE(N <= LS_MAX_N, UD_ERR_UNSUPPORTED, "factorization: per-rank batch %d > %d unsupported", N, LS_MAX_N);
  UD_REQUIRE(emb_a && emb_b && loss && ws, UD_ERR_INVALID, "factorization: null pointer");
  UD_REQUIRE(ws_bytes >= ud_factorization_workspace_bytes(N, F), UD_ERR_WORKSPACE, "factorization: workspace too small");
  const int nch = ud_cdiv(F, FC_CHUNK);
  float* p = static_cast<float*>(ws);
  float* ah = p; p += (size_t)N * F;
  float* bh = p; p += (size_t)N * F;
  float* ra = p; p += F;
  float* sa = p; p += F;
  float* cdiag = p; p += F;
  float* part_on = p; p += nch;
  float* part_dd = p; p += nch;
  float* gram_part = p; p += (size_t)nch * 2 * N * N;
  float* GB = p;
  const size_t smem = sizeof(float) * 2ull * N * FC_CHUNK;
  if (smem > (48u << 10)) {
    UD_CUDA(cudaFuncSetAttribute(ls_fac_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  ls_fac_stats_kernel<<<nch, 256, smem, stream>>>(emb_a, emb_b, ah, bh, ra, sa, cdiag, part_on, part_dd, gram_part, N,
                                                  F, eps);
  int rc = ud_check_launch("fac_stats");
  if (rc != UD_OK) return rc;
  ls_fac_loss_kernel<<<1, 256, 0, stream>>>(part_on, part_dd, gram_part, GB, loss, N, F, nch, off_diag_weight);
  if ((rc = ud_check_launch("fac_loss")) != UD_OK) return rc;
  if (g_a != nullptr) {
    ls_fac_grad_kernel<<<nch, 256, sizeof(float) * (size_t)N * FC_CHUNK, stream>>>(emb_a, ah, bh, ra, sa, cdiag, GB,
                                                                                   g_a, N, F, off_diag_weight);
    rc = ud_check_launch("fac_grad");
  }
  return rc;
}

// ---------------------------------------------------------------- a11 mask KL
// loss = sum_n sum_m softmax(gt)_m (log_softmax(gt)_m - log_softmax(pred)_m) / N ; g_pred = (softmax(pred) - softmax(gt))/N
__global__ void __launch_bounds__(128)
ls_mask_kl_kernel(const float* __restrict__ pred, const float* __restrict__ gt, float* __restrict__ row_loss,
                  float* __restrict__ g_pred, int N, int M) {
  __shared__ float red[33];
  const int n = blockIdx.x;
  const float* p = pred + (long long)n * M;
  const float* q = gt + (long long)n * M;
  float mp = -INFINITY, mq = -INFINITY;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    mp = fmaxf(mp, p[i]);
    mq = fmaxf(mq, q[i]);
  }
  // block max via the sum helper on a warp-max: do a two-step reduce
  mp = ud_warp_max(mp);
  mq = ud_warp_max(mq);
  __shared__ float smax[2][4];
  if ((threadIdx.x & 31) == 0) {
    smax[0][threadIdx.x >> 5] = mp;
    smax[1][threadIdx.x >> 5] = mq;
  }
  __syncthreads();
  mp = fmaxf(fmaxf(smax[0][0], smax[0][1]), fmaxf(smax[0][2], smax[0][3]));
  mq = fmaxf(fmaxf(smax[1][0], smax[1][1]), fmaxf(smax[1][2], smax[1][3]));
  float sp = 0.f, sq = 0.f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    sp += expf(p[i] - mp);
    sq += expf(q[i] - mq);
  }
  sp = ud_block_sum(sp, red);
  sq = ud_block_sum(sq, red);
  const float lzp = mp + logf(sp), lzq = mq + logf(sq);
  float l = 0.f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const float lp = p[i] - lzp, lq = q[i] - lzq;
    const float eq = expf(lq);
    l = fmaf(eq, lq - lp, l);
    if (g_pred) g_pred[(long long)n * M + i] = (expf(lp) - eq) / (float)N;
  }
  l = ud_block_sum(l, red);
  if (threadIdx.x == 0) row_loss[n] = l / (float)N;
}

__global__ void ls_sum_kernel(const float* __restrict__ v, float* __restrict__ out, int n) {
  __shared__ float red[33];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += v[i];
  a = ud_block_sum(a, red);
  if (threadIdx.x == 0) *out = a;
}

extern "C" size_t ud_mask_kl_workspace_bytes(int N) { return sizeof(float) * (size_t)(N > 0 ? N : 1); }

extern "C" int ud_mask_kl_fwd(const float* pred, const float* gt, float* loss, float* g_pred, void* ws, size_t ws_bytes,
                              int N, int M, cudaStream_t stream) {
  UD_REQUIRE(N >= 1 && M >= 1, UD_ERR_INVALID, "mask_kl: bad shape N=%d M=%d", N, M);
  UD_REQUIRE(pred && gt && loss && ws, UD_ERR_INVALID, "mask_kl: null pointer");
  UD_REQUIRE(ws_bytes >= ud_mask_kl_workspace_bytes(N), UD_ERR_WORKSPACE, "mask_kl: workspace too small");
  float* rows = static_cast<float*>(ws);
  ls_mask_kl_kernel<<<N, 128, 0, stream>>>(pred, gt, rows, g_pred, N, M);
  int rc = ud_check_launch("mask_kl");
  if (rc != UD_OK) return rc;
  ls_sum_kernel<<<1, 128, 0, stream>>>(rows, loss, N);
  return ud_check_launch("mask_kl_sum");
}

// nn.KLDivLoss(reduction="batchmean", log_target=True) on already-normalised log-probabilities, the
// signature the engine calls (engine/abstract_engine.py:337,345): loss = sum exp(t)*(t - x) / N,
// d loss / d x = -exp(t) / N.
__global__ void __launch_bounds__(128)
ls_kl_log_kernel(const float* __restrict__ x, const float* __restrict__ t, float* __restrict__ row_loss,
                 float* __restrict__ g_x, int N, int M) {
  __shared__ float red[33];
  const int n = blockIdx.x;
  float l = 0.f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const float tv = t[(long long)n * M + i], xv = x[(long long)n * M + i];
    const float e = expf(tv);
    l = fmaf(e, tv - xv, l);
    if (g_x) g_x[(long long)n * M + i] = -e / (float)N;
  }
  l = ud_block_sum(l, red);
  if (threadIdx.x == 0) row_loss[n] = l / (float)N;
}

extern "C" int ud_kl_div_log_target_fwd(const float* log_pred, const float* log_target, float* loss, float* g_pred,
                                        void* ws, size_t ws_bytes, int N, int M, cudaStream_t stream) {
  UD_REQUIRE(N >= 1 && M >= 1, UD_ERR_INVALID, "kl_div: bad shape N=%d M=%d", N, M);
  UD_REQUIRE(log_pred && log_target && loss && ws, UD_ERR_INVALID, "kl_div: null pointer");
  UD_REQUIRE(ws_bytes >= ud_mask_kl_workspace_bytes(N), UD_ERR_WORKSPACE, "kl_div: workspace too small");
  float* rows = static_cast<float*>(ws);
  ls_kl_log_kernel<<<N, 128, 0, stream>>>(log_pred, log_target, rows, g_pred, N, M);
  int rc = ud_check_launch("kl_div");
  if (rc != UD_OK) return rc;
  ls_sum_kernel<<<1, 128, 0, stream>>>(rows, loss, N);
  return ud_check_launch("kl_div_sum");
}

// ---------------------------------------------------------------- a12 classification loss
// nn.CrossEntropyLoss() (mean over the batch) on logits [N,K] with int64 targets, and nn.BCEWithLogitsLoss() on
// logits [N] with float targets -- the two branches of engine/abstract_engine.py:256-259 / :325-328.  One CTA:
// value and d loss / d logits in the same pass (the head is [N,2] or [N]: latency, not bandwidth).
__global__ void __launch_bounds__(128)
ls_ce_kernel(const float* __restrict__ logits, const long long* __restrict__ target, float* __restrict__ loss,
             float* __restrict__ g, int N, int K) {
  __shared__ float red[33];
  float l = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float* z = logits + (long long)n * K;
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, z[k]);
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += expf(z[k] - m);
    const float lse = m + logf(s);
    const int t = (int)target[n];
    l += lse - z[t];
    if (g) {
      for (int k = 0; k < K; ++k) g[(long long)n * K + k] = (expf(z[k] - lse) - (k == t ? 1.f : 0.f)) / (float)N;
    }
  }
  l = ud_block_sum(l, red);
  if (threadIdx.x == 0) *loss = l / (float)N;
}

__global__ void __launch_bounds__(128)
ls_bce_logits_kernel(const float* __restrict__ logits, const float* __restrict__ target, float* __restrict__ loss,
                     float* __restrict__ g, int N) {
  __shared__ float red[33];
  float l = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float z = logits[n], t = target[n];
    l += fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z)));          // torch's stable form
    if (g) g[n] = (1.f / (1.f + expf(-z)) - t) / (float)N;
  }
  l = ud_block_sum(l, red);
  if (threadIdx.x == 0) *loss = l / (float)N;
}

extern "C" int ud_cross_entropy_fwd(const float* logits, const long long* target, float* loss, float* g_logits, int N,
                                    int K, cudaStream_t stream) {
  UD_REQUIRE(N >= 1 && K >= 1, UD_ERR_INVALID, "cross_entropy: bad shape N=%d K=%d", N, K);
  UD_REQUIRE(logits && target && loss, UD_ERR_INVALID, "cross_entropy: null pointer");
  ls_ce_kernel<<<1, 128, 0, stream>>>(logits, target, loss, g_logits, N, K);
  return ud_check_launch("cross_entropy");
}

extern "C" int ud_bce_with_logits_fwd(const float* logits, const float* target, float* loss, float* g_logits, int N,
                                      cudaStream_t stream) {
  UD_REQUIRE(N >= 1, UD_ERR_INVALID, "bce_with_logits: bad shape N=%d", N);
  UD_REQUIRE(logits && target && loss, UD_ERR_INVALID, "bce_with_logits: null pointer");
  ls_bce_logits_kernel<<<1, 128, 0, stream>>>(logits, target, loss, g_logits, N);
  return ud_check_launch("bce_with_logits");
}

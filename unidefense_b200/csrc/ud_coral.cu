// a15 -- coral colour-statistics transfer, batched and fused (no grad)            
//
// Reference: utils/operation.py:15-45, called per sample in a Python loop (model/unidefense.py:189-191):
//     f  = (img - mean_c) / std_c            (per channel, UNBIASED std), cov = f f^T + I        (3x3, not / HW)
//     out = sqrt~(cov_t) @ inv(sqrt~(cov_s)) @ f_s * std_t + mean_t
// where sqrt~(M) = U diag(sqrt(D)) Vh^T with (U, D, Vh) = torch.linalg.svd(M)  -- Vh transposed AGAIN, i.e.
// U sqrt(D) V with V's COLUMNS the right singular vectors: not a matrix square root.
// For the symmetric positive definite M = cov + I, V = U, so sqrt~(M) = U sqrt(D) U.  That expression depends on the
// SIGN of every singular vector, which no SVD defines: LAPACK (the reference on CPU) and cuSOLVER (the reference on
// GPU) already disagree with each other.  Parity is therefore stated modulo that gauge (tests/test_perturb_gpu.py
// asserts that the reference fixture is one of the 2^3 x 2^3 sign patterns and that this kernel is the pattern of
// its documented convention): each eigenvector is signed so that its largest-magnitude component is positive,
// eigenvalues in descending order.
//
// Three launches for the whole batch instead of 2N host-synchronising 3x3 SVDs:
//   coral_moments : per (image, pixel chunk) shifted raw moments  sum (x-k)_c, sum (x-k)_c (x-k)_d      (HBM-bound)
//   coral_solve   : one warp per sample: reduce the chunks, 3x3 algebra in fp64 (cyclic Jacobi), emit the affine
//                   colour map  out_c = sum_d A[c][d] s_d + b[c]
//   coral_apply   : the per-pixel 3x3 map                                                               (HBM-bound)
// Algorithmic bytes: read source twice + target once, write out once = 4 * 3*H*W*4 per sample.
#include "../../include/unidefense_b200.h"
#include "ud_common.cuh"

#define CR_CHUNK 4096
#define CR_THREADS 256

// grid (chunks, 2N): image index < N: source, else target.  part[(img*chunks + chunk)*9 + {s0,s1,s2,q00,q01,q02,q11,q12,q22}]
__global__ void __launch_bounds__(CR_THREADS)
coral_moments_kernel(const float* __restrict__ src, const float* __restrict__ tgt, float* __restrict__ part, int N, int HW,
                     int chunks) {
  __shared__ float red[9][CR_THREADS / 32];
  const int img = blockIdx.y;
  const float* p = (img < N ? src + (long long)img * 3 * HW : tgt + (long long)(img - N) * 3 * HW);
  const float k0 = __ldg(p), k1 = __ldg(p + HW), k2 = __ldg(p + 2 * (long long)HW);      // pivot: the first pixel
  const int i0 = blockIdx.x * CR_CHUNK;
  const int i1 = min(i0 + CR_CHUNK, HW);
  float a[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = i0 + threadIdx.x; i < i1; i += CR_THREADS) {
    const float x = __ldg(p + i) - k0, y = __ldg(p + HW + i) - k1, z = __ldg(p + 2 * (long long)HW + i) - k2;
    a[0] += x; a[1] += y; a[2] += z;
    a[3] = fmaf(x, x, a[3]); a[4] = fmaf(x, y, a[4]); a[5] = fmaf(x, z, a[5]);
    a[6] = fmaf(y, y, a[6]); a[7] = fmaf(y, z, a[7]); a[8] = fmaf(z, z, a[8]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const float v = ud_warp_sum(a[j]);
    if (lane == 0) red[j][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    float v = 0.f;
    for (int w = 0; w < CR_THREADS / 32; ++w) v += red[threadIdx.x][w];
    part[((long long)img * chunks + blockIdx.x) * 9 + threadIdx.x] = v;
  }
}

// cyclic Jacobi on a symmetric 3x3 (fp64): a -> diag(d), columns of v = eigenvectors
__device__ void cr_jacobi3(double a[3][3], double d[3], double v[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p][q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {            // A <- J^T A J
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) d[i] = a[i][i];
  // descending eigenvalues (singular-value order), columns follow
  for (int i = 0; i < 2; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (d[j] > d[i]) {
        const double td = d[i]; d[i] = d[j]; d[j] = td;
        for (int k = 0; k < 3; ++k) { const double tv = v[k][i]; v[k][i] = v[k][j]; v[k][j] = tv; }
      }
  // gauge: largest-magnitude component of every eigenvector positive
  for (int j = 0; j < 3; ++j) {
    int m = 0;
    for (int k = 1; k < 3; ++k) if (fabs(v[k][j]) > fabs(v[m][j])) m = k;
    if (v[m][j] < 0.0) for (int k = 0; k < 3; ++k) v[k][j] = -v[k][j];
  }
}

// R = U sqrt(D) U   (the reference's U sqrt(D) Vh^T for symmetric positive definite input; NOT U sqrt(D) U^T)
__device__ void cr_quirk_sqrt(const double cov[3][3], double R[3][3]) {
  double a[3][3], d[3], u[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = cov[i][j];
  cr_jacobi3(a, d, u);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += u[i][k] * sqrt(fmax(d[k], 0.0)) * u[k][j];
      R[i][j] = s;
    }
}

// mean, unbiased std and f f^T + I of one image from the summed shifted moments
__device__ void cr_stats(const double m[9], const float piv[3], int HW, double mean[3], double sd[3], double cov[3][3]) {
  const double n = (double)HW;
  double mu[3] = {m[0] / n, m[1] / n, m[2] / n};
  const double q[3][3] = {{m[3], m[4], m[5]}, {m[4], m[6], m[7]}, {m[5], m[7], m[8]}};
  double c2[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) c2[i][j] = q[i][j] - n * mu[i] * mu[j];   // sum (x-mean)_i (x-mean)_j
  for (int i = 0; i < 3; ++i) {
    mean[i] = mu[i] + (double)piv[i];
    sd[i] = sqrt(fmax(c2[i][i], 0.0) / (n - 1.0));
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov[i][j] = c2[i][j] / (sd[i] * sd[j]) + (i == j ? 1.0 : 0.0);
}

// one warp per sample; coef[n*12 + {A row-major 9, b 3}]
__global__ void __launch_bounds__(32)
coral_solve_kernel(const float* __restrict__ src, const float* __restrict__ tgt, const float* __restrict__ part,
                   float* __restrict__ coef, int N, int HW, int chunks) {
  const int n = blockIdx.x, lane = threadIdx.x;
  double ms[9], mt[9];
  for (int j = 0; j < 9; ++j) {
    double a = 0.0, b = 0.0;
    for (int c = lane; c < chunks; c += 32) {
      a += (double)part[((long long)n * chunks + c) * 9 + j];
      b += (double)part[((long long)(N + n) * chunks + c) * 9 + j];
    }
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    ms[j] = a;
    mt[j] = b;
  }
  if (lane != 0) return;
  const float* ps = src + (long long)n * 3 * HW;
  const float* pt = tgt + (long long)n * 3 * HW;
  const float pivs[3] = {ps[0], ps[HW], ps[2 * (long long)HW]};
  const float pivt[3] = {pt[0], pt[HW], pt[2 * (long long)HW]};
  double mean_s[3], sd_s[3], cov_s[3][3], mean_t[3], sd_t[3], cov_t[3][3];
  cr_stats(ms, pivs, HW, mean_s, sd_s, cov_s);
  cr_stats(mt, pivt, HW, mean_t, sd_t, cov_t);
  double Rs[3][3], Rt[3][3];
  cr_quirk_sqrt(cov_s, Rs);
  cr_quirk_sqrt(cov_t, Rt);
  // inverse of Rs by the adjugate
  const double det = Rs[0][0] * (Rs[1][1] * Rs[2][2] - Rs[1][2] * Rs[2][1]) - Rs[0][1] * (Rs[1][0] * Rs[2][2] - Rs[1][2] * Rs[2][0]) +
                     Rs[0][2] * (Rs[1][0] * Rs[2][1] - Rs[1][1] * Rs[2][0]);
  double inv[3][3];
  inv[0][0] = (Rs[1][1] * Rs[2][2] - Rs[1][2] * Rs[2][1]) / det;
  inv[0][1] = (Rs[0][2] * Rs[2][1] - Rs[0][1] * Rs[2][2]) / det;
  inv[0][2] = (Rs[0][1] * Rs[1][2] - Rs[0][2] * Rs[1][1]) / det;
  inv[1][0] = (Rs[1][2] * Rs[2][0] - Rs[1][0] * Rs[2][2]) / det;
  inv[1][1] = (Rs[0][0] * Rs[2][2] - Rs[0][2] * Rs[2][0]) / det;
  inv[1][2] = (Rs[0][2] * Rs[1][0] - Rs[0][0] * Rs[1][2]) / det;
  inv[2][0] = (Rs[1][0] * Rs[2][1] - Rs[1][1] * Rs[2][0]) / det;
  inv[2][1] = (Rs[0][1] * Rs[2][0] - Rs[0][0] * Rs[2][1]) / det;
  inv[2][2] = (Rs[0][0] * Rs[1][1] - Rs[0][1] * Rs[1][0]) / det;
  // out = (Rt inv(Rs)) ((s - mean_s)/sd_s) * sd_t + mean_t  =  A s + b
  for (int i = 0; i < 3; ++i) {
    double bi = mean_t[i];
    for (int j = 0; j < 3; ++j) {
      double mij = 0.0;
      for (int k = 0; k < 3; ++k) mij += Rt[i][k] * inv[k][j];
      const double aij = sd_t[i] * mij / sd_s[j];
      coef[n * 12 + i * 3 + j] = (float)aij;
      bi -= aij * mean_s[j];
    }
    coef[n * 12 + 9 + i] = (float)bi;
  }
}

__global__ void __launch_bounds__(256)
coral_apply_kernel(const float* __restrict__ src, const float* __restrict__ coef, float* __restrict__ out, int HW) {
  const int n = blockIdx.y;
  __shared__ float c[12];
  if (threadIdx.x < 12) c[threadIdx.x] = coef[n * 12 + threadIdx.x];
  __syncthreads();
  const float* p = src + (long long)n * 3 * HW;
  float* o = out + (long long)n * 3 * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float x = __ldcs(p + i), y = __ldcs(p + HW + i), z = __ldcs(p + 2 * (long long)HW + i);
    __stcs(o + i, fmaf(c[0], x, fmaf(c[1], y, fmaf(c[2], z, c[9]))));
    __stcs(o + HW + i, fmaf(c[3], x, fmaf(c[4], y, fmaf(c[5], z, c[10]))));
    __stcs(o + 2 * (long long)HW + i, fmaf(c[6], x, fmaf(c[7], y, fmaf(c[8], z, c[11]))));
  }
}

extern "C" size_t ud_coral_workspace_bytes(int N, int HW) {
  if (N < 1 || HW < 1) return 0;
  const size_t chunks = (size_t)ud_cdiv(HW, CR_CHUNK);
  return (2ull * N * chunks * 9 + 12ull * N) * sizeof(float) + 256;
}

// source, target, out: [N, 3, H, W] fp32 (source[n] takes the colour statistics of target[n]); HW = H*W >= 2
extern "C" int ud_coral(const float* source, const float* target, float* out, void* ws, size_t ws_bytes, int N, int HW,
                        cudaStream_t stream) {
  UD_REQUIRE(N >= 0 && HW >= 2, UD_ERR_INVALID, "coral: bad shape N=%d HW=%d (unbiased std needs HW >= 2)", N, HW);
  if (N == 0) return UD_OK;
  UD_REQUIRE(source && target && out && ws, UD_ERR_INVALID, "coral: null pointer");
  UD_REQUIRE(ws_bytes >= ud_coral_workspace_bytes(N, HW), UD_ERR_WORKSPACE, "coral: workspace too small");
  UD_REQUIRE(2 * N <= 65535, UD_ERR_UNSUPPORTED, "coral: batch too large");
  const int chunks = ud_cdiv(HW, CR_CHUNK);
  float* part = static_cast<float*>(ws);
  float* coef = part + 2ull * N * chunks * 9;
  coral_moments_kernel<<<dim3(chunks, 2 * N), CR_THREADS, 0, stream>>>(source, target, part, N, HW, chunks);
  int rc = ud_check_launch("coral_moments");
  if (rc != UD_OK) return rc;
  coral_solve_kernel<<<N, 32, 0, stream>>>(source, target, part, coef, N, HW, chunks);
  if ((rc = ud_check_launch("coral_solve")) != UD_OK) return rc;
  const int gx = ud_cdiv(HW, 256 * 4) < 1 ? 1 : ud_cdiv(HW, 256 * 4);
  coral_apply_kernel<<<dim3(gx, N), 256, 0, stream>>>(source, coef, out, HW);
  return ud_check_launch("coral_apply");
}

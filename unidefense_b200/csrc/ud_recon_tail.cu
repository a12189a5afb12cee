// a1 -- reconstruction-loss tail, forward and backward        
//
// Reference: model/unidefense.py:244-253 (Eb4), :423-433 (Res18), :618-628 (Res50):
//     rec     = bilinear(dec -> HxW, align_corners=True)
//     spatial = mean_{c,h,w} |rec - x|
//     freq    = mean_{c,h,k} |Re dF| + |Im dF|,   dF = rfft2(rec) - rfft2(x)
// By linearity dF = rfft2(rec - x), so ONE 2-D real FFT of d = rec - x per
// plane is computed instead of two, and neither spectrum is ever written to HBM.
//
// Forward  = rows kernel (upsample + diff + |d| partial sums + packed real row FFTs -> Y)
//          + cols kernel (column FFTs of Y + |Re|+|Im| partial sums + sign bytes)
//          + finalize   (fixed-order sum of the partials -> spatial[n], freq[n]).
// Backward = cols-inverse kernel (sign spectrum -> inverse column FFTs -> T)
//          + rows-inverse kernel (Hermitian-folded packed inverse row FFTs + sign(d) term
//            + transposed bilinear into g_dec).
// Y/T is an L2-resident workspace holding `chunk` samples' half spectra [plane][H][WhP].
//
// HBM traffic per sample (fp32): fwd reads dec+x, writes rec (+1 B/bin signs);
// bwd reads dec+x(+signs), writes g_dec.
#include "../../include/unidefense_b200.h"
#include "ud_fft.cuh"

#define RT_THREADS 256
#define RT_LINES 12   // row pairs per CTA (rows kernels) / columns per CTA (cols kernels)

static inline int rt_whp(int W) { return ((W / 2 + 1) + 3) & ~3; }  // padded half width (float2)

// ------------------------------------------------------------------------------------------
// forward rows: grid (row_tiles, planes_in_chunk)
// ------------------------------------------------------------------------------------------
template <class Plan>
__global__ void __launch_bounds__(RT_THREADS)
rt_rows_fwd_kernel(Plan plan, const float* __restrict__ dec, const float* __restrict__ x,
                   float* __restrict__ rec, float2* __restrict__ Y, float* __restrict__ part_spatial,
                   const float2* __restrict__ tw_g, int plane0, int h, int w, int H, int W, float sy, float sx,
                   int row_tiles) {
  extern __shared__ float2 smem[];
  const int n = plan.n();  // == W
  const int LS = n | 1;
  float2* tw = smem;
  float2* buf0 = tw + n;
  float2* buf1 = buf0 + RT_LINES * LS;
  __shared__ float red[33];

  const int tile = blockIdx.x;
  const int pl = blockIdx.y;             // plane within chunk
  const long long plane = plane0 + pl;   // global (n*C + c)
  const int r0 = tile * (2 * RT_LINES);
  const int Wh = W / 2 + 1;
  const int WhP = ((Wh + 3) & ~3);

  for (int t = threadIdx.x; t < n; t += blockDim.x) tw[t] = tw_g[t];

  const float* decp = dec + plane * (long long)h * w;
  const float* xp = x + plane * (long long)H * W;
  float* recp = rec + plane * (long long)H * W;

  float acc = 0.f;
  for (int p = 0; p < RT_LINES; ++p) {
    const int ra = r0 + 2 * p, rb = ra + 1;
    float2* line = buf0 + p * LS;
    if (ra >= H) {
      for (int c = threadIdx.x; c < W; c += blockDim.x) line[c] = make_float2(0.f, 0.f);
      continue;
    }
    const UdLerp ya = ud_lerp_ac(ra, h, sy);
    const bool has_b = rb < H;
    const UdLerp yb = ud_lerp_ac(has_b ? rb : ra, h, sy);
    const float* da0 = decp + ya.i0 * w;
    const float* da1 = decp + ya.i1 * w;
    const float* db0 = decp + yb.i0 * w;
    const float* db1 = decp + yb.i1 * w;
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
      const UdLerp xc = ud_lerp_ac(c, w, sx);
      const float va = ya.l0 * (xc.l0 * __ldg(da0 + xc.i0) + xc.l1 * __ldg(da0 + xc.i1)) +
                       ya.l1 * (xc.l0 * __ldg(da1 + xc.i0) + xc.l1 * __ldg(da1 + xc.i1));
      const float xa = __ldg(xp + (long long)ra * W + c);
      recp[(long long)ra * W + c] = va;
      const float dA = va - xa;
      float dB = 0.f;
      if (has_b) {
        const float vb = yb.l0 * (xc.l0 * __ldg(db0 + xc.i0) + xc.l1 * __ldg(db0 + xc.i1)) +
                         yb.l1 * (xc.l0 * __ldg(db1 + xc.i0) + xc.l1 * __ldg(db1 + xc.i1));
        const float xb = __ldg(xp + (long long)rb * W + c);
        recp[(long long)rb * W + c] = vb;
        dB = vb - xb;
      }
      acc += fabsf(dA) + fabsf(dB);
      line[c] = make_float2(dA, dB);
    }
  }
  __syncthreads();
  float2* res = plan.run(buf0, buf1, tw, RT_LINES, LS);

  // unpack the two real rows of each pair: A[k] = (Z[k]+conj Z[n-k])/2, B[k] = (Z[k]-conj Z[n-k])/(2i)
  float2* Yp = Y + (long long)pl * H * WhP;
  for (int t = threadIdx.x; t < RT_LINES * Wh; t += blockDim.x) {
    const int p = t / Wh, k = t - p * Wh;
    const int ra = r0 + 2 * p;
    if (ra >= H) continue;
    const float2 z = res[p * LS + k];
    const float2 zn = res[p * LS + (k == 0 ? 0 : n - k)];
    Yp[(long long)ra * WhP + k] = make_float2(0.5f * (z.x + zn.x), 0.5f * (z.y - zn.y));
    if (ra + 1 < H) Yp[(long long)(ra + 1) * WhP + k] = make_float2(0.5f * (z.y + zn.y), -0.5f * (z.x - zn.x));
  }
  const float tot = ud_block_sum(acc, red);
  if (threadIdx.x == 0) part_spatial[plane * row_tiles + tile] = tot;
}

// ------------------------------------------------------------------------------------------
// forward cols: grid (col_tiles, planes_in_chunk).  Column FFT length n = H.
// ------------------------------------------------------------------------------------------
template <class Plan>
__global__ void __launch_bounds__(RT_THREADS)
rt_cols_fwd_kernel(Plan plan, const float2* __restrict__ Y, float* __restrict__ part_freq,
                   uint8_t* __restrict__ signs, const float2* __restrict__ tw_g, int plane0, int H, int W,
                   int col_tiles) {
  extern __shared__ float2 smem[];
  const int n = plan.n();  // == H
  const int LS = n | 1;
  float2* tw = smem;
  float2* buf0 = tw + n;
  float2* buf1 = buf0 + RT_LINES * LS;
  __shared__ float red[33];

  const int tile = blockIdx.x;
  const int pl = blockIdx.y;
  const long long plane = plane0 + pl;
  const int Wh = W / 2 + 1;
  const int WhP = ((Wh + 3) & ~3);
  const int k0 = tile * RT_LINES;
  const int ncols = min(RT_LINES, Wh - k0);

  for (int t = threadIdx.x; t < n; t += blockDim.x) tw[t] = tw_g[t];
  const float2* Yp = Y + (long long)pl * H * WhP;
  for (int t = threadIdx.x; t < H * RT_LINES; t += blockDim.x) {
    const int r = t / RT_LINES, cc = t - r * RT_LINES;
    float2 v = make_float2(0.f, 0.f);
    if (cc < ncols) v = Yp[(long long)r * WhP + k0 + cc];
    buf0[cc * LS + r] = v;
  }
  __syncthreads();
  float2* res = plan.run(buf0, buf1, tw, RT_LINES, LS);

  float acc = 0.f;
  for (int t = threadIdx.x; t < ncols * H; t += blockDim.x) {
    const int cc = t / H, j = t - cc * H;
    const float2 v = res[cc * LS + j];
    acc += fabsf(v.x) + fabsf(v.y);
    if (signs != nullptr) {
      const uint8_t sr = v.x > 0.f ? 1 : (v.x < 0.f ? 2 : 0);
      const uint8_t si = v.y > 0.f ? 1 : (v.y < 0.f ? 2 : 0);
      signs[(plane * Wh + (k0 + cc)) * (long long)H + j] = (uint8_t)(sr | (si << 2));
    }
  }
  const float tot = ud_block_sum(acc, red);
  if (threadIdx.x == 0) part_freq[plane * col_tiles + tile] = tot;
}

// fixed-order reduction of the partial sums: grid N, block 128
__global__ void rt_finalize_kernel(const float* __restrict__ part_spatial, const float* __restrict__ part_freq,
                                   float* __restrict__ spatial, float* __restrict__ freq, int C, int row_tiles,
                                   int col_tiles, float inv_sp, float inv_fr) {
  __shared__ float red[33];
  const int nsmp = blockIdx.x;
  float a = 0.f, b = 0.f;
  const int ns = C * row_tiles, nf = C * col_tiles;
  for (int i = threadIdx.x; i < ns; i += blockDim.x) a += part_spatial[(long long)nsmp * ns + i];
  for (int i = threadIdx.x; i < nf; i += blockDim.x) b += part_freq[(long long)nsmp * nf + i];
  a = ud_block_sum(a, red);
  b = ud_block_sum(b, red);
  if (threadIdx.x == 0) {
    spatial[nsmp] = a * inv_sp;
    freq[nsmp] = b * inv_fr;
  }
}

// ------------------------------------------------------------------------------------------
// backward cols (inverse column FFT of the sign spectrum): grid (col_tiles, planes_in_chunk)
// T[r,k] = sum_j S[j,k] e^{+2 pi i j r / H},  S = gf * (sgn Re D + i sgn Im D)
// inverse via the swap trick: IFFT(z) = swap(FFT(swap(z))).
// ------------------------------------------------------------------------------------------
template <class Plan>
__global__ void __launch_bounds__(RT_THREADS)
rt_cols_bwd_kernel(Plan plan, const uint8_t* __restrict__ signs, const float* __restrict__ g_freq,
                   float2* __restrict__ T, const float2* __restrict__ tw_g, int plane0, int C, int H, int W,
                   float gscale) {
  extern __shared__ float2 smem[];
  const int n = plan.n();
  const int LS = n | 1;
  float2* tw = smem;
  float2* buf0 = tw + n;
  float2* buf1 = buf0 + RT_LINES * LS;

  const int tile = blockIdx.x;
  const int pl = blockIdx.y;
  const long long plane = plane0 + pl;
  const int smp = (int)(plane / C);
  if (g_freq[smp] == 0.f) return;  // whole plane contributes nothing; the rows kernel skips its FFT too
  const float gf = g_freq[smp] * gscale;
  const int Wh = W / 2 + 1;
  const int WhP = ((Wh + 3) & ~3);
  const int k0 = tile * RT_LINES;
  const int ncols = min(RT_LINES, Wh - k0);

  for (int t = threadIdx.x; t < n; t += blockDim.x) tw[t] = tw_g[t];
  for (int t = threadIdx.x; t < RT_LINES * H; t += blockDim.x) {
    const int cc = t / H, j = t - cc * H;
    float2 v = make_float2(0.f, 0.f);
    if (cc < ncols) {
      const uint8_t s = signs[(plane * Wh + (k0 + cc)) * (long long)H + j];
      const float sr = (s & 1) ? gf : ((s & 2) ? -gf : 0.f);
      const float si = (s & 4) ? gf : ((s & 8) ? -gf : 0.f);
      v = make_float2(si, sr);  // swapped
    }
    buf0[cc * LS + j] = v;
  }
  __syncthreads();
  float2* res = plan.run(buf0, buf1, tw, RT_LINES, LS);
  float2* Tp = T + (long long)pl * H * WhP;
  for (int t = threadIdx.x; t < H * RT_LINES; t += blockDim.x) {
    const int r = t / RT_LINES, cc = t - r * RT_LINES;
    if (cc < ncols) {
      const float2 v = res[cc * LS + r];
      Tp[(long long)r * WhP + k0 + cc] = make_float2(v.y, v.x);  // swap back
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward rows: grid (row_tiles, planes_in_chunk).
// g_rec[r,c] = gs*sign(d[r,c]) + Re sum_{k<Wh} T[r,k] e^{+2 pi i k c / W}; then the transposed
// bilinear resize accumulates into g_dec (pre-zeroed).  Two rows are packed per complex line by
// folding T into its Hermitian part Th (Re IFFT(T) == IFFT(Th)) and transforming Th_a + i Th_b.
// ------------------------------------------------------------------------------------------
template <class Plan>
__global__ void __launch_bounds__(RT_THREADS)
rt_rows_bwd_kernel(Plan plan, const float* __restrict__ dec, const float* __restrict__ x,
                   const float2* __restrict__ T, const float* __restrict__ g_spatial,
                   const float* __restrict__ g_freq, float* __restrict__ g_dec,
                   const float2* __restrict__ tw_g, int plane0, int C, int h, int w, int H, int W, float sy,
                   float sx, float sp_scale) {
  extern __shared__ float2 smem[];
  const int n = plan.n();  // == W
  const int LS = n | 1;
  float2* tw = smem;
  float2* buf0 = tw + n;
  float2* buf1 = buf0 + RT_LINES * LS;
  float* hrow = reinterpret_cast<float*>(buf1 + RT_LINES * LS);  // [2*RT_LINES][w]

  const int tile = blockIdx.x;
  const int pl = blockIdx.y;
  const long long plane = plane0 + pl;
  const int smp = (int)(plane / C);
  const float gs = g_spatial[smp] * sp_scale;
  const bool has_f = g_freq[smp] != 0.f;
  if (gs == 0.f && !has_f) return;
  const int r0 = tile * (2 * RT_LINES);
  const int Wh = W / 2 + 1;
  const int WhP = ((Wh + 3) & ~3);
  const int nrows = min(2 * RT_LINES, H - r0);

  float2* res = buf0;
  if (has_f) {
    for (int t = threadIdx.x; t < n; t += blockDim.x) tw[t] = tw_g[t];
    const float2* Tp = T + (long long)pl * H * WhP;
    // build V = Th_a + i Th_b (swapped for the inverse) for k in [0, Wh) and its mirror n-k
    for (int t = threadIdx.x; t < RT_LINES * Wh; t += blockDim.x) {
      const int p = t / Wh, k = t - p * Wh;
      const int ra = r0 + 2 * p;
      float2 ta = make_float2(0.f, 0.f), tb = make_float2(0.f, 0.f);
      if (ra < H) ta = Tp[(long long)ra * WhP + k];
      if (ra + 1 < H) tb = Tp[(long long)(ra + 1) * WhP + k];
      float2* line = buf0 + p * LS;
      const bool self_paired = (k == 0) || (2 * k == n);
      if (self_paired) {
        // Th = Re T ; V = (Re Ta, Re Tb) ; swapped -> (Re Tb, Re Ta)
        line[k] = make_float2(tb.x, ta.x);
      } else {
        const float ar = 0.5f * ta.x, ai = 0.5f * ta.y, br = 0.5f * tb.x, bi = 0.5f * tb.y;
        // V[k] = (ar - bi, ai + br) ; V[n-k] = (ar + bi, br - ai) ; store swapped
        line[k] = make_float2(ai + br, ar - bi);
        line[n - k] = make_float2(br - ai, ar + bi);
      }
    }
    __syncthreads();
    res = plan.run(buf0, buf1, tw, RT_LINES, LS);
    // res[p][c] (swapped back) = (g_a[c], g_b[c])  ->  read as (.y, .x)
  }

  // pass 1: per output row, g = gs*sign(d) + g_fft ; horizontal transposed lerp by gather
  const float* decp = dec + plane * (long long)h * w;
  const float* xp = x + plane * (long long)H * W;
  // write g rows into the free ping-pong buffer as plain floats [row][W]
  float* grow = reinterpret_cast<float*>(res == buf0 ? buf1 : buf0);
  for (int t = threadIdx.x; t < nrows * W; t += blockDim.x) {
    const int rr = t / W, c = t - rr * W;
    const int r = r0 + rr;
    float g = 0.f;
    if (has_f) {
      const float2 v = res[(rr >> 1) * LS + c];
      g = (rr & 1) ? v.x : v.y;
    }
    if (gs != 0.f) {
      const UdLerp yr = ud_lerp_ac(r, h, sy);
      const UdLerp xc = ud_lerp_ac(c, w, sx);
      const float* d0 = decp + yr.i0 * w;
      const float* d1 = decp + yr.i1 * w;
      const float v = yr.l0 * (xc.l0 * __ldg(d0 + xc.i0) + xc.l1 * __ldg(d0 + xc.i1)) +
                      yr.l1 * (xc.l0 * __ldg(d1 + xc.i0) + xc.l1 * __ldg(d1 + xc.i1));
      g += gs * ud_sign(v - __ldg(xp + (long long)r * W + c));
    }
    grow[rr * W + c] = g;
  }
  __syncthreads();
  const float inv_sx = sx > 0.f ? 1.f / sx : 0.f;
  for (int t = threadIdx.x; t < nrows * w; t += blockDim.x) {
    const int rr = t / w, j = t - rr * w;
    int c_lo, c_hi;
    if (sx > 0.f) {
      c_lo = max(0, (int)floorf((float)(j - 1) * inv_sx) - 1);
      c_hi = min(W - 1, (int)ceilf((float)(j + 1) * inv_sx) + 1);
    } else {
      c_lo = 0;
      c_hi = W - 1;
    }
    float a = 0.f;
    for (int c = c_lo; c <= c_hi; ++c) {
      const UdLerp xc = ud_lerp_ac(c, w, sx);
      float wgt = 0.f;
      if (xc.i0 == j) wgt += xc.l0;
      if (xc.i1 == j) wgt += xc.l1;
      a = fmaf(wgt, grow[rr * W + c], a);
    }
    hrow[rr * w + j] = a;
  }
  __syncthreads();
  // pass 2: vertical transposed lerp; this tile's rows touch dec rows [i_lo, i_hi]
  const int i_lo = ud_lerp_ac(r0, h, sy).i0;
  const int i_hi = ud_lerp_ac(r0 + nrows - 1, h, sy).i1;
  float* gd = g_dec + plane * (long long)h * w;
  for (int t = threadIdx.x; t < (i_hi - i_lo + 1) * w; t += blockDim.x) {
    const int ii = t / w, j = t - ii * w;
    const int i = i_lo + ii;
    float a = 0.f;
    for (int rr = 0; rr < nrows; ++rr) {
      const UdLerp yr = ud_lerp_ac(r0 + rr, h, sy);
      float wgt = 0.f;
      if (yr.i0 == i) wgt += yr.l0;
      if (yr.i1 == i) wgt += yr.l1;
      if (wgt != 0.f) a = fmaf(wgt, hrow[rr * w + j], a);
    }
    if (a != 0.f) atomicAdd(gd + (long long)i * w + j, a);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static size_t rt_fft_smem(int n) { return sizeof(float2) * ((size_t)n + 2ull * RT_LINES * (n | 1)); }

static int rt_chunk_samples(int N, int C, int H, int W) {
  // keep the Y/T workspace (plus the streamed images) comfortably inside the 126 MB L2
  const size_t per_sample = (size_t)C * H * rt_whp(W) * sizeof(float2);
  int chunk = (int)((24ull << 20) / (per_sample ? per_sample : 1));
  if (chunk < 1) chunk = 1;
  if (chunk > N) chunk = N;
  return chunk;
}

extern "C" size_t ud_recon_tail_workspace_bytes(int N, int C, int h, int w, int H, int W) {
  (void)h; (void)w;
  const int chunk = rt_chunk_samples(N, C, H, W);
  const int row_tiles = ud_cdiv(H, 2 * RT_LINES), col_tiles = ud_cdiv(W / 2 + 1, RT_LINES);
  size_t y = ud_align_up((size_t)chunk * C * H * rt_whp(W) * sizeof(float2), 256);
  size_t ps = ud_align_up((size_t)N * C * row_tiles * sizeof(float), 256);
  size_t pf = ud_align_up((size_t)N * C * col_tiles * sizeof(float), 256);
  return y + ps + pf;
}

extern "C" size_t ud_recon_tail_signs_bytes(int N, int C, int H, int W) {
  return (size_t)N * C * (W / 2 + 1) * H;
}

// Runs the body with PLANVAR bound to the static plan for n when one exists, else a dynamic plan.
#define RT_DISPATCH_PLAN(n, PLANVAR, ...)                                          \
  do {                                                                             \
    if ((n) == 380) { UdPlan380 PLANVAR; __VA_ARGS__; }                            \
    else if ((n) == 256) { UdPlan256 PLANVAR; __VA_ARGS__; }                       \
    else if ((n) == 224) { UdPlan224 PLANVAR; __VA_ARGS__; }                       \
    else if ((n) == 299) { UdPlan299 PLANVAR; __VA_ARGS__; }                       \
    else { UdDynPlan PLANVAR; ud_make_dyn_plan((n), &PLANVAR); __VA_ARGS__; }      \
  } while (0)

template <class K>
static int rt_set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    ud_set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", bytes, cudaGetErrorString(e));
    return UD_ERR_CUDA;
  }
  return UD_OK;
}

static int rt_validate(int N, int C, int h, int w, int H, int W) {
  UD_REQUIRE(N >= 0 && C >= 1 && h >= 1 && w >= 1 && H >= 1 && W >= 1, UD_ERR_INVALID,
             "recon_tail: bad shape N=%d C=%d h=%d w=%d H=%d W=%d", N, C, h, w, H, W);
  UD_REQUIRE(ud_fft_size_supported(H) && ud_fft_size_supported(W), UD_ERR_UNSUPPORTED,
             "recon_tail: FFT size %dx%d unsupported (prime factors must be <= 23, n <= %d)", H, W, UD_FFT_MAX_N);
  return UD_OK;
}

extern "C" int ud_recon_tail_fwd(const float* dec, const float* x, float* rec, float* spatial, float* freq,
                                 uint8_t* signs, void* ws, size_t ws_bytes, int N, int C, int h, int w, int H,
                                 int W, int norm_ortho, cudaStream_t stream) {
  int rc = rt_validate(N, C, h, w, H, W);
  if (rc != UD_OK) return rc;
  if (N == 0) return UD_OK;
  UD_REQUIRE(dec && x && rec && spatial && freq && ws, UD_ERR_INVALID, "recon_tail_fwd: null pointer");
  UD_REQUIRE(ws_bytes >= ud_recon_tail_workspace_bytes(N, C, h, w, H, W), UD_ERR_WORKSPACE,
             "recon_tail_fwd: workspace too small (%zu < %zu)", ws_bytes,
             ud_recon_tail_workspace_bytes(N, C, h, w, H, W));
  const int chunk = rt_chunk_samples(N, C, H, W);
  const int Wh = W / 2 + 1;
  const int row_tiles = ud_cdiv(H, 2 * RT_LINES), col_tiles = ud_cdiv(Wh, RT_LINES);
  char* p = static_cast<char*>(ws);
  float2* Y = reinterpret_cast<float2*>(p);
  p += ud_align_up((size_t)chunk * C * H * rt_whp(W) * sizeof(float2), 256);
  float* part_sp = reinterpret_cast<float*>(p);
  p += ud_align_up((size_t)N * C * row_tiles * sizeof(float), 256);
  float* part_fr = reinterpret_cast<float*>(p);
  const float2* twW = ud_twiddles(W);
  const float2* twH = ud_twiddles(H);
  if (!twW || !twH) return UD_ERR_CUDA;
  const float sy = ud_ac_scale(h, H), sx = ud_ac_scale(w, W);
  const size_t smW = rt_fft_smem(W), smH = rt_fft_smem(H);

  for (int s0 = 0; s0 < N; s0 += chunk) {
    const int ns = (N - s0 < chunk) ? (N - s0) : chunk;
    const int planes = ns * C, plane0 = s0 * C;
    RT_DISPATCH_PLAN(W, plan, {
      auto k = rt_rows_fwd_kernel<decltype(plan)>;
      if ((rc = rt_set_smem(k, smW)) != UD_OK) return rc;
      k<<<dim3(row_tiles, planes), RT_THREADS, smW, stream>>>(plan, dec, x, rec, Y, part_sp, twW, plane0, h, w, H,
                                                               W, sy, sx, row_tiles);
    });
    if ((rc = ud_check_launch("rt_rows_fwd")) != UD_OK) return rc;
    RT_DISPATCH_PLAN(H, plan, {
      auto k = rt_cols_fwd_kernel<decltype(plan)>;
      if ((rc = rt_set_smem(k, smH)) != UD_OK) return rc;
      k<<<dim3(col_tiles, planes), RT_THREADS, smH, stream>>>(plan, Y, part_fr, signs, twH, plane0, H, W,
                                                               col_tiles);
    });
    if ((rc = ud_check_launch("rt_cols_fwd")) != UD_OK) return rc;
  }
  const float nrm = norm_ortho ? 1.f / sqrtf((float)H * (float)W) : 1.f;
  rt_finalize_kernel<<<N, 128, 0, stream>>>(part_sp, part_fr, spatial, freq, C, row_tiles, col_tiles,
                                            1.f / ((float)C * H * W), nrm / ((float)C * H * Wh));
  return ud_check_launch("rt_finalize");
}

extern "C" int ud_recon_tail_bwd(const float* dec, const float* x, const uint8_t* signs, const float* g_spatial,
                                 const float* g_freq, float* g_dec, void* ws, size_t ws_bytes, int N, int C,
                                 int h, int w, int H, int W, int norm_ortho, cudaStream_t stream) {
  int rc = rt_validate(N, C, h, w, H, W);
  if (rc != UD_OK) return rc;
  if (N == 0) return UD_OK;
  UD_REQUIRE(dec && x && signs && g_spatial && g_freq && g_dec && ws, UD_ERR_INVALID,
             "recon_tail_bwd: null pointer");
  UD_REQUIRE(ws_bytes >= ud_recon_tail_workspace_bytes(N, C, h, w, H, W), UD_ERR_WORKSPACE,
             "recon_tail_bwd: workspace too small");
  const int chunk = rt_chunk_samples(N, C, H, W);
  const int Wh = W / 2 + 1;
  const int row_tiles = ud_cdiv(H, 2 * RT_LINES), col_tiles = ud_cdiv(Wh, RT_LINES);
  float2* T = reinterpret_cast<float2*>(ws);
  const float2* twW = ud_twiddles(W);
  const float2* twH = ud_twiddles(H);
  if (!twW || !twH) return UD_ERR_CUDA;
  const float sy = ud_ac_scale(h, H), sx = ud_ac_scale(w, W);
  const size_t smH = rt_fft_smem(H);
  const size_t smW = rt_fft_smem(W) + sizeof(float) * 2ull * RT_LINES * w;
  // freq[n] = nrm/(C*H*Wh) * sum |D'| with D' the unnormalised spectrum, so the sign spectrum
  // carries nrm/(C*H*Wh) and the adjoint of the unnormalised forward is the unnormalised inverse.
  const float nrm = norm_ortho ? 1.f / sqrtf((float)H * (float)W) : 1.f;
  const float gscale = nrm / ((float)C * H * Wh);
  const float sp_scale = 1.f / ((float)C * H * W);
  UD_CUDA(cudaMemsetAsync(g_dec, 0, sizeof(float) * (size_t)N * C * h * w, stream));
  for (int s0 = 0; s0 < N; s0 += chunk) {
    const int ns = (N - s0 < chunk) ? (N - s0) : chunk;
    const int planes = ns * C, plane0 = s0 * C;
    RT_DISPATCH_PLAN(H, plan, {
      auto k = rt_cols_bwd_kernel<decltype(plan)>;
      if ((rc = rt_set_smem(k, smH)) != UD_OK) return rc;
      k<<<dim3(col_tiles, planes), RT_THREADS, smH, stream>>>(plan, signs, g_freq, T, twH, plane0, C, H, W, gscale);
    });
    if ((rc = ud_check_launch("rt_cols_bwd")) != UD_OK) return rc;
    RT_DISPATCH_PLAN(W, plan, {
      auto k = rt_rows_bwd_kernel<decltype(plan)>;
      if ((rc = rt_set_smem(k, smW)) != UD_OK) return rc;
      k<<<dim3(row_tiles, planes), RT_THREADS, smW, stream>>>(plan, dec, x, T, g_spatial, g_freq, g_dec, twW,
                                                               plane0, C, h, w, H, W, sy, sx, sp_scale);
    });
    if ((rc = ud_check_launch("rt_rows_bwd")) != UD_OK) return rc;
  }
  return UD_OK;
}

// Shared-memory line FFT building blocks (mixed radix Stockham autosort) for sm_100a.
//
// A "line" is one complex sequence of length n held in shared memory as float2.  A CTA keeps
// L lines resident (line stride LS float2, LS odd so transposing loads are bank-conflict
// free) and runs every stage over all of them; one thread owns one radix-R butterfly at a
// time.  n = R0*R1*...; stage s (radix R, Ns = product of earlier radices):
//     v[r]  = in[j + r*n/R] * W_n^(r * (j mod Ns) * n/(Ns*R))       r = 0..R-1
//     v     = DFT_R(v)
//     out[(j div Ns)*Ns*R + (j mod Ns) + r*Ns] = v[r]
// Natural order in, natural order out, ping-pong between two buffers.  Forward sign only;
// the inverse is obtained by the callers by swapping re/im on load and store.
//
// Plans: StaticPlan<N, R0, R1, ...> (all index arithmetic constant-folded) for the sizes the
// reference's configs use (224, 256, 299, 380) and DynPlan for any other
// n whose prime factors are <= 23.
#pragma once
#include "ud_common.cuh"
#include "ud_fft_bfly_gen.cuh"

#define UD_FFT_MAX_STAGES 10

__device__ __forceinline__ float2 ud_cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

template <int R>
struct UdBfly;
template <>
struct UdBfly<2> {
  static __device__ __forceinline__ void run(float2 (&v)[2]) {
    const float2 a = v[0], b = v[1];
    v[0] = make_float2(a.x + b.x, a.y + b.y);
    v[1] = make_float2(a.x - b.x, a.y - b.y);
  }
};
template <>
struct UdBfly<4> {
  static __device__ __forceinline__ void run(float2 (&v)[4]) {
    const float2 s02 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y);
    const float2 d02 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
    const float2 s13 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
    const float2 d13 = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
    v[0] = make_float2(s02.x + s13.x, s02.y + s13.y);
    v[2] = make_float2(s02.x - s13.x, s02.y - s13.y);
    // forward: X1 = d02 - i*d13, X3 = d02 + i*d13;  -i*(a+bi) = b - ai
    v[1] = make_float2(d02.x + d13.y, d02.y - d13.x);
    v[3] = make_float2(d02.x - d13.y, d02.y + d13.x);
  }
};
#define UD_BFLY_PRIME(P)                                                              \
  template <>                                                                         \
  struct UdBfly<P> {                                                                  \
    static __device__ __forceinline__ void run(float2 (&v)[P]) { ud_bfly##P(v); }     \
  };
UD_BFLY_PRIME(3)
UD_BFLY_PRIME(5)
UD_BFLY_PRIME(7)
UD_BFLY_PRIME(11)
UD_BFLY_PRIME(13)
UD_BFLY_PRIME(17)
UD_BFLY_PRIME(19)
UD_BFLY_PRIME(23)
#undef UD_BFLY_PRIME

// One Stockham stage over L lines.  tw[t] = exp(-2 pi i t / n), t in [0, n).
template <int R>
__device__ __forceinline__ void ud_fft_stage(const float2* __restrict__ in, float2* __restrict__ out,
                                             const float2* __restrict__ tw, const int n, const int Ns,
                                             const int L, const int LS) {
  const int T = n / R;          // butterflies per line
  const int tstride = T / Ns;   // n / (Ns*R)
  const int total = L * T;
  for (int wi = threadIdx.x; wi < total; wi += blockDim.x) {
    const int line = wi / T;
    const int j = wi - line * T;
    const int q = j / Ns;
    const int k = j - q * Ns;
    const float2* src = in + line * LS + j;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = src[r * T];
    if (Ns > 1) {
      const int t1 = k * tstride;
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = ud_cmul(v[r], tw[r * t1]);
    }
    UdBfly<R>::run(v);
    float2* dst = out + line * LS + q * Ns * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) dst[r * Ns] = v[r];
  }
}

// ---- plans ---------------------------------------------------------------------------
template <int N, int NS, int... Rs>
struct UdStaticStages;
template <int N, int NS>
struct UdStaticStages<N, NS> {
  static __device__ __forceinline__ float2* run(float2* a, float2*, const float2*, int, int) { return a; }
};
template <int N, int NS, int R0, int... Rest>
struct UdStaticStages<N, NS, R0, Rest...> {
  static __device__ __forceinline__ float2* run(float2* a, float2* b, const float2* tw, int L, int LS) {
    ud_fft_stage<R0>(a, b, tw, N, NS, L, LS);
    __syncthreads();
    return UdStaticStages<N, NS * R0, Rest...>::run(b, a, tw, L, LS);
  }
};

template <int N, int... Rs>
struct UdStaticPlan {
  static constexpr bool kInPlace = false;
  static constexpr int kN = N;
  __device__ __forceinline__ int n() const { return N; }
  __device__ __forceinline__ int line_len() const { return N; }      // points a line buffer must hold
  // Runs all stages; data must be in `a` and visible (caller synced).  Returns the buffer
  // holding the result; ends with a __syncthreads().
  __device__ __forceinline__ float2* run(float2* a, float2* b, const float2* tw, int L, int LS) const {
    return UdStaticStages<N, 1, Rs...>::run(a, b, tw, L, LS);
  }
};

struct UdDynPlan {
  static constexpr bool kInPlace = false;
  int n_;
  int nstages;
  int radix[UD_FFT_MAX_STAGES];
  __device__ __forceinline__ int n() const { return n_; }
  __device__ __forceinline__ int line_len() const { return n_; }
  __device__ __forceinline__ float2* run(float2* a, float2* b, const float2* tw, int L, int LS) const {
    int Ns = 1;
    for (int s = 0; s < nstages; ++s) {
      const int R = radix[s];
      switch (R) {
        case 2: ud_fft_stage<2>(a, b, tw, n_, Ns, L, LS); break;
        case 3: ud_fft_stage<3>(a, b, tw, n_, Ns, L, LS); break;
        case 4: ud_fft_stage<4>(a, b, tw, n_, Ns, L, LS); break;
        case 5: ud_fft_stage<5>(a, b, tw, n_, Ns, L, LS); break;
        case 7: ud_fft_stage<7>(a, b, tw, n_, Ns, L, LS); break;
        case 11: ud_fft_stage<11>(a, b, tw, n_, Ns, L, LS); break;
        case 13: ud_fft_stage<13>(a, b, tw, n_, Ns, L, LS); break;
        case 17: ud_fft_stage<17>(a, b, tw, n_, Ns, L, LS); break;
        case 19: ud_fft_stage<19>(a, b, tw, n_, Ns, L, LS); break;
        default: ud_fft_stage<23>(a, b, tw, n_, Ns, L, LS); break;
      }
      __syncthreads();
      Ns *= R;
      float2* t = a;
      a = b;
      b = t;
    }
    return a;
  }
};

// ---- in-place static plans ------------------------------------------------------------------
// Same Stockham stages, but ONE buffer: every thread first pulls all of its butterflies for the stage
// into registers, the CTA synchronises, then results are written back in autosort order.  Halves the
// shared memory per CTA (=> twice the resident CTAs per SM) at the price of one extra barrier per
// stage.  N, L (lines per CTA) and THREADS are compile-time so every index is constant-folded.
template <int R, int N, int NS, int L, int THREADS>
__device__ __forceinline__ void ud_fft_stage_ip(float2* __restrict__ buf, const float2* __restrict__ tw) {
  constexpr int T = N / R;
  constexpr int TOTAL = L * T;
  constexpr int PASSES = (TOTAL + THREADS - 1) / THREADS;
  constexpr int LS = N | 1;
  constexpr int TS = T / NS;
  constexpr bool FULL = (PASSES * THREADS == TOTAL);
  float2 v[PASSES][R];
#pragma unroll
  for (int p = 0; p < PASSES; ++p) {
    const int wi = (int)threadIdx.x + p * THREADS;
    if (FULL || wi < TOTAL) {
      const int line = wi / T;
      const int j = wi - line * T;
      const float2* src = buf + line * LS + j;
#pragma unroll
      for (int r = 0; r < R; ++r) v[p][r] = src[r * T];
    }
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < PASSES; ++p) {
    const int wi = (int)threadIdx.x + p * THREADS;
    if (FULL || wi < TOTAL) {
      const int line = wi / T;
      const int j = wi - line * T;
      const int q = j / NS;
      const int k = j - q * NS;
      if (NS > 1) {
        const int t1 = k * TS;
#pragma unroll
        for (int r = 1; r < R; ++r) v[p][r] = ud_cmul(v[p][r], tw[r * t1]);
      }
      UdBfly<R>::run(v[p]);
      float2* dst = buf + line * LS + q * (NS * R) + k;
#pragma unroll
      for (int r = 0; r < R; ++r) dst[r * NS] = v[p][r];
    }
  }
  __syncthreads();
}

template <int N, int NS, int L, int THREADS, int... Rs>
struct UdIpStages;
template <int N, int NS, int L, int THREADS>
struct UdIpStages<N, NS, L, THREADS> {
  static __device__ __forceinline__ void run(float2*, const float2*) {}
};
template <int N, int NS, int L, int THREADS, int R0, int... Rest>
struct UdIpStages<N, NS, L, THREADS, R0, Rest...> {
  static __device__ __forceinline__ void run(float2* buf, const float2* tw) {
    ud_fft_stage_ip<R0, N, NS, L, THREADS>(buf, tw);
    UdIpStages<N, NS * R0, L, THREADS, Rest...>::run(buf, tw);
  }
};

// Plan concept used by the kernels:  kInPlace, n(), line_len() (>= n: the points a line buffer and the staged twiddle
// table must hold -- Bluestein plans work at a longer length), run(a, b, tw, L, LS) -> buffer holding the result.
template <int N, int L, int THREADS, int... Rs>
struct UdStaticPlanIP {
  static constexpr bool kInPlace = true;
  static constexpr int kN = N;
  __device__ __forceinline__ int n() const { return N; }
  __device__ __forceinline__ int line_len() const { return N; }
  __device__ __forceinline__ float2* run(float2* a, float2*, const float2* tw, int, int) const {
    UdIpStages<N, 1, L, THREADS, Rs...>::run(a, tw);
    return a;
  }
};

typedef UdStaticPlan<380, 19, 5, 4> UdPlan380;
typedef UdStaticPlan<256, 4, 4, 4, 4> UdPlan256;
typedef UdStaticPlan<224, 7, 4, 4, 2> UdPlan224;
typedef UdStaticPlan<299, 23, 13> UdPlan299;

// ---- host side -----------------------------------------------------------------------
// Factorises n into supported radices (largest odd primes first, then 4s, then a 2).
// Returns false when n has a prime factor > 23.
bool ud_make_dyn_plan(int n, UdDynPlan* plan);
// Device twiddle table exp(-2 pi i t/n), t in [0,n), cached per (device, n) for the process
// lifetime (mutex-guarded, never freed).  nullptr on failure (error set).
const float2* ud_twiddles(int n);
// align_corners=True bilinear table for resizing `in` -> `out` samples along one axis:
// entry d = { __int_as_float(i0), l1 } with ATen's fp32 arithmetic (scale=(in-1)/(out-1), src=scale*d,
// i0=(int)src, l1=src-i0); i1 = i0 + (i0 < in-1), l0 = 1-l1.  Cached per (device, in, out) like the twiddles.
const float2* ud_lerp_table(int in, int out);
// For the transposed resize: entry j (0 <= j < in) = {first, last} output sample whose taps (i0 or i1)
// include input sample j; first > last when none does.  Derived from the same fp32 table.
const int2* ud_lerp_ranges(int in, int out);
// padded line stride (odd) for n-point lines
static inline int ud_line_stride(int n) { return n | 1; }

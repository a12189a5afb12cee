// Register-resident two-stage line FFT (N = R1 * R2) for sm_100a.
//
// A line of N complex points is owned by TL = max(R1, R2) threads.  Cooley-Tukey, decimation in time:
//     n = R2*n1 + n2,  k = k1 + R1*k2
//     stage A: thread n2 (< R2) holds x[n2 + R2*n1], n1 = 0..R1-1 in registers, runs ONE radix-R1 butterfly
//              (generated straight-line code with literal twiddles), multiplies by W_N^(n2*k1) and writes
//              S[k1*P + n2] to shared memory (P = R2|1: conflict-free for the column reads that follow);
//     stage B: thread k1 (< R1) reads S[k1*P + n2], n2 = 0..R2-1, runs ONE radix-R2 butterfly and holds
//              X[k1 + R1*k2], k2 = 0..R2-1.
// One shared-memory exchange and one barrier per line instead of one per radix stage, no index arithmetic inside
// the butterflies (a thread's role is fixed), and global loads/stores happen straight from/to the registers the
// butterflies use: for a fixed n1 (k2) the TL threads of a line touch R2 (R1) consecutive elements.
// Forward sign only; callers obtain the inverse with the re/im swap trick.
// 380 = 20*19, 256 = 16*16, 224 = 16*14, 299 = 13*23 (SURVEY.md App. A.1 image sizes).
#pragma once
#include "ud_common.cuh"
#include "ud_fft_bfly_gen.cuh"

template <int R>
struct UdB;
#define UD_B(R)                                                                  \
  template <>                                                                    \
  struct UdB<R> {                                                                \
    static __device__ __forceinline__ void run(float2 (&v)[R]) { ud_bfly##R(v); } \
  };
UD_B(2) UD_B(3) UD_B(4) UD_B(5) UD_B(7) UD_B(8) UD_B(11) UD_B(12) UD_B(13) UD_B(14) UD_B(16) UD_B(17) UD_B(19) UD_B(20) UD_B(23)
#undef UD_B

template <int N_, int R1_, int R2_>
struct Ud2S {
  static_assert(R1_ * R2_ == N_, "N = R1 * R2");
  static constexpr int N = N_, R1 = R1_, R2 = R2_;
  static constexpr int TL = R1_ > R2_ ? R1_ : R2_;                     // threads per line
  static constexpr int P = R2_ | 1;                                    // pitch of the exchange layout
  static constexpr int LS = ((R1_ * P > N_ ? R1_ * P : N_) | 1);       // float2 slots of one line buffer
};

// tw2[k1*P + n2] = exp(-2 pi i n2 k1 / N), from the library's global table tw_g[t] = exp(-2 pi i t / N)
template <class PL>
__device__ __forceinline__ void ud2s_build_tw(float2* __restrict__ tw2, const float2* __restrict__ tw_g) {
  for (int t = threadIdx.x; t < PL::R1 * PL::P; t += blockDim.x) {
    const int k1 = t / PL::P, n2 = t - k1 * PL::P;
    tw2[t] = (n2 < PL::R2) ? __ldg(tw_g + n2 * k1) : make_float2(0.f, 0.f);
  }
}

// stage A on the thread's R1 values (z[n1] = x[n2 + R2*n1]); results go to the line's exchange buffer S
template <class PL>
__device__ __forceinline__ void ud2s_stage_a(float2 (&z)[PL::R1], int n2, const float2* __restrict__ tw2,
                                             float2* __restrict__ S) {
  UdB<PL::R1>::run(z);
  S[n2] = z[0];
#pragma unroll
  for (int k1 = 1; k1 < PL::R1; ++k1) {
    const float2 t = tw2[k1 * PL::P + n2];
    S[k1 * PL::P + n2] = make_float2(z[k1].x * t.x - z[k1].y * t.y, z[k1].x * t.y + z[k1].y * t.x);
  }
}

// stage B: u[k2] = X[k1 + R1*k2]
template <class PL>
__device__ __forceinline__ void ud2s_stage_b(float2 (&u)[PL::R2], int k1, const float2* __restrict__ S) {
#pragma unroll
  for (int n2 = 0; n2 < PL::R2; ++n2) u[n2] = S[k1 * PL::P + n2];
  UdB<PL::R2>::run(u);
}

// a17 (first step) -- glue kernels of the SFConv frequency branch
//
// Reference: SFConv2dStaticSamePadding.forward (model/efficientnet/exp.py:46-65), SFConv2d.forward
// (model/resnet/exp.py:36-54):
//     F = rfft2(x); P = cat([F.real, F.imag], 1); Q = conv1x1(P); y = irfft2(complex(*tensor_split(Q, 2, 1)))
//     out = (1 - sigmoid(sf_coef)) * conv(x) + sigmoid(sf_coef) * pool(y)
// The FFTs (cuFFT) and the 1x1 convolution (cuDNN / cuBLAS) stay library calls for now; these kernels replace
// the ~10 elementwise / layout / dtype passes torch makes between them with ONE pass each:
//   sf_pack   : interleaved complex64 [N,C,P] -> planar cat([re, im]) [N,2C,P] in the convolution's dtype
//               (fp32 | bf16) and memory format (NCHW | channels-last), through a shared-memory tile transpose
//   sf_unpack : the inverse (also each other's autograd adjoint)
//   sf_mix    : the sigmoid-gated blend of the spatial branch (conv dtype / layout) with the fp32 NCHW frequency
//               branch, forward and backward (gate gradient reduced deterministically).
#include <cuda_bf16.h>

#include "../../include/unidefense_b200.h"
#include "ud_common.cuh"

#define SF_TC 64   // channels per tile
#define SF_TP 32   // positions per tile

template <bool BF16>
struct SfIO;
template <>
struct SfIO<false> {
  typedef float T;
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  static __device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
};
template <>
struct SfIO<true> {
  typedef __nv_bfloat16 T;
  static __device__ __forceinline__ float ld(const T* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(T* p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float2 ld2(const T* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
  }
  static __device__ __forceinline__ void st2(T* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
  }
};

// ---- NCHW planar: no transpose needed --------------------------------------------------------------
template <bool BF16, bool PACK>
__global__ void sf_pack_nchw_kernel(float2* __restrict__ spec, typename SfIO<BF16>::T* __restrict__ planar, long long total,
                                    int C, int P) {
  typedef SfIO<BF16> IO;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int p = (int)(i % P);
    const long long nc = i / P;
    const int c = (int)(nc % C);
    const long long n = nc / C;
    typename IO::T* re = planar + ((n * 2 * C + c) * (long long)P + p);
    typename IO::T* im = re + (long long)C * P;
    if (PACK) {
      const float2 v = spec[i];
      IO::st(re, v.x);
      IO::st(im, v.y);
    } else {
      spec[i] = make_float2(IO::ld(re), IO::ld(im));
    }
  }
}

// ---- channels-last planar [N, P, 2C]: tile transpose (64 channels x 32 positions) ----------------------
//   grid (ceil(P/32), ceil(C/64), N), block 256.  C must be even.
template <bool BF16, bool PACK>
__global__ void __launch_bounds__(256)
sf_pack_nhwc_kernel(float2* __restrict__ spec, typename SfIO<BF16>::T* __restrict__ planar, int C, int P) {
  typedef SfIO<BF16> IO;
  __shared__ float2 tile[SF_TC][SF_TP + 1];
  const int p0 = blockIdx.x * SF_TP, c0 = blockIdx.y * SF_TC;
  const long long n = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* sp = spec + n * (long long)C * P;
  typename IO::T* pl = planar + n * (long long)P * 2 * C;
  if (PACK) {
#pragma unroll
    for (int r = 0; r < SF_TC / 8; ++r) {
      const int c = c0 + warp * (SF_TC / 8) + r, p = p0 + lane;
      if (c < C && p < P) tile[warp * (SF_TC / 8) + r][lane] = sp[(long long)c * P + p];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SF_TP / 8; ++r) {
      const int pp = warp * (SF_TP / 8) + r, p = p0 + pp;
      const int c = c0 + 2 * lane;
      if (p < P && c < C) {
        const float2 a = tile[2 * lane][pp], b = tile[2 * lane + 1][pp];
        typename IO::T* dst = pl + (long long)p * 2 * C + c;
        IO::st2(dst, a.x, b.x);
        IO::st2(dst + C, a.y, b.y);
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < SF_TP / 8; ++r) {
      const int pp = warp * (SF_TP / 8) + r, p = p0 + pp;
      const int c = c0 + 2 * lane;
      if (p < P && c < C) {
        const typename IO::T* src = pl + (long long)p * 2 * C + c;
        const float2 re = IO::ld2(src), im = IO::ld2(src + C);
        tile[2 * lane][pp] = make_float2(re.x, im.x);
        tile[2 * lane + 1][pp] = make_float2(re.y, im.y);
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SF_TC / 8; ++r) {
      const int c = c0 + warp * (SF_TC / 8) + r, p = p0 + lane;
      if (c < C && p < P) sp[(long long)c * P + p] = tile[warp * (SF_TC / 8) + r][lane];
    }
  }
}

template <bool PACK>
static int sf_pack_launch(float2* spec, void* planar, int N, int C, int P, int nhwc, int bf16, cudaStream_t stream,
                          const char* what) {
  UD_REQUIRE(N >= 0 && C >= 1 && P >= 1, UD_ERR_INVALID, "%s: bad shape N=%d C=%d P=%d", what, N, C, P);
  if (N == 0) return UD_OK;
  UD_REQUIRE(spec && planar, UD_ERR_INVALID, "%s: null pointer", what);
  if (nhwc) {
    UD_REQUIRE(C % 2 == 0, UD_ERR_UNSUPPORTED, "%s: channels-last planar layout needs an even channel count (C=%d)", what, C);
    UD_REQUIRE(N <= 65535 && ud_cdiv(C, SF_TC) <= 65535, UD_ERR_UNSUPPORTED, "%s: grid too large", what);
    dim3 grid(ud_cdiv(P, SF_TP), ud_cdiv(C, SF_TC), N);
    if (bf16) sf_pack_nhwc_kernel<true, PACK><<<grid, 256, 0, stream>>>(spec, static_cast<__nv_bfloat16*>(planar), C, P);
    else sf_pack_nhwc_kernel<false, PACK><<<grid, 256, 0, stream>>>(spec, static_cast<float*>(planar), C, P);
  } else {
    const long long total = (long long)N * C * P;
    const int blocks = (int)min((long long)UD_NUM_SMS * 16, (total + 255) / 256);
    if (bf16) sf_pack_nchw_kernel<true, PACK><<<blocks, 256, 0, stream>>>(spec, static_cast<__nv_bfloat16*>(planar), total, C, P);
    else sf_pack_nchw_kernel<false, PACK><<<blocks, 256, 0, stream>>>(spec, static_cast<float*>(planar), total, C, P);
  }
  return ud_check_launch(what);
}

extern "C" int ud_sf_pack(const void* spec, void* planar, int N, int C, int P, int nhwc, int bf16, cudaStream_t stream) {
  return sf_pack_launch<true>(static_cast<float2*>(const_cast<void*>(spec)), planar, N, C, P, nhwc, bf16, stream, "sf_pack");
}
extern "C" int ud_sf_unpack(const void* planar, void* spec, int N, int C, int P, int nhwc, int bf16, cudaStream_t stream) {
  return sf_pack_launch<false>(static_cast<float2*>(spec), const_cast<void*>(planar), N, C, P, nhwc, bf16, stream, "sf_unpack");
}

// ---- gated blend ------------------------------------------------------------------------------------
// spat / out / g_out / g_spat: dtype T, element (n,c,p) at NCHW or channels-last position; freq / g_freq: fp32 NCHW.
//   grid (ceil(P/32), ceil(C/64), N), block 256; tile transpose only on the fp32 NCHW side when NHWC.
template <bool BF16, bool NHWC, bool BWD>
__global__ void __launch_bounds__(256)
sf_mix_kernel(const typename SfIO<BF16>::T* __restrict__ spat, const float* __restrict__ freq, const float* __restrict__ coef,
              typename SfIO<BF16>::T* __restrict__ out, const typename SfIO<BF16>::T* __restrict__ g_out,
              typename SfIO<BF16>::T* __restrict__ g_spat, float* __restrict__ g_freq, float* __restrict__ part, int C, int P) {
  typedef SfIO<BF16> IO;
  __shared__ float tf[SF_TC][SF_TP + 1];    // freq tile (fwd) / g_freq tile (bwd)
  __shared__ float red[33];
  const float s = ud_sigmoid(__ldg(coef));
  const int p0 = blockIdx.x * SF_TP, c0 = blockIdx.y * SF_TC;
  const long long n = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* fq = freq + n * (long long)C * P;
  float acc = 0.f;
  if (!NHWC) {
    // everything NCHW: thread (c, p) with p fastest
#pragma unroll
    for (int r = 0; r < SF_TC / 8; ++r) {
      const int c = c0 + warp * (SF_TC / 8) + r, p = p0 + lane;
      if (c < C && p < P) {
        const long long i = (n * C + c) * (long long)P + p;
        const float sv = IO::ld(spat + i), fv = fq[(long long)c * P + p];
        if (!BWD) {
          IO::st(out + i, (1.f - s) * sv + s * fv);
        } else {
          const float g = IO::ld(g_out + i);
          IO::st(g_spat + i, (1.f - s) * g);
          g_freq[i] = s * g;
          acc = fmaf(g, fv - sv, acc);
        }
      }
    }
  } else {
    // freq side (NCHW) through the tile; spat side channels-last, 2 channels per lane
    if (!BWD) {
#pragma unroll
      for (int r = 0; r < SF_TC / 8; ++r) {
        const int cc = warp * (SF_TC / 8) + r, c = c0 + cc, p = p0 + lane;
        if (c < C && p < P) tf[cc][lane] = fq[(long long)c * P + p];
      }
      __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < SF_TP / 8; ++r) {
      const int pp = warp * (SF_TP / 8) + r, p = p0 + pp;
      const int c = c0 + 2 * lane;
      if (p < P && c < C) {
        const long long i = (n * P + p) * (long long)C + c;
        const float2 sv = IO::ld2(spat + i);
        if (!BWD) {
          IO::st2(out + i, (1.f - s) * sv.x + s * tf[2 * lane][pp], (1.f - s) * sv.y + s * tf[2 * lane + 1][pp]);
        } else {
          const float2 g = IO::ld2(g_out + i);
          IO::st2(g_spat + i, (1.f - s) * g.x, (1.f - s) * g.y);
          tf[2 * lane][pp] = g.x;       // staged: g_freq = s*g and the gate term need freq in NCHW order
          tf[2 * lane + 1][pp] = g.y;
          // gate term uses spat here and freq below: acc -= g*spat now
          acc = fmaf(-g.x, sv.x, acc);
          acc = fmaf(-g.y, sv.y, acc);
        }
      }
    }
    if (BWD) {
      __syncthreads();
      float* gf = g_freq + n * (long long)C * P;
#pragma unroll
      for (int r = 0; r < SF_TC / 8; ++r) {
        const int cc = warp * (SF_TC / 8) + r, c = c0 + cc, p = p0 + lane;
        if (c < C && p < P) {
          const float g = tf[cc][lane];
          gf[(long long)c * P + p] = s * g;
          acc = fmaf(g, fq[(long long)c * P + p], acc);
        }
      }
    }
  }
  if (BWD) {
    const float tot = ud_block_sum(acc, red);
    if (threadIdx.x == 0)
      part[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot * s * (1.f - s);
  }
}

__global__ void sf_sum_kernel(const float* __restrict__ part, float* __restrict__ out, long long n) {
  __shared__ float red[33];
  float a = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) a += part[i];
  a = ud_block_sum(a, red);
  if (threadIdx.x == 0) *out = a;
}

extern "C" size_t ud_sf_mix_bwd_workspace_bytes(int N, int C, int P) {
  return sizeof(float) * (size_t)N * ud_cdiv(C, SF_TC) * ud_cdiv(P, SF_TP);
}

template <bool BWD>
static int sf_mix_launch(const void* spat, const float* freq, const float* coef, void* out, const void* g_out, void* g_spat,
                         float* g_freq, float* part, int N, int C, int P, int nhwc, int bf16, cudaStream_t stream) {
  dim3 grid(ud_cdiv(P, SF_TP), ud_cdiv(C, SF_TC), N);
#define SF_MIX(B, H)                                                                                                     \
  sf_mix_kernel<B, H, BWD><<<grid, 256, 0, stream>>>(static_cast<const typename SfIO<B>::T*>(spat), freq, coef,          \
                                                     static_cast<typename SfIO<B>::T*>(out),                             \
                                                     static_cast<const typename SfIO<B>::T*>(g_out),                     \
                                                     static_cast<typename SfIO<B>::T*>(g_spat), g_freq, part, C, P)
  if (bf16 && nhwc) SF_MIX(true, true);
  else if (bf16) SF_MIX(true, false);
  else if (nhwc) SF_MIX(false, true);
  else SF_MIX(false, false);
#undef SF_MIX
  return ud_check_launch(BWD ? "sf_mix_bwd" : "sf_mix_fwd");
}

static int sf_mix_validate(const char* what, int N, int C, int P, int nhwc) {
  UD_REQUIRE(N >= 0 && C >= 1 && P >= 1, UD_ERR_INVALID, "%s: bad shape N=%d C=%d P=%d", what, N, C, P);
  UD_REQUIRE(!nhwc || C % 2 == 0, UD_ERR_UNSUPPORTED, "%s: channels-last layout needs an even channel count (C=%d)", what, C);
  UD_REQUIRE(N <= 65535 && ud_cdiv(C, SF_TC) <= 65535, UD_ERR_UNSUPPORTED, "%s: grid too large", what);
  return UD_OK;
}

extern "C" int ud_sf_mix_fwd(const void* spat, const float* freq, const float* coef, void* out, int N, int C, int P,
                             int nhwc, int bf16, cudaStream_t stream) {
  int rc = sf_mix_validate("sf_mix_fwd", N, C, P, nhwc);
  if (rc != UD_OK) return rc;
  if (N == 0) return UD_OK;
  UD_REQUIRE(spat && freq && coef && out, UD_ERR_INVALID, "sf_mix_fwd: null pointer");
  return sf_mix_launch<false>(spat, freq, coef, out, nullptr, nullptr, nullptr, nullptr, N, C, P, nhwc, bf16, stream);
}

extern "C" int ud_sf_mix_bwd(const void* g_out, const void* spat, const float* freq, const float* coef, void* g_spat,
                             float* g_freq, float* g_coef, void* ws, size_t ws_bytes, int N, int C, int P, int nhwc,
                             int bf16, cudaStream_t stream) {
  int rc = sf_mix_validate("sf_mix_bwd", N, C, P, nhwc);
  if (rc != UD_OK) return rc;
  UD_REQUIRE(g_coef, UD_ERR_INVALID, "sf_mix_bwd: null pointer");
  if (N == 0) {
    UD_CUDA(cudaMemsetAsync(g_coef, 0, sizeof(float), stream));
    return UD_OK;
  }
  UD_REQUIRE(g_out && spat && freq && coef && g_spat && g_freq && ws, UD_ERR_INVALID, "sf_mix_bwd: null pointer");
  UD_REQUIRE(ws_bytes >= ud_sf_mix_bwd_workspace_bytes(N, C, P), UD_ERR_WORKSPACE, "sf_mix_bwd: workspace too small");
  float* part = static_cast<float*>(ws);
  rc = sf_mix_launch<true>(spat, freq, coef, nullptr, g_out, g_spat, g_freq, part, N, C, P, nhwc, bf16, stream);
  if (rc != UD_OK) return rc;
  sf_sum_kernel<<<1, 256, 0, stream>>>(part, g_coef, (long long)N * ud_cdiv(C, SF_TC) * ud_cdiv(P, SF_TP));
  return ud_check_launch("sf_mix_bwd_sum");
}

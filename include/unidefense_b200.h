/* libunidefense_b200.so -- C ABI of the B200-native UniDefense dual-space reconstruction path.
 *
 * The reference (VISION-SJTU/UniDefense) is pure Python/PyTorch and has no FFI; its boundary for
 * this path is the nn.Module contract of model/unidefense.py (forward() -> cls_out/rec/loss_dict).
 * Each entry point below replaces the torch call sequence cited next to it; the Python mirror
 * of the reference interface (unidefense_b200/model, /loss) binds these through ctypes.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer into caller-owned memory (torch's caching allocator);
 *    the library never allocates or frees caller tensors.  It owns only immutable per-size
 *    twiddle tables, created lazily, cached for the process lifetime (mutex-guarded).
 *  - tensors are fp32, NCHW, contiguous.  Spectra use the reference's channel-planar
 *    cat([re, im], dim=1) layout ([N, 2C, H, W/2+1]), never interleaved complex.
 *  - `stream` is the CUDA stream to launch on (the shim passes torch's current stream);
 *    entry points are re-entrant and may be called from autograd's worker threads.
 *  - return 0 on success, a negative UD_ERR_* code otherwise; ud_last_error() returns a
 *    thread-local message.  No exceptions cross this boundary.
 *  - workspaces are passed in by the caller; *_workspace_bytes() say how large.
 */
#ifndef UNIDEFENSE_B200_H
#define UNIDEFENSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if !defined(__CUDA_RUNTIME_H__) && !defined(__DRIVER_TYPES_H__)
typedef struct CUstream_st* cudaStream_t;
#endif

#if defined(__GNUC__)
#define UD_API __attribute__((visibility("default")))
#else
#define UD_API
#endif

#define UD_B200_VERSION 100
#define UD_FFT_MAX_N 1024

#define UD_ACT_NONE 0
#define UD_ACT_RELU 1   /* nn.ReLU            (model/unidefense.py:277,457) */
#define UD_ACT_SWISH 2  /* MemoryEfficientSwish (model/unidefense.py:56; efficientnet/utils.py:66-82) */

UD_API const char* ud_last_error(void);
UD_API int ud_version(void);
/* 1 when n-point line FFTs are supported (prime factors <= 23, n <= UD_FFT_MAX_N). */
UD_API int ud_fft_size_supported(int n);

/* ---- a1: reconstruction-loss tail --------------------------------------------------------
 * Replaces model/unidefense.py:244-253 (Eb4), :423-433 (Res18), :618-628 (Res50):
 *   interpolate(dec, x.shape[-2:]) ; abs(rec-x).mean ; rfft2 x2 ; cat ; abs ; tensor_split ; mean.
 * dec [N,C,h,w], x [N,C,H,W] -> rec [N,C,H,W], spatial [N], freq [N].
 * signs (nullable; ud_recon_tail_signs_bytes) receives 1 byte per spectral bin (sign codes of
 * Re/Im dF) and is what backward needs instead of the two saved spectra.                    */
UD_API size_t ud_recon_tail_workspace_bytes(int N, int C, int h, int w, int H, int W);
UD_API size_t ud_recon_tail_signs_bytes(int N, int C, int H, int W);
UD_API int ud_recon_tail_fwd(const float* dec, const float* x, float* rec, float* spatial, float* freq,
                             uint8_t* signs, void* ws, size_t ws_bytes, int N, int C, int h, int w, int H,
                             int W, int norm_ortho, cudaStream_t stream);
/* g_dec [N,C,h,w] = d(sum_n g_spatial[n]*spatial[n] + g_freq[n]*freq[n]) / d dec
 * (torch autograd of the lines above: abs -> sgn, fft_r2c_backward, upsample_bilinear2d_backward).
 * Samples with g_spatial[n]==g_freq[n]==0 (the engine's fake rows, abstract_engine.py:241,249)
 * cost nothing.                                                                              */
UD_API int ud_recon_tail_bwd(const float* dec, const float* x, const uint8_t* signs, const float* g_spatial,
                             const float* g_freq, float* g_dec, void* ws, size_t ws_bytes, int N, int C, int h,
                             int w, int H, int W, int norm_ortho, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIDEFENSE_B200_H */

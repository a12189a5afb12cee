// a5 / a6 -- the dense projections of the dynamic filters on the 5th-generation tensor cores
//                                                                     
// Reference: model/modules.py:82-85  FrequencyDynamicFilter.layer1[0]  nn.Conv2d(2C, 2C, 1)
//            model/modules.py:111-114 SpatialDynamicFilter.layer1[0]   nn.Conv2d(C, C, 3, 1, 1)
// followed by nn.BatchNorm2d (layer1[1]) whose batch statistics this kernel produces in its epilogue.
//
// One kernel for both: an implicit GEMM  D[pixel, co] = sum_{tap, ci} X[pixel + tap, ci] * Wk[co, tap, ci]
//   * A (activations, channels-last fp32) is fetched by TMA straight from the [N,H,W,C] tensor: a 4-D box
//     {32 channels, W, bh rows, bn samples} whose coordinates are shifted by the tap; the zero padding of the
//     3x3 convolution is the TMA's out-of-bounds fill.  The 1x1 projection uses the flat [N*H*W, C] view.
//   * B (weights [Cout, taps, Cin], K contiguous) by TMA boxes {32, BLOCK_N}.
//   * both land in 128-byte-swizzled shared memory, a 3-stage mbarrier ring feeds `tcgen05.mma.kind::tf32`
//     (M=128, N=128, K=8 per instruction) issued by one thread; the fp32 accumulator lives in TMEM.
//   * precision: TF32 (what cuDNN runs for torch's default allow_tf32=True), or "3xTF32": operands are
//     pre-split into hi (exactly representable in TF32) + lo = x - hi and D += Alo*Bhi + Ahi*Blo + Ahi*Bhi,
//     which restores fp32-grade accuracy (~1e-6 relative) on the tensor pipe.
//   * epilogue (4 warps, tcgen05.ld 32 lanes x 32 columns): writes D as NCHW (a warp stores 32 consecutive
//     pixels of one channel: coalesced) and stages the tile in the drained pipeline buffers to emit the
//     BatchNorm partial statistics of the tile, per channel: (mean, M2 = sum (x-mean)^2) over its valid
//     pixels -- two-pass inside the tile, merged across tiles with Chan's formula by ud_bn_merge_partials.
//     The separate full read of `proj` by ud_bn_stats disappears.
#include <cuda.h>
#include <string.h>

#include <mutex>

#include "../../include/unidefense_b200.h"
#include "ud_common.cuh"

#define PJ_BLOCK_M 128
#define PJ_BLOCK_N 128
#define PJ_BLOCK_K 32            // fp32 elements = one 128-byte swizzle row
#define PJ_UMMA_K 8              // tf32
#define PJ_STAGES 3
#define PJ_THREADS 256           // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-7 epilogue
#define PJ_A_BYTES (PJ_BLOCK_M * PJ_BLOCK_K * 4)
#define PJ_B_BYTES (PJ_BLOCK_N * PJ_BLOCK_K * 4)
#define PJ_TMEM_COLS 128
// 3xTF32 only: the tensor core adds into its fp32 accumulator with truncation, a bias that grows with the number of
// accumulation steps (K = 18432 for the ResNet-50 3x3 projection).  Rotating over 4 accumulators keeps each partial
// sum 4x smaller (so its ulp, and the bias, shrink accordingly); the epilogue adds them in round-to-nearest fp32.
#define PJ_SPLIT_ACCS 4

struct PjGeom {
  int N, H, W, Cin, Cout, taps;   // taps = 1 (1x1) or 9 (3x3, pad 1)
  int flat;                       // 1: A is the flat [N*H*W, Cin] matrix (1x1 only)
  int bh, bn;                     // 4-D box: bh image rows x bn samples (x W columns) per M tile
  int tiles_per_sample;           // 4-D, bn == 1: ceil(H / bh)
  int m_tiles, n_tiles, kc;       // kc = ceil(Cin / 32) k-blocks per tap
  int a_rows, b_rows;             // rows the A / B boxes really carry (<= 128): the TMA transaction size
  int wgrad;                      // 1: weight-gradient GEMM of the 1x1 projection (see ud_proj_wgrad_1x1): A and B are both
                                  // NCHW activations [n][channel][pixel], K runs over (sample, 32-pixel chunk)
  int kp;                         // wgrad: 32-pixel chunks per sample = ceil(P / 32)
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pj_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void pj_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pj_smem(bar)), "r"(count));
}
__device__ __forceinline__ void pj_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pj_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pj_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(pj_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void pj_tma_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
                   "r"(pj_smem(dst)), "l"(map), "r"(pj_smem(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void pj_tma_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
                   "r"(pj_smem(dst)), "l"(map), "r"(pj_smem(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void pj_tma_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
                   "r"(pj_smem(dst)), "l"(map), "r"(pj_smem(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t pj_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void pj_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void pj_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pj_smem(bar)) : "memory");
}
__device__ __forceinline__ void pj_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// tile row r -> (sample, pixel) of the output it produces; false for padding rows of the tile
__device__ __forceinline__ bool pj_row(const PjGeom& g, int m_tile, int r, int& n, int& p) {
  const int P = g.H * g.W;
  if (g.wgrad) {                   // D'[ci, co]: "pixel" = ci of a single pseudo-sample with P = Cin rows
    const int m = m_tile * PJ_BLOCK_M + r;
    if (m >= g.H) return false;
    n = 0;
    p = m;
    return true;
  }
  if (g.flat) {
    const long long m = (long long)m_tile * PJ_BLOCK_M + r;
    if (m >= (long long)g.N * P) return false;
    n = (int)(m / P);
    p = (int)(m - (long long)n * P);
    return true;
  }
  const int per_n = g.bh * g.W;
  if (r >= g.bn * per_n) return false;
  const int dn = r / per_n, rem = r - dn * per_n;
  const int dh = rem / g.W, dw = rem - dh * g.W;
  int n0, h0;
  if (g.bn > 1) { n0 = m_tile * g.bn; h0 = 0; }
  else { n0 = m_tile / g.tiles_per_sample; h0 = (m_tile - n0 * g.tiles_per_sample) * g.bh; }
  n = n0 + dn;
  const int h = h0 + dh;
  if (n >= g.N || h >= g.H) return false;
  p = h * g.W + dw;
  return true;
}

template <bool SPLIT3>
__global__ void __launch_bounds__(PJ_THREADS)
pj_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               float* __restrict__ y, float* __restrict__ part_mean, float* __restrict__ part_m2,
               float* __restrict__ part_cnt, const PjGeom g) {
  constexpr int NOPS = SPLIT3 ? 2 : 1;
  constexpr int NACC = SPLIT3 ? PJ_SPLIT_ACCS : 1;
  constexpr uint32_t STAGE_BYTES = NOPS * (PJ_A_BYTES + PJ_B_BYTES);
  extern __shared__ uint8_t pj_smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>(((uintptr_t)pj_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + PJ_STAGES * STAGE_BYTES);
  uint64_t* empty = full + PJ_STAGES;
  uint64_t* acc_full = empty + PJ_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y;
  const int num_kb = g.wgrad ? g.N * g.kp : g.taps * g.kc;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PJ_STAGES; ++s) {
      pj_mbar_init(full + s, 1);
      pj_mbar_init(empty + s, 1);
    }
    pj_mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pj_smem(tmem_slot)), "r"(PJ_TMEM_COLS * NACC));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    int n0 = 0, h0 = 0;
    if (!g.flat) {
      if (g.bn > 1) n0 = m_tile * g.bn;
      else { n0 = m_tile / g.tiles_per_sample; h0 = (m_tile - n0 * g.tiles_per_sample) * g.bh; }
    }
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % PJ_STAGES;
      const uint32_t ph = (kb / PJ_STAGES) & 1;
      pj_mbar_wait(empty + s, ph ^ 1);
      uint8_t* st = tiles + s * STAGE_BYTES;
      pj_mbar_expect_tx(full + s, (uint32_t)NOPS * (uint32_t)(g.a_rows + g.b_rows) * 128u);
      const int tap = kb / g.kc, c0 = (kb - tap * g.kc) * PJ_BLOCK_K;
      const int dy = (g.taps == 9) ? tap / 3 - 1 : 0, dx = (g.taps == 9) ? tap % 3 - 1 : 0;
#pragma unroll
      for (int o = 0; o < NOPS; ++o) {
        const CUtensorMap* ma = o ? &map_a_lo : &map_a_hi;
        const CUtensorMap* mb = o ? &map_b_lo : &map_b_hi;
        uint8_t* sa = st + o * (PJ_A_BYTES + PJ_B_BYTES);
        if (g.wgrad) {             // K block = 32 pixels of sample smp; rows = channels (pixels beyond P: zero fill)
          const int smp = kb / g.kp, p0 = (kb - smp * g.kp) * PJ_BLOCK_K;
          pj_tma_3d(sa, ma, full + s, p0, m_tile * PJ_BLOCK_M, smp);
          pj_tma_3d(sa + PJ_A_BYTES, mb, full + s, p0, n_tile * PJ_BLOCK_N, smp);
          continue;
        }
        if (g.flat) pj_tma_2d(sa, ma, full + s, c0, m_tile * PJ_BLOCK_M);
        else pj_tma_4d(sa, ma, full + s, c0, dx, h0 + dy, n0);
        pj_tma_2d(sa + PJ_A_BYTES, mb, full + s, tap * g.Cin + c0, n_tile * PJ_BLOCK_N);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(PJ_BLOCK_N >> 3) << 17) |
                           ((uint32_t)(PJ_BLOCK_M >> 4) << 24);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % PJ_STAGES;
      const uint32_t ph = (kb / PJ_STAGES) & 1;
      pj_mbar_wait(full + s, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = pj_smem(tiles + s * STAGE_BYTES);
      const uint64_t a_hi = pj_desc(sa), b_hi = pj_desc(sa + PJ_A_BYTES);
#pragma unroll
      for (int k = 0; k < PJ_BLOCK_K / PJ_UMMA_K; ++k) {
        const uint64_t adv = (uint64_t)((k * PJ_UMMA_K * 4) >> 4);
        if (SPLIT3) {
          const uint64_t a_lo = pj_desc(sa + PJ_A_BYTES + PJ_B_BYTES), b_lo = pj_desc(sa + 2 * PJ_A_BYTES + PJ_B_BYTES);
          const uint32_t acc = tmem_base + (uint32_t)((kb % NACC) * PJ_TMEM_COLS);
          const uint32_t first = (kb >= NACC || k != 0) ? 1u : 0u;      // the first k-block of each accumulator overwrites
          pj_mma_tf32(acc, a_lo + adv, b_hi + adv, idesc, first);
          pj_mma_tf32(acc, a_hi + adv, b_lo + adv, idesc, 1u);
          pj_mma_tf32(acc, a_hi + adv, b_hi + adv, idesc, 1u);
        } else {
          const uint32_t first = (kb | k) != 0;
          pj_mma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, first);
        }
      }
      pj_commit(empty + s);                       // frees the stage once these MMAs have read it
    }
    pj_commit(acc_full);                          // accumulator complete
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> y (NCHW) + staging tile for the BatchNorm partials =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int n = 0, p = 0;
    const bool valid = pj_row(g, m_tile, r, n, p);
    const int P = g.H * g.W;
    float* stage = reinterpret_cast<float*>(tiles);               // [128][PJ_BLOCK_N + 1], pipeline is drained
    pj_mbar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int co0 = n_tile * PJ_BLOCK_N;
    float* yrow = y + ((long long)n * g.Cout + co0) * P + p;
#pragma unroll 1
    for (int c = 0; c < PJ_BLOCK_N; c += 32) {
      uint32_t v[32];
      pj_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      if (SPLIT3) {
        const int used = num_kb < NACC ? num_kb : NACC;           // accumulators that received at least one k-block
        for (int a = 1; a < used; ++a) {
          uint32_t u[32];
          pj_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * PJ_TMEM_COLS + c), u);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float f = valid ? __uint_as_float(v[j]) : 0.f;
        stage[r * (PJ_BLOCK_N + 1) + c + j] = f;
        if (valid && co0 + c + j < g.Cout) yrow[(long long)(c + j) * P] = f;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, 128;" ::: "memory");               // the four epilogue warps
    if (part_mean != nullptr) {
      const int col = threadIdx.x - 128;                          // one channel per thread
      float cnt;
      if (g.flat) {
        const long long left = (long long)g.N * P - (long long)m_tile * PJ_BLOCK_M;
        cnt = (float)(left < PJ_BLOCK_M ? left : PJ_BLOCK_M);
      } else {
        int n0, h0;
        if (g.bn > 1) { n0 = m_tile * g.bn; h0 = 0; }
        else { n0 = m_tile / g.tiles_per_sample; h0 = (m_tile - n0 * g.tiles_per_sample) * g.bh; }
        const int vn = min(g.bn, g.N - n0), vh = min(g.bh, g.H - h0);
        cnt = (float)(vn * vh * g.W);
      }
      if (co0 + col < g.Cout) {
        float s = 0.f;
        for (int i = 0; i < PJ_BLOCK_M; ++i) s += stage[i * (PJ_BLOCK_N + 1) + col];   // padding rows hold 0
        const float mean = s / cnt;
        // padding rows would contribute mean^2 each: remove them in closed form
        float m2 = 0.f;
        for (int i = 0; i < PJ_BLOCK_M; ++i) {
          const float d = stage[i * (PJ_BLOCK_N + 1) + col] - mean;
          m2 = fmaf(d, d, m2);
        }
        m2 -= ((float)PJ_BLOCK_M - cnt) * mean * mean;
        part_mean[(long long)m_tile * g.Cout + co0 + col] = mean;
        part_m2[(long long)m_tile * g.Cout + co0 + col] = fmaxf(m2, 0.f);
      }
      if (n_tile == 0 && col == 0) part_cnt[m_tile] = cnt;
    }
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(PJ_TMEM_COLS * NACC));
  }
}

// ---- operand preparation ---------------------------------------------------------------------
__device__ __forceinline__ float pj_tf32_hi(float x) {        // round-to-nearest-even onto the TF32 grid
  uint32_t b = __float_as_uint(x);
  b += 0xFFFu + ((b >> 13) & 1u);
  return __uint_as_float(b & 0xFFFFE000u);
}

// x [N, C, P] (NCHW) -> hi (and lo) [N, P, C] (channels-last); 32x32 shared-memory tile transpose
__global__ void pj_prep_x_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int C, int P) {
  __shared__ float t[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* xn = x + (long long)n * C * P;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    t[i][threadIdx.x] = (c < C && p < P) ? xn[(long long)c * P + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < P && c < C) {
      const float v = t[threadIdx.x][i];
      const long long o = ((long long)n * P + p) * C + c;
      if (lo != nullptr) {
        const float h = pj_tf32_hi(v);
        hi[o] = h;
        lo[o] = pj_tf32_hi(v - h);     // exactly representable: the tensor core's truncation of lo is then a no-op
      } else {
        hi[o] = pj_tf32_hi(v);         // round to nearest (the tensor core itself would truncate)
      }
    }
  }
}

// w [Cout, Cin, T] -> hi (and lo) [Cout, T, Cin]
__global__ void pj_prep_w_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long long total,
                                 int Cin, int T) {
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const long long co = o / ((long long)T * Cin);
    const int rem = (int)(o - co * T * Cin);
    const int t = rem / Cin, ci = rem - t * Cin;
    const float v = __ldg(w + (co * Cin + ci) * T + t);
    if (lo != nullptr) {
      const float h = pj_tf32_hi(v);
      hi[o] = h;
      lo[o] = pj_tf32_hi(v - h);
    } else {
      hi[o] = pj_tf32_hi(v);
    }
  }
}

// Chan merge of the per-tile partials -> mean [C], m2 [C]  (same outputs as ud_bn_stats)
__global__ void pj_merge_kernel(const float* __restrict__ pm, const float* __restrict__ p2, const float* __restrict__ pc,
                                float* __restrict__ mean, float* __restrict__ m2, int tiles, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float n = 0.f, mu = 0.f, M2 = 0.f;
  for (int t = 0; t < tiles; ++t) {
    const float nb = pc[t];
    if (nb <= 0.f) continue;
    const float mb = pm[(long long)t * C + c], sb = p2[(long long)t * C + c];
    const float tot = n + nb, d = mb - mu;
    mu += d * (nb / tot);
    M2 += sb + d * d * (n * nb / tot);
    n = tot;
  }
  mean[c] = mu;
  m2[c] = M2;
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*PjEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PjEncodeFn pj_encode_fn() {
  static PjEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PjEncodeFn>(p);
  });
  return fn;
}

static int pj_make_map(CUtensorMap* m, const float* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                       const cuuint32_t* box) {
  PjEncodeFn enc = pj_encode_fn();
  UD_REQUIRE(enc != nullptr, UD_ERR_CUDA, "proj: cuTensorMapEncodeTiled is unavailable in this driver");
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims, strides_bytes, box,
                   ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UD_REQUIRE(r == CUDA_SUCCESS, UD_ERR_CUDA, "proj: cuTensorMapEncodeTiled failed (CUresult %d, rank %d)", (int)r, rank);
  return UD_OK;
}

static int pj_geometry(int N, int H, int W, int Cin, int Cout, int ksize, PjGeom* g) {
  UD_REQUIRE(N >= 1 && H >= 1 && W >= 1 && Cin >= 1 && Cout >= 1, UD_ERR_INVALID, "proj: bad shape");
  UD_REQUIRE(ksize == 1 || ksize == 3, UD_ERR_UNSUPPORTED, "proj: kernel size %d (1 or 3)", ksize);
  UD_REQUIRE(Cin % 4 == 0 && Cin >= PJ_BLOCK_K, UD_ERR_UNSUPPORTED,
             "proj: Cin=%d must be a multiple of 4 (16-byte TMA row pitch) and >= %d", Cin, PJ_BLOCK_K);
  g->N = N; g->H = H; g->W = W; g->Cin = Cin; g->Cout = Cout;
  g->taps = ksize * ksize;
  g->kc = ud_cdiv(Cin, PJ_BLOCK_K);
  g->n_tiles = ud_cdiv(Cout, PJ_BLOCK_N);
  g->flat = (ksize == 1);
  g->wgrad = 0;
  g->kp = 0;
  g->bh = g->bn = g->tiles_per_sample = 1;
  if (g->flat) {
    g->m_tiles = ud_cdiv((long long)N * H * W, PJ_BLOCK_M);
  } else {
    UD_REQUIRE(W <= PJ_BLOCK_M, UD_ERR_UNSUPPORTED, "proj: W=%d > %d", W, PJ_BLOCK_M);
    if (H * W <= PJ_BLOCK_M) {
      g->bh = H;
      g->bn = PJ_BLOCK_M / (H * W);
      if (g->bn > N) g->bn = N;
      g->m_tiles = ud_cdiv(N, g->bn);
    } else {
      g->bh = PJ_BLOCK_M / W;
      g->tiles_per_sample = ud_cdiv(H, g->bh);
      g->m_tiles = N * g->tiles_per_sample;
    }
  }
  UD_REQUIRE(g->m_tiles <= 65535, UD_ERR_UNSUPPORTED, "proj: too many M tiles (%d)", g->m_tiles);
  const long long Mtot = (long long)N * H * W;
  g->a_rows = g->flat ? (int)(Mtot < PJ_BLOCK_M ? Mtot : PJ_BLOCK_M) : g->bn * g->bh * W;
  g->b_rows = Cout < PJ_BLOCK_N ? Cout : PJ_BLOCK_N;
  return UD_OK;
}

extern "C" int ud_proj_m_tiles(int N, int H, int W, int ksize) {
  PjGeom g;
  if (pj_geometry(N, H, W, PJ_BLOCK_K, PJ_BLOCK_K, ksize, &g) != UD_OK) return -1;
  return g.m_tiles;
}

extern "C" int ud_proj_prep_x(const float* x_nchw, float* hi, float* lo, int N, int C, int P, cudaStream_t stream) {
  UD_REQUIRE(x_nchw && hi && N >= 1 && C >= 1 && P >= 1 && N <= 65535, UD_ERR_INVALID, "proj_prep_x: bad arguments");
  pj_prep_x_kernel<<<dim3(ud_cdiv(P, 32), ud_cdiv(C, 32), N), dim3(32, 8), 0, stream>>>(x_nchw, hi, lo, C, P);
  return ud_check_launch("pj_prep_x");
}

extern "C" int ud_proj_prep_w(const float* w, float* hi, float* lo, int Cout, int Cin, int taps, cudaStream_t stream) {
  UD_REQUIRE(w && hi && Cout >= 1 && Cin >= 1 && taps >= 1, UD_ERR_INVALID, "proj_prep_w: bad arguments");
  const long long total = (long long)Cout * Cin * taps;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  pj_prep_w_kernel<<<blocks, 256, 0, stream>>>(w, hi, lo, total, Cin, taps);
  return ud_check_launch("pj_prep_w");
}

extern "C" int ud_proj_fwd(const float* x_hi, const float* x_lo, const float* w_hi, const float* w_lo, float* y,
                           float* part_mean, float* part_m2, float* part_cnt, int N, int H, int W, int Cin, int Cout,
                           int ksize, cudaStream_t stream) {
  PjGeom g;
  int rc = pj_geometry(N, H, W, Cin, Cout, ksize, &g);
  if (rc != UD_OK) return rc;
  UD_REQUIRE(x_hi && w_hi && y, UD_ERR_INVALID, "proj_fwd: null pointer");
  UD_REQUIRE((x_lo == nullptr) == (w_lo == nullptr), UD_ERR_INVALID, "proj_fwd: x_lo and w_lo go together (3xTF32)");
  UD_REQUIRE((part_mean == nullptr) == (part_m2 == nullptr) && (part_mean == nullptr) == (part_cnt == nullptr),
             UD_ERR_INVALID, "proj_fwd: part_mean/part_m2/part_cnt go together");
  const bool split = x_lo != nullptr;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  for (int o = 0; o < (split ? 2 : 1); ++o) {
    const float* xa = o ? x_lo : x_hi;
    const float* wb = o ? w_lo : w_hi;
    CUtensorMap* ma = o ? &ma_lo : &ma_hi;
    CUtensorMap* mb = o ? &mb_lo : &mb_hi;
    if (g.flat) {
      cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)N * H * W};
      cuuint64_t str[1] = {(cuuint64_t)Cin * 4};
      cuuint32_t box[2] = {PJ_BLOCK_K, (cuuint32_t)g.a_rows};
      if ((rc = pj_make_map(ma, xa, 2, dims, str, box)) != UD_OK) return rc;
    } else {
      cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      cuuint64_t str[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)W * Cin * 4, (cuuint64_t)H * W * Cin * 4};
      cuuint32_t box[4] = {PJ_BLOCK_K, (cuuint32_t)W, (cuuint32_t)g.bh, (cuuint32_t)g.bn};
      if ((rc = pj_make_map(ma, xa, 4, dims, str, box)) != UD_OK) return rc;
    }
    cuuint64_t dims[2] = {(cuuint64_t)g.taps * Cin, (cuuint64_t)Cout};
    cuuint64_t str[1] = {(cuuint64_t)g.taps * Cin * 4};
    cuuint32_t box[2] = {PJ_BLOCK_K, (cuuint32_t)g.b_rows};
    if ((rc = pj_make_map(mb, wb, 2, dims, str, box)) != UD_OK) return rc;
  }
  if (!split) { ma_lo = ma_hi; mb_lo = mb_hi; }
  const size_t stage = (size_t)(split ? 2 : 1) * (PJ_A_BYTES + PJ_B_BYTES);
  const size_t smem = 1024 + PJ_STAGES * stage + 128;
  static_assert(PJ_STAGES * (PJ_A_BYTES + PJ_B_BYTES) >= PJ_BLOCK_M * (PJ_BLOCK_N + 1) * 4, "staging tile must fit the ring");
  // boxes smaller than 128 rows leave the tail of a tile unwritten (garbage rows/columns of D that the epilogue
  // masks): the expected transaction count is the box size (a_rows + b_rows rows of 128 bytes)
  dim3 grid(g.n_tiles, g.m_tiles);
  if (split) {
    UD_CUDA(cudaFuncSetAttribute(pj_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pj_gemm_kernel<true><<<grid, PJ_THREADS, smem, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, y, part_mean, part_m2, part_cnt, g);
  } else {
    UD_CUDA(cudaFuncSetAttribute(pj_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pj_gemm_kernel<false><<<grid, PJ_THREADS, smem, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, y, part_mean, part_m2, part_cnt, g);
  }
  return ud_check_launch("pj_gemm");
}

// ---- backward of the projections ------------------------------------------------------------------------
// elementwise hi/lo split (no re-layout): the NCHW operands of the weight-gradient GEMM in 3xTF32
__global__ void pj_split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(x + i), h = pj_tf32_hi(v);
    hi[i] = h;
    lo[i] = pj_tf32_hi(v - h);
  }
}

// w [Cout, Cin, k, k] -> wt [Cin, k*k, Cout] with the taps FLIPPED: the weights of the data-gradient convolution
// dX = conv(dY, flip(W)^T), an implicit GEMM of the same form as the forward one.
__global__ void pj_prep_wt_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long long total,
                                  int Cout, int Cin, int T) {
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const long long ci = o / ((long long)T * Cout);
    const int rem = (int)(o - ci * T * Cout);
    const int t = rem / Cout, co = rem - t * Cout;
    const float v = __ldg(w + ((long long)co * Cin + ci) * T + (T - 1 - t));
    if (lo != nullptr) {
      const float h = pj_tf32_hi(v);
      hi[o] = h;
      lo[o] = pj_tf32_hi(v - h);
    } else {
      hi[o] = pj_tf32_hi(v);
    }
  }
}

extern "C" int ud_proj_split(const float* x, float* hi, float* lo, long long total, cudaStream_t stream) {
  UD_REQUIRE(x && hi && lo && total >= 1, UD_ERR_INVALID, "proj_split: bad arguments");
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  pj_split_kernel<<<blocks, 256, 0, stream>>>(x, hi, lo, total);
  return ud_check_launch("pj_split");
}

extern "C" int ud_proj_prep_wt(const float* w, float* hi, float* lo, int Cout, int Cin, int taps, cudaStream_t stream) {
  UD_REQUIRE(w && hi && Cout >= 1 && Cin >= 1 && taps >= 1, UD_ERR_INVALID, "proj_prep_wt: bad arguments");
  const long long total = (long long)Cout * Cin * taps;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  pj_prep_wt_kernel<<<blocks, 256, 0, stream>>>(w, hi, lo, total, Cout, Cin, taps);
  return ud_check_launch("pj_prep_wt");
}

// Weight gradient of the 1x1 projection: dw [Cout, Cin] = sum_{n,p} dy[n,co,p] * x[n,ci,p], both operands NCHW
// ([N, C, P], P = h*w pixels, P % 4 == 0).  D'[ci, co] is accumulated over K = (sample, 32-pixel chunk) blocks fetched
// by 3-D TMA boxes {32 pixels, 128 channels, 1 sample} (pixels beyond P are zero-filled) and the epilogue's
// column-major store lands it as dw[co][ci].  x_lo / dy_lo: nullable 3xTF32 low parts (ud_proj_split).
extern "C" int ud_proj_wgrad_1x1(const float* x_hi, const float* x_lo, const float* dy_hi, const float* dy_lo, float* dw,
                                 int N, int P, int Cin, int Cout, cudaStream_t stream) {
  UD_REQUIRE(N >= 1 && P >= 1 && Cin >= 1 && Cout >= 1, UD_ERR_INVALID, "proj_wgrad: bad shape");
  UD_REQUIRE(P % 4 == 0 && P >= PJ_BLOCK_K, UD_ERR_UNSUPPORTED,
             "proj_wgrad: P=%d pixels per plane must be a multiple of 4 (TMA row pitch) and >= %d", P, PJ_BLOCK_K);
  UD_REQUIRE(x_hi && dy_hi && dw, UD_ERR_INVALID, "proj_wgrad: null pointer");
  UD_REQUIRE((x_lo == nullptr) == (dy_lo == nullptr), UD_ERR_INVALID, "proj_wgrad: x_lo and dy_lo go together (3xTF32)");
  PjGeom g;
  memset(&g, 0, sizeof(g));
  g.N = N; g.H = Cin; g.W = 1; g.Cin = P; g.Cout = Cout; g.taps = 1; g.flat = 0; g.wgrad = 1;
  g.bh = g.bn = g.tiles_per_sample = 1;
  g.kp = ud_cdiv(P, PJ_BLOCK_K);
  g.kc = g.kp;
  g.m_tiles = ud_cdiv(Cin, PJ_BLOCK_M);
  g.n_tiles = ud_cdiv(Cout, PJ_BLOCK_N);
  g.a_rows = Cin < PJ_BLOCK_M ? Cin : PJ_BLOCK_M;
  g.b_rows = Cout < PJ_BLOCK_N ? Cout : PJ_BLOCK_N;
  const bool split = x_lo != nullptr;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  for (int o = 0; o < (split ? 2 : 1); ++o) {
    cuuint64_t da[3] = {(cuuint64_t)P, (cuuint64_t)Cin, (cuuint64_t)N};
    cuuint64_t sa[2] = {(cuuint64_t)P * 4, (cuuint64_t)Cin * P * 4};
    cuuint32_t ba[3] = {PJ_BLOCK_K, (cuuint32_t)g.a_rows, 1};
    if ((rc = pj_make_map(o ? &ma_lo : &ma_hi, o ? x_lo : x_hi, 3, da, sa, ba)) != UD_OK) return rc;
    cuuint64_t db[3] = {(cuuint64_t)P, (cuuint64_t)Cout, (cuuint64_t)N};
    cuuint64_t sb[2] = {(cuuint64_t)P * 4, (cuuint64_t)Cout * P * 4};
    cuuint32_t bb[3] = {PJ_BLOCK_K, (cuuint32_t)g.b_rows, 1};
    if ((rc = pj_make_map(o ? &mb_lo : &mb_hi, o ? dy_lo : dy_hi, 3, db, sb, bb)) != UD_OK) return rc;
  }
  if (!split) { ma_lo = ma_hi; mb_lo = mb_hi; }
  const size_t stage = (size_t)(split ? 2 : 1) * (PJ_A_BYTES + PJ_B_BYTES);
  const size_t smem = 1024 + PJ_STAGES * stage + 128;
  dim3 grid(g.n_tiles, g.m_tiles);
  if (split) {
    UD_CUDA(cudaFuncSetAttribute(pj_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pj_gemm_kernel<true><<<grid, PJ_THREADS, smem, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, dw, nullptr, nullptr, nullptr, g);
  } else {
    UD_CUDA(cudaFuncSetAttribute(pj_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pj_gemm_kernel<false><<<grid, PJ_THREADS, smem, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, dw, nullptr, nullptr, nullptr, g);
  }
  return ud_check_launch("pj_wgrad");
}

extern "C" int ud_bn_merge_partials(const float* part_mean, const float* part_m2, const float* part_cnt, float* mean,
                                    float* m2, int tiles, int C, cudaStream_t stream) {
  UD_REQUIRE(part_mean && part_m2 && part_cnt && mean && m2 && tiles >= 1 && C >= 1, UD_ERR_INVALID,
             "bn_merge_partials: bad arguments");
  pj_merge_kernel<<<ud_cdiv(C, 128), 128, 0, stream>>>(part_mean, part_m2, part_cnt, mean, m2, tiles, C);
  return ud_check_launch("pj_merge");
}

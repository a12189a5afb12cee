// §8(e) -- the one exchange step of the path: SyncBatchNorm statistics across the ranks of one NVLink/NVSwitch box
//
// Reference: SyncBatchNorm.convert_sync_batchnorm (engine/forgery_engine.py:142) turns every BatchNorm -- the three on
// the hot path (freq_filter.layer1.1, spat_filter.layer1.1, bottleneck) and the 96 of the EfficientNet-B4 backbone --
// into two tiny collectives per layer and step: all_gather of [mean, invstd|M2, count] (2C+1 floats) forward,
// all_reduce of [sum dy, sum dy*xmu] (2C floats) backward.  198 dependent NCCL launches per step are pure latency.
//
// Here every rank owns one communication buffer in its HBM, mapped into all peers with CUDA IPC (NVLink P2P).  A
// gather is ONE single-CTA kernel per rank:
//     1. store my vector into slot[seq % 2][my_rank] of EVERY rank's buffer (remote stores over NVLink),
//     2. __threadfence_system(); publish flag[seq % 2][my_rank] = seq in every rank's buffer (st.release.sys),
//     3. spin (ld.acquire.sys, local HBM) until all `world` flags of my own buffer show `seq`,
//     4. copy the `world` vectors from my own buffer to the caller's tensor (or sum them: the all-reduce flavour).
// No host involvement, no NCCL: the kernels are plain launches and can be captured into a CUDA graph; the sequence
// number lives on the device so a replayed graph keeps counting.  Two slots suffice: a peer can be at most one
// gather ahead of me, because it cannot finish gather n+1 before I have published my part of n+1, which I only do
// after I have consumed gather n (stream order).
#include <mutex>

#include "../../include/unidefense_b200.h"
#include "ud_common.cuh"

#define CM_MAX_WORLD 16
#define CM_SPIN_LIMIT (1u << 22)      // a few seconds of system-scope polling, then give up loudly instead of hanging the GPU

struct UdCommDev {
  float* peer[CM_MAX_WORLD];          // base of every rank's buffer as mapped in THIS process (peer[rank] = my own)
  int rank, world, max_count;
  unsigned int* seq;                  // device counter of gathers issued so far
  unsigned int* error;                // set to 1 when a wait timed out
};

struct UdComm {
  UdCommDev d;
  void* own;
  size_t bytes;
};

// buffer layout (floats): slot s, rank r: data at ((s*world + r) * max_count), flags after all data
__device__ __forceinline__ size_t cm_data_off(const UdCommDev& c, int s, int r) {
  return ((size_t)s * c.world + r) * c.max_count;
}
__device__ __forceinline__ size_t cm_flag_off(const UdCommDev& c, int s, int r) {
  return (size_t)2 * c.world * c.max_count + (size_t)s * c.world + r;
}

__global__ void __launch_bounds__(256)
cm_gather_kernel(const UdCommDev c, const float* __restrict__ src, float* __restrict__ dst, int count, int reduce) {
  __shared__ unsigned int s_seq;
  if (threadIdx.x == 0) s_seq = *c.seq + 1;
  __syncthreads();
  const unsigned int seq = s_seq;
  const int s = seq & 1;
  // 1. scatter my vector to every rank (own buffer included)
  for (int q = 0; q < c.world; ++q) {
    float* d = c.peer[q] + cm_data_off(c, s, c.rank);
    for (int i = threadIdx.x; i < count; i += blockDim.x) d[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish; 3. wait for everyone
  if (threadIdx.x < c.world) {
    unsigned int* f = reinterpret_cast<unsigned int*>(c.peer[threadIdx.x] + cm_flag_off(c, s, c.rank));
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
    const unsigned int* mine = reinterpret_cast<const unsigned int*>(c.peer[c.rank] + cm_flag_off(c, s, threadIdx.x));
    unsigned int v = 0, spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    } while (v != seq && ++spins < CM_SPIN_LIMIT);
    if (v != seq) *c.error = 1;
  }
  __syncthreads();
  // 4. consume
  const float* base = c.peer[c.rank] + cm_data_off(c, s, 0);
  if (reduce) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
      float a = 0.f;
      for (int q = 0; q < c.world; ++q) a += __ldcg(base + (size_t)q * c.max_count + i);   // fixed rank order: deterministic
      dst[i] = a;
    }
  } else {
    for (int q = 0; q < c.world; ++q)
      for (int i = threadIdx.x; i < count; i += blockDim.x) dst[(size_t)q * count + i] = __ldcg(base + (size_t)q * c.max_count + i);
  }
  __syncthreads();
  if (threadIdx.x == 0) *c.seq = seq;
}

extern "C" size_t ud_comm_buffer_bytes(int world, int max_count) {
  if (world < 1 || world > CM_MAX_WORLD || max_count < 1) return 0;
  return ((size_t)2 * world * max_count + (size_t)2 * world + 64) * sizeof(float);
}

// cudaMalloc (IPC needs a cudaMalloc allocation, not the caching allocator's) + zero fill + IPC handle (64 bytes)
extern "C" int ud_comm_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out) {
  UD_REQUIRE(dev_ptr && ipc_handle_out && bytes > 0, UD_ERR_INVALID, "comm_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  UD_CUDA(cudaMalloc(dev_ptr, bytes));
  UD_CUDA(cudaMemset(*dev_ptr, 0, bytes));
  UD_CUDA(cudaDeviceSynchronize());
  UD_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(ipc_handle_out), *dev_ptr));
  return UD_OK;
}

extern "C" int ud_comm_open(const void* ipc_handle, void** peer_ptr) {
  UD_REQUIRE(ipc_handle && peer_ptr, UD_ERR_INVALID, "comm_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  UD_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return UD_OK;
}

// peers[q] = pointer to rank q's buffer in this process (peers[rank] = the pointer ud_comm_alloc returned)
extern "C" int ud_comm_create(void* const* peers, int rank, int world, int max_count, void** comm_out) {
  UD_REQUIRE(peers && comm_out && world >= 1 && world <= CM_MAX_WORLD && rank >= 0 && rank < world && max_count >= 1,
             UD_ERR_INVALID, "comm_create: bad arguments (world <= %d)", CM_MAX_WORLD);
  UdComm* c = new UdComm();
  for (int q = 0; q < world; ++q) c->d.peer[q] = static_cast<float*>(peers[q]);
  c->d.rank = rank;
  c->d.world = world;
  c->d.max_count = max_count;
  unsigned int* state = nullptr;
  UD_CUDA(cudaMalloc(&state, 2 * sizeof(unsigned int)));
  UD_CUDA(cudaMemset(state, 0, 2 * sizeof(unsigned int)));
  c->d.seq = state;
  c->d.error = state + 1;
  c->own = peers[rank];
  *comm_out = c;
  return UD_OK;
}

// dst [world, count] (reduce = 0) or [count] = sum over ranks (reduce = 1); every rank must make the same calls
extern "C" int ud_comm_gather(void* comm, const float* src, float* dst, int count, int reduce, cudaStream_t stream) {
  UD_REQUIRE(comm && src && dst, UD_ERR_INVALID, "comm_gather: null pointer");
  UdComm* c = static_cast<UdComm*>(comm);
  UD_REQUIRE(count >= 1 && count <= c->d.max_count, UD_ERR_INVALID, "comm_gather: count %d exceeds the buffer (%d)", count,
             c->d.max_count);
  cm_gather_kernel<<<1, 256, 0, stream>>>(c->d, src, dst, count, reduce);
  return ud_check_launch("cm_gather");
}

// 1 when a gather timed out waiting for a peer since creation (synchronises the device)
extern "C" int ud_comm_error(void* comm) {
  if (!comm) return -1;
  UdComm* c = static_cast<UdComm*>(comm);
  unsigned int e = 0;
  if (cudaMemcpy(&e, c->d.error, sizeof(e), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)e;
}

// K6p — exchange of the per-GPU top-k keys over NVLink peer memory, fused with the k-way merge.
//
// The distributed form of np.argsort(-scores)[:k] (Tool/rank_chunks_optimized.py:225) needs one
// exchange step: every GPU's B x k best keys must reach every GPU.  The payload is tiny (80 B for one
// query, 12.8 KB for BASELINE config 5), so the step is pure latency; an NCCL all-gather plus a separate
// merge launch costs ~50 us of a 0.38 ms single-query step on 8 GPUs.  Here each rank owns a small
// exchange buffer that every peer maps through CUDA IPC:
//
//   push kernel : stores this rank's keys straight into slot [rank] of EVERY rank's buffer (st.global on
//                 NVLink-mapped peer pointers), fences at system scope, and the last CTA raises
//                 flag[rank] = seq in every buffer;
//   merge kernel: one CTA per query spins (bounded) on its OWN buffer until all `world` flags show seq —
//                 the only cross-GPU wait, satisfied by remote push kernels that never depend on this
//                 GPU — then merges the `world` lists exactly like ss_topk_merge.
//
// Buffers are double-buffered by seq parity: a rank can be at most one search ahead of its slowest peer
// (to push search s+2 it must have merged s+1, which needs every peer's push of s+1).
#include <algorithm>

#include "ss_common.cuh"
#include "topk_merge.cuh"

namespace ss {

constexpr int kPeerMaxWorld = 32;
constexpr size_t kPeerFlagBytes = 256;   // [2][kPeerMaxWorld] uint32
constexpr size_t kPeerHeaderBytes = 512; // flags, then this rank's push ticket at +256
constexpr int kPeerMergeThreads = 512;
constexpr size_t kPeerStatusOffset = kPeerFlagBytes + 16;   // 0 = healthy, r + 1 = gave up waiting for rank r's push
constexpr size_t kPeerSeqOffset = kPeerFlagBytes + 24;      // auto-mode sequence counter of this rank

__host__ __device__ inline size_t peer_slot_elems(int max_queries, int k) { return static_cast<size_t>(max_queries) * k; }

struct PeerParams {
  const uint64_t* local_keys;  // [n_queries][k]
  int n_queries, k, rank, world, max_queries;
  uint32_t seq;                 // explicit sequence number, or 0: take it from *seq_counter
  uint32_t* seq_counter;        // device word holding the sequence number of the LAST finished search (auto mode: the
                                // launch parameters never change, so the pair of kernels can live in a CUDA graph)
  unsigned char* const* peers;  // device array: base of every rank's exchange buffer as mapped here
};

// 1, 2, ..., 0xFFFFFFFF, 2, 3, ...: never 0 (= "no push yet") and always of alternating parity (the double buffer)
__device__ __forceinline__ uint32_t next_seq(uint32_t last) { return last == 0xFFFFFFFFu ? 2u : last + 1u; }
__device__ __forceinline__ uint32_t current_seq(const PeerParams& p) {
  return p.seq != 0u ? p.seq : next_seq(*reinterpret_cast<volatile uint32_t*>(p.seq_counter));
}

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) peer_push_kernel(const PeerParams p) {
  const uint32_t seq = current_seq(p);
  const int parity = static_cast<int>(seq & 1u);
  const size_t slot = peer_slot_elems(p.max_queries, p.k);
  const size_t n = static_cast<size_t>(p.n_queries) * p.k;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint64_t key = p.local_keys[i];
    for (int dst = 0; dst < p.world; ++dst) {
      uint64_t* data = reinterpret_cast<uint64_t*>(p.peers[dst] + kPeerHeaderBytes);
      data[(static_cast<size_t>(parity) * p.world + p.rank) * slot + i] = key;
    }
  }
  __threadfence_system();  // this thread's peer stores are visible system-wide before the ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(p.peers[p.rank] + kPeerFlagBytes);
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      *ticket = 0u;
      __threadfence_system();
      for (int dst = 0; dst < p.world; ++dst) {
        uint32_t* flags = reinterpret_cast<uint32_t*>(p.peers[dst]);
        st_release_sys_u32(flags + parity * kPeerMaxWorld + p.rank, seq);
      }
    }
  }
}

__global__ void __launch_bounds__(kPeerMergeThreads) peer_merge_kernel(const PeerParams p, uint64_t* __restrict__ out_keys,
                                                                       float* __restrict__ out_scores,
                                                                       long long* __restrict__ out_indices) {
  extern __shared__ __align__(16) unsigned char peer_smem[];
  __shared__ uint64_t scratch[2];
  const uint32_t seq = current_seq(p);
  const int parity = static_cast<int>(seq & 1u);
  unsigned char* mine = p.peers[p.rank];
  if (threadIdx.x < p.world) {
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(mine) + parity * kPeerMaxWorld + threadIdx.x;
    uint32_t spins = 0;
    while (ld_acquire_sys_u32(flag) != seq) {
      __nanosleep(64);
      if (++spins > (1u << 26)) {
        // many seconds: a peer died or the ranks lost lock-step.  The context stays usable: the failure is recorded in the
        // buffer's status word (ss_peer_status) and this search's output must be discarded.
        atomicExch(reinterpret_cast<unsigned int*>(mine + kPeerStatusOffset), 1u + static_cast<unsigned int>(threadIdx.x));
        break;
      }
    }
  }
  __syncthreads();
  const int q = blockIdx.x;
  const size_t slot = peer_slot_elems(p.max_queries, p.k);
  const uint64_t* base = reinterpret_cast<const uint64_t*>(mine + kPeerHeaderBytes) + static_cast<size_t>(parity) * p.world * slot +
                         static_cast<size_t>(q) * p.k;
  uint64_t* surv = reinterpret_cast<uint64_t*>(peer_smem);
  uint64_t* sl = surv + kMergeSurvivorCap;
  for (int c = threadIdx.x; c < p.world * p.k; c += blockDim.x) {
    const int src = c / p.k, i = c - src * p.k;
    sl[c] = __ldcv(base + static_cast<size_t>(src) * slot + i);  // written by remote GPUs: never from a stale cache line
  }
  __syncthreads();
  MergeOut out;
  out.keys = out_keys ? out_keys + static_cast<size_t>(q) * p.k : nullptr;
  out.scores = out_scores ? out_scores + static_cast<size_t>(q) * p.k : nullptr;
  out.indices = out_indices ? out_indices + static_cast<size_t>(q) * p.k : nullptr;
  block_merge_lists(sl, p.world, p.k, p.k, p.k, out, scratch, surv);
  if (p.seq == 0u) {  // auto mode: the last CTA to finish publishes the sequence number for the next search of this stream
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int* ticket = reinterpret_cast<unsigned int*>(mine + kPeerFlagBytes + 8);
      if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
        *ticket = 0u;
        *reinterpret_cast<volatile uint32_t*>(p.seq_counter) = seq;
        __threadfence();
      }
    }
  }
}

}  // namespace ss

using namespace ss;

extern "C" size_t ss_peer_buffer_bytes(int world, int max_queries, int k) {
  if (world <= 0 || world > kPeerMaxWorld || max_queries <= 0 || k <= 0) return 0;
  return kPeerHeaderBytes + 2 * static_cast<size_t>(world) * peer_slot_elems(max_queries, k) * 8;
}

extern "C" int ss_peer_alloc(size_t bytes, void** dev_ptr_out, unsigned char* handle_out64) {
  if (!dev_ptr_out || !handle_out64 || bytes == 0) return fail(SS_ERR_INVALID_ARG, "ss_peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* ptr = nullptr;
  SS_CUDA_CHECK(cudaMalloc(&ptr, bytes));
  SS_CUDA_CHECK(cudaMemset(ptr, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) {
    cudaFree(ptr);
    return cuda_fail(e, "cudaIpcGetMemHandle");
  }
  memcpy(handle_out64, &h, 64);
  *dev_ptr_out = ptr;
  return SS_OK;
}

extern "C" int ss_peer_open(const unsigned char* handle64, void** dev_ptr_out) {
  if (!handle64 || !dev_ptr_out) return fail(SS_ERR_INVALID_ARG, "ss_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  SS_CUDA_CHECK(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return SS_OK;
}

extern "C" int ss_peer_close(void* dev_ptr) {
  if (dev_ptr) SS_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
  return SS_OK;
}

extern "C" int ss_peer_free(void* dev_ptr) {
  if (dev_ptr) SS_CUDA_CHECK(cudaFree(dev_ptr));
  return SS_OK;
}

static int peer_exchange_impl(const uint64_t* local_keys, int n_queries, int k, int rank, int world, void* const* peer_bases_device,
                              int max_queries, uint32_t seq, void* local_buffer, uint64_t* out_keys, float* out_scores,
                              int64_t* out_indices, void* stream) {
  if (!local_keys || !peer_bases_device) return fail(SS_ERR_INVALID_ARG, "ss_topk_peer_exchange_merge: null pointer");
  if (world <= 0 || world > kPeerMaxWorld || rank < 0 || rank >= world || n_queries <= 0 || k <= 0 || n_queries > max_queries)
    return fail(SS_ERR_INVALID_ARG, "ss_topk_peer_exchange_merge: bad sizes");
  if (seq == 0 && !local_buffer) return fail(SS_ERR_INVALID_ARG, "ss_topk_peer_exchange_merge: seq starts at 1");
  const size_t dyn = (static_cast<size_t>(kMergeSurvivorCap) + static_cast<size_t>(world) * k) * 8;
  if (dyn + 1024 > smem_optin()) return fail(SS_ERR_UNSUPPORTED, "ss_topk_peer_exchange_merge: world * k too large for shared memory");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PeerParams p;
  p.local_keys = local_keys;
  p.n_queries = n_queries;
  p.k = k;
  p.rank = rank;
  p.world = world;
  p.max_queries = max_queries;
  p.seq = seq;
  p.seq_counter = local_buffer ? reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(local_buffer) + kPeerSeqOffset) : nullptr;
  p.peers = reinterpret_cast<unsigned char* const*>(peer_bases_device);
  const size_t n = static_cast<size_t>(n_queries) * k;
  const int push_blocks = static_cast<int>(std::max<size_t>(1, std::min<size_t>((n + 255) / 256, 64)));
  peer_push_kernel<<<push_blocks, 256, 0, st>>>(p);
  SS_CUDA_CHECK(cudaGetLastError());
  if (dyn > 48 * 1024)
    SS_CUDA_CHECK(cudaFuncSetAttribute(peer_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn)));
  peer_merge_kernel<<<n_queries, kPeerMergeThreads, dyn, st>>>(p, out_keys, out_scores, reinterpret_cast<long long*>(out_indices));
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}

extern "C" int ss_topk_peer_exchange_merge(const uint64_t* local_keys, int n_queries, int k, int rank, int world,
                                           void* const* peer_bases_device, int max_queries, uint32_t seq, uint64_t* out_keys,
                                           float* out_scores, int64_t* out_indices, void* stream) {
  if (seq == 0) return fail(SS_ERR_INVALID_ARG, "ss_topk_peer_exchange_merge: seq starts at 1");
  return peer_exchange_impl(local_keys, n_queries, k, rank, world, peer_bases_device, max_queries, seq, nullptr, out_keys, out_scores,
                            out_indices, stream);
}

extern "C" int ss_topk_peer_exchange_merge_auto(const uint64_t* local_keys, int n_queries, int k, int rank, int world,
                                                void* const* peer_bases_device, int max_queries, void* local_buffer,
                                                uint64_t* out_keys, float* out_scores, int64_t* out_indices, void* stream) {
  if (!local_buffer) return fail(SS_ERR_INVALID_ARG, "ss_topk_peer_exchange_merge_auto: null local buffer");
  return peer_exchange_impl(local_keys, n_queries, k, rank, world, peer_bases_device, max_queries, 0u, local_buffer, out_keys,
                            out_scores, out_indices, stream);
}

extern "C" int ss_peer_status(const void* local_buffer, int* status_out_host) {
  if (!local_buffer || !status_out_host) return fail(SS_ERR_INVALID_ARG, "ss_peer_status: null pointer");
  unsigned int v = 0;
  SS_CUDA_CHECK(cudaMemcpy(&v, static_cast<const unsigned char*>(local_buffer) + kPeerStatusOffset, 4, cudaMemcpyDeviceToHost));
  *status_out_host = static_cast<int>(v);
  return SS_OK;
}

// Shared device/host helpers for libsemsearch_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>
#include <string>

#include "../../include/semsearch_b200.h"

namespace ss {

// ---------------------------------------------------------------------------------------------
// Host-side error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();
size_t smem_optin();

// RAII marker around the launch of a path's dominant kernel (see ss_profile_begin()).
class ProfileScope {
 public:
  explicit ProfileScope(cudaStream_t st);
  ~ProfileScope();

 private:
  cudaStream_t st_;
  int slot_;
};

#define SS_CUDA_CHECK(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::ss::cuda_fail(_e, #expr); \
  } while (0)

inline size_t dtype_size(int dt) { return dt == SS_F32 ? 4 : 2; }
inline bool dtype_ok(int dt) { return dt == SS_F32 || dt == SS_BF16 || dt == SS_F16; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// Packed top-k keys: (order-preserving fp32 bits << 32) | (0xFFFFFFFF - row index)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  uint32_t b;
  memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t global_row) {
  if (!(score == score)) score = -INFINITY;  // NaN ranks last, like np.argsort(-x)
  return (static_cast<uint64_t>(float_to_ordered(score)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - global_row);
}
__device__ __forceinline__ float key_score(uint64_t key) { return ordered_to_float(static_cast<uint32_t>(key >> 32)); }
__device__ __forceinline__ int64_t key_index(uint64_t key) {
  return static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull));
}

// ---------------------------------------------------------------------------------------------
// Element conversion: one 16-byte chunk -> fp32 lanes
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Chunk;  // EPC = elements per 16-byte chunk
template <>
struct Chunk<float> {
  static constexpr int EPC = 4;
  __device__ __forceinline__ static void unpack(const uint4& r, float* x) {
    x[0] = __uint_as_float(r.x);
    x[1] = __uint_as_float(r.y);
    x[2] = __uint_as_float(r.z);
    x[3] = __uint_as_float(r.w);
  }
};
template <>
struct Chunk<__nv_bfloat16> {
  static constexpr int EPC = 8;
  __device__ __forceinline__ static void unpack(const uint4& r, float* x) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[2 * i] = __uint_as_float(w[i] << 16);
      x[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
};
template <>
struct Chunk<__half> {
  static constexpr int EPC = 8;
  __device__ __forceinline__ static void unpack(const uint4& r, float* x) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      const float2 f = __half22float2(h);
      x[2 * i] = f.x;
      x[2 * i + 1] = f.y;
    }
  }
};

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// ---------------------------------------------------------------------------------------------
// Warp helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t other = __shfl_xor_sync(0xffffffffu, v, o);
    v = other < v ? other : v;
  }
  return v;
}

// ---------------------------------------------------------------------------------------------
// Packed-pair row math shared by the streaming kernels
// ---------------------------------------------------------------------------------------------
// Packed fp32 pair math (Blackwell FFMA2): halves the FMA instruction count of the row pass.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float sum2(unsigned long long v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}

// One 16-byte chunk -> EPC/2 packed fp32 pairs.
template <typename T>
struct Pairs;
template <>
struct Pairs<float> {
  static constexpr int NP = 2;
  __device__ __forceinline__ static void unpack(const uint4& r, unsigned long long* x) {
    x[0] = (static_cast<unsigned long long>(r.y) << 32) | r.x;
    x[1] = (static_cast<unsigned long long>(r.w) << 32) | r.z;
  }
};
template <>
struct Pairs<__nv_bfloat16> {
  static constexpr int NP = 4;
  __device__ __forceinline__ static void unpack(const uint4& r, unsigned long long* x) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = pack2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xFFFF0000u));
  }
};
template <>
struct Pairs<__half> {
  static constexpr int NP = 4;
  __device__ __forceinline__ static void unpack(const uint4& r, unsigned long long* x) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      x[i] = pack2(f.x, f.y);
    }
  }
};

constexpr int next_pow2(int v) { return v <= 1 ? 1 : (v <= 2 ? 2 : (v <= 4 ? 4 : (v <= 8 ? 8 : (v <= 16 ? 16 : 32)))); }
constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v / 2); }

// Transposing butterfly: NVP per-lane partial sums -> one fully reduced value per lane.
// Afterwards lane L holds value index (L >> (5 - log2 NVP)); ~NVP + log2(32/NVP) shuffles in
// total instead of 5 * NVP.
template <int NVP>
__device__ __forceinline__ float transpose_reduce(float (&v)[NVP], int lane) {
  int off = 16;
#pragma unroll
  for (int n = NVP; n > 1; n >>= 1) {
    const int half = n >> 1;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  float r = v[0];
#pragma unroll
  for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  return r;
}

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA engine, 1-D contiguous form)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (CUDA error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace ss

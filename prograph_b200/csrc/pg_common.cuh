// Shared helpers for the prograph_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/prograph_b200.h"

namespace pg {

// ---- error plumbing --------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch();

#define PG_CHECK_ARG(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) { pg::set_error(__VA_ARGS__); return PG_ERR_INVALID; } \
  } while (0)

#define PG_CUDA(call)                                                          \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      pg::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,       \
                    cudaGetErrorString(e__));                                  \
      return PG_ERR_CUDA;                                                      \
    }                                                                          \
  } while (0)

#define PG_LAUNCH_CHECK()                                                      \
  do {                                                                         \
    pg::count_launch();                                                        \
    cudaError_t e__ = cudaGetLastError();                                      \
    if (e__ != cudaSuccess) {                                                  \
      pg::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,   \
                    cudaGetErrorString(e__));                                  \
      return PG_ERR_CUDA;                                                      \
    }                                                                          \
  } while (0)

int num_sms();
// stream-ordered scratch for CUB; the pool keeps its memory between calls
cudaError_t temp_alloc(void** ptr, size_t bytes, cudaStream_t s);

// ---- geometry of the packed table -------------------------------------------
constexpr int kStreamRowPad = 512;  // packed tables are padded to a multiple of this many rows

__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- PTX: mbarrier + bulk async copy (TMA engine, 1-D) -----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ unsigned atom_add_acq_rel_cta(unsigned* addr, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(addr)), "r"(v) : "memory");
  return old;
}
// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- small device utilities ---------------------------------------------------
// i -> (i / d, i % d) without the ~100-instruction 64-bit division of the generic path: d is a
// power of two (shift >= 0) for every packed width up to 512 residues, and small indices divide in 32 bits.
__device__ __forceinline__ void split_index(long long i, int d, int shift, long long* q, int* r) {
  if (shift >= 0) {
    *q = i >> shift;
    *r = static_cast<int>(i & (d - 1));
  } else if (i < (1ll << 32)) {
    const unsigned u = static_cast<unsigned>(i);
    *q = u / static_cast<unsigned>(d);
    *r = static_cast<int>(u - static_cast<unsigned>(*q) * static_cast<unsigned>(d));
  } else {
    *q = i / d;
    *r = static_cast<int>(i - *q * d);
  }
}
__host__ __device__ inline int pow2_shift(int d) {
  if (d <= 0 || (d & (d - 1)) != 0) return -1;
  int s = 0;
  while ((1 << s) < d) ++s;
  return s;
}

__device__ __forceinline__ float sim_f32(int d) { return __fdiv_rn(1.0f, (float)(1 + d)); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace pg

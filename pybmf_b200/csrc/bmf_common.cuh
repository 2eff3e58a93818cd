// Shared host/device helpers for libbmf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pybmf_b200.h"

namespace bmf {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail_arg(const char* what);
int check_cuda(cudaError_t e, const char* where);
int num_sms();

#define BMF_REQUIRE(cond, what)          \
  do {                                   \
    if (!(cond)) return bmf::fail_arg(what); \
  } while (0)

#define BMF_LAUNCH_CHECK(where)                                   \
  do {                                                            \
    int _rc = bmf::check_cuda(cudaGetLastError(), where);         \
    if (_rc) return _rc;                                          \
  } while (0)

static inline cudaStream_t as_stream(bmf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------
#ifdef __CUDACC__
// 128-bit streaming load of two bit-words (rows are 16-byte aligned by contract)
__device__ __forceinline__ ulonglong2 ld_words2(const uint64_t* p) {
  return __ldg(reinterpret_cast<const ulonglong2*>(p));
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- mbarrier / bulk-copy (TMA) primitives shared by the tensor-core and the streaming kernels ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must fault the launch, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > (1ll << 33)) {
      printf("bmf: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar,
             parity);
      __trap();
    }
  }
}
// 1-D bulk copy global -> shared (TMA engine, no registers held while in flight); 16-byte granularity
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// One elected lane of a CONVERGED warp (elect.sync).  Warps that issue TMA / tcgen05 / bulk-copy instructions run their
// loops with the whole warp so that every operand is warp-uniform and lives in uniform registers; issuing from inside an
// `if (lane == 0)` region instead makes the compiler wrap every such instruction in an ELECT + R2UR.BROADCAST loop (7
// moves per MMA, 4 per bulk copy), which made the ISSUER the bottleneck of the FP4 GEMM (150 instead of 128 cycles per
// kind::mxf4 instruction).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ uint64_t warp_uniform64(uint64_t v) {
  return ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(v >> 32), 0) << 32) | __shfl_sync(0xffffffffu, (uint32_t)v, 0);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// The reference's row test (PyBMF/models/Asso.py:181 on top of PyBMF/utils/metrics.py:201):
//   s = (-w_fp) * FP + w_fn * TP   -- three separate IEEE operations, never fused.
__device__ __forceinline__ double cover_score_f64(double neg_w_fp, double w_fn, int fp, int tp) {
  return __dadd_rn(__dmul_rn(neg_w_fp, (double)fp), __dmul_rn(w_fn, (double)tp));
}
// use(i,j): does adding (P new true positives, N new false positives) strictly raise row i's score?
__device__ __forceinline__ bool row_uses(int wa, int wb, double neg_w_fp, double w_fn, int tp_old,
                                         int fp_old, int P, int N) {
  if (wa | wb) return (wb * P - wa * N) > 0;  // integer mode: exact by construction
  return cover_score_f64(neg_w_fp, w_fn, fp_old + N, tp_old + P) >
         cover_score_f64(neg_w_fp, w_fn, fp_old, tp_old);
}
#endif

}  // namespace bmf

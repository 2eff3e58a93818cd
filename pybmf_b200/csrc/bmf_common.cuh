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

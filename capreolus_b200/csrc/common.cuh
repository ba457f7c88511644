// Shared device/host helpers for the capr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "capr_b200.h"

namespace capr {

// ---- host-side error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define CAPR_CHECK_CUDA(expr)                                 \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return capr::cuda_fail(_e, #expr); \
  } while (0)

#define CAPR_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      capr::set_error(__VA_ARGS__);   \
      return (code);                  \
    }                                 \
  } while (0)

int sm_count();

// Profiling switches (CAPR_DEBUG_FLAGS, CAPR_ATTN_DEBUG, CAPR_PACRR_DEBUG: skip the pooling / the MMAs / the gathers, results
// invalid) exist only in the debug build of the library (libcapr_b200_dbg.so, -DCAPR_DEBUG_BUILD); in the product library
// every test of them folds to a compile-time 0.
#ifdef CAPR_DEBUG_BUILD
#define CAPR_DBG(expr) (expr)
#else
#define CAPR_DBG(expr) 0
#endif

// The kernels, cudaFuncSetAttribute, sm_count() and the occupancy queries all act on the CURRENT device, while a caller hands
// the C ABI a stream and pointers that belong to its tensors' device (a model on cuda:1 while cuda:0 is current).  Every entry
// point that enqueues work therefore makes the device that owns one of its device pointers (or the one stored in its handle)
// current for the duration of the call and restores the previous one on return.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const void* device_ptr);
  explicit DeviceGuard(int device);
  ~DeviceGuard();
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;

 private:
  void enter(int device);
};
int device_of(const void* device_ptr);  // -1: not a device pointer (or null)

// The row gathers of the KNRM-family producers.  .cg (default) keeps the rows out of L1; -DCAPR_GATHER_CA builds the A/B variant that
// lets L1 keep them (hot zipf rows; measured, see DESIGN.md "What limits the gather").
#ifdef CAPR_GATHER_CA
#define CAPR_GATHER_CP "cp.async.ca.shared.global"
#else
#define CAPR_GATHER_CP "cp.async.cg.shared.global"
#endif

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
// 16-byte copy, or -- live == false -- 16 bytes of zeros without touching global memory (src-size 0).  Used for <pad> / OOV
// tokens, which all map to table row 0: reading it like any other row makes every SM hammer one L2 line.
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool live) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(live ? 16u : 0u) : "memory");
}
// Predicated 16-byte copy (PTX predicate, no branch: a C++ `if` around the asm statement makes ptxas emit a divergent branch per
// copy, which halved the producer's issue rate -- round-2 A/B) with src-size `nbytes` (0 = zero-fill without a global read).
__device__ __forceinline__ void cp_async16_pred(uint32_t smem_dst, const void* gmem_src, uint32_t nbytes, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %3, 0;\n\t"
      "@p " CAPR_GATHER_CP " [%0], [%1], 16, %2;\n\t}\n" ::"r"(smem_dst),
      "l"(gmem_src), "r"(nbytes), "r"((uint32_t)pred)
      : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Packed fp32 FMA of sm_100 (FFMA2): {w,w} * b + c on both halves with one instruction.  With w coming from constant
// memory ptxas emits the scalar-broadcast uniform-register form (FFMA2 R, R.F32x2, UR.F32, R.F32x2).
__device__ __forceinline__ float2 fma2_bcast(float w, float2 b, float2 c) {
  const float2 ww = make_float2(w, w);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(rd)
      : "l"(*reinterpret_cast<const unsigned long long*>(&ww)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
        "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Token id (int64 in the reference layout) -> table row: ids <= 0 (pad, OOV) and ids >= V read the
// all-zero <pad> row 0, which makes their cosine exactly 0 (capreolus/reranker/common.py:149-153,180).
__device__ __forceinline__ int table_row(long long id, int V) { return (id > 0 && id < (long long)V) ? (int)id : 0; }
// ids are compared for the OOV exact match (common.py:155-158,179); keep them as int (they are small).
__device__ __forceinline__ int id_as_int(long long id) {
  return id > 2147483647LL ? 2147483647 : (id < -2147483647LL ? -2147483647 : (int)id);
}

}  // namespace capr

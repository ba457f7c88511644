// Micro-benchmark (debug library only): the L2 -> SM row-gather rate of the KNRM-family producer, with nothing else in the kernel.
//
// Same access pattern as simtc3::producer_loop: 4 warps per SM copy 128-row units of a bf16 (hi, lo) table, one 64-element K atom
// of one plane per 16 KB stage, with 16-byte cp.async (8 lanes per 128-byte row segment) into the SWIZZLE_128B layout and
// cp.async.mbarrier.arrive.noinc on the stage's barrier; a consumer warp frees every stage as soon as it has landed.  bench.py
// reports the product kernel's gather traffic against this rate as a second roofline (`roofline.l2_gather`): it is the ceiling of
// THIS access pattern on THIS machine (zipf row ids over a 36 MB L2-resident table), which no amount of overlap can beat.
#include "common.cuh"
#include "tc_common.cuh"

namespace capr {

constexpr int GB_PROD_THREADS = 128;
constexpr int GB_THREADS = GB_PROD_THREADS + 32;
constexpr int GB_STAGE_BYTES = 128 * 128;
constexpr int GB_MAX_STAGES = 13;

__global__ void __launch_bounds__(GB_THREADS, 1) gather_bench_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int pitch,
                                                                     const int* __restrict__ rows, int n_units, int n_stages) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* ring = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[GB_MAX_STAGES], empty[GB_MAX_STAGES];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int atoms = (pitch + 63) / 64;
  const int last_chunks = (pitch - (atoms - 1) * 64) / 8;
  if (tid == 0) {
    for (int i = 0; i < n_stages; ++i) tc::mbar_init(&full[i], GB_PROD_THREADS), tc::mbar_init(&empty[i], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (warp < 4) {
    const int sub = tid & 7, rsub = tid >> 3;
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      unsigned off[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) off[j] = (unsigned)rows[(size_t)u * 128 + rsub + 16 * j] * (unsigned)pitch + (unsigned)(sub * 8);
      for (int a = 0; a < atoms; ++a) {
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
          const __nv_bfloat16* tab = (plane == 0 ? hi : lo) + a * 64;
          tc::mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t base = tc::smem_u32(ring + stage * GB_STAGE_BYTES);
          if (a + 1 < atoms || sub < last_chunks) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int r = rsub + 16 * j;
              asm volatile(CAPR_GATHER_CP " [%0], [%1], 16;" ::"r"(base + r * 128 + ((sub ^ (r & 7)) << 4)), "l"(tab + off[j]) : "memory");
            }
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(&full[stage])) : "memory");
          if (++stage == n_stages) stage = 0, phase ^= 1;
        }
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
  } else {
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x)
      for (int i = 0; i < 2 * atoms; ++i) {
        tc::mbar_wait(&full[stage], phase);
        __syncwarp();
        if ((tid & 31) == 0) tc::mbar_arrive(&empty[stage]);
        if (++stage == n_stages) stage = 0, phase ^= 1;
      }
  }
}


// Variant for the hand-off question: WARP-OWNED stages.  Stage i (global counter over units, K atoms and
// planes) is filled by producer warp i % 4 alone -- 32 lanes x 32 copies, full-barrier count 32 instead of 128 -- and with a ring
// size that is a multiple of 4 a warp only ever revisits its own slots, so its phase bookkeeping cannot alias.  mode bit 1: no
// copies at all (the arrive / wait / free skeleton only).
__global__ void __launch_bounds__(GB_THREADS, 1) gather_bench_owned_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int pitch,
                                                                           const int* __restrict__ rows, int n_units, int n_stages, int skip, int lockstep) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* ring = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[16], empty[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int atoms = (pitch + 63) / 64;
  const int last_chunks = (pitch - (atoms - 1) * 64) / 8;
  if (tid == 0) {
    for (int i = 0; i < n_stages; ++i) tc::mbar_init(&full[i], lockstep ? GB_PROD_THREADS : 32), tc::mbar_init(&empty[i], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  const int per_unit = 2 * atoms;
  if (warp < 4) {
    if (lockstep) {  // the product kernels' scheme, for an A/B inside the same kernel: every warp takes 32 rows of every stage
      const int sub = tid & 7, rsub = tid >> 3;
      int i = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        unsigned off[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) off[j] = (unsigned)rows[(size_t)u * 128 + rsub + 16 * j] * (unsigned)pitch + (unsigned)(sub * 8);
        for (int s2 = 0; s2 < per_unit; ++s2, ++i) {
          const int a = s2 >> 1, slot = i % n_stages;
          const __nv_bfloat16* tab = ((s2 & 1) == 0 ? hi : lo) + a * 64;
          tc::mbar_wait(&empty[slot], (uint32_t)(((i / n_stages) & 1) ^ 1));
          const uint32_t base = tc::smem_u32(ring + slot * GB_STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = rsub + 16 * j;
            cp_async16_pred(base + r * 128 + ((sub ^ (r & 7)) << 4), tab + off[j], 16u, !skip && (a + 1 < atoms || sub < last_chunks));
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(&full[slot])) : "memory");
        }
      }
    } else {
      const int sub = lane & 7, rsub = lane >> 3;  // 4 rows per pass, 32 passes
      int base_i = 0;                              // global index of the unit's first stage
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, base_i += per_unit) {
        // the stages of this unit that belong to this warp: i = base_i + s2 with i % 4 == warp
        for (int s2 = ((warp - base_i) & 3); s2 < per_unit; s2 += 4) {
          const int i = base_i + s2, a = s2 >> 1, slot = i % n_stages;
          const __nv_bfloat16* tab = ((s2 & 1) == 0 ? hi : lo) + a * 64;
          tc::mbar_wait(&empty[slot], (uint32_t)(((i / n_stages) & 1) ^ 1));
          const uint32_t base = tc::smem_u32(ring + slot * GB_STAGE_BYTES);
          const bool on = !skip && (a + 1 < atoms || sub < last_chunks);
#pragma unroll 8
          for (int j = 0; j < 32; ++j) {
            const int r = rsub + 4 * j;
            const unsigned off = (unsigned)rows[(size_t)u * 128 + r] * (unsigned)pitch + (unsigned)(sub * 8);
            cp_async16_pred(base + r * 128 + ((sub ^ (r & 7)) << 4), tab + off, 16u, on);
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(&full[slot])) : "memory");
        }
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
  } else {
    int i = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x)
      for (int s2 = 0; s2 < per_unit; ++s2, ++i) {
        const int slot = i % n_stages;
        tc::mbar_wait(&full[slot], (uint32_t)((i / n_stages) & 1));
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[slot]);
      }
  }
}


// How does the gather rate scale with the number of PRODUCER WARPS?  Lock-step stages, PW warps (4, 8 or 16) share the 128 rows of a
// stage (128 / (4 PW) rows per thread).  If a warp's LDGSTS issue latency (~40 cycles per instruction in the traces) is the limit,
// the rate grows with PW; if an SM-wide unit is, it does not.
template <int PW>
__global__ void __launch_bounds__(PW * 32 + 32, 1) gather_bench_warps_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int pitch,
                                                                              const int* __restrict__ rows, int n_units, int n_stages) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* ring = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[16], empty[16];
  constexpr int RS = PW * 4;        // rows covered by one pass of the producer threads
  constexpr int RPT = 128 / RS;     // rows per thread
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int atoms = (pitch + 63) / 64;
  const int last_chunks = (pitch - (atoms - 1) * 64) / 8;
  if (tid == 0) {
    for (int i = 0; i < n_stages; ++i) tc::mbar_init(&full[i], PW * 32), tc::mbar_init(&empty[i], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  const int per_unit = 2 * atoms;
  if (warp < PW) {
    const int sub = tid & 7, rsub = tid >> 3;
    int i = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      unsigned off[RPT];
#pragma unroll
      for (int j = 0; j < RPT; ++j) off[j] = (unsigned)rows[(size_t)u * 128 + rsub + RS * j] * (unsigned)pitch + (unsigned)(sub * 8);
      for (int s2 = 0; s2 < per_unit; ++s2, ++i) {
        const int a = s2 >> 1, slot = i % n_stages;
        const __nv_bfloat16* tab = ((s2 & 1) == 0 ? hi : lo) + a * 64;
        tc::mbar_wait(&empty[slot], (uint32_t)(((i / n_stages) & 1) ^ 1));
        const uint32_t base = tc::smem_u32(ring + slot * GB_STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const int r = rsub + RS * j;
          cp_async16_pred(base + r * 128 + ((sub ^ (r & 7)) << 4), tab + off[j], 16u, a + 1 < atoms || sub < last_chunks);
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(&full[slot])) : "memory");
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
  } else {
    int i = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x)
      for (int s2 = 0; s2 < per_unit; ++s2, ++i) {
        const int slot = i % n_stages;
        tc::mbar_wait(&full[slot], (uint32_t)((i / n_stages) & 1));
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[slot]);
      }
  }
}

}  // namespace capr

using namespace capr;

// rows: [n_rows] int32 table rows in [0, V); n_rows is truncated to a multiple of 128.  stages: 16 KB stages in flight per SM (2..13).
extern "C" int capr_debug_gather_bench(const void* table_hi, const void* table_lo, int V, int pitch, const int* rows, int n_rows, int stages,
                                       capr_stream_t stream) {
  capr::DeviceGuard device_guard(table_hi);
  const char* fn = "capr_debug_gather_bench";
  CAPR_REQUIRE(table_hi && table_lo && rows, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(V > 0 && pitch > 0 && pitch % 16 == 0 && pitch <= 320 && n_rows >= 128 && stages >= 2 && stages <= GB_MAX_STAGES, CAPR_ERR_BAD_SHAPE,
               "%s: bad arguments V=%d pitch=%d n_rows=%d stages=%d", fn, V, pitch, n_rows, stages);
  CAPR_REQUIRE((long long)V * pitch < (1ll << 31), CAPR_ERR_UNSUPPORTED, "%s: table too large for 32-bit offsets", fn);
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  const size_t smem = 1024 + (size_t)stages * GB_STAGE_BYTES;
  CAPR_CHECK_CUDA(cudaFuncSetAttribute(gather_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_units = n_rows / 128;
  gather_bench_kernel<<<n_units < sms ? n_units : sms, GB_THREADS, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)table_hi, (const __nv_bfloat16*)table_lo, pitch, rows,
                                                                                               n_units, stages);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

// mode: bit 0 = warp-owned stages (else the lock-step scheme inside the same kernel), bit 1 = no copies (hand-off skeleton only).
// stages must be a multiple of 4 (4, 8 or 12).
extern "C" int capr_debug_gather_bench2(const void* table_hi, const void* table_lo, int V, int pitch, const int* rows, int n_rows, int stages, int mode,
                                        capr_stream_t stream) {
  capr::DeviceGuard device_guard(table_hi);
  const char* fn = "capr_debug_gather_bench2";
  CAPR_REQUIRE(table_hi && table_lo && rows, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(V > 0 && pitch > 0 && pitch % 16 == 0 && pitch <= 320 && n_rows >= 128 && stages >= 4 && stages <= 12 && stages % 4 == 0, CAPR_ERR_BAD_SHAPE,
               "%s: bad arguments V=%d pitch=%d n_rows=%d stages=%d", fn, V, pitch, n_rows, stages);
  CAPR_REQUIRE((long long)V * pitch < (1ll << 31), CAPR_ERR_UNSUPPORTED, "%s: table too large for 32-bit offsets", fn);
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  const size_t smem = 1024 + (size_t)stages * GB_STAGE_BYTES;
  CAPR_CHECK_CUDA(cudaFuncSetAttribute(gather_bench_owned_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_units = n_rows / 128;
  const char* grid_env = getenv("CAPR_GB_GRID");  // fewer CTAs than SMs: is the rate limited per SM or by the shared L2 / fabric?
  const int want = grid_env ? atoi(grid_env) : sms;
  const int grid = want > 0 && want < sms ? want : sms;
  gather_bench_owned_kernel<<<n_units < grid ? n_units : grid, GB_THREADS, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)table_hi, (const __nv_bfloat16*)table_lo, pitch,
                                                                                                     rows, n_units, stages, (mode >> 1) & 1, (mode & 1) ? 0 : 1);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

// prod_warps: 4, 8 or 16 producer warps sharing every stage (lock-step); stages: 2..12 x 16 KB.
extern "C" int capr_debug_gather_bench3(const void* table_hi, const void* table_lo, int V, int pitch, const int* rows, int n_rows, int stages, int prod_warps,
                                        capr_stream_t stream) {
  capr::DeviceGuard device_guard(table_hi);
  const char* fn = "capr_debug_gather_bench3";
  CAPR_REQUIRE(table_hi && table_lo && rows, CAPR_ERR_BAD_POINTER, "%s: null pointer", fn);
  CAPR_REQUIRE(V > 0 && pitch > 0 && pitch % 16 == 0 && pitch <= 320 && n_rows >= 128 && stages >= 2 && stages <= 12, CAPR_ERR_BAD_SHAPE,
               "%s: bad arguments V=%d pitch=%d n_rows=%d stages=%d", fn, V, pitch, n_rows, stages);
  CAPR_REQUIRE(prod_warps == 4 || prod_warps == 8 || prod_warps == 16, CAPR_ERR_BAD_SHAPE, "%s: prod_warps=%d (4, 8 or 16)", fn, prod_warps);
  CAPR_REQUIRE((long long)V * pitch < (1ll << 31), CAPR_ERR_UNSUPPORTED, "%s: table too large for 32-bit offsets", fn);
  const int sms = sm_count();
  CAPR_REQUIRE(sms > 0, CAPR_ERR_NO_DEVICE, "%s: no CUDA device", fn);
  const size_t smem = 1024 + (size_t)stages * GB_STAGE_BYTES;
  const int n_units = n_rows / 128;
  const int grid = n_units < sms ? n_units : sms;
  cudaStream_t st = (cudaStream_t)stream;
  auto launch = [&](auto kernel, int threads) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, threads, smem, st>>>((const __nv_bfloat16*)table_hi, (const __nv_bfloat16*)table_lo, pitch, rows, n_units, stages);
    return cudaGetLastError();
  };
  if (prod_warps == 4) CAPR_CHECK_CUDA(launch(gather_bench_warps_kernel<4>, 4 * 32 + 32));
  else if (prod_warps == 8) CAPR_CHECK_CUDA(launch(gather_bench_warps_kernel<8>, 8 * 32 + 32));
  else CAPR_CHECK_CUDA(launch(gather_bench_warps_kernel<16>, 16 * 32 + 32));
  return CAPR_OK;
}

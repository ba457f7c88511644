// Micro-benchmark (debug entry, not on any product path): cycles per tcgen05.mma for a given shape, issued back to
// back by one thread with shared-memory operands, to size the tile shapes of simtc.cuh / bert_attn.cuh.
#include "common.cuh"
#include "tc_common.cuh"

namespace capr {

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int M, int N, int n_mma, int n_acc, int reps, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    // warp-uniform loop, one elected lane issues (the pattern of simtc::mma_loop): a divergent `if (lane == 0)` issuer
    // measures ~150 cycles per instruction of issue overhead instead of the tensor pipe
    const uint32_t idesc = tc::make_instr_desc(tc::FMT_BF16, M, N);
    const uint32_t a = tc::smem_u32(smem), b = a + 16384;
    const uint64_t adesc = tc::make_sw128_kmajor_desc(a), bdesc = tc::make_sw128_kmajor_desc(b);
    long long best = 1ll << 60;
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (tc::elect_one()) {
        for (int i = 0; i < n_mma; ++i) {
          const uint64_t k = (uint64_t)((i & 3) * 2);
          tc::umma_f16(tmem + (uint32_t)((i % n_acc) * N), adesc + k, bdesc + k, idesc, i >= n_acc);
        }
        tc::umma_commit(&bar);
      }
      __syncwarp();
      tc::mbar_wait(&bar, phase);
      phase ^= 1;
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = best;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

}  // namespace capr

// cycles[grid]: best-of-reps cycles for n_mma back-to-back MMAs of shape M x N x 16 (bf16), cycling over n_acc accumulators.
extern "C" int capr_debug_mma_bench(int M, int N, int n_mma, int n_acc, int reps, int grid, long long* cycles, capr_stream_t stream) {
  CAPR_REQUIRE((M == 64 || M == 128) && N >= 16 && N <= 256 && N % 16 == 0 && n_acc >= 1 && n_acc * N <= 512 && grid > 0 && cycles, CAPR_ERR_BAD_SHAPE, "capr_debug_mma_bench: bad arguments");
  const size_t smem = 1024 + 16384 + 32768;
  CAPR_CHECK_CUDA(cudaFuncSetAttribute(capr::mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  capr::mma_bench_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(M, N, n_mma, n_acc, reps, cycles);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

// ---- FFMA2 issue-rate micro-benchmark (debug entry): does the packed fp32 FMA run at the same rate when its scalar operand
// comes from a uniform register / constant bank (the form PACRR's conv uses) as with a vector-register pair?
namespace capr {
__constant__ float c_ffma2_taps[64];

template <int MODE>  // 0: constant-bank scalar broadcast (fma2_bcast), 1: vector-register {w,w} pairs, 2: plain FFMA (two per pair),
                     // 3: mode 0 + FMNMX in PACRR's ratio (24 filter-max operations per 56 FFMA2: does the max share the FMA pipe's slots?)
__global__ void __launch_bounds__(256) ffma2_bench_kernel(int iters, const float* __restrict__ gw, float* out, long long* cycles) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, 1.f);
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float2(1.0f + i * 1e-3f, 1.0f - i * 1e-3f);
  float wreg[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) wreg[t] = gw[t];
  float best[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 16; ++t) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0 || MODE == 3) {
          acc[i] = fma2_bcast(c_ffma2_taps[t], x[i], acc[i]);
          if (MODE == 3 && (i < 3 || (i == 3 && (t & 1) == 0))) best[i] = fmaxf(best[i], acc[(i + 4) & 7].y);  // 56 FMNMX per 128 FFMA2
        } else if (MODE == 1) {
          const float2 ww = make_float2(wreg[t], wreg[t]);
          unsigned long long rd;
          asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd)
              : "l"(*reinterpret_cast<const unsigned long long*>(&ww)), "l"(*reinterpret_cast<const unsigned long long*>(&x[i])),
                "l"(*reinterpret_cast<const unsigned long long*>(&acc[i])));
          acc[i] = *reinterpret_cast<float2*>(&rd);
        } else {
          acc[i].x = fmaf(wreg[t], x[i].x, acc[i].x);
          acc[i].y = fmaf(wreg[t], x[i].y, acc[i].y);
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  if (MODE == 3) s += (best[0] + best[1]) + (best[2] + best[3]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
}  // namespace capr

// cycles[grid]: SM cycles for iters x 16 taps x 8 packed FMAs per thread, 256 threads per CTA (8 warps, 2 per scheduler).
extern "C" int capr_debug_ffma2_bench(int mode, int iters, int grid, float* scratch, long long* cycles, capr_stream_t stream) {
  CAPR_REQUIRE(mode >= 0 && mode <= 3 && iters > 0 && grid > 0 && scratch && cycles, CAPR_ERR_BAD_SHAPE, "capr_debug_ffma2_bench: bad arguments");
  float h[64];
  for (int i = 0; i < 64; ++i) h[i] = 1.0f + 1e-4f * i;
  CAPR_CHECK_CUDA(cudaMemcpyToSymbolAsync(capr::c_ffma2_taps, h, sizeof(h), 0, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  CAPR_CHECK_CUDA(cudaMemcpyAsync(scratch, h, sizeof(h), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  if (mode == 0) capr::ffma2_bench_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(iters, scratch, scratch + 64, cycles);
  else if (mode == 1) capr::ffma2_bench_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(iters, scratch, scratch + 64, cycles);
  else if (mode == 3) capr::ffma2_bench_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(iters, scratch, scratch + 64, cycles);
  else capr::ffma2_bench_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(iters, scratch, scratch + 64, cycles);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

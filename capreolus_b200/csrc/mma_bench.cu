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

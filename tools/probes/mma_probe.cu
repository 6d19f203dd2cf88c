// How fast does one SM execute tcgen05.mma.kind::tf32 (M = 128, K = 8 per instruction) as a function of N, with both
// operands in shared memory (SS) and with A in tensor memory (TS)?  One CTA per SM issues `iters` back-to-back MMAs into
// one accumulator from (uninitialised) operands, commits, waits; cycles per MMA = clock64 difference / iters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I ../../distributed-drl_b200/csrc -I ../../include -o mma_probe mma_probe.cu
#include <cstdio>
#include "sac_gemm_tc.cuh"
using namespace ddrl::tc;
__global__ void __launch_bounds__(128, 1) probe(int n, int ts, int b_mn, int iters, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { bar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(&slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(0, b_mn, n);
    const uint64_t da = make_sdesc(s_addr(base), false), db = make_sdesc(s_addr(base) + 2 * TILE_BYTES, b_mn != 0);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint64_t off = (uint64_t)(((i & 3) * (b_mn ? 1024 : 32)) >> 4);
        if (ts) mma_tf32_ts(tmem + 256, tmem + (uint32_t)((i & 3) * 8), db + off, idesc, 1u);
        else mma_tf32(tmem + 256, da + (uint64_t)(((i & 3) * 32) >> 4), db + off, idesc, 1u);
      }
      mma_commit(&bar);
    }
    __syncwarp();
    bar_wait(&bar, 0);
    t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  printf("tcgen05.mma kind::tf32, M = 128, K = 8 per instruction; cycles per MMA (floor at the 1.1 PF/s tf32 peak: N / 2)\n");
  for (int grid : {1, 148})
    for (int ts = 0; ts < 2; ++ts)
      for (int b_mn = 0; b_mn < 2; ++b_mn)
        for (int n : {64, 128, 256}) {
          if (b_mn && n > 128 && 0) continue;
          probe<<<grid, 128, smem>>>(n, ts, b_mn, iters, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("CTAs %3d  A from %s  B %s-major  N %3d : %6.1f cycles per MMA  (%s)\n", grid, ts ? "TMEM" : "smem", b_mn ? "MN" : "K ", n,
                 (double)c / iters, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
          if (e != cudaSuccess) return 1;
        }
  return 0;
}

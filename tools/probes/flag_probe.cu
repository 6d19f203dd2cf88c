// One-way latency of a cross-GPU flag over NVLink peer memory, as the data-parallel exchange uses it (2 GPUs, one process,
// peer access enabled).  GPU0 and GPU1 each run a kernel that, `iters` times: (optionally) rewrites a 0.9 MB local buffer
// from all its CTAs, publishes flag = it on the PEER with the chosen fence, then waits for the peer's flag = it.  The time
// per iteration is one exchange round; variants isolate the fence and the preceding writes.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void __launch_bounds__(256) ping(int variant, int iters, int nwrite4, float4* local_buf, volatile unsigned int* my_flag, unsigned int* peer_flag,
                     unsigned int* ticket, unsigned int* go, unsigned int* relay) {
  __shared__ bool s_last;
  for (int it = 1; it <= iters; ++it) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nwrite4; i += gridDim.x * blockDim.x) local_buf[i] = make_float4(it, it, it, it);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == (unsigned)(gridDim.x * it - 1);
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
      if (variant == 0) { __threadfence_system(); asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag), "r"((unsigned)it) : "memory"); }
      else if (variant == 1) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag), "r"((unsigned)it) : "memory"); }
      else if (variant == 2) { __threadfence_system(); asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(peer_flag), "r"((unsigned)it) : "memory"); }
      else if (variant == 3) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(peer_flag), "r"((unsigned)it) : "memory"); }
      else { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag), "r"((unsigned)it) : "memory"); }
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(go), "r"((unsigned)it) : "memory");
    }
    if (variant < 4) {
      // every CTA waits for the peer's flag itself (thread 0 polls at system scope, like the first wait_flags)
      if (threadIdx.x == 0) {
        unsigned int v;
        do { asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flag) : "memory"); } while (v < (unsigned)it);
        asm volatile("fence.acq_rel.sys;" ::: "memory");
      }
    } else {
      // ONE thread of the grid polls at system scope and re-publishes at GPU scope; every other CTA waits on that local word
      if (threadIdx.x == 0) {
        unsigned int v;
        if (blockIdx.x == 0) {
          do { asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flag) : "memory"); } while (v < (unsigned)it);
          asm volatile("fence.acq_rel.sys;" ::: "memory");
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(relay), "r"((unsigned)it) : "memory");
        } else {
          do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(relay) : "memory"); } while (v < (unsigned)it);
        }
      }
    }
    __syncthreads();
  }
}
int main() {
  int n = 0; CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  float4* buf[2]; unsigned int *flag[2], *ticket[2], *go[2], *relay[2]; cudaStream_t st[2]; cudaEvent_t e0[2], e1[2];
  for (int d = 0; d < 2; ++d) {
    CK(cudaSetDevice(d)); CK(cudaDeviceEnablePeerAccess(1 - d, 0));
    CK(cudaMalloc(&buf[d], 1 << 20)); CK(cudaMalloc(&flag[d], 256)); CK(cudaMalloc(&ticket[d], 4)); CK(cudaMalloc(&go[d], 4)); CK(cudaMalloc(&relay[d], 4));
    CK(cudaStreamCreate(&st[d])); CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d]));
  }
  const char* names[5] = {"threadfence_system + st.release.sys (the exchange's publish)", "st.release.sys alone", "threadfence_system + st.relaxed.sys", "st.relaxed.sys alone (no ordering: timing only)", "st.release.sys alone; ONE system-scope poller relays at GPU scope"};
  for (int grid : {1, 216}) for (int nw : {0, 55296}) for (int v = 0; v < 5; ++v) {
    const int iters = 2000;
    for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaMemset(flag[d], 0, 256)); CK(cudaMemset(ticket[d], 0, 4)); CK(cudaMemset(relay[d], 0, 4)); CK(cudaDeviceSynchronize()); }
    for (int d = 0; d < 2; ++d) {
      CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d]));
      ping<<<grid, 256, 0, st[d]>>>(v, iters, nw, buf[d], flag[d], flag[1 - d], ticket[d], go[d], relay[d]);
      CK(cudaEventRecord(e1[d], st[d]));
    }
    float ms[2];
    for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaEventSynchronize(e1[d])); CK(cudaEventElapsedTime(&ms[d], e0[d], e1[d])); }
    printf("grid %3d, %6d float4 rewritten per round, %-62s: %.2f us per exchange round\n", grid, nw, names[v], ms[0] * 1e3 / iters);
  }
  return 0;
}

// Does the time of back-to-back dependent launches on one stream grow smoothly with the kernel's duration, or in steps?
// (tools/probes, GPU box only: nvcc -arch=sm_100a -o tick_probe tick_probe.cu && ./tick_probe)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(long long ns, int* sink) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while ((long long)(t - t0) < ns);
  if (ns < 0) *sink = 1;
}
int main() {
  int* d; cudaMalloc(&d, 4);
  cudaStream_t s; cudaStreamCreate(&s);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int grid : {1, 128}) {
    for (long long ns = 0; ns <= 12000; ns += 500) {
      for (int i = 0; i < 20; ++i) spin<<<grid, 256, 0, s>>>(ns, d);
      cudaStreamSynchronize(s);
      cudaEventRecord(e0, s);
      const int reps = 200;
      for (int i = 0; i < reps; ++i) spin<<<grid, 256, 0, s>>>(ns, d);
      cudaEventRecord(e1, s);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("grid %3d spin %5lld ns: %.2f us per launch (overhead %.2f)\n", grid, ns, ms * 1e3 / reps, ms * 1e3 / reps - ns * 1e-3);
    }
  }
  // the same inside a CUDA graph of 10 dependent kernel nodes
  for (long long ns : {4000LL, 5000LL, 6000LL, 7000LL, 8000LL}) {
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < 10; ++i) spin<<<128, 256, 0, s>>>(ns, d);
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&ge, g, 0);
    for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, s);
    cudaStreamSynchronize(s);
    cudaEventRecord(e0, s);
    for (int i = 0; i < 50; ++i) cudaGraphLaunch(ge, s);
    cudaEventRecord(e1, s); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("graph of 10 x spin %lld ns: %.2f us per node\n", ns, ms * 1e3 / 500);
  }
  return 0;
}

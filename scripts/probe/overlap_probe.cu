// Probe: how does a short "test" kernel overlap the next long "scan" kernel?  (round 2, session 6)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o overlap_probe overlap_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(256, 3) scan_like(unsigned long long ns, int pdl_nowait, unsigned long long *sink) {
  extern __shared__ unsigned char sm[];
  const unsigned long long t0 = gtime();
  while (gtime() - t0 < ns) { if (sm[threadIdx.x] == 255 && ns == 1) sink[0] = t0; }
}
__global__ void __launch_bounds__(128) test_like(unsigned long long ns, int pdl, unsigned long long *sink) {
  if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
  const unsigned long long t0 = gtime();
  while (gtime() - t0 < ns) { if (ns == 1) sink[1] = t0; }
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
  const int N = 60, grid = 148 * 3 * 6;     // six waves of CTAs
  const unsigned long long nsA = 25000, nsT = 20000;   // 25 us per CTA -> ~150 us per scan_like; 20 us test_like
  unsigned long long *sink; CK(cudaMalloc(&sink, 64));
  CK(cudaFuncSetAttribute(scan_like, cudaFuncAttributeMaxDynamicSharedMemorySize, 43000));
  cudaStream_t A, B; int lo, hi; cudaDeviceGetStreamPriorityRange(&lo, &hi);
  CK(cudaStreamCreateWithFlags(&A, cudaStreamNonBlocking)); CK(cudaStreamCreateWithPriority(&B, cudaStreamNonBlocking, hi));
  cudaEvent_t t0, t1, ev[2][3], evd[2]; cudaEventCreate(&t0); cudaEventCreate(&t1);
  for (auto &e : ev) for (auto &x : e) cudaEventCreate(&x);
  for (auto &x : evd) cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
  float base = 0;
  for (int variant = 0; variant < 6; variant++) {
    for (int rep = 0; rep < 2; rep++) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(t0, A));
      for (int k = 0; k < N; k++) {
        const int s = k & 1;
        if (variant == 0) {                       // scans only
          scan_like<<<grid, 256, 43000, A>>>(nsA, 0, sink);
        } else if (variant == 1) {                // same stream: scan, test, scan, test
          scan_like<<<grid, 256, 43000, A>>>(nsA, 0, sink);
          test_like<<<25, 128, 0, A>>>(nsT, 0, sink);
        } else if (variant == 2) {                // the library's scheme: timing events in A, test on B behind an event
          cudaEventRecord(ev[s][0], A);
          scan_like<<<grid, 256, 43000, A>>>(nsA, 0, sink);
          cudaEventRecord(ev[s][1], A);
          cudaStreamWaitEvent(B, ev[s][1], 0);
          test_like<<<25, 128, 0, B>>>(nsT, 0, sink);
          cudaEventRecord(ev[s][2], B);
        } else if (variant == 3) {                // same, dependency through a no-timing event only
          scan_like<<<grid, 256, 43000, A>>>(nsA, 0, sink);
          cudaEventRecord(evd[s], A);
          cudaStreamWaitEvent(B, evd[s], 0);
          test_like<<<25, 128, 0, B>>>(nsT, 0, sink);
        } else if (variant == 4) {                // programmatic dependent launch in one stream: test waits, next scan does not
          scan_like<<<grid, 256, 43000, A>>>(nsA, 0, sink);
          cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(25); cfg.blockDim = dim3(128); cfg.stream = A;
          cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          CK(cudaLaunchKernelEx(&cfg, test_like, nsT, 1, sink));
          if (k + 1 < N) {
            cudaLaunchConfig_t c2 = {}; c2.gridDim = dim3(grid); c2.blockDim = dim3(256); c2.dynamicSmemBytes = 43000; c2.stream = A;
            c2.attrs = at; c2.numAttrs = 1;
            CK(cudaLaunchKernelEx(&c2, scan_like, nsA, 1, sink));
            k++;
            CK(cudaLaunchKernelEx(&cfg, test_like, nsT, 1, sink));
          }
        } else {                                  // variant 2 with the host confirming step k-1 after enqueuing step k (as the library does)
          cudaEventRecord(ev[s][0], A);
          scan_like<<<grid, 256, 43000, A>>>(nsA, 0, sink);
          cudaEventRecord(ev[s][1], A);
          cudaStreamWaitEvent(B, ev[s][1], 0);
          test_like<<<25, 128, 0, B>>>(nsT, 0, sink);
          cudaEventRecord(ev[s][2], B);
          if (k > 0) cudaEventSynchronize(ev[1 - s][2]);
        }
      }
      CK(cudaEventRecord(evd[0], B)); CK(cudaStreamWaitEvent(A, evd[0], 0));
      CK(cudaEventRecord(t1, A)); CK(cudaEventSynchronize(t1));
      float ms; cudaEventElapsedTime(&ms, t0, t1);
      if (variant == 0) base = ms / N;
      if (rep == 1) printf("variant %d: %.1f us per step (scan alone %.1f us): +%.1f us\n", variant, 1e3 * ms / N, 1e3 * base, 1e3 * (ms / N - base));
    }
  }
  return 0;
}

// Kernel-to-kernel dependency latency on this GPU, measured with %globaltimer: the time from
// the last instruction of kernel A's only CTA to the first instruction of dependent kernel B,
// in stream order and inside a CUDA graph, for (a) tiny kernels, (b) B with 4 KB of kernel
// parameters, (c) B with 96 KB dynamic shared memory after an A with none (carve-out change),
// (d) both with the maximum carve-out preferred, (e) A with 32 x 128-thread CTAs.
// Diagnostic for DESIGN.md section 7 (where a half-step's time goes); not part of the library.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct Big { double pad[500]; };

__global__ void kA(unsigned long long* tl, int it) {
  if (threadIdx.x == 0) atomicMax(&tl[2 * it], now_ns());
}
__global__ void kB(unsigned long long* tl, int it) {
  extern __shared__ double sm[];
  if (threadIdx.x == 0 && blockIdx.x == 0) tl[2 * it + 1] = now_ns();
  if (tl == nullptr) sm[threadIdx.x] = 0;
}
__global__ void kBbig(unsigned long long* tl, int it, const __grid_constant__ Big b) {
  if (threadIdx.x == 0 && blockIdx.x == 0) tl[2 * it + 1] = now_ns();
  if (tl == nullptr) tl[0] = (unsigned long long)b.pad[it];
}

static double median(std::vector<double> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }

int main() {
  const int n = 200;
  unsigned long long* tl;
  cudaMalloc(&tl, 2 * n * 8);
  std::vector<unsigned long long> h(2 * n);
  cudaStream_t st;
  cudaStreamCreate(&st);
  Big big{};
  cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int variant = 0; variant < 5; ++variant) {
    for (int use_graph = 0; use_graph < 2; ++use_graph) {
      int carve = variant == 3 ? 100 : -1;
      cudaFuncSetAttribute(kA, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      cudaFuncSetAttribute(kB, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      cudaMemset(tl, 0, 2 * n * 8);
      size_t smemB = (variant == 2 || variant == 3) ? 96 * 1024 : 0;
      int gridA = variant == 4 ? 32 : 1;
      auto enqueue = [&](int it) {
        kA<<<gridA, 128, 0, st>>>(tl, it);
        if (variant == 1) kBbig<<<1, 128, 0, st>>>(tl, it, big);
        else kB<<<variant == 4 ? 128 : 1, 256, smemB, st>>>(tl, it);
      };
      if (!use_graph) {
        for (int it = 0; it < n; ++it) enqueue(it);
      } else {
        // graphs of 2 pairs each (like one ensemble step: two half-steps)
        for (int it = 0; it < n; it += 2) {
          cudaGraph_t g; cudaGraphExec_t ge;
          cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
          enqueue(it); enqueue(it + 1);
          cudaStreamEndCapture(st, &g);
          cudaGraphInstantiate(&ge, g, 0);
          cudaGraphLaunch(ge, st);
          cudaStreamSynchronize(st);
          cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
        }
      }
      cudaStreamSynchronize(st);
      cudaMemcpy(h.data(), tl, 2 * n * 8, cudaMemcpyDeviceToHost);
      std::vector<double> ab, ba;
      for (int it = 8; it < n; ++it) ab.push_back((double)(h[2 * it + 1] - h[2 * it]) / 1e3);
      for (int it = 8; it < n - 1; ++it) ba.push_back((double)((long long)h[2 * it + 2] - (long long)h[2 * it + 1]) / 1e3);
      const char* names[] = {"tiny", "B with 4 KB params", "B with 96 KB dyn smem", "96 KB, max carve-out both",
                             "A 32 CTAs, B 128 CTAs"};
      printf("%-28s %-6s  A end -> B start: median %.2f us   (B start -> next A end %.2f us)\n", names[variant],
             use_graph ? "graph" : "stream", median(ab), median(ba));
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}

// How fast can the L2 absorb coalesced FP64 reductions (red.global.add.f64) with the locality of the per-landmark
// Schur formulation?  500k "landmarks", 55 pair products each, 36 doubles per product added into one of ~16k slots of
// the current chunk (slots 288 B apart).  Run under gpurun.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_red(int n_lm, int pairs_per_lm, int slots_per_chunk, int lm_per_chunk, double* __restrict__ C) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
  for (int l = warp; l < n_lm; l += nw) {
    const int chunk = l / lm_per_chunk;
    unsigned h = (unsigned)l * 2654435761u;
    for (int p = 0; p < pairs_per_lm; ++p) {
      h = h * 1664525u + 1013904223u;
      const size_t slot = (size_t)chunk * slots_per_chunk + (h >> 8) % slots_per_chunk;
      double* out = C + slot * 36;
      atomicAdd(out + lane, 1.0 + lane);
      if (lane < 4) atomicAdd(out + 32 + lane, 2.0);
    }
  }
}
int main() {
  const int n_lm = 500000, ppl = 55, spc = 16384, lpc = 8192;
  const size_t n_slots = (size_t)(n_lm / lpc + 1) * spc;
  double* C; cudaMalloc(&C, n_slots * 36 * 8); cudaMemset(C, 0, n_slots * 36 * 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int ctas : {148 * 4, 148 * 8, 148 * 16}) {
    k_red<<<ctas, 256>>>(n_lm, ppl, spc, lpc, C);
    cudaEventRecord(a);
    k_red<<<ctas, 256>>>(n_lm, ppl, spc, lpc, C);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("ctas %5d: %.3f ms for %.1f M pair products (%.2f G fp64 reductions/s)\n", ctas, ms, n_lm * (double)ppl / 1e6, n_lm * (double)ppl * 36 / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Micro-benchmarks of the FP64 building blocks used by the tile Cholesky (run under gpurun): dependent-chain
// latency and independent-issue throughput of DMMA m8n8k4 / m16n8k8 variants, DFMA, 1.0/x, rsqrt, barriers.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
#define N 256
__global__ void bench(long long* out, double* sink, double seed) {
  const int tid = threadIdx.x;
  double a = seed + tid * 1e-3, b = 1.0 + tid * 1e-4;
  long long t0, t1;
  // 0: dependent DMMA 884
  { double c0 = 0, c1 = 0; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) dmma884(c0, c1, a, b);
    t1 = clock64(); if (tid == 0) out[0] = t1 - t0; sink[tid] = c0 + c1; }
  // 1: 8 independent DMMA 884 chains
  { double c[8][2] = {}; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma884(c[j][0], c[j][1], a, b); }
    t1 = clock64(); if (tid == 0) out[1] = t1 - t0; double s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1]; sink[tid] += s; }
  // 2: dependent m16n8k8
  { double d[4] = {}; double aa[4] = {a, b, a, b}; double bb[2] = {b, a}; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) dmma1688(d, aa, bb);
    t1 = clock64(); if (tid == 0) out[2] = t1 - t0; sink[tid] += d[0] + d[1] + d[2] + d[3]; }
  // 3: 4 independent m16n8k8
  { double d[4][4] = {}; double aa[4] = {a, b, a, b}; double bb[2] = {b, a}; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma1688(d[j], aa, bb); }
    t1 = clock64(); if (tid == 0) out[3] = t1 - t0; double s = 0; for (int j = 0; j < 4; ++j) s += d[j][0] + d[j][3]; sink[tid] += s; }
  // 4: dependent m16n8k16 ; 5: 4 independent
  { double d[4] = {}; double aa[8] = {a, b, a, b, a, b, a, b}; double bb[4] = {b, a, b, a}; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) dmma16816(d, aa, bb);
    t1 = clock64(); if (tid == 0) out[4] = t1 - t0; sink[tid] += d[0] + d[1] + d[2] + d[3]; }
  { double d[4][4] = {}; double aa[8] = {a, b, a, b, a, b, a, b}; double bb[4] = {b, a, b, a}; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma16816(d[j], aa, bb); }
    t1 = clock64(); if (tid == 0) out[5] = t1 - t0; double s = 0; for (int j = 0; j < 4; ++j) s += d[j][0] + d[j][3]; sink[tid] += s; }
  // 6: dependent DFMA ; 7: dependent 1.0/x ; 8: dependent rsqrt ; 9: dependent sqrt; 10: __drcp_rn
  { double x = a; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) x = fma(x, b, a);
    t1 = clock64(); if (tid == 0) out[6] = t1 - t0; sink[tid] += x; }
  { double x = a + 2.0; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) x = 1.0 / x + 1.5;
    t1 = clock64(); if (tid == 0) out[7] = t1 - t0; sink[tid] += x; }
  { double x = a + 2.0; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.5;
    t1 = clock64(); if (tid == 0) out[8] = t1 - t0; sink[tid] += x; }
  { double x = a + 2.0; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) x = sqrt(x) + 1.5;
    t1 = clock64(); if (tid == 0) out[9] = t1 - t0; sink[tid] += x; }
  { double x = a + 2.0; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) x = __drcp_rn(x) + 1.5;
    t1 = clock64(); if (tid == 0) out[10] = t1 - t0; sink[tid] += x; }
  // 11: custom rcp: rcp.approx.ftz.f64 + 2 Newton
  { double x = a + 2.0; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
      double e = fma(-x, y, 1.0); y = fma(y, e, y); e = fma(-x, y, 1.0); y = fma(y, e, y); x = y + 1.5; }
    t1 = clock64(); if (tid == 0) out[11] = t1 - t0; sink[tid] += x; }
  // 12: __syncthreads ; 13: named barrier 96
  { __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) __syncthreads();
    t1 = clock64(); if (tid == 0) out[12] = t1 - t0; }
  { __syncthreads(); t0 = clock64();
    if (blockDim.x >= 96 && tid < 96) for (int i = 0; i < N; ++i) asm volatile("barrier.sync 1, 96;" ::: "memory");
    t1 = clock64(); if (tid == 0) out[13] = t1 - t0; }
  // 14: smem store->sync->load round trip
  { __shared__ double sm[256]; double x = a; __syncthreads(); t0 = clock64();
    for (int i = 0; i < N; ++i) { sm[tid] = x; __syncthreads(); x = sm[(tid + 1) % blockDim.x] + 1.0; __syncthreads(); }
    t1 = clock64(); if (tid == 0) out[14] = t1 - t0; sink[tid] += x; }
}
int main() {
  long long* d; double* s; cudaMalloc(&d, 16 * 8); cudaMalloc(&s, 1024 * 8);
  const char* names[] = {"dmma884 dependent", "dmma884 x8 independent (per 8)", "dmma16n8k8 dependent", "dmma16n8k8 x4 indep (per 4)",
                         "dmma16n8k16 dependent", "dmma16n8k16 x4 indep (per 4)", "dfma dependent", "1.0/x dependent (+add)", "rsqrt dependent (+add)",
                         "sqrt dependent (+add)", "__drcp_rn dependent (+add)", "rcp.approx+2 newton (+add)", "__syncthreads", "barrier.sync 96", "sts-sync-lds-sync"};
  for (int threads : {32, 192}) {
    cudaMemset(d, 0, 16 * 8);
    bench<<<1, threads>>>(d, s, 1.25);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost);
    printf("---- %d threads: cycles per iteration\n", threads);
    for (int i = 0; i < 15; ++i) printf("  %-34s %8.1f\n", names[i], (double)h[i] / N);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Micro-benchmarks of dependent-issue latencies on B200 (FP64 pipe, shuffles, shared memory, division).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -O3 -o lat lat.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
__device__ __forceinline__ double rcp_seed(double b) {
  double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b)); return y;
}
template <int MODE>
__global__ void chain(double* out, double a, double b, long long* cyc, int active) {
  __shared__ double sm[64];
  sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
  __syncwarp();
  double x = a + threadIdx.x * 1e-9, y = b;
  int idx = threadIdx.x;
  long long t0 = clock64();
  if (threadIdx.x < active) {
#pragma unroll 16
  for (int i = 0; i < N_IT; i++) {
    if (MODE == 0) x = __dadd_rn(x, y);
    if (MODE == 1) x = __dmul_rn(x, y);
    if (MODE == 2) x = __fma_rn(x, y, y);
    if (MODE == 3) x = x / y;
    if (MODE == 4) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    if (MODE == 5) { idx = (int)sm[idx & 63] & 31; }
    if (MODE == 6) x = rcp_seed(x);
    if (MODE == 7) x = sqrt(x);
    if (MODE == 8) { x = __dadd_rn(__dmul_rn(x, y), y); }
    if (MODE == 9) { x = sm[((int)__double2hiint(x)) & 31]; }
  }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + idx;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
// throughput: many independent warps each doing independent DFMA / DADD streams
template <int MODE>
__global__ void thr(double* out, double a, double b, long long* cyc) {
  double x0 = a + threadIdx.x, x1 = a * 2 + threadIdx.x, x2 = a * 3, x3 = a * 4, x4 = a * 5, x5 = a * 6, x6 = a * 7, x7 = a * 8;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_IT; i++) {
    if (MODE == 0) { x0 = __fma_rn(x0, b, b); x1 = __fma_rn(x1, b, b); x2 = __fma_rn(x2, b, b); x3 = __fma_rn(x3, b, b); x4 = __fma_rn(x4, b, b); x5 = __fma_rn(x5, b, b); x6 = __fma_rn(x6, b, b); x7 = __fma_rn(x7, b, b); }
    if (MODE == 1) { x0 = __dadd_rn(x0, b); x1 = __dadd_rn(x1, b); x2 = __dadd_rn(x2, b); x3 = __dadd_rn(x3, b); x4 = __dadd_rn(x4, b); x5 = __dadd_rn(x5, b); x6 = __dadd_rn(x6, b); x7 = __dadd_rn(x7, b); }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const char* names[] = {"dadd", "dmul", "dfma", "ddiv", "shfl64", "lds->idx", "mufu.rcp64h", "dsqrt", "dmul+dadd", "lds64 dep"};
  long long h;
#define RUN(M, act) chain<M><<<1, 32>>>(out, 1.000001, 0.9999999, cyc, act); chain<M><<<1, 32>>>(out, 1.000001, 0.9999999, cyc, act); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-14s active=%2d  %.1f cycles/op\n", names[M], act, (double)h / N_IT);
  RUN(0, 32) RUN(0, 2) RUN(1, 32) RUN(2, 32) RUN(2, 2) RUN(3, 32) RUN(3, 2) RUN(4, 32) RUN(5, 32) RUN(6, 32) RUN(7, 32) RUN(8, 32) RUN(9, 32)
  for (int warps = 1; warps <= 16; warps *= 2) {
    thr<0><<<1, 32 * warps>>>(out, 1.0, 0.999, cyc); thr<0><<<1, 32 * warps>>>(out, 1.0, 0.999, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dfma throughput %2d warps/SM: %.2f warp-instr/cycle/SM\n", warps, 8.0 * N_IT * warps / h);
    thr<1><<<1, 32 * warps>>>(out, 1.0, 0.999, cyc); thr<1><<<1, 32 * warps>>>(out, 1.0, 0.999, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dadd throughput %2d warps/SM: %.2f warp-instr/cycle/SM\n", warps, 8.0 * N_IT * warps / h);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

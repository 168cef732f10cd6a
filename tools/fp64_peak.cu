// FP64 pipe microbenchmark for B200 (sm_100a): DFMA vs DMMA shapes.  Measurement tool only (not product).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int NT>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters) {
  double c[NT][2];
  double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int NT>
__global__ void __launch_bounds__(256) k_dmma1684(double* out, int iters) {
  double c[NT][4];
  double a0 = threadIdx.x * 1e-6, a1 = a0 + 1e-3, b = 1.0 + threadIdx.x * 1e-7;
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = i; c[i][3] = 1; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int NT>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters) {
  double c[NT][4];
  double a0 = threadIdx.x * 1e-6, a1 = a0 + 1e-3, a2 = a0 * 2, a3 = a1 * 2, b0 = 1.0 + threadIdx.x * 1e-7, b1 = b0 * 0.5;
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = i; c[i][3] = 1; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int NT>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters) {
  double c[NT][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-6 * (i + 1);
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = 1.0 + threadIdx.x * 1e-7 * (i + 1);
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = i; c[i][3] = 1; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: DFMA and DMMA concurrently in the same warp (do the pipes add up?)
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double fa, double fb) {
  double c[8][2]; double acc[16];
  double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = i + threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      acc[2*i] = fma(acc[2*i], fa, fb); acc[2*i+1] = fma(acc[2*i+1], fa, fb);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) out[i] = in[i];
}

template<typename F> float timeit(F f, int rep = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < rep; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 20000;
  for (int bps = 1; bps <= 4; bps *= 2) {
    int grid = sms * bps, thr = 256; double nthr = (double)grid * thr;
    float ms;
    ms = timeit([&]{ k_dfma<16><<<grid, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf("bps %d DFMA x16        : %8.3f ms  %7.2f TFLOP/s\n", bps, ms, 2.0 * nthr * 16 * iters / ms / 1e9);
    ms = timeit([&]{ k_dmma884<8><<<grid, thr>>>(out, iters); });
    printf("bps %d DMMA m8n8k4 x8  : %8.3f ms  %7.2f TFLOP/s\n", bps, ms, 2.0 * (nthr / 32) * 8 * 256 * iters / ms / 1e9);
    ms = timeit([&]{ k_dmma1684<8><<<grid, thr>>>(out, iters); });
    printf("bps %d DMMA m16n8k4 x8 : %8.3f ms  %7.2f TFLOP/s\n", bps, ms, 2.0 * (nthr / 32) * 8 * 512 * iters / ms / 1e9);
    ms = timeit([&]{ k_dmma1688<8><<<grid, thr>>>(out, iters); });
    printf("bps %d DMMA m16n8k8 x8 : %8.3f ms  %7.2f TFLOP/s\n", bps, ms, 2.0 * (nthr / 32) * 8 * 1024 * iters / ms / 1e9);
    ms = timeit([&]{ k_dmma16816<8><<<grid, thr>>>(out, iters / 2); });
    printf("bps %d DMMA m16n8k16 x8: %8.3f ms  %7.2f TFLOP/s\n", bps, ms, 2.0 * (nthr / 32) * 8 * 2048 * (iters / 2) / ms / 1e9);
    ms = timeit([&]{ k_mixed<<<grid, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf("bps %d MIXED 8dmma+16dfma: %8.3f ms  %7.2f TFLOP/s (dmma part %7.2f, dfma part %7.2f)\n", bps, ms,
           2.0 * ((nthr / 32) * 8 * 256 + nthr * 16) * iters / ms / 1e9, 2.0 * (nthr / 32) * 8 * 256 * iters / ms / 1e9, 2.0 * nthr * 16 * iters / ms / 1e9);
  }
  // fewer warps: latency / ILP sensitivity of DMMA (1 block of 128 thr per SM => 1 warp per SMSP)
  {
    int grid = sms, thr = 128; double nthr = (double)grid * thr; float ms;
    ms = timeit([&]{ k_dmma884<8><<<grid, thr>>>(out, iters); });
    printf("1 warp/SMSP DMMA m8n8k4 x8 : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * (nthr / 32) * 8 * 256 * iters / ms / 1e9);
    ms = timeit([&]{ k_dmma884<2><<<grid, thr>>>(out, iters); });
    printf("1 warp/SMSP DMMA m8n8k4 x2 : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * (nthr / 32) * 2 * 256 * iters / ms / 1e9);
    ms = timeit([&]{ k_dmma884<1><<<grid, thr>>>(out, iters); });
    printf("1 warp/SMSP DMMA m8n8k4 x1 (latency): %8.3f ms  => %.1f ns per dependent mma\n", ms, ms * 1e6 / iters);
    ms = timeit([&]{ k_dfma<1><<<grid, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf("1 warp/SMSP DFMA x1 (latency): %8.3f ms  => %.2f ns per dependent dfma\n", ms, ms * 1e6 / iters);
    ms = timeit([&]{ k_dfma<16><<<grid, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf("1 warp/SMSP DFMA x16 : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * nthr * 16 * iters / ms / 1e9);
  }
  // HBM copy
  {
    size_t n = (size_t)1 << 28; // 268M double2 = 4 GiB each
    double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16)); CK(cudaMemset(a, 1, n * 16));
    float ms = timeit([&]{ k_copy<<<sms * 16, 512>>>(a, b, n); });
    printf("copy double2 4GiB: %8.3f ms  %7.1f GB/s (r+w)\n", ms, 2.0 * n * 16 / ms / 1e6);
    ms = timeit([&]{ cudaMemcpyAsync(b, a, n * 16, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D 4GiB: %8.3f ms  %7.1f GB/s (r+w)\n", ms, 2.0 * n * 16 / ms / 1e6);
  }
  return 0;
}

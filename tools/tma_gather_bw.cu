// TMA bulk-copy gather feed-rate microbenchmark (measurement tool, not product).
// Mimics the producer side of k_apply_dmma: per stage 8 x 5184 B blocks gathered through a bcc-like stencil from a
// large vector + one 10368 B block from a small (L2-resident) table, into a ring of NSTAGE smem slots; consumers only
// wait and release.  Reports aggregate GB/s and per-SM B/clk.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define BLKD 648
#define HBLK 1296
#define STAGE_D (HBLK + 8 * BLKD)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DN;\nbra LW;\nDN:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
template <int NSTAGE>
__global__ void __launch_bounds__(288, 1) k_feed(const double* vec, const double* H, int kk, int ntiles, int nslot, int nx, int ny, int delay) {
  extern __shared__ __align__(128) unsigned char raw[];
  double* st = (double*)raw;
  uint64_t* full = (uint64_t*)(raw + (size_t)NSTAGE * STAGE_D * 8);
  uint64_t* empty = full + NSTAGE;
  int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int s = 0; s < NSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int offs[15] = {0, 1, -1, 2 * nx, -2 * nx, 2 * nx * ny, -2 * nx * ny, 3, -3, 2 * nx + 1, -2 * nx - 1, 2 * nx * ny + 1, -2 * nx * ny - 1, 2 * nx * ny - 2 * nx, -2 * nx * ny + 2 * nx};
  uint32_t it = 0;
  if (warp == 8) {
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      for (int j = 0; j < nslot; j++, it++) {
        int slot = it % NSTAGE;
        mbar_wait(&empty[slot], ((it / NSTAGE) & 1) ^ 1);
        double* sm = st + (size_t)slot * STAGE_D;
        if (lane == 0) mbar_expect(&full[slot], STAGE_D * 8);
        __syncwarp();
        if (lane < 8) {
          long site = (long)tile * 8 + lane + offs[j % 15];
          site = ((site % kk) + kk) % kk;
          bulk(sm + HBLK + lane * BLKD, vec + site * BLKD, BLKD * 8, &full[slot]);
        } else if (lane == 8) bulk(sm, H + (size_t)(j % 15) * HBLK, HBLK * 8, &full[slot]);
      }
  } else {
    double sink = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      for (int j = 0; j < nslot; j++, it++) {
        int slot = it % NSTAGE;
        mbar_wait(&full[slot], (it / NSTAGE) & 1);
        const double* sm = st + (size_t)slot * STAGE_D;
        sink += sm[tid];
        if (delay > 0) { long long t0 = clock64(); while (clock64() - t0 < delay) {} }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
      }
    if (sink == 12345.678) printf("x");
  }
}
template <int NSTAGE> void run(const double* vec, const double* H, int kk, int nx, int ny, int sms, int delay) {
  size_t smem = (size_t)NSTAGE * STAGE_D * 8 + 128;
  cudaFuncSetAttribute(k_feed<NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int ntiles = kk / 8;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k_feed<NSTAGE><<<sms, 288, smem>>>(vec, H, kk, ntiles, 15, nx, ny, delay);
  cudaEventRecord(a);
  k_feed<NSTAGE><<<sms, 288, smem>>>(vec, H, kk, ntiles, 15, nx, ny, delay);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double bytes = (double)ntiles * 15 * STAGE_D * 8;
  printf("stages %d delay %5d clk: %8.3f ms  %8.1f GB/s  %6.1f B/clk/SM (@1.965GHz)  err=%s\n", NSTAGE, delay, ms, bytes / ms / 1e6, bytes / ms / 1e6 / sms / 1.965, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  int nx = 100, ny = 100, nz = 50; int kk = 2 * nx * ny * nz;
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *vec, *H; cudaMalloc(&vec, (size_t)(kk + 1) * BLKD * 8); cudaMalloc(&H, 16 * HBLK * 8);
  cudaMemset(vec, 0, (size_t)(kk + 1) * BLKD * 8); cudaMemset(H, 0, 16 * HBLK * 8);
  for (int delay : {0, 1500, 3000, 3400}) {
    run<2>(vec, H, kk, nx, ny, p.multiProcessorCount, delay);
    run<3>(vec, H, kk, nx, ny, p.multiProcessorCount, delay);
    run<4>(vec, H, kk, nx, ny, p.multiProcessorCount, delay);
  }
  return 0;
}

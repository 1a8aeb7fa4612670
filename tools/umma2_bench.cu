// cta_group::2 probe: a cluster of two CTAs issues ONE tcgen05.mma M256 x N x K16 (bf16, K-major, no swizzle) from the
// leader CTA and both CTAs dump their TMEM accumulators.  The host checks the operand split this repository's round-2
// plan assumes (each CTA supplies its own 128 rows of A and HALF of B's N rows; CTA r's TMEM holds output rows
// [128 r, 128 r + 128) x all N columns), then times back-to-back MMAs for N = 64 / 128 / 256.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma2_bench umma2_bench.cu      (run under `timeout`)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// out: [2 ranks][128 lanes][N] fp32;  cyc: cycles of the timed loop (leader)
__global__ void __cluster_dims__(2, 1, 1) probe(int N, int iters, float* out, long long* cyc) {
  __shared__ __align__(1024) uint8_t sA[2 * 128 * 16];        // [k group][128 rows][8 bf16]
  __shared__ __align__(1024) uint8_t sB[2 * 128 * 16];        // [k group][N/2 rows][8 bf16]  (N/2 <= 128)
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NH = N / 2;
  // A[m][k] (global row m = 128*rank + row), B[n][k] (global row n = NH*rank + row): small integers, exact in bf16
  for (int i = tid; i < 128 * 16; i += blockDim.x) {
    const int row = i / 16, k = i % 16, m = 128 * (int)rank + row;
    const float v = (float)((m % 7) - 3) * (float)((k % 3) + 1);
    reinterpret_cast<__nv_bfloat16*>(sA)[((k / 8) * 128 + row) * 8 + (k % 8)] = __float2bfloat16(v);
  }
  for (int i = tid; i < NH * 16; i += blockDim.x) {
    const int row = i / 16, k = i % 16, n = NH * (int)rank + row;
    const float v = (float)((n % 5) - 2) + (float)(k % 2);
    reinterpret_cast<__nv_bfloat16*>(sB)[((k / 8) * NH + row) * 8 + (k % 8)] = __float2bfloat16(v);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tslot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  if (rank == 0 && warp == 1 && elect_one()) {
    const uint64_t da = desc(smem_u32(sA), 128 * 16, 128), db = desc(smem_u32(sB), NH * 16, 128);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t acc = i ? 1u : 0u;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(iters > 1 ? 0u : acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    while (!mbar_try(smem_u32(&bar), 0)) {}
    cyc[0] = clock64() - t0;
  }
  // every thread of both CTAs waits for the commit that the leader multicast to both barriers
  while (!mbar_try(smem_u32(&bar), 0)) {}
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp < 4) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) out[((size_t)rank * 128 + warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 2 * 128 * 256 * sizeof(float));
  cudaMalloc(&cyc, 8);
  for (int N : {64, 128, 256}) {
    cudaMemset(out, 0, 2 * 128 * 256 * sizeof(float));
    probe<<<2, 128>>>(N, 1, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N %d: error %s\n", N, cudaGetErrorString(e)); return 1; }
    static float h[2 * 128 * 256];
    cudaMemcpy(h, out, 2 * 128 * N * sizeof(float), cudaMemcpyDeviceToHost);
    long bad = 0; int shown = 0;
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int k = 0; k < 16; ++k) ref += (float)((m % 7) - 3) * (float)((k % 3) + 1) * ((float)((n % 5) - 2) + (float)(k % 2));
        const float got = h[(size_t)m * N + n];          // rank = m / 128, lane = m % 128
        if (got != ref) { ++bad; if (shown < 4) { printf("  N %d mismatch m %d n %d got %g want %g\n", N, m, n, got, ref); ++shown; } }
      }
    printf("N %3d: operand-split hypothesis %s (%ld mismatches of %d)\n", N, bad ? "REJECTED" : "confirmed", bad, 256 * N);
    const int iters = 4096;
    probe<<<2, 128>>>(N, iters, out, cyc);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N %d timing: error %s\n", N, cudaGetErrorString(e)); return 1; }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("N %3d: %.1f cycles per tcgen05.mma.cta_group::2 M256 x N%d x K16 (one cluster; 1-CTA M128 takes 48.3 / 64.3 / 128.3)\n", N, (double)c / iters, N);
  }
  return 0;
}

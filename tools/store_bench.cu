// Micro-benchmark: how long does a CTA take to push one output tile (planes x positions x 16 B, the C8P layout of
// csrc/conv.cuh) to global memory -- (a) 16-byte stores from registers by 512 threads (the conv epilogue today),
// (b) the same values written to shared memory, then one cp.async.bulk shared -> global per plane issued by one thread.
// grid CTAs run at once (one per SM); per-CTA clock64 cycles from first store to completion (wait_group.read for (b)).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bench store_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) store_kernel(uint4* __restrict__ out, long long plane_stride16, int planes, int mt,
                                                       int tiles_per_cta, long long* cycles, int stagger) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint4* st = reinterpret_cast<uint4*>(smem);
  const int tid = threadIdx.x;
  if (stagger) {                                  // de-synchronise the CTAs (what the three concurrent branches do in a stage)
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)(blockIdx.x % 16) * stagger) {}
  }
  __syncthreads();
  const long long t_begin = clock64();
  for (int it = 0; it < tiles_per_cta; ++it) {
    const long long base = ((long long)(blockIdx.x * tiles_per_cta + it)) * mt;
    if (MODE == 0) {
      for (int i = tid; i < planes * mt; i += 512) {
        const int g = i / mt, p = i - g * mt;
        out[(long long)g * plane_stride16 + base + p] = make_uint4(i, it, g, p);
      }
    } else {
      for (int i = tid; i < planes * mt; i += 512) st[i] = make_uint4(i, it, i / mt, i % mt);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        for (int g = 0; g < planes; ++g)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (long long)g * plane_stride16 + base),
                       "r"(smem_u32(st + (long long)g * mt)), "r"(mt * 16)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t_begin;
}

int main() {
  const int grid = 148;
  const int mt = 512;
  for (int planes : {8, 16, 32}) {
    for (int tiles : {1, 4}) {
      const long long plane_stride16 = (long long)grid * tiles * mt + 1024;
      uint4* out;
      long long* cyc;
      cudaMalloc(&out, sizeof(uint4) * plane_stride16 * planes);
      cudaMalloc(&cyc, sizeof(long long) * grid);
      const size_t smem = (size_t)planes * mt * 16;
      cudaFuncSetAttribute(store_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      for (int stagger : {0, 2000}) {
        for (int mode = 0; mode < 2; ++mode) {
          float best = 1e30f;
          double avg_cyc = 0;
          for (int rep = 0; rep < 5; ++rep) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            if (mode == 0) store_kernel<0><<<grid, 512, 0>>>(out, plane_stride16, planes, mt, tiles, cyc, stagger);
            else store_kernel<1><<<grid, 512, smem>>>(out, plane_stride16, planes, mt, tiles, cyc, stagger);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
            long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0;
            for (int i = 0; i < grid; ++i) s += (double)h[i];
            avg_cyc = s / grid;
          }
          printf("planes %2d (%3d KB/tile) tiles/CTA %d stagger %4d  %-22s  %8.0f cycles per CTA (%.1f B/cycle/SM)  kernel %.1f us  err=%d\n",
                 planes, planes * mt * 16 / 1024, tiles, stagger, mode == 0 ? "st.global.v4 x512 thr" : "smem + cp.async.bulk", avg_cyc,
                 (double)planes * mt * 16 * tiles / avg_cyc, best * 1e3, (int)cudaGetLastError());
        }
      }
      cudaFree(out); cudaFree(cyc);
    }
  }
  return 0;
}

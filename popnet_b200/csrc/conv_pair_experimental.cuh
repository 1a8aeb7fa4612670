// EXPERIMENTAL, off by default (PopnetNetConfig.tuning, POPNET_TUNE_PAIR): the CTA-pair convolution kernel.
// Included twice by conv_kernels.cu -- once for the device code (inside its anonymous namespace, after the shared
// tcgen05 / mbarrier helpers), once with POPNET_PAIR_LAUNCH_PART for the host launcher.  Validated bit-identical against the
// default kernels (tests/test_forward.py); measured: the two non-residual 64 -> 64 layers 57 -> 51 us and the forward alone
// 0.994 -> 0.984 ms at burst clocks, but the pipelined step 4-6 % SLOWER (a cluster needs both SMs of a TPC free at once and
// the overlapped decode's CTAs split pairs).  Kept as the working reference for cta_group::2.
#ifndef POPNET_PAIR_LAUNCH_PART

// ------------------------------------------------------------------------------------------------
// CTA-pair kernel for the 64 -> 64 3x3 layers: a cluster of two CTAs issues
// tcgen05.mma.cta_group::2 (M 256 x N 64 x K 16) from the leader.  Each CTA stages its OWN tile of positions (the A
// operand, same shared-memory offsets in both CTAs) and only HALF of the weights (32 of the 64 output channels of every
// tap: the B operand of a pair MMA is split over the two CTAs), so the operand bytes read per MMA and SM drop from 6 KB
// to 5 KB (tools/umma2_bench.cu: 48.3 -> 43.1 cycles) and the resident weights from 72 KB to 36 KB.  Each CTA's TMEM
// receives its own 128 rows x 64 columns, so the epilogue is the single-CTA one.
//   leader waits for: its own A stage, the peer's A stage (relayed by the peer's otherwise idle MMA thread with a remote
//   mbarrier arrive), the accumulator stage released by the epilogue warps of BOTH CTAs (remote arrives);
//   leader's commits (A stage consumed, accumulators ready) are multicast to both CTAs' barriers.
// Both CTAs walk the same number of tile slots; a slot past the last tile recomputes tile 0 and stores nothing.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t target_rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(target_rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
template <int A_OFF, int B_OFF, int D_OFF>
__device__ __forceinline__ void umma2_off(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 ta, tb, td;\n\t"
      ".reg .b64 da, db;\n\t"
      "add.u32 ta, %1, %7;\n\t"
      "add.u32 tb, %3, %8;\n\t"
      "add.u32 td, %0, %9;\n\t"
      "mov.b64 da, {ta, %2};\n\t"
      "mov.b64 db, {tb, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [td], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "n"(A_OFF), "n"(B_OFF), "n"(D_OFF)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

template <int NACC>
__global__ void __launch_bounds__(kTcThreads, 1) conv_pair64_kernel(const ConvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NT = 64, MT = NACC * 128, AS = 2;
  constexpr uint32_t kAccCols = NACC * NT;
  static_assert(AS * kAccCols <= 512, "accumulators exceed TMEM");
  constexpr uint32_t kCols = 512;
  constexpr uint32_t kBTapBytes = 8u * 32u * 16u;           // one tap, this CTA's 32 output channels: [k8][32][8]
  const int halo = a.Wp + 1;
  const int apos = MT + 2 * halo;
  const uint32_t a_plane_bytes = (uint32_t)apos * 16u;
  const uint32_t a_stage_bytes = 8u * a_plane_bytes;
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((2 * (size_t)a_stage_bytes + 127) & ~(size_t)127);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 9 * kBTapBytes);
  // barriers: a_full[2], a_full_peer[2] (leader only), a_empty[2], b_full, acc_full[2], acc_empty[2] (leader only)
  constexpr int kNumBars = 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  float* s_shift = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(bars + kNumBars + 1) + 15) & ~(uintptr_t)15);
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_full_peer = [&](int s) { return bar0 + 8u * (2 + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (4 + s); };
  const uint32_t b_full = bar0 + 8u * 6;
  auto acc_full = [&](int s) { return bar0 + 8u * (7 + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (9 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_tiles = (a.P + MT - 1) / MT;
  const int tile_end = (int)((num_tiles + gridDim.x - 1) / gridDim.x * gridDim.x);      // same slot count in every CTA
  trace_min(a.trace, 0);
  const unsigned long long t_enter = trace_enter(a.trace);
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumBars; ++i) mbar_init(bar0 + 8u * i, i >= 9 ? 2u * kEpiWarps : 1u);     // acc_empty: the epilogue warps of both CTAs
    fence_mbar_init();
  }
  cluster_sync_all();
  for (int i = threadIdx.x; i < NT; i += kTcThreads) s_shift[i] = a.shift[i];
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp == 0 && elect_one()) {
    // this CTA's half of the weights (constant data: fetched before waiting for the previous layer)
    mbar_expect_tx(b_full, 9u * kBTapBytes);
    for (int t = 0; t < 9; ++t)
      for (int g = 0; g < 8; ++g)
        bulk_g2s(smem_u32(sB) + (uint32_t)(t * 8 + g) * 512u, a.w + ((long long)(t * 8 + g) * 64 + 32 * (long long)rank) * 8, 512u, b_full);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  trace_min(a.trace, 1);

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- producer: this CTA's A tiles ----------------
      int ia = 0;
      for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++ia) {
        const int t0 = (tile < num_tiles ? tile : 0) * MT;
        const int as = ia & 1;
        if (ia >= 2) mbar_wait_relaxed(a_empty(as), ((ia >> 1) - 1) & 1);
        mbar_expect_tx(a_full(as), a_stage_bytes);
        for (int g = 0; g < 8; ++g)
          bulk_g2s(smem_u32(sA + (size_t)as * a_stage_bytes + (size_t)g * a_plane_bytes),
                   a.in + (long long)g * a.in_plane_stride + (long long)(t0 - halo) * 8, a_plane_bytes, a_full(as));
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      if (rank != 0) {
        // ---------------- peer: tell the leader when this CTA's A stage (and, first, its weights) have landed ----------------
        mbar_wait(b_full, 0);
        int ia = 0;
        for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++ia) {
          const int as = ia & 1;
          mbar_wait(a_full(as), (ia >> 1) & 1);
          mbar_arrive_remote(a_full_peer(as), 0u);
        }
      } else {
        // ---------------- leader: MMA issue for the pair ----------------
        const uint32_t idesc = (1u << 4) | ((a.fmt == 0 ? 1u : 0u) << 7) | ((a.fmt == 0 ? 1u : 0u) << 10) | ((uint32_t)(NT >> 3) << 17) |
                               ((uint32_t)(256 >> 4) << 24);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lo0 = ((smem_u32(sA) >> 4) & 0x3FFFu) | (((a_plane_bytes >> 4) & 0x3FFFu) << 16);
        const uint32_t b_lo0 = ((smem_u32(sB) >> 4) & 0x3FFFu) | (((512u >> 4) & 0x3FFFu) << 16);      // LBO: 32 rows x 16 B
        const uint32_t a_kstep = (2u * a_plane_bytes) >> 4;
        constexpr uint32_t b_kstep = (2u * 512u) >> 4;
        mbar_wait(b_full, 0);
        int ia = 0;
        for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++ia) {
          const int s = ia & 1, as = ia & 1;
          const uint32_t ph = (ia >> 1) & 1;
          if (ia >= AS) { while (!mbar_try_wait_cluster(acc_empty(s), ph ^ 1)) {} }
          mbar_wait(a_full(as), ph);
          while (!mbar_try_wait_cluster(a_full_peer(as), ph)) {}
          tc_fence_after();
          const uint32_t tmem_acc = tmem_base + (uint32_t)s * kAccCols;
          const uint32_t a_lo_stage = a_lo0 + ((as * a_stage_bytes) >> 4);
          int row0 = halo - a.Wp;
#pragma unroll 1
          for (int r = 0; r < 3; ++r, row0 += a.Wp) {
            auto tap = [&](auto d_c) {
              constexpr int d = decltype(d_c)::value;
              const uint32_t a_lo_tap = a_lo_stage + (uint32_t)(row0 + d - 1);
              const uint32_t b_lo_tap = b_lo0 + (uint32_t)((r * 3 + d) * (int)(kBTapBytes >> 4));
              const uint32_t first = (r == 0 && d == 0) ? 0u : 1u;
              static_for<4>([&](auto kk_c) {
                constexpr int kk = decltype(kk_c)::value;
                const uint32_t a_lo_k = a_lo_tap + kk * a_kstep;
                static_for<NACC>([&](auto acc_c) {
                  constexpr int acc = decltype(acc_c)::value;
                  umma2_off<acc * 128, kk * (int)b_kstep, acc * NT>(tmem_acc, a_lo_k, desc_hi, b_lo_tap, desc_hi, idesc,
                                                                    (kk == 0) ? first : 1u);
                });
              });
            };
            static_for<3>(tap);
          }
          umma2_commit_mc(a_empty(as));          // both CTAs' A stages are free again once these MMAs retire
          umma2_commit_mc(acc_full(s));          // both CTAs' accumulators are ready
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ---------------- epilogue: this CTA's 128-row accumulators (as in conv_tc_kernel) ----------------
    const int q = warp & 3;
    const int sub = (warp - kEpiWarp0) >> 2;
    const float slope = a.act == kActRelu ? 0.f : (a.act == kActLeaky ? 0.1f : 1.f);
    const uint32_t mHs = div_magic(a.Hs), mWp = div_magic(a.Wp);
    const int epi_variant = (a.fmt != 0 ? 2 : 0) + (slope == 0.f ? 0 : 1);
    constexpr int kItems = NACC * (NT / 32);
    int ia = 0;
    for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x, ++ia) {
      const int s = ia & 1;
      const bool live = tile < num_tiles;
      const int t0 = tile * MT;
      if (live && a.res != nullptr && tile + (int)gridDim.x < num_tiles && (lane & 7) == 0) {      // residual of the next tile -> L2
        const int tn0 = (tile + (int)gridDim.x) * MT;
#pragma unroll 1
        for (int it = sub; it < kItems; it += 4) {
          const int acc = it / (NT / 32), j = it - acc * (NT / 32);
          const int pos = tn0 + acc * 128 + q * 32 + lane;
          if (pos < a.P) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + (long long)(j * 4 + g) * a.res_plane_stride + (long long)pos * 8));
          }
        }
      }
      mbar_wait_relaxed(acc_full(s), (ia >> 1) & 1);
      tc_fence_after();
      if (live) {
#pragma unroll 1
        for (int it = sub; it < kItems; it += 4) {
          const int acc = it / (NT / 32), j = it - acc * (NT / 32);
          const int pos = t0 + acc * 128 + q * 32 + lane;
          const PosInfo pi = c8p_locate_fast(pos, a.P, a.Hs, a.Wp, mHs, mWp);
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kAccCols + acc * NT + j * 32), r);
          const int plane = (j * 32) >> 3;
          uint4 rs[4];
          const bool has_res = a.res != nullptr && pi.interior;
          if (has_res) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              rs[g] = *reinterpret_cast<const uint4*>(a.res + (long long)(plane + g) * a.res_plane_stride + (long long)pos * 8);
          }
          tmem_ld_wait();
          if (pi.in_range) {
            auto store4 = [&](auto fmt_c, auto mode_c) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 o = finish8_fast<decltype(fmt_c)::value, decltype(mode_c)::value>(
                    r + g * 8, s_shift + j * 32 + g * 8, has_res, rs[g], slope, pi.interior);
                *reinterpret_cast<uint4*>(a.out + (long long)(plane + g) * a.out_plane_stride + (long long)pos * 8) = o;
              }
            };
            using std::integral_constant;
            switch (epi_variant) {
              case 0: store4(integral_constant<int, 0>{}, integral_constant<int, 0>{}); break;
              case 1: store4(integral_constant<int, 0>{}, integral_constant<int, 1>{}); break;
              case 2: store4(integral_constant<int, 1>{}, integral_constant<int, 0>{}); break;
              default: store4(integral_constant<int, 1>{}, integral_constant<int, 1>{}); break;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                           // the leader owns the accumulator-stage barrier of the pair
        if (rank == 0) mbar_arrive(acc_empty(s));
        else mbar_arrive_remote(acc_empty(s), 0u);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  trace_max(a.trace, 2);
  trace_exit(a.trace, t_enter);
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kCols) : "memory");
  }
}


#else  // POPNET_PAIR_LAUNCH_PART

namespace {
template <int NACC>
int launch_pair64(const ConvArgs& a, cudaStream_t st) {
  auto kern = conv_pair64_kernel<NACC>;
  const int halo = a.Wp + 1;
  const size_t a_bytes = (((size_t)2 * 8 * (NACC * 128 + 2 * halo) * 16) + 127) & ~(size_t)127;
  const size_t smem = a_bytes + 9 * 4096 + 512;
  if (smem > kSmemLimit) return POPNET_ERR_UNSUPPORTED;
  POPNET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  POPNET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int tiles = (a.P + NACC * 128 - 1) / (NACC * 128);
  int grid = tiles < kNumSMs ? tiles : kNumSMs;
  grid = (grid + 1) & ~1;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = 2; attrs[1].val.clusterDim.y = 1; attrs[1].val.clusterDim.z = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cfg.attrs = attrs; cfg.numAttrs = 2;
  ConvArgs at = a;
  at.trace = next_trace_slot(64 * 1000 + NACC * 100 + 90 + 7);
  POPNET_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, at));
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}
}  // namespace


#endif

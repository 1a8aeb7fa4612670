// Convolution kernels of the rtpose_light3d forward (reference: third_party_methods/lib/network/
// rtpose_light3d.py:24-32, 56-72, 145-158, 222-246 -- nn.Conv2d + eval-mode BatchNorm2d + ReLU /
// LeakyReLU(0.1) + residual add + AvgPool2d(3,2,1), which the reference runs through cuDNN/ATen).
//
//   conv_tc_kernel   the product path: implicit GEMM on the 5th-generation tensor cores.
//                    D[128 positions x NT channels] (fp32, TMEM) += A[positions x 16 cin] * B[NT x 16 cin]
//                    issued as tcgen05.mma.cta_group::1.kind::f16 by one thread; operands are staged in
//                    shared memory by bulk-async copies (cp.async.bulk + mbarrier complete_tx) in the
//                    canonical no-swizzle K-major core-matrix layout, which is exactly the C8P activation
//                    layout (conv.cuh), so a 3x3 tap is a descriptor start-address shift and every input
//                    tile is fetched once per 64 input channels instead of once per tap.
//                    Warp roles: 0 = copy producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue
//                    (tcgen05.ld -> +shift, residual, activation -> 16-bit C8P store and/or fp32 NCHW head).
//                    Operands are bf16 (default) or fp16 -- a runtime field of the instruction descriptor.
//   conv_simt_kernel a plain CUDA-core evaluation of the same packed operands, used by tests / bring-up
//                    to localise tensor-core descriptor mistakes (POPNET_FWD_IMPL_SIMT); not a product path.
//   stem_kernel      model0.conv1 (7x7, stride 2, C_in = 1): direct fp32 evaluation, 0.6 % of the FLOPs.
//   pool_kernel      AvgPool2d(3, 2, 1) with count_include_pad (always / 9) on C8P planes.
//
// Roofline: tensor pipe for conv_tc_kernel (13.34 GFLOP per frame in total); HBM for the
// stem and the pools.
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>
#include <cstdlib>
#include <type_traits>
#include <utility>

#include "conv.cuh"

namespace popnet {
// ---- timeline tracing: the debug entry points in forward.cu hand out one 4-word slot per launch
unsigned long long* g_trace_buf = nullptr;
int g_trace_cap = 0;
std::atomic<int> g_trace_next{0};
int g_trace_tags[1024];
namespace {

// ------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// for waits that are expected to be long (epilogue warps waiting for a whole tile of MMAs): back off so
// that the polling does not compete with the tensor core for shared-memory bandwidth
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(128);      // (32 ns measured equal)
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0, 16 B aligned)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// ---- thread-block cluster of two CTAs (MC kernels): B (weight) stages are loaded half by each CTA and multicast to both
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bulk copy delivered to the same shared-memory offset (and signalled on the same mbarrier offset) of every CTA in `mask`
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
// One lane of a fully converged warp.  Unlike `lane == 0`, the compiler knows the elected region runs on exactly one
// thread and issues the uniform-datapath tcgen05 instructions directly; behind a plain lane test it wraps EVERY
// tcgen05.mma / commit in an ELECT + BRA.U.ANY serialisation loop (4 extra dependent instructions per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The MMA with its two descriptors given as (lo, hi) 32-bit halves -- only `lo` (the start address) changes between the
// MMAs of a tile -- and compile-time offsets added to the start addresses and the TMEM column INSIDE the asm block: the compiler cannot hoist the descriptor arithmetic of a whole unrolled tap group above its first MMA (which
// made it spill uniform registers between the MMAs); each MMA is preceded by exactly its own three adds.
template <int A_OFF, int B_OFF, int D_OFF>
__device__ __forceinline__ void umma_off(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
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
      "tcgen05.mma.cta_group::1.kind::f16 [td], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "n"(A_OFF), "n"(B_OFF), "n"(D_OFF)
      : "memory");
}
// compile-time loop: f(std::integral_constant<int, 0>{}) ... f(std::integral_constant<int, N-1>{})
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}
// mbarrier arrive when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// the same arrive delivered to the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// two fp32 -> one packed pair of 16-bit operands (single cvt.rn instruction per pair)
__device__ __forceinline__ uint32_t pack2(float lo, float hi, int fmt) {
  if (fmt == 0) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  }
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u, int fmt) {
  if (fmt == 0) return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}

// UMMA shared-memory descriptor, SWIZZLE_NONE, K-major: 8 rows x 16 B core matrices, rows 16 B apart inside
// a core matrix; SBO = byte distance between 8-row groups (M/N direction), LBO = byte distance between the
// two core matrices one K=16 instruction consumes (cute/atom/mma_traits_sm100.hpp, "INTERLEAVE" K-major).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor for kind::f16: D = F32, A = B = BF16 (fmt 0) or F16 (fmt 1), both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t umma_idesc(int n, int fmt) {
  return (1u << 4) | ((fmt == 0 ? 1u : 0u) << 7) | ((fmt == 0 ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// epilogue shared by the tensor-core and the SIMT kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ PosInfo locate(int pos, const ConvArgs& a) { return c8p_locate(pos, a.P, a.Hs, a.Wp); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// v: raw accumulators of channels ch0 .. ch0+7 at `pos`
__device__ __forceinline__ void finish8(const ConvArgs& a, int pos, const PosInfo& pi, int ch0, const float (&v)[8]) {
  if (!pi.in_range) return;
  __align__(16) h16 ob[8];
  if (pi.interior) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = 0.f;
    if (a.res) {
      const uint4 q = *reinterpret_cast<const uint4*>(a.res + (long long)(ch0 >> 3) * a.res_plane_stride + (long long)pos * 8);
      const h16* rb = reinterpret_cast<const h16*>(&q);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = h162f(rb[i], a.fmt);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float x = v[i] + __ldg(a.shift + ch0 + i) + r[i];
      switch (a.act) {
        case kActRelu: x = fmaxf(x, 0.f); break;
        case kActLeaky: x = x > 0.f ? x : 0.1f * x; break;
        case kActHeadPaf: x = (sigmoidf_(x) - 0.5f) * 4.f; break;      // rtpose_light3d.py:335,337
        case kActHeadHeat: x = sigmoidf_(x); break;                    // rtpose_light3d.py:336
        default: break;
      }
      if (a.head_out && ch0 + i < a.cout) {
        const int H = a.Hs - 1, W = a.Wp - 1;
        a.head_out[(((long long)pi.n * a.cout + ch0 + i) * H + pi.h) * W + pi.w] = x;
      }
      if (ch0 + i >= a.cout) x = 0.f;                                  // padded channels stay zero
      ob[i] = f2h16(x, a.fmt);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) ob[i] = 0;
  }
  if (a.out)
    *reinterpret_cast<uint4*>(a.out + (long long)(ch0 >> 3) * a.out_plane_stride + (long long)pos * 8) =
        *reinterpret_cast<const uint4*>(ob);
}

// hot-path epilogue of the hidden layers: 8 channels (one C8P vector) of one position.
// FMT (operand format) and MODE (0 = ReLU, 1 = slope multiply: LeakyReLU 0.1 / identity 1) are compile-time so that the
// warp-uniform choices cost no predicated-off instructions; adds and multiplies run as packed f32x2 operations
// (add.rn.f32x2 / mul.rn.f32x2, sm_100): same fp32 results, half the issue slots.
template <int FMT, int MODE>
__device__ __forceinline__ uint4 finish8_fast(const uint32_t* acc, const float* s_shift8, bool has_res, uint4 res,
                                              float slope, bool interior) {
  const float4 s0 = *reinterpret_cast<const float4*>(s_shift8), s1 = *reinterpret_cast<const float4*>(s_shift8 + 4);
  float2 x[4];
  x[0] = __fadd2_rn(make_float2(__uint_as_float(acc[0]), __uint_as_float(acc[1])), make_float2(s0.x, s0.y));
  x[1] = __fadd2_rn(make_float2(__uint_as_float(acc[2]), __uint_as_float(acc[3])), make_float2(s0.z, s0.w));
  x[2] = __fadd2_rn(make_float2(__uint_as_float(acc[4]), __uint_as_float(acc[5])), make_float2(s1.x, s1.y));
  x[3] = __fadd2_rn(make_float2(__uint_as_float(acc[6]), __uint_as_float(acc[7])), make_float2(s1.z, s1.w));
  if (has_res) {
    const uint32_t rw[4] = {res.x, res.y, res.z, res.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = __fadd2_rn(x[i], unpack2(rw[i], FMT));
  }
  if (MODE == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = make_float2(fmaxf(x[i].x, 0.f), fmaxf(x[i].y, 0.f));
  } else {
    const float2 sl = make_float2(slope, slope);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 y = __fmul2_rn(x[i], sl);                          // slope in (0,1]: LeakyReLU 0.1, identity 1
      x[i] = make_float2(fmaxf(x[i].x, y.x), fmaxf(x[i].y, y.y));
    }
  }
  uint4 o;
  o.x = pack2(x[0].x, x[0].y, FMT); o.y = pack2(x[1].x, x[1].y, FMT); o.z = pack2(x[2].x, x[2].y, FMT); o.w = pack2(x[3].x, x[3].y, FMT);
  if (!interior) o = make_uint4(0u, 0u, 0u, 0u);                     // zero cells stay zero
  return o;
}

// ------------------------------------------------------------------------------------------------
// tensor-core implicit GEMM
// ------------------------------------------------------------------------------------------------
constexpr int kNumSMs = 148;           // B200
constexpr int kTcThreads = 640;       // 4 control warps + 16 epilogue warps
constexpr int kEpiWarps = 16;
constexpr int kEpiWarp0 = 4;

// Programmatic dependent launch: every kernel of the forward lets its successor be scheduled early
// (launch_dependents) and touches global memory only after its predecessor has completed (wait), so the
// successor's launch latency and prologue (barrier init, TMEM allocation) overlap the predecessor's tail.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Timeline tracing (tools/forward_timeline.py): per launch three globaltimer stamps, reduced over the CTAs with atomics.
// `tr` is nullptr in normal operation.
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_min(unsigned long long* tr, int slot) {
  if (tr && threadIdx.x == 0) atomicMin(tr + slot, gtime());
}
__device__ __forceinline__ void trace_max(unsigned long long* tr, int slot) {
  if (tr && threadIdx.x == 0) atomicMax(tr + slot, gtime());
}
// slot 3: sum over the CTAs of (exit - entry) = the SM time the launch occupied, whatever queued in front of its CTAs
__device__ __forceinline__ unsigned long long trace_enter(unsigned long long* tr) { return (tr && threadIdx.x == 0) ? gtime() : 0ull; }
__device__ __forceinline__ void trace_exit(unsigned long long* tr, unsigned long long t_in) {
  if (tr && threadIdx.x == 0) atomicAdd(tr + 3, gtime() - t_in);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent kernel: one CTA per SM walks the output tiles (tile = MT = NACC*128 positions x NT channels).
// Three pipelines, all mbarrier-based:
//   A ring  (a_full / a_empty, `ast` stages)  : one 64-input-channel slab of the tile + halo, all 9 taps read it
//   B ring  (b_full / b_empty, BST stages)    : one (tap, 64-channel) weight slab
//   accumulators (acc_full / acc_empty, AS)   : AS = 2 lets the epilogue of tile i run under the MMAs of tile i+1
template <int NT, int NACC>
constexpr int acc_stages() { return (2 * NACC * NT <= 512) ? 2 : 1; }

//   BRES: the layer has a single 64-channel chunk and all of its taps' weights stay resident in shared memory
//         for the whole kernel (loaded once): the MMA thread then never waits on the B ring, which matters
//         because the tensor core's instruction queue is shallow -- every instruction the issuing thread
//         spends between two MMAs is exposed (measured: tools/umma_bench.cu).
//   DBG:  bring-up instantiation with clock64 probes and debug switches; the product path compiles them out.
//   MC:   clusters of two CTAs share the weight stream: each CTA loads half of every B stage and multicasts it to both, a
//         stage is released when BOTH CTAs' MMAs have consumed it (commit multicast to both b_empty barriers).  Halves
//         the L2 weight traffic per output, which is what limits how small a tile may be: with it the N = 256 layers
//         run 128-position tiles with double-buffered accumulators (epilogue hidden) instead of 256-position tiles
//         whose two accumulators fill the TMEM.  Both CTAs walk the same number of tiles (the last may be a dummy
//         that only keeps the B ring in step).
// Layer chain (CHAIN): several layers of the same geometry run side by side inside ONE launch, each on its own slice of
// the CTAs, as a spatial pipeline: the CTAs of layer r+1 consume the tiles of layer r a few tiles behind their
// production, so a tensor travels from producer to consumer through the L2 instead of making a round trip through HBM
// (the 105 MB tensors of the 112 x 112 block do not survive in the 126 MB L2 from one launch to the next).
//   done[t]  (this layer, global memory)  raised to kEpiWarps when physical tile t is stored: every epilogue warp arrives on a
//            shared-memory mbarrier after its stores; the publisher warp (warp 3) waits for it, then proxy fence +
//            __threadfence + red.release.gpu -- the fence's round trip to the L2 stays out of the epilogue warps' path
//   wait[t]  (= done[] of the producing layer)  the copy thread takes tile t only when tiles t-1, t, t+1 (its halo reaches
//            at most one tile into either neighbour: halo <= MT) are complete: ld.acquire.gpu spin, then a proxy fence in
//            front of the bulk copies (generic-proxy stores of another SM -> async-proxy reads).
// Write-after-read on the ping-pong buffers is covered by the same condition: layer r+1 overwrites tile t of a buffer
// only after the tiles t-1 .. t+1 of layer r, the only readers of those positions, have completed.
// Spins are bounded (kChainTimeoutNs): a consumer gives up, sets *err and runs on with whatever is there -- a scheduling
// surprise must not hang the GPU.  All CTAs of the launch are co-resident (one per SM) and start in blockIdx order, so
// producers never wait for an SM held by their own consumers.
struct ChainLink {
  const unsigned int* wait;     // done[] of the producing layer, nullptr for the first layer of the chain
  unsigned int* done;           // this layer's counters [num_tiles]
  unsigned int* err;            // set to 1 when a wait timed out
};
constexpr unsigned long long kChainTimeoutNs = 200ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int NT, int NACC, int TAPS, int BST, bool BRES, bool DBG, bool MC, bool CHAIN>
__device__ __forceinline__ void conv_tc_body(const ConvArgs& a, uint8_t* smem, const int cta, const int ncta, const ChainLink link) {
  static_assert(!(MC && BRES), "multicast applies to the B ring");
  static_assert(!(MC && CHAIN), "chains run the single-CTA kernel");
  constexpr int MT = NACC * 128;
  constexpr int AS = acc_stages<NT, NACC>();
  constexpr int kBSlots = BRES ? TAPS : BST;                 // B slabs held in shared memory
  constexpr uint32_t kBStageBytes = 8u * NT * 16u;          // 64 input channels x NT output channels
  constexpr uint32_t kAccCols = NACC * NT;
  constexpr uint32_t kCols = (AS * kAccCols <= 32) ? 32 : (AS * kAccCols <= 64) ? 64 : (AS * kAccCols <= 128) ? 128
                             : (AS * kAccCols <= 256) ? 256 : 512;
  static_assert(AS * kAccCols <= 512, "accumulators exceed TMEM");
  constexpr int kNumBars = 4 + 2 * BST + 2 * AS + (CHAIN ? 2 : 0);
  const int halo = (TAPS == 9) ? a.Wp + 1 : 0;
  const int apos = MT + 2 * halo;                            // positions per staged plane
  const uint32_t a_plane_bytes = (uint32_t)apos * 16u;
  const uint32_t a_stage_bytes = 8u * a_plane_bytes;
  const int ast = a.a_stages;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (((size_t)ast * a_stage_bytes + 127) & ~(size_t)127);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)kBSlots * kBStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  float* s_shift = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(bars + kNumBars + 1) + 15) & ~(uintptr_t)15);   // [NT], 16 B aligned
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (4 + BST + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (4 + 2 * BST + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (4 + 2 * BST + AS + s); };
  auto st_done = [&](int s) { return bar0 + 8u * (4 + 2 * BST + 2 * AS + s); };      // CHAIN: a tile's outputs are stored (2, alternating)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (a.P + MT - 1) / MT;
  // MC: every CTA walks the same number of tile slots (the B ring of a cluster must stay in step); slots past the last
  // tile are dummies: their B stages are loaded and released, nothing else happens
  const int tile_end = MC ? (int)((num_tiles + ncta - 1) / ncta * ncta) : num_tiles;
  const uint32_t cta_rank = MC ? cluster_ctarank() : 0u;
  long long* probe = (DBG && a.probe) ? a.probe + (long long)cta * 16 : nullptr;
  const int dbg = DBG ? a.dbg : 0;
  if (probe && threadIdx.x == 0) probe[0] = clock64();
  trace_min(a.trace, 0);
  const unsigned long long t_enter = trace_enter(a.trace);
  const int chunks = a.chunks;
  const int chunks_all = chunks + (BRES ? 0 : a.chunks2);
  const int k8_total = chunks * 8;
  // first fill issued from the prologue (see below): all taps' weights when they are resident, else the first kPreB tap
  // stages of chunk 0, and the first input slab of the CTA's first tile (the grid never exceeds the tile count)
  constexpr bool kPrefill = !MC && !CHAIN && !DBG;
  constexpr int kPreB = (BST < TAPS) ? BST : TAPS;
  pdl_launch_dependents();

  if (warp == 0 && elect_one()) {       // (elect.sync, not `threadIdx.x == 0`: the bulk copies below are uniform-datapath instructions)
    for (int i = 0; i < kNumBars; ++i) {
      uint32_t cnt = 1u;
      if (i >= 4 + 2 * BST + AS) cnt = (uint32_t)kEpiWarps;                    // acc_empty (and st_done): one arrive per epilogue warp
      else if (MC && i >= 4 + BST && i < 4 + 2 * BST) cnt = 2u;                // b_empty: the MMA threads of both CTAs
      mbar_init(bar0 + 8u * i, cnt);
    }
    fence_mbar_init();
    if (kPrefill && !a.no_prefill && cta < num_tiles) {
      // Early first fill, by the thread that initialised the barriers, while the rest of the CTA loads the shift vector,
      // allocates TMEM and meets at the barrier below: a single-tile CTA of a 28 x 28 stage layer lives 12-17 us, of which the
      // L2 latency + transfer of its first operands (after that barrier) was more than one.  Weights first -- they do not
      // depend on the previous layer, so they go out BEFORE griddepcontrol.wait --, then the first input slab.
      if (BRES) {
        mbar_expect_tx(b_full(0), (uint32_t)TAPS * kBStageBytes);
        for (int t = 0; t < TAPS; ++t)
          bulk_g2s(smem_u32(sB + (size_t)t * kBStageBytes), a.w + (long long)t * k8_total * NT * 8, kBStageBytes, b_full(0));
      } else {
#pragma unroll
        for (int t = 0; t < kPreB; ++t) {
          mbar_expect_tx(b_full(t), kBStageBytes);
          bulk_g2s(smem_u32(sB + (size_t)t * kBStageBytes), a.w + (long long)t * k8_total * NT * 8, kBStageBytes, b_full(t));
        }
      }
      pdl_wait();
      const int t0 = (a.reverse ? num_tiles - 1 - cta : cta) * MT;
      mbar_expect_tx(a_full(0), a_stage_bytes);
      for (int g = 0; g < 8; ++g)
        bulk_g2s(smem_u32(sA + (size_t)g * a_plane_bytes), a.in + (long long)g * a.in_plane_stride + (long long)(t0 - halo) * 8,
                 a_plane_bytes, a_full(0));
    }
  }
  if (MC) cluster_sync_all();       // the peer's multicast copies / commits may reach this CTA's barriers from here on
  for (int i = threadIdx.x; i < NT; i += kTcThreads) s_shift[i] = a.shift[i];
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), kCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // activations of the previous layer are complete and visible from here on
  trace_min(a.trace, 1);
  if (probe && threadIdx.x == 0) probe[1] = clock64();

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- producer ----------------
      int ia = 0, ib = 0;
      bool chain_gave_up = false;
      if (BRES && !(kPrefill && !a.no_prefill && cta < num_tiles)) {   // all taps of the (single) chunk, once
        mbar_expect_tx(b_full(0), (uint32_t)TAPS * kBStageBytes);
        for (int t = 0; t < TAPS; ++t)
          bulk_g2s(smem_u32(sB + (size_t)t * kBStageBytes), a.w + (long long)t * k8_total * NT * 8, kBStageBytes, b_full(0));
      }
      for (int tile = cta; tile < tile_end; tile += ncta) {
        const int t0 = ((a.reverse && tile < num_tiles) ? num_tiles - 1 - tile : tile) * MT;
        const bool live = tile < num_tiles;                 // (always true without MC)
        if (CHAIN && link.wait != nullptr) {
          // the producing layer's tiles under this tile and its halo are complete (see ChainLink)
          const int pt = t0 / MT;
          unsigned long long ts = 0;
          for (int q = (pt > 0 ? pt - 1 : 0); q <= pt + 1 && q < num_tiles && !chain_gave_up; ++q) {
            while (ld_acquire_gpu_u32(link.wait + q) < (unsigned)kEpiWarps) {
              const unsigned long long now = gtime();
              if (ts == 0) ts = now;
              if (now - ts > kChainTimeoutNs) { atomicExch(link.err, 1u); chain_gave_up = true; break; }
              __nanosleep(64);
            }
          }
          fence_proxy_async_all();
        }
        for (int c = 0; c < chunks_all; ++c) {
          const bool ex = c >= chunks;                      // extra 1x1 chunk of the fused shortcut
          if (live) {
            const h16* src = ex ? a.in2 + (long long)((c - chunks) * 8) * a.in2_plane_stride
                                : a.in + (long long)(c * 8) * a.in_plane_stride;
            const long long pstride = ex ? a.in2_plane_stride : a.in_plane_stride;
            const int as = ia % ast;
            if (ia >= ast) mbar_wait_relaxed(a_empty(as), ((ia / ast) - 1) & 1);
            if (kPrefill && !a.no_prefill && ia == 0) {
              // (issued from the prologue)
            } else if ((dbg & 16) && ia >= ast) mbar_arrive(a_full(as));      // tuning: no copy traffic after the first fill
            else {
              mbar_expect_tx(a_full(as), a_stage_bytes);
              for (int g = 0; g < 8; ++g)
                bulk_g2s(smem_u32(sA + (size_t)as * a_stage_bytes + (size_t)g * a_plane_bytes),
                         src + (long long)g * pstride + (long long)(t0 - halo) * 8, a_plane_bytes, a_full(as));
            }
            ++ia;
          }
          const int ntap = ex ? 1 : TAPS;
          for (int t = 0; t < ntap && !BRES; ++t, ++ib) {
            const int bs = ib % BST;
            if (ib >= BST) mbar_wait_relaxed(b_empty(bs), ((ib / BST) - 1) & 1);
            if (kPrefill && !a.no_prefill && ib < kPreB) {
              // (issued from the prologue: chunk 0, taps 0 .. kPreB-1 of the first tile)
            } else if ((dbg & 16) && ib >= BST) mbar_arrive(b_full(bs));
            else {
              const long long woff = ex ? ((long long)TAPS * k8_total + (c - chunks) * 8) * NT * 8
                                        : ((long long)t * k8_total + c * 8) * NT * 8;
              mbar_expect_tx(b_full(bs), kBStageBytes);
              if (MC) {     // this CTA's half (input-channel groups 4*rank .. 4*rank+3) to both CTAs; the peer sends the other
                bulk_g2s_mc(smem_u32(sB + (size_t)bs * kBStageBytes) + cta_rank * (kBStageBytes / 2),
                            a.w + woff + (long long)cta_rank * (kBStageBytes / 4), kBStageBytes / 2, b_full(bs), (uint16_t)3);
              } else {
                bulk_g2s(smem_u32(sB + (size_t)bs * kBStageBytes), a.w + woff, kBStageBytes, b_full(bs));
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc = umma_idesc(NT, a.fmt);
      const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB);
      // descriptor halves: hi = SBO (128 B between 8-row groups) | version 1; lo = start address | LBO << 16
      // (LBO = byte distance between the two K core matrices of one K=16 instruction), all in 16-byte units
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t a_lo0 = ((sA0 >> 4) & 0x3FFFu) | (((a_plane_bytes >> 4) & 0x3FFFu) << 16);
      const uint32_t b_lo0 = ((sB0 >> 4) & 0x3FFFu) | ((((uint32_t)NT * 16u >> 4) & 0x3FFFu) << 16);
      const uint32_t a_kstep = (2u * a_plane_bytes) >> 4;       // two K core matrices per instruction
      constexpr uint32_t b_kstep = (2u * NT * 16u) >> 4;
      long long wait_a = 0, wait_b = 0, wait_acc = 0;
      // The tensor core's instruction queue is shallow: whatever this thread does between two MMAs is exposed.
      // Hence (1) no divisions / 64-bit math in the loop, (2) the B-ring wait of the NEXT tap is taken before the
      // MMAs of the current tap are issued (always deadlock-free: that slot was released by a tap already issued).
      int ti = 0;
      int as = 0, bs = 0;                         // A / B ring positions and their barrier phases
      uint32_t aph = 0, bph = 0;
      bool b_ready = false;                       // b_full(bs) already observed
      if (BRES) { mbar_wait(b_full(0), 0); b_ready = true; }
      // B stage consumed: arrive on its empty barrier when the MMAs issued so far retire (MC: in both CTAs of the cluster)
      auto release_b = [&](int stage) {
        if (MC) umma_commit_mc(b_empty(stage), (uint16_t)3);
        else umma_commit(b_empty(stage));
      };
      for (int tile = cta; tile < tile_end; tile += ncta) {
        if (MC && tile >= num_tiles) {
          // dummy slot: keep the shared B ring in step with the peer -- take every stage and release it again
          for (int c = 0; c < chunks_all; ++c) {
            const int ntap = c >= chunks ? 1 : TAPS;
            for (int t = 0; t < ntap; ++t) {
              if (!b_ready) mbar_wait(b_full(bs), bph);
              b_ready = false;
              release_b(bs);
              if (++bs == BST) { bs = 0; bph ^= 1; }
            }
          }
          continue;
        }
        const int s = ti % AS;
        if (ti >= AS) {
          long long tw = probe ? clock64() : 0;
          mbar_wait(acc_empty(s), ((ti / AS) - 1) & 1);
          if (probe) wait_acc += clock64() - tw;
        }
        const uint32_t tmem_acc = tmem_base + (uint32_t)s * kAccCols;
        uint32_t accumulate = 0;
        for (int c = 0; c < chunks_all; ++c) {
          const bool ex = c >= chunks;            // extra 1x1 chunk: centre tap only
          long long tw = probe ? clock64() : 0;
          mbar_wait(a_full(as), aph);
          if (probe) wait_a += clock64() - tw;
          const uint32_t a_lo_stage = a_lo0 + ((as * a_stage_bytes) >> 4);
          // The issuing thread runs at most about one MMA ahead of the tensor core, so every instruction between two
          // MMAs beyond ~1 MMA time is exposed.  Hence: taps fully unrolled (their A shifts fold to constants), the
          // release of the previous tap's B stage and the look-ahead probe of the next one are issued BETWEEN the first
          // MMAs of the current tap (hidden behind their execution) instead of at the tap boundary.
          int prev_bs = -1;                        // B stage whose release (commit) is still owed
          auto tap_body = [&](const int shift, const int t) {
            int nbs = bs + 1;
            uint32_t nbph = bph;
            if (!BRES) {
              if (!b_ready) {
                tw = probe ? clock64() : 0;
                mbar_wait(b_full(bs), bph);
                if (probe) wait_b += clock64() - tw;
              }
              if (nbs == BST) { nbs = 0; nbph ^= 1; }
              tc_fence_after();
            } else if (t == 0) {
              tc_fence_after();
            }
            const uint32_t a_lo_tap = a_lo_stage + (uint32_t)shift;          // one position = one 16-byte unit
            const uint32_t b_lo_tap = b_lo0 + (((BRES ? t : bs) * kBStageBytes) >> 4);
            if (!(dbg & 1)) {
              static_for<4>([&](auto kk_c) {
                constexpr int kk = decltype(kk_c)::value;
                const uint32_t a_lo_k = a_lo_tap + kk * a_kstep;
                static_for<NACC>([&](auto acc_c) {
                  constexpr int acc = decltype(acc_c)::value;
                  umma_off<acc * 128, kk * (int)b_kstep, acc * NT>(tmem_acc, a_lo_k, desc_hi, b_lo_tap, desc_hi, idesc,
                                                                   (kk == 0) ? accumulate : 1u);
                  if (!BRES && kk == 0 && acc == 0 && prev_bs >= 0) release_b(prev_bs);   // (also covers the MMA above)
                  if (!BRES && (NACC > 1 ? (kk == 0 && acc == 1) : (kk == 1))) {
                    // look ahead: next B stage (also across chunk / tile boundaries); one probe only -- if it is not there
                    // yet, block at the top of the next tap.  Deadlock-free: that slot was released by a tap already issued
                    // or is released by the commit just above.
                    b_ready = mbar_try_wait(b_full(nbs), nbph);
                  }
                });
              });
            } else if (!BRES) {
              if (prev_bs >= 0) release_b(prev_bs);
              b_ready = mbar_try_wait(b_full(nbs), nbph);
            }
            accumulate = 1u;
            prev_bs = bs;
            bs = nbs; bph = nbph;
          };
          if (ex || TAPS == 1) {
            tap_body(halo, 0);
          } else {
            // resident weights: one kernel row (36 MMAs) per unrolled body -- all 108 at once make the compiler spill
            // uniform registers between the MMAs (measured 70 vs 56 cycles per MMA)
            int row0 = halo - a.Wp;
#pragma unroll(BRES ? 1 : 3)
            for (int r = 0; r < 3; ++r, row0 += a.Wp) {
#pragma unroll
              for (int d = 0; d < 3; ++d) tap_body(row0 + d - 1, r * 3 + d);
            }
          }
          if (!BRES) release_b(prev_bs);                 // last tap of the chunk
          umma_commit(a_empty(as));
          if (++as == ast) { as = 0; aph ^= 1; }
        }
        umma_commit(acc_full(s));
        ++ti;
      }
      if (probe) { probe[3] = wait_a; probe[4] = wait_b; probe[5] = clock64(); probe[10] = wait_acc; probe[9] = ti; }
    }
  } else if (CHAIN && warp == 3) {
    // ---------------- publisher (CHAIN): when all epilogue warps have stored a tile, make it visible GPU-wide and count it.
    // A warp of its own: the fence waits for the stores to reach the L2, which must not sit in the epilogue's path.
    if (elect_one()) {
      int ti = 0;
      for (int tile = cta; tile < num_tiles; tile += ncta, ++ti) {
        const int t0 = (a.reverse ? num_tiles - 1 - tile : tile) * MT;
        mbar_wait_relaxed(st_done(ti & 1), (ti >> 1) & 1);
        fence_proxy_async_all();
        __threadfence();
        red_release_gpu_add(link.done + t0 / MT, (unsigned)kEpiWarps);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ---------------- epilogue: TMEM lane quarter q, thread = one output position; the two warps of a
    // quarter split the column chunks between them (a lone warp per scheduler is issue-latency bound) ----
    const int q = warp & 3;
    const int sub = (warp - kEpiWarp0) >> 2;            // 0..3: the four warps of a lane quarter split the work items
    const bool head = a.act == kActHeadPaf || a.act == kActHeadHeat;
    const float slope = a.act == kActRelu ? 0.f : (a.act == kActLeaky ? 0.1f : 1.f);
    const uint32_t mHs = div_magic(a.Hs), mWp = div_magic(a.Wp);
    const int epi_variant = (a.fmt != 0 ? 2 : 0) + (slope == 0.f ? 0 : 1);
    long long wait_full = 0, busy = 0;
    int ti = 0;
    for (int tile = cta; tile < num_tiles; tile += ncta, ++ti) {
      const int s = ti % AS;
      const int t0 = ((a.reverse && tile < num_tiles) ? num_tiles - 1 - tile : tile) * MT;
      long long tw = probe ? clock64() : 0;
      // Balanced split for the 64-channel, three-accumulator tiles of the 112 x 112 layers (6 work items of 32 channels do not
      // divide over the 4 warps of a lane quarter: two warps would do twice the work of the others and the epilogue, not
      // the MMAs, would set the tile time): warp (q, sub) owns channels [16 sub, 16 sub + 16) of all three accumulators.
      // The residual vectors of the tile are requested BEFORE the wait for its accumulators, so their latency runs under
      // the MMAs.  (CHAIN: the residual was written during this launch -- it is complete once this tile's A slab has
      // landed, which a single non-blocking probe of its a_full barrier tells; if the probe fails the loads happen after
      // the wait as before.  The probe cannot report a slab that has not landed: see the phase argument in DESIGN.md.)
      constexpr bool kSplit16 = (NT == 64 && NACC == 3 && BRES);
      uint4 rs16[kSplit16 ? 6 : 1];
      bool res_early = false;
      if constexpr (kSplit16) {
        if (a.res != nullptr && !head) {
          res_early = !CHAIN || mbar_try_wait(a_full(ti % ast), (uint32_t)((ti / ast) & 1));
          if (res_early) {
#pragma unroll
            for (int acc = 0; acc < 3; ++acc) {
              const int pos = t0 + acc * 128 + q * 32 + lane;
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                const uint4* rp = reinterpret_cast<const uint4*>(a.res + (long long)(sub * 2 + g) * a.res_plane_stride + (long long)pos * 8);
                rs16[acc * 2 + g] = pos < a.P ? (CHAIN ? __ldcg(rp) : *rp) : make_uint4(0u, 0u, 0u, 0u);
              }
            }
          }
        }
      }
      mbar_wait_relaxed(acc_full(s), (ti / AS) & 1);
      tc_fence_after();
      long long tb = probe ? clock64() : 0;
      if (probe) wait_full += tb - tw;
      if constexpr (kSplit16) {
        if (!head) {
          if (a.res != nullptr && tile + ncta < num_tiles && (lane & 7) == 0) {      // next tile's residual -> L2
            const int tn0 = (a.reverse ? num_tiles - 1 - (tile + ncta) : tile + ncta) * MT;
#pragma unroll
            for (int acc = 0; acc < 3; ++acc) {
              const int pos = tn0 + acc * 128 + q * 32 + lane;
              if (pos < a.P) {
#pragma unroll
                for (int g = 0; g < 2; ++g)
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + (long long)(sub * 2 + g) * a.res_plane_stride + (long long)pos * 8));
              }
            }
          }
#pragma unroll
          for (int acc = 0; acc < 3; ++acc) {
            const int pos = t0 + acc * 128 + q * 32 + lane;
            const PosInfo pi = c8p_locate_fast(pos, a.P, a.Hs, a.Wp, mHs, mWp);
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kAccCols + acc * NT + sub * 16), r);
            const bool has_res = a.res != nullptr && pi.interior;
            if (a.res != nullptr && !res_early && pi.in_range) {
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                const uint4* rp = reinterpret_cast<const uint4*>(a.res + (long long)(sub * 2 + g) * a.res_plane_stride + (long long)pos * 8);
                rs16[acc * 2 + g] = CHAIN ? __ldcg(rp) : *rp;
              }
            }
            tmem_ld_wait();
            if (pi.in_range && !(dbg & 2)) {
              auto store2 = [&](auto fmt_c, auto mode_c) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                  const uint4 o = finish8_fast<decltype(fmt_c)::value, decltype(mode_c)::value>(
                      r + g * 8, s_shift + sub * 16 + g * 8, has_res, rs16[acc * 2 + g], slope, pi.interior);
                  *reinterpret_cast<uint4*>(a.out + (long long)(sub * 2 + g) * a.out_plane_stride + (long long)pos * 8) = o;
                }
              };
              using std::integral_constant;
              switch (epi_variant) {
                case 0: store2(integral_constant<int, 0>{}, integral_constant<int, 0>{}); break;
                case 1: store2(integral_constant<int, 0>{}, integral_constant<int, 1>{}); break;
                case 2: store2(integral_constant<int, 1>{}, integral_constant<int, 0>{}); break;
                default: store2(integral_constant<int, 1>{}, integral_constant<int, 1>{}); break;
              }
            }
          }
        }
      }
      if (kSplit16 && !head) {
        // (done above)
      } else if (head) {
        // output heads (NT = 16 / 32): fp32 NCHW maps with the sigmoid scaling, optional 16-bit copy
        constexpr int kItems = NACC * (NT / 16);
#pragma unroll 1
        for (int it = sub; it < kItems; it += 4) {
          const int acc = it / (NT / 16), j = it - acc * (NT / 16);
          const int pos = t0 + acc * 128 + q * 32 + lane;
          const PosInfo pi = c8p_locate_fast(pos, a.P, a.Hs, a.Wp, mHs, mWp);
          uint32_t r[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kAccCols + acc * NT + j * 16), r);
          tmem_ld_wait();
          float v[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[h * 8 + i]);
            if (!(dbg & 2)) finish8(a, pos, pi, j * 16 + h * 8, v);
          }
        }
      } else if constexpr (NT >= 32) {
        constexpr int kItems = NACC * (NT / 32);
        // Residual layers: this warp's residual vectors of its NEXT tile are pulled into L2 now (one 128-byte line per
        // 8 lanes), one epilogue ahead of their use -- the residual tensor was written two layers ago and has left the
        // L2; without this the epilogue of the 64-channel residual layers waits on DRAM and outlasts the MMA phase.
        if (a.res != nullptr && tile + ncta < num_tiles && (lane & 7) == 0) {
          const int tn0 = (a.reverse ? num_tiles - 1 - (tile + ncta) : tile + ncta) * MT;
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
            for (int g = 0; g < 4; ++g) {
              const uint4* rp = reinterpret_cast<const uint4*>(a.res + (long long)(plane + g) * a.res_plane_stride + (long long)pos * 8);
              rs[g] = CHAIN ? __ldcg(rp) : *rp;       // chains: written by another SM during this launch -> read at the L2
            }
          }
          tmem_ld_wait();
          if (pi.in_range && !(dbg & 2)) {
            auto store4 = [&](auto fmt_c, auto mode_c) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 o = finish8_fast<decltype(fmt_c)::value, decltype(mode_c)::value>(
                    r + g * 8, s_shift + j * 32 + g * 8, has_res, rs[g], slope, pi.interior);
                *reinterpret_cast<uint4*>(a.out + (long long)(plane + g) * a.out_plane_stride + (long long)pos * 8) = o;
              }
            };
            using std::integral_constant;
            switch (epi_variant) {                    // warp-uniform: (operand format, ReLU | slope)
              case 0: store4(integral_constant<int, 0>{}, integral_constant<int, 0>{}); break;
              case 1: store4(integral_constant<int, 0>{}, integral_constant<int, 1>{}); break;
              case 2: store4(integral_constant<int, 1>{}, integral_constant<int, 0>{}); break;
              default: store4(integral_constant<int, 1>{}, integral_constant<int, 1>{}); break;
            }
          }
        }
      }
      // all TMEM reads of this accumulator stage are complete (tcgen05.wait::ld): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      // CHAIN: this warp's part of the tile is stored (all lanes: __syncwarp above).  Tell the publisher warp -- BEFORE the
      // accumulator is handed back, so that no warp can run two tiles (the two st_done barriers) ahead of the slowest one.
      if (CHAIN && lane == 0) mbar_arrive(st_done(ti & 1));
      if (lane == 0) mbar_arrive(acc_empty(s));
      if (probe) busy += clock64() - tb;
    }
    if (probe && threadIdx.x == kEpiWarp0 * 32) { probe[6] = wait_full; probe[7] = busy; }
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();        // the peer may still multicast into this CTA's shared memory / barriers until it is done too
  if (probe && threadIdx.x == 0) probe[8] = clock64();
  trace_max(a.trace, 2);
  trace_exit(a.trace, t_enter);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kCols);
  }
}

template <int NT, int NACC, int TAPS, int BST, bool BRES, bool DBG, bool MC = false>
__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_kernel(const ConvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  conv_tc_body<NT, NACC, TAPS, BST, BRES, DBG, MC, false>(a, smem, (int)blockIdx.x, (int)gridDim.x, ChainLink{nullptr, nullptr, nullptr});
}

// A chain of up to kMaxChain single-chunk 3x3 layers with resident weights and identical geometry (the four 64 -> 64
// convolutions of the 112 x 112 block): CTA slice r of the grid runs layer r (see ChainLink).
constexpr int kMaxChain = 4;
struct ChainArgs {
  ConvArgs base;                        // geometry, activation, operand format: common to all layers
  const h16* in[kMaxChain];
  const h16* w[kMaxChain];
  const float* shift[kMaxChain];
  h16* out[kMaxChain];
  const h16* res[kMaxChain];            // nullptr = no residual
  unsigned long long* trace[kMaxChain];
  unsigned int* flags;                  // [1 + nlayers * num_tiles]: word 0 = error, then done[] per layer; zeroed before the launch
  int nlayers;
};

template <int NT, int NACC>
__global__ void __launch_bounds__(kTcThreads, 1) conv_chain_kernel(const ChainArgs c) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int per = (int)gridDim.x / c.nlayers;
  const int role = (int)blockIdx.x / per;               // host: gridDim.x is a multiple of nlayers
  ConvArgs a = c.base;
#pragma unroll
  for (int i = 0; i < kMaxChain; ++i)
    if (role == i) { a.in = c.in[i]; a.w = c.w[i]; a.shift = c.shift[i]; a.out = c.out[i]; a.res = c.res[i]; a.trace = c.trace[i]; }
  const int num_tiles = (a.P + NACC * 128 - 1) / (NACC * 128);
  ChainLink link;
  link.err = c.flags;
  link.done = c.flags + 1 + (size_t)role * num_tiles;
  link.wait = role > 0 ? c.flags + 1 + (size_t)(role - 1) * num_tiles : nullptr;
  conv_tc_body<NT, NACC, 9, 2, true, false, false, true>(a, smem, (int)blockIdx.x - role * per, per, link);
}

// ------------------------------------------------------------------------------------------------
// CUDA-core check kernel: one thread = one position x 8 output channels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_simt_kernel(const ConvArgs a) {
  const int pos = blockIdx.x * 128 + threadIdx.x;
  const int ch0 = blockIdx.y * 8;
  const PosInfo pi = locate(pos, a);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (pi.interior) {
    const int k8_total = a.chunks * 8;
    const int nti = ch0 / a.nt, nn = ch0 - nti * a.nt;
    for (int t = 0; t < a.taps; ++t) {
      const int shift = (a.taps == 9) ? ((t / 3 - 1) * a.Wp + (t % 3 - 1)) : 0;
      for (int g = 0; g < k8_total; ++g) {
        const uint4 xa = *reinterpret_cast<const uint4*>(a.in + (long long)g * a.in_plane_stride + (long long)(pos + shift) * 8);
        const h16* xb = reinterpret_cast<const h16*>(&xa);
        const h16* wrow = a.w + ((((long long)nti * a.taps + t) * k8_total + g) * a.nt + nn) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 wa = *reinterpret_cast<const uint4*>(wrow + i * 8);
          const h16* wb = reinterpret_cast<const h16*>(&wa);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i] = fmaf(h162f(xb[j], a.fmt), h162f(wb[j], a.fmt), acc[i]);
        }
      }
    }
    for (int g = 0; g < a.chunks2 * 8; ++g) {              // fused 1x1 shortcut on the second input
      const uint4 xa = *reinterpret_cast<const uint4*>(a.in2 + (long long)g * a.in2_plane_stride + (long long)pos * 8);
      const h16* xb = reinterpret_cast<const h16*>(&xa);
      const h16* wrow = a.w + (((long long)a.taps * k8_total + g) * a.nt + nn) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 wa = *reinterpret_cast<const uint4*>(wrow + i * 8);
        const h16* wb = reinterpret_cast<const h16*>(&wa);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i] = fmaf(h162f(xb[j], a.fmt), h162f(wb[j], a.fmt), acc[i]);
      }
    }
  }
  finish8(a, pos, pi, ch0, acc);
}

// (the opt-in CTA-pair kernel, tcgen05.mma.cta_group::2, lives in conv_pair_experimental.cuh: validated, bit-identical,
//  measured slower inside the pipelined step -- POPNET_TUNE_PAIR)
#include "conv_pair_experimental.cuh"

// ------------------------------------------------------------------------------------------------
// stem: model0.conv1 (7x7, stride 2, pad 3, one input channel) + folded BN + ReLU -> C8P (64 channels).
// Also on the tensor cores: each CTA im2col's 128 output positions into the K-major core-matrix layout
// (K = 8 input rows x 8 input columns = 64, of which the 7x7 kernel uses 49), issues M128 x N64 x K16 MMAs against the
// resident weights and runs the same TMEM epilogue; the input image is read through L1 (every pixel feeds ~12 taps).
// INPUT PRECISION: the depth frame is the one fp32 tensor of the forward, and rounding it to 16 bits was measured to be
// 85 % of the forward's whole output error (tools/e2e_sensitivity.py --skip input: max-abs 3.6e-3 -> 5.4e-4 with fp16
// operands; the frame is large flat areas whose features are small differences of O(1) values, which an 11-bit mantissa
// erases).  The im2col therefore splits every pixel into hi = rn16(x) and lo = rn16(x - hi) and the tile is multiplied
// twice against the same weights into the same accumulator (K = 64 hi + 64 lo): 22 (fp16) / 16 (bf16) mantissa bits of the
// input reach the fp32 accumulator, for four more N = 64 MMAs per 128 positions (0.6 % of the forward's FLOPs).
// ------------------------------------------------------------------------------------------------
constexpr int kStemThreads = 128;
#ifndef POPNET_STEM_BATCHES
#define POPNET_STEM_BATCHES 2
#endif
constexpr int kStemBatches = POPNET_STEM_BATCHES;   // rows loaded per batch = 8 / kStemBatches: 2 -> 16 float2 loads in flight per thread
constexpr int kStemCtasPerSm = 5;                   // 40.3 KB of shared memory per CTA (hi + lo tiles, weights)

__global__ void __launch_bounds__(kStemThreads, kStemCtasPerSm) stem_kernel(const StemArgs a) {
  __shared__ __align__(128) h16 sA[2 * 8 * 128 * 8];      // [hi | lo][k8][row][8]
  __shared__ __align__(128) h16 sB[8 * 64 * 8];           // [k8][cout][8]
  __shared__ __align__(16) float s_shift[64];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ho = a.H / 2, Wo = a.W / 2, Hs = Ho + 1, Wp = Wo + 1;
  const int P = (int)c8p_positions(a.N, Ho, Wo);
  const int tiles = (P + 127) / 128;
  for (int i = tid; i < 8 * 64; i += kStemThreads)
    reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(a.w)[i];
  if (tid < 64) s_shift[tid] = a.shift[tid];
  const uint32_t bar = smem_u32(&s_bar);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&s_tmem), 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  const uint32_t idesc = umma_idesc(64, a.fmt);
  uint32_t phase = 0;
  const uint32_t mHs = div_magic(Hs), mWp = div_magic(Wp);
  trace_min(a.trace, 0);
  const unsigned long long t_enter = trace_enter(a.trace);
  pdl_launch_dependents();
  pdl_wait();
  trace_min(a.trace, 1);
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int pos = (a.reverse ? tiles - 1 - tile : tile) * 128 + tid;
    const PosInfo pi = c8p_locate_fast(pos, P, Hs, Wp, mHs, mWp);
    const int n = pi.n;
    const bool interior = pi.interior;
    // ---- im2col: K = 8 input rows x 8 input columns (rows 2oy-3 .. 2oy+4, columns 2ox-4 .. 2ox+3; the 7x7 kernel
    // occupies rows 0..6 / columns 1..7, the extra row and column carry zero weights).  One k8 group = one input
    // row = four aligned float2 loads, packed straight into the 16-byte operand vector.
    {
      const int iy0 = pi.h * 2 - 3, ix0 = pi.w * 2 - 4;
      // one base pointer (possibly outside the image: only dereferenced under its predicate), four column predicates
      // shared by all rows, one unsigned compare per row
      const float* p0 = a.x + ((long long)n * a.H + iy0) * a.W + ix0;
      bool cok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) cok[j] = interior && (unsigned)(ix0 + 2 * j) < (unsigned)a.W;   // even: the pair is in or out together
      // kStemBatches batches of rows: all loads of a batch are issued before their first use
      constexpr int RB = 8 / kStemBatches;
#pragma unroll
      for (int half = 0; half < kStemBatches; ++half) {
        float2 v[RB][4];
#pragma unroll
        for (int r4 = 0; r4 < RB; ++r4) {
          const int ry = half * RB + r4;
          const bool rok = (unsigned)(iy0 + ry) < (unsigned)a.H;
          const float* row = p0 + ry * a.W;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            v[r4][j] = (rok && cok[j]) ? __ldg(reinterpret_cast<const float2*>(row + 2 * j)) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int r4 = 0; r4 < RB; ++r4) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            hi[j] = pack2(v[r4][j].x, v[r4][j].y, a.fmt);
            const float2 h = unpack2(hi[j], a.fmt);                     // what the 16-bit value holds: the rest goes into lo
            lo[j] = pack2(v[r4][j].x - h.x, v[r4][j].y - h.y, a.fmt);
          }
          reinterpret_cast<uint4*>(sA)[(half * RB + r4) * 128 + tid] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          reinterpret_cast<uint4*>(sA)[(8 + half * RB + r4) * 128 + tid] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
    // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (tid < 32 && elect_one()) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)       // K steps 0-3: hi tile, 4-7: lo tile, both against the same weights
        umma_bf16(tmem_base, umma_desc(a0 + (2 * kk) * 128 * 16, 128 * 16, 128), umma_desc(b0 + (2 * (kk & 3)) * 64 * 16, 64 * 16, 128),
                  idesc, kk != 0 ? 1u : 0u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: thread = position (TMEM lane), 64 channels
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32), r);
      tmem_ld_wait();
      if (pos < P) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 o = a.fmt == 0
              ? finish8_fast<0, 0>(r + g * 8, s_shift + j * 32 + g * 8, false, make_uint4(0, 0, 0, 0), 0.f, interior)
              : finish8_fast<1, 0>(r + g * 8, s_shift + j * 32 + g * 8, false, make_uint4(0, 0, 0, 0), 0.f, interior);
          *reinterpret_cast<uint4*>(a.out + (long long)(j * 4 + g) * a.out_plane_stride + (long long)pos * 8) = o;
        }
      }
    }
    tc_fence_before();
    __syncthreads();          // TMEM and sA are reused by the next tile
    tc_fence_after();
  }
  __syncthreads();
  trace_max(a.trace, 2);
  trace_exit(a.trace, t_enter);
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------------
// AvgPool2d(3, stride 2, pad 1), divisor always 9; the zero ring of the input IS the padding
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pool_kernel(const PoolArgs a) {
  const int Ho = a.H / 2, Wo = a.W / 2, Wpi = a.W + 1, Hsi = a.H + 1;
  const int P = (int)c8p_positions(a.N, Ho, Wo);
  const int pos = (a.reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * 128 + threadIdx.x;
  trace_min(a.trace, 0);
  pdl_launch_dependents();
  pdl_wait();
  trace_min(a.trace, 1);
  trace_max(a.trace, 2);              // (entry of the last CTA; a pool CTA lives ~1 us)
  if (pos >= P) return;
  const int g = blockIdx.y;
  const PosInfo pi = c8p_locate(pos, P, Ho + 1, Wo + 1);
  __align__(16) h16 ob[8];
  if (pi.interior) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // output (oy, ox) covers input rows 2oy-1 .. 2oy+1, cols 2ox-1 .. 2ox+1; the zero row above / below an image and
    // the zero cell that ends every row ARE the padding (col -1 of a row is the zero cell of the row before it)
    const h16* src = a.in + (long long)g * a.in_plane_stride;
    const long long row0 = 2 + (long long)pi.n * Hsi + 2 * pi.h - 1;
    const int col0 = 2 * pi.w - 1;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint4 q = *reinterpret_cast<const uint4*>(src + ((row0 + dy) * Wpi + col0 + dx) * 8);
        const h16* qb = reinterpret_cast<const h16*>(&q);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += h162f(qb[i], a.fmt);
      }
#pragma unroll
    for (int i = 0; i < 8; ++i) ob[i] = f2h16(acc[i] * (1.f / 9.f), a.fmt);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) ob[i] = 0;
  }
  *reinterpret_cast<uint4*>(a.out + (long long)g * a.out_plane_stride + (long long)pos * 8) = *reinterpret_cast<const uint4*>(ob);
}

unsigned long long* next_trace_slot(int tag) {
  if (!g_trace_buf) return nullptr;
  const int i = g_trace_next.fetch_add(1);
  if (i >= g_trace_cap || i >= 1024) return nullptr;
  g_trace_tags[i] = tag;
  return g_trace_buf + 4 * (size_t)i;
}

constexpr size_t kSmemLimit = 227 * 1024;

// launch configuration with programmatic stream serialization (PDL) enabled; the attribute lives in the caller's frame
struct PdlConfig {
  cudaLaunchAttribute attr[2];
  cudaLaunchConfig_t cfg;
  // cluster2: launch as clusters of two CTAs (the kernel itself does not use the cluster; the grid must be even)
  PdlConfig(dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool cluster2 = false) : cfg{} {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = cluster2 ? 2 : 1;
  }
  PdlConfig(const PdlConfig&) = delete;
};

template <int NT, int NACC, int TAPS, int BST, bool BRES, bool DBG>
int launch_tc_inst(const ConvArgs& a, size_t smem, cudaStream_t st) {
  auto kern = conv_tc_kernel<NT, NACC, TAPS, BST, BRES, DBG>;
  POPNET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // ask for the full shared-memory carveout so that two CTAs can be co-resident where their tiles allow it
  POPNET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int MT = NACC * 128;
  if (a.cout_pad != NT) return POPNET_ERR_UNSUPPORTED;     // one N tile per layer (true for every rtpose layer)
  const int tiles = (a.P + MT - 1) / MT;
  const int cap = (a.grid_cap > 0 && a.grid_cap < kNumSMs) ? a.grid_cap : kNumSMs;
  int grid = tiles < cap ? tiles : cap;                      // persistent: one CTA per SM walks the tiles
  if (a.balance && tiles > cap) {
    // equal tile counts: ceil(tiles / cap) tiles in every CTA (211 tiles: 106 CTAs x 2 instead of 63 x 2 + 85 x 1) -- the layer
    // takes as long either way, but the SMs a launch does not need are free for the concurrent branches of the stage
    const int per = (tiles + cap - 1) / cap;
    grid = (tiles + per - 1) / per;
  }
  // tuning (POPNET_TUNE_CLUSTER_ALL): every stage launch as clusters of two, so that SM pairs are taken and released together
  // and the multicast kernels of the concurrent PAF branch always find whole pairs; an odd grid gets one CTA without tiles
  if (a.cluster2) grid = (grid + 1) & ~1;
  PdlConfig pc(dim3(grid), dim3(kTcThreads), smem, st, a.cluster2 != 0);
  ConvArgs at = a;
  at.trace = next_trace_slot(NT * 1000 + NACC * 100 + TAPS * 10);
  POPNET_CUDA_TRY(cudaLaunchKernelEx(&pc.cfg, kern, at));
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}

template <int NT, int NACC, int TAPS, bool BRES>
int launch_tc_bst(const ConvArgs& a, int bst, size_t smem, cudaStream_t st) {
  const bool dbgk = a.probe != nullptr || a.dbg != 0;
  if (BRES) {
    if (a.chunks != 1) return POPNET_ERR_UNSUPPORTED;
    return dbgk ? launch_tc_inst<NT, NACC, TAPS, 2, BRES, true>(a, smem, st) : launch_tc_inst<NT, NACC, TAPS, 2, BRES, false>(a, smem, st);
  }
  if (dbgk) return launch_tc_inst<NT, NACC, TAPS, 4, BRES, true>(a, smem, st);      // bring-up: 4 stages only
  switch (bst) {
    case 2: return launch_tc_inst<NT, NACC, TAPS, 2, BRES, false>(a, smem, st);
    case 3: return launch_tc_inst<NT, NACC, TAPS, 3, BRES, false>(a, smem, st);
    default: return launch_tc_inst<NT, NACC, TAPS, 4, BRES, false>(a, smem, st);
  }
}

}  // namespace


size_t conv_tc_smem_bytes(int nt, int nacc, int taps, int a_stages, int Wp, int* b_stages_out, bool b_resident) {
  const int halo = taps == 9 ? Wp + 1 : 0;
  const size_t a_bytes = (((size_t)a_stages * 8 * (nacc * 128 + 2 * halo) * 16) + 127) & ~(size_t)127;
  const size_t b_stage = (size_t)8 * nt * 16;
  const size_t misc = 256 + (size_t)nt * 4;
  if (b_resident) {
    if (b_stages_out) *b_stages_out = taps;
    return a_bytes + taps * b_stage + misc;
  }
  int bst = 4;
  while (bst > 2 && a_bytes + bst * b_stage + misc > kSmemLimit) --bst;
  if (b_stages_out) *b_stages_out = bst;
  return a_bytes + bst * b_stage + misc;
}

namespace {
// cluster-of-two launch of the multicast variant (programmatic dependent launch + cluster dimension)
template <int NT, int NACC, int TAPS, int BST>
int launch_tc_mc(const ConvArgs& a, cudaStream_t st) {
  auto kern = conv_tc_kernel<NT, NACC, TAPS, BST, false, false, true>;
  const int halo = TAPS == 9 ? a.Wp + 1 : 0;
  const size_t a_bytes = (((size_t)a.a_stages * 8 * (NACC * 128 + 2 * halo) * 16) + 127) & ~(size_t)127;
  const size_t smem = a_bytes + (size_t)BST * 8 * NT * 16 + 256 + (size_t)NT * 4;
  if (smem > kSmemLimit || a.cout_pad != NT || a.chunks2 != 0) return POPNET_ERR_UNSUPPORTED;
  POPNET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  POPNET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int tiles = (a.P + NACC * 128 - 1) / (NACC * 128);
  const int cap = (a.grid_cap > 0 && a.grid_cap < kNumSMs) ? (a.grid_cap & ~1) : kNumSMs;
  int grid = tiles < cap ? tiles : cap;
  grid = (grid + 1) & ~1;                                      // whole clusters (148 is even)
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = 2; attrs[1].val.clusterDim.y = 1; attrs[1].val.clusterDim.z = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cfg.attrs = attrs; cfg.numAttrs = 2;
  ConvArgs at = a;
  at.trace = next_trace_slot(NT * 1000 + NACC * 100 + TAPS * 10 + 5);
  POPNET_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, at));
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}
}  // namespace

#define POPNET_PAIR_LAUNCH_PART
#include "conv_pair_experimental.cuh"
#undef POPNET_PAIR_LAUNCH_PART

int launch_conv_tc(const ConvArgs& a, int nacc, cudaStream_t st) {
  // CTA-pair kernel (cta_group::2) for the large single-chunk 64 -> 64 3x3 layers: opt-in with PopnetNetConfig.tuning
  // POPNET_TUNE_PAIR(4) (512-position tiles) or (3) (384).  Measured, same box, A/B/A/B: the two non-residual layers go from
  // 57 to 51 us and the forward ALONE from 0.994 to 0.984 ms -- but inside the pipelined step, where the decode of the
  // previous batch shares the SMs, the step gets 1.6 % SLOWER (1.035 vs 1.019 ms: a cluster needs both SMs of a pair free at
  // once), so it is off by default.  The residual layers are HBM-bound (210 MB read + 105 MB written in 66 us -- the 105 MB
  // tensors do not survive in the L2 from one layer to the next) and the pair's coupled accumulator release makes them
  // 4 us slower (POPNET_TUNE_PAIR_RES includes them).
  const int pair = a.pair;
  const bool pair_res = a.pair_res != 0;
  if (pair && !a.mc && a.nt == 64 && a.taps == 9 && a.chunks == 1 && a.chunks2 == 0 && a.head_out == nullptr &&
      a.cout_pad == 64 && a.out != nullptr && a.probe == nullptr && a.dbg == 0 && (pair_res || a.res == nullptr) &&
      a.P >= 148 * 512) {
    const int rc = pair == 3 ? launch_pair64<3>(a, st) : launch_pair64<4>(a, st);
    if (rc != POPNET_ERR_UNSUPPORTED) return rc;
  }
  if (a.mc) {
    if (a.probe != nullptr || a.dbg != 0) return POPNET_ERR_UNSUPPORTED;
    if (a.nt == 256 && nacc == 1 && a.taps == 9) return launch_tc_mc<256, 1, 9, 5>(a, st);
    return POPNET_ERR_UNSUPPORTED;
  }
  int bst = 0;
  // weights resident in shared memory whenever the layer is a single chunk and everything fits
  // (64 output channels: 72 KB of weights; 128 output channels: 144 KB, next to two 128-position A stages -- the 64 -> 128
  //  layer at 56 x 56, which with streamed weights and 256-position tiles pulled 40 B per cycle and SM from the L2: the
  //  chip-wide L2 limit, 28 us instead of the 13 us its MMAs take)
  const bool bres = a.chunks == 1 && a.chunks2 == 0 && a.taps == 9 && (a.nt == 64 || (a.nt == 128 && nacc == 1)) &&
                    conv_tc_smem_bytes(a.nt, nacc, a.taps, a.a_stages, a.Wp, nullptr, true) <= kSmemLimit;
  const size_t smem = conv_tc_smem_bytes(a.nt, nacc, a.taps, a.a_stages, a.Wp, &bst, bres);
  if (smem > kSmemLimit) return POPNET_ERR_UNSUPPORTED;
  if ((a.probe != nullptr || a.dbg != 0) && !bres && bst != 4) return POPNET_ERR_UNSUPPORTED;
#define POPNET_TC_CASE(NT_, NACC_, TAPS_, BRES_) \
  if (a.nt == NT_ && nacc == NACC_ && a.taps == TAPS_ && bres == BRES_) return launch_tc_bst<NT_, NACC_, TAPS_, BRES_>(a, bst, smem, st);
  POPNET_TC_CASE(64, 2, 9, true)
  POPNET_TC_CASE(64, 3, 9, true)
  POPNET_TC_CASE(64, 4, 9, true)
  POPNET_TC_CASE(128, 1, 9, true)
  POPNET_TC_CASE(64, 2, 9, false)
  POPNET_TC_CASE(64, 4, 9, false)
  POPNET_TC_CASE(128, 2, 9, false)
  POPNET_TC_CASE(128, 4, 9, false)
  POPNET_TC_CASE(128, 3, 9, false)
  POPNET_TC_CASE(128, 2, 1, false)
  POPNET_TC_CASE(32, 2, 1, false)
  POPNET_TC_CASE(16, 2, 9, false)
  POPNET_TC_CASE(128, 3, 1, false)
  POPNET_TC_CASE(64, 3, 9, false)
  POPNET_TC_CASE(32, 3, 1, false)
  POPNET_TC_CASE(16, 3, 9, false)
  POPNET_TC_CASE(128, 4, 1, false)
  POPNET_TC_CASE(256, 2, 9, false)
  POPNET_TC_CASE(32, 4, 1, false)
  POPNET_TC_CASE(16, 4, 9, false)
#undef POPNET_TC_CASE
  return POPNET_ERR_UNSUPPORTED;
}

size_t conv_chain_flag_words(int P, int nacc, int nlayers) {
  const int tiles = (P + nacc * 128 - 1) / (nacc * 128);
  return 1 + (size_t)nlayers * tiles;
}

// One launch for `n` consecutive single-chunk 64 -> 64 3x3 layers of the same geometry (ChainArgs / ChainLink above).
// `flags`: conv_chain_flag_words() words of device memory, zeroed in stream order before this launch.
int launch_conv_chain(const ConvArgs* layers, int n, int nacc, unsigned int* flags, cudaStream_t st) {
  if (n < 2 || n > kMaxChain || !flags) return POPNET_ERR_INVALID_ARG;
  const ConvArgs& a0 = layers[0];
  for (int i = 0; i < n; ++i) {
    const ConvArgs& a = layers[i];
    if (a.nt != 64 || a.cout_pad != 64 || a.taps != 9 || a.chunks != 1 || a.chunks2 != 0 || a.head_out || !a.out || a.mc ||
        a.probe || a.dbg || a.P != a0.P || a.Hs != a0.Hs || a.Wp != a0.Wp || a.act != a0.act || a.fmt != a0.fmt ||
        a.in_plane_stride != a0.in_plane_stride || a.out_plane_stride != a0.out_plane_stride ||
        (a.res && a.res_plane_stride != a0.out_plane_stride) || a.reverse != a0.reverse)
      return POPNET_ERR_UNSUPPORTED;
  }
  if (a0.Wp + 1 > nacc * 128) return POPNET_ERR_UNSUPPORTED;          // the halo must stay inside the neighbouring tile
  const size_t smem = conv_tc_smem_bytes(64, nacc, 9, 2, a0.Wp, nullptr, true);
  if (smem > kSmemLimit) return POPNET_ERR_UNSUPPORTED;
  ChainArgs c{};
  c.base = a0;
  c.base.a_stages = 2;
  c.base.res_plane_stride = a0.out_plane_stride;
  for (int i = 0; i < n; ++i) {
    c.in[i] = layers[i].in; c.w[i] = layers[i].w; c.shift[i] = layers[i].shift; c.out[i] = layers[i].out; c.res[i] = layers[i].res;
    c.trace[i] = next_trace_slot(64 * 1000 + nacc * 100 + 90 + 1 + i);
  }
  c.flags = flags;
  c.nlayers = n;
  const int grid = kNumSMs / n * n;                                 // equal CTA slices (4 x 37 on 148 SMs)
  auto launch = [&](auto kern) -> int {
    POPNET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    POPNET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    PdlConfig pc(dim3(grid), dim3(kTcThreads), smem, st);
    POPNET_CUDA_TRY(cudaLaunchKernelEx(&pc.cfg, kern, c));
    POPNET_AFTER_LAUNCH();
    return POPNET_OK;
  };
  if (nacc == 3) return launch(conv_chain_kernel<64, 3>);
  if (nacc == 2) return launch(conv_chain_kernel<64, 2>);
  return POPNET_ERR_UNSUPPORTED;
}

int launch_conv_simt(const ConvArgs& a, cudaStream_t st) {
  dim3 grid((a.P + 127) / 128, a.cout_pad / 8);
  conv_simt_kernel<<<grid, 128, 0, st>>>(a);
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}

int launch_stem(const StemArgs& a, cudaStream_t st) {
  const int P = (int)c8p_positions(a.N, a.H / 2, a.W / 2);
  const int tiles = (P + 127) / 128;
  const int cps = kStemCtasPerSm;       // CTAs per SM (shared memory: 5 x 40.3 KB)
  // (a function attribute is per device context: set on every call, it is a host-side table write)
  POPNET_CUDA_TRY(cudaFuncSetAttribute(stem_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int grid = tiles < 148 * cps ? tiles : 148 * cps;      // persistent: `cps` CTAs per SM walk the tiles
  PdlConfig pc(dim3(grid), dim3(kStemThreads), 0, st);
  StemArgs at = a;
  at.trace = next_trace_slot(1);
  POPNET_CUDA_TRY(cudaLaunchKernelEx(&pc.cfg, stem_kernel, at));
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}

int launch_pool(const PoolArgs& a, cudaStream_t st) {
  const int P = (int)c8p_positions(a.N, a.H / 2, a.W / 2);
  dim3 grid((P + 127) / 128, a.planes);
  PdlConfig pc(grid, dim3(128), 0, st);
  PoolArgs at = a;
  at.trace = next_trace_slot(2);
  POPNET_CUDA_TRY(cudaLaunchKernelEx(&pc.cfg, pool_kernel, at));
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}

}  // namespace popnet

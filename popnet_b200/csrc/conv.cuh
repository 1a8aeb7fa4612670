// Shared declarations of the forward path (conv_kernels.cu: kernels + launchers, forward.cu: layer plan).
//
// ACTIVATION LAYOUT ("C8P"): a tensor of C channels (C % 8 == 0) over N images of H x W pixels is
// stored as C/8 planes; plane g holds, for every position of a zero-separated image stack, the 8 channels
// 8g .. 8g+7 as one 16-byte vector:
//
//     plane[g][ guard | P positions, rounded up to 512, + slack | guard ][8]   bf16 / fp16
//
// Positions are linear over rows of Wp = W + 1 cells: W pixels followed by ONE zero cell (it is the right
// neighbour of the row's last pixel and, through linearity, the left neighbour of the next row's first).
// Rows: two zero rows, then for every image its H pixel rows followed by ONE zero row (bottom neighbour of
// image n = top neighbour of image n+1): P = (2 + N * (H + 1)) * (W + 1).  Every producer writes the zero
// cells, which makes a 3x3 tap a pure shift by dh*Wp + dw positions: the A operand of tap (dh, dw) is the
// same shared-memory tile read through a UMMA descriptor whose start address is moved by that many
// 16-byte rows.  Outputs computed at zero / slack positions are garbage and are replaced by zeros or
// dropped in the epilogue.  Guards are never zeroed: rows of the MMA are independent, so whatever they
// hold only reaches zero-cell / slack outputs.  Work wasted on zero cells: 1.8 % at 112x112, 7.4 % at 28x28.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <atomic>

#include "common.cuh"

namespace popnet {

constexpr int kGuard = 128;        // positions in front of / behind every plane (>= W + 3)
constexpr int kPosRound = 512;     // planes hold a multiple of 512 positions ...
constexpr int kPosSlack = 512;     // ... plus slack so that 384-position tiles may overrun the last multiple

// 16-bit operand storage: bf16 (fmt 0, the north-star default) or fp16 (fmt 1).  Both run at the same
// tensor-core rate and byte count; fp16 carries 3 more mantissa bits ("operand format").
typedef uint16_t h16;
__host__ __device__ inline h16 f2h16(float x, int fmt) {
  if (fmt == 0) return __bfloat16_as_ushort(__float2bfloat16(x));
  return __half_as_ushort(__float2half(x));
}
__host__ __device__ inline float h162f(h16 v, int fmt) {
  if (fmt == 0) return __bfloat162float(__ushort_as_bfloat16(v));
  return __half2float(__ushort_as_half(v));
}

enum Act : int { kActNone = 0, kActRelu = 1, kActLeaky = 2, kActHeadPaf = 3, kActHeadHeat = 4 };

struct ConvArgs {
  const h16* in;                // position 0 of the first input plane
  long long in_plane_stride;    // 16-bit elements between planes
  const h16* in2;               // optional second input (same geometry) feeding `chunks2` extra 1x1 chunks: the
  long long in2_plane_stride;   //   fused projection shortcut of a residual block (K-concatenation), else nullptr
  const h16* w;                 // packed [tap][cin_pad/8][NT][8] (+ [chunks2*8][NT][8] for the extra chunks), scale folded in
  const float* shift;           // [cout_pad] folded BN shift + conv bias
  h16* out;                     // position 0 of the first output plane, or nullptr
  long long out_plane_stride;
  const h16* res;               // residual (same geometry as out) or nullptr
  long long res_plane_stride;
  float* head_out;              // fp32 [N][cout][H][W] or nullptr
  int P;                        // (2 + N * Hs) * Wp positions
  int Hs, Wp;                   // row period of an image (H + 1) and row pitch (W + 1)
  int chunks;                   // cin_pad / 64
  int chunks2;                  // extra 64-channel chunks read from in2 with the centre tap only
  int a_stages;                 // 1 or 2
  int act;                      // Act
  int cout;                     // logical output channels (head_out bound)
  int cout_pad;
  int nt;                       // N tile (template argument of the launched kernel)
  int taps;                     // 1 or 9
  int fmt;                      // 0 = bf16, 1 = fp16 operands
  long long* probe;             // optional [gridDim.x][16] clock64 stamps (bring-up / tuning), else nullptr
  int dbg;                      // bring-up switches: 1 = skip MMA issue, 2 = skip epilogue stores
  int reverse;                  // 1 = walk the tiles from the last to the first (zig-zag between consecutive layers: a layer then
                                //   starts on the positions its producer wrote last, which are the ones still in the L2)
  int mc;                       // 1 = cluster-of-two kernel with multicast weight stages (conv_kernels.cu, "MC")
  int pair, pair_res;           // tuning (PopnetNetConfig.tuning): CTA-pair kernel for the 64 -> 64 layers, 0 = off, 3 / 4 = tile size / 128
  unsigned long long* trace;    // optional [4] globaltimer stamps {first CTA in, first CTA past griddepcontrol.wait, last CTA out, sum of CTA lifetimes}
  int grid_cap;                 // persistent grid size limit (0 = all 148 SMs): tuning, leaves SMs to the concurrently running decode
  int balance;                  // tuning: size the persistent grid so that every CTA walks the same number of tiles
  int cluster2;                 // tuning: launch as clusters of two CTAs (see launch_tc_inst)
  int no_prefill;               // tuning: first operand loads from the producer loop (after the CTA-wide barrier) instead of the prologue
};

struct StemArgs {
  const float* x;               // [N][H][W] fp32
  const h16* w;                 // [8 kernel rows (8th zero)][64 cout][8 column slots (slot 0 zero)], scale folded
  const float* shift;           // [64]
  h16* out;                     // C8P, 64 channels at (H/2, W/2)
  long long out_plane_stride;
  int N, H, W;                  // input size
  int fmt;
  int reverse;
  unsigned long long* trace;
};

struct PoolArgs {
  const h16* in;
  long long in_plane_stride;
  h16* out;
  long long out_plane_stride;
  int planes;
  int N, H, W;                  // input spatial size (output is H/2 x W/2)
  int fmt;
  int reverse;
  unsigned long long* trace;
};

// timeline tracing state (conv_kernels.cu); set through popnet_debug_trace (forward.cu)
extern unsigned long long* g_trace_buf;
extern int g_trace_cap;
extern std::atomic<int> g_trace_next;
extern int g_trace_tags[1024];
int launch_conv_tc(const ConvArgs& a, int nacc, cudaStream_t st);
int launch_conv_simt(const ConvArgs& a, cudaStream_t st);
// n consecutive 64 -> 64 3x3 layers of one geometry as ONE launch (spatial pipeline through the L2, conv_kernels.cu "CHAIN")
int launch_conv_chain(const ConvArgs* layers, int n, int nacc, unsigned int* flags, cudaStream_t st);
size_t conv_chain_flag_words(int P, int nacc, int nlayers);
int launch_stem(const StemArgs& a, cudaStream_t st);
int launch_pool(const PoolArgs& a, cudaStream_t st);
size_t conv_tc_smem_bytes(int nt, int nacc, int taps, int a_stages, int Wp, int* b_stages_out, bool b_resident = false);

}  // namespace popnet

namespace popnet {
// position -> (image, row, col) of the C8P layout; interior = a real pixel (not a zero cell / slack)
struct PosInfo {
  bool in_range, interior;
  int n, h, w;
};
__host__ __device__ inline long long c8p_positions(int N, int H, int W) { return (long long)(2 + (long long)N * (H + 1)) * (W + 1); }
// Division by a loop-invariant divisor d (2 <= d < 2^16): q = umulhi(n, magic(d)) over-estimates n / d by at most one for
// every 32-bit n; one multiply-subtract and a conditional fix-up make it exact (4 instructions instead of ~20).
__host__ __device__ inline uint32_t div_magic(int d) { return 0xFFFFFFFFu / (uint32_t)d + 1u; }
__device__ __forceinline__ void fast_divmod(int n, int d, uint32_t magic, int& q, int& r) {
  q = (int)__umulhi((uint32_t)n, magic);
  r = n - q * d;
  if (r < 0) { --q; r += d; }
}
// same as c8p_locate with the two divisions by magic numbers (mWp = div_magic(Wp), mHs = div_magic(Hs))
__device__ __forceinline__ PosInfo c8p_locate_fast(int pos, int P, int Hs, int Wp, uint32_t mHs, uint32_t mWp) {
  PosInfo r;
  r.in_range = pos < P;
  int row, c;
  fast_divmod(pos, Wp, mWp, row, c);
  const int rr = row - 2;
  int n = 0, h = rr;
  if (rr >= 0) fast_divmod(rr, Hs, mHs, n, h);
  r.interior = r.in_range && rr >= 0 && c < Wp - 1 && h < Hs - 1;
  r.n = n; r.h = h; r.w = c;
  return r;
}
__device__ __forceinline__ PosInfo c8p_locate(int pos, int P, int Hs, int Wp) {
  PosInfo r;
  r.in_range = pos < P;
  const int row = pos / Wp, c = pos - row * Wp;
  const int rr = row - 2;
  const int n = rr >= 0 ? rr / Hs : 0, h = rr - n * Hs;
  r.interior = r.in_range && rr >= 0 && c < Wp - 1 && h < Hs - 1;
  r.n = n; r.h = h; r.w = c;
  return r;
}
}  // namespace popnet

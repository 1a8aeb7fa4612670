// Shared declarations of the forward path (conv_kernels.cu: kernels + launchers, forward.cu: layer plan).
//
// ACTIVATION LAYOUT ("C8P"): a tensor of C channels (C % 8 == 0) over N images of H x W pixels is
// stored as C/8 planes; plane g holds, for every position of the zero-padded images, the 8 channels
// 8g .. 8g+7 as one 16-byte vector:
//
//     plane[g][ guard | N * (H+2) * (W+2) positions, rounded up to 512 | guard ][8]   bf16 / fp16
//
// Positions are linear over (n, h_pad, w_pad).  The one-pixel ring around every image is ZERO (every
// producer writes it), which makes a 3x3 tap a pure shift by dh*(W+2)+dw positions: the A operand of
// tap (dh, dw) is the same shared-memory tile read through a UMMA descriptor whose start address is
// moved by that many 16-byte rows.  Outputs computed at ring / slack positions are garbage and are
// replaced by zeros (ring) or dropped (slack) in the epilogue.  Guards are never zeroed: rows of the
// MMA are independent, so whatever they hold only reaches ring / slack outputs.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

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
  const h16* w;                 // packed [n_tile][tap][cin_pad/8][NT][8], BN scale folded in
  const float* shift;           // [cout_pad] folded BN shift + conv bias
  h16* out;                     // position 0 of the first output plane, or nullptr
  long long out_plane_stride;
  const h16* res;               // residual (same geometry as out) or nullptr
  long long res_plane_stride;
  float* head_out;              // fp32 [N][cout][H][W] or nullptr
  int P;                        // N * Hp * Wp
  int Hp, Wp;                   // padded image size
  int chunks;                   // cin_pad / 64
  int a_stages;                 // 1 or 2
  int act;                      // Act
  int cout;                     // logical output channels (head_out bound)
  int cout_pad;
  int nt;                       // N tile (template argument of the launched kernel)
  int taps;                     // 1 or 9
  int fmt;                      // 0 = bf16, 1 = fp16 operands
  long long* probe;             // optional [gridDim.x][16] clock64 stamps (bring-up / tuning), else nullptr
  int dbg;                      // bring-up switches: 1 = skip MMA issue, 2 = skip epilogue stores
};

struct StemArgs {
  const float* x;               // [N][H][W] fp32
  const h16* w;                 // [8 k8][64 cout][8]: K = ky*7+kx (49 taps, zero padded to 64), scale folded
  const float* shift;           // [64]
  h16* out;                     // C8P, 64 channels at (H/2, W/2)
  long long out_plane_stride;
  int N, H, W;                  // input size
  int fmt;
};

struct PoolArgs {
  const h16* in;
  long long in_plane_stride;
  h16* out;
  long long out_plane_stride;
  int planes;
  int N, H, W;                  // input spatial size (output is H/2 x W/2)
  int fmt;
};

int launch_conv_tc(const ConvArgs& a, int nacc, cudaStream_t st);
int launch_conv_simt(const ConvArgs& a, cudaStream_t st);
int launch_stem(const StemArgs& a, cudaStream_t st);
int launch_pool(const PoolArgs& a, cudaStream_t st);
size_t conv_tc_smem_bytes(int nt, int nacc, int taps, int a_stages, int Wp, int* b_stages_out, bool b_resident = false);

}  // namespace popnet

// Heat-map / PAF decode and 2D->3D lift on the device.
//
// Reference being replaced (third_party_methods/): lib/utils/paf_to_pose.py:33-377 (find_peaks, NMS with
// cv2 bicubic refinement, find_connected_joints, group_limbs_of_same_person), lib/utils/common.py:5-32,
// 272-293 (paf_to_human_list, retrieve_depth_heat_weighted) and the per-frame glue of
// evaluate/evaluation_rtpose_light3d_kdh3d_mpreal_ablation.py:179-263.
//
// Three kernels per batch, all reading the network's channel-major fp32 maps in place; each is PERSISTENT with one WARP
// per work item and at most `max_ctas` CTAs (PopnetDecodeParams.max_ctas):
//   peaks_kernel     item = (frame, joint type): the map goes to the warp's shared-memory slice, 4-neighbour NMS with
//                    ordered compaction, then the 8x bicubic of the clipped 5x5 patch of every peak and its first arg-max.
//   limbs_kernel     item = (frame, limb): the limb's two PAF planes go to the warp's slice, every (src, dst) pair is scored
//                    from 10 on-the-fly bicubic PAF samples (the 224x224x28 upsample the reference materialises is never
//                    built), then greedy one-to-one matching.  Items whose score matrix exceeds the warp's pool are done by
//                    the whole CTA afterwards.
//   assemble_kernel  item = frame: sequential person assembly, pruning, depth lift, rescale and back-projection; writes the
//                    pose records (and pushes them to the peers).
// Why few, fat CTAs: the work per frame is tiny and latency-bound, and in the pipelined step the decode of batch i runs
// under the forward of batch i + 1, whose convolution CTAs fill an SM's register file -- every SM that holds a decode CTA
// at a layer boundary delays that layer.  With max_ctas = 8 the decode lives on the 8 SMs the conv grids leave free.
// Compiled with -fmad=false so fp32/fp64 expressions round exactly like OpenCV's C++ path and NumPy;
// the two places the reference goes through BLAS use explicit fma().
//
// Bound: HBM in principle (181,888 B of maps per frame); in practice latency.
#include <math_constants.h>

#include <algorithm>
#include <type_traits>
#include <cstring>

#include "common.cuh"

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxCells = 64 * 64;      // largest supported grid_h * grid_w

// OpenCV INTER_CUBIC at scale 8: phase r = dst % 8 -> first-tap offset and the four Keys(A=-0.75)
// weights, cv::interpolateCubic evaluated in fp32 at x = frac((r + 0.5) / 8 - 0.5).  All sixteen distinct values are
// integers / 2^14, exact in fp32 (oracle/decode_np.py::phase_table computes them; tests compare).  Statically
// initialised __constant__ data: present on every device of the process from module load, no upload call, no state.
#define POPNET_Q14(n) ((float)(n) / 16384.0f)
__constant__ int c_ofs[8] = {-1, -1, -1, -1, 0, 0, 0, 0};
__constant__ float c_coef[8][4] = {
    {POPNET_Q14(-1323), POPNET_Q14(8365), POPNET_Q14(11043), POPNET_Q14(-1701)},
    {POPNET_Q14(-825), POPNET_Q14(5615), POPNET_Q14(13409), POPNET_Q14(-1815)},
    {POPNET_Q14(-351), POPNET_Q14(3033), POPNET_Q14(15223), POPNET_Q14(-1521)},
    {POPNET_Q14(-45), POPNET_Q14(859), POPNET_Q14(16245), POPNET_Q14(-675)},
    {POPNET_Q14(-675), POPNET_Q14(16245), POPNET_Q14(859), POPNET_Q14(-45)},
    {POPNET_Q14(-1521), POPNET_Q14(15223), POPNET_Q14(3033), POPNET_Q14(-351)},
    {POPNET_Q14(-1815), POPNET_Q14(13409), POPNET_Q14(5615), POPNET_Q14(-825)},
    {POPNET_Q14(-1701), POPNET_Q14(11043), POPNET_Q14(8365), POPNET_Q14(-1323)}};
#undef POPNET_Q14

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// value of the 8x bicubic upsample of an H x W array (row pitch `pitch`) at integer point (X, Y):
// horizontal ((s0*a0+s1*a1)+s2*a2)+s3*a3 on the four source rows, then vertical r0*b0+(r1*b1+(r2*b2+r3*b3))
__device__ __forceinline__ float bicubic_at(const float* m, int pitch, int H, int W, int X, int Y) {
  const int rx = X & 7, ry = Y & 7;
  const int bx = (X >> 3) + c_ofs[rx], by = (Y >> 3) + c_ofs[ry];
  const float a0 = c_coef[rx][0], a1 = c_coef[rx][1], a2 = c_coef[rx][2], a3 = c_coef[rx][3];
  const int x0 = clampi(bx - 1, 0, W - 1), x1 = clampi(bx, 0, W - 1), x2 = clampi(bx + 1, 0, W - 1),
            x3 = clampi(bx + 2, 0, W - 1);
  float rows[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float* row = m + clampi(by - 1 + j, 0, H - 1) * pitch;
    rows[j] = ((row[x0] * a0 + row[x1] * a1) + row[x2] * a2) + row[x3] * a3;
  }
  return rows[0] * c_coef[ry][0] + (rows[1] * c_coef[ry][1] + (rows[2] * c_coef[ry][2] + rows[3] * c_coef[ry][3]));
}

// the same value for TWO maps of identical geometry (the x and y planes of a PAF limb) at one point: phases, clamped tap
// columns / rows and weights are computed once; per map the arithmetic is bicubic_at's, statement for statement
__device__ __forceinline__ void bicubic2_at(const float* mx, const float* my, int pitch, int H, int W, int X, int Y, float& vx,
                                            float& vy) {
  const int rx = X & 7, ry = Y & 7;
  const int bx = (X >> 3) + c_ofs[rx], by = (Y >> 3) + c_ofs[ry];
  const float a0 = c_coef[rx][0], a1 = c_coef[rx][1], a2 = c_coef[rx][2], a3 = c_coef[rx][3];
  const int x0 = clampi(bx - 1, 0, W - 1), x1 = clampi(bx, 0, W - 1), x2 = clampi(bx + 1, 0, W - 1),
            x3 = clampi(bx + 2, 0, W - 1);
  float rowsx[4], rowsy[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ro = clampi(by - 1 + j, 0, H - 1) * pitch;
    const float* rwx = mx + ro;
    const float* rwy = my + ro;
    rowsx[j] = ((rwx[x0] * a0 + rwx[x1] * a1) + rwx[x2] * a2) + rwx[x3] * a3;
    rowsy[j] = ((rwy[x0] * a0 + rwy[x1] * a1) + rwy[x2] * a2) + rwy[x3] * a3;
  }
  const float b0 = c_coef[ry][0], b1 = c_coef[ry][1], b2 = c_coef[ry][2], b3 = c_coef[ry][3];
  vx = rowsx[0] * b0 + (rowsx[1] * b1 + (rowsx[2] * b2 + rowsx[3] * b3));
  vy = rowsy[0] * b0 + (rowsy[1] * b1 + (rowsy[2] * b2 + rowsy[3] * b3));
}

// ------------------------------------------------------------------------------------------------
// D2 for ONE peak by one warp (paf_to_pose.py:96-118): the 8x bicubic of the clipped 5x5 patch around `cell` and the FIRST
// arg-max of the (<= 40 x 40) result in row-major order.
// Lane = output column (dx = lane; the last 8 columns of a 40-wide patch are spread over all lanes, see below): its horizontal
// phase rx = dx & 7 = lane & 7 never changes, so the four horizontal weights `hc` and the tap offset `hofs` are per-lane
// constants of the kernel; the (<= 5) horizontally interpolated values of every column go to the warp's scratch `tmp` [5][40]
// and a lane walks down its column: per source row cy the five scratch rows cy-2 .. cy+2 (clamped) give the 8 output rows
// 8 cy .. 8 cy + 7, whose vertical weights are compile-time constants after unrolling.  Same expressions and association as
// cv::resize (see bicubic_at), so every value is bit-identical to the full upsample's.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void refine_peak(const float* __restrict__ s_map, int W, int H, int cell, float* __restrict__ tmp,
                                            int lane, const float hc0, const float hc1, const float hc2, const float hc3,
                                            const int hofs, int& X, int& Y, float& score) {
  const int y = cell / W, x = cell - y * W;
  const int x0 = max(x - 2, 0), y0 = max(y - 2, 0), x1 = min(x + 2, W - 1), y1 = min(y + 2, H - 1);
  const int pw = x1 - x0 + 1, ph = y1 - y0 + 1, uw = pw * 8;
  const float* patch = s_map + y0 * W + x0;
  // horizontal pass: column dx of source row r -> tmp[r][dx]
  auto hpass = [&](int dx, int r_first, int r_step) {
    const int bx = (dx >> 3) + hofs;
    const int i0 = clampi(bx - 1, 0, pw - 1), i1 = clampi(bx, 0, pw - 1), i2 = clampi(bx + 1, 0, pw - 1), i3 = clampi(bx + 2, 0, pw - 1);
    for (int r = r_first; r < ph; r += r_step) {
      const float* row = patch + r * W;
      tmp[r * 40 + dx] = ((row[i0] * hc0 + row[i1] * hc1) + row[i2] * hc2) + row[i3] * hc3;
    }
  };
  // columns 0 .. 31: lane = column, all rows; columns 32 .. 39 (a full 5-wide patch): lane = 8 g + c -> column 32 + c (the
  // same phase c = lane & 7, hence the same weights), rows g, g + 4 -- all 32 lanes stay busy
  if (lane < uw) hpass(lane, 0, 1);
  if (uw > 32) hpass(32 + (lane & 7), lane >> 3, 4);
  __syncwarp();
  float best = -CUDART_INF_F;
  int bidx = 0x7fffffff;
  // vertical pass: the 8 output rows 8 cy .. 8 cy + 7 of column dx from the scratch rows cy - 2 .. cy + 2 (clamped)
  // (a lane meets the candidates of its first column in ascending index order, so a strict > keeps the first maximum there;
  //  the spread-out last columns come afterwards with smaller and larger indices: `ordered` = false compares the index too)
  auto vpass = [&](int dx, int cy_first, int cy_step, auto ordered) {
    for (int cy = cy_first; cy < ph; cy += cy_step) {
      const float tm2 = tmp[clampi(cy - 2, 0, ph - 1) * 40 + dx], tm1 = tmp[clampi(cy - 1, 0, ph - 1) * 40 + dx];
      const float t00 = tmp[cy * 40 + dx];
      const float tp1 = tmp[clampi(cy + 1, 0, ph - 1) * 40 + dx], tp2 = tmp[clampi(cy + 2, 0, ph - 1) * 40 + dx];
      int i = (8 * cy) * uw + dx;
#pragma unroll
      for (int ry = 0; ry < 8; ++ry, i += uw) {
        // ry < 4: first tap row cy - 2 (c_ofs = -1); ry >= 4: first tap row cy - 1
        const float r0 = ry < 4 ? tm2 : tm1, r1 = ry < 4 ? tm1 : t00, r2 = ry < 4 ? t00 : tp1, r3 = ry < 4 ? tp1 : tp2;
        const float v = r0 * c_coef[ry][0] + (r1 * c_coef[ry][1] + (r2 * c_coef[ry][2] + r3 * c_coef[ry][3]));
        if (decltype(ordered)::value ? (v > best) : (v > best || (v == best && i < bidx))) { best = v; bidx = i; }
      }
    }
  };
  if (lane < uw) vpass(lane, 0, 1, std::true_type{});
  if (uw > 32) vpass(32 + (lane & 7), lane >> 3, 4, std::false_type{});
#pragma unroll
  for (int ofs = 16; ofs > 0; ofs >>= 1) {
    const float ov = __shfl_xor_sync(kFull, best, ofs);
    const int oi = __shfl_xor_sync(kFull, bidx, ofs);
    if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
  }
  const int ay = bidx / uw, ax = bidx - ay * uw;
  X = 8 * x0 + ax; Y = 8 * y0 + ay; score = best;
  __syncwarp();                  // `tmp` is rewritten by the next peak
}

// ------------------------------------------------------------------------------------------------
// D1 + D2: peaks
// ------------------------------------------------------------------------------------------------
// per-warp shared-memory slice of peaks_kernel: [cells (padded to 4)] map | [5*40] scratch | [MP] peak cells
__host__ __device__ inline size_t peaks_warp_bytes(int cells, int MP) {
  return (((((size_t)cells + 3) & ~(size_t)3) + 200 + (size_t)MP) * 4 + 15) & ~(size_t)15;      // slices stay 16-byte aligned (float4 fills)
}

__global__ void __launch_bounds__(1024) peaks_kernel(const float* __restrict__ heat, int batch, PopnetDecodeParams p,
                                                     PopnetDecodeOut o) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const int H = p.grid_h, W = p.grid_w, cells = H * W, cells4 = (cells + 3) & ~3;
  const int K = p.num_joints, MP = p.max_peaks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  float* s_map = reinterpret_cast<float*>(s_dyn + (size_t)warp * peaks_warp_bytes(cells, MP));
  float* tmp = s_map + cells4;
  int* s_cell = reinterpret_cast<int*>(tmp + 200);
  const float hc0 = c_coef[lane & 7][0], hc1 = c_coef[lane & 7][1], hc2 = c_coef[lane & 7][2], hc3 = c_coef[lane & 7][3];
  const int hofs = c_ofs[lane & 7];
  const int items = batch * K;
  for (int it = blockIdx.x * wpc + warp; it < items; it += gridDim.x * wpc) {
    const int b = it / K, k = it - b * K;
    const float* m = heat + ((size_t)b * (K + 1) + k) * cells;
    if ((cells & 3) == 0 && (reinterpret_cast<uintptr_t>(m) & 15u) == 0) {
      const float4* m4 = reinterpret_cast<const float4*>(m);
      float4* d4 = reinterpret_cast<float4*>(s_map);
      for (int i = lane; i < cells / 4; i += 32) d4[i] = __ldg(m4 + i);
    } else {
      for (int i = lane; i < cells; i += 32) s_map[i] = m[i];
    }
    __syncwarp();
    // D1: peak cells in row-major order (ordered compaction by ballot)
    int cnt = 0;
    for (int base = 0; base < cells; base += 32) {
      const int i = base + lane;
      bool pk = false;
      if (i < cells) {
        const int y = i / W, x = i - y * W;
        const float v = s_map[i];
        pk = v > p.thresh_heat;
        if (pk && y > 0) pk = !(s_map[i - W] > v);
        if (pk && y < H - 1) pk = !(s_map[i + W] > v);
        if (pk && x > 0) pk = !(s_map[i - 1] > v);
        if (pk && x < W - 1) pk = !(s_map[i + 1] > v);
      }
      const unsigned bal = __ballot_sync(kFull, pk);
      if (pk) {
        const int slot = cnt + __popc(bal & ((1u << lane) - 1u));
        if (slot < MP) s_cell[slot] = i;
      }
      cnt += __popc(bal);
    }
    if (cnt > MP) {
      if (lane == 0) atomicOr(o.flags + b, POPNET_FLAG_PEAK_OVERFLOW);
      cnt = MP;
    }
    if (lane == 0) o.peak_count[(size_t)b * K + k] = cnt;
    __syncwarp();
    // D2: refinement, peak by peak
    for (int pi = 0; pi < cnt; ++pi) {
      int X, Y;
      float best;
      refine_peak(s_map, W, H, s_cell[pi], tmp, lane, hc0, hc1, hc2, hc3, hofs, X, Y, best);
      if (lane == 0) {
        const size_t slot = ((size_t)b * K + k) * MP + pi;
        *reinterpret_cast<int*>(o.peak_xy + slot * 2) = (X & 0xffff) | (Y << 16);
        o.peak_score[slot] = best;
      }
    }
    __syncwarp();                    // the slice is refilled by the warp's next item
  }
}

// ------------------------------------------------------------------------------------------------
// D4: limb scoring + greedy matching
// ------------------------------------------------------------------------------------------------
// round(linspace(a, b, n))[i] for integer a, b (paf_to_pose.py:214-217)
__device__ __forceinline__ int line_point(int a, int b, int i, int n) {
  if (n == 10) {
    const int num = 2 * (9 * a + i * (b - a)) + 9;
    int q = num / 18;
    if (num % 18 != 0 && num < 0) --q;
    return q;
  }
  if (n == 1) return a;
  const double step = ((double)b - (double)a) / (double)(n - 1);
  const double v = (i == n - 1) ? (double)b : (double)a + (double)i * step;
  return (int)rint(v);
}

// Scores of one limb's candidate pairs (paf_to_pose.py:196-239) by `nwarps` warps (this is warp `warp` of them);
// s_score[i * pitch + j] = score or -inf (not a candidate).  Two lane mappings, chosen per 32 pairs:
//   * full groups of 32 pairs: lane = pair.  The lane walks its NP intermediate points itself (two on-the-fly bicubic PAF
//     samples and the dot product with the unit vector each), summing them in NumPy's pairwise order on the fly (8 running
//     sums + tail, no per-point storage): the fewest instructions per pair -- what counts when the decode is limited to a
//     few SMs and a crowded frame has hundreds of pairs per limb;
//   * the remaining (< 32) pairs -- all of them for the typical 2 x 2 .. 6 x 6 candidates: lane = (pair, point), 32 / NP
//     pairs per round (3 for the reference's 10 points); every lane evaluates ONE point and parks it in the warp's scratch,
//     the pair's first lane sums.  One lane per pair would leave most lanes idle and make the 2 NP samples a serial chain.
__device__ __forceinline__ double pair_point(const float* __restrict__ s_px, const float* __restrict__ s_py, int W, int H,
                                             int ax, int ay, int bx, int by, double ux, double uy, int t, int NP, int body) {
  const int X = line_point(ax, bx, t, NP), Y = line_point(ay, by, t, NP);
  float fx, fy;
  bicubic2_at(s_px, s_py, W, H, W, X, Y, fx, fy);
  const double px = (double)fx, py = (double)fy;
  // ndarray.dot -> OpenBLAS dgemv: vector body fma(px,ux,py*uy), scalar tail fma(py,uy,px*ux)
  return (t < body) ? fma(px, ux, py * uy) : fma(py, uy, px * ux);
}
__device__ __forceinline__ double pair_finish(double sum, int above, int NP, double Hup, double dist) {
  double pen = 0.5 * Hup / dist - 1;
  if (!(pen < 0)) pen = 0;
  const double sc = sum / (double)NP + pen;
  return ((double)above > 0.8 * (double)NP && sc > 0) ? sc : -CUDART_INF;
}
__device__ __forceinline__ void score_pairs(const float* __restrict__ s_px, const float* __restrict__ s_py, int W, int H,
                                            const int16_t (*xa)[2], const int16_t (*xb)[2], int na, int nb, int NP,
                                            double thresh_paf, double Hup, double* __restrict__ s_score, int pitch,
                                            double* __restrict__ scratch, int warp, int nwarps, int lane) {
  const int npairs = na * nb;
  const int body = NP & ~3;
  const int full = npairs & ~31;
  // ---- lane = pair
  for (int base = warp * 32; base < full; base += nwarps * 32) {
    const int pr = base + lane;
    const int i = pr / nb, j = pr - i * nb;
    const int ax = xa[i][0], ay = xa[i][1], bx = xb[j][0], by = xb[j][1];
    const double dx = (double)bx - (double)ax, dy = (double)by - (double)ay;
    const double dist = sqrt(dx * dx + dy * dy) + 1e-8;
    const double ux = dx / dist, uy = dy / dist;
    int above = 0;
    double sum = 0.0;
    int t = 0;
    if (NP >= 8) {                                   // np.mean: NumPy pairwise sum = 8 running sums over the 8-blocks ...
      double r8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        r8[u] = pair_point(s_px, s_py, W, H, ax, ay, bx, by, ux, uy, u, NP, body);
        above += r8[u] > thresh_paf;
      }
      for (t = 8; t + 8 <= NP; t += 8)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double v = pair_point(s_px, s_py, W, H, ax, ay, bx, by, ux, uy, t + u, NP, body);
          above += v > thresh_paf;
          r8[u] += v;
        }
      sum = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]));
    }
    for (; t < NP; ++t) {                            // ... + tail (a plain loop below 8 elements)
      const double v = pair_point(s_px, s_py, W, H, ax, ay, bx, by, ux, uy, t, NP, body);
      above += v > thresh_paf;
      sum += v;
    }
    s_score[i * pitch + j] = pair_finish(sum, above, NP, Hup, dist);
  }
  // ---- lane = (pair, point)
  const int G = 32 / NP;                             // NP <= 32 (checked on the host): G >= 1
  const int pl = lane / NP, t = lane - pl * NP;
  for (int base = full + warp * G; base < npairs; base += nwarps * G) {
    const int pr = base + pl;
    const bool act = pl < G && pr < npairs;
    int i = 0, j = 0;
    double dist = 1.0, sv = 0.0;
    if (act) {
      i = pr / nb; j = pr - i * nb;
      const int ax = xa[i][0], ay = xa[i][1], bx = xb[j][0], by = xb[j][1];
      const double dx = (double)bx - (double)ax, dy = (double)by - (double)ay;
      dist = sqrt(dx * dx + dy * dy) + 1e-8;
      const double ux = dx / dist, uy = dy / dist;
      sv = pair_point(s_px, s_py, W, H, ax, ay, bx, by, ux, uy, t, NP, body);
    }
    scratch[lane] = sv;
    __syncwarp();
    if (act && t == 0) {
      const double* sp = scratch + pl * NP;
      int above = 0;
      for (int u = 0; u < NP; ++u) above += sp[u] > thresh_paf;
      double sum;
      if (NP < 8) {
        sum = 0.0;
        for (int u = 0; u < NP; ++u) sum += sp[u];
      } else {
        double r8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) r8[u] = sp[u];
        int u0 = 8;
        for (; u0 + 8 <= NP; u0 += 8)
#pragma unroll
          for (int u = 0; u < 8; ++u) r8[u] += sp[u0 + u];
        sum = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]));
        for (; u0 < NP; ++u0) sum += sp[u0];
      }
      s_score[i * pitch + j] = pair_finish(sum, above, NP, Hup, dist);
    }
    __syncwarp();
  }
}

struct Best { double v; int idx; };

// Shared-memory layout of limbs_kernel (bytes from the start of the dynamic block), computed on the host:
//   planes  [wpc][2][cells4] float    the limb's x / y PAF planes of each warp's current item
//   pool    [wpc][pool_doubles] double the warp's score matrix [na][nb]; all pools together hold one max_peaks^2 matrix
//   dot     [wpc][32] double           score_pairs scratch
//   xab     [wpc][2][MP][2] int16      end-point coordinates of the src / dst peaks
//   used    [wpc][2][MP] uint8         greedy: peak already connected
//   best    [wpc] Best                 block-wide arg-max (big items)
//   big     [big_words] uint32         bitmap over the CTA's items (local index n * wpc + warp): left to the whole CTA
struct LimbsLayout {
  unsigned planes, pool, dot, xab, used, best, big, total;
  int pool_doubles, big_words, cells4;
};
inline LimbsLayout limbs_layout(int cells, int MP, int wpc, int items_per_cta) {
  LimbsLayout y{};
  y.cells4 = (cells + 3) & ~3;
  y.pool_doubles = 512;
  while ((long long)y.pool_doubles * wpc < (long long)MP * MP) y.pool_doubles *= 2;
  size_t off = 0;
  y.planes = (unsigned)off; off += (size_t)wpc * 2 * y.cells4 * sizeof(float);
  off = (off + 15) & ~(size_t)15;
  y.pool = (unsigned)off; off += (size_t)wpc * y.pool_doubles * sizeof(double);
  y.dot = (unsigned)off; off += (size_t)wpc * 32 * sizeof(double);
  y.best = (unsigned)off; off += (size_t)wpc * sizeof(Best);
  y.xab = (unsigned)off; off += (size_t)wpc * 2 * MP * 2 * sizeof(int16_t);
  y.used = (unsigned)off; off += (size_t)wpc * 2 * MP;
  off = (off + 3) & ~(size_t)3;
  y.big_words = (items_per_cta + 31) / 32;
  y.big = (unsigned)off; off += (size_t)y.big_words * sizeof(unsigned int);
  y.total = (unsigned)((off + 15) & ~(size_t)15);
  return y;
}

// stage one item's PAF planes and end-point coordinates with `nthr` threads (thread `t` of them)
__device__ __forceinline__ void limbs_stage(const float* __restrict__ paf, const PopnetDecodeOut& o, int b, int l, int ta, int tb,
                                            int na, int nb, int K, int L, int MP, int cells, int cells4, float* s_px,
                                            int16_t (*xa)[2], int16_t (*xb)[2], unsigned char* used_a, unsigned char* used_b,
                                            int t, int nthr) {
  const float* mx = paf + ((size_t)b * 2 * L + 2 * l) * cells;             // planes 2l (x) and 2l + 1 (y) are adjacent
  if ((cells & 3) == 0 && (reinterpret_cast<uintptr_t>(mx) & 15u) == 0) {
    const float4* m4 = reinterpret_cast<const float4*>(mx);
    float4* d4 = reinterpret_cast<float4*>(s_px);
    for (int i = t; i < cells / 2; i += nthr) d4[i] = __ldg(m4 + i);       // 2 * cells / 4 vectors; cells4 == cells here
  } else {
    for (int i = t; i < cells; i += nthr) { s_px[i] = mx[i]; s_px[cells4 + i] = mx[cells + i]; }
  }
  for (int i = t; i < na; i += nthr) {
    *reinterpret_cast<int*>(xa[i]) = *reinterpret_cast<const int*>(o.peak_xy + (((size_t)b * K + ta) * MP + i) * 2);
    used_a[i] = 0;
  }
  for (int i = t; i < nb; i += nthr) {
    *reinterpret_cast<int*>(xb[i]) = *reinterpret_cast<const int*>(o.peak_xy + (((size_t)b * K + tb) * MP + i) * 2);
    used_b[i] = 0;
  }
}

__global__ void __launch_bounds__(512) limbs_kernel(const float* __restrict__ paf, int batch, PopnetDecodeParams p,
                                                    PopnetDecodeOut o, LimbsLayout lay) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const int H = p.grid_h, W = p.grid_w, cells = H * W, cells4 = lay.cells4;
  const int K = p.num_joints, L = p.num_limbs, MP = p.max_peaks, NP = p.num_intermed_pts;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wpc = blockDim.x >> 5;
  float* s_px = reinterpret_cast<float*>(s_dyn + lay.planes) + (size_t)warp * 2 * cells4;
  double* s_pool = reinterpret_cast<double*>(s_dyn + lay.pool) + (size_t)warp * lay.pool_doubles;
  double* s_dot = reinterpret_cast<double*>(s_dyn + lay.dot) + warp * 32;
  int16_t (*xa)[2] = reinterpret_cast<int16_t (*)[2]>(s_dyn + lay.xab) + (size_t)warp * 2 * MP;
  int16_t (*xb)[2] = xa + MP;
  unsigned char* used_a = s_dyn + lay.used + (size_t)warp * 2 * MP;
  unsigned char* used_b = used_a + MP;
  Best* s_best = reinterpret_cast<Best*>(s_dyn + lay.best);
  unsigned int* s_big = reinterpret_cast<unsigned int*>(s_dyn + lay.big);
  for (int i = tid; i < lay.big_words; i += blockDim.x) s_big[i] = 0u;
  __syncthreads();
  const double Hup = (double)(H * p.stride);
  const int items = batch * L;

  // ---- pass 1: one warp per item
  int li = warp;                                         // local item index n * wpc + warp
  for (int it = blockIdx.x * wpc + warp; it < items; it += gridDim.x * wpc, li += wpc) {
    const int b = it / L, l = it - b * L;
    const int ta = p.limbs[l][0], tb = p.limbs[l][1];
    const int na = o.peak_count[(size_t)b * K + ta], nb = o.peak_count[(size_t)b * K + tb];
    if (na == 0 || nb == 0) {
      if (lane == 0) o.conn_count[(size_t)b * L + l] = 0;
      continue;
    }
    const int npairs = na * nb;
    if (npairs > lay.pool_doubles) {                   // too many candidates for one warp's pool: the whole CTA takes it in pass 2
      if (lane == 0) atomicOr(&s_big[li >> 5], 1u << (li & 31));
      continue;
    }
    limbs_stage(paf, o, b, l, ta, tb, na, nb, K, L, MP, cells, cells4, s_px, xa, xb, used_a, used_b, lane, 32);
    __syncwarp();
    score_pairs(s_px, s_px + cells4, W, H, xa, xb, na, nb, NP, p.thresh_paf, Hup, s_pool, nb, s_dot, 0, 1, lane);
    // Stable descending sort + greedy (paf_to_pose.py:241-261) == repeatedly take the best remaining pair whose ends are
    // both free, ties to the smallest (i, j)
    const int maxc = min(na, nb);
    int nconn = 0;
    while (nconn < maxc) {
      double bv = -CUDART_INF;
      int bi = 0x7fffffff;
      for (int pr = lane; pr < npairs; pr += 32) {
        const int i = pr / nb, j = pr - i * nb;
        if (used_a[i] || used_b[j]) continue;
        const double v = s_pool[pr];
        if (v > bv) { bv = v; bi = pr; }
      }
#pragma unroll
      for (int ofs = 16; ofs > 0; ofs >>= 1) {
        const double ov = __shfl_xor_sync(kFull, bv, ofs);
        const int oi = __shfl_xor_sync(kFull, bi, ofs);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (bi == 0x7fffffff) break;
      const int i = bi / nb, j = bi - i * nb;
      if (lane == 0) {
        used_a[i] = 1; used_b[j] = 1;
        const size_t slot = ((size_t)b * L + l) * MP + nconn;
        *reinterpret_cast<int*>(o.conn_ij + slot * 2) = (i & 0xffff) | (j << 16);
        o.conn_score[slot] = bv;
      }
      ++nconn;
      __syncwarp();
    }
    if (lane == 0) o.conn_count[(size_t)b * L + l] = nconn;
    __syncwarp();
  }
  __syncthreads();

  // ---- pass 2: items with more candidates than a warp's pool, one at a time by the whole CTA; the score matrix spans all
  // pools (wpc * pool_doubles >= max_peaks^2), planes and tables are warp 0's
  float* b_px = reinterpret_cast<float*>(s_dyn + lay.planes);
  double* b_score = reinterpret_cast<double*>(s_dyn + lay.pool);
  int16_t (*bxa)[2] = reinterpret_cast<int16_t (*)[2]>(s_dyn + lay.xab);
  int16_t (*bxb)[2] = bxa + MP;
  unsigned char* bused_a = s_dyn + lay.used;
  unsigned char* bused_b = bused_a + MP;
  for (int w = 0; w < lay.big_words; ++w)
  for (unsigned int word = s_big[w]; word != 0u; word &= word - 1u) {
    const int bli = w * 32 + __ffs((int)word) - 1;
    const int it = blockIdx.x * wpc + bli % wpc + (bli / wpc) * (int)gridDim.x * wpc;
    const int b = it / L, l = it - b * L;
    const int ta = p.limbs[l][0], tb = p.limbs[l][1];
    const int na = o.peak_count[(size_t)b * K + ta], nb = o.peak_count[(size_t)b * K + tb];
    const int npairs = na * nb;
    limbs_stage(paf, o, b, l, ta, tb, na, nb, K, L, MP, cells, cells4, b_px, bxa, bxb, bused_a, bused_b, tid, (int)blockDim.x);
    __syncthreads();
    score_pairs(b_px, b_px + cells4, W, H, bxa, bxb, na, nb, NP, p.thresh_paf, Hup, b_score, nb,
                reinterpret_cast<double*>(s_dyn + lay.dot) + warp * 32, warp, wpc, lane);
    __syncthreads();
    const int maxc = min(na, nb);
    int nconn = 0;
    while (nconn < maxc) {
      double bv = -CUDART_INF;
      int bi = 0x7fffffff;
      for (int pr = tid; pr < npairs; pr += blockDim.x) {
        const int i = pr / nb, j = pr - i * nb;
        if (bused_a[i] || bused_b[j]) continue;
        const double v = b_score[pr];
        if (v > bv) { bv = v; bi = pr; }
      }
#pragma unroll
      for (int ofs = 16; ofs > 0; ofs >>= 1) {
        const double ov = __shfl_xor_sync(kFull, bv, ofs);
        const int oi = __shfl_xor_sync(kFull, bi, ofs);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { s_best[warp].v = bv; s_best[warp].idx = bi; }
      __syncthreads();
      bv = s_best[0].v; bi = s_best[0].idx;
      for (int w = 1; w < wpc; ++w) {
        const double ov = s_best[w].v;
        const int oi = s_best[w].idx;
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (bi == 0x7fffffff) break;                       // uniform: nothing admissible is left
      const int i = bi / nb, j = bi - i * nb;
      if (tid == 0) {
        bused_a[i] = 1; bused_b[j] = 1;
        const size_t slot = ((size_t)b * L + l) * MP + nconn;
        *reinterpret_cast<int*>(o.conn_ij + slot * 2) = (i & 0xffff) | (j << 16);
        o.conn_score[slot] = bv;
      }
      ++nconn;
      __syncthreads();
    }
    if (tid == 0) o.conn_count[(size_t)b * L + l] = nconn;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// D5 + D7..D9: assembly, pruning, lift (one warp per frame)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sum_pairwise_f32(const float* a, int n) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r += a[i];
    return r;
  }
  float r = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  for (int i = 8; i < n; ++i) r += a[i];
  return r;
}


// Multi-GPU record push (PopnetPeerPush, include/popnet_b200.h): every record value is stored locally AND at the same
// offset of this rank's chunk in every peer's gather buffer -- plain stores to peer-mapped memory (NVLink).
struct PushCtx {
  int world, rank;
  char* peer_chunk[POPNET_MAX_PEERS];      // gather_base[p] + rank * records_bytes (entry `rank` unused)
  const char* local_chunk;                 // gather_base[rank] + rank * records_bytes: where the local record fields live
};
template <typename T>
__device__ __forceinline__ void rec_store(const PushCtx& pc, T* ptr, T v) {
  *ptr = v;
  if (pc.world > 1) {
    const size_t off = (size_t)(reinterpret_cast<const char*>(ptr) - pc.local_chunk);
#pragma unroll 1
    for (int q = 0; q < pc.world; ++q)
      if (q != pc.rank) *reinterpret_cast<T*>(pc.peer_chunk[q] + off) = v;
  }
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}


__device__ __forceinline__ PushCtx make_push_ctx(const PopnetPeerPush& push) {
  PushCtx pc;
  pc.world = push.world; pc.rank = push.rank;
  pc.local_chunk = nullptr;
  if (push.world > 1) {
    for (int q = 0; q < push.world; ++q)
      pc.peer_chunk[q] = static_cast<char*>(push.gather_base[q]) + (size_t)push.rank * push.records_bytes;
    pc.local_chunk = pc.peer_chunk[push.rank];
  }
  return pc;
}

// publish (multi-GPU): when the LAST CTA of the grid has pushed its frames, tag this rank's slot in every rank's arrive[]
// array.  Called by all threads of the CTA after their last record store.
__device__ __forceinline__ void publish_push(const PopnetPeerPush& push) {
  if (push.world <= 1) return;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(push.done_counter, 1u);
    if (ticket == gridDim.x - 1) {
      *push.done_counter = 0u;
      const unsigned long long tag = *push.step + 1ull;
      *push.step = tag;
      __threadfence_system();
      for (int q = 0; q < push.world; ++q) st_release_sys(push.arrive[q] + push.rank, tag);
    }
  }
}

// Shared-memory views of ONE frame's tables (row pitches in elements): what the assembly and the lift read and write.
struct FrameTables {
  double* cs;        // [L][MP]       connection scores
  int16_t* ci;       // [L][MP][2]    connection (src peak, dst peak)
  float* pk;         // [K][MP]       peak scores
  int16_t* xy;       // [K][MP][2]    peak coordinates (upsampled pixels)
  int* nc;           // [L]
  int* npk;          // [K]
  int16_t* pj;       // [MM][POPNET_MAX_JOINTS]  person -> peak index per joint type, -1 = none
  double* ps;        // [MM]          person score
  int* pc;           // [MM]          person joint count
  int* keep;         // [MM]          surviving persons, in order
  int MP;           // row pitch of the [..][MP] tables (elements)
};

// D5 + pruning by ONE warp (paf_to_pose.py:267-351): the frame's connections are taken limb by limb in list order; the
// chain of dependent steps is what bounds a frame, so it is kept short:
//   * lane c pre-loads connection c of the limb (ends, score, end-point peak scores) and the steps get them by shuffle -- no
//     shared-memory load sits between two steps except the person rows themselves;
//   * lane q OWNS person q (and q + 32): it tests its own row, and in the common step (exactly one person matches) the owner
//     alone updates row, count and score -- no other lane reads them before the next warp-wide event, so no barrier;
//   * merges and new persons touch rows across lanes and are fenced with __syncwarp().
// Returns the number of surviving persons (T.keep[0..n)); *flags_out gets POPNET_FLAG_PERSON_OVERFLOW if the table filled.
__device__ __forceinline__ int assemble_persons(const FrameTables& T, const PopnetDecodeParams& p, int lane, unsigned* flags_out) {
  const int K = p.num_joints, L = p.num_limbs, MM = p.max_persons, MP = T.MP;
  constexpr int PJ = POPNET_MAX_JOINTS;
  int np_ = 0;
  unsigned flags = 0;
  for (int l = 0; l < L; ++l) {
    const int ta = p.limbs[l][0], tb = p.limbs[l][1];
    const int nc = T.nc[l];
    for (int c0 = 0; c0 < nc; c0 += 32) {
      int my_ia = 0, my_ib = 0;
      double my_ls = 0.0, my_sa = 0.0, my_sb = 0.0;
      if (c0 + lane < nc) {
        const int c = l * MP + c0 + lane;
        my_ia = T.ci[c * 2]; my_ib = T.ci[c * 2 + 1];
        my_ls = T.cs[c];
        my_sa = (double)T.pk[ta * MP + my_ia];
        my_sb = (double)T.pk[tb * MP + my_ib];
      }
      const int n = min(32, nc - c0);
      for (int c = 0; c < n; ++c) {
        const int ia = __shfl_sync(kFull, my_ia, c), ib = __shfl_sync(kFull, my_ib, c);
        const double ls = __shfl_sync(kFull, my_ls, c), sa = __shfl_sync(kFull, my_sa, c), sb = __shfl_sync(kFull, my_sb, c);
        // persons whose src or dst slot already holds this joint (paf_to_pose.py:285-287)
        unsigned long long hits = 0;
        for (int q0 = 0; q0 < np_; q0 += 32) {
          const int q = q0 + lane;
          const bool h = q < np_ && (T.pj[q * PJ + ta] == ia || T.pj[q * PJ + tb] == ib);
          hits |= (unsigned long long)__ballot_sync(kFull, h) << q0;
        }
        const int nh = __popcll(hits);
        if (nh == 1) {
          const int q = __ffsll((long long)hits) - 1;
          if (lane == (q & 31) && T.pj[q * PJ + tb] != ib) {
            T.pj[q * PJ + tb] = (int16_t)ib;
            T.pc[q] += 1;
            T.ps[q] += sb + ls;
          }
        } else if (nh == 2) {
          __syncwarp();                                   // the owners' updates are visible to the lanes that merge
          const int q1 = __ffsll((long long)hits) - 1;
          const int q2 = __ffsll((long long)(hits & (hits - 1))) - 1;
          const bool ov = lane < K && T.pj[q1 * PJ + lane] >= 0 && T.pj[q2 * PJ + lane] >= 0;
          if (!__any_sync(kFull, ov)) {
            if (lane < K) T.pj[q1 * PJ + lane] = (int16_t)(T.pj[q1 * PJ + lane] + T.pj[q2 * PJ + lane] + 1);
            if (lane == 0) {
              T.ps[q1] += T.ps[q2];
              T.pc[q1] += T.pc[q2];
              T.ps[q1] += ls;
            }
            __syncwarp();
            for (int q = q2; q + 1 < np_; ++q) {        // list.pop(q2): later persons move up
              if (lane < K) T.pj[q * PJ + lane] = T.pj[(q + 1) * PJ + lane];
              if (lane == 0) { T.ps[q] = T.ps[q + 1]; T.pc[q] = T.pc[q + 1]; }
              __syncwarp();
            }
            --np_;
          } else if (lane == 0) {
            T.pj[q1 * PJ + tb] = (int16_t)ib;
            T.pc[q1] += 1;
            T.ps[q1] += sb + ls;
          }
          __syncwarp();
        } else {                                          // 0 or >= 3 matches: a new person
          if (np_ >= MM) flags |= POPNET_FLAG_PERSON_OVERFLOW;
          else {
            if (lane < PJ) T.pj[np_ * PJ + lane] = (lane == ta) ? (int16_t)ia : (lane == tb) ? (int16_t)ib : (int16_t)-1;
            if (lane == 0) {
              T.pc[np_] = 2;
              T.ps[np_] = ((0 + sa) + sb) + ls;
            }
            ++np_;
          }
          __syncwarp();
        }
      }
    }
  }
  __syncwarp();
  // prune (paf_to_pose.py:338-346): ordered compaction of the survivors
  int nout = 0;
  for (int q0 = 0; q0 < np_; q0 += 32) {
    const int q = q0 + lane;
    bool keep = false;
    if (q < np_) {
      const double cnt = (double)T.pc[q], sc = T.ps[q];
      keep = !(cnt < 3 || sc / cnt < 0.2);
    }
    const unsigned bal = __ballot_sync(kFull, keep);
    if (keep) T.keep[nout + __popc(bal & ((1u << lane) - 1u))] = q;
    nout += __popc(bal);
  }
  __syncwarp();
  *flags_out = flags;
  return nout;
}

// D7 .. D9 for the frame's `nout` surviving persons by threads t = tid, tid + nthreads, ...: every (person, joint) pair is
// independent.  Heat-weighted depth over the clipped 3 x 3 window (common.py:272-293, fp32, NumPy pairwise order), rescale to
// the original image, pinhole back-projection; all record values go through rec_store (local + peers).
__device__ __forceinline__ void lift_and_store(const FrameTables& T, const PopnetDecodeParams& p, const PopnetDecodeOut& o,
                                               const PushCtx& pc, const float* __restrict__ heat, const float* __restrict__ depth,
                                               int b, int nout, int tid, int nthreads) {
  const int K = p.num_joints, MP = T.MP, MM = p.max_persons;
  const int H = p.grid_h, W = p.grid_w, cells = H * W;
  constexpr int PJ = POPNET_MAX_JOINTS;
  for (int i = tid; i < nout; i += nthreads) {
    const size_t row = (size_t)b * MM + i;
    const int q = T.keep[i];
    if (o.person_score) rec_store(pc, o.person_score + row, T.ps[q]);
    if (o.person_njoint) rec_store(pc, o.person_njoint + row, (int32_t)T.pc[q]);
  }
  for (int i = tid; i < nout * K; i += nthreads) {
    const int pi_ = i / K, k = i - pi_ * K;
    const int q = T.keep[pi_], idx = T.pj[q * PJ + k];
    const size_t row = (size_t)b * MM + pi_;
    if (o.person_peak) rec_store(pc, o.person_peak + row * K + k, (int16_t)idx);
    double x2 = -1, y2 = -1, Z = -1, conf = 0;
    if (idx >= 0) {
      const int X = T.xy[(k * MP + idx) * 2], Y = T.xy[(k * MP + idx) * 2 + 1];
      conf = (double)T.pk[k * MP + idx];
      if (depth) {
        const int cx = X / p.stride, cy = Y / p.stride;
        const float* hm = heat + ((size_t)b * (K + 1) + k) * cells;
        const float* dm = depth + ((size_t)b * (p.depth_channels > 0 ? p.depth_channels : K) + k) * cells;
        if (cx >= 1 && cx <= W - 2 && cy >= 1 && cy <= H - 2) {
          // interior window (the usual case): all 18 loads are issued before the first use
          float hv[9], dd[9];
#pragma unroll
          for (int u = 0; u < 9; ++u) {
            const int at = (cy - 1 + u / 3) * W + cx - 1 + u % 3;
            hv[u] = __ldg(hm + at); dd[u] = __ldg(dm + at);
          }
          float wv[9], dv[9];
#pragma unroll
          for (int u = 0; u < 9; ++u) {
            float h = hv[u];
            if (h < 0) h = 0;
            const float w = h + 0.000000001f;
            float d = dd[u] * p.depth_std;
            d = d + p.depth_mean;
            wv[u] = w; dv[u] = d * w;
          }
          const float sd = (((dv[0] + dv[1]) + (dv[2] + dv[3])) + ((dv[4] + dv[5]) + (dv[6] + dv[7]))) + dv[8];
          const float sw = (((wv[0] + wv[1]) + (wv[2] + wv[3])) + ((wv[4] + wv[5]) + (wv[6] + wv[7]))) + wv[8];
          Z = (double)(sd / sw);
        } else {
          const int x0 = clampi(cx - 1, 0, W - 1), x1 = clampi(cx + 1, 0, W - 1);
          const int y0 = clampi(cy - 1, 0, H - 1), y1 = clampi(cy + 1, 0, H - 1);
          float wv[9], dv[9];
          int n = 0;
          for (int yy = y0; yy <= y1; ++yy)
            for (int xx = x0; xx <= x1; ++xx) {
              float hv = hm[yy * W + xx];
              if (hv < 0) hv = 0;
              const float w = hv + 0.000000001f;
              float d = dm[yy * W + xx] * p.depth_std;
              d = d + p.depth_mean;
              wv[n] = w; dv[n] = d * w; ++n;
            }
          Z = (double)(sum_pairwise_f32(dv, n) / sum_pairwise_f32(wv, n));
        }
      }
      x2 = (double)X / p.input_size * p.w_org;
      y2 = (double)Y / p.input_size * p.h_org;
    }
    if (o.pose2d) { rec_store(pc, o.pose2d + (row * K + k) * 2, x2); rec_store(pc, o.pose2d + (row * K + k) * 2 + 1, y2); }
    if (o.pose_conf) rec_store(pc, o.pose_conf + row * K + k, conf);
    if (o.pose3d && depth) {
      double X3 = (x2 - p.cx) * Z / p.fx, Y3 = (y2 - p.cy) * Z / p.fy;
      if (p.flip_y) Y3 = -Y3;
      double* d3 = o.pose3d + (row * K + k) * 3;
      rec_store(pc, d3, X3); rec_store(pc, d3 + 1, Y3); rec_store(pc, d3 + 2, Z);
    }
  }
}

// bytes of one frame's tables in the assembly kernel's dynamic shared memory (8-byte fields first)
__host__ __device__ inline size_t asm_frame_bytes(int K, int L, int MP, int MM) {
  size_t n = 0;
  n += (size_t)L * MP * sizeof(double);                  // cs
  n += (size_t)MM * sizeof(double);                      // ps
  n += (size_t)K * MP * sizeof(float);                   // pk
  n += (size_t)(MM + MM + L + K) * sizeof(int);          // pc, keep, nc, npk
  n += (size_t)L * MP * 2 * sizeof(int16_t);             // ci
  n += (size_t)K * MP * 2 * sizeof(int16_t);             // xy
  n += (size_t)MM * POPNET_MAX_JOINTS * sizeof(int16_t); // pj
  return (n + 15) & ~(size_t)15;
}

// One WARP per frame, `blockDim.x / 32` frames per CTA: the assembly of a frame is a serial chain that one warp runs at
// its latency; what counts is how many SMs the kernel holds while it does so (its CTAs share the GPU with the next
// batch's convolutions, which cannot co-reside with anything: they fill the register file).  Eight frames per CTA: a batch
// of 64 is 8 CTAs instead of 64.
__global__ void __launch_bounds__(256) assemble_kernel(const float* __restrict__ heat, const float* __restrict__ depth, int batch,
                                                       PopnetDecodeParams p, PopnetDecodeOut o, PopnetPeerPush push,
                                                       unsigned int frame_bytes) {
  extern __shared__ __align__(16) unsigned char s_asm[];
  const PushCtx pc = make_push_ctx(push);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = p.num_joints, L = p.num_limbs, MP = p.max_peaks, MM = p.max_persons;
  const int wpc = blockDim.x >> 5;
  for (int b = blockIdx.x * wpc + warp; b < batch; b += gridDim.x * wpc) {
    FrameTables T;
    unsigned char* base = s_asm + (size_t)warp * frame_bytes;
    T.cs = reinterpret_cast<double*>(base); base += (size_t)L * MP * sizeof(double);
    T.ps = reinterpret_cast<double*>(base); base += (size_t)MM * sizeof(double);
    T.pk = reinterpret_cast<float*>(base); base += (size_t)K * MP * sizeof(float);
    T.pc = reinterpret_cast<int*>(base); base += (size_t)MM * sizeof(int);
    T.keep = reinterpret_cast<int*>(base); base += (size_t)MM * sizeof(int);
    T.nc = reinterpret_cast<int*>(base); base += (size_t)L * sizeof(int);
    T.npk = reinterpret_cast<int*>(base); base += (size_t)K * sizeof(int);
    T.ci = reinterpret_cast<int16_t*>(base); base += (size_t)L * MP * 2 * sizeof(int16_t);
    T.xy = reinterpret_cast<int16_t*>(base); base += (size_t)K * MP * 2 * sizeof(int16_t);
    T.pj = reinterpret_cast<int16_t*>(base);
    T.MP = MP;
    // ---- stage the frame's connection lists and peaks (written by the two kernels before this one).  Items are (list,
    // entry) pairs up to the longest list; four per lane are LOADED before the first is stored, so a round costs one
    // global-memory latency, not four.
    int mc = 0, mk = 0;
    for (int l = lane; l < L; l += 32) { const int v = o.conn_count[(size_t)b * L + l]; T.nc[l] = v; mc = max(mc, v); }
    for (int k = lane; k < K; k += 32) { const int v = o.peak_count[(size_t)b * K + k]; T.npk[k] = v; mk = max(mk, v); }
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
      mc = max(mc, __shfl_xor_sync(kFull, mc, ofs));
      mk = max(mk, __shfl_xor_sync(kFull, mk, ofs));
    }
    __syncwarp();
    for (int i0 = 0; i0 < L * mc; i0 += 128) {
      int cij[4]; double csv[4]; int at[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 32 + lane;
        at[u] = -1;
        if (i < L * mc) {
          const int l = i / mc, c = i - l * mc;
          if (c < T.nc[l]) {
            const size_t slot = ((size_t)b * L + l) * MP + c;
            at[u] = l * MP + c;
            cij[u] = *reinterpret_cast<const int*>(o.conn_ij + slot * 2);
            csv[u] = o.conn_score[slot];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (at[u] >= 0) { *reinterpret_cast<int*>(T.ci + at[u] * 2) = cij[u]; T.cs[at[u]] = csv[u]; }
    }
    for (int i0 = 0; i0 < K * mk; i0 += 128) {
      int pxy[4]; float pkv[4]; int at[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 32 + lane;
        at[u] = -1;
        if (i < K * mk) {
          const int k = i / mk, c = i - k * mk;
          if (c < T.npk[k]) {
            const size_t slot = ((size_t)b * K + k) * MP + c;
            at[u] = k * MP + c;
            pxy[u] = *reinterpret_cast<const int*>(o.peak_xy + slot * 2);
            pkv[u] = o.peak_score[slot];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (at[u] >= 0) { *reinterpret_cast<int*>(T.xy + at[u] * 2) = pxy[u]; T.pk[at[u]] = pkv[u]; }
    }
    __syncwarp();
    unsigned flags = 0;
    const int nout = assemble_persons(T, p, lane, &flags);
    if (lane == 0) {
      rec_store(pc, o.n_person + b, nout);
      // the frame's flag word is complete here (peaks / limbs kernels have finished): fold in this kernel's bits
      const uint32_t fl = o.flags[b] | flags;
      rec_store(pc, o.flags + b, fl);
    }
    lift_and_store(T, p, o, pc, heat, depth, b, nout, lane, 32);
    __syncwarp();                  // the tables are rewritten by the warp's next frame
  }
  publish_push(push);
}

// one warp: lane q waits for rank q's tag of the current step (bounded: a dead peer sets *status instead of hanging)
__global__ void __launch_bounds__(32) p2p_wait_kernel(const unsigned long long* arrive, int world, const unsigned long long* step,
                                                      unsigned int* status, unsigned long long timeout_ns) {
  const int q = threadIdx.x;
  if (q >= world) return;
  const unsigned long long want = *step;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (ld_acquire_sys(arrive + q) < want) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) { atomicExch(status, 1u); return; }
    __nanosleep(200);
  }
}

// NumPy's pairwise summation of n <= 128 contiguous fp32 values (numpy/core/src/umath/loops_utils.h, pairwise_sum):
// below 8 a plain loop from 0; else eight running sums over whole blocks of 8, combined as a tree, then the tail in order
__device__ float sum_pairwise_f32_n(const float* a, int n) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r += a[i];
    return r;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] += a[i + j];
  float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += a[i];
  return res;
}

constexpr int kMaxLiftRadius = 5;                                   // (2*5+1)^2 = 121 <= 128: one pairwise block
constexpr int kMaxLiftWin = (2 * kMaxLiftRadius + 1) * (2 * kMaxLiftRadius + 1);

// retrieve_depth_heat_weighted / retrieve_depth_weighted / retrieve_depth_heat_max (lib/utils/common.py:251-318) for
// arbitrary query points and window radius: one thread per query
__global__ void __launch_bounds__(128) lift_points_kernel(const float* __restrict__ heat, const float* __restrict__ depth,
                                                          const int32_t* __restrict__ queries, int n, int H, int W,
                                                          float depth_mean, float depth_std, int mode, int radius,
                                                          float* __restrict__ out) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  const int plane = queries[3 * i], cx = queries[3 * i + 1], cy = queries[3 * i + 2];
  // min(max(c - r, 0), g - 1) .. max(min(c + r, g - 1), 0)   (common.py:279-282)
  const int x0 = clampi(cx - radius, 0, W - 1), x1 = clampi(cx + radius, 0, W - 1);
  const int y0 = clampi(cy - radius, 0, H - 1), y1 = clampi(cy + radius, 0, H - 1);
  const float* hm = heat ? heat + (size_t)plane * H * W : nullptr;
  const float* dm = depth + (size_t)plane * H * W;
  float wv[kMaxLiftWin], dv[kMaxLiftWin];
  int m = 0;
  float best_w = 0.f, best_d = 0.f;
  for (int yy = y0; yy <= y1; ++yy)
    for (int xx = x0; xx <= x1; ++xx) {
      float hv = hm ? hm[yy * W + xx] : 0.f;
      if (hv < 0) hv = 0;
      float d = dm[yy * W + xx] * depth_std;
      d = d + depth_mean;
      if (mode == POPNET_LIFT_HEAT_WEIGHTED) {
        const float w = hv + 0.000000001f;
        wv[m] = w; dv[m] = d * w;
      } else {
        dv[m] = d;
        if (m == 0 || hv > best_w) { best_w = hv; best_d = d; }     // np.argmax: first maximum, row-major (NaN-free maps)
      }
      ++m;
    }
  if (mode == POPNET_LIFT_HEAT_WEIGHTED) out[i] = sum_pairwise_f32_n(dv, m) / sum_pairwise_f32_n(wv, m);
  else if (mode == POPNET_LIFT_MEAN) out[i] = sum_pairwise_f32_n(dv, m) / (float)m;     // np.mean of an fp32 window
  else out[i] = best_d;
}

}  // namespace

namespace {
int decode_impl(const float* heat, const float* paf, const float* depth, int batch, const PopnetDecodeParams* p,
                const PopnetDecodeOut* o, const PopnetPeerPush* push, void* stream) {
  if (!heat || !paf || !p || !o || batch < 0) return POPNET_ERR_INVALID_ARG;
  PopnetPeerPush pp{};
  if (push) {
    pp = *push;
    if (pp.world < 1 || pp.world > POPNET_MAX_PEERS || pp.rank < 0 || pp.rank >= pp.world || !pp.step || !pp.done_counter)
      return POPNET_ERR_INVALID_ARG;
    // every record field must live inside this rank's chunk of its own gather buffer
    const char* lo = static_cast<const char*>(pp.gather_base[pp.rank]) + (size_t)pp.rank * pp.records_bytes;
    const char* hi = lo + pp.records_bytes;
    const void* fields[] = {o->n_person, o->flags, o->person_peak, o->person_score, o->person_njoint, o->pose2d, o->pose3d, o->pose_conf};
    for (const void* f : fields)
      if (f && (static_cast<const char*>(f) < lo || static_cast<const char*>(f) >= hi)) return POPNET_ERR_INVALID_ARG;
    for (int q = 0; q < pp.world; ++q)
      if (!pp.gather_base[q] || !pp.arrive[q]) return POPNET_ERR_INVALID_ARG;
  }
  // peak_* and conn_* double as the stage-to-stage storage and are therefore mandatory
  if (!o->peak_count || !o->peak_xy || !o->peak_score || !o->conn_count || !o->conn_ij || !o->conn_score ||
      !o->n_person || !o->flags)
    return POPNET_ERR_INVALID_ARG;
  // (x, y) / (src, dst) int16 pairs are moved as one 32-bit word
  if ((reinterpret_cast<uintptr_t>(o->peak_xy) & 3u) || (reinterpret_cast<uintptr_t>(o->conn_ij) & 3u)) return POPNET_ERR_INVALID_ARG;
  if (p->num_joints < 1 || p->num_joints > POPNET_MAX_JOINTS || p->num_limbs < 1 || p->num_limbs > POPNET_MAX_LIMBS ||
      p->stride != 8 || p->max_peaks < 1 || p->max_peaks > POPNET_MAX_PEAKS || p->max_persons < 1 ||
      p->max_persons > POPNET_MAX_PERSONS || p->grid_h < 1 || p->grid_w < 1 || p->grid_h * p->grid_w > kMaxCells ||
      p->num_intermed_pts < 1 || p->num_intermed_pts > 32 || p->num_joints > 32)
    return POPNET_ERR_UNSUPPORTED;
  if (p->depth_channels != 0 && p->depth_channels < p->num_joints) return POPNET_ERR_INVALID_ARG;
  for (int l = 0; l < p->num_limbs; ++l)
    if (p->limbs[l][0] < 0 || p->limbs[l][0] >= p->num_joints || p->limbs[l][1] < 0 || p->limbs[l][1] >= p->num_joints)
      return POPNET_ERR_INVALID_ARG;
  if (p->max_ctas < 0) return POPNET_ERR_INVALID_ARG;
  if (batch == 0) return POPNET_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cells = p->grid_h * p->grid_w;
  int dev = 0, sms = 148;
  POPNET_CUDA_TRY(cudaGetDevice(&dev));
  POPNET_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int cap = p->max_ctas > 0 ? p->max_ctas : sms;
  constexpr size_t kSmemBudget = 220 * 1024;
  POPNET_CUDA_TRY(cudaMemsetAsync(o->flags, 0, sizeof(uint32_t) * batch, st));
  {
    // warps (= items in flight) per CTA: as many map slices as fit, at most 32
    const size_t wb = peaks_warp_bytes(cells, p->max_peaks);
    int wpc = 32;
    while (wpc > 1 && wb * wpc > kSmemBudget) wpc >>= 1;
    const int items = batch * p->num_joints;
    const int grid = std::min(cap, (items + wpc - 1) / wpc);
    if (wb * wpc > 48 * 1024)
      POPNET_CUDA_TRY(cudaFuncSetAttribute(peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(wb * wpc)));
    peaks_kernel<<<grid, 32 * wpc, wb * wpc, st>>>(heat, batch, *p, *o);
    POPNET_AFTER_LAUNCH();
  }
  {
    const int items = batch * p->num_limbs;
    int wpc = 16, grid = 1;
    LimbsLayout lay{};
    for (;; wpc >>= 1) {
      grid = std::min(cap, (items + wpc - 1) / wpc);
      const int per_cta = (items + grid * wpc - 1) / (grid * wpc) * wpc;
      lay = limbs_layout(cells, p->max_peaks, wpc, per_cta);
      if (lay.total <= kSmemBudget || wpc == 1) break;
    }
    if (lay.total > 227 * 1024) return POPNET_ERR_UNSUPPORTED;
    if (lay.total > 48 * 1024)
      POPNET_CUDA_TRY(cudaFuncSetAttribute(limbs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
    limbs_kernel<<<grid, 32 * wpc, lay.total, st>>>(paf, batch, *p, *o, lay);
    POPNET_AFTER_LAUNCH();
  }
  {
    // frames per CTA (one warp each): as many as fit into 200 KB of tables, at most 8
    const size_t fb = asm_frame_bytes(p->num_joints, p->num_limbs, p->max_peaks, p->max_persons);
    int fpc = 8;
    while (fpc > 1 && fb * fpc > 200 * 1024) fpc >>= 1;
    if (fb * fpc > 48 * 1024)
      POPNET_CUDA_TRY(cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fb * fpc)));
    assemble_kernel<<<std::min(cap, (batch + fpc - 1) / fpc), 32 * fpc, fb * fpc, st>>>(heat, depth, batch, *p, *o, pp, (unsigned int)fb);
  }
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}
}  // namespace

extern "C" int popnet_decode(const float* heat, const float* paf, const float* depth, int batch,
                             const PopnetDecodeParams* p, const PopnetDecodeOut* o, void* stream) {
  return decode_impl(heat, paf, depth, batch, p, o, nullptr, stream);
}

extern "C" int popnet_decode_push(const float* heat, const float* paf, const float* depth, int batch,
                                  const PopnetDecodeParams* p, const PopnetDecodeOut* o, const PopnetPeerPush* push,
                                  void* stream) {
  if (!push) return POPNET_ERR_INVALID_ARG;
  return decode_impl(heat, paf, depth, batch, p, o, push, stream);
}

extern "C" int popnet_p2p_wait(const unsigned long long* local_arrive, int world, const unsigned long long* step,
                               unsigned int* status, int timeout_ms, void* stream) {
  if (!local_arrive || !step || !status || world < 1 || world > POPNET_MAX_PEERS || timeout_ms < 1) return POPNET_ERR_INVALID_ARG;
  p2p_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(local_arrive, world, step, status,
                                                                   (unsigned long long)timeout_ms * 1000000ull);
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}

extern "C" int popnet_p2p_alloc(size_t bytes, void** dev_ptr_out, unsigned char* handle_out) {
  if (!dev_ptr_out || !handle_out || bytes == 0) return POPNET_ERR_INVALID_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* ptr = nullptr;
  POPNET_CUDA_TRY(cudaMalloc(&ptr, bytes));
  cudaError_t e = cudaMemset(ptr, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) { cudaFree(ptr); return popnet::record_cuda_error(e); }
  memcpy(handle_out, &h, sizeof(h));
  *dev_ptr_out = ptr;
  return POPNET_OK;
}

extern "C" int popnet_p2p_open(const unsigned char* handle, void** peer_ptr_out) {
  if (!handle || !peer_ptr_out) return POPNET_ERR_INVALID_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  POPNET_CUDA_TRY(cudaIpcOpenMemHandle(peer_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return POPNET_OK;
}

extern "C" int popnet_p2p_close(void* peer_ptr) {
  if (!peer_ptr) return POPNET_ERR_INVALID_ARG;
  POPNET_CUDA_TRY(cudaIpcCloseMemHandle(peer_ptr));
  return POPNET_OK;
}

extern "C" int popnet_p2p_free(void* dev_ptr) {
  if (!dev_ptr) return POPNET_ERR_INVALID_ARG;
  POPNET_CUDA_TRY(cudaFree(dev_ptr));
  return POPNET_OK;
}

extern "C" int popnet_lift_depth_window(const float* heat, const float* depth, const int32_t* queries, int n, int grid_h,
                                        int grid_w, float depth_mean, float depth_std, int mode, int radius, float* out_z,
                                        void* stream) {
  if (mode < POPNET_LIFT_HEAT_WEIGHTED || mode > POPNET_LIFT_HEAT_MAX) return POPNET_ERR_INVALID_ARG;
  if ((!heat && mode != POPNET_LIFT_MEAN) || !depth || !queries || !out_z || n < 0 || grid_h < 1 || grid_w < 1 || radius < 0)
    return POPNET_ERR_INVALID_ARG;
  if (radius > kMaxLiftRadius) return POPNET_ERR_UNSUPPORTED;
  if (n == 0) return POPNET_OK;
  lift_points_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(heat, depth, queries, n, grid_h, grid_w,
                                                                                    depth_mean, depth_std, mode, radius, out_z);
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}

extern "C" int popnet_lift_depth_mode(const float* heat, const float* depth, const int32_t* queries, int n, int grid_h,
                                      int grid_w, float depth_mean, float depth_std, int mode, float* out_z, void* stream) {
  return popnet_lift_depth_window(heat, depth, queries, n, grid_h, grid_w, depth_mean, depth_std, mode, 1, out_z, stream);
}

extern "C" int popnet_lift_depth(const float* heat, const float* depth, const int32_t* queries, int n, int grid_h, int grid_w,
                                 float depth_mean, float depth_std, float* out_z, void* stream) {
  return popnet_lift_depth_window(heat, depth, queries, n, grid_h, grid_w, depth_mean, depth_std, POPNET_LIFT_HEAT_WEIGHTED, 1,
                                  out_z, stream);
}

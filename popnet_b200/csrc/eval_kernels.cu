// PCK / mAP matching kernels (fp64, one warp per frame).
//
// Replaces the per-frame bodies of util/eval_pck.py:266-310, 377-475 (match_humans_2d/3d,
// compute_bbox_from_humans, bbox_ious) and util/eval_mAP.py:60-157 (assignGTmulti) of the reference.
// This translation unit is compiled with -fmad=false: every a*b+c below rounds twice like NumPy's
// element-wise arithmetic; the one place the reference goes through BLAS (np.linalg.norm -> ddot,
// eval_mAP.py:116) is written with explicit fma().
//
// Roofline: HBM-bound in principle (120*(6P+5G) bytes read per frame, section 8(d)); in practice
// latency-bound -- a frame is a few KB -- so the grid is sized to keep every SM busy with frames.
#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxCachedPred = 64;   // predicted boxes cached per warp; larger frames recompute
constexpr unsigned kFull = 0xffffffffu;

struct Box { double x0, y0, x1, y1; };

// bbox over joints != (-1,-1); valid = false if the human has no valid joint (eval_pck.py:433-449)
__device__ __forceinline__ Box bbox_of(const double* __restrict__ h2d, int K, bool& valid) {
  Box b{0, 0, 0, 0};
  int nv = 0;
  for (int k = 0; k < K; ++k) {
    const double x = h2d[2 * k], y = h2d[2 * k + 1];
    if (x == -1.0 && y == -1.0) continue;
    if (nv == 0) { b.x0 = b.x1 = x; b.y0 = b.y1 = y; }
    else {
      if (x < b.x0) b.x0 = x;
      if (x > b.x1) b.x1 = x;
      if (y < b.y0) b.y0 = y;
      if (y > b.y1) b.y1 = y;
    }
    ++nv;
  }
  valid = nv > 0;
  return b;
}

// eval_pck.py:452-475; 0/0 -> NaN is kept
__device__ __forceinline__ double iou_of(const Box& a, const Box& b) {
  double dx = fmin(a.x1, b.x1) - fmax(a.x0, b.x0);
  double dy = fmin(a.y1, b.y1) - fmax(a.y0, b.y0);
  if (!(dx > 0)) dx = 0;
  if (!(dy > 0)) dy = 0;
  const double inter = dx * dy;
  const double a1 = (a.x1 - a.x0) * (a.y1 - a.y0), a2 = (b.x1 - b.x0) * (b.y1 - b.y0);
  return inter / ((a1 + a2) - inter);
}

// np.max / np.argmax semantics over (value, index): the first NaN wins, otherwise the first maximum
__device__ __forceinline__ bool better(double v, int i, double bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (bi < 0) return i >= 0;
  if (i < 0) return false;
  if (vn || bn) return vn && (!bn || i < bi);
  return v > bv || (v == bv && i < bi);
}

__device__ __forceinline__ void warp_argmax(double& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(kFull, v, o);
    const int oi = __shfl_xor_sync(kFull, i, o);
    if (better(ov, oi, v, i)) { v = ov; i = oi; }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) pck_kernel(PopnetPckArgs a) {
  __shared__ Box s_box[kWarpsPerBlock][kMaxCachedPred];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = a.num_joints;
  Box* pbox = s_box[warp];
  long long hit_acc = 0, valid_acc = 0;   // lane k accumulates joint k
  const int warps_total = gridDim.x * kWarpsPerBlock;
  for (int f = blockIdx.x * kWarpsPerBlock + warp; f < a.num_frames; f += warps_total) {
    const int g0 = a.gt_off[f], G = a.gt_off[f + 1] - g0;
    const int p0 = a.pred_off[f], P = a.pred_off[f + 1] - p0;
    if (a.status && lane == 0) a.status[f] = 0;
    if (G == 0) continue;
    bool pred_ok = P > 0, gt_ok = true;
    if (P > 0) {
      bool ok = true;
      for (int q = lane; q < P; q += 32) {
        bool v;
        const Box b = bbox_of(a.pred2d + (size_t)(p0 + q) * K * 2, K, v);
        ok &= v;
        if (q < kMaxCachedPred) pbox[q] = b;
      }
      pred_ok = __all_sync(kFull, ok);
      ok = true;
      for (int g = lane; g < G; g += 32) {
        bool v;
        bbox_of(a.gt2d + (size_t)(g0 + g) * K * 2, K, v);
        ok &= v;
      }
      gt_ok = __all_sync(kFull, ok);
      if (!gt_ok && a.status && lane == 0) a.status[f] = 1;
      __syncwarp();
    }
    for (int g = 0; g < G; ++g) {
      int m = -1;
      if (pred_ok && gt_ok) {
        bool v;
        const Box bg = bbox_of(a.gt2d + (size_t)(g0 + g) * K * 2, K, v);
        double best = 0;
        int bi = -1;
        for (int q = lane; q < P; q += 32) {
          Box bp;
          if (q < kMaxCachedPred) bp = pbox[q];
          else { bool pv; bp = bbox_of(a.pred2d + (size_t)(p0 + q) * K * 2, K, pv); }
          const double u = iou_of(bg, bp);
          if (better(u, q, best, bi)) { best = u; bi = q; }
        }
        warp_argmax(best, bi);
        if (!(best < a.iou_th)) m = bi;      // NaN < th is False -> matched (eval_pck.py:296)
      }
      if (a.matched_pred && lane == 0) a.matched_pred[g0 + g] = m;
      if (lane < K) {
        const int k = lane;
        double d = -1.0;
        if (m >= 0) {
          const double* p2 = a.pred2d + ((size_t)(p0 + m) * K + k) * 2;
          const double* g2 = a.gt2d + ((size_t)(g0 + g) * K + k) * 2;
          if (a.pred3d) {
            const double* p3 = a.pred3d + ((size_t)(p0 + m) * K + k) * 3;
            const double* g3 = a.gt3d + ((size_t)(g0 + g) * K + k) * 3;
            const double e0 = g3[0] - p3[0], e1 = g3[1] - p3[1], e2 = g3[2] - p3[2];
            d = sqrt((e0 * e0 + e1 * e1) + e2 * e2);
            if (g2[0] == -1.0 && g2[1] == -1.0) d = -1.0;
          } else {
            const double e0 = g2[0] - p2[0], e1 = g2[1] - p2[1];
            d = sqrt(e0 * e0 + e1 * e1);
          }
          if (p2[0] == -1.0 && p2[1] == -1.0) d = -1.0;
        }
        if (a.gt_vis && a.gt_vis[(size_t)(g0 + g) * K + k] == 0) d = -1.0;
        a.dists[(size_t)(g0 + g) * K + k] = d;
        const double th = a.gt_thresh ? a.gt_thresh[g0 + g] : a.dist_th;
        const bool hit = (d >= 0) && (d < th);
        if (a.hit) a.hit[(size_t)(g0 + g) * K + k] = hit ? 1 : 0;
        hit_acc += hit;
        valid_acc += (d >= 0);
      }
    }
  }
  if (lane < K) {
    if (hit_acc) atomicAdd(reinterpret_cast<unsigned long long*>(a.hit_cnt) + lane, (unsigned long long)hit_acc);
    if (valid_acc) atomicAdd(reinterpret_cast<unsigned long long*>(a.valid_cnt) + lane, (unsigned long long)valid_acc);
  }
}

// np.linalg.norm(pred - gt) for 2 or 3 components: sqrt(ddot(v, v)); OpenBLAS accumulates with FMA
__device__ __forceinline__ double norm_blas(const double* __restrict__ p, const double* __restrict__ g, int D) {
  const double v0 = p[0] - g[0], v1 = p[1] - g[1];
  double acc = fma(v1, v1, v0 * v0);
  if (D == 3) { const double v2 = p[2] - g[2]; acc = fma(v2, v2, acc); }
  return sqrt(acc);
}

constexpr int kWinnerFlag = 0x40000000;

__global__ void __launch_bounds__(kWarpsPerBlock * 32) map_assign_kernel(PopnetMapArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = a.num_joints, D = a.dim;
  long long ngt_acc = 0, npos_acc = 0;
  const int warps_total = gridDim.x * kWarpsPerBlock;
  const unsigned kmask = (K >= 32) ? kFull : ((1u << K) - 1u);
  for (int f = blockIdx.x * kWarpsPerBlock + warp; f < a.num_frames; f += warps_total) {
    const int g0 = a.gt_off[f], G = a.gt_off[f + 1] - g0;
    const int p0 = a.pred_off[f], P = a.pred_off[f + 1] - p0;
    if (P == 0) continue;   // quirk kept: frames without predictions add nothing to nGT (eval_mAP.py:93-155)
    if (G == 0) {           // the reference raises here; the host wrapper does too
      for (int q = 0; q < P; ++q) {
        if (lane == 0) a.matched_gt[p0 + q] = -1;
        if (lane < K) a.labels[(size_t)(p0 + q) * K + lane] = 0;
      }
      continue;
    }
    if (lane < K)
      for (int g = 0; g < G; ++g) ngt_acc += a.gt_vis ? (a.gt_vis[(size_t)(g0 + g) * K + lane] > 0) : 1;
    // phase 1: each prediction keeps its best GT (np.argmax(pck, 1), eval_mAP.py:122-126)
    for (int q = 0; q < P; ++q) {
      double best = 0;
      int bi = -1;
      unsigned bmask = 0;
      for (int g = 0; g < G; ++g) {
        bool has = false, mt = false;
        if (lane < K) {
          has = a.gt_vis ? (a.gt_vis[(size_t)(g0 + g) * K + lane] > 0) : true;
          if (has) {
            const double d = norm_blas(a.pred + ((size_t)(p0 + q) * K + lane) * D,
                                       a.gt + ((size_t)(g0 + g) * K + lane) * D, D) / a.ref_dist[g0 + g];
            mt = d <= a.thresh;
          }
        }
        const unsigned mm = __ballot_sync(kFull, mt) & kmask;
        const int ngt = __popc(__ballot_sync(kFull, has) & kmask);
        const double pck = (double)__popc(mm) / (double)ngt;
        if (better(pck, g, best, bi)) { best = pck; bi = g; bmask = mm; }
      }
      if (lane == 0) a.matched_gt[p0 + q] = bi;                       // provisional
      if (lane < K) a.labels[(size_t)(p0 + q) * K + lane] = (bmask >> lane) & 1u;
    }
    __syncwarp();
    // phase 2: each GT takes the prediction with the highest pck among those that chose it
    // (eval_mAP.py:127-129).  Within a column pck = cnt / nGT[g], so comparing cnt is exact; a column
    // with nGT == 0 is all-NaN and np.argmax returns its first entry.
    for (int g = 0; g < G; ++g) {
      int ngt = 0;
      {
        bool has = (lane < K) && (a.gt_vis ? (a.gt_vis[(size_t)(g0 + g) * K + lane] > 0) : true);
        ngt = __popc(__ballot_sync(kFull, has) & kmask);
      }
      int bcnt = -1, bq = 0x7fffffff;
      for (int q = lane; q < P; q += 32) {
        if ((a.matched_gt[p0 + q] & ~kWinnerFlag) != g || (a.matched_gt[p0 + q] & kWinnerFlag)) continue;
        int cnt = 0;
        for (int k = 0; k < K; ++k) cnt += a.labels[(size_t)(p0 + q) * K + k];
        if (ngt == 0) cnt = 1 << 20;                                   // NaN column: first chooser wins
        if (cnt > bcnt) { bcnt = cnt; bq = q; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const int oc = __shfl_xor_sync(kFull, bcnt, o), oq = __shfl_xor_sync(kFull, bq, o);
        if (oc > bcnt || (oc == bcnt && oq < bq)) { bcnt = oc; bq = oq; }
      }
      if (bcnt > 0 && lane == 0) a.matched_gt[p0 + bq] = g | kWinnerFlag;
      __syncwarp();
    }
    // phase 3: winners keep their labels, everybody else is a false positive (eval_mAP.py:132-150)
    for (int q = 0; q < P; ++q) {
      const int mg = a.matched_gt[p0 + q];
      __syncwarp();
      if (mg & kWinnerFlag) {
        if (lane == 0) a.matched_gt[p0 + q] = mg & ~kWinnerFlag;
        if (lane < K) npos_acc += a.labels[(size_t)(p0 + q) * K + lane];
      } else {
        if (lane == 0) a.matched_gt[p0 + q] = -1;
        if (lane < K) a.labels[(size_t)(p0 + q) * K + lane] = 0;
      }
    }
  }
  if (lane < K) {
    if (ngt_acc) atomicAdd(reinterpret_cast<unsigned long long*>(a.n_gt) + lane, (unsigned long long)ngt_acc);
    if (npos_acc) atomicAdd(reinterpret_cast<unsigned long long*>(a.n_pos) + lane, (unsigned long long)npos_acc);
  }
}


// ------------------------------------------------------------------------------------------------
// AP tail: per joint a bitonic sort of (score descending, prediction index ascending), then one CTA per joint scans
// the sorted labels (cumulative true positives -> precision / recall), takes the suffix maximum of the precision
// (VOC envelope) and sums the recall steps.
// Workspace per joint: keys f64 [N2], idx i32 [N2] with N2 = SP rounded up to a power of two (>= 2048).
// ------------------------------------------------------------------------------------------------
constexpr int kSortBlock = 2048;          // elements sorted inside one CTA's shared memory (1024 threads)

__device__ __forceinline__ bool ap_before(double ka, int ia, double kb, int ib) {   // a sorts before b
  return ka > kb || (ka == kb && ia < ib);
}
__device__ __forceinline__ void ap_cmpswap(double& ka, int& ia, double& kb, int& ib, bool ascending_block) {
  // within an "ascending" bitonic block the smaller position must hold the element that sorts first
  const bool swap = ascending_block ? ap_before(kb, ib, ka, ia) : ap_before(ka, ia, kb, ib);
  if (swap) { const double tk = ka; ka = kb; kb = tk; const int ti = ia; ia = ib; ib = ti; }
}

__global__ void ap_init_kernel(const double* __restrict__ conf, int SP, int K, int N2, double* __restrict__ keys, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= N2) return;
  // padding sorts last: -inf score and indices beyond every real prediction
  keys[(size_t)j * N2 + i] = i < SP ? conf[(size_t)i * K + j] : -INFINITY;
  idx[(size_t)j * N2 + i] = i;
}

// all (k, j) stages with j < kSortBlock for k in [k_lo, k_hi] on one 2048-element block held in shared memory
__global__ void __launch_bounds__(kSortBlock / 2) ap_sort_smem_kernel(double* __restrict__ keys, int* __restrict__ idx, int N2, int k_lo, int k_hi) {
  __shared__ double sk[kSortBlock];
  __shared__ int si[kSortBlock];
  const int jn = blockIdx.y;
  const size_t base = (size_t)jn * N2 + (size_t)blockIdx.x * kSortBlock;
  for (int t = threadIdx.x; t < kSortBlock; t += blockDim.x) { sk[t] = keys[base + t]; si[t] = idx[base + t]; }
  __syncthreads();
  for (int k = k_lo; k <= k_hi; k <<= 1) {
    for (int j = (k >> 1) < kSortBlock ? (k >> 1) : (kSortBlock >> 1); j > 0; j >>= 1) {
      const int t = threadIdx.x;
      const int lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;               // pair (lo, lo + j) inside the block
      const size_t glo = (size_t)blockIdx.x * kSortBlock + lo;
      const bool asc = (glo & (size_t)k) == 0;
      ap_cmpswap(sk[lo], si[lo], sk[hi], si[hi], asc);
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < kSortBlock; t += blockDim.x) { keys[base + t] = sk[t]; idx[base + t] = si[t]; }
}

// one (k, j) stage with j >= kSortBlock: partners live in different blocks
__global__ void ap_sort_global_kernel(double* __restrict__ keys, int* __restrict__ idx, int N2, int k, int j) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, jn = blockIdx.y;
  if (t >= N2 / 2) return;
  const int lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
  double* kk = keys + (size_t)jn * N2;
  int* ii = idx + (size_t)jn * N2;
  double ka = kk[lo], kb = kk[hi];
  int ia = ii[lo], ib = ii[hi];
  const bool asc = (lo & k) == 0;
  const double oa = ka;
  const int oi = ia;
  ap_cmpswap(ka, ia, kb, ib, asc);
  if (ia != oi || ka != oa) { kk[lo] = ka; kk[hi] = kb; ii[lo] = ia; ii[hi] = ib; }
}

constexpr int kApThreads = 1024;
// one CTA per joint over the sorted order; every thread owns one contiguous chunk of positions
__global__ void __launch_bounds__(kApThreads) ap_scan_kernel(PopnetApArgs a, const int* __restrict__ idx, int N2) {
  __shared__ long long s_cnt[kApThreads];
  __shared__ double s_max[kApThreads];
  __shared__ double s_sum[kApThreads];
  const int j = blockIdx.x, t = threadIdx.x, SP = a.num_preds, K = a.num_joints;
  const int* order = idx + (size_t)j * N2;
  const int chunk = (SP + kApThreads - 1) / kApThreads;
  const int i0 = min(t * chunk, SP), i1 = min(i0 + chunk, SP);
  // pass 1: true positives per chunk -> exclusive prefix
  long long c = 0;
  for (int i = i0; i < i1; ++i) c += a.labels[(size_t)order[i] * K + j];
  s_cnt[t] = c;
  __syncthreads();
  if (t == 0) {
    long long run = 0;
    for (int q = 0; q < kApThreads; ++q) { const long long v = s_cnt[q]; s_cnt[q] = run; run += v; }
  }
  __syncthreads();
  const long long before = s_cnt[t];
  // pass 2: suffix maximum of the precision, chunk-local then across chunks (precision[i] = npos_i / (i + 1))
  double m = 0.0;                                   // mpre[-1] = 0 (eval_mAP.py:201)
  {
    long long np = before + c;
    for (int i = i1 - 1; i >= i0; --i) {
      const double prec = (double)np / (double)(i + 1);
      m = fmax(m, prec);
      np -= a.labels[(size_t)order[i] * K + j];
    }
  }
  s_max[t] = m;
  __syncthreads();
  if (t == 0) {
    double run = 0.0;
    for (int q = kApThreads - 1; q >= 0; --q) { const double v = s_max[q]; s_max[q] = run; run = fmax(run, v); }   // max over LATER chunks
  }
  __syncthreads();
  // pass 3: sum over the recall steps (true positives) of (recall_i - recall_{i-1}) * envelope_i
  const double T = (double)a.n_gt[j];
  double acc = 0.0;
  {
    // envelope inside the chunk needs the running suffix max from the right: walk backwards again
    double env = s_max[t];
    long long np = before + c;
    for (int i = i1 - 1; i >= i0; --i) {
      const double prec = (double)np / (double)(i + 1);
      env = fmax(env, prec);
      if (a.labels[(size_t)order[i] * K + j]) {
        const double step = (double)np / T - (double)(np - 1) / T;
        if (step > 0) acc += step * env;
        --np;
      }
    }
  }
  s_sum[t] = acc;
  __syncthreads();
  if (t == 0) {
    double tot = 0.0;
    for (int q = 0; q < kApThreads; ++q) tot += s_sum[q];
    a.ap[j] = tot * 100;
  }
}

__global__ void ap_mean_kernel(double* ap, int K) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    // np.mean of K doubles: pairwise sum (K < 128: 8-way unrolled order) divided by K
    double r;
    if (K < 8) { r = 0.0; for (int i = 0; i < K; ++i) r += ap[i]; }
    else {
      double q[8];
      for (int i = 0; i < 8; ++i) q[i] = ap[i];
      int i = 8;
      for (; i < K - (K % 8); i += 8) for (int u = 0; u < 8; ++u) q[u] += ap[i + u];
      r = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
      for (; i < K; ++i) r += ap[i];
    }
    ap[K] = r / (double)K;
  }
}

int ap_pow2(int n) {
  int p = kSortBlock;
  while (p < n) p <<= 1;
  return p;
}

int grid_for(int frames) {
  int blocks = (frames + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int cap = 148 * 16;   // B200: 148 SMs x 16 resident 128-thread CTAs
  return blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
}

}  // namespace

extern "C" int popnet_eval_pck(const PopnetPckArgs* args, void* stream) {
  if (!args || !args->pred2d || !args->pred_off || !args->gt2d || !args->gt_off || !args->dists ||
      !args->hit_cnt || !args->valid_cnt || args->num_frames < 0 || args->num_joints < 1 ||
      args->num_joints > 32 || ((args->pred3d == nullptr) != (args->gt3d == nullptr)))
    return POPNET_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  POPNET_CUDA_TRY(cudaMemsetAsync(args->hit_cnt, 0, sizeof(long long) * args->num_joints, st));
  POPNET_CUDA_TRY(cudaMemsetAsync(args->valid_cnt, 0, sizeof(long long) * args->num_joints, st));
  if (args->num_frames == 0) return POPNET_OK;
  pck_kernel<<<grid_for(args->num_frames), kWarpsPerBlock * 32, 0, st>>>(*args);
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}

extern "C" int popnet_eval_map_assign(const PopnetMapArgs* args, void* stream) {
  if (!args || !args->pred || !args->pred_off || !args->gt || !args->gt_off || !args->ref_dist ||
      !args->labels || !args->matched_gt || !args->n_gt || !args->n_pos || args->num_frames < 0 ||
      args->num_joints < 1 || args->num_joints > 32 || (args->dim != 2 && args->dim != 3))
    return POPNET_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  POPNET_CUDA_TRY(cudaMemsetAsync(args->n_gt, 0, sizeof(long long) * args->num_joints, st));
  POPNET_CUDA_TRY(cudaMemsetAsync(args->n_pos, 0, sizeof(long long) * args->num_joints, st));
  if (args->num_frames == 0) return POPNET_OK;
  map_assign_kernel<<<grid_for(args->num_frames), kWarpsPerBlock * 32, 0, st>>>(*args);
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}

extern "C" size_t popnet_eval_ap_workspace_bytes(int num_preds, int num_joints) {
  if (num_preds < 0 || num_joints < 1 || num_joints > 32 || num_preds > (1 << 28)) return 0;
  const size_t N2 = (size_t)ap_pow2(num_preds);
  return (size_t)num_joints * N2 * (sizeof(double) + sizeof(int));
}

extern "C" int popnet_eval_ap(const PopnetApArgs* args, void* stream) {
  if (!args || !args->n_gt || !args->ap || !args->workspace || args->num_preds < 0 || args->num_preds > (1 << 28) ||
      args->num_joints < 1 || args->num_joints > 32 || (args->num_preds > 0 && (!args->conf || !args->labels)))
    return POPNET_ERR_INVALID_ARG;
  if (args->workspace_bytes < popnet_eval_ap_workspace_bytes(args->num_preds, args->num_joints)) return POPNET_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int K = args->num_joints, SP = args->num_preds, N2 = ap_pow2(SP);
  double* keys = static_cast<double*>(args->workspace);
  int* idx = reinterpret_cast<int*>(keys + (size_t)K * N2);
  ap_init_kernel<<<dim3((N2 + 255) / 256, K), 256, 0, st>>>(args->conf, SP, K, N2, keys, idx);
  POPNET_AFTER_LAUNCH();
  // k = 2 .. kSortBlock entirely in shared memory, then per k: the global stages j >= kSortBlock, the rest in shared memory
  ap_sort_smem_kernel<<<dim3(N2 / kSortBlock, K), kSortBlock / 2, 0, st>>>(keys, idx, N2, 2, kSortBlock);
  POPNET_AFTER_LAUNCH();
  for (int k = kSortBlock * 2; k <= N2; k <<= 1) {
    for (int j = k >> 1; j >= kSortBlock; j >>= 1) {
      ap_sort_global_kernel<<<dim3((N2 / 2 + 255) / 256, K), 256, 0, st>>>(keys, idx, N2, k, j);
      POPNET_AFTER_LAUNCH();
    }
    ap_sort_smem_kernel<<<dim3(N2 / kSortBlock, K), kSortBlock / 2, 0, st>>>(keys, idx, N2, k, k);
    POPNET_AFTER_LAUNCH();
  }
  ap_scan_kernel<<<K, kApThreads, 0, st>>>(*args, idx, N2);
  POPNET_AFTER_LAUNCH();
  ap_mean_kernel<<<1, 32, 0, st>>>(args->ap, K);
  POPNET_AFTER_LAUNCH();
  return POPNET_OK;
}

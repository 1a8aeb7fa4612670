/*
 * popnet_b200 -- C ABI of the B200-native depth-pose hot path
 * (rtpose_light3d forward -> heat-map/PAF decode -> 2D-to-3D lift -> best-match PCK / mAP matching).
 *
 * The reference (oppo-us-research/PoP-Net, MP-3DHP release) has no FFI / plugin layer for this path:
 * its only native code (third_party_methods/lib/pafprocess/pafprocess.h:53-59, SWIG) is COCO-only,
 * keeps results in process globals (pafprocess.cpp:12-13) and is not called (paf_to_pose.py:7).  The
 * boundary that callers see is a set of Python signatures; every entry point below is the native
 * half of one of them and cites it.  the popnet_b200 Python package holds the Python half with the reference's
 * names and argument meaning; INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; all pointers are DEVICE pointers unless
 * the name ends in _host; `stream` is a cudaStream_t passed as void* (NULL = default stream); calls
 * enqueue work and return without synchronising; no allocation, no global state, re-entrant.
 * Return value: POPNET_OK or a negative PopnetStatus.  Nothing throws.
 */
#ifndef POPNET_B200_H_
#define POPNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POPNET_ABI_VERSION 5

#if defined(__GNUC__)
#define POPNET_API __attribute__((visibility("default")))
#else
#define POPNET_API
#endif

#define POPNET_MAX_JOINTS 24      /* K upper bound (15 MP-3DHP/ITOP, 18 COCO)                      */
#define POPNET_MAX_LIMBS 24       /* L upper bound (14 / 19)                                       */
#define POPNET_MAX_PEAKS 64       /* per joint type and frame (reference: unbounded)               */
#define POPNET_MAX_PERSONS 64     /* assembled persons per frame before pruning (ref.: unbounded)  */

typedef enum PopnetStatus {
  POPNET_OK = 0,
  POPNET_ERR_INVALID_ARG = -1,
  POPNET_ERR_UNSUPPORTED = -2,    /* shape / topology outside the compiled capacities              */
  POPNET_ERR_WORKSPACE = -3,      /* workspace too small                                           */
  POPNET_ERR_CUDA = -4,           /* a CUDA call failed; popnet_last_cuda_error() has the code     */
  POPNET_ERR_NO_DEVICE = -5
} PopnetStatus;

/* per-frame overflow flags written by popnet_decode (reference lists are unbounded; a flagged frame
 * is a parity failure by definition and is counted as such by the tests) */
#define POPNET_FLAG_PEAK_OVERFLOW 1u
#define POPNET_FLAG_PERSON_OVERFLOW 2u

POPNET_API int popnet_abi_version(void);
POPNET_API int popnet_last_cuda_error(void);
/* number of kernels this library has launched since load (all entry points); bench.py reports the
 * delta over its timed region as "gpu_launches" */
POPNET_API long long popnet_launch_count(void);
/* The only objects the library keeps between calls are plumbing: per (device, caller stream) two auxiliary streams and four
 * events that popnet_forward forks its branch chains onto (created on first use).  This destroys those of the CURRENT device
 * (after synchronising them); a later popnet_forward re-creates them.  Captured CUDA graphs that contain a forward keep
 * working (graph nodes do not reference the capture streams).  Returns the number of stream sets released. */
POPNET_API int popnet_release_streams(void);

/* ------------------------------------------------------------------------------------------------
 * Decode + lift.  Replaces, batched and on the device:
 *   paf_to_pose(heatmaps, pafs, config)              third_party_methods/lib/utils/paf_to_pose.py:354-377
 *     NMS / find_peaks / bicubic refinement          paf_to_pose.py:33-153
 *     find_connected_joints                          paf_to_pose.py:156-264
 *     group_limbs_of_same_person                     paf_to_pose.py:267-351
 *   paf_to_human_list(joint_list, assoc)             third_party_methods/lib/utils/common.py:5-32
 *   retrieve_depth_heat_weighted(c, depth, heat, 1)  third_party_methods/lib/utils/common.py:272-293
 *   de-normalise / rescale / back-project            evaluate/evaluation_rtpose_light3d_kdh3d_mpreal_ablation.py:179-263
 * The five yacs fields the reference reads (MODEL.NUM_KEYPOINTS, MODEL.DOWNSAMPLE, TEST.THRESH_HEATMAP,
 * TEST.THRESH_PAF, TEST.NUM_INTERMED_PTS_BETWEEN_KEYPOINTS; lib/config/default.py:128-130), the skeleton
 * that paf_to_pose.py:28-30 binds at import, and the camera constants (util/util_functions.py:4,11-13)
 * travel in one POD block.
 * ---------------------------------------------------------------------------------------------- */
typedef struct PopnetDecodeParams {
  int32_t num_joints;                       /* K                                                  */
  int32_t num_limbs;                        /* L                                                  */
  int32_t limbs[POPNET_MAX_LIMBS][2];       /* (src type, dst type); PAF channels (2l, 2l+1)      */
  int32_t grid_h, grid_w;                   /* 28 x 28                                            */
  int32_t stride;                           /* MODEL.DOWNSAMPLE = 8 (the only compiled value)     */
  int32_t num_intermed_pts;                 /* 10                                                 */
  float thresh_heat;                        /* 0.1, compared in fp32 like the reference           */
  float depth_mean, depth_std;              /* applied in fp32: d * std + mean                    */
  double thresh_paf;                        /* 0.05                                               */
  double input_size;                        /* 224: 2D rescale is X / input_size * w_org          */
  double w_org, h_org;
  double fx, fy, cx, cy;
  int32_t flip_y;                           /* ITOP: Y3 negated (evaluation_rtpose_light3d_itop.py:206) */
  int32_t max_peaks;                        /* <= POPNET_MAX_PEAKS                                 */
  int32_t max_persons;                      /* <= POPNET_MAX_PERSONS                               */
  int32_t depth_channels;                   /* planes per frame in `depth`: the network's third head has L + 1
                                               (rtpose_light3d.py:299-309), joint j reads plane j; 0 means K        */
  int32_t max_ctas;                         /* 0: every decode kernel may use all SMs; n > 0: at most n CTAs per kernel -- the
                                               pipelined step passes 8, the SMs its convolution grids leave free
                                               (POPNET_TUNE_RESERVE_SMS): same records either way                    */
} PopnetDecodeParams;

/* Output buffers (see popnet_decode for which may be NULL).
 * Strides use the capacities in PopnetDecodeParams (P = max_peaks, M = max_persons, K, L). */
typedef struct PopnetDecodeOut {
  int32_t* peak_count;     /* [B][K]                                                              */
  int16_t* peak_xy;        /* [B][K][P][2]   refined (X, Y) in input pixels, integers; 4-byte aligned */
  float* peak_score;       /* [B][K][P]      bicubic value at the refined maximum                  */
  int32_t* conn_count;     /* [B][L]                                                              */
  int16_t* conn_ij;        /* [B][L][P][2]   (src index, dst index) within their joint types; 4-byte aligned */
  double* conn_score;      /* [B][L][P]                                                           */
  int32_t* n_person;       /* [B]            persons that survive pruning                          */
  int16_t* person_peak;    /* [B][M][K]      peak index within the joint type, -1 = missing        */
  double* person_score;    /* [B][M]         row[-2] of person_to_joint_assoc                      */
  int32_t* person_njoint;  /* [B][M]         row[-1]                                               */
  double* pose2d;          /* [B][M][K][2]   original-resolution pixels; (-1,-1) = missing         */
  double* pose3d;          /* [B][M][K][3]   metres (camera frame)                                 */
  double* pose_conf;       /* [B][M][K]      peak score, 0 = missing                               */
  uint32_t* flags;         /* [B]            POPNET_FLAG_*                                         */
} PopnetDecodeOut;

/* heat [B][K+1][gh][gw], paf [B][2L][gh][gw], depth [B][depth_channels][gh][gw]: fp32, channel-major (the layout
 * the network writes; the reference transposes to HWC on the host, ...mpreal_ablation.py:176-178).
 * depth may be NULL (2D only: pose3d / Z are then not written).
 * peak_* and conn_* double as the stage-to-stage storage of the three decode kernels and are
 * mandatory, as are n_person and flags; person_* and pose* may be NULL. */
POPNET_API int popnet_decode(const float* heat, const float* paf, const float* depth, int batch,
                  const PopnetDecodeParams* params_host, const PopnetDecodeOut* out_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU record exchange, fused into the decode ("write straight into the registered all-gather
 * send buffer").  Frames are sharded by batch over the GPUs of one NVSwitch box, one process per GPU; the only data
 * that crosses GPUs is the pose records of every step.  Instead of an all-gather kernel behind the decode (which holds
 * SMs while it waits for the slowest rank), the assembly kernel stores every record value it produces into the gather
 * buffer of EVERY rank (plain stores to peer-mapped memory over NVLink), then publishes a step tag; a one-warp kernel
 * waits until all ranks' tags of the step have arrived.  The reference has no equivalent (single-GPU DataParallel,
 * evaluation_rtpose_light3d_kdh3d_mpreal_ablation.py:140-141).
 *
 * Memory: each rank owns ONE cudaMalloc block (popnet_p2p_alloc) that it exports with a CUDA IPC handle; peers map it
 * (popnet_p2p_open).  The host side (popnet_b200/p2p.py) lays it out as, per pipeline slot,
 *   gather[world][records_bytes]   rank r's records of the step, identical on every rank once the tags have arrived
 *   arrive[world]                  uint64 step tags, arrive[r] written by rank r
 * and points PopnetDecodeOut's record fields INTO gather[rank] of its own block, so the local record block and the
 * rank's chunk of the gathered result are the same bytes.
 * ---------------------------------------------------------------------------------------------- */
#define POPNET_MAX_PEERS 8
typedef struct PopnetPeerPush {
  int32_t world, rank;
  void* gather_base[POPNET_MAX_PEERS];            /* rank p's gather buffer of this slot, as mapped in THIS process   */
  unsigned long long* arrive[POPNET_MAX_PEERS];   /* rank p's arrive[] array of this slot, as mapped in this process  */
  size_t records_bytes;                           /* size of one rank's record block (chunk stride in gather)         */
  unsigned long long* step;                       /* local: steps completed on this slot (tag = *step + 1)            */
  unsigned int* done_counter;                     /* local, zero: CTAs of the assembly kernel that finished pushing   */
  unsigned int* status;                           /* local: set to 1 by popnet_p2p_wait when a peer's tag timed out   */
} PopnetPeerPush;

/* popnet_decode whose assembly kernel also pushes the records to the peers.  The record fields of out_host
 * (n_person, flags, person_*, pose*) must point into gather_base[rank] + rank * records_bytes. */
POPNET_API int popnet_decode_push(const float* heat, const float* paf, const float* depth, int batch,
                                  const PopnetDecodeParams* params_host, const PopnetDecodeOut* out_host,
                                  const PopnetPeerPush* push_host, void* stream);
/* one warp: returns (in stream order) when arrive[p] >= *step for every rank p of the local arrive array, or sets
 * *status = 1 after timeout_ms without progress (a dead peer must not hang the GPU) */
POPNET_API int popnet_p2p_wait(const unsigned long long* local_arrive, int world, const unsigned long long* step,
                               unsigned int* status, int timeout_ms, void* stream);
/* peer-visible device memory: cudaMalloc + cudaIpcGetMemHandle (handle_out: 64 bytes, host); popnet_p2p_open maps a
 * peer's block into this process (cudaIpcOpenMemHandle, peer access enabled lazily) */
POPNET_API int popnet_p2p_alloc(size_t bytes, void** dev_ptr_out, unsigned char* handle_out);
POPNET_API int popnet_p2p_open(const unsigned char* handle, void** peer_ptr_out);
POPNET_API int popnet_p2p_close(void* peer_ptr);
POPNET_API int popnet_p2p_free(void* dev_ptr);

/* retrieve_depth_heat_weighted(center, depthmap, heatmap, radius=1)   lib/utils/common.py:272-293, for n query points:
 * queries [n][3] = (map plane index, cx, cy) in grid cells; heat / depth are stacks of [grid_h][grid_w] fp32 planes;
 * out_z[i] = sum(d*w)/sum(w) over the clipped 3x3 window, w = max(heat,0)+1e-9, d = depth*std+mean, fp32 in NumPy's
 * pairwise order (pass mean 0, std 1 for maps that are already de-normalised, like the reference's call site). */
POPNET_API int popnet_lift_depth(const float* heat, const float* depth, const int32_t* queries, int n, int grid_h, int grid_w,
                                 float depth_mean, float depth_std, float* out_z, void* stream);

/* The sibling depth reads of the same helper family, same window, queries and de-normalisation:
 *   POPNET_LIFT_HEAT_WEIGHTED (0)  retrieve_depth_heat_weighted   lib/utils/common.py:272-293  (= popnet_lift_depth)
 *   POPNET_LIFT_MEAN          (1)  retrieve_depth_weighted        lib/utils/common.py:251-269  np.mean of the fp32 window
 *                                  (pairwise sum in fp32, divided by the count; `heat` may be NULL)
 *   POPNET_LIFT_HEAT_MAX      (2)  retrieve_depth_heat_max        lib/utils/common.py:296-318  depth at the first
 *                                  (row-major) maximum of max(heat, 0) in the window */
#define POPNET_LIFT_HEAT_WEIGHTED 0
#define POPNET_LIFT_MEAN 1
#define POPNET_LIFT_HEAT_MAX 2
POPNET_API int popnet_lift_depth_mode(const float* heat, const float* depth, const int32_t* queries, int n, int grid_h,
                                      int grid_w, float depth_mean, float depth_std, int mode, float* out_z, void* stream);
/* The same three reads with the helpers' `radius` argument (lib/utils/common.py:251,272,296): window
 * [c - radius, c + radius] clipped to the grid, 0 <= radius <= 5 (up to 121 cells: NumPy's pairwise sum is reproduced for
 * one 128-element block; larger windows return POPNET_ERR_UNSUPPORTED).  radius = 1 is what every reference call site
 * passes and what popnet_decode uses for the assembled joints. */
POPNET_API int popnet_lift_depth_window(const float* heat, const float* depth, const int32_t* queries, int n, int grid_h,
                                        int grid_w, float depth_mean, float depth_std, int mode, int radius, float* out_z,
                                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluator.  Ragged lists-of-lists are CSR-packed by the host: humans of frame f are rows
 * off[f] .. off[f+1]-1; a human is K joints of D doubles; a missing joint is (-1, -1).
 *
 * popnet_eval_pck replaces the per-frame matching of
 *   eval_human_dataset_2d / _2d_PCKh / _3d            util/eval_pck.py:20-77, 80-154, 313-374
 *   match_humans_2d / match_humans_3d                 util/eval_pck.py:266-310, 377-430
 *   compute_bbox_from_humans, bbox_ious               util/eval_pck.py:433-475
 * (identical copies: third_party_methods/evaluate/eval_pose_mp.py).
 * ---------------------------------------------------------------------------------------------- */
typedef struct PopnetPckArgs {
  const double* pred2d;        /* [SP][K][2]                                                       */
  const double* pred3d;        /* [SP][K][3] or NULL -> 2D distances                               */
  const int32_t* pred_off;     /* [N+1]                                                            */
  const double* gt2d;          /* [SG][K][2]                                                       */
  const double* gt3d;          /* [SG][K][3] or NULL                                               */
  const int32_t* gt_off;       /* [N+1]                                                            */
  const uint8_t* gt_vis;       /* [SG][K] or NULL (= all visible); 0 forces distance -1            */
  const double* gt_thresh;     /* [SG] per-GT hit threshold (PCKh: hsz*h_th, host-computed) or NULL */
  double dist_th;              /* used when gt_thresh == NULL; hit iff 0 <= d < th                  */
  double iou_th;
  int32_t num_frames;          /* N                                                                */
  int32_t num_joints;          /* K                                                                */
  double* dists;               /* out [SG][K]: matched joint distance or -1                        */
  uint8_t* hit;                /* out [SG][K] or NULL                                              */
  int32_t* matched_pred;       /* out [SG] or NULL: frame-local index of the matched prediction, -1 */
  long long* hit_cnt;          /* out [K], zeroed by the call                                      */
  long long* valid_cnt;        /* out [K], zeroed by the call: #(d >= 0)                           */
  int32_t* status;             /* out [N] or NULL: 1 = a GT human has no valid joint (the reference
                                  raises IndexError there, eval_pck.py:441-443,462)                */
} PopnetPckArgs;

POPNET_API int popnet_eval_pck(const PopnetPckArgs* args_host, void* stream);

/* popnet_eval_map_assign replaces assignGTmulti          util/eval_mAP.py:60-157
 * (callers eval_ap_mpii_v2 :272-332 and eval_ap_3D :335-395; copy in evaluate/eval_ap_mpii.py).
 * D = 2 (refDist = head size, host-computed with the reference's own expression) or 3 (refDist = 1). */
typedef struct PopnetMapArgs {
  const double* pred;          /* [SP][K][D]                                                       */
  const int32_t* pred_off;     /* [N+1]                                                            */
  const double* gt;            /* [SG][K][D]                                                       */
  const int32_t* gt_off;       /* [N+1]                                                            */
  const uint8_t* gt_vis;       /* [SG][K] or NULL (= all visible)                                  */
  const double* ref_dist;      /* [SG]                                                             */
  double thresh;               /* match iff dist / ref_dist <= thresh                              */
  int32_t num_frames, num_joints, dim;
  uint8_t* labels;             /* out [SP][K]: 1 = this predicted joint is a true positive         */
  int32_t* matched_gt;         /* out [SP]: frame-local GT index this prediction was assigned, -1   */
  long long* n_gt;             /* out [K], zeroed by the call: annotated (visible) GT joints        */
  long long* n_pos;            /* out [K], zeroed by the call: #labels == 1                         */
} PopnetMapArgs;

POPNET_API int popnet_eval_map_assign(const PopnetMapArgs* args_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * AP tail on the device.  Replaces, per joint,
 *   getRPC      util/eval_mAP.py:160-191   sort the scores of ALL predictions descending, cumulative precision / recall
 *   VOCap       util/eval_mAP.py:194-207   monotone precision envelope, area under the recall steps
 * and the `ap[j] = VOCap(...) * 100; ap[-1] = mean` lines of eval_ap_mpii / eval_ap_mpii_v2 / eval_ap_3D (:262-265).
 * Tie order: the reference sorts with np.argsort (introsort, order of equal scores unspecified); here equal scores keep
 * their input order (stable by prediction index), which is one of the orders the reference can produce.  AP is a float64:
 * compared with a tolerance (summation order differs from NumPy's pairwise sum), see tests/test_gpu_eval.py.
 * conf / labels are the [SP][K] arrays of the assignment step (labels = PopnetMapArgs.labels, n_gt = PopnetMapArgs.n_gt).
 * ---------------------------------------------------------------------------------------------- */
typedef struct PopnetApArgs {
  const double* conf;          /* [SP][K] prediction scores                                        */
  const uint8_t* labels;       /* [SP][K] 1 = true positive                                        */
  const long long* n_gt;       /* [K] annotated GT joints (recall denominator)                     */
  int32_t num_preds;           /* SP                                                               */
  int32_t num_joints;          /* K <= 32                                                          */
  double* ap;                  /* out [K+1]: AP * 100 per joint, their mean last                   */
  void* workspace;             /* popnet_eval_ap_workspace_bytes(SP, K) bytes of device memory     */
  size_t workspace_bytes;
} PopnetApArgs;

POPNET_API size_t popnet_eval_ap_workspace_bytes(int num_preds, int num_joints);
POPNET_API int popnet_eval_ap(const PopnetApArgs* args_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Network forward.  Replaces rtpose_light3d.forward   third_party_methods/lib/network/rtpose_light3d.py:326-356
 * (ResPreprocessNet :201-216, BasicBlock :56-72, make_stages :222-246) for num_stages = 2.
 * Weights are folded (eval-mode BatchNorm into scale/shift, then bf16) and packed once by
 * popnet_pack_weights from the 234 tensors of the reference state dict, passed in the canonical
 * order returned by popnet_weight_manifest (the Python mirror keeps the reference key names so a
 * reference checkpoint loads unchanged).
 * ---------------------------------------------------------------------------------------------- */
typedef struct PopnetNetConfig {
  int32_t num_parts;           /* 15                                                               */
  int32_t num_limbs;           /* 14                                                               */
  int32_t input_dim;           /* 1 (depth); the only compiled value                               */
  int32_t height, width;       /* 224 x 224; must be multiples of 8                                */
  int32_t operand_dtype;       /* POPNET_OPERAND_FP16 (what the Python mirror defaults to: holds the 1e-2 map tolerance
                                  on trained weights) or POPNET_OPERAND_BF16: storage format of weights and
                                  inter-layer activations; accumulation is always fp32 and the six output maps
                                  are always fp32                                                          */
  uint32_t tuning;             /* POPNET_TUNE_* bits; 0 = the product defaults.  Results are bit-identical for every value:
                                  the bits only choose between validated launch schedules (tests/test_forward.py) */
} PopnetNetConfig;

/* Launch-schedule switches (A/B measurements).  They replace the environment variables of ABI 2:
 * the library reads no environment. */
#define POPNET_TUNE_NO_ZIGZAG 0x1u          /* walk the tiles of every 112 x 112 layer front to back                        */
#define POPNET_TUNE_MC 0x2u                 /* N = 256 stage layers as cluster-of-two kernels with multicast weight stages */
#define POPNET_TUNE_STAGE_NACC(v) (((uint32_t)(v) & 3u) << 2)   /* 2 / 3: 256- / 384-position tiles in the 28 x 28 stages (0 = 512) */
#define POPNET_TUNE_PAIR(v) (((uint32_t)(v) & 7u) << 4)         /* 3 / 4: cta_group::2 pair kernel for the 64 -> 64 layers (0 = off) */
#define POPNET_TUNE_PAIR_RES 0x80u          /* ... including the residual layers                                           */
#define POPNET_TUNE_RESERVE_SMS(v) (((uint32_t)(v) & 7u) << 9)  /* persistent conv grids use 148 - 4 v SMs: the rest stays free for the
                                               decode of the previous batch, which runs concurrently on its own stream       */
#define POPNET_TUNE_BALANCE 0x1000u         /* persistent grids sized so that every CTA walks the same number of tiles        */
#define POPNET_TUNE_NO_PREFILL 0x2000u      /* first operand loads after the CTA-wide prologue barrier instead of before it  */
#define POPNET_TUNE_CLUSTER_ALL 0x4000u     /* every 28 x 28 stage launch as clusters of two CTAs (pairs of SMs are taken and freed together) */
#define POPNET_TUNE_CHAIN 0x100u            /* the four 64 -> 64 layers of the 112 x 112 block as ONE spatially pipelined launch
                                               (CTA slices linked by per-tile progress counters; tensors travel through the L2)
                                               instead of four launches: bit-identical, measured 3 % slower per forward      */

#define POPNET_OPERAND_BF16 0
#define POPNET_OPERAND_FP16 1

/* number of conv layers (39) and, per layer l, the element counts the packer expects */
POPNET_API int popnet_num_conv_layers(const PopnetNetConfig* cfg);
/* packed-weight blob size (device) */
POPNET_API size_t popnet_packed_weight_bytes(const PopnetNetConfig* cfg);
/* workspace (activations) size for a batch */
POPNET_API size_t popnet_workspace_bytes(const PopnetNetConfig* cfg, int batch);

/* Host-side description of one conv layer after BN folding (all host pointers, fp32):
 * weight [cout][cin][kh][kw] (PyTorch OIHW), scale[cout] and shift[cout] so that
 * y = act(scale * conv(x, weight) + shift). */
typedef struct PopnetConvHost {
  const float* weight_host;
  const float* scale_host;
  const float* shift_host;
  int32_t cout, cin, ksize;
} PopnetConvHost;

/* packs (host -> device blob): bf16 weights with `scale` folded in, fp32 shift; synchronous */
POPNET_API int popnet_pack_weights(const PopnetNetConfig* cfg, const PopnetConvHost* layers_host, int num_layers,
                        void* packed_dev, size_t packed_bytes, void* stream);

#define POPNET_FWD_IMPL_TCGEN05 0   /* product path: tcgen05/TMEM implicit GEMM                    */
#define POPNET_FWD_IMPL_SIMT 1      /* plain CUDA-core check kernel (tests / bring-up only)        */

/* x [B][1][H][W] fp32 (normalised depth).  Outputs fp32 channel-major like the reference's NCHW:
 * paf [B][2L][H/8][W/8], heat [B][K+1][..], depth [B][L+1][..]; stage1_* are the first-stage maps
 * (saved_for_loss[0..2], rtpose_light3d.py:340-342) and may be NULL. */
POPNET_API int popnet_forward(const PopnetNetConfig* cfg, const void* packed_dev, const float* x, int batch,
                   float* paf, float* heat, float* depth,
                   float* stage1_paf, float* stage1_heat, float* stage1_depth,
                   void* workspace, size_t workspace_bytes, int impl, void* stream);

/* ------------------------------------------------------------------------------------------------
 * "Next" row 1: eval-time preprocessing on the device.  Replaces
 *   Cvt2ndarray + Resize(224) (cv2.INTER_LINEAR)   lib/datasets/data_augmentation_2d3d.py:70-89, 497-522
 *   clamp [0, depth_max], (x - mean) / std          lib/datasets/datasets_kdh3d_rtpose_mpreal.py:CR229-246
 * src [B][src_h][src_w] fp32 metres -> dst [B][1][dst_h][dst_w] fp32 normalised.
 * ---------------------------------------------------------------------------------------------- */
POPNET_API int popnet_preprocess_depth(const float* src, int batch, int src_h, int src_w, float* dst, int dst_h,
                            int dst_w, float depth_max, float depth_mean, float depth_std, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POPNET_B200_H_ */

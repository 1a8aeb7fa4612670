"""ctypes mirror of include/popnet_b200.h (struct layouts and function prototypes).

Pure declarations: loading the library and failing loudly when it is missing is popnet_b200._lib's job.
"""
import ctypes as C

MAX_JOINTS = 24
MAX_LIMBS = 24
LIFT_HEAT_WEIGHTED, LIFT_MEAN, LIFT_HEAT_MAX = 0, 1, 2
MAX_PEAKS = 64
MAX_PERSONS = 64
ABI_VERSION = 5
MAX_PEERS = 8

OK = 0
STATUS_NAMES = {0: "POPNET_OK", -1: "POPNET_ERR_INVALID_ARG", -2: "POPNET_ERR_UNSUPPORTED",
                -3: "POPNET_ERR_WORKSPACE", -4: "POPNET_ERR_CUDA", -5: "POPNET_ERR_NO_DEVICE"}
FLAG_PEAK_OVERFLOW = 1
FLAG_PERSON_OVERFLOW = 2
FWD_IMPL_TCGEN05 = 0
FWD_IMPL_SIMT = 1
OPERAND_BF16 = 0
OPERAND_FP16 = 1
# PopnetNetConfig.tuning bits (include/popnet_b200.h, POPNET_TUNE_*); 0 = product defaults
TUNE_NO_ZIGZAG = 0x1
TUNE_MC = 0x2
TUNE_PAIR_RES = 0x80
TUNE_CHAIN = 0x100
TUNE_BALANCE = 0x1000
TUNE_NO_PREFILL = 0x2000
TUNE_CLUSTER_ALL = 0x4000


def TUNE_RESERVE_SMS(v):
    """persistent conv grids use 148 - 4 v SMs (v = 0..7)"""
    return (int(v) & 7) << 9


def tune_stage_nacc(v):
    return (v & 3) << 2


def tune_pair(v):
    return (v & 7) << 4


vp = C.c_void_p


class DecodeParams(C.Structure):
    _fields_ = [
        ("num_joints", C.c_int32), ("num_limbs", C.c_int32),
        ("limbs", (C.c_int32 * 2) * MAX_LIMBS),
        ("grid_h", C.c_int32), ("grid_w", C.c_int32), ("stride", C.c_int32),
        ("num_intermed_pts", C.c_int32),
        ("thresh_heat", C.c_float), ("depth_mean", C.c_float), ("depth_std", C.c_float),
        ("thresh_paf", C.c_double), ("input_size", C.c_double),
        ("w_org", C.c_double), ("h_org", C.c_double),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("flip_y", C.c_int32), ("max_peaks", C.c_int32), ("max_persons", C.c_int32),
        ("depth_channels", C.c_int32), ("max_ctas", C.c_int32),
    ]


class DecodeOut(C.Structure):
    _fields_ = [(n, vp) for n in (
        "peak_count", "peak_xy", "peak_score", "conn_count", "conn_ij", "conn_score", "n_person",
        "person_peak", "person_score", "person_njoint", "pose2d", "pose3d", "pose_conf", "flags")]


class PeerPush(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32),
                ("gather_base", vp * MAX_PEERS), ("arrive", vp * MAX_PEERS),
                ("records_bytes", C.c_size_t), ("step", vp), ("done_counter", vp), ("status", vp)]


class PckArgs(C.Structure):
    _fields_ = [
        ("pred2d", vp), ("pred3d", vp), ("pred_off", vp), ("gt2d", vp), ("gt3d", vp), ("gt_off", vp),
        ("gt_vis", vp), ("gt_thresh", vp), ("dist_th", C.c_double), ("iou_th", C.c_double),
        ("num_frames", C.c_int32), ("num_joints", C.c_int32),
        ("dists", vp), ("hit", vp), ("matched_pred", vp), ("hit_cnt", vp), ("valid_cnt", vp),
        ("status", vp),
    ]


class MapArgs(C.Structure):
    _fields_ = [
        ("pred", vp), ("pred_off", vp), ("gt", vp), ("gt_off", vp), ("gt_vis", vp), ("ref_dist", vp),
        ("thresh", C.c_double),
        ("num_frames", C.c_int32), ("num_joints", C.c_int32), ("dim", C.c_int32),
        ("labels", vp), ("matched_gt", vp), ("n_gt", vp), ("n_pos", vp),
    ]


class ApArgs(C.Structure):
    _fields_ = [
        ("conf", vp), ("labels", vp), ("n_gt", vp),
        ("num_preds", C.c_int32), ("num_joints", C.c_int32),
        ("ap", vp), ("workspace", vp), ("workspace_bytes", C.c_size_t),
    ]


class NetConfig(C.Structure):
    _fields_ = [("num_parts", C.c_int32), ("num_limbs", C.c_int32), ("input_dim", C.c_int32),
                ("height", C.c_int32), ("width", C.c_int32), ("operand_dtype", C.c_int32), ("tuning", C.c_uint32)]


class ConvHost(C.Structure):
    _fields_ = [("weight_host", vp), ("scale_host", vp), ("shift_host", vp),
                ("cout", C.c_int32), ("cin", C.c_int32), ("ksize", C.c_int32)]


#: every symbol include/popnet_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "popnet_abi_version": (C.c_int, []),
    "popnet_last_cuda_error": (C.c_int, []),
    "popnet_launch_count": (C.c_longlong, []),
    "popnet_release_streams": (C.c_int, []),
    "popnet_decode": (C.c_int, [vp, vp, vp, C.c_int, C.POINTER(DecodeParams), C.POINTER(DecodeOut), vp]),
    "popnet_decode_push": (C.c_int, [vp, vp, vp, C.c_int, C.POINTER(DecodeParams), C.POINTER(DecodeOut),
                                     C.POINTER(PeerPush), vp]),
    "popnet_p2p_wait": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, vp]),
    "popnet_p2p_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp), C.c_char_p]),
    "popnet_p2p_open": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    "popnet_p2p_close": (C.c_int, [vp]),
    "popnet_p2p_free": (C.c_int, [vp]),
    "popnet_eval_ap_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "popnet_eval_ap": (C.c_int, [C.POINTER(ApArgs), vp]),
    "popnet_lift_depth": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, vp, vp]),
    "popnet_lift_depth_mode": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, vp, vp]),
    "popnet_lift_depth_window": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                           vp, vp]),
    "popnet_eval_pck": (C.c_int, [C.POINTER(PckArgs), vp]),
    "popnet_eval_map_assign": (C.c_int, [C.POINTER(MapArgs), vp]),
    "popnet_num_conv_layers": (C.c_int, [C.POINTER(NetConfig)]),
    "popnet_packed_weight_bytes": (C.c_size_t, [C.POINTER(NetConfig)]),
    "popnet_workspace_bytes": (C.c_size_t, [C.POINTER(NetConfig), C.c_int]),
    "popnet_pack_weights": (C.c_int, [C.POINTER(NetConfig), C.POINTER(ConvHost), C.c_int, vp, C.c_size_t, vp]),
    "popnet_forward": (C.c_int, [C.POINTER(NetConfig), vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp,
                                 C.c_size_t, C.c_int, vp]),
    "popnet_preprocess_depth": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int,
                                          C.c_float, C.c_float, C.c_float, vp]),
}


def bind(lib):
    """Attach restype/argtypes to every declared symbol; raises AttributeError on a missing export."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def make_decode_params(cfg, cam, *, input_size=224, grid_hw=None, depth_channels=0, max_peaks=MAX_PEAKS,
                       max_persons=MAX_PERSONS, max_ctas=0):
    """DecodeParams from a topology.DecodeConfig and a topology.Camera.  ``grid_hw``: (rows, columns) of the maps --
    default: the square grid of a square ``input_size`` network input; ``depth_channels``: planes per frame of the
    depth tensor handed to the decode (0 = num_keypoints; the network's third head emits num_limbs + 1); ``max_ctas``: CTA
    limit per decode kernel (0 = all SMs; the pipelined step passes the SMs its convolution grids leave free)."""
    p = DecodeParams()
    p.num_joints = cfg.num_keypoints
    p.num_limbs = len(cfg.limbs)
    for l, (a, b) in enumerate(cfg.limbs):
        p.limbs[l][0] = a
        p.limbs[l][1] = b
    if grid_hw is None:
        grid_hw = (input_size // cfg.downsample, input_size // cfg.downsample)
    p.grid_h, p.grid_w = int(grid_hw[0]), int(grid_hw[1])
    p.depth_channels = int(depth_channels)
    p.max_ctas = int(max_ctas)
    p.stride = cfg.downsample
    p.num_intermed_pts = cfg.num_intermed_pts
    p.thresh_heat = cfg.thresh_heatmap
    p.depth_mean = cam.depth_mean
    p.depth_std = cam.depth_std
    p.thresh_paf = cfg.thresh_paf
    p.input_size = float(input_size)
    p.w_org = float(cam.w_org)
    p.h_org = float(cam.h_org)
    p.fx, p.fy, p.cx, p.cy = cam.fx, cam.fy, cam.cx, cam.cy
    p.flip_y = int(cam.flip_y)
    p.max_peaks = max_peaks
    p.max_persons = max_persons
    return p

"""rtpose_light3d -- host half.

Mirror of third_party_methods/lib/network/rtpose_light3d.py:249-362 of the reference: same constructor,
same ``forward(x) -> ((paf, heat, depth), saved_for_loss[6])`` contract, same module attribute names
(``model0, model1_1 .. model2_3``) and the same 234 state-dict keys, so a reference checkpoint
(``torch.save(DataParallel(model).state_dict())``, keys prefixed ``module.``) loads unchanged.

The module holds ordinary fp32 ``nn.Parameter``s only as the checkpoint container; ``forward`` never
runs a torch op on them.  On first use (and after every ``load_state_dict``) the parameters are folded
(eval-mode BatchNorm -> per-channel scale/shift) and handed to ``popnet_pack_weights``; every
convolution then runs in popnet_b200/csrc (tcgen05 implicit GEMM, 16-bit operands -- fp16 by default, bf16 selectable --,
fp32 accumulate).
No torch / cuDNN fallback exists: without the CUDA library ``forward`` raises.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import _abi, _lib

BN_EPS = 1e-5


# ---------------------------------------------------------------------------------------------
# parameter containers with the reference's attribute names (no forward of their own)
# ---------------------------------------------------------------------------------------------
def _conv(cin, cout, k, bias):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=1, padding=k // 2, bias=bias)


class _BasicBlock(nn.Module):          # rtpose_light3d.py:35-72
    def __init__(self, inplanes, planes, downsample=None):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 3, False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = _conv(planes, planes, 3, False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample


class _ResPreprocessNet(nn.Module):    # rtpose_light3d.py:124-219 with layers=[2, 1]
    def __init__(self, input_dim):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.layer1 = nn.Sequential(_BasicBlock(64, 64), _BasicBlock(64, 64))
        self.layer2 = nn.Sequential(_BasicBlock(64, 128, nn.Sequential(_conv(64, 128, 1, False), nn.BatchNorm2d(128))))
        self.conv2 = _conv(128, 128, 1, False)
        self.bn2 = nn.BatchNorm2d(128)


def _stage(spec):                      # make_stages, rtpose_light3d.py:222-246
    layers = []
    for cin, cout, k in spec[:-1]:
        layers += [_conv(cin, cout, k, True), nn.BatchNorm2d(cout), nn.LeakyReLU(0.1, inplace=True)]
    cin, cout, k = spec[-1]
    layers += [_conv(cin, cout, k, True)]
    return nn.Sequential(*layers)


def branch_specs(num_parts, num_limbs, stage):
    """(cin, cout, k) of the five convs of the three branches (rtpose_light3d.py:263-309)."""
    cin = 128 if stage == 1 else 128 + 2 * num_limbs + num_parts + 1 + num_limbs + 1
    return (
        [(cin, 256, 3), (256, 256, 3), (256, 256, 3), (256, 128, 1), (128, 2 * num_limbs, 1)],
        [(cin, 128, 3), (128, 128, 3), (128, 128, 3), (128, 128, 3), (128, num_parts + 1, 3)],
        [(cin, 128, 3), (128, 64, 3), (64, 64, 3), (64, 64, 3), (64, num_limbs + 1, 3)],
    )


class rtpose_light3d(nn.Module):
    def __init__(self, num_parts=18, num_limbs=19, num_stages=2, input_dim=3):
        super().__init__()
        if num_stages != 2:
            raise ValueError("the reference hard-wires two stages (rtpose_light3d.py:314-322)")
        # The reference's own defaults (18 parts, 19 limbs, 3 input channels: the COCO RGB configuration) are accepted by
        # its constructor but are not on the depth path this package replaces; the compiled layer plan (csrc/forward.cu,
        # make_plan) takes one depth channel and up to 15 parts / 15 limbs.  Fail here, not at the first forward.
        if input_dim != 1 or not 1 <= num_parts <= 15 or not 1 <= num_limbs <= 15:
            raise ValueError("popnet_b200 compiles the depth configuration only: input_dim=1, num_parts<=15, num_limbs<=15 "
                             "(got input_dim=%d, num_parts=%d, num_limbs=%d); the depth path uses "
                             "rtpose_light3d(15, 14, 2, input_dim=1)" % (input_dim, num_parts, num_limbs))
        self.num_parts = num_parts
        self.num_stages = num_stages
        self.num_limbs = num_limbs
        self.input_dim = input_dim
        self.model0 = _ResPreprocessNet(input_dim)
        for s in (1, 2):
            for b, spec in enumerate(branch_specs(num_parts, num_limbs, s), start=1):
                setattr(self, "model%d_%d" % (s, b), _stage(spec))
        self._initialize_weights_norm()
        self.impl = _abi.FWD_IMPL_TCGEN05
        # 16-bit storage format of weights and inter-layer activations (fp32 accumulate either way, same tensor-core rate
        # and bytes).  fp16 is the default: on a trained checkpoint it holds the 1e-2 bound of the north star with margin
        # (4.5e-3 measured on the fixture checkpoint) where bf16's 8-bit mantissa does not (3.6e-2) -- DESIGN.md section 2.
        self.operand_dtype = _abi.OPERAND_FP16     # or _abi.OPERAND_BF16; repacked automatically on the next forward
        #: launch-schedule switches (_abi.TUNE_*; include/popnet_b200.h POPNET_TUNE_*): 0 = product defaults.  Every value
        #: gives bit-identical maps; PoseEstimator.refresh() after a change (captured graphs hold the old schedule).
        self.tuning = 0
        self._packed = None        # (device blob, config key)
        self._workspace = None
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()

    def _initialize_weights_norm(self):
        # rtpose_light3d.py:358-362 (overrides the Kaiming init of block0, :160-165)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, mean=0, std=0.01)

    # -------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        # accept DataParallel checkpoints the way the eval scripts do (...mpreal_ablation.py:136-139)
        if state_dict and all(k.startswith("module.") for k in state_dict):
            state_dict = OrderedDict((k[len("module."):], v) for k, v in state_dict.items())
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._packed = None
        return out

    def conv_layers(self):
        """[(conv, bn-or-None)] in the canonical order of popnet_pack_weights (csrc/forward.cu)."""
        m0 = self.model0
        out = [(m0.conv1, m0.bn1)]
        for blk in m0.layer1:
            out += [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2)]
        blk = m0.layer2[0]
        out += [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2), (blk.downsample[0], blk.downsample[1]), (m0.conv2, m0.bn2)]
        for s in (1, 2):
            for b in (1, 2, 3):
                seq = getattr(self, "model%d_%d" % (s, b))
                for i in range(5):
                    out.append((seq[3 * i], seq[3 * i + 1] if i < 4 else None))
        return out

    @staticmethod
    def fold(conv, bn):
        """y = scale * conv_nobias(x) + shift with eval-mode BN statistics (fp32)."""
        w = conv.weight.detach().float().cpu().contiguous()
        cout = w.shape[0]
        bias = conv.bias.detach().float().cpu() if conv.bias is not None else torch.zeros(cout)
        if bn is None:
            return w, torch.ones(cout), bias.clone()
        scale = bn.weight.detach().float().cpu() / torch.sqrt(bn.running_var.detach().float().cpu() + bn.eps)
        shift = (bias - bn.running_mean.detach().float().cpu()) * scale + bn.bias.detach().float().cpu()
        return w, scale, shift

    def _net_config(self, h, w):
        return _abi.NetConfig(num_parts=self.num_parts, num_limbs=self.num_limbs, input_dim=self.input_dim,
                              height=h, width=w, operand_dtype=int(self.operand_dtype), tuning=int(self.tuning))

    def pack(self, h=224, w=224):
        """Fold + pack the current parameters into the device blob (idempotent until load_state_dict)."""
        lib = _lib.get()
        if not torch.cuda.is_available():
            raise _lib.PopnetError("popnet_b200 needs a CUDA device (no CPU fallback exists)")
        cfg = self._net_config(h, w)
        nbytes = lib.popnet_packed_weight_bytes(C.byref(cfg))
        if nbytes == 0:
            raise _lib.PopnetError("unsupported network configuration: parts=%d limbs=%d input_dim=%d %dx%d" %
                                   (self.num_parts, self.num_limbs, self.input_dim, h, w))
        layers = self.conv_layers()
        n = lib.popnet_num_conv_layers(C.byref(cfg))
        assert n == len(layers), (n, len(layers))
        arr = (_abi.ConvHost * n)()
        keep = []
        for i, (conv, bn) in enumerate(layers):
            wt, scale, shift = self.fold(conv, bn)
            wt, scale, shift = (np.ascontiguousarray(t.numpy(), np.float32) for t in (wt, scale, shift))
            keep += [wt, scale, shift]
            arr[i].weight_host = wt.ctypes.data
            arr[i].scale_host = scale.ctypes.data
            arr[i].shift_host = shift.ctypes.data
            arr[i].cout, arr[i].cin, arr[i].ksize = wt.shape[0], wt.shape[1], wt.shape[2]
        blob = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        rc = lib.popnet_pack_weights(C.byref(cfg), arr, n, C.c_void_p(blob.data_ptr()), nbytes,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "popnet_pack_weights")
        self._packed = blob
        self._packed_dtype = int(self.operand_dtype)
        self._packed_version = self._param_version()
        return blob

    def _param_version(self):
        """Changes whenever a parameter or BatchNorm statistic is modified in place (copy_, per-module loads, .to())
        or replaced: sum of the tensors' autograd version counters plus their identities."""
        ts = self.__dict__.get("_version_tensors")
        if ts is None or len(ts) != 234:
            ts = self.__dict__["_version_tensors"] = list(self.parameters()) + list(self.buffers())
        live = self._parameters_and_buffers_ids()
        if live != self.__dict__.get("_version_ids"):
            ts = self.__dict__["_version_tensors"] = list(self.parameters()) + list(self.buffers())
            self.__dict__["_version_ids"] = live
        return sum(t._version for t in ts) + live

    def _parameters_and_buffers_ids(self):
        # replaced tensors (model.float(), .to(), per-module load with assign=True) change identity, not version; the
        # first conv's weight and the last BatchNorm's statistics are representative and cheap to look at
        return id(self.model0.conv1.weight) ^ id(self.model2_3[12].weight) ^ id(self.model0.bn1.running_var)

    def map_shapes(self, B, H, W):
        """Shapes of the six fp32 output maps [paf1, heat1, depth1, paf2, heat2, depth2] for a [B, 1, H, W] input."""
        g = (H // 8, W // 8)
        K1, L2, L1 = self.num_parts + 1, 2 * self.num_limbs, self.num_limbs + 1
        return [(B, c) + g for c in (L2, K1, L1, L2, K1, L1)]

    def alloc_maps(self, B, H, W):
        return [torch.empty(s, dtype=torch.float32, device="cuda") for s in self.map_shapes(B, H, W)]

    def prepare(self, B, H, W):
        """Pack the weights if they changed and size the activation workspace; returns (config, workspace bytes).
        Everything ``forward_into`` needs besides its arguments -- call before capturing it in a CUDA graph."""
        lib = _lib.get()
        if (self._packed is None or self._packed_dtype != int(self.operand_dtype)
                or self._packed_version != self._param_version()):
            self.pack(H, W)
        cfg = self._net_config(H, W)
        ws_bytes = lib.popnet_workspace_bytes(C.byref(cfg), B)
        if ws_bytes == 0:
            raise _lib.PopnetError("unsupported input size %dx%d" % (H, W))
        if self._workspace is None or self._workspace.numel() < ws_bytes:
            self._workspace = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        return cfg, ws_bytes

    def forward_into(self, x, maps, _prepared=None):
        """popnet_forward on the current stream: x [B, 1, H, W] contiguous fp32 CUDA, maps = the six preallocated output
        tensors of ``alloc_maps``.  No allocation, no synchronisation: capturable in a CUDA graph after ``prepare``."""
        lib = _lib.get()
        B, _, H, W = x.shape
        cfg, ws_bytes = _prepared if _prepared is not None else self.prepare(B, H, W)
        p = lambda t: C.c_void_p(t.data_ptr())
        paf1, heat1, depth1, paf2, heat2, depth2 = maps
        rc = lib.popnet_forward(C.byref(cfg), p(self._packed), p(x), B, p(paf2), p(heat2), p(depth2),
                                p(paf1), p(heat1), p(depth1), p(self._workspace), ws_bytes, int(self.impl),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "popnet_forward")
        return maps

    def forward(self, x):
        """x [B, 1, H, W] fp32 CUDA tensor -> ((paf, heat, depth), [paf1, heat1, depth1, paf2, heat2, depth2])."""
        _lib.get()
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise _lib.PopnetError("rtpose_light3d.forward needs a CUDA tensor (no CPU fallback exists)")
        if x.dim() != 4 or x.shape[1] != self.input_dim:
            raise ValueError("expected [B, %d, H, W], got %s" % (self.input_dim, tuple(x.shape)))
        x = x.contiguous().float()
        B, _, H, W = x.shape
        maps = self.forward_into(x, self.alloc_maps(B, H, W))
        return (maps[3], maps[4], maps[5]), list(maps)


# ---------------------------------------------------------------------------------------------
# synthetic checkpoints (no trained weights ship with the reference)
# ---------------------------------------------------------------------------------------------
def synth_state_dict(seed=0, style="trained_like", num_parts=15, num_limbs=14, input_dim=1):
    """Deterministic state dict with the reference's keys.

    "reference": the reference's own init (N(0, 0.01) conv weights, identity BatchNorm, default biases).
    "trained_like": He-scaled conv weights, random BatchNorm statistics and affine terms, head biases that
    keep the sigmoids in their sensitive range -- activations of O(1) at every layer, so that folding,
    residual adds, LeakyReLU and the bf16 operand rounding are all exercised.
    """
    g = torch.Generator().manual_seed(seed)
    model = rtpose_light3d(num_parts, num_limbs, 2, input_dim)
    sd = OrderedDict()
    for k, v in model.state_dict().items():
        v = v.clone()
        if style == "reference":
            if v.dim() == 4:
                v = torch.randn(v.shape, generator=g) * 0.01
                fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            elif k.endswith(".bias") and not _is_bn_key(model, k):
                # nn.Conv2d default: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) of the conv this bias belongs to
                v = (torch.rand(v.shape, generator=g) * 2 - 1) / fan_in ** 0.5
        else:
            if v.dim() == 4:
                fan_in = v.shape[1] * v.shape[2] * v.shape[3]
                v = torch.randn(v.shape, generator=g) * (1.6 / fan_in) ** 0.5
            elif k.endswith("running_mean"):
                v = torch.randn(v.shape, generator=g) * 0.1
            elif k.endswith("running_var"):
                v = 0.6 + torch.rand(v.shape, generator=g) * 0.8
            elif k.endswith("num_batches_tracked"):
                pass
            elif _is_bn_key(model, k) and k.endswith(".weight"):
                v = 0.8 + torch.rand(v.shape, generator=g) * 0.4
            elif k.endswith(".bias"):
                v = torch.randn(v.shape, generator=g) * 0.1
        sd[k] = v.numpy()
    return sd


def _is_bn_key(model, key):
    mod = model
    for part in key.split(".")[:-1]:
        mod = getattr(mod, part) if not part.isdigit() else mod[int(part)]
    return isinstance(mod, nn.BatchNorm2d)

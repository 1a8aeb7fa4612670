"""ORACLE (test infrastructure, NOT product code): fp32 PyTorch restatement of the reference forward.

Follows third_party_methods/lib/network/rtpose_light3d.py of the reference:
  ResPreprocessNet._forward_impl :201-216, BasicBlock.forward :56-72, make_stages :222-246,
  rtpose_light3d.forward :326-356 (sigmoid heads, torch.cat order).
Written functionally on a plain state dict (the reference's 234 keys), so it needs neither the
reference nor popnet_b200.network.  Pinned against the reference module's own outputs through
tests/golden/forward_golden.npz (tests/golden/make_golden.py).  This is the "plain PyTorch fp32
reference of the same op" the CUDA convolution path is compared with; tolerance is stated in the tests
(max-abs <= 1e-2 on the three output maps, BASELINE.json north_star).
"""
import torch
import torch.nn.functional as F


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(v)


def _bn(sd, x, prefix):
    return F.batch_norm(x, _t(sd, prefix + ".running_mean"), _t(sd, prefix + ".running_var"),
                        _t(sd, prefix + ".weight"), _t(sd, prefix + ".bias"), False, 0.0, 1e-5)


def _block(sd, x, prefix, has_down):
    out = F.relu(_bn(sd, F.conv2d(x, _t(sd, prefix + ".conv1.weight"), None, 1, 1), prefix + ".bn1"))
    out = _bn(sd, F.conv2d(out, _t(sd, prefix + ".conv2.weight"), None, 1, 1), prefix + ".bn2")
    idt = x
    if has_down:
        idt = _bn(sd, F.conv2d(x, _t(sd, prefix + ".downsample.0.weight")), prefix + ".downsample.1")
    return F.relu(out + idt)


def _stage(sd, x, name):
    for i in range(5):
        w = _t(sd, "%s.%d.weight" % (name, 3 * i))
        x = F.conv2d(x, w, _t(sd, "%s.%d.bias" % (name, 3 * i)), 1, w.shape[2] // 2)
        if i < 4:
            x = F.leaky_relu(_bn(sd, x, "%s.%d" % (name, 3 * i + 1)), 0.1)
    return x


@torch.no_grad()
def forward(sd, x):
    """sd: state dict (numpy or torch values), x: [B,1,H,W] fp32 -> ((paf, heat, depth), saved[6])."""
    x = x if isinstance(x, torch.Tensor) else torch.from_numpy(x)
    x = x.float()
    dev = x.device
    sd = {k: _t(sd, k).to(dev) for k in sd}
    y = F.relu(_bn(sd, F.conv2d(x, sd["model0.conv1.weight"], None, 2, 3), "model0.bn1"))
    y = _block(sd, y, "model0.layer1.0", False)
    y = _block(sd, y, "model0.layer1.1", False)
    y = F.avg_pool2d(y, 3, 2, 1)
    y = _block(sd, y, "model0.layer2.0", True)
    y = F.relu(_bn(sd, F.conv2d(y, sd["model0.conv2.weight"]), "model0.bn2"))
    out1 = F.avg_pool2d(y, 3, 2, 1)
    paf1 = (_stage(sd, out1, "model1_1").sigmoid() - 0.5) * 4
    heat1 = _stage(sd, out1, "model1_2").sigmoid()
    dep1 = (_stage(sd, out1, "model1_3").sigmoid() - 0.5) * 4
    out2 = torch.cat([paf1, heat1, dep1, out1], 1)
    paf2 = (_stage(sd, out2, "model2_1").sigmoid() - 0.5) * 4
    heat2 = _stage(sd, out2, "model2_2").sigmoid()
    dep2 = (_stage(sd, out2, "model2_3").sigmoid() - 0.5) * 4
    return (paf2, heat2, dep2), [paf1, heat1, dep1, paf2, heat2, dep2]

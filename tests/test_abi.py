"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/popnet_b200.h declares,
the ctypes mirror has the same struct layouts as the C header, and the product refuses to run without CUDA."""
import ctypes as C
import os
import re
import subprocess

import pytest

from popnet_b200 import _abi, _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "popnet_b200.h")


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.get()


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"POPNET_API\s+[\w\s\*]+?\b(popnet_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
        assert s in _abi.PROTOTYPES, "no ctypes prototype for " + s
    assert sorted(_abi.PROTOTYPES) == syms
    assert lib.popnet_abi_version() == _abi.ABI_VERSION


def test_struct_layouts_match_the_header(tmp_path):
    names = {"PopnetDecodeParams": _abi.DecodeParams, "PopnetDecodeOut": _abi.DecodeOut, "PopnetPckArgs": _abi.PckArgs,
             "PopnetMapArgs": _abi.MapArgs, "PopnetNetConfig": _abi.NetConfig, "PopnetConvHost": _abi.ConvHost}
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "popnet_b200.h"\nint main(void){\n'
    for n in names:
        prog += 'printf("%s %%zu\\n", sizeof(%s));\n' % (n, n)
    prog += ('printf("off_thresh_paf %zu\\n", offsetof(PopnetDecodeParams, thresh_paf));\n'
             'printf("off_max_peaks %zu\\n", offsetof(PopnetDecodeParams, max_peaks));\n'
             'printf("off_dists %zu\\n", offsetof(PopnetPckArgs, dists));\n'
             'printf("off_labels %zu\\n", offsetof(PopnetMapArgs, labels));\nreturn 0;}\n')
    src = tmp_path / "t.c"
    src.write_text(prog)
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for n, cls in names.items():
        assert int(out[n]) == C.sizeof(cls), n
    assert int(out["off_thresh_paf"]) == _abi.DecodeParams.thresh_paf.offset
    assert int(out["off_max_peaks"]) == _abi.DecodeParams.max_peaks.offset
    assert int(out["off_dists"]) == _abi.PckArgs.dists.offset
    assert int(out["off_labels"]) == _abi.MapArgs.labels.offset


def test_argument_validation_without_a_device(lib):
    """No compute: invalid arguments are rejected before any CUDA call."""
    assert lib.popnet_decode(None, None, None, 0, None, None, None) == -1
    assert lib.popnet_eval_pck(None, None) == -1
    assert lib.popnet_eval_map_assign(None, None) == -1
    cfg = _abi.NetConfig(num_parts=15, num_limbs=14, input_dim=1, height=224, width=224, operand_dtype=0)
    assert lib.popnet_num_conv_layers(C.byref(cfg)) == 39
    assert lib.popnet_packed_weight_bytes(C.byref(cfg)) > 10_000_000
    assert lib.popnet_workspace_bytes(C.byref(cfg), 64) > 500_000_000
    bad = _abi.NetConfig(num_parts=18, num_limbs=19, input_dim=3, height=224, width=224, operand_dtype=0)
    assert lib.popnet_num_conv_layers(C.byref(bad)) == -2          # COCO / RGB: outside the compiled capacities


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from popnet_b200 import decode, evaluate
    with pytest.raises(_lib.PopnetError):
        evaluate._backend = None
        evaluate.eval_human_dataset_2d([[]], [[[[1.0, 2.0]] * 15]], 15, 10.0, 0.5)
    with pytest.raises(_lib.PopnetError):
        decode._backend = None
        import numpy as np
        decode.paf_to_pose(np.zeros((28, 28, 16), np.float32), np.zeros((28, 28, 28), np.float32),
                           __import__("popnet_b200.topology", fromlist=["DecodeConfig"]).DecodeConfig())


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under popnet_b200/ may import, load or execute it."""
    pkg = os.path.join(ROOT, "popnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt and "libpopnet_oracle" not in txt, f

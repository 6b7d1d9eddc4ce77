"""CPU: the C-ABI library loads and exports every symbol include/deepwmh_b200.h declares; its host-only
entry points (sliding-window steps, Gaussian closed form) agree with the oracle; compute entry points
fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import oracle as O
from deepwmh_b200 import _lib, build as dbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "deepwmh_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dwmh_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = dbuild.build()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SYMBOLS.keys()) == declared, "ctypes binding and header disagree"
    assert _lib.load().dwmh_version() >= 100


def test_steps_match_oracle():
    rng = np.random.default_rng(0)
    cases = [((128,) * 3, (182, 218, 182), 0.5), ((128,) * 3, (512, 512, 320), 0.5), ((128,) * 3, (128,) * 3, 0.5)]
    for _ in range(300):
        patch = tuple(int(v) for v in rng.integers(8, 160, size=3))
        img = tuple(int(p + e) for p, e in zip(patch, rng.integers(0, 300, size=3)))
        cases.append((patch, img, float(rng.choice([0.25, 0.5, 0.75, 1.0, 0.33, 0.1]))))
    for patch, img, step in cases:
        assert _lib.compute_steps(patch, img, step) == O.compute_steps_for_sliding_window(patch, img, step), (patch, img, step)


def test_steps_reject_bad_input():
    with pytest.raises(_lib.DwmhError):
        _lib.compute_steps((32, 32, 32), (16, 40, 40), 0.5)       # image smaller than patch: caller must pad
    with pytest.raises(_lib.DwmhError):
        _lib.compute_steps((32, 32, 32), (40, 40, 40), 1.5)


@pytest.mark.parametrize("patch", [(128, 128, 128), (16, 64, 48), (32, 32, 32), (40, 56, 24)])
def test_gaussian_closed_form_matches_scipy(patch):
    ref = O.get_gaussian(patch)
    got = _lib.gaussian_map(patch)
    assert got.dtype == np.float32 and got.shape == tuple(patch)
    assert got.max() == 1.0 and got.min() > 0
    rel = np.abs(got.astype(np.float64) - ref) / ref
    assert rel.max() < 3e-7, rel.max()          # at most ~2 fp32 ulp from scipy's own arithmetic


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    lib = _lib.load()
    ctx = C.c_void_p()
    d = _lib.NetDesc()
    rc = lib.dwmh_create(C.byref(ctx), 0, C.byref(d))
    assert rc != 0 and b"no CUDA device" in lib.dwmh_last_error()
    import deepwmh_b200
    with pytest.raises(_lib.DwmhError):
        deepwmh_b200.nnUNetTrainerV2(deepwmh_b200.benchmark_plans())


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = "import sys; import deepwmh_b200, deepwmh_b200.parallel; assert not any(m.split('.')[0]=='oracle' for m in sys.modules), 'oracle imported'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "deepwmh_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f


def test_net_desc_layout_matches_the_header_and_the_integration_stub(tmp_path):
    """dwmh_create copies the caller's dwmh_net_desc by value: the ctypes mirror, the header and the stub INTEGRATION.md
    shows must agree field for field (round 1's stub was one int32 short)."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "deepwmh_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu\\n", sizeof(dwmh_net_desc), offsetof(dwmh_net_desc, struct_size), offsetof(dwmh_net_desc, conv_kernel_sizes)); return 0; }\n')
    exe = str(tmp_path / "sz")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src)])
    size, off_ss, off_ck = (int(v) for v in subprocess.check_output([exe]).split())
    assert C.sizeof(_lib.NetDesc) == size and _lib.NetDesc.struct_size.offset == off_ss and _lib.NetDesc.conv_kernel_sizes.offset == off_ck
    # the stub of INTEGRATION.md, executed
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class NetDesc\(ctypes\.Structure\):.*?\n(?=\nmodel_dir)", doc, flags=re.S)
    assert m, "INTEGRATION.md no longer shows the NetDesc stub"
    ns = {}
    exec("import ctypes\n" + re.sub(r"#.*", "", m.group(0)), ns)
    assert C.sizeof(ns["NetDesc"]) == size and ns["NetDesc"].struct_size.offset == off_ss
    assert [f[0] for f in ns["NetDesc"]._fields_] == [f[0] for f in _lib.NetDesc._fields_]

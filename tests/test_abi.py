"""CPU-only checks of the drop-in boundary: struct layouts, defaults, and that the C-ABI library loads and exports
every symbol include/wildcat_b200.h declares.  No compute call is made (there is no GPU here)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from wildcat_slam_b200 import abi
from wildcat_slam_b200 import types as T

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(abi.SO_PATH):
        from wildcat_slam_b200 import build

        build.build()
    return abi.load()


def test_library_exports_every_declared_symbol(lib):
    names = abi.declared_symbols()
    assert len(names) >= 20
    out = subprocess.check_output(["nm", "-D", "--defined-only", abi.SO_PATH]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in wildcat_b200.h but not exported: {missing}"
    assert lib.wc_abi_version() == 2


def test_struct_sizes_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include "wildcat_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
        "sizeof(wc_point48),sizeof(wc_surfel),sizeof(wc_corr_idx),sizeof(wc_sample_state),sizeof(wc_imu_state),"
        "sizeof(wc_point_assign),sizeof(wc_params),sizeof(wc_solve_opts),sizeof(wc_solve_summary));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [T.POINT48.itemsize, T.SURFEL.itemsize, T.CORR.itemsize, T.SAMPLE.itemsize, T.IMU.itemsize, T.ASSIGN.itemsize,
            C.sizeof(T.Params), C.sizeof(T.SolveOpts), C.sizeof(T.SolveSummary)]
    assert got == want
    assert T.POINT48.itemsize == 48 and T.POINT48.fields["time"][1] == 24 and T.POINT48.fields["ring"][1] == 32


def test_defaults_are_the_reference_constants(lib):
    p = T.Params()
    lib.wc_default_params(C.byref(p))
    q = T.default_params()
    for name, _ in T.Params._fields_:
        a, b = getattr(p, name), getattr(q, name)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), name
        else:
            assert a == b, name
    assert np.float32(p.voxel_size) == np.float32(0.8) and p.max_layer == 2 and p.cluster_min_points == 20
    assert p.knn_candidates == 10 and p.time_diff_threshold == 0.06 and p.cauchy_a == 0.4
    o = T.SolveOpts()
    lib.wc_default_solve_opts(C.byref(o))
    r = T.default_solve_opts()
    for name, _ in T.SolveOpts._fields_:
        assert getattr(o, name) == getattr(r, name), name
    assert o.max_num_iterations == 100 and o.initial_trust_region_radius == 1e4


def test_no_cpu_fallback(lib):
    """Without a CUDA device wc_create must fail with WC_ECUDA, not fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    p = T.default_params()
    assert lib.wc_create(C.byref(p), 0, C.byref(h)) == T.WC_ECUDA and not h.value
    from wildcat_slam_b200 import odometry

    with pytest.raises(abi.WildcatError):
        odometry.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "wildcat_slam_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "wc_oracle" not in txt and "oracle/" not in txt, f

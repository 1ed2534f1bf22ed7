"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/alore_b200.h
declares, its POD layouts match the ctypes mirror, defaults equal the reference's yaml, and the library
fails loudly (no fallback) without a GPU.  No compute call is made here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from alore_legged_manipulator_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    header = (ROOT / "include" / "alore_b200.h").read_text()
    declared = set(re.findall(r"\b(alore_[a-z0-9_]+)\s*\(", header))
    declared -= {"alore_ctx", "alore_batch"}
    assert declared, "no declarations parsed"
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(capi.EXPORTED_SYMBOLS) == declared


def test_param_defaults_match_reference_yaml():
    p = capi.default_params()
    # planning_ddr_opt/back_end/config/global_planning3ms.yaml
    assert (p.smoothEps, p.safeDis, p.finalMinSafeDis, p.finalSafeDisCheckNum, p.safeReplanMaxTime) == (0.01, 0.6, 0.10, 16, 3)
    assert (p.pw_time, p.pw_acc, p.pw_domega, p.pw_collision, p.pw_moment, p.pw_mean_time, p.pw_cen_acc) == (50, 300, 300, 500000, 300, 300, 300)
    assert (p.ppw_time, p.ppw_bigpath_sdf, p.ppw_mean_time, p.ppw_moment, p.ppw_acc, p.ppw_domega) == (20, 200000, 100, 1000, 100, 100)
    assert list(p.energyWeights) == [0.33, 1.0]
    assert list(p.EqualRho) == [1e4, 1e4] and list(p.EqualRhoMax) == [1e10, 1e10] and list(p.EqualGamma) == [9.0, 9.0]
    assert list(p.EqualTolerance) == [0.01, 0.0] and list(p.CutEqualRho) == [1e3, 1e3] and list(p.CutEqualGamma) == [5.0, 5.0]
    assert list(p.CutEqualTolerance) == [0.5, 0.0]
    pl, l = p.path_lbfgs, p.lbfgs
    assert (pl.mem_size, pl.past, pl.g_epsilon, pl.min_step, pl.delta, pl.max_iterations) == (256, 2, 0.0, 0.0, 5e-2, 8000)
    assert (p.normal_past, p.shot_path_past, p.shot_path_horizon) == (2, 8, 0.5)
    assert (l.mem_size, l.past, l.g_epsilon, l.min_step, l.delta, l.max_iterations) == (256, 3, 0.0, 1e-32, 5e-4, 8000)
    # lbfgs.hpp defaults that the yaml does not override
    assert (l.max_linesearch, l.max_step, l.f_dec_coeff, l.s_curv_coeff, l.cautious_factor, l.machine_prec) == (64, 1e20, 1e-4, 0.9, 1e-6, 1e-16)
    assert p.sparseResolution == 8
    # plan_tester/config/car3ms.yaml + launch
    assert (p.max_vel, p.min_vel, p.max_acc, p.max_omega, p.max_domega, p.max_centripetal_acc) == (3.0, -3.0, 2.0, 3.0, 4.0, 50.0)
    assert p.if_directly_constrain_v_omega == 0 and p.if_standard_diff == 1 and p.n_checkpoints == 1
    assert list(p.ICR) == [0.3, -0.3, 0.2]


def test_struct_layout_matches_header_sizes():
    # sizes the C compiler produced for the PODs (catches ctypes / header drift)
    src = r'''
    #include "alore_b200.h"
    #include <stdio.h>
    int main(void) { printf("%zu %zu %zu %zu %zu\n", sizeof(alore_lbfgs_params_t), sizeof(alore_map_geom_t),
                            sizeof(alore_params_t), sizeof(alore_candidates_t), sizeof(alore_results_t)); return 0; }
    '''
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "t.c").write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", f"{d}/t", f"{d}/t.c"], check=True)
        out = subprocess.run([f"{d}/t"], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes == [C.sizeof(capi.LbfgsParams), C.sizeof(capi.MapGeom), C.sizeof(capi.Params), C.sizeof(capi.Candidates),
                     C.sizeof(capi.Results)]


def test_fails_loudly_without_gpu_or_library(tmp_path):
    with pytest.raises(RuntimeError):
        capi.load_library(tmp_path / "nope.so")
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(capi.AloreError):
            capi.Context(0)            # no CUDA device -> error, never a CPU fallback


def test_null_and_bad_arguments_return_error_codes():
    lib = capi.load_library()
    assert lib.alore_create(0, None) == -1
    assert lib.alore_esdf_update(None, None, None, 0, 0, 0, 0, None, 1) == -1
    assert lib.alore_opt_batch(None, None, None, None) == -1
    assert lib.alore_batch_run(None, None, None, None) == -1
    assert lib.alore_launch_count(None) == 0


def test_oracle_states_the_same_defaults():
    """bench.py's CPU arm takes its parameters from the oracle (it must not load the product library)."""
    import ctypes as C
    import oracle_lib
    a, b = capi.default_params(), oracle_lib.default_params()
    assert bytes(C.string_at(C.addressof(a), C.sizeof(a))) == bytes(C.string_at(C.addressof(b), C.sizeof(b)))


def test_result_store_binds_views_of_one_allocation():
    r = capi.ResultBatch(capacity=(10, 100))
    r.bind(4, 30)
    assert r.ok.shape == (4,) and r.coeffs.shape == (30, 6, 2) and r.inner_pts.shape == (26, 2) and r.piece_T.shape == (30,)
    assert r.coeffs.ctypes.data == r._base["coeffs"].ctypes.data
    r.bind(10, 100)
    assert r.inner_pts.shape == (90, 2)
    with pytest.raises(ValueError):
        r.bind(11, 30)

"""The C-ABI library loads without a GPU and exports every symbol include/asc.h declares; struct mirrors match;
compute entry points fail loudly (ASC_ERR_CUDA) instead of falling back when no B200 is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "asc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(asc_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(asc):
    lib = asc.lib()
    names = header_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_signature_table_covers_header(asc):
    from arm_slam_calib_b200.capi import SIGNATURES
    assert sorted(SIGNATURES) == header_symbols()


def test_struct_mirrors_match_the_header(asc, tmp_path):
    """sizeof/offsetof from the real header (compiled with gcc) against the ctypes mirrors."""
    from arm_slam_calib_b200 import _ctypes_defs as D
    src = tmp_path / "sz.c"
    fields = {"asc_params": ["sigma_px", "sigma_encoder", "sigma_extrinsic", "use_dead_band", "dead_band", "pose3_chart", "max_iterations",
                             "relative_error_tol", "lambda_initial", "delta_initial", "pcg_rel_tol", "pcg_max_iters", "device",
                             "pcg_cluster_max", "pcg_cluster_covis", "pcg_coarse_levels", "pcg_ml_degree", "pcg_ml_agg", "pcg_ml_top_dim"],
              "asc_iter_log": ["iteration", "pcg_iterations", "error", "solve_ms", "prepare_ms", "coarse_ms", "pcg_ms"],
              "asc_linear_stats": ["error", "cheirality"],
              "asc_sim_config": ["trajectory_size", "landmark_size_x", "max_range", "seed", "min_observations", "sim_extrinsic", "extrinsic_guess",
                                 "segment_length"]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "asc.h"', 'int main(void){']
    for st, fs in fields.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (st, st))
        for f in fs:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (st, f, st, f))
    lines.append("return 0;}")
    src.write_text("\n".join(lines))
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    mirror = {"asc_params": D.AscParams, "asc_iter_log": D.AscIterLog, "asc_linear_stats": D.AscLinearStats, "asc_sim_config": D.AscSimConfig}
    for st, cls in mirror.items():
        assert C.sizeof(cls) == int(out[st]), st
        for f in fields[st]:
            assert getattr(cls, f).offset == int(out["%s.%s" % (st, f)]), (st, f)


def test_defaults_are_the_reference_defaults(asc):
    p = asc.default_params()
    # ArmSlamCalib.h:79-86, ArmSlamCalib.cpp:197; GTSAM 3.2 optimiser defaults
    assert (p.sigma_px, p.sigma_landmark, p.cauchy_k) == (1.0, 50.0, 3.0)
    assert list(p.sigma_encoder)[:6] == [0.1] * 6 and list(p.sigma_extrinsic) == [0.1] * 6
    assert p.use_dead_band == 0 and p.dead_band == 0.05
    assert (p.max_iterations, p.relative_error_tol, p.absolute_error_tol, p.error_tol) == (100, 1e-5, 1e-5, 0.0)
    assert (p.lambda_initial, p.lambda_factor, p.lambda_upper_bound, p.lambda_lower_bound, p.min_model_fidelity) == (1e-5, 10.0, 1e5, 0.0, 1e-3)
    assert p.delta_initial == 1.0


def test_no_cpu_fallback(asc):
    """Without a GPU asc_create must fail with ASC_ERR_CUDA and say why (the product path never routes to the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = asc.lib()
    h = C.c_void_p()
    p = asc.default_params()
    rc = lib.asc_create(C.byref(p), C.byref(h))
    assert rc == asc.ASC_ERR_CUDA
    assert b"no CPU fallback" in lib.asc_last_error()
    with pytest.raises(asc.AscError):
        asc.BatchOptimizer(asc.make_sim_problem(trajectory_size=5), p)


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "arm_slam_calib_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "asc_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dirpath, f)

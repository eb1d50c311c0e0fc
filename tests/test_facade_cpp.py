"""The header-compatible C++ facade (include/cupoch/...): compiles against the C ABI with plain g++
(CPU check), and on the GPU box its RegistrationICP / KDTreeFlann results equal the oracle's."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "tests", "cpp", "_build", "facade_smoke")


def build_facade(name="facade_smoke"):
    import __graft_entry__
    __graft_entry__.build()
    exe = os.path.join(os.path.dirname(EXE), name)
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    lib = os.path.join(ROOT, "cupoch_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe, "-L" + lib,
                           "-lcupoch_b200", "-Wl,-rpath," + lib])
    return exe


def test_facade_compiles_and_links():
    for name in ("facade_smoke", "facade_filters"):
        exe = build_facade(name)
        out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
        assert "libcupoch_b200.so" in out and "not found" not in out


@pytest.mark.gpu
def test_facade_matches_oracle(orc):
    from cupoch_b200.testing import datagen
    exe = build_facade()
    tgt, tn = datagen.surface(20000, 11)
    src = datagen.make_source(tgt, datagen.gt_transform((-1.0, 1.5, 2.0), (0.01, -0.005, 0.008)), 13, 14, 5e-4)
    with tempfile.TemporaryDirectory() as d:
        src.tofile(os.path.join(d, "src.f32"))
        tgt.tofile(os.path.join(d, "tgt.f32"))
        tn.tofile(os.path.join(d, "tgt_nrm.f32"))
        r = subprocess.run([exe, d], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        T = np.fromfile(os.path.join(d, "T.f32"), np.float32).reshape(2, 4, 4)
        corr = np.fromfile(os.path.join(d, "corr.i32"), np.int32).reshape(-1, 2)
        ridx = np.fromfile(os.path.join(d, "radius_idx.i32"), np.int32)
        fit = np.fromfile(os.path.join(d, "fit.f32"), np.float32)
    ref = orc.registration_icp(orc.P2PLANE, src, tgt, 0.03, tgt_nrm=tn, relative_fitness=0, relative_rmse=0, max_iteration=6)
    assert np.linalg.norm(T[0].astype(np.float64) - ref["transformation"]) <= 1e-5
    # the generic virtual-dispatch loop (user-defined estimator) reaches the same pose as the fused path
    assert np.linalg.norm(T[1].astype(np.float64) - T[0]) <= 1e-5
    np.testing.assert_array_equal(corr, ref["correspondence_set"])
    assert abs(fit[0] - ref["fitness"]) < 1e-6 and abs(fit[2] - ref["fitness"]) < 1e-6
    oi, _, _ = orc.search(tgt, src, 1, radius=0.03, kdtree=True)
    np.testing.assert_array_equal(ridx, oi[:, 0])

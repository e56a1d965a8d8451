"""CPU-only checks of the host side: the C ABI library loads and exports every symbol the header
declares, the compat module mirrors the reference's signature, inputs generators are seeded."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    header = open(os.path.join(ROOT, "include", "gicp_b200.h")).read()
    declared = set(re.findall(r"\b(gicp[A-Z]\w+)\s*\(", header))
    from generalized_icp_b200 import _lib
    assert declared == set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.gicpVersion() >= 100
    p = _lib.GicpParams()
    assert lib.gicpDefaultParams(ctypes.byref(p)) == 0
    # defaults of gicp.py:78, :5, :11, :24
    assert (p.k, p.max_iterations, p.tolerance) == (6, 100, 1e-6)
    assert (p.max_distance_correspondence, p.max_distance_nearest_neighbors) == (150.0, 50.0)
    assert (p.lambda_tangent, p.lambda_normal) == (100.0, 10.0)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gicp as shim
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        shim.gicp(np.zeros((5, 2)), np.ones((5, 2)))
    from generalized_icp_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.gicpCreate(ctypes.byref(h), 0, 2, 1) != 0
    assert b"no CUDA device" in lib.gicpGetLastError()


def test_signature_matches_reference():
    import gicp as shim
    sig = inspect.signature(shim.gicp)
    assert list(sig.parameters) == ["source_points", "target_points", "max_iterations", "tolerance",
                                    "max_distance_correspondence", "max_distance_nearest_neighbors"]
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect._empty}
    assert d == dict(max_iterations=100, tolerance=1e-6, max_distance_correspondence=150,
                     max_distance_nearest_neighbors=50)
    T = np.array([[0.0, -1.0, 2.0], [1.0, 0.0, 3.0], [0, 0, 1.0]])
    cloud = np.array([[1.0, 0.0, 9.0], [0.0, 2.0, 9.0]])          # (N, >=2) accepted, gicp.py:176-177
    assert np.allclose(shim.apply_transformation(cloud, T), [[2.0, 4.0], [0.0, 3.0]])


def test_product_path_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "generalized-icp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                assert "oracle" not in open(os.path.join(root, f)).read().replace("the oracle", "").replace("CPU oracle", ""), f
    assert "oracle" not in open(os.path.join(ROOT, "gicp.py")).read()


def test_generators_are_seeded(golden):
    from generalized_icp_b200 import synthetic
    g = golden("config1_seed0")
    s, t = synthetic.config1_pair(0)
    assert np.array_equal(s, g["src"]) and np.array_equal(t, g["tgt"])
    assert s.shape == (90, 2) and t.shape == (87, 2)
    scans, poses = synthetic.lidar_sequence(seed=1, num_rays=90, n_scans=3)
    g2 = golden("config2_rays90_pair0")
    assert np.array_equal(np.asarray(scans[0]), g2["src"]) and np.array_equal(np.asarray(scans[1]), g2["tgt"])
    a, b, T = synthetic.patches3d_pair(n=1000, seed=5)
    a2, b2, T2 = synthetic.patches3d_pair(n=1000, seed=5)
    assert a.dtype == np.float32 and np.array_equal(a, a2) and np.array_equal(b, b2) and np.array_equal(T, T2)

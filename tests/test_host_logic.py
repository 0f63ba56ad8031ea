"""CPU: host-side logic that needs no GPU — chunking of the streamed host-input path, bench.py's workload and flop
accounting, the synthetic generators' shapes and reproducibility."""
import os
import sys

import numpy as np
import pytest

import davo_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("B,chunk", [(65536, 4096), (65536, 8192), (100, 4096), (9000, 4096), (4097, 4096), (1, 16),
                                     (1 << 20, 4096)])
def test_stream_chunks_tile_the_batch(B, chunk):
    spans = davo_b200.BFGSSolver._chunk_spans(B, chunk)
    assert spans[0][0] == 0 and spans[-1][1] == B
    for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
        assert a1 == b0 and a0 < a1
    assert all(hi - lo <= chunk for lo, hi in spans)
    if B >= 4 * chunk:  # the tail is halved down to stream_tail problems: the last solve after the last copy is short
        assert spans[-1][1] - spans[-1][0] <= 2 * davo_b200.BFGSSolver.stream_tail


def test_bench_workloads_and_flop_accounting():
    for cfg in ("cfg2", "cfg3", "cfg4", "ba"):
        b = bench.make_batch(cfg, 8, 1)
        assert b.B == 8 and b.x0.shape == (8, b.n) and b.x0.dtype == np.float32
        assert str(bench.default_B(cfg)) in bench.workload_name(cfg, bench.default_B(cfg))
    b = bench.make_batch("cfg2", 4, 1)
    fe, it = np.array([10, 20, 30, 40]), np.array([3, 5, 7, 9])
    assert bench.algorithmic_flops(b, fe, it) == 100 * 256 * 93.0 + 24 * (12 * 100 + 100)
    j = bench.make_batch("cfg3", 2, 1)
    assert bench.algorithmic_flops(j, np.array([1, 1]), np.array([0, 0])) == 2 * 1024 * 145.0
    assert bench.default_B("cfg5", 8) * 8 == bench.CFG5_TOTAL
    assert bench.make_batch("ba", 4, 1, np.float64).x0.dtype == np.float64


def test_generators_are_reproducible_and_consistent():
    a = davo_b200.synthetic.make_angle_ba(6, 8, 4, seed=5, dtype=np.float64)
    b = davo_b200.synthetic.make_angle_ba(6, 8, 4, seed=5, dtype=np.float32)
    assert a.digest() == b.digest() and a.n == 45 and a.obs.shape == (6, 4, 8, 2) and a.weights.shape == (6, 4, 8)
    assert set(np.unique(a.weights)) <= {0.0, 1.0} and a.weights[:, 0].mean() > 0.5
    s = a.slice(2, 5)
    assert s.B == 3 and np.array_equal(s.weights, a.weights[2:5]) and s.points_3d is None
    d = davo_b200.synthetic.make_distort10(5, 17, seed=3, dtype=np.float64)
    up, vp = davo_b200.synthetic.forward_numpy(d.points_3d, np.concatenate([d.truth, d.pose], axis=1))
    assert np.allclose(np.stack([up, vp], -1), d.obs, atol=1e-12)   # observations are the forward model at the truth
    with pytest.raises(ValueError):
        davo_b200.synthetic.make_angle_ba(2, 8, 1)


def test_new_entry_points_fail_loudly_without_a_gpu():
    """No CPU fallback anywhere: the round-2 surface (training solve, interpolate_alpha, the fused estimator) raises
    DavoError on a box without a CUDA device instead of computing something else."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    from davo_b200 import _lib
    with pytest.raises(_lib.DavoError):
        davo_b200.interpolate_alpha(torch.zeros(3), torch.ones(3), torch.ones(3), -torch.ones(3))
    with pytest.raises(_lib.DavoError):
        davo_b200.AnalyticObjective("sphere", (2,), 3)
    net = davo_b200.CalibrationNetwork(4, 8).eval()
    x = torch.zeros(5, 64)
    with torch.no_grad():
        y = net.estimate(x)   # CPU tensors: the fused kernel does not apply, the torch modules run
    assert y.shape == (5, 45)


def test_descriptor_layouts_match_the_header():
    """ctypes mirrors of the C structs (include/davo_b200.h): sizes and a few offsets."""
    import ctypes
    from davo_b200 import _lib
    assert ctypes.sizeof(_lib.ProblemDesc) == 80 and _lib.ProblemDesc.zoom_interpolation.offset == 40
    assert ctypes.sizeof(_lib.TrainingDesc) == 32 and _lib.TrainingDesc.drop_path_p.offset == 8
    assert _lib.TrainingDesc.seed.offset == 16 and _lib.TrainingDesc.hvp_rel_step.offset == 24
    assert ctypes.sizeof(_lib.MlpDesc) == 16
    d = _lib.make_desc(4, 8, 1, 10, "distort10", __import__("torch").float32, zoom_interpolation=True)
    assert d.zoom_interpolation == 1 and d.reserved0 == 0


def test_drop_path_restatement_is_uniform_and_reproducible():
    from oracle import train_oracle
    u = np.array([train_oracle.drop_path_uniform(1234, b, k) for b in range(200) for k in range(20)])
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.03 and abs((u > 0.1).mean() - 0.9) < 0.02
    assert train_oracle.drop_path_uniform(1234, 7, 3) == train_oracle.drop_path_uniform(1234, 7, 3)
    assert train_oracle.drop_path_uniform(1234, 7, 3) != train_oracle.drop_path_uniform(1235, 7, 3)


def test_pending_solve_handle():
    h = davo_b200.PendingSolve(value=5)
    assert h.result() == 5
    calls = []
    h = davo_b200.PendingSolve(finalize=lambda: calls.append(1) or 7)
    assert h.result() == 7 and h.result() == 7 and calls == [1]   # finalised once

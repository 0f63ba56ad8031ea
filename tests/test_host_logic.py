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

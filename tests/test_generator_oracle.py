"""The synthetic generator: numpy restatement (oracle/gen_oracle.py) pinned on CPU — Philox
known-answer vectors, the distributions the reference's dataset states, and that generated observations are the
forward model of the generated truth — and, on the GPU, csrc/generate_kernels.cu against that restatement."""
import numpy as np
import pytest
import torch

import davo_b200
from oracle import c_oracle, gen_oracle


def test_philox_known_answer():
    """Random123's kat_vectors for philox4x32-10."""
    kat = [([0, 0, 0, 0], (0, 0), [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, (0xffffffff, 0xffffffff), [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0),
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for c, k, want in kat:
        got = gen_oracle.philox4x32(np.array([c], dtype=np.uint32), k)[0]
        assert [int(v) for v in got] == want


def test_oracle_generator_distributions_and_consistency():
    r = gen_oracle.generate_distort10(4096, 32, seed=5)
    t = r["truth"]
    assert 1.0 <= t[:, 7].min() and t[:, 7].max() <= 1.5 and abs(t[:, 7].mean() - 1.25) < 0.02   # fx ~ U(1, 1.5)
    assert abs(t[:, 2].std() - 0.05) < 0.005 and abs(t[:, 5].std() - 0.005) < 0.0005              # k1, p1
    z = r["points_3d"][..., 2]
    assert z.min() >= 1.0 and abs(z.mean() - 5.0) < 0.05                                           # |4 + N| + 1
    assert np.abs(r["points_3d"][..., 0] / z).max() <= 0.5
    # observations are the forward model of the truth: the cost at the truth is zero (C oracle evaluator)
    staged = c_oracle.stage(r["points_3d"], r["obs"], r["pose"])
    f, _ = c_oracle.eval_cost_grad("distort10", r["truth"], staged, N=32)
    assert f.max() <= 1e-24
    j = gen_oracle.generate_joint(256, 24, 3, seed=6)
    f, _ = c_oracle.eval_cost_grad("joint", j["truth"], j["points_3d"], j["obs"], N=24, V=3)
    assert f.max() <= 1e-22
    f0, _ = c_oracle.eval_cost_grad("joint", j["x0"], j["points_3d"], j["obs"], N=24, V=3)
    assert f0.min() > 1e-4


def test_oracle_views_and_points_follow_the_reference_dataset():
    """data/camera_and_parameters_dataset.py:85-94,147-151: world points xy ~ 3 N, z ~ |20 + 5 N|, f' = 1/tan(U(30,
    120 deg)/2), centre ~ clamp(0.2 N, +-0.5); rotations are proper and reproduce the projections; the angular
    error of the truth is zero."""
    v = gen_oracle.generate_views_and_points(2048, 8, 4, seed=9)
    W = v["world_points"]
    assert abs(W[..., 0].std() - 3.0) < 0.1 and abs(W[..., 2].mean() - 20.0) < 0.3 and W[..., 2].min() >= 0.0
    fp = v["camera_intrinsics"][:, 0]
    assert fp.min() >= 1 / np.tan(np.pi / 3) - 1e-9 and fp.max() <= 1 / np.tan(np.pi / 12) + 1e-9
    assert np.abs(v["camera_intrinsics"][:, 1:]).max() <= 0.5
    assert 0.5 < v["visibility_mask"].mean() <= 1.0
    ang = np.linalg.norm(v["camera_orientations"], axis=-1)
    # cameras look at a common target from nearby positions: moderate rotations (rare large rolls when the 'up'
    # reference point happens to fall near a camera)
    assert np.median(ang) < 0.6 and ang.max() < np.pi
    f, g = c_oracle.eval_cost_grad("angle_ba", v["truth"], v["projected_points"], None, v["visibility_mask"], N=8, V=4)
    assert f.max() <= 1e-6, f.max()  # sum of angles (not squares): ~1e-8 per visible point in float64
    f0, _ = c_oracle.eval_cost_grad("angle_ba", v["x0"], v["projected_points"], None, v["visibility_mask"], N=8, V=4)
    assert np.median(f0) > 0.05


def test_sharded_generation_is_a_slice_of_the_global_batch():
    whole = gen_oracle.generate_distort10(64, 8, seed=3)
    part = gen_oracle.generate_distort10(16, 8, seed=3, first=32)
    for k in whole:
        assert np.array_equal(whole[k][32:48], part[k])


# ---- the CUDA generator against the restatement -------------------------------------------------------------------

def _close(a, b, tol):
    """|a - b| <= tol * max(1, |b|); values beyond the float32 range (config 4's points at z -> 0+ give
    observations of order 1e40) must have overflowed to the same infinity."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    big = np.abs(b) > 3e38
    ok = np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))
    ok = np.where(big, (a == b) | (np.isinf(a) & (np.sign(a) == np.sign(b))) | ok, ok)
    if not np.all(ok):
        i = np.unravel_index(np.argmax(~ok), ok.shape)
        print("first mismatch at", i, "device", a[i], "restatement", b[i], "mismatches", int((~ok).sum()), "of", ok.size)
    return bool(np.all(ok))


@pytest.mark.gpu
@pytest.mark.parametrize("dt,tol", [(torch.float64, 1e-10), (torch.float32, 2e-6)])
def test_device_distort10_and_joint_match_restatement(dt, tol):
    for kw in (dict(), dict(ill_conditioned=True, pathological=0.3), dict(random_pose=True, noise=1e-3)):
        d = davo_b200.synthetic.generate_distort10(300, 40, seed=11, dtype=dt, **kw)
        r = gen_oracle.generate_distort10(300, 40, seed=11, **kw)
        # rows whose points sit at z = 1e-6 have observations of order 1e12: compared relatively like the rest
        for k in ("points_3d", "obs", "pose", "x0", "truth"):
            assert _close(getattr(d, k).cpu().numpy(), r[k], tol), (k, kw)
    d = davo_b200.synthetic.generate_joint(100, 33, 3, seed=12, dtype=dt, noise=1e-3)
    r = gen_oracle.generate_joint(100, 33, 3, seed=12, noise=1e-3)
    for k in ("points_3d", "obs", "x0", "truth"):
        assert _close(getattr(d, k).cpu().numpy(), r[k], tol), k


@pytest.mark.gpu
@pytest.mark.parametrize("M,N", [(4, 8), (2, 5), (6, 12)])
def test_device_views_and_points_match_restatement(M, N):
    d = davo_b200.synthetic.generate_views_and_points(500, N, M, seed=13, dtype=torch.float64)
    r = gen_oracle.generate_views_and_points(500, N, M, seed=13)
    s = d.sample
    assert isinstance(s, davo_b200.CameraViewsAndPoints)
    assert s.projected_points.shape == (500, M, N, 2) and s.visibility_mask.shape == (500, M, N)
    assert s.camera_orientations.shape == (500, M - 1, 3) and s.world_points.shape == (500, N, 3)
    for k in ("projected_points", "camera_intrinsics", "camera_orientations", "camera_translations", "world_points"):
        assert _close(getattr(s, k).cpu().numpy(), r[k], 1e-9), k
    assert (s.visibility_mask.cpu().numpy() == r["visibility_mask"]).mean() > 0.9995
    assert _close(d.x0.cpu().numpy(), r["x0"], 1e-9) and _close(d.truth.cpu().numpy(), r["truth"], 1e-9)
    # the generated truth has zero angular error under the CUDA evaluator, and the solver recovers it from x0
    obj = davo_b200.AngleDistanceObjective(s.projected_points, s.visibility_mask)
    cost, _ = obj.evaluate(d.truth)
    assert float(cost.max()) <= 1e-6


@pytest.mark.gpu
def test_device_generation_shards_and_feeds_the_solver():
    whole = davo_b200.synthetic.generate_distort10(4096, 64, seed=21)
    part = davo_b200.synthetic.generate_distort10(1024, 64, seed=21, first_problem=2048)
    for k in ("points_3d", "obs", "x0", "truth"):
        assert torch.equal(getattr(whole, k)[2048:3072], getattr(part, k))
    obj = davo_b200.DistortionObjective(whole.points_3d, whole.obs)
    info = davo_b200.BFGSSolver(error_threshold=1e-5).eval()(whole.x0, obj, return_info=True)
    assert float(info.converged.float().mean()) > 0.99
    rel = ((info.parameters - whole.truth).abs()[:, [0, 1, 7, 9]]).max()   # cx, cy, fx, fy are well observed
    assert float(rel) < 5e-2

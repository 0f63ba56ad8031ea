"""Synthetic oracle-match generators for the BASELINE.json configurations (host side, numpy).

The distributions follow SURVEY.md §8(d), which in turn follows the reference's generator
(data/camera_and_parameters_dataset.py:48-61,85-94,147-151; that file does not parse at HEAD).
Everything is generated in float64 with ``numpy.random.default_rng(seed)`` and rounded to the
requested dtype at the end, so the float32 and float64 variants of one seed are the same problems.

This is input generation for tests and bench.py, not part of the solve path.
"""
from __future__ import annotations

from dataclasses import dataclass
import hashlib

import numpy as np

# index layout of the 16-parameter model (include/davo_b200.h, DAVO_CX..DAVO_TZ)
CX, CY, K1, K2, K3, P1, P2, FX, S, FY, RX, RY, RZ, TX, TY, TZ = range(16)


def euler_matrix(rx, ry, rz):
    """R = Rz(rz) Ry(ry) Rx(rx), camera_model/distorted_camera_model.py:38-55.  Inputs [...]; output [...,3,3]."""
    sx, cx = np.sin(rx), np.cos(rx)
    sy, cy = np.sin(ry), np.cos(ry)
    sz, cz = np.sin(rz), np.cos(rz)
    rows = [
        [cy * cz, sx * sy * cz - cx * sz, cx * sy * cz + sx * sz],
        [cy * sz, sx * sy * sz + cx * cz, cx * sy * sz - sx * cz],
        [-sy, sx * cy, cx * cy],
    ]
    return np.stack([np.stack(r, axis=-1) for r in rows], axis=-2)


def forward_numpy(points_3d: np.ndarray, params16: np.ndarray):
    """Float64 numpy statement of camera_model/distorted_camera_model.py:24-103, used only to make
    noise-free observations.  points_3d [B,N,3], params16 [B,16] -> (u'[B,N], v'[B,N])."""
    p = params16[:, None, :]
    rot = euler_matrix(params16[:, RX], params16[:, RY], params16[:, RZ])  # [B,3,3]
    xp = np.einsum("bij,bnj->bni", rot, points_3d) + params16[:, None, TX : TZ + 1]
    z = xp[..., 2].copy()
    z[z == 0] += 1e-8
    a = xp[..., 0] / z
    b = xp[..., 1] / z
    u = p[..., FX] * a + p[..., S] * b
    v = p[..., FY] * b
    r2 = u * u + v * v
    rad = 1.0 + p[..., K1] * r2 + p[..., K2] * r2 * r2 + p[..., K3] * r2 * r2 * r2
    up = u * rad + 2.0 * p[..., P1] * u * v + p[..., P2] * (r2 + 2 * u * u) + p[..., CX]
    vp = v * rad + 2.0 * p[..., P2] * u * v + p[..., P1] * (r2 + 2 * v * v) + p[..., CY]
    return up, vp


@dataclass
class CalibrationBatch:
    """One batch of independent calibration problems.

    model "distort10": points_3d [B,N,3], obs [B,N,2], pose [B,6] fixed, x0/truth [B,10].
    model "joint":     points_3d [B,N,3] world points shared by V views, obs [B,V,N,2],
                       x0/truth [B,10+6V] = intrinsics then (rx,ry,rz,tx,ty,tz) per view.
    model "angle_ba":  the driver's bundle-adjustment objective (networks/calibration_network.py:58-67):
                       points_3d None (the world points are parameters), obs [B,V,N,2] pixel coordinates,
                       weights [B,V,N] visibility mask, x0/truth [B, 3 + 3N + 6(V-1)] =
                       (f, cx, cy), N world points, V-1 translations, V-1 axis-angle rotations.
    """

    model: str
    points_3d: np.ndarray
    obs: np.ndarray
    pose: np.ndarray | None
    x0: np.ndarray
    truth: np.ndarray
    views: int = 1
    weights: np.ndarray | None = None

    @property
    def B(self) -> int:
        return self.obs.shape[0]

    @property
    def N(self) -> int:
        return self.obs.shape[-2]

    @property
    def n(self) -> int:
        return self.x0.shape[1]

    def astype(self, dtype) -> "CalibrationBatch":
        cast = lambda a: None if a is None else np.ascontiguousarray(a, dtype=dtype)
        return CalibrationBatch(self.model, cast(self.points_3d), cast(self.obs), cast(self.pose),
                                cast(self.x0), cast(self.truth), self.views, cast(self.weights))

    def slice(self, lo: int, hi: int) -> "CalibrationBatch":
        cut = lambda a: None if a is None else np.ascontiguousarray(a[lo:hi])
        return CalibrationBatch(self.model, cut(self.points_3d), cut(self.obs), cut(self.pose),
                                cut(self.x0), cut(self.truth), self.views, cut(self.weights))

    def digest(self) -> str:
        """sha256 over the float32 image of the inputs: golden fixtures store it so that a drifting
        generator is detected instead of silently comparing different problems."""
        h = hashlib.sha256()
        for a in (self.points_3d, self.obs, self.pose, self.x0, self.weights):
            if a is not None:
                h.update(np.ascontiguousarray(a, dtype=np.float32).tobytes())
        return h.hexdigest()


def _points(rng, B, N, fov):
    z = np.abs(4.0 + rng.standard_normal((B, N))) + 1.0
    xy = z[..., None] * fov * rng.uniform(-1.0, 1.0, size=(B, N, 2))
    return np.concatenate([xy, z[..., None]], axis=-1)


def _intrinsics(rng, B, k_scale=(0.05, 0.005, 0.0005), p_scale=0.005):
    th = np.zeros((B, 10))
    th[:, FX] = rng.uniform(1.0, 1.5, size=B)
    th[:, FY] = th[:, FX] * (1.0 + 0.02 * rng.standard_normal(B))
    th[:, CX] = np.clip(0.1 * rng.standard_normal(B), -0.5, 0.5)
    th[:, CY] = np.clip(0.1 * rng.standard_normal(B), -0.5, 0.5)
    th[:, K1] = k_scale[0] * rng.standard_normal(B)
    th[:, K2] = k_scale[1] * rng.standard_normal(B)
    th[:, K3] = k_scale[2] * rng.standard_normal(B)
    th[:, P1] = p_scale * rng.standard_normal(B)
    th[:, P2] = p_scale * rng.standard_normal(B)
    return th


def make_distort10(B: int, N: int = 256, seed: int = 0xB200, dtype=np.float32, fov: float = 0.5,
                   noise: float = 0.0, ill_conditioned: bool = False, pathological: float = 0.0,
                   random_pose: bool = False) -> CalibrationBatch:
    """BASELINE.json configs 2, 4 and 5.

    ill_conditioned=True is config 4: heavy distortion, doubled field of view and a start focal
    length off by a log-uniform factor in [0.3, 3]; ``pathological`` is the fraction of problems that
    additionally get points at z -> 0+ (first half) or an ascent-inducing start (second half).
    """
    rng = np.random.default_rng(seed)
    if ill_conditioned:
        fov = 1.0 if fov == 0.5 else fov
        truth = _intrinsics(rng, B, k_scale=(0.5, 0.2, 0.1), p_scale=0.05)
    else:
        truth = _intrinsics(rng, B)
    pts = _points(rng, B, N, fov)
    pose = np.zeros((B, 6))
    if random_pose:
        pose[:, :3] = 0.2 * rng.standard_normal((B, 3))
        pose[:, 3:] = 0.3 * rng.standard_normal((B, 3))
    x0 = np.zeros((B, 10))
    if ill_conditioned:
        f0 = truth[:, FX] * np.exp(rng.uniform(np.log(0.3), np.log(3.0), size=B))
    else:
        f0 = truth[:, FX] * (1.0 + 0.2 * rng.uniform(-1.0, 1.0, size=B))
    x0[:, FX] = f0
    x0[:, FY] = f0
    n_path = int(round(pathological * B))
    if n_path:
        idx = rng.permutation(B)[:n_path]
        near, ascent = idx[: n_path // 2], idx[n_path // 2 :]
        # a handful of points a hair in front of the camera plane: huge a,b -> overflow / NaN paths
        pts[near[:, None], np.arange(4)[None, :], 2] = 1e-6
        # a start on the far side of a pole of the radial polynomial: first directions are poor
        x0[ascent, K1] = 5.0
        x0[ascent, FX] *= -1.0
    full = np.concatenate([truth, pose], axis=1)
    up, vp = forward_numpy(pts, full)
    obs = np.stack([up, vp], axis=-1)
    if noise > 0.0:
        obs = obs + noise * rng.standard_normal(obs.shape)
    return CalibrationBatch("distort10", pts, obs, pose, x0, truth, 1).astype(dtype)


def make_joint(B: int, N: int = 256, V: int = 4, seed: int = 0xB200, dtype=np.float32,
               fov: float = 0.5, noise: float = 0.0) -> CalibrationBatch:
    """BASELINE.json config 3: intrinsics + a 6-DoF pose per view, V views of N shared world points."""
    rng = np.random.default_rng(seed)
    intr = _intrinsics(rng, B)
    pts = _points(rng, B, N, fov)
    poses = np.concatenate([0.2 * rng.standard_normal((B, V, 3)), 0.3 * rng.standard_normal((B, V, 3))], axis=-1)
    obs = np.empty((B, V, N, 2))
    for v in range(V):
        full = np.concatenate([intr, poses[:, v]], axis=1)
        up, vp = forward_numpy(pts, full)
        obs[:, v, :, 0] = up
        obs[:, v, :, 1] = vp
    if noise > 0.0:
        obs = obs + noise * rng.standard_normal(obs.shape)
    x0 = np.zeros((B, 10 + 6 * V))
    f0 = intr[:, FX] * (1.0 + 0.2 * rng.uniform(-1.0, 1.0, size=B))
    x0[:, FX] = f0
    x0[:, FY] = f0
    start = poses.copy()
    start[..., :3] += 0.05 * rng.standard_normal((B, V, 3))
    start[..., 3:] += 0.1 * rng.standard_normal((B, V, 3))
    x0[:, 10:] = start.reshape(B, 6 * V)
    truth = np.concatenate([intr, poses.reshape(B, 6 * V)], axis=1)
    return CalibrationBatch("joint", pts, obs, None, x0, truth, V).astype(dtype)


def rotate_axis_angle_numpy(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Rodrigues rotation of x [...,3] by the axis-angle vector w [...,3] (geometry/axis_angle_rotation.py:25-54),
    float64, for generating observations only."""
    ang = np.sqrt((w * w).sum(-1, keepdims=True))
    safe = np.where(ang > 1e-12, ang, 1.0)
    s_on = np.where(ang > 1e-12, np.sin(ang) / safe, 1.0)
    omc = np.where(ang > 1e-12, (1.0 - np.cos(ang)) / (safe * safe), 0.5)
    dot = (x * w).sum(-1, keepdims=True)
    wb, xb = np.broadcast_arrays(w, x)
    return x * np.cos(ang) + omc * dot * w + np.cross(wb, xb) * s_on


def make_angle_ba(B: int, N: int = 8, V: int = 4, seed: int = 0xB200, dtype=np.float32,
                  start_noise: float = 1.0) -> CalibrationBatch:
    """The entry script's problem (camera_calibration_from_oracle_matches.py:29-75 -> CalibrationNetwork):
    V views of N world points, unknown focal length / principal point, world points and the pose of views
    1..V-1 relative to view 0.  Distributions follow the reference generator
    (data/camera_and_parameters_dataset.py:48-61,85-94,147-151; that file does not parse at HEAD): world points
    xy ~ 3 N(0,1), z ~ |20 + 5 N(0,1)|, camera offsets ~ 3 N(0,1), f' = 1/tan(U(30,120 deg)/2), principal point
    ~ clamp(0.2 N(0,1), +-0.5); rotations are small (0.3 N(0,1) rad) so that most points stay in view.  The start
    is the truth perturbed by start_noise * (0.1, 0.05, 0.05 | 0.3 | 0.3 | 0.03) (the reference starts from an
    MLP's guess).  `weights` is the visibility mask |u|,|v| < 1 (:165-169 of the generator)."""
    if V < 2:
        raise ValueError("the bundle-adjustment objective needs at least two views")
    rng = np.random.default_rng(seed)
    X = np.concatenate([3.0 * rng.standard_normal((B, N, 2)), np.abs(20.0 + 5.0 * rng.standard_normal((B, N, 1)))], -1)
    t = 3.0 * rng.standard_normal((B, V - 1, 3))
    w = 0.3 * rng.standard_normal((B, V - 1, 3))
    fov = 3 * np.pi / 18 + (9 * np.pi / 18) * rng.uniform(size=B)
    fp = 1.0 / np.tan(fov / 2.0)
    c = np.clip(0.2 * rng.standard_normal((B, 2)), -0.5, 0.5)
    f = np.where(fp > 1.0, fp - 1.0, np.log(fp))  # f' = elu(f) + 1, geometry/homogeneous_projection.py:37
    truth = np.concatenate([f[:, None], c, X.reshape(B, -1), t.reshape(B, -1), w.reshape(B, -1)], -1)
    rel = np.concatenate([X[:, None], rotate_axis_angle_numpy(X[:, None], w[:, :, None]) + t[:, :, None]], 1)
    uv = fp[:, None, None, None] * rel[..., :2] / rel[..., 2:3] + c[:, None, None, :]
    vis = ((np.abs(uv) < 1.0).all(-1) & (rel[..., 2] > 0)).astype(np.float64)
    sig = np.concatenate([[0.1, 0.05, 0.05], np.full(3 * N, 0.3), np.full(3 * (V - 1), 0.3), np.full(3 * (V - 1), 0.03)])
    x0 = truth + start_noise * sig * rng.standard_normal(truth.shape)
    return CalibrationBatch("angle_ba", None, uv, None, x0, truth, V, vis).astype(dtype)


def stage_numpy(batch: CalibrationBatch) -> np.ndarray:
    """Host statement of davo_stage_matches for tests: [B,N,4] = {x'/z', y'/z', u*, v*} in float64."""
    assert batch.model == "distort10"
    pts = batch.points_3d.astype(np.float64)
    pose = np.zeros((batch.B, 6)) if batch.pose is None else batch.pose.astype(np.float64)
    rot = euler_matrix(pose[:, 0], pose[:, 1], pose[:, 2])
    xp = np.einsum("bij,bnj->bni", rot, pts) + pose[:, None, 3:]
    z = xp[..., 2].copy()
    z[z == 0] += 1e-8
    return np.stack([xp[..., 0] / z, xp[..., 1] / z, batch.obs[..., 0].astype(np.float64),
                     batch.obs[..., 1].astype(np.float64)], axis=-1)


def make_distort10_torch(B: int, N: int = 256, seed: int = 0xB200, device="cpu", fov: float = 0.5,
                         chunk: int = 65536) -> CalibrationBatch:
    """make_distort10 (well-conditioned, identity pose) with torch on `device`, for batches too large to
    generate with single-threaded numpy in reasonable time (BASELINE.json config 5, 1M problems).  Same
    distributions, float64 arithmetic, float32 result on the host; torch's generator, so the problems are
    not the numpy generator's."""
    import torch

    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    f64 = dict(dtype=torch.float64, device=dev, generator=gen)
    pts_out = np.empty((B, N, 3), np.float32)
    obs_out = np.empty((B, N, 2), np.float32)
    x0_out = np.zeros((B, 10), np.float32)
    truth_out = np.zeros((B, 10), np.float32)
    for lo in range(0, B, chunk):
        b = min(chunk, B - lo)
        th = torch.zeros(b, 10, dtype=torch.float64, device=dev)
        th[:, FX] = 1.0 + 0.5 * torch.rand(b, **f64)
        th[:, FY] = th[:, FX] * (1.0 + 0.02 * torch.randn(b, **f64))
        th[:, CX] = (0.1 * torch.randn(b, **f64)).clamp(-0.5, 0.5)
        th[:, CY] = (0.1 * torch.randn(b, **f64)).clamp(-0.5, 0.5)
        for col, scale in ((K1, 0.05), (K2, 0.005), (K3, 0.0005), (P1, 0.005), (P2, 0.005)):
            th[:, col] = scale * torch.randn(b, **f64)
        z = (4.0 + torch.randn(b, N, **f64)).abs() + 1.0
        xy = z[..., None] * fov * (2.0 * torch.rand(b, N, 2, **f64) - 1.0)
        a, bb = xy[..., 0] / z, xy[..., 1] / z
        p = th[:, None, :]
        u = p[..., FX] * a + p[..., S] * bb
        v = p[..., FY] * bb
        r2 = u * u + v * v
        rad = 1.0 + p[..., K1] * r2 + p[..., K2] * r2 * r2 + p[..., K3] * r2 * r2 * r2
        up = u * rad + 2.0 * p[..., P1] * u * v + p[..., P2] * (r2 + 2 * u * u) + p[..., CX]
        vp = v * rad + 2.0 * p[..., P2] * u * v + p[..., P1] * (r2 + 2 * v * v) + p[..., CY]
        f0 = th[:, FX] * (1.0 + 0.2 * (2.0 * torch.rand(b, **f64) - 1.0))
        pts_out[lo:lo + b] = torch.cat([xy, z[..., None]], dim=-1).float().cpu().numpy()
        obs_out[lo:lo + b] = torch.stack([up, vp], dim=-1).float().cpu().numpy()
        x0_out[lo:lo + b, FX] = f0.float().cpu().numpy()
        x0_out[lo:lo + b, FY] = x0_out[lo:lo + b, FX]
        truth_out[lo:lo + b] = th.float().cpu().numpy()
    return CalibrationBatch("distort10", pts_out, obs_out, np.zeros((B, 6), np.float32), x0_out, truth_out, 1)


# ---- on-device generation (csrc/generate_kernels.cu through the C-ABI) ---------------------------------------------

@dataclass
class DeviceBatch:
    """A CalibrationBatch whose arrays are torch tensors resident on the GPU that generated them."""

    model: str
    points_3d: "object"
    obs: "object"
    pose: "object"
    x0: "object"
    truth: "object"
    views: int = 1
    weights: "object" = None
    sample: "object" = None  # CameraViewsAndPoints for model "angle_ba"

    @property
    def B(self) -> int:
        return self.obs.shape[0]

    @property
    def N(self) -> int:
        return self.obs.shape[-2]

    @property
    def n(self) -> int:
        return self.x0.shape[1]

    def numpy(self) -> CalibrationBatch:
        h = lambda t: None if t is None else t.cpu().numpy()
        return CalibrationBatch(self.model, h(self.points_3d), h(self.obs), h(self.pose), h(self.x0), h(self.truth),
                                self.views, h(self.weights))


def _gen_desc(B, N, V, dtype, seed, first, *, fov=0.5, noise=0.0, ill_conditioned=False, pathological=0.0,
              random_pose=False, start_noise=1.0, min_camera_distance=0.1):
    from . import _lib
    return _lib.GeneratorDesc(int(B), int(N), int(V), _lib.dtype_code(dtype), int(bool(ill_conditioned)),
                              int(bool(random_pose)), int(seed) & (2 ** 64 - 1), int(first), float(fov), float(noise),
                              float(pathological), float(start_noise), float(min_camera_distance))


def generate_distort10(B: int, N: int = 256, seed: int = 0xB200, dtype=None, device=None, first_problem: int = 0,
                       fov: float = 0.5, noise: float = 0.0, ill_conditioned: bool = False, pathological: float = 0.0,
                       random_pose: bool = False) -> DeviceBatch:
    """BASELINE configs 2, 4, 5 generated in HBM (davo_generate_distort10): the distributions of make_distort10
    (SURVEY.md 8(d)), Philox randomness keyed by (seed, first_problem + row) so that shards are rows of one batch."""
    import ctypes

    import torch

    from . import _lib
    dtype = dtype or torch.float32
    device = _lib.require_cuda() if device is None else torch.device(device)
    if ill_conditioned and fov == 0.5:
        fov = 1.0
    with torch.cuda.device(device):
        new = lambda *shape: torch.empty(*shape, dtype=dtype, device=device)
        pts, obs, pose, x0, truth = new(B, N, 3), new(B, N, 2), new(B, 6), new(B, 10), new(B, 10)
        d = _gen_desc(B, N, 1, dtype, seed, first_problem, fov=fov, noise=noise, ill_conditioned=ill_conditioned,
                      pathological=pathological, random_pose=random_pose)
        st = _lib.lib().davo_generate_distort10(ctypes.byref(d), _lib.ptr(pts), _lib.ptr(obs), _lib.ptr(pose),
                                                _lib.ptr(x0), _lib.ptr(truth), _lib.stream_ptr())
    _lib.check(st, "davo_generate_distort10")
    return DeviceBatch("distort10", pts, obs, pose, x0, truth, 1)


def generate_joint(B: int, N: int = 256, V: int = 4, seed: int = 0xB200, dtype=None, device=None,
                   first_problem: int = 0, fov: float = 0.5, noise: float = 0.0) -> DeviceBatch:
    """BASELINE config 3 generated in HBM (davo_generate_joint), distributions of make_joint."""
    import ctypes

    import torch

    from . import _lib
    dtype = dtype or torch.float32
    device = _lib.require_cuda() if device is None else torch.device(device)
    with torch.cuda.device(device):
        new = lambda *shape: torch.empty(*shape, dtype=dtype, device=device)
        pts, obs, x0, truth = new(B, N, 3), new(B, V, N, 2), new(B, 10 + 6 * V), new(B, 10 + 6 * V)
        d = _gen_desc(B, N, V, dtype, seed, first_problem, fov=fov, noise=noise)
        st = _lib.lib().davo_generate_joint(ctypes.byref(d), _lib.ptr(pts), _lib.ptr(obs), _lib.ptr(x0),
                                            _lib.ptr(truth), _lib.stream_ptr())
    _lib.check(st, "davo_generate_joint")
    return DeviceBatch("joint", pts, obs, None, x0, truth, V)


def generate_views_and_points(B: int, num_points: int = 8, num_views: int = 4, seed: int = 0xB200, dtype=None,
                              device=None, first_problem: int = 0, start_noise: float = 1.0,
                              min_camera_distance: float = 0.1) -> DeviceBatch:
    """The entry script's batches generated in HBM (davo_generate_views_and_points): the reference dataset's
    distributions (data/camera_and_parameters_dataset.py:85-151) in its CameraViewsAndPoints layout (`.sample`),
    plus the solver's start / truth parameter vectors (the reference starts from an MLP's guess)."""
    import ctypes

    import torch

    from . import _lib
    from .base_types import CameraViewsAndPoints
    dtype = dtype or torch.float32
    device = _lib.require_cuda() if device is None else torch.device(device)
    M, N = int(num_views), int(num_points)
    n = 3 + 3 * N + 6 * (M - 1)
    with torch.cuda.device(device):
        new = lambda *shape: torch.empty(*shape, dtype=dtype, device=device)
        proj, vis, intr, orient, trans, world = new(B, M, N, 2), new(B, M, N), new(B, 3), new(B, M - 1, 3), \
            new(B, M - 1, 3), new(B, N, 3)
        x0, truth = new(B, n), new(B, n)
        d = _gen_desc(B, N, M, dtype, seed, first_problem, start_noise=start_noise,
                      min_camera_distance=min_camera_distance)
        st = _lib.lib().davo_generate_views_and_points(
            ctypes.byref(d), _lib.ptr(proj), _lib.ptr(vis), _lib.ptr(intr), _lib.ptr(orient), _lib.ptr(trans),
            _lib.ptr(world), _lib.ptr(x0), _lib.ptr(truth), _lib.stream_ptr())
    _lib.check(st, "davo_generate_views_and_points")
    sample = CameraViewsAndPoints(proj, vis, intr, orient, trans, world)
    return DeviceBatch("angle_ba", None, proj, None, x0, truth, M, vis, sample)

"""numpy restatement of the on-device synthetic generator — TEST INFRASTRUCTURE (checker only).

Follows the same published pieces as csrc/generate_kernels.cu: Philox4x32-10 (Salmon, Moraes, Dror, Shaw,
"Parallel random numbers: as easy as 1, 2, 3", SC'11: multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments
0x9E3779B9 / 0xBB67AE85), Box-Muller, and the distributions of SURVEY.md 8(d) / the reference dataset
(data/camera_and_parameters_dataset.py:85-151, camera axes made right-handed; see the kernel's header comment).
Counter = (problem index lo, hi, stream, element); key = seed.  Vectorised over problems, float64.

Pinned by test_philox_known_answer (the Random123 known-answer vectors) in tests/test_generator_oracle.py.
"""
from __future__ import annotations

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
S_SCALARS, S_POINTS, S_NOISE, S_VIEWS, S_START = 0, 1, 2, 3, 4
CX, CY, K1, K2, K3, P1, P2, FX, S, FY, RX, RY, RZ, TX, TY, TZ = range(16)


def philox4x32(c, key):
    """c: uint32 array [..., 4], key: (k0, k1) -> uint32 array [..., 4]."""
    c = [np.asarray(c[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & mask, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & mask]
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def draw(seed, problems, stream, element):
    """uint32 [..., 4] for broadcastable arrays of global problem indices and element indices."""
    problems, element = np.broadcast_arrays(np.asarray(problems, np.uint64), np.asarray(element, np.uint64))
    c = np.stack([problems & np.uint64(0xFFFFFFFF), problems >> np.uint64(32),
                  np.full(problems.shape, stream, np.uint64), element], axis=-1)
    return philox4x32(c, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))


def u01(r):
    return (r.astype(np.float64) + 0.5) / 4294967296.0


def normal4(w):
    out = np.empty(w.shape, np.float64)
    for k in (0, 2):
        r = np.sqrt(-2.0 * np.log(u01(w[..., k])))
        ang = 2.0 * np.pi * u01(w[..., k + 1])
        out[..., k] = r * np.cos(ang)
        out[..., k + 1] = r * np.sin(ang)
    return out


def euler_rows(rx, ry, rz):
    sx, cx, sy, cy, sz, cz = np.sin(rx), np.cos(rx), np.sin(ry), np.cos(ry), np.sin(rz), np.cos(rz)
    return np.stack([cy * cz, sx * sy * cz - cx * sz, cx * sy * cz + sx * sz,
                     cy * sz, sx * sy * sz + cx * cz, cx * sy * sz - sx * cz,
                     -sy, sx * cy, cx * cy], axis=-1).reshape(rx.shape + (3, 3))


def forward16(th, R, P):
    """th [B,16], R [B,3,3], P [B,N,3] -> (u', v') [B,N] (camera_model/distorted_camera_model.py:24-103)."""
    xp = np.einsum("bij,bnj->bni", R, P) + th[:, None, TX:TZ + 1]
    z = xp[..., 2].copy()
    z[z == 0] += 1e-8
    a, b = xp[..., 0] / z, xp[..., 1] / z
    p = th[:, None, :]
    u = p[..., FX] * a + p[..., S] * b
    v = p[..., FY] * b
    r2 = u * u + v * v
    rad = 1.0 + p[..., K1] * r2 + p[..., K2] * r2 * r2 + p[..., K3] * r2 * r2 * r2
    return (u * rad + 2.0 * p[..., P1] * u * v + p[..., P2] * (r2 + 2.0 * u * u) + p[..., CX],
            v * rad + 2.0 * p[..., P2] * u * v + p[..., P1] * (r2 + 2.0 * v * v) + p[..., CY])


def _intrinsics(seed, gb, ill):
    s0 = draw(seed, gb, S_SCALARS, 0)
    n1 = normal4(draw(seed, gb, S_SCALARS, 1))
    n2 = normal4(draw(seed, gb, S_SCALARS, 2))
    k1s, k2s, k3s, ps = (0.5, 0.2, 0.1, 0.05) if ill else (0.05, 0.005, 0.0005, 0.005)
    B = len(gb)
    th = np.zeros((B, 16))
    fx = 1.0 + 0.5 * u01(s0[:, 0])
    th[:, FX] = fx
    th[:, FY] = fx * (1.0 + 0.02 * n1[:, 0])
    th[:, CX] = np.clip(0.1 * n1[:, 1], -0.5, 0.5)
    th[:, CY] = np.clip(0.1 * n1[:, 2], -0.5, 0.5)
    th[:, K1], th[:, K2], th[:, K3] = k1s * n1[:, 3], k2s * n2[:, 0], k3s * n2[:, 1]
    th[:, P1], th[:, P2] = ps * n2[:, 2], ps * n2[:, 3]
    uf = u01(s0[:, 1])
    f0 = fx * np.exp(np.log(0.3) + uf * (np.log(3.0) - np.log(0.3))) if ill else fx * (1.0 + 0.2 * (2.0 * uf - 1.0))
    x0 = np.zeros((B, 10))
    x0[:, FX] = x0[:, FY] = f0
    return th, x0, u01(s0[:, 2])


def _points(seed, gb, N, fov):
    w = draw(seed, gb[:, None], S_POINTS, np.arange(N)[None, :])
    nn = normal4(w)
    Z = np.abs(4.0 + nn[..., 0]) + 1.0
    X = Z * fov * (2.0 * u01(w[..., 2]) - 1.0)
    Y = Z * fov * (2.0 * u01(w[..., 3]) - 1.0)
    return np.stack([X, Y, Z], axis=-1)


def generate_distort10(B, N, seed, first=0, fov=0.5, noise=0.0, ill_conditioned=False, pathological=0.0,
                       random_pose=False):
    gb = first + np.arange(B, dtype=np.uint64)
    if ill_conditioned and fov == 0.5:
        fov = 1.0
    th, x0, u_path = _intrinsics(seed, gb, ill_conditioned)
    if random_pose:
        a = normal4(draw(seed, gb, S_SCALARS, 3))
        c = normal4(draw(seed, gb, S_SCALARS, 4))
        th[:, RX:RZ + 1] = 0.2 * a[:, :3]
        th[:, TX], th[:, TY], th[:, TZ] = 0.3 * a[:, 3], 0.3 * c[:, 0], 0.3 * c[:, 1]
    near = u_path < 0.5 * pathological
    ascent = ~near & (u_path < pathological)
    x0[ascent, K1] = 5.0
    x0[ascent, FX] *= -1.0
    P = _points(seed, gb, N, fov)
    P[near, :4, 2] = 1e-6
    up, vp = forward16(th, euler_rows(th[:, RX], th[:, RY], th[:, RZ]), P)
    obs = np.stack([up, vp], axis=-1)
    if noise > 0:
        nn = normal4(draw(seed, gb[:, None], S_NOISE, np.arange(N)[None, :]))
        obs = obs + noise * nn[..., :2]
    return dict(points_3d=P, obs=obs, pose=th[:, 10:], x0=x0, truth=th[:, :10])


def generate_joint(B, N, V, seed, first=0, fov=0.5, noise=0.0):
    gb = first + np.arange(B, dtype=np.uint64)
    th, x0_10, _ = _intrinsics(seed, gb, False)
    P = _points(seed, gb, N, fov)
    n = 10 + 6 * V
    x0, truth, obs = np.zeros((B, n)), np.zeros((B, n)), np.zeros((B, V, N, 2))
    x0[:, :10], truth[:, :10] = x0_10, th[:, :10]
    for v in range(V):
        a = normal4(draw(seed, gb, S_VIEWS, 2 * v))
        c = normal4(draw(seed, gb, S_VIEWS, 2 * v + 1))
        sa = normal4(draw(seed, gb, S_START, 2 * v))
        sc = normal4(draw(seed, gb, S_START, 2 * v + 1))
        pose = np.concatenate([0.2 * a[:, :3], 0.3 * a[:, 3:4], 0.3 * c[:, :2]], axis=1)
        start = pose + np.concatenate([0.05 * sa[:, :3], 0.1 * sa[:, 3:4], 0.1 * sc[:, :2]], axis=1)
        th[:, 10:] = pose
        truth[:, 10 + 6 * v:16 + 6 * v] = pose
        x0[:, 10 + 6 * v:16 + 6 * v] = start
        up, vp = forward16(th, euler_rows(pose[:, 0], pose[:, 1], pose[:, 2]), P)
        o = np.stack([up, vp], axis=-1)
        if noise > 0:
            nn = normal4(draw(seed, gb[:, None], S_NOISE, (v * N + np.arange(N))[None, :]))
            o = o + noise * nn[..., :2]
        obs[:, v] = o
    return dict(points_3d=P, obs=obs, x0=x0, truth=truth)


def generate_views_and_points(B, N, M, seed, first=0, start_noise=1.0, min_camera_distance=0.1):
    gb = first + np.arange(B, dtype=np.uint64)
    nn = normal4(draw(seed, gb[:, None], S_POINTS, np.arange(N)[None, :]))
    Xw = np.stack([3.0 * nn[..., 0], 3.0 * nn[..., 1], np.abs(20.0 + 5.0 * nn[..., 2])], axis=-1)  # [B,N,3]
    s0 = draw(seed, gb, S_SCALARS, 0)
    ns = normal4(draw(seed, gb, S_SCALARS, 1))
    nt = normal4(draw(seed, gb, S_SCALARS, 2))
    fov = 3.0 * np.pi / 18.0 + (9.0 * np.pi / 18.0) * u01(s0[:, 0])
    fp = 1.0 / np.tan(0.5 * fov)
    pc = np.clip(0.2 * ns[:, :2], -0.5, 0.5)
    up_distance = np.abs(20.0 + 5.0 * ns[:, 2])
    tb = Xw.mean(axis=1) * (1.0 + u01(s0[:, 1]))[:, None] + 1.5 * nt[:, :3]
    n = 3 + 3 * N + 6 * (M - 1)
    R = np.tile(np.eye(3), (B, M, 1, 1))
    t = np.zeros((B, M, 3))
    om = np.zeros((B, M - 1, 3))
    truth, x0 = np.zeros((B, n)), np.zeros((B, n))
    for v in range(1, M):
        a = normal4(draw(seed, gb, S_VIEWS, 3 * v))
        c = normal4(draw(seed, gb, S_VIEWS, 3 * v + 1))
        e = normal4(draw(seed, gb, S_VIEWS, 3 * v + 2))
        L = 3.0 * a[:, :3]
        tg = tb + 3.0 * np.stack([a[:, 3], c[:, 0], c[:, 1]], axis=1)
        ub = np.stack([3.0 * c[:, 2], -up_distance + 3.0 * c[:, 3], 3.0 * e[:, 0]], axis=1)
        f = tg - L
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        u = ub - L
        u -= f * (f * u).sum(1, keepdims=True)
        y = -u / np.linalg.norm(u, axis=1, keepdims=True)
        x = np.cross(y, f)
        z = ((Xw - L[:, None]) * f[:, None]).sum(-1) - min_camera_distance
        zmin = z.min(axis=1)
        L = L + np.where(zmin < 1e-3, zmin, 0.0)[:, None] * f
        Rv = np.stack([x, y, f], axis=1)
        R[:, v] = Rv
        t[:, v] = -np.einsum("bij,bj->bi", Rv, L)
        w = np.stack([Rv[:, 2, 1] - Rv[:, 1, 2], Rv[:, 0, 2] - Rv[:, 2, 0], Rv[:, 1, 0] - Rv[:, 0, 1]], axis=1)
        s2 = np.linalg.norm(w, axis=1)
        ang = np.arctan2(0.5 * s2, 0.5 * (np.trace(Rv, axis1=1, axis2=2) - 1.0))
        k = np.where(s2 > 1e-12, ang / np.where(s2 > 1e-12, s2, 1.0), 0.5)
        om[:, v - 1] = k[:, None] * w
        sn = normal4(draw(seed, gb, S_START, 2 * v))
        sr = normal4(draw(seed, gb, S_START, 2 * v + 1))
        pt = 3 + 3 * N + 3 * (v - 1)
        pr = pt + 3 * (M - 1)
        truth[:, pt:pt + 3], truth[:, pr:pr + 3] = t[:, v], om[:, v - 1]
        x0[:, pt:pt + 3] = t[:, v] + start_noise * 0.3 * sn[:, :3]
        x0[:, pr:pr + 3] = om[:, v - 1] + start_noise * 0.03 * sr[:, :3]
    fpar = np.where(fp > 1.0, fp - 1.0, np.log(fp))
    sn0 = normal4(draw(seed, gb, S_START, 0))
    tv = np.stack([fpar, pc[:, 0], pc[:, 1]], axis=1)
    truth[:, :3] = tv
    x0[:, :3] = tv + start_noise * np.array([0.1, 0.05, 0.05]) * sn0[:, :3]
    snp = normal4(draw(seed, gb[:, None], S_START, (64 + np.arange(N))[None, :]))
    truth[:, 3:3 + 3 * N] = Xw.reshape(B, -1)
    x0[:, 3:3 + 3 * N] = (Xw + start_noise * 0.3 * snp[..., :3]).reshape(B, -1)
    rel = np.einsum("bvij,bnj->bvni", R, Xw) + t[:, :, None, :]
    zc = np.maximum(rel[..., 2], 1e-8)
    proj = fp[:, None, None, None] * rel[..., :2] / zc[..., None] + pc[:, None, None, :]
    vis = ((proj > -1.0) & (proj < 1.0)).all(-1) & (rel[..., 2] > 0.0)
    return dict(projected_points=proj, visibility_mask=vis.astype(np.float64), camera_intrinsics=np.stack([fp, pc[:, 0], pc[:, 1]], 1),
                camera_orientations=om, camera_translations=t[:, 1:], world_points=Xw, x0=x0, truth=truth)

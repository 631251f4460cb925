"""Synthetic organised LiDAR scans (SURVEY.md §8d generator).

Scene: box room x∈[-15,15], y∈[-10,10], floor z=-1.5, ceiling z=+4, plus 24 vertical
cylinders (radius 0.3 m) centred on a circle of radius 6 m.  Sensor: R rings with
elevation linearly spaced in [-fov/2, +fov/2], P columns with azimuth 2πc/P; point index
r*P + c (row-major, the layout the reference requires, README.md:92 of the reference).
Range noise N(0, sigma); xyz rounded to float32 and stored {x, y, z, 0} (16-byte point).

The generator is written against a tiny array-namespace shim so the same code runs on
numpy (tests, fixtures) and on torch (bench: bulk generation on the GPU).
"""
from __future__ import annotations

import math

import numpy as np

ROOM = (-15.0, 15.0, -10.0, 10.0, -1.5, 4.0)
N_CYL = 24
CYL_RADIUS = 0.3
CYL_RING = 6.0


def fov_for_rings(rings: int) -> float:
    """Vertical field of view in degrees for the named sensor shapes."""
    if rings <= 16:
        return 30.0
    if rings >= 128:
        return 45.0
    return 32.0


def trajectory(k):
    """Sensor pose of scan k: (x, y, z, yaw).  ≈6–9 cm and ≈0.8° per step."""
    x = 3.0 * math.sin(2 * math.pi * k / 300.0)
    y = 2.0 * math.sin(2 * math.pi * k / 190.0)
    yaw = 0.3 * math.sin(2 * math.pi * k / 140.0)
    return x, y, 0.0, yaw


def pose_of_scan(k):
    """world_T_sensor(k) as (qx,qy,qz,qw,tx,ty,tz)."""
    x, y, z, yaw = trajectory(k)
    return np.array([0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2), x, y, z])


def relative_pose(k_target: int, k_source: int):
    """Ground-truth target_T_source = world_T_target^-1 * world_T_source (yaw-only motion)."""
    xt, yt, _, yawt = trajectory(k_target)
    xs, ys, _, yaws = trajectory(k_source)
    dyaw = yaws - yawt
    c, s = math.cos(-yawt), math.sin(-yawt)
    dx, dy = xs - xt, ys - yt
    return np.array([0.0, 0.0, math.sin(dyaw / 2), math.cos(dyaw / 2), c * dx - s * dy, s * dx + c * dy, 0.0])


def _ranges_numpy(rings, cols, pos, yaw, fov_deg):
    el = np.deg2rad(np.linspace(-fov_deg / 2, fov_deg / 2, rings))[:, None]
    az = (2 * np.pi * np.arange(cols) / cols)[None, :]
    ds = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el) * np.ones_like(az)], axis=-1)
    cy, sy = math.cos(yaw), math.sin(yaw)
    dw = np.stack([cy * ds[..., 0] - sy * ds[..., 1], sy * ds[..., 0] + cy * ds[..., 1], ds[..., 2]], axis=-1)
    t = np.full(dw.shape[:2], np.inf)
    lo = (ROOM[0], ROOM[2], ROOM[4])
    hi = (ROOM[1], ROOM[3], ROOM[5])
    with np.errstate(divide="ignore", invalid="ignore"):
        for a in range(3):
            d = dw[..., a]
            ta = np.where(d > 0, (hi[a] - pos[a]) / d, np.where(d < 0, (lo[a] - pos[a]) / d, np.inf))
            t = np.minimum(t, ta)
        # vertical cylinders
        dxy2 = dw[..., 0] ** 2 + dw[..., 1] ** 2
        for k in range(N_CYL):
            ang = 2 * np.pi * k / N_CYL + 0.1
            cx, cyy = CYL_RING * math.cos(ang), CYL_RING * math.sin(ang)
            ox, oy = pos[0] - cx, pos[1] - cyy
            b = ox * dw[..., 0] + oy * dw[..., 1]
            c = ox * ox + oy * oy - CYL_RADIUS**2
            disc = b * b - dxy2 * c
            tc = (-b - np.sqrt(np.maximum(disc, 0))) / dxy2
            ok = (disc > 0) & (tc > 0)
            t = np.where(ok, np.minimum(t, tc), t)
    return t, ds


def make_scan(rings: int, cols: int, k: int = 0, sigma: float = 0.01, seed: int | None = None,
              dropout: float = 0.0, static: bool = False) -> np.ndarray:
    """Scan k of the synthetic sequence as float32 [rings*cols, 4] = {x, y, z, 0} in the sensor frame."""
    x, y, z, yaw = (0.0, 0.0, 0.0, 0.0) if static else trajectory(k)
    t, ds = _ranges_numpy(rings, cols, (x, y, z), yaw, fov_for_rings(rings))
    rng = np.random.RandomState(1000 + k if seed is None else seed)
    t = t + rng.normal(0.0, sigma, size=t.shape) if sigma > 0 else t
    pts = (ds * t[..., None]).astype(np.float32)
    if dropout > 0:
        drop = rng.uniform(size=t.shape) < dropout
        pts[drop] = 0.0
    out = np.zeros((rings * cols, 4), dtype=np.float32)
    out[:, :3] = pts.reshape(-1, 3)
    return out


def sweep_poses(cols: int, start_T_end):
    """Per-column sensor pose relative to the sweep start, the model loamgpu_extract_dewarped inverts: column c is
    measured at s = c / cols, rotation = normalised linear interpolation of the quaternion, translation linear.
    Returns [cols, 7] (qx,qy,qz,qw,tx,ty,tz)."""
    m = np.asarray(start_T_end, dtype=np.float64)
    q = m[:4] if m[3] >= 0 else -m[:4]
    s = (np.arange(cols, dtype=np.float64) / cols)[:, None]
    qs = s * q[None, :]
    qs[:, 3] = (1.0 - s[:, 0]) + s[:, 0] * q[3]
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    return np.concatenate([qs, s * m[None, 4:7]], axis=1)


def make_warped_scan(rings: int, cols: int, k: int, start_T_end, sigma: float = 0.0, seed: int | None = None,
                     start_pose=None):
    """Scan k as a sensor that keeps moving during the sweep would record it (yaw + planar motions only): column c is
    cast from world_T_sensor(k) ∘ sweep_poses[c] and stored in that column's own instantaneous frame.  float32
    [rings*cols, 4].  De-warping it with start_T_end gives the scene as seen from the pose at the sweep start."""
    m = np.asarray(start_T_end, dtype=np.float64)
    assert m[0] == 0.0 and m[1] == 0.0 and m[6] == 0.0, "generator handles yaw + planar motion only"
    x0, y0, yaw0 = start_pose if start_pose is not None else (trajectory(k)[0], trajectory(k)[1], trajectory(k)[3])
    sp = sweep_poses(cols, m)
    yaw = yaw0 + 2.0 * np.arctan2(sp[:, 2], sp[:, 3])  # [cols]
    c0, s0 = math.cos(yaw0), math.sin(yaw0)
    px = x0 + c0 * sp[:, 4] - s0 * sp[:, 5]
    py = y0 + s0 * sp[:, 4] + c0 * sp[:, 5]
    fov = fov_for_rings(rings)
    el = np.deg2rad(np.linspace(-fov / 2, fov / 2, rings))[:, None]
    az = (2 * np.pi * np.arange(cols) / cols)[None, :]
    ds = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el) * np.ones_like(az)], axis=-1)
    cy, sy = np.cos(yaw)[None, :], np.sin(yaw)[None, :]
    dw = np.stack([cy * ds[..., 0] - sy * ds[..., 1], sy * ds[..., 0] + cy * ds[..., 1], ds[..., 2]], axis=-1)
    pos = (px[None, :], py[None, :], np.zeros((1, cols)))
    t = np.full(dw.shape[:2], np.inf)
    lo = (ROOM[0], ROOM[2], ROOM[4])
    hi = (ROOM[1], ROOM[3], ROOM[5])
    with np.errstate(divide="ignore", invalid="ignore"):
        for a in range(3):
            d = dw[..., a]
            ta = np.where(d > 0, (hi[a] - pos[a]) / d, np.where(d < 0, (lo[a] - pos[a]) / d, np.inf))
            t = np.minimum(t, ta)
        dxy2 = dw[..., 0] ** 2 + dw[..., 1] ** 2
        for j in range(N_CYL):
            ang = 2 * np.pi * j / N_CYL + 0.1
            ox, oy = pos[0] - CYL_RING * math.cos(ang), pos[1] - CYL_RING * math.sin(ang)
            b = ox * dw[..., 0] + oy * dw[..., 1]
            c = ox * ox + oy * oy - CYL_RADIUS**2
            disc = b * b - dxy2 * c
            tc = (-b - np.sqrt(np.maximum(disc, 0))) / dxy2
            ok = (disc > 0) & (tc > 0)
            t = np.where(ok, np.minimum(t, tc), t)
    if sigma > 0:
        rng = np.random.RandomState(1000 + k if seed is None else seed)
        t = t + rng.normal(0.0, sigma, size=t.shape)
    out = np.zeros((rings * cols, 4), dtype=np.float32)
    out[:, :3] = (ds * t[..., None]).astype(np.float32).reshape(-1, 3)
    return out


def distance_to_scene(world_pts):
    """Distance of world-frame points to the nearest scene surface (room faces, cylinders) — what a correctly
    de-warped noise-free scan drives to ~0."""
    p = np.asarray(world_pts, dtype=np.float64)
    d = np.minimum.reduce([np.abs(p[:, 0] - ROOM[0]), np.abs(p[:, 0] - ROOM[1]), np.abs(p[:, 1] - ROOM[2]),
                           np.abs(p[:, 1] - ROOM[3]), np.abs(p[:, 2] - ROOM[4]), np.abs(p[:, 2] - ROOM[5])])
    for j in range(N_CYL):
        ang = 2 * np.pi * j / N_CYL + 0.1
        r = np.hypot(p[:, 0] - CYL_RING * math.cos(ang), p[:, 1] - CYL_RING * math.sin(ang))
        d = np.minimum(d, np.abs(r - CYL_RADIUS))
    return d


def make_scans_torch(rings: int, cols: int, k0: int, count: int, device, sigma: float = 0.01, chunk: int = 64):
    """Bulk generation on `device` with torch: float32 [count, rings*cols, 4].

    Same scene/trajectory as make_scan (noise stream differs: torch generator seeded 1000+k0).
    Used by bench.py to create the 10,000-scan sequence without a slow host loop.
    """
    import torch

    fov = fov_for_rings(rings)
    f64 = torch.float64
    el = torch.deg2rad(torch.linspace(-fov / 2, fov / 2, rings, dtype=f64, device=device))[:, None]
    az = (2 * math.pi * torch.arange(cols, dtype=f64, device=device) / cols)[None, :]
    ds = torch.stack([torch.cos(el) * torch.cos(az), torch.cos(el) * torch.sin(az),
                      torch.sin(el) * torch.ones_like(az)], dim=-1)  # [R,P,3]
    gen = torch.Generator(device=device)
    gen.manual_seed(1000 + k0)
    out = torch.zeros((count, rings * cols, 4), dtype=torch.float32, device=device)
    angs = 2 * math.pi * torch.arange(N_CYL, dtype=f64, device=device) / N_CYL + 0.1
    ccx, ccy = CYL_RING * torch.cos(angs), CYL_RING * torch.sin(angs)
    inf = float("inf")
    for c0 in range(0, count, chunk):
        n = min(chunk, count - c0)
        ks = torch.arange(k0 + c0, k0 + c0 + n, dtype=f64, device=device)
        px = 3.0 * torch.sin(2 * math.pi * ks / 300.0)
        py = 2.0 * torch.sin(2 * math.pi * ks / 190.0)
        yaw = 0.3 * torch.sin(2 * math.pi * ks / 140.0)
        cy, sy = torch.cos(yaw)[:, None, None], torch.sin(yaw)[:, None, None]
        dwx = cy * ds[None, ..., 0] - sy * ds[None, ..., 1]
        dwy = sy * ds[None, ..., 0] + cy * ds[None, ..., 1]
        dwz = ds[None, ..., 2].expand(n, -1, -1)
        pxb, pyb = px[:, None, None], py[:, None, None]
        t = torch.full_like(dwx, inf)
        for d, lo, hi, p in ((dwx, ROOM[0], ROOM[1], pxb), (dwy, ROOM[2], ROOM[3], pyb), (dwz, ROOM[4], ROOM[5], 0.0)):
            ta = torch.where(d > 0, (hi - p) / d, torch.where(d < 0, (lo - p) / d, torch.full_like(d, inf)))
            t = torch.minimum(t, ta)
        dxy2 = dwx * dwx + dwy * dwy
        for j in range(N_CYL):
            ox, oy = pxb - ccx[j], pyb - ccy[j]
            b = ox * dwx + oy * dwy
            c = ox * ox + oy * oy - CYL_RADIUS**2
            disc = b * b - dxy2 * c
            tc = (-b - torch.sqrt(torch.clamp(disc, min=0))) / dxy2
            ok = (disc > 0) & (tc > 0)
            t = torch.where(ok, torch.minimum(t, tc), t)
        if sigma > 0:
            t = t + sigma * torch.randn(t.shape, dtype=f64, device=device, generator=gen)
        pts = (ds[None] * t[..., None]).to(torch.float32)
        out[c0:c0 + n, :, :3] = pts.reshape(n, rings * cols, 3)
    return out

"""Synthetic scenes of BASELINE.md section 4: analytic fields, the TSD default colour map and the orbit camera.

All generators are deterministic closed-form functions (no RNG, no files).  ``numpy`` variants feed
the parity tests (the oracle and the CUDA path get the SAME host array); ``torch`` variants create
the large bench volumes directly in HBM.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


# ----------------------------------------------------------------------------------------------------
# fields
# ----------------------------------------------------------------------------------------------------
def marschner_lobb_np(n: int, alpha: float = 0.25, f_m: float = 6.0) -> np.ndarray:
    """Marschner-Lobb test signal on [-1,1]^3, n^3 float32, x fastest."""
    c = np.linspace(-1.0, 1.0, n, dtype=np.float64)
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    r = np.sqrt(x * x + y * y)
    rho_r = np.cos(2.0 * np.pi * f_m * np.cos(np.pi * r / 2.0))
    v = (1.0 - np.sin(np.pi * z / 2.0) + alpha * (1.0 + rho_r)) / (2.0 * (1.0 + alpha))
    return np.ascontiguousarray(v.astype(np.float32))


def marschner_lobb_torch(n: int, device, alpha: float = 0.25, f_m: float = 6.0, z_begin: int = 0,
                         z_end: int | None = None, nz_total: int | None = None):
    """Same signal generated slab-wise on the GPU (float32 math); returns a (z,y,x) float32 tensor."""
    import torch
    nz_total = n if nz_total is None else nz_total
    z_end = nz_total if z_end is None else z_end
    cx = torch.linspace(-1.0, 1.0, n, device=device, dtype=torch.float32)
    cz_all = torch.linspace(-1.0, 1.0, nz_total, device=device, dtype=torch.float32)
    r = torch.sqrt(cx[None, :] ** 2 + cx[:, None] ** 2)  # (y,x)
    rho_r = torch.cos(2.0 * math.pi * f_m * torch.cos(math.pi * r / 2.0))
    plane = alpha * (1.0 + rho_r)  # (y,x)
    out = torch.empty((z_end - z_begin, n, n), device=device, dtype=torch.float32)
    chunk = 32
    for z0 in range(z_begin, z_end, chunk):
        z1 = min(z0 + chunk, z_end)
        zz = cz_all[z0:z1]
        out[z0 - z_begin:z1 - z_begin] = ((1.0 - torch.sin(math.pi * zz / 2.0))[:, None, None] + plane[None]) / (
            2.0 * (1.0 + alpha))
    return out


def blobs_np(n: int, sigma_frac: float = 1.0 / 16.0) -> np.ndarray:
    """Sum of 8 Gaussians centred on the (1/4,3/4)^3 lattice, sigma = n*sigma_frac voxels (C2 "blobs")."""
    c = (np.arange(n, dtype=np.float64) + 0.0) / (n - 1)
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    v = np.zeros_like(x)
    s2 = (sigma_frac) ** 2
    for cz in (0.25, 0.75):
        for cy in (0.25, 0.75):
            for cx in (0.25, 0.75):
                v += np.exp(-((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) / (2.0 * s2))
    return np.ascontiguousarray(np.clip(v, 0.0, 1.0).astype(np.float32))


def shells_np(n: int, n_shells: int = 12, radius_frac: float = 96.0 / 2048.0, width_frac: float = 8.0 / 2048.0,
              dtype=np.float32) -> np.ndarray:
    """Sparse field: thin Gaussian shells, zero elsewhere (C3: ~4 % non-empty macrocells)."""
    c = np.arange(n, dtype=np.float64) / (n - 1)
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    v = np.zeros_like(x)
    for k in range(n_shells):
        # deterministic low-discrepancy centres
        cx = (0.15 + 0.7 * ((k * 0.6180339887498949) % 1.0))
        cy = (0.15 + 0.7 * ((k * 0.7548776662466927) % 1.0))
        cz = (0.15 + 0.7 * ((k * 0.5698402909980532) % 1.0))
        d = np.sqrt((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2)
        s = np.exp(-((d - radius_frac) ** 2) / (2.0 * width_frac ** 2))
        s[np.abs(d - radius_frac) > 4.0 * width_frac] = 0.0
        v = np.maximum(v, s)
    if dtype == np.uint16:
        return np.ascontiguousarray(np.round(v * 65535.0).astype(np.uint16))
    if dtype == np.uint8:
        return np.ascontiguousarray(np.round(v * 255.0).astype(np.uint8))
    return np.ascontiguousarray(v.astype(np.float32))


def shells_torch(n: int, device, n_shells: int = 12, radius_frac: float = 96.0 / 2048.0,
                 width_frac: float = 8.0 / 2048.0, dtype="uint16"):
    """GPU generator of the sparse shells field, slab-wise; returns (z,y,x) tensor (uint16 stored as int16 bits)."""
    import torch
    c = torch.arange(n, device=device, dtype=torch.float32) / (n - 1)
    out = torch.zeros((n, n, n), device=device, dtype=torch.float32 if dtype == "float32" else torch.int16)
    chunk = 16
    yy = c[None, :, None]
    xx = c[None, None, :]
    for z0 in range(0, n, chunk):
        z1 = min(z0 + chunk, n)
        zz = c[z0:z1][:, None, None]
        v = torch.zeros((z1 - z0, n, n), device=device, dtype=torch.float32)
        for k in range(n_shells):
            cx = (0.15 + 0.7 * ((k * 0.6180339887498949) % 1.0))
            cy = (0.15 + 0.7 * ((k * 0.7548776662466927) % 1.0))
            cz = (0.15 + 0.7 * ((k * 0.5698402909980532) % 1.0))
            if abs(float(c[z0]) - cz) > radius_frac + 5 * width_frac and abs(float(c[z1 - 1]) - cz) > radius_frac + 5 * width_frac \
                    and not (float(c[z0]) < cz < float(c[z1 - 1])):
                continue
            d = torch.sqrt((xx - cx) ** 2 + (yy - cy) ** 2 + (zz - cz) ** 2)
            s = torch.exp(-((d - radius_frac) ** 2) / (2.0 * width_frac ** 2))
            s = torch.where((d - radius_frac).abs() > 4.0 * width_frac, torch.zeros_like(s), s)
            v = torch.maximum(v, s)
        if dtype == "float32":
            out[z0:z1] = v
        else:
            q = torch.round(v * 65535.0).to(torch.int32)
            out[z0:z1] = torch.where(q > 32767, q - 65536, q).to(torch.int16)
    return out


# ----------------------------------------------------------------------------------------------------
# transfer functions
# ----------------------------------------------------------------------------------------------------
def tsd_default_colormap(size: int = 256) -> np.ndarray:
    """tsd::makeDefaultColorMap (tsd/src/tsd/core/ColorMapUtil.hpp:70-95): (1,0,0,0)->(0,1,0,.5)->(0,0,1,1)."""
    ctrl = np.array([[1, 0, 0, 0.0], [0, 1, 0, 0.5], [0, 0, 1, 1.0]], dtype=np.float32)
    out = np.empty((size, 4), dtype=np.float32)
    scale = np.float32(len(ctrl) - 1) / np.float32(size - 1)
    for i in range(size):
        x = np.float32(i) * scale
        idx = int(x)
        t = np.float32(x - np.float32(idx))
        if idx + 1 < len(ctrl):
            out[i] = (np.float32(1.0) - t) * ctrl[idx] + t * ctrl[idx + 1]
        else:
            out[i] = ctrl[idx]
    return out


def sparse_colormap(size: int = 256, threshold: float = 0.5) -> np.ndarray:
    """Default map with alpha forced to 0 below `threshold` (exercises macrocell skipping, C3)."""
    cm = tsd_default_colormap(size)
    pos = np.arange(size, dtype=np.float32) / np.float32(size - 1)
    a = np.clip((pos - threshold) / (1.0 - threshold), 0.0, 1.0).astype(np.float32)
    cm[:, 3] = a
    return cm


# ----------------------------------------------------------------------------------------------------
# camera
# ----------------------------------------------------------------------------------------------------
@dataclass
class OrbitPose:
    position: tuple
    direction: tuple
    up: tuple
    fovy: float
    aspect: float


def orbit_camera(bounds_lo, bounds_hi, width: int, height: int, az_deg: float = 30.0, el_deg: float = 20.0,
                 fovy_deg: float = 60.0, dist_scale: float = 2.0) -> OrbitPose:
    """Orbit pose at dist_scale*|diag| looking at the centre (tsd/apps/tools/tsdRender.cpp:183-200)."""
    lo = np.asarray(bounds_lo, dtype=np.float64)
    hi = np.asarray(bounds_hi, dtype=np.float64)
    center = 0.5 * (lo + hi)
    dist = dist_scale * float(np.linalg.norm(hi - lo))
    az, el = math.radians(az_deg), math.radians(el_deg)
    offs = np.array([math.sin(az) * math.cos(el), math.sin(el), math.cos(az) * math.cos(el)]) * dist
    eye = center + offs
    d = center - eye
    d /= np.linalg.norm(d)
    return OrbitPose(tuple(float(v) for v in eye.astype(np.float32)), tuple(float(v) for v in d.astype(np.float32)),
                     (0.0, 1.0, 0.0), math.radians(fovy_deg), float(width) / float(height))

"""Synthetic point clouds for the five BASELINE.json configurations.

All generators are counter-based and use integer arithmetic only (torch int64 tensors), so the same
(seed, index) produces bit-identical coordinates on the CPU (oracle side) and on the GPU.
Coordinates are LAS-like: int32 X,Y,Z in millimetres, then ``p = offset + X * scale`` as a separate
multiply and add in float64 (reference io/LASFile.cpp:82-84, scale 0.001) — torch eager launches
one kernel per op, so no FMA contraction can happen.

This is benchmark/test input generation, not part of the tiler hot path.
"""
from __future__ import annotations

import math

import torch

M32 = 0xFFFFFFFF
SCALE = 0.001  # LAS scale: millimetres


def _mix(x):
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & M32
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & M32
    return x ^ (x >> 16)


def _rand_u32(i, stream, seed):
    """32-bit hash of (index tensor i, stream id, seed): values in [0, 2^32) as int64."""
    salt = _mix(torch.tensor((seed * 0x9E3779B1 + stream * 0x85EBCA6B) & M32, dtype=torch.int64, device=i.device))
    h = _mix((i & M32) ^ salt)
    return _mix((h + ((i >> 32) & M32) * 0x27D4EB2F + stream) & M32)


def _uniform_int(i, stream, seed, rng):
    """uniform integer in [0, rng), rng < 2^30"""
    return (_rand_u32(i, stream, seed) * int(rng)) >> 32


def _to_xyz(X, Y, Z, offset):
    out = torch.empty((X.numel(), 3), dtype=torch.float64, device=X.device)
    out[:, 0] = X.to(torch.float64) * SCALE + offset[0]
    out[:, 1] = Y.to(torch.float64) * SCALE + offset[1]
    out[:, 2] = Z.to(torch.float64) * SCALE + offset[2]
    return out


def _value_noise_mm(X, Y, seed, base_shift, octaves, amplitude_mm):
    """Multi-octave bilinear value noise on an integer lattice; returns int64 heights in mm."""
    total = torch.zeros_like(X)
    for o in range(octaves):
        sh = base_shift - o
        ix, iy = X >> sh, Y >> sh
        fx, fy = X & ((1 << sh) - 1), Y & ((1 << sh) - 1)
        one = 1 << sh

        def corner(cx, cy):
            return _rand_u32(cx * 1000003 + cy, 100 + o, seed) & 0xFFFF

        v00, v10 = corner(ix, iy), corner(ix + 1, iy)
        v01, v11 = corner(ix, iy + 1), corner(ix + 1, iy + 1)
        top = (v00 * (one - fx) + v10 * fx) >> sh
        bot = (v01 * (one - fx) + v11 * fx) >> sh
        val = (top * (one - fy) + bot * fy) >> sh
        total = total + (val >> o)
    # total < 2^17: scale to the requested amplitude
    return (total * int(amplitude_mm)) >> 17


def uniform_cube(n, seed=1, side_m=1000.0, device="cpu", start=0):
    """C1: uniform points in a cube."""
    i = torch.arange(start, start + n, dtype=torch.int64, device=device)
    r = int(round(side_m / SCALE))
    X, Y, Z = (_uniform_int(i, k, seed, r) for k in range(3))
    return _to_xyz(X, Y, Z, (0.0, 0.0, 0.0))


def terrain(n, seed=2, side_m=10000.0, amplitude_m=300.0, device="cpu", start=0):
    """C2 / C4: x,y uniform, z = value-noise heightfield + +-0.5 m jitter."""
    i = torch.arange(start, start + n, dtype=torch.int64, device=device)
    r = int(round(side_m / SCALE))
    X = _uniform_int(i, 0, seed, r)
    Y = _uniform_int(i, 1, seed, r)
    base_shift = max(8, int(math.floor(math.log2(r))) - 1)
    Z = _value_noise_mm(X, Y, seed, base_shift, 7, amplitude_m / SCALE)
    Z = Z + _uniform_int(i, 2, seed, 1001) - 500
    return _to_xyz(X, Y, Z, (400000.0, 5600000.0, 200.0))


def urban(n, seed=3, side_m=4000.0, height_m=200.0, n_primitives=2000, device="cpu", start=0):
    """C3: axis-aligned facades / roofs / ground patches with 2 cm Gaussian-like thickness."""
    i = torch.arange(start, start + n, dtype=torch.int64, device=device)
    r = int(round(side_m / SCALE))
    hmax = int(round(height_m / SCALE))
    k = torch.arange(n_primitives, dtype=torch.int64, device=device)
    cx = _uniform_int(k, 10, seed, r)
    cy = _uniform_int(k, 11, seed, r)
    w = 5000 + _uniform_int(k, 12, seed, 55000)        # 5..60 m
    d = 5000 + _uniform_int(k, 13, seed, 55000)
    hgt = 3000 + _uniform_int(k, 14, seed, hmax - 3000)
    kind = _uniform_int(k, 15, seed, 4)                 # 0 ground, 1 roof, 2 facade-x, 3 facade-y
    prim = _uniform_int(i, 0, seed, n_primitives)
    u = _rand_u32(i, 1, seed) & 0xFFFF
    v = _rand_u32(i, 2, seed) & 0xFFFF
    g = ((_rand_u32(i, 3, seed) & 0xFFFF) + (_rand_u32(i, 4, seed) & 0xFFFF) + (_rand_u32(i, 5, seed) & 0xFFFF) +
         (_rand_u32(i, 6, seed) & 0xFFFF) - 2 * 65535)
    thick = torch.div(g * 20, 37837, rounding_mode="trunc")  # ~N(0, 20 mm)
    pcx, pcy, pw, pd, ph, pk = cx[prim], cy[prim], w[prim], d[prim], hgt[prim], kind[prim]
    du = ((u * pw) >> 16) - (pw >> 1)
    dv = ((v * pd) >> 16) - (pd >> 1)
    hv = (v * ph) >> 16
    hu = (u * ph) >> 16
    zero = torch.zeros_like(u)
    X = pcx + torch.where(pk == 3, thick, du)
    Y = pcy + torch.where(pk == 2, thick, torch.where(pk == 3, du, dv))
    Z = torch.where(pk == 0, thick, torch.where(pk == 1, ph + thick, torch.where(pk == 2, hv, hu)))
    del zero
    X = X.clamp(0, r - 1)
    Y = Y.clamp(0, r - 1)
    Z = Z.clamp(-100, hmax)
    return _to_xyz(X, Y, Z, (30000.0, 60000.0, 100.0))


def skewed(n, seed=5, side_m=1000.0, device="cpu", start=0):
    """C5: 95 % of the points uniform in a sub-cube holding 1 % of the volume, 5 % elsewhere."""
    i = torch.arange(start, start + n, dtype=torch.int64, device=device)
    r = int(round(side_m / SCALE))
    sub = int(r * 0.2154)
    dense = _uniform_int(i, 7, seed, 100) < 95
    ox, oy, oz = int(r * 0.31), int(r * 0.55), int(r * 0.17)

    def coord(stream, off):
        full = _uniform_int(i, stream, seed, r)
        part = _uniform_int(i, stream + 3, seed, sub) + off
        return torch.where(dense, part, full)

    return _to_xyz(coord(0, ox), coord(1, oy), coord(2, oz), (0.0, 0.0, 0.0))


GENERATORS = {"uniform": uniform_cube, "terrain": terrain, "urban": urban, "skewed": skewed}


def generate(kind, n, seed, device="cpu", chunk=1 << 24, **kw):
    """Generates `n` points in chunks (bounded temporaries).  Returns a (n, 3) float64 tensor."""
    gen = GENERATORS[kind]
    out = torch.empty((n, 3), dtype=torch.float64, device=device)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        out[s:s + m] = gen(m, seed=seed, device=device, start=s, **kw)
    return out


def tight_bounds(xyz):
    mn = xyz.amin(dim=0).cpu().numpy()
    mx = xyz.amax(dim=0).cpu().numpy()
    return mn, mx


def shift_to_centre_float32(xyz, cubic_min, cubic_max):
    """The 3DTILES pre-transform: p -= cubic centre; p = (float)p (process/TilerProcess.cpp:552-559)."""
    centre = torch.tensor([cubic_min[a] + (cubic_max[a] - cubic_min[a]) / 2 for a in range(3)], dtype=torch.float64,
                          device=xyz.device)
    return (xyz - centre).to(torch.float32).to(torch.float64)

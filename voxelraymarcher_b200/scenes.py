"""Deterministic, integer-only procedural voxel scenes (synthetic stand-ins for the missing
``resources/scene.vox`` and the BASELINE.json configs).

Every generator returns ``(xyz, rgb)``: ``xyz`` int32 ``[n, 3]`` and ``rgb`` uint32 ``[n]`` (``r<<16 | g<<8 | b``), in
INSERTION ORDER -- the scene semantics are last-write-wins on duplicate coordinates
(reference: VoxelRaymarcher/src/geometry/VoxelSceneCPU.cuh:45), so order is part of the data.

The cube / sphere generators restate the reference's own (unused by its ``main``) generators:
VoxelRaymarcher/src/geometry/VoxelCube.cuh:10-39 and VoxelSphere.cuh:10-66, including their uint32 colour
arithmetic; tests/test_oracle.py checks them against the reference's own functions (run through the host build of the reference).
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32


def pack_rgb(r, g, b):
    return (np.asarray(r, U32) << U32(16)) | (np.asarray(g, U32) << U32(8)) | np.asarray(b, U32)


def _cat(parts):
    xyz = np.concatenate([p[0] for p in parts]).astype(np.int32, copy=False)
    rgb = np.concatenate([p[1] for p in parts]).astype(U32, copy=False)
    return np.ascontiguousarray(xyz), np.ascontiguousarray(rgb)


def hollow_cube(xp: int, yp: int, zp: int, hw: int):
    """VoxelCube::generateVoxelCube (VoxelCube.cuh:10-39): z faces red, then x faces green, then y faces blue."""
    rng = np.arange(-hw, hw, dtype=np.int32)
    a, b = np.meshgrid(rng, rng, indexing="ij")  # outer loop first
    a = a.ravel()
    b = b.ravel()
    n = a.size

    def inter(p, q):  # the reference inserts the +face and the -face alternately
        out = np.empty((2 * n, 3), np.int32)
        out[0::2] = p
        out[1::2] = q
        return out

    red, green, blue = pack_rgb(255, 0, 0), pack_rgb(0, 255, 0), pack_rgb(0, 0, 255)
    zf = inter(np.stack([xp + a, yp + b, np.full(n, zp + hw)], 1), np.stack([xp + a, yp + b, np.full(n, zp - hw)], 1))
    xf = inter(np.stack([np.full(n, xp - hw), yp + a, zp + b], 1), np.stack([np.full(n, xp + hw), yp + a, zp + b], 1))
    yf = inter(np.stack([xp + a, np.full(n, yp - hw), zp + b], 1), np.stack([xp + a, np.full(n, yp + hw), zp + b], 1))
    return _cat([(zf, np.full(2 * n, red, U32)), (xf, np.full(2 * n, green, U32)), (yf, np.full(2 * n, blue, U32))])


def sphere_shell(xp: int, yp: int, zp: int, radius: int, checkered: bool = False):
    """VoxelSphere::generateVoxelSphere / generateCheckeredVoxelSphere (VoxelSphere.cuh:10-66).

    All arithmetic is uint32 as in the reference (``green`` really is computed from ``y - xMin`` and the blue
    divisor really is ``xMax - zMin``).  Requires pos >= radius on every axis (otherwise the reference's
    unsigned loop bounds wrap and it generates nothing).
    """
    assert min(xp, yp, zp) >= radius > 1
    xmin, xmax = xp - radius, xp + radius
    ymin, ymax = yp - radius, yp + radius
    zmin, zmax = zp - radius, zp + radius
    xs = np.arange(xmin, xmax, 2 if checkered else 1, dtype=np.int64)
    ys = np.arange(ymin, ymax, dtype=np.int64)
    zs = np.arange(zmin, zmax, dtype=np.int64)
    x, y, z = np.meshgrid(xs, ys, zs, indexing="ij")
    x, y, z = x.ravel(), y.ravel(), z.ravel()
    d2 = (x - xp) ** 2 + (y - yp) ** 2 + (z - zp) ** 2
    keep = (d2 < radius * radius) & (d2 > (radius - 1) * (radius - 1))
    x, y, z = x[keep], y[keep], z[keep]
    m = np.int64(0xFFFFFFFF)
    kx = 200 // (xmax - xmin)
    ky = 200 // (ymax - ymin)
    kz = 200 // ((xmax - zmin) & m)
    red = (50 + (((x - xmin) & m) * kx & m)) & m
    green = (50 + (((y - xmin) & m) * ky & m)) & m
    blue = (50 + (((z - zmin) & m) * kz & m)) & m
    rgb = pack_rgb(np.minimum(red, 255), np.minimum(green, 255), np.minimum(blue, 255))
    return np.stack([x, y, z], 1).astype(np.int32), rgb


def checker_floor(x0: int, x1: int, z0: int, z1: int, y: int = 0, cell: int = 8):
    xs = np.arange(x0, x1, dtype=np.int32)
    zs = np.arange(z0, z1, dtype=np.int32)
    x, z = np.meshgrid(xs, zs, indexing="ij")
    x, z = x.ravel(), z.ravel()
    par = ((x // cell) + (z // cell)) & 1
    rgb = np.where(par == 0, pack_rgb(200, 200, 200), pack_rgb(60, 60, 90)).astype(U32)
    return np.stack([x, np.full_like(x, y), z], 1), rgb


def probe_scene():
    """Stand-in for the stripped ``resources/scene.vox``: sphere shell, hollow cube in the
    negative-z regions, checkered sphere shell, checker floor spanning negative regions.  Use with scale 8 and the
    reference camera (Main.cu:199)."""
    return _cat([
        sphere_shell(32, 32, 32, 20),
        hollow_cube(20, 10, -40, 12),
        sphere_shell(100, 40, 40, 30, checkered=True),
        checker_floor(-64, 128, -128, 64),
    ])


def mini_scene():
    """A few hundred voxels incl. duplicates with different colours (last write wins) and negative coordinates."""
    a = hollow_cube(3, 3, 3, 3)
    b = sphere_shell(12, 12, 12, 6)
    c = checker_floor(-20, 20, -20, 20, y=-2, cell=4)
    dup_xyz = np.array([[0, 0, 0], [0, 0, 0], [-1, -1, -1], [-1, -1, -1], [63, 63, 63], [64, 64, 64], [-64, 0, 0], [-65, 0, 0]], np.int32)
    dup_rgb = np.array([0x010203, 0x0A0B0C, 0x111111, 0x222222, 0x333333, 0x444444, 0x555555, 0x666666], U32)
    return _cat([a, b, c, (dup_xyz, dup_rgb)])


def _hash2(ix, iz, seed):
    h = (ix.astype(U32) * U32(0x9E3779B1)) ^ (iz.astype(U32) * U32(0x85EBCA77)) ^ U32((seed * 0xC2B2AE3D) & 0xFFFFFFFF)
    h ^= h >> U32(15)
    h *= U32(0x2C1B3C6D)
    h ^= h >> U32(12)
    h *= U32(0x297A2D39)
    h ^= h >> U32(15)
    return (h & U32(0xFFFF)).astype(np.int64)


def terrain_heights(size: int = 512, seed: int = 1234, base: int = 32, amps=(96, 48, 24, 12), lattices=(128, 64, 32, 16)):
    """Fixed-point 4-octave value noise height map h[x, z]; integer arithmetic only."""
    xs = np.arange(size, dtype=np.int64)
    x, z = np.meshgrid(xs, xs, indexing="ij")
    h = np.full((size, size), base << 16, np.int64)
    for octave, (amp, lat) in enumerate(zip(amps, lattices)):
        ix, iz = x // lat, z // lat
        fx, fz = (x % lat) * 256 // lat, (z % lat) * 256 // lat  # 8-bit fractions
        s = seed + octave
        v00, v10 = _hash2(ix, iz, s), _hash2(ix + 1, iz, s)
        v01, v11 = _hash2(ix, iz + 1, s), _hash2(ix + 1, iz + 1, s)
        top = v00 * (256 - fx) + v10 * fx
        bot = v01 * (256 - fx) + v11 * fx
        val = (top * (256 - fz) + bot * fz) >> 16  # back to 16-bit
        h += amp * val
    return (h >> 16).astype(np.int32)


def _height_colour(y):
    y = np.asarray(y)
    r = np.select([y < 40, y < 90, y < 140], [194, 60, 120], 240)
    g = np.select([y < 40, y < 90, y < 140], [178, 160, 120], 240)
    b = np.select([y < 40, y < 90, y < 140], [128, 70, 125], 250)
    shade = (y % 8) * 2
    return pack_rgb(r - shade, g - shade, b - shade)


def terrain(size: int = 512, seed: int = 1234, max_height: int | None = None):
    """Solid-filled height field in [0,size)^3 (BASELINE.json config 3: size 512 -> ~30 M voxels, 8^3 regions)."""
    h = terrain_heights(size, seed)
    h = np.clip(h, 1, (max_height or size) - 1)
    counts = h.ravel().astype(np.int64)
    total = int(counts.sum())
    col = np.repeat(np.arange(size * size, dtype=np.int64), counts)
    starts = np.cumsum(counts) - counts
    y = (np.arange(total, dtype=np.int64) - np.repeat(starts, counts)).astype(np.int32)
    xyz = np.empty((total, 3), np.int32)
    xyz[:, 0] = (col // size).astype(np.int32)
    xyz[:, 1] = y
    xyz[:, 2] = (col % size).astype(np.int32)
    return xyz, _height_colour(y).astype(U32)


def _hash3(ix, iy, iz, seed):
    h = (ix.astype(U32) * U32(0x9E3779B1)) ^ (iy.astype(U32) * U32(0x7FEB352D)) ^ (iz.astype(U32) * U32(0x85EBCA77)) ^ U32((seed * 0xC2B2AE3D) & 0xFFFFFFFF)
    h ^= h >> U32(16)
    h *= U32(0x21F0AAAD)
    h ^= h >> U32(15)
    h *= U32(0x735A2D97)
    h ^= h >> U32(15)
    return h


def menger_surface(size: int = 243, offset=(0, 0, 0)):
    """Surface voxels of a Menger sponge of edge ``size`` (a power of 3), coordinates offset by ``offset``
    (BASELINE.json config 4 family).  A cell is solid unless, at some base-3 digit position, at least two of its
    coordinates have digit 1.  Only solid cells with an empty (or outside) 6-neighbour are kept."""
    levels = 0
    s = 1
    while s < size:
        s *= 3
        levels += 1
    assert s == size, "size must be a power of 3"
    ax = np.arange(size, dtype=np.int32)

    def digit_is_one(level):
        return ((ax // (3 ** level)) % 3) == 1

    solid = np.ones((size, size, size), bool)
    for lv in range(levels):
        d = digit_is_one(lv)
        dx, dy, dz = d[:, None, None], d[None, :, None], d[None, None, :]
        hole = (dx & dy) | (dx & dz) | (dy & dz)
        solid &= ~hole
    pad = np.pad(solid, 1, constant_values=False)
    interior = (pad[:-2, 1:-1, 1:-1] & pad[2:, 1:-1, 1:-1] & pad[1:-1, :-2, 1:-1] & pad[1:-1, 2:, 1:-1]
                & pad[1:-1, 1:-1, :-2] & pad[1:-1, 1:-1, 2:])
    surf = solid & ~interior
    x, y, z = np.nonzero(surf)
    rgb = pack_rgb(80 + (x * 160 // size), 80 + (y * 160 // size), 80 + (z * 160 // size))
    xyz = np.stack([x, y, z], 1).astype(np.int32) + np.asarray(offset, np.int32)
    return xyz, rgb.astype(U32)


def sparse_shells(size: int = 1024, cell: int = 64, seed: int = 7, fill_pct: int = 35):
    """Sparse scene of sphere shells scattered on a coarse lattice inside [0,size)^3: roughly ``fill_pct`` percent
    of the ``cell``-sized lattice cells get one shell of a hashed radius (BASELINE.json configs 4/5 family;
    many regions, most clusters empty).  Integer-only."""
    n = size // cell
    g = np.arange(n, dtype=np.int64)
    cx, cy, cz = np.meshgrid(g, g, g, indexing="ij")
    cx, cy, cz = cx.ravel(), cy.ravel(), cz.ravel()
    h = _hash3(cx, cy, cz, seed)
    keep = (h % U32(100)) < U32(fill_pct)
    cx, cy, cz, h = cx[keep], cy[keep], cz[keep], h[keep]
    radii = (cell // 8 + ((h >> U32(8)) % U32(cell // 4)).astype(np.int64))
    parts = []
    cache = {}
    for r in np.unique(radii):
        r = int(r)
        ax = np.arange(-r, r + 1, dtype=np.int64)
        x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
        d2 = x * x + y * y + z * z
        m = (d2 < r * r) & (d2 >= (r - 1) * (r - 1))
        cache[r] = np.stack([x[m], y[m], z[m]], 1)
    for r in np.unique(radii):
        sel = radii == r
        centres = np.stack([cx[sel], cy[sel], cz[sel]], 1) * cell + cell // 2
        off = cache[int(r)]
        pts = (centres[:, None, :] + off[None, :, :]).reshape(-1, 3)
        hh = np.repeat(h[sel], off.shape[0])
        rgb = pack_rgb(64 + (hh & U32(127)), 64 + ((hh >> U32(7)) & U32(127)), 64 + ((hh >> U32(14)) & U32(127)))
        parts.append((pts.astype(np.int32), rgb))
    return _cat(parts)


def random_rays(n: int, origin, seed: int = 42):
    """Incoherent stress rays (BASELINE.json config 5): one ray per index from a fixed origin,
    direction from a counter-based integer hash -> three 24-bit uniforms in [-1,1)^3, rejected while
    len^2 > 1 or < 1e-4 (re-hash), normalised with IEEE fp32 divide / sqrt.  Returns float32 [n, 6]."""
    idx = np.arange(n, dtype=np.int64)
    d = np.zeros((n, 3), np.float32)
    todo = np.ones(n, bool)
    attempt = 0
    while todo.any():
        i = idx[todo]
        comps = []
        for c in range(3):
            h = _hash3(i, np.full_like(i, attempt), np.full_like(i, c), seed)
            u = (h >> U32(8)).astype(np.float32)  # 24 bits, exact in fp32
            comps.append(u * np.float32(2.0 ** -23) - np.float32(1.0))
        v = np.stack(comps, 1).astype(np.float32)
        l2 = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]).astype(np.float32) + v[:, 2] * v[:, 2]
        ok = (l2 <= np.float32(1.0)) & (l2 >= np.float32(1e-4))
        ln = np.sqrt(l2[ok], dtype=np.float32)
        where = np.nonzero(todo)[0][ok]
        d[where] = (v[ok] / ln[:, None]).astype(np.float32)
        todo[where] = False
        attempt += 1
    rays = np.empty((n, 6), np.float32)
    rays[:, 0:3] = np.asarray(origin, np.float32)
    rays[:, 3:6] = d
    return rays


def write_csv(path: str, xyz, rgb):
    """The reference's ``scene.vox`` text format: one ``x,y,z,color`` line per voxel (VoxelFile.cuh:9-35)."""
    with open(path, "w") as f:
        for (x, y, z), c in zip(np.asarray(xyz).tolist(), np.asarray(rgb).tolist()):
            f.write(f"{x},{y},{z},{c}\n")

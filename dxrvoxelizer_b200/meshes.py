"""Seeded synthetic watertight meshes for the build-dominated / batch configurations
(BASELINE.json configs 4 and 5): subdivided icospheres and
torus-knot tubes with a smooth seeded radial displacement, so Morton keys are not degenerate.

Pure numpy host code (input generation only -- nothing here is on the voxelization path).
"""
import numpy as np

from .voxelizer import Mesh


def _vertex_normals(pos, tri):
    """Unit face normals summed per vertex, then normalised (what ObjLoader::recomputeNormals does
    for files without `vn`, XUSGObjLoader.cpp:337-384; order of summation differs, so this is for
    synthetic inputs only, not a loader substitute)."""
    p0, p1, p2 = pos[tri[:, 0]], pos[tri[:, 1]], pos[tri[:, 2]]
    n = np.cross(p1 - p0, p2 - p1)
    l = np.linalg.norm(n, axis=1, keepdims=True)
    n = n / np.where(l > 0, l, 1)
    out = np.zeros_like(pos)
    for c in range(3):
        np.add.at(out, tri[:, c], n)
    l = np.linalg.norm(out, axis=1, keepdims=True)
    return (out / np.where(l > 0, l, 1)).astype(np.float32)


def _displace(unit_dirs, seed, amplitude=0.15, waves=4):
    """Smooth seeded radial displacement: a few random plane waves over the unit sphere."""
    rng = np.random.default_rng(seed)
    r = np.ones(len(unit_dirs), np.float64)
    for _ in range(waves):
        k = rng.normal(size=3)
        k *= rng.uniform(2.0, 6.0) / np.linalg.norm(k)
        r += amplitude / waves * np.sin(unit_dirs @ k + rng.uniform(0, 2 * np.pi))
    return r


def icosphere(subdivisions, seed=1234, amplitude=0.15, rotate=False, normals=True):
    """20 * 4**k triangles, 10 * 4**k + 2 vertices, outward winding, closed 2-manifold."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(subdivisions):
        # midpoint of every undirected edge, shared between the two triangles that use it
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        e.sort(axis=1)
        key = e[:, 0] * (len(v) + 1) + e[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // (len(v) + 1), uniq % (len(v) + 1)
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid], axis=0)
        n = len(f)
        m01, m12, m20 = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=0)
    pos = v * _displace(v, seed, amplitude)[:, None]
    if rotate:
        rng = np.random.default_rng(seed + 7919)
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        pos = pos @ q.T
    pos = pos.astype(np.float32)
    tri = f.astype(np.uint32)
    return Mesh.from_arrays(pos, tri, _vertex_normals(pos, tri) if normals else None)


def torus_knot(nu, nv, p=2, q=3, tube=0.18, seed=1234, amplitude=0.1, normals=True):
    """Tube around a (p,q) torus knot: 2*nu*nv triangles, nu*nv vertices, closed 2-manifold."""
    u = np.linspace(0, 2 * np.pi, nu, endpoint=False)
    r = np.cos(q * u) + 2.0
    c = np.stack([r * np.cos(p * u), r * np.sin(p * u), -np.sin(q * u)], 1)
    tang = np.roll(c, -1, 0) - np.roll(c, 1, 0)
    tang /= np.linalg.norm(tang, axis=1, keepdims=True)
    # rotation-minimising-enough frame: project a fixed axis
    ref = np.array([0.0, 0.0, 1.0])
    n1 = np.cross(tang, ref)
    n1 /= np.linalg.norm(n1, axis=1, keepdims=True)
    n2 = np.cross(tang, n1)
    w = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, 2)
    rad = tube * (1.0 + amplitude * np.sin(5 * u[:, None] + ph[0]) * np.cos(3 * w[None, :] + ph[1]))
    pos = c[:, None, :] + rad[..., None] * (np.cos(w)[None, :, None] * n1[:, None, :] + np.sin(w)[None, :, None] * n2[:, None, :])
    pos = pos.reshape(-1, 3).astype(np.float32)
    i = np.arange(nu)[:, None]
    j = np.arange(nv)[None, :]
    i1, j1 = (i + 1) % nu, (j + 1) % nv
    a, b, c2, d = i * nv + j, i1 * nv + j, i1 * nv + j1, i * nv + j1
    tri = np.concatenate([np.stack([a, c2, b], -1).reshape(-1, 3), np.stack([a, d, c2], -1).reshape(-1, 3)], 0).astype(np.uint32)
    return Mesh.from_arrays(pos, tri, _vertex_normals(pos, tri) if normals else None)


def cube(half=0.75):
    """Axis-aligned cube, 12 triangles, outward winding; exact analytic occupancy."""
    s = half
    pos = np.array([[-s, -s, -s], [s, -s, -s], [s, s, -s], [-s, s, -s], [-s, -s, s], [s, -s, s], [s, s, s], [-s, s, s]], np.float32)
    tri = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [2, 3, 7], [2, 7, 6],
                    [1, 2, 6], [1, 6, 5], [3, 0, 4], [3, 4, 7]], np.uint32)
    return Mesh.from_arrays(pos, tri, _vertex_normals(pos, tri))

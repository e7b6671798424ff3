"""LBVH build stages on the GPU: bounds, Morton sort (onesweep),
Karras hierarchy, atomic refit -- each checked against an independent numpy restatement."""
import numpy as np
import pytest

from dxrvoxelizer_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 31, 4096, 4097, 100000, 1 << 20, 3000001])
def test_onesweep_sorts_pairs_stably(vox, n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    if n > 1000:
        keys[: n // 3] &= np.uint32(0x3FF)          # many duplicates: stability matters
    vals = np.arange(n, dtype=np.uint32)
    k, v = vox.debug_sort_pairs(keys, vals)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])


def test_onesweep_all_equal_and_presorted(vox):
    n = 50000
    k, v = vox.debug_sort_pairs(np.full(n, 7, np.uint32), np.arange(n, dtype=np.uint32))
    assert np.array_equal(v, np.arange(n, dtype=np.uint32)) and (k == 7).all()
    keys = np.arange(n, dtype=np.uint32)[::-1].copy()
    k, v = vox.debug_sort_pairs(keys, keys)
    assert np.array_equal(k, np.arange(n, dtype=np.uint32)) and np.array_equal(v, k)


def _scene(mesh, bound):
    p = mesh.vertices[:, :3]
    b = bound.astype(np.float32)
    return ((p - b[:3]) / b[3]).astype(np.float32)


def _check_tree(vox, mesh):
    T = mesh.num_triangles
    bound = vox.bound()
    keys = vox.debug_read(L.DBG_MORTON_SORTED, np.uint32, T)
    prims = vox.debug_read(L.DBG_PRIM_SORTED, np.uint32, T)
    assert (np.diff(keys.astype(np.int64)) >= 0).all()
    assert np.array_equal(np.sort(prims), np.arange(T, dtype=np.uint32))
    # equal keys keep input order (stable sort)
    same = keys[1:] == keys[:-1]
    assert (prims[1:][same] > prims[:-1][same]).all()

    tris = vox.debug_read(L.DBG_TRIS, np.float32, T * 12).reshape(T, 3, 4)
    assert np.array_equal(tris[:, 0, 3].copy().view(np.uint32), prims)
    scene = _scene(mesh, bound)
    want = scene[mesh.indices.reshape(-1, 3)[prims]]
    assert np.array_equal(tris[:, :, :3], want)          # p' = (p - c) / w, bit exact

    if T < 2:
        return
    nodes = vox.debug_read(L.DBG_NODES, np.uint32, (T - 1) * 16).reshape(T - 1, 16)
    f = nodes[:, :12].copy().view(np.float32)            # yz0 | yz1 | x01 (csrc/common.cuh BvhNode)
    boxes = np.empty((T - 1, 2, 2, 3), np.float32)        # [node, child, lo/hi, xyz]
    for c in range(2):
        boxes[:, c, 0, 0], boxes[:, c, 1, 0] = f[:, 8 + 2 * c], f[:, 9 + 2 * c]
        boxes[:, c, 0, 1], boxes[:, c, 1, 1] = f[:, 4 * c + 0], f[:, 4 * c + 1]
        boxes[:, c, 0, 2], boxes[:, c, 1, 2] = f[:, 4 * c + 2], f[:, 4 * c + 3]
    child = nodes[:, 12:14]
    is_leaf = (child & 0x80000000) != 0
    idx = child & 0x7FFFFFFF
    # every leaf and every inner node (but the root) is referenced exactly once
    assert np.array_equal(np.sort(idx[is_leaf]), np.arange(T, dtype=np.uint32))
    assert np.array_equal(np.sort(idx[~is_leaf]), np.arange(1, T - 1, dtype=np.uint32))
    # boxes: bottom-up restatement (children are exact min/max unions)
    lo = np.empty((2 * T - 1, 3), np.float32)
    hi = np.empty((2 * T - 1, 3), np.float32)
    lo[T - 1:] = tris[:, :, :3].min(1)
    hi[T - 1:] = tris[:, :, :3].max(1)
    ref = np.where(is_leaf, idx + (T - 1), idx).astype(np.int64)
    # process inner nodes in an order where children come first (BFS from the root, then reversed);
    # every node also records the contiguous range of sorted leaves it covers
    rng = nodes[:, 14:16].astype(np.int64)
    depth = np.zeros(T - 1, np.int64)
    order = [0]
    assert tuple(rng[0]) == (0, T - 1)
    for i in order:
        split = [int(idx[i, 0]), int(idx[i, 1])]
        for c in range(2):
            if not is_leaf[i, c]:
                j = int(idx[i, c])
                depth[j] = depth[i] + 1
                order.append(j)
                want_rng = (rng[i, 0], split[0]) if c == 0 else (split[1], rng[i, 1])
                assert tuple(rng[j]) == tuple(want_rng)
        assert split[1] == split[0] + 1 and rng[i, 0] <= split[0] < rng[i, 1]
    assert len(order) == T - 1
    for i in reversed(order):
        lo[i] = np.minimum(lo[ref[i, 0]], lo[ref[i, 1]])
        hi[i] = np.maximum(hi[ref[i, 0]], hi[ref[i, 1]])
    assert np.array_equal(boxes[:, 0, 0], lo[ref[:, 0]]) and np.array_equal(boxes[:, 0, 1], hi[ref[:, 0]])
    assert np.array_equal(boxes[:, 1, 0], lo[ref[:, 1]]) and np.array_equal(boxes[:, 1, 1], hi[ref[:, 1]])
    root = vox.debug_read(L.DBG_ROOT_BOX, np.float32, 6)
    assert np.array_equal(root[:3], lo[0]) and np.array_equal(root[3:], hi[0])
    assert depth.max() <= 62


@pytest.mark.parametrize("name", ["dragon.obj", "TuringBowl.obj"])
def test_lbvh_invariants_on_shipped_meshes(vox, assets, name):
    m = assets(name)
    vox.build_bvh(m)
    assert np.array_equal(vox.bound(), m.bound)            # device bound == Voxelizer.cpp:52-57 on the host
    _check_tree(vox, m)
    vox.build_bvh(m)                                        # rebuild reuses the flag parity trick
    _check_tree(vox, m)


@pytest.mark.parametrize("shape", ["knot600", "knot16384", "ico81920"])
def test_lbvh_invariants_on_synthetic_meshes(vox, meshes_mod, shape):
    """Triangle counts that put 33..64 entries on the top level of the box pyramid (600 -> 38, 16384 -> 64: the
    two-entries-per-lane branch of the warp-wide range query) and one with four levels."""
    m = {"knot600": lambda: meshes_mod.torus_knot(30, 10, seed=1), "knot16384": lambda: meshes_mod.torus_knot(256, 32, seed=4),
         "ico81920": lambda: meshes_mod.icosphere(6, seed=2, rotate=True)}[shape]()
    vox.build_bvh(m)
    _check_tree(vox, m)


@pytest.mark.parametrize("tris", [1, 2, 3, 12])
def test_lbvh_tiny_meshes(vox, meshes_mod, tris):
    from dxrvoxelizer_b200 import Mesh
    c = meshes_mod.cube()
    m = Mesh(c.vertex_bytes, c.indices[: 3 * tris], c.stride)
    vox.build_bvh(m)
    _check_tree(vox, m)


def test_lbvh_duplicate_keys(vox):
    """Many triangles with identical Morton keys: the index tie-break must still give a tree."""
    from dxrvoxelizer_b200 import Mesh
    rng = np.random.default_rng(0)
    base = rng.uniform(-1e-4, 1e-4, size=(300, 3)).astype(np.float32)
    pos = np.concatenate([base, [[-1, -1, -1], [1, 1, 1]]]).astype(np.float32)
    tri = rng.integers(0, 300, size=(500, 3)).astype(np.uint32)
    m = Mesh.from_arrays(pos, tri)
    vox.build_bvh(m)
    _check_tree(vox, m)


def test_bad_index_is_reported(vox, meshes_mod):
    from dxrvoxelizer_b200 import Mesh, DxrvError
    c = meshes_mod.cube()
    idx = c.indices.copy()
    idx[5] = 1000
    vox.build_bvh(Mesh(c.vertex_bytes, idx, c.stride))
    with pytest.raises(DxrvError):
        vox.synchronize()
    vox.build_bvh(c)          # the context stays usable
    vox.synchronize()


def _build_snapshot(d, mesh, fused, tree, bound=None, N=64):
    """Sorted keys / triangle order / records of a fresh context; `tree`: the previous consumer traversed the hierarchy."""
    import os
    os.environ["DXRV_FUSED_BUILD"] = "1" if fused else "0"
    try:
        v = d.Voxelizer(0)
        v.build_bvh(mesh, bound)
        v.voxelize(N, d.MODE_SHADER if tree else d.MODE_PARITY)     # sets what the NEXT build includes
        v.build_bvh(mesh, bound)
        v.synchronize()
        T = mesh.num_triangles
        snap = (v.bound(), v.debug_read(L.DBG_MORTON_SORTED, np.uint32, T), v.debug_read(L.DBG_PRIM_SORTED, np.uint32, T),
                v.debug_read(L.DBG_TRIS, np.uint32, T * 12))
        v.voxelize(N, d.MODE_PARITY)
        grid_p = v.fetch_bits()
        v.voxelize(N, d.MODE_SHADER)                                # (fused, no tree: redoes the leaves for the pyramid)
        grid_s = v.fetch_bits()
        nodes = v.debug_read(L.DBG_NODES, np.uint32, (T - 1) * 16) if T > 1 else np.zeros(0, np.uint32)
        v.close()
        return snap + (grid_p, grid_s, nodes)
    finally:
        os.environ.pop("DXRV_FUSED_BUILD", None)


@pytest.mark.parametrize("shape", ["dragon", "bowl", "cube", "ico20480", "knot200k", "ico327680"])
@pytest.mark.parametrize("tree", [False, True])
def test_fused_build_equals_multi_kernel_build(assets, meshes_mod, shape, tree):
    """k_build_fused (one cooperative kernel: bounds, keys, stable radix passes, sorted records) against the
    multi-kernel build: same bound, keys, order, records, hierarchy and grids, bit for bit.  ico327680 is beyond
    the fused kernel's capacity (two triangles per thread of 148 CTAs): both runs take the multi-kernel path."""
    import dxrvoxelizer_b200 as d
    m = {"dragon": lambda: assets("dragon.obj"), "bowl": lambda: assets("TuringBowl.obj"), "cube": meshes_mod.cube,
         "ico20480": lambda: meshes_mod.icosphere(5, seed=3, rotate=True), "knot200k": lambda: meshes_mod.torus_knot(1000, 100, seed=5),
         "ico327680": lambda: meshes_mod.icosphere(7, seed=1)}[shape]()
    a = _build_snapshot(d, m, True, tree)
    b = _build_snapshot(d, m, False, tree)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_fused_build_with_given_bound(assets):
    import dxrvoxelizer_b200 as d
    m = assets("bunny.obj")
    bound = np.array([0.01, 0.1, 0.0, 0.13], np.float32)
    a = _build_snapshot(d, m, True, False, bound)
    b = _build_snapshot(d, m, False, False, bound)
    assert np.array_equal(a[0], bound)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_fused_builds_of_two_contexts_run_concurrently(assets):
    """Two contexts on one GPU, two host threads, each rebuilding and voxelizing a 100 k-triangle mesh back to back: the
    cooperative one-kernel builds of different streams must not starve each other's grid barriers (a barrier that gives
    up raises a device error, so a failure shows up here as an exception, not as a hang)."""
    import threading
    import dxrvoxelizer_b200 as d
    meshes = [assets("dragon.obj"), assets("bunny.obj")]
    want = []
    for m in meshes:
        v = d.Voxelizer(0)
        v.build_bvh(m); v.voxelize(640, d.MODE_PARITY); want.append(v.fetch_bits()); v.close()
    errors = []

    def work(k):
        try:
            v = d.Voxelizer(0)
            for _ in range(40):
                v.build_bvh(meshes[k])
                v.voxelize(640, d.MODE_PARITY)
            if not np.array_equal(v.fetch_bits(), want[k]):
                errors.append("grid of thread %d differs" % k)
            v.close()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in threads: t.start()
    for t in threads: t.join()
    assert not errors, errors

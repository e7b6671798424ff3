"""GPU parity tests proper: the CUDA path, called through
the C ABI, must equal the CPU oracle bit for bit -- zero mismatched voxels -- per mode, mesh, N, slab."""
import numpy as np
import pytest

import dxrvoxelizer_b200 as d
from conftest import popcount
from dxrvoxelizer_b200 import _lib as L

pytestmark = pytest.mark.gpu

SHIPPED = ["dragon.obj", "bunny.obj", "TuringBowl.obj"]


def _run(vox, mesh, N, mode, z0=0, z1=None, bound=None, texels=False):
    vox.build_bvh(mesh, bound=bound)
    vox.voxelize(N, mode, z0, z1, texels=texels)
    return vox.fetch_bits()


@pytest.mark.parametrize("name", SHIPPED)
@pytest.mark.parametrize("N", [64, 128])
def test_parity_mode_matches_oracle(vox, assets, oracle_mod, name, N):
    m = assets(name)
    got = _run(vox, m, N, d.MODE_PARITY)
    ref = oracle_mod.voxelize(m.vertices, m.indices, N, oracle_mod.MODE_PARITY)
    assert popcount(got ^ ref["bits"]) == 0
    assert vox.info(L.INFO_CROSSINGS) == ref["crossings"]
    assert vox.count_inside() == popcount(ref["bits"])


@pytest.mark.parametrize("name", SHIPPED)
def test_shader_mode_matches_oracle_at_reference_grid_size(vox, assets, oracle_mod, name):
    """C1: the reference's own configuration, GRID_SIZE 64 (Voxelizer.cpp:8), texels included."""
    m = assets(name)
    vox.build_bvh(m)
    vox.voxelize(64, d.MODE_SHADER, texels=True)
    ref = oracle_mod.voxelize(m.vertices, m.indices, 64, oracle_mod.MODE_SHADER, texels=True)
    assert popcount(vox.fetch_bits() ^ ref["bits"]) == 0
    assert np.array_equal(vox.fetch_texels(), ref["texels"])     # R10G10B10A2 exactly as the UAV holds it
    assert np.array_equal(vox.fetch_u8(), d.unpack_bits(ref["bits"], 64))


@pytest.mark.parametrize("N", [1, 5, 31, 33, 96, 100])
@pytest.mark.parametrize("mode", [d.MODE_SHADER, d.MODE_PARITY])
def test_ragged_grid_sizes(vox, meshes_mod, oracle_mod, N, mode):
    m = meshes_mod.icosphere(3, seed=11)
    got = _run(vox, m, N, mode)
    ref = oracle_mod.voxelize(m.vertices, m.indices, N, mode)
    assert got.shape == ref["bits"].shape
    assert popcount(got ^ ref["bits"]) == 0


@pytest.mark.parametrize("mode", [d.MODE_SHADER, d.MODE_PARITY])
def test_slabs_equal_single_shot(vox, assets, mode):
    """Multi-GPU decomposition run sequentially on one device."""
    m = assets("bunny.obj")
    N = 96
    full = _run(vox, m, N, mode)
    for k in (2, 3, 8):
        edges = [N * i // k for i in range(k + 1)]
        parts = []
        for a, b in zip(edges[:-1], edges[1:]):
            vox.voxelize(N, mode, a, b)
            parts.append(vox.fetch_bits())
        assert np.array_equal(np.concatenate(parts, 0), full)


def test_synthetic_meshes_both_modes(vox, meshes_mod, oracle_mod):
    for m in (meshes_mod.torus_knot(256, 32, seed=4), meshes_mod.icosphere(5, seed=9, rotate=True)):
        for mode in (d.MODE_SHADER, d.MODE_PARITY):
            got = _run(vox, m, 64, mode)
            assert popcount(got ^ oracle_mod.voxelize(m.vertices, m.indices, 64, mode)["bits"]) == 0


def test_exact_edge_and_vertex_hits(vox, meshes_mod, oracle_mod):
    """Columns through shared edges/vertices (exact zeros of the edge functions): double-precision
    fallback + symbolic tie rule must agree with the oracle and keep the fill watertight."""
    N = 16
    h = float((np.float32(11.5) / np.float32(16)) * np.float32(2) - np.float32(1))
    m = meshes_mod.cube(h)
    for mode in (d.MODE_SHADER, d.MODE_PARITY):
        got = _run(vox, m, N, mode, bound=[0, 0, 0, 1])
        ref = oracle_mod.voxelize(m.vertices, m.indices, N, mode, bound=[0, 0, 0, 1], tier=oracle_mod.TIER_BRUTE)
        assert popcount(got ^ ref["bits"]) == 0


def test_explicit_bound_and_positions_only_mesh(vox, meshes_mod, oracle_mod):
    c = meshes_mod.icosphere(2, seed=1)
    m = d.Mesh.from_arrays(c.vertices[:, :3], c.indices)          # stride 12: no normals
    got = _run(vox, m, 48, d.MODE_PARITY, bound=[0.1, -0.2, 0.05, 1.5])
    ref = oracle_mod.voxelize(m.vertices, m.indices, 48, 1, bound=[0.1, -0.2, 0.05, 1.5])
    assert popcount(got ^ ref["bits"]) == 0
    with pytest.raises(d.DxrvError):                               # MODE_SHADER needs normals
        vox.voxelize(48, d.MODE_SHADER)


def test_empty_mesh_gives_empty_grid(vox, meshes_mod):
    c = meshes_mod.cube()
    m = d.Mesh(c.vertex_bytes, np.zeros(0, np.uint32), c.stride)
    for mode in (d.MODE_SHADER, d.MODE_PARITY):
        assert popcount(_run(vox, m, 40, mode)) == 0


def test_error_behaviour(vox, meshes_mod):
    fresh = d.Voxelizer(0)
    with pytest.raises(d.DxrvError) as e:
        fresh.voxelize(64)
    assert e.value.code == L.ERR_NO_BVH
    fresh.build_bvh(meshes_mod.cube())
    with pytest.raises(d.DxrvError) as e:
        fresh.fetch_bits(np.empty((1, 1, 1), np.uint32))
    assert e.value.code == L.ERR_NO_GRID
    for bad in ((0, 1, 0, 0), (64, 9, 0, 64), (64, 1, 11, 10), (64, 1, 0, 65), (64, L.MODE_PARITY | L.EMIT_TEXELS, 0, 64)):
        with pytest.raises(d.DxrvError) as e:
            fresh._check(fresh._lib.dxrv_voxelize(fresh._h, *bad))
        assert e.value.code == L.ERR_INVALID_ARG
    fresh.voxelize(64)
    with pytest.raises(d.DxrvError):
        fresh.fetch_bits(np.empty((3,), np.uint32))               # wrong byte count
    fresh.close()


def test_external_grid_target_and_device_pointer(vox, assets):
    import torch
    m = assets("bunny.obj")
    N = 64
    full = _run(vox, m, N, d.MODE_PARITY)
    buf = torch.full((N * N * 2,), -1, dtype=torch.int32, device="cuda:0")
    torch.cuda.synchronize()
    half = N * N * 2 // 2 * 4
    vox.set_grid_target(buf.data_ptr(), half)
    vox.voxelize(N, d.MODE_PARITY, 0, N // 2)
    vox.set_grid_target(buf.data_ptr() + half, half)
    vox.voxelize(N, d.MODE_PARITY, N // 2, N)
    vox.synchronize()
    got = buf.cpu().numpy().view(np.uint32).reshape(N, N, 2)
    vox.set_grid_target(None, 0)
    assert np.array_equal(got, full)


@pytest.mark.parametrize("N,layers", [(128, 128), (256, 256), (1024, 40), (1280, 24), (2048, 24), (4096, 16)])
def test_every_word_of_the_slab_is_written(vox, assets, N, layers):
    """MODE_PARITY promises to write every word of the slab exactly once with no clear pass (empty tiles by
    the TMA writer CTAs, the rest by the tracing CTAs; 4, 8 or 16 warps per CTA by row length).  Run it
    into a buffer poisoned with ones and into one poisoned with zeros: a word nobody wrote would differ."""
    import torch
    m = assets("dragon.obj")
    vox.build_bvh(m)
    P = (N + 31) // 32
    z0 = (N - layers) // 2
    words = layers * N * P
    out = []
    for poison in (-1, 0):
        buf = torch.full((words,), poison, dtype=torch.int32, device="cuda:0")
        torch.cuda.synchronize()
        vox.set_grid_target(buf.data_ptr(), words * 4)
        vox.voxelize(N, d.MODE_PARITY, z0, z0 + layers)
        vox.synchronize()
        out.append(buf.cpu().numpy().view(np.uint32))
        vox.set_grid_target(None, 0)
    assert np.array_equal(out[0], out[1])
    assert 0 < popcount(out[0]) < words * 32


@pytest.mark.parametrize("name,N", [("dragon.obj", 256), ("TuringBowl.obj", 192), ("cube", 100), ("knot", 320)])
def test_scatter_and_tile_paths_agree(vox, assets, meshes_mod, oracle_mod, monkeypatch, name, N):
    """MODE_PARITY has two implementations (tile kernels / triangle-parallel scatter, chosen by triangle density):
    force each on the same input -- fine meshes, a coarse one (12 huge triangles) and a ragged N -- and
    require identical grids and crossing counts, equal to the oracle's."""
    m = {"cube": meshes_mod.cube, "knot": lambda: meshes_mod.torus_knot(192, 24, seed=3)}.get(name, lambda: assets(name))()
    ref = oracle_mod.voxelize(m.vertices, m.indices, N, oracle_mod.MODE_PARITY)
    for path in ("tiles", "scatter"):
        monkeypatch.setenv("DXRV_PARITY_PATH", path)
        got = _run(vox, m, N, d.MODE_PARITY)
        assert popcount(got ^ ref["bits"]) == 0, path
        assert vox.info(L.INFO_CROSSINGS) == ref["crossings"], path
        half = _run(vox, m, N, d.MODE_PARITY, N // 3, N // 3 + 17)        # a slab
        assert np.array_equal(half, got[N // 3:N // 3 + 17]), path


@pytest.mark.parametrize("seed,N", [(1, 37), (2, 64), (3, 160), (4, 256)])
def test_random_triangle_soup_both_paths(vox, oracle_mod, monkeypatch, seed, N):
    """Fuzz: an open soup of triangles of every size -- slivers, sub-voxel specks, sheets spanning the grid,
    axis-aligned ones whose edges run through column centres -- through both MODE_PARITY paths.  Nothing is
    watertight here, so columns have odd crossing counts; the toggling semantics (Spec H) still define every bit."""
    from dxrvoxelizer_b200 import Mesh
    rng = np.random.default_rng(seed)
    n = 400
    centre = rng.uniform(-1, 1, size=(n, 1, 3))
    size = 10.0 ** rng.uniform(-3.5, 0.3, size=(n, 1, 1))
    tri = centre + size * rng.uniform(-1, 1, size=(n, 3, 3))
    snap = rng.random(n) < 0.2                                   # some triangles on the voxel lattice: exact ties
    tri[snap] = np.round(tri[snap] * N / 2) * 2 / N
    pos = np.concatenate([tri.reshape(-1, 3), [[-1, -1, -1], [1, 1, 1]]]).astype(np.float32)   # pin the bound to the cube
    m = Mesh.from_arrays(pos, np.arange(3 * n, dtype=np.uint32).reshape(n, 3))
    ref = oracle_mod.voxelize(m.vertices, m.indices, N, oracle_mod.MODE_PARITY)
    for path in ("tiles", "scatter"):
        monkeypatch.setenv("DXRV_PARITY_PATH", path)
        got = _run(vox, m, N, d.MODE_PARITY)
        assert popcount(got ^ ref["bits"]) == 0, path
        assert vox.info(L.INFO_CROSSINGS) == ref["crossings"], path


def test_fine_mesh_with_huge_triangles(vox, assets, oracle_mod, monkeypatch):
    """A detailed object over a two-triangle ground sheet: the triangle density sends it down the scatter path,
    where the sheet's triangles cover a quarter of all columns each -- they are listed and rasterised by the whole
    grid (k_scatter_huge) instead of by one warp.  Same bits as the tile path and the oracle."""
    from dxrvoxelizer_b200 import Mesh
    m = assets("bunny.obj")
    pos = m.vertices[:, :3]
    lo, hi = pos.min(0), pos.max(0)
    y = lo[1] - 0.05 * (hi[1] - lo[1])
    e = 0.5 * (hi - lo).max()
    c = 0.5 * (lo + hi)
    sheet = np.array([[c[0] - e, y, c[2] - e], [c[0] + e, y, c[2] - e], [c[0] + e, y + 0.3 * e, c[2] + e], [c[0] - e, y + 0.3 * e, c[2] + e]], np.float32)
    n0 = pos.shape[0]
    tris = np.concatenate([m.indices.reshape(-1, 3), [[n0, n0 + 1, n0 + 2], [n0, n0 + 2, n0 + 3]]]).astype(np.uint32)
    mm = Mesh.from_arrays(np.concatenate([pos, sheet]), tris)
    N = 192
    ref = oracle_mod.voxelize(mm.vertices, mm.indices, N, oracle_mod.MODE_PARITY)
    for path in ("scatter", "tiles", "scatter"):
        monkeypatch.setenv("DXRV_PARITY_PATH", path)
        got = _run(vox, mm, N, d.MODE_PARITY)
        assert popcount(got ^ ref["bits"]) == 0, path
        assert vox.info(L.INFO_CROSSINGS) == ref["crossings"], path


def test_full_size_1024_parity_against_oracle(vox, assets, oracle_mod):
    """C3 at full size: dragon, N = 1024 (128 MiB bit grid).  The accelerated oracle finishes in
    seconds at this size, so the check is still a full bit-exact comparison, plus the size-independent
    properties: even crossing counts, inside fraction ~ mesh volume / 8, idempotence."""
    m = assets("dragon.obj")
    N = 1024
    got = _run(vox, m, N, d.MODE_PARITY)
    ref = oracle_mod.voxelize(m.vertices, m.indices, N, oracle_mod.MODE_PARITY)
    assert ref["odd_columns"] == 0
    assert popcount(got ^ ref["bits"]) == 0
    assert vox.info(L.INFO_CROSSINGS) == ref["crossings"]
    frac = vox.count_inside() / N ** 3
    assert abs(frac - 0.0556) < 0.001          # dragon volume / cube volume
    again = _run(vox, m, N, d.MODE_PARITY)
    assert np.array_equal(again, got)


def test_shader_mode_slab_at_512_matches_oracle(vox, assets, oracle_mod):
    """C2: TuringBowl at the 'hi-res' grid (N = 512); the oracle checks a few z-slabs at full width."""
    m = assets("TuringBowl.obj")
    N = 512
    vox.build_bvh(m)
    for z0 in (200, 255, 300):
        vox.voxelize(N, d.MODE_SHADER, z0, z0 + 2)
        ref = oracle_mod.voxelize(m.vertices, m.indices, N, oracle_mod.MODE_SHADER, z0=z0, z1=z0 + 2)
        assert popcount(vox.fetch_bits() ^ ref["bits"]) == 0


def test_kernel_timing_api(vox, assets):
    m = assets("bunny.obj")
    vox.build_bvh(m)
    with pytest.raises(d.DxrvError):
        vox.info(L.INFO_LAST_FILL_NS)          # profiling not enabled yet
    vox.set_profiling(True)
    vox.voxelize(256, d.MODE_PARITY)
    walk, fill = vox.info(L.INFO_LAST_WALK_NS), vox.info(L.INFO_LAST_FILL_NS)
    vox.set_profiling(False)
    assert 0 < walk < 50_000_000 and 0 < fill < 50_000_000


def test_independent_contexts_and_streams(meshes_mod, oracle_mod):
    """C5 shape: several contexts (one per stream) on one GPU, different meshes in flight at once."""
    ctxs = [d.Voxelizer(0) for _ in range(4)]
    ms = [meshes_mod.icosphere(4, seed=i, rotate=True) for i in range(8)]
    N = 64
    for round_ in range(2):
        for i, c in enumerate(ctxs):
            c.build_bvh(ms[round_ * 4 + i])
            c.voxelize(N, d.MODE_PARITY)
        for i, c in enumerate(ctxs):
            m = ms[round_ * 4 + i]
            assert popcount(c.fetch_bits() ^ oracle_mod.voxelize(m.vertices, m.indices, N, 1)["bits"]) == 0
    for c in ctxs:
        c.close()


def test_large_synthetic_mesh_dense_tiles(vox, meshes_mod, oracle_mod):
    """Build-dominated regime in miniature: 327k triangles at 128^3 puts thousands of candidates into
    every super-tile (candidate-list overflow -> in-kernel walk) and at 512^3 exercises split tiles."""
    m = meshes_mod.icosphere(7, seed=3, normals=False)
    for N in (128, 512):
        got = _run(vox, m, N, d.MODE_PARITY)
        ref = oracle_mod.voxelize(m.vertices, m.indices, N, 1)
        assert ref["odd_columns"] == 0
        assert popcount(got ^ ref["bits"]) == 0
        assert vox.info(L.INFO_CROSSINGS) == ref["crossings"]


def test_rebuild_is_deterministic(vox, assets):
    m = assets("TuringBowl.obj")
    a = _run(vox, m, 256, d.MODE_PARITY)
    k1 = vox.debug_read(L.DBG_PRIM_SORTED, np.uint32, m.num_triangles)
    b = _run(vox, m, 256, d.MODE_PARITY)
    k2 = vox.debug_read(L.DBG_PRIM_SORTED, np.uint32, m.num_triangles)
    assert np.array_equal(a, b) and np.array_equal(k1, k2)


@pytest.mark.parametrize("N,z0,z1", [(64, 0, 64), (96, 32, 64), (100, 0, 100), (256, 0, 256)])
def test_occupancy_pyramid(vox, assets, N, z0, z1):
    """dxrv_build_mips: a coarse voxel is set iff any of its 2x2x2 children is (numpy restatement)."""
    m = assets("bunny.obj")
    vox.build_bvh(m)
    vox.voxelize(N, d.MODE_PARITY, z0, z1)
    level0 = d.unpack_bits(vox.fetch_bits(), N).astype(bool)
    levels = vox.build_mips()
    want_levels, n, l = 1, N, z1 - z0
    while n % 2 == 0 and l % 2 == 0:
        n //= 2; l //= 2; want_levels += 1
    assert levels == want_levels
    cur = level0
    for lev in range(1, levels):
        lz, ly, lx = cur.shape
        cur = cur.reshape(lz // 2, 2, ly // 2, 2, lx // 2, 2).any(axis=(1, 3, 5))
        got = vox.fetch_mip(lev)
        assert got.shape[:2] == cur.shape[:2]
        assert np.array_equal(d.unpack_bits(got, cur.shape[2]).astype(bool), cur)
        pad = np.unpackbits(got.view(np.uint8), axis=-1, bitorder="little")[..., cur.shape[2]:]
        assert not pad.any()                       # bits beyond N_l stay clear
    with pytest.raises(d.DxrvError):
        vox.fetch_mip(levels)


def test_cli_matches_oracle(tmp_path, oracle_mod, assets):
    """The headless CLI (Bin/TuringBowl.bat argument list + new flags) end to end, raw grid dump."""
    import os, subprocess, json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "grid.bin"
    r = subprocess.run([os.path.join(root, "dxrvoxelizer_b200", "dxrvoxelizer"), "-mesh", d.asset_path("TuringBowl.obj"),
                        "0.0", "2.8", "0.0", "0.03", "-GRID", "128", "/mode", "parity", "-out", str(out)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout)
    m = assets("TuringBowl.obj")
    ref = oracle_mod.voxelize(m.vertices, m.indices, 128, 1)["bits"]
    got = np.fromfile(out, np.uint32).reshape(ref.shape)
    assert popcount(got ^ ref) == 0 and info["inside"] == popcount(ref) and info["triangles"] == m.num_triangles


@pytest.mark.parametrize("N", [384, 1280, 1664, 2048, 4096])
def test_weak_scaling_grid_sizes_slab(vox, assets, oracle_mod, N):
    """Grids of the multi-GPU bench (N % 128 == 0 but not a power of two: shared row pitch > global pitch,
    IEEE division in centre()) and the two larger CTA shapes (8 and 16 warps, N = 2048 / 4096 -- also the sizes
    at which the float index estimates of the conservative culling are least accurate): a few z-slabs
    against the oracle."""
    m = assets("dragon.obj")
    vox.build_bvh(m)
    for z0 in (N // 2 - 8, N // 2 + 40):
        vox.voxelize(N, d.MODE_PARITY, z0, z0 + 16)
        ref = oracle_mod.voxelize(m.vertices, m.indices, N, 1, z0=z0, z1=z0 + 16)
        assert popcount(vox.fetch_bits() ^ ref["bits"]) == 0
        assert vox.info(L.INFO_CROSSINGS) == ref["crossings"]


def test_degenerate_inputs_do_not_hang(vox, meshes_mod, oracle_mod):
    base = meshes_mod.icosphere(2, seed=8)
    # zero-area and duplicated triangles (finite coordinates): still bit-exact against the oracle
    idx = base.indices.reshape(-1, 3).copy()
    extra = np.array([[0, 0, 0], [1, 1, 2], [3, 4, 3], idx[5], idx[5]], np.uint32)
    m = d.Mesh(base.vertex_bytes, np.concatenate([idx, extra]).reshape(-1), base.stride)
    for mode in (d.MODE_PARITY, d.MODE_SHADER):
        got = _run(vox, m, 40, mode)
        assert popcount(got ^ oracle_mod.voxelize(m.vertices, m.indices, 40, mode)["bits"]) == 0
    # all vertices identical: w = 0, scene coordinates are NaN -> nothing can be hit, nothing may hang
    v = base.vertices.copy(); v[:, :3] = 1.5
    flat = d.Mesh(v, base.indices, base.stride)
    for mode in (d.MODE_PARITY, d.MODE_SHADER):
        assert popcount(_run(vox, flat, 32, mode)) == 0
    # a NaN vertex: the call must complete (results near the NaN triangles are unspecified)
    v = base.vertices.copy(); v[7, :3] = np.nan
    vox.build_bvh(d.Mesh(v, base.indices, base.stride))
    vox.voxelize(32, d.MODE_PARITY); vox.fetch_bits()
    vox.voxelize(32, d.MODE_SHADER); vox.fetch_bits()


def test_million_triangle_mesh_uses_atomic_refit(vox, meshes_mod, oracle_mod):
    """C4 regime: 1.3 M triangles (> 2^19: 30-bit keys, four radix passes with the big tile, bottom-up refit
    with atomics instead of range unions) -- still bit-exact, and the root box is the scene box."""
    m = meshes_mod.icosphere(8, seed=1234, normals=False)
    assert m.num_triangles == 1310720
    got = _run(vox, m, 256, d.MODE_PARITY)
    ref = oracle_mod.voxelize(m.vertices, m.indices, 256, 1)
    assert ref["odd_columns"] == 0
    assert popcount(got ^ ref["bits"]) == 0
    assert vox.info(L.INFO_CROSSINGS) == ref["crossings"]
    root = vox.debug_read(L.DBG_ROOT_BOX, np.float32, 6)
    b = vox.bound()
    scene = ((m.vertices[:, :3] - b[:3]) / b[3]).astype(np.float32)
    assert np.array_equal(root[:3], scene.min(0)) and np.array_equal(root[3:], scene.max(0))
    keys = vox.debug_read(L.DBG_MORTON_SORTED, np.uint32, m.num_triangles)
    assert (np.diff(keys.astype(np.int64)) >= 0).all()


@pytest.mark.parametrize("N,z0,z1,chunks", [(256, 0, 256, 8), (256, 40, 200, 3), (128, 0, 128, 1), (100, 10, 90, 4), (33, 0, 33, 8), (64, 5, 8, 16)])
def test_pipelined_voxelize_to_host(vox, assets, N, z0, z1, chunks):
    """dxrv_voxelize_to_host: the slab computed and read back in z sub-slabs equals voxelize + fetch, and leaves the
    context describing the whole slab (N = 100: 16-byte-aligned layers; N = 33: unaligned -> a single chunk)."""
    import torch
    m = assets("bunny.obj")
    vox.build_bvh(m)
    for mode in (d.MODE_PARITY, d.MODE_SHADER):
        vox.voxelize(N, mode, z0, z1)
        want = vox.fetch_bits()
        host = torch.empty(want.nbytes, dtype=torch.uint8).pin_memory()
        host.fill_(0xA5)
        vox.voxelize_to_host(N, mode, z0, z1, host.data_ptr(), want.nbytes, chunks)
        assert np.array_equal(host.numpy().view(np.uint32).reshape(want.shape), want)
        assert np.array_equal(vox.fetch_bits(), want)              # the context holds the whole slab afterwards
        assert vox.count_inside() == popcount(want)


@pytest.mark.parametrize("name,N,slab", [("dragon.obj", 512, None), ("TuringBowl.obj", 384, (100, 180)), ("cube", 200, None), ("bunny.obj", 1024, (500, 540))])
def test_walked_and_binned_candidates_agree(vox, assets, meshes_mod, oracle_mod, monkeypatch, name, N, slab):
    """The tile path of MODE_PARITY finds its candidates by triangle-parallel binning (default, no hierarchy) or by the
    LBVH walk (DXRV_PARITY_CANDIDATES=walk): same candidate sets, so identical grids and crossing counts, equal to the
    oracle's; 12 scene-sized triangles (cube) exercise the warp-cooperative rectangles."""
    m = meshes_mod.cube() if name == "cube" else assets(name)
    z0, z1 = slab if slab else (0, N)
    ref = oracle_mod.voxelize(m.vertices, m.indices, N, oracle_mod.MODE_PARITY, z0=z0, z1=z1)
    monkeypatch.setenv("DXRV_PARITY_PATH", "tiles")
    for how in ("bins", "walk", "bins"):
        monkeypatch.setenv("DXRV_PARITY_CANDIDATES", how)
        got = _run(vox, m, N, d.MODE_PARITY, z0, z1)
        assert popcount(got ^ ref["bits"]) == 0, how
        assert vox.info(L.INFO_CROSSINGS) == ref["crossings"], how


def test_binned_candidates_overflow_scans_all_triangles(vox, meshes_mod, oracle_mod, monkeypatch):
    """327 k triangles forced through the tile path at 64^3: 32 tiles, far more candidates per tile than a list holds
    -> the fill kernel's tree-free fallback (every CTA scans all triangle boxes) -- still the oracle's bits."""
    m = meshes_mod.icosphere(7, seed=3, normals=False)
    monkeypatch.setenv("DXRV_PARITY_PATH", "tiles")
    for how in ("bins", "walk"):
        monkeypatch.setenv("DXRV_PARITY_CANDIDATES", how)
        got = _run(vox, m, 64, d.MODE_PARITY)
        ref = oracle_mod.voxelize(m.vertices, m.indices, 64, 1)
        assert popcount(got ^ ref["bits"]) == 0, how
        assert vox.info(L.INFO_CROSSINGS) == ref["crossings"], how


def test_candidate_lists_are_reused_for_the_same_structure_grid_and_slab(assets):
    """MODE_PARITY tile path: a second voxelize of the same acceleration structure with the same grid and slab starts at
    k_file_columns (the candidate lists are still in the scratch buffer); a rebuild, another grid / slab or another
    mesh bins again.  Same grids every time."""
    import dxrvoxelizer_b200 as d
    from dxrvoxelizer_b200 import _lib as L
    v = d.Voxelizer(0)
    launches = lambda: v.info(L.INFO_KERNEL_LAUNCHES)
    for mesh in (assets("dragon.obj"), assets("bunny.obj")):
        v.build_bvh(mesh)
        for N, z0, z1 in ((640, 0, 640), (704, 13, 150)):        # (4 T < N^2: the tile path, not the scatter path)
            n0 = launches(); v.voxelize(N, d.MODE_PARITY, z0, z1); first = v.fetch_bits(); n1 = launches()
            v.voxelize(N, d.MODE_PARITY, z0, z1); second = v.fetch_bits(); n2 = launches()
            assert n1 - n0 == 3 and n2 - n1 == 2                   # k_bin_columns skipped the second time
            assert np.array_equal(first, second)
            v.voxelize(N, d.MODE_PARITY, z0, z1 - 1)               # another slab: bins again
            assert launches() - n2 == 3
            assert np.array_equal(v.fetch_bits(), first[:-1])
            v.voxelize(N, d.MODE_SHADER, z0, z0 + 2)               # (does not touch the lists)
            v.voxelize(N, d.MODE_PARITY, z0, z1 - 1)
            assert np.array_equal(v.fetch_bits(), first[:-1])
            v.build_bvh(mesh)                                      # a new structure: bins again
            n3 = launches(); v.voxelize(N, d.MODE_PARITY, z0, z1 - 1); n4 = launches()
            assert n4 - n3 == 3 and np.array_equal(v.fetch_bits(), first[:-1])
    v.close()

"""DXRV_FORMAT_SPARSE_BRICKS (include/dxrv.h): lossless compact read-back format.  The decoder is host code and is
checked here against an independent numpy ENCODER written from the format description; the GPU encoder is checked
against the decoder and against that numpy encoder (byte-identical blobs)."""
import numpy as np
import pytest

import dxrvoxelizer_b200 as d
from conftest import popcount


def numpy_encode(bits, N, z0):
    """uint32[(layers, N, P)] -> blob, straight from the description in dxrv.h."""
    layers, _, P = bits.shape
    BY, BZ = (N + 3) // 4, (layers + 3) // 4
    pad = np.zeros((BZ * 4, BY * 4, P), np.uint32)
    pad[:layers, :N] = bits
    exists = np.zeros((BZ * 4, BY * 4), bool)
    exists[:layers, :N] = True
    bricks = pad.reshape(BZ, 4, BY, 4, P).transpose(0, 2, 4, 1, 3).reshape(-1, 16)        # [b][4k + j]
    ex = np.broadcast_to(exists.reshape(BZ, 4, BY, 4)[:, :, :, :, None], (BZ, 4, BY, 4, P)).transpose(0, 2, 4, 1, 3).reshape(-1, 16)
    tail = np.uint32((1 << (N & 31)) - 1) if N & 31 else np.uint32(0xffffffff)
    full_word = np.where(np.arange(bricks.shape[0]) % P == P - 1, tail, np.uint32(0xffffffff)).astype(np.uint32)
    any_ = bricks.any(axis=1)
    full = ((bricks == full_word[:, None]) | ~ex).all(axis=1) & any_
    state = np.where(~any_, 0, np.where(full, 1, 2)).astype(np.uint32)
    nb = bricks.shape[0]
    sw = np.zeros((nb + 15) // 16, np.uint32)
    np.bitwise_or.at(sw, np.arange(nb) >> 4, state << (2 * (np.arange(nb) & 15)).astype(np.uint32))
    payload = bricks[state == 2].reshape(-1)
    off_states = 64
    off_payload = (off_states + sw.size * 4 + 63) & ~63
    header = np.array([0x42525844, 1, N, z0, z0 + layers, P, BY, BZ, nb, int((state == 2).sum()), off_states, off_payload, 32, 4, 4, 0], np.uint32)
    blob = np.zeros(off_payload + payload.size * 4, np.uint8)
    blob[:64] = header.view(np.uint8)
    blob[off_states:off_states + sw.size * 4] = sw.view(np.uint8)
    blob[off_payload:] = payload.view(np.uint8)
    return blob


@pytest.mark.parametrize("delay_us", [0, 3000])
@pytest.mark.parametrize("N,layers,seed", [(64, 64, 1), (100, 37, 2), (33, 5, 3), (128, 2, 4), (31, 31, 5), (256, 256, 6), (512, 42, 7)])
def test_decoder_against_a_numpy_encoder(N, layers, seed, delay_us, monkeypatch):
    # The decoder is ONE pass of the host pool over the dense grid (csrc/sparse_host.cpp, the same pass dxrv_voxelize_to_host
    # runs): brick layers reached before the blob is published are zeroed and expanded afterwards, the others are written
    # with their final contents at once.  delay_us = 3000 publishes the blob after the pool has zeroed everything.
    # (N = 512: rows of whole 16-word groups, the AVX-512 line composer where the host has it.)
    monkeypatch.setenv("DXRV_HOST_FILL_DELAY_US", str(delay_us))
    rng = np.random.default_rng(seed)
    P = (N + 31) // 32
    occ = np.zeros((layers, N, N), bool)
    zz, yy, xx = np.meshgrid(np.arange(layers), np.arange(N), np.arange(N), indexing="ij")
    occ |= (xx - N / 2) ** 2 + (yy - N / 2) ** 2 + (zz - layers / 2) ** 2 < (0.4 * N) ** 2          # a solid ball: empty, full, mixed bricks
    occ ^= rng.random(occ.shape) < 0.001                                                          # speckle
    bits = np.packbits(np.pad(occ, ((0, 0), (0, 0), (0, P * 32 - N))), axis=-1, bitorder="little").view(np.uint32).reshape(layers, N, P)
    blob = numpy_encode(bits, N, 7)
    assert np.array_equal(d.sparse_decode(blob), bits)
    bad = blob.copy(); bad[0] ^= 1
    with pytest.raises(d.DxrvError):
        d.sparse_decode(bad)
    with pytest.raises(d.DxrvError):
        d.sparse_decode(blob[:-8] if blob.size > 72 and blob[:64].view(np.uint32)[9] else blob[:60])
    bad = blob.copy(); bad[64] |= 3                                                               # brick 0 in the undefined state 3
    with pytest.raises(d.DxrvError):
        d.sparse_decode(bad)


@pytest.mark.parametrize("N,layers,z0,seed", [(64, 64, 0, 1), (100, 37, 7, 2), (33, 5, 20, 3), (128, 2, 126, 4), (31, 31, 0, 5), (256, 256, 0, 6),
                                              (512, 42, 300, 7), (1, 1, 0, 8), (5, 3, 2, 9)])
def test_host_encoder_against_the_numpy_encoder(N, layers, z0, seed):
    """dxrv_sparse_encode (host code): the blob of a dense slab, byte for byte what the format description gives (and so what
    the device encoder writes: test_gpu_encoder_round_trip); the decoder inverts it; an empty and a full grid; capacity."""
    import ctypes
    from dxrvoxelizer_b200 import _lib as L
    rng = np.random.default_rng(seed)
    P = (N + 31) // 32
    zz, yy, xx = np.meshgrid(np.arange(layers), np.arange(N), np.arange(N), indexing="ij")
    occ = (xx - N / 2) ** 2 + (yy - N / 2) ** 2 + (zz - layers / 2) ** 2 < (0.4 * N) ** 2
    occ ^= rng.random(occ.shape) < 0.001
    for grid in (occ, np.zeros_like(occ), np.ones_like(occ)):
        bits = np.packbits(np.pad(grid, ((0, 0), (0, 0), (0, P * 32 - N))), axis=-1, bitorder="little").view(np.uint32).reshape(layers, N, P)
        blob = d.sparse_encode(bits, N, z0)
        assert np.array_equal(blob, numpy_encode(bits, N, z0))
        assert np.array_equal(d.sparse_decode(blob), bits)
    n = ctypes.c_size_t()
    small = np.empty(blob.size - 1, np.uint8)
    lib = L.lib()
    assert lib.dxrv_sparse_encode(bits.ctypes.data, bits.nbytes, N, z0, z0 + layers, small.ctypes.data, small.size, ctypes.byref(n)) == L.ERR_INVALID_ARG
    assert n.value == blob.size                                                # the size it needs
    assert lib.dxrv_sparse_encode(bits.ctypes.data, bits.nbytes - 4, N, z0, z0 + layers, blob.ctypes.data, blob.size, ctypes.byref(n)) == L.ERR_INVALID_ARG
    assert lib.dxrv_sparse_encode(bits.ctypes.data, bits.nbytes, N, z0, N + 1, blob.ctypes.data, blob.size, ctypes.byref(n)) == L.ERR_INVALID_ARG


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,z0,z1,mode", [("dragon.obj", 256, 0, 256, 1), ("bunny.obj", 100, 10, 47, 1), ("TuringBowl.obj", 192, 0, 192, 0),
                                               ("dragon.obj", 1024, 300, 560, 1), ("bunny.obj", 33, 0, 33, 1)])
def test_gpu_encoder_round_trip(vox, assets, name, N, z0, z1, mode):
    m = assets(name)
    vox.build_bvh(m)
    vox.voxelize(N, mode, z0, z1)
    dense = vox.fetch_bits()
    blob = vox.fetch_sparse()
    assert np.array_equal(d.sparse_decode(blob), dense)
    assert np.array_equal(blob, numpy_encode(dense, N, z0))                 # byte-identical to the format description
    assert np.array_equal(blob, d.sparse_encode(dense, N, z0))              # ... and to the host encoder
    h = blob[:64].view(np.uint32)
    assert h[9] > 0 and blob.size < dense.nbytes                            # a solid object: far smaller than the dense grid
    small = np.empty(128, np.uint8)
    with pytest.raises(d.DxrvError):
        vox.fetch_sparse_into(small.ctypes.data, small.size)               # capacity too small


@pytest.mark.gpu
def test_empty_and_full_grids(vox, meshes_mod):
    c = meshes_mod.cube()
    vox.build_bvh(d.Mesh(c.vertex_bytes, np.zeros(0, np.uint32), c.stride))
    vox.voxelize(64, d.MODE_PARITY)
    blob = vox.fetch_sparse()
    assert blob[:64].view(np.uint32)[9] == 0 and popcount(d.sparse_decode(blob)) == 0
    vox.build_bvh(meshes_mod.cube(2.0), bound=[0, 0, 0, 1])                 # the cube contains the whole grid
    vox.voxelize(64, d.MODE_PARITY)
    dense = vox.fetch_bits()
    blob = vox.fetch_sparse()
    assert np.array_equal(d.sparse_decode(blob), dense)


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,z0,z1,mode", [("dragon.obj", 512, 0, 512, 1), ("bunny.obj", 1000, 100, 901, 1), ("TuringBowl.obj", 192, 0, 192, 0),
                                               ("dragon.obj", 1024, 0, 1024, 1), ("bunny.obj", 33, 0, 33, 1)])
def test_to_host_transports_agree(vox, assets, name, N, z0, z1, mode):
    """dxrv_voxelize_to_host: the pipelined dense copy, the sparse transport (device encoder + host expansion into a
    buffer full of garbage) and the automatic choice all leave the same dense grid as dxrv_voxelize + dxrv_fetch_grid."""
    from dxrvoxelizer_b200 import _lib as L
    m = assets(name)
    vox.build_bvh(m)
    vox.voxelize(N, mode, z0, z1)
    want = vox.fetch_bits()
    try:
        for transport in (L.READ_BACK_DENSE, L.READ_BACK_SPARSE, L.READ_BACK_AUTO):
            vox.set_read_back(transport)
            got = np.full(want.shape, 0xDEADBEEF, np.uint32)
            vox.voxelize_to_host(N, mode, z0, z1, got.ctypes.data, got.nbytes, chunks=4)
            assert np.array_equal(got, want), transport
            assert np.array_equal(vox.fetch_bits(), want)       # the context describes the whole slab afterwards
    finally:
        vox.set_read_back(L.READ_BACK_AUTO)


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,z0,z1,mode", [("dragon.obj", 1024, 0, 1024, 1), ("bunny.obj", 512, 37, 500, 1), ("TuringBowl.obj", 192, 0, 192, 0),
                                               ("bunny.obj", 100, 10, 47, 1)])
def test_mesh_to_host_is_build_plus_to_host(vox, assets, name, N, z0, z1, mode):
    """dxrv_voxelize_mesh_to_host (upload + build + voxelize + read-back as one call; the host pool starts on the buffer
    before the upload) leaves the same host grid as dxrv_build_bvh + dxrv_voxelize + dxrv_fetch_grid, with every transport;
    a call that fails (bad arguments, bad mesh) leaves the pool usable."""
    from dxrvoxelizer_b200 import _lib as L
    m = assets(name)
    other = assets("bunny.obj" if name != "bunny.obj" else "dragon.obj")
    vox.build_bvh(m)
    vox.voxelize(N, mode, z0, z1)
    want = vox.fetch_bits()
    vb, ib = np.ascontiguousarray(m.vertex_bytes), np.ascontiguousarray(m.indices)
    try:
        for transport in (L.READ_BACK_SPARSE, L.READ_BACK_DENSE, L.READ_BACK_AUTO):
            vox.set_read_back(transport)
            vox.build_bvh(other)                                  # (the call must rebuild: another mesh is resident)
            got = np.full(want.shape, 0xDEADBEEF, np.uint32)
            vox.voxelize_mesh_to_host(vb.ctypes.data, m.num_vertices, m.stride, ib.ctypes.data, ib.size, N, mode, z0, z1, got.ctypes.data, got.nbytes, chunks=4)
            assert np.array_equal(got, want), transport
            assert np.array_equal(vox.fetch_bits(), want)
        vox.set_read_back(L.READ_BACK_SPARSE)
        got = np.full(want.shape, 0xDEADBEEF, np.uint32)
        with pytest.raises(d.DxrvError):                          # wrong byte count: refused before the pool starts
            vox.voxelize_mesh_to_host(vb.ctypes.data, m.num_vertices, m.stride, ib.ctypes.data, ib.size, N, mode, z0, z1, got.ctypes.data, got.nbytes - 4)
        bad = ib.copy(); bad[5] = m.num_vertices + 7                 # an index out of range: the build reports it after the pool has started
        with pytest.raises(d.DxrvError):
            vox.voxelize_mesh_to_host(vb.ctypes.data, m.num_vertices, m.stride, bad.ctypes.data, bad.size, N, mode, z0, z1, got.ctypes.data, got.nbytes)
        vox.voxelize_mesh_to_host(vb.ctypes.data, m.num_vertices, m.stride, ib.ctypes.data, ib.size, N, mode, z0, z1, got.ctypes.data, got.nbytes)
        assert np.array_equal(got, want)
    finally:
        vox.set_read_back(L.READ_BACK_AUTO)


@pytest.mark.gpu
def test_to_host_sparse_transport_falls_back_when_the_grid_does_not_compress(vox):
    """A soup of random triangles: most bricks are mixed, the blob would exceed half the dense size -> dense copy."""
    from dxrvoxelizer_b200 import _lib as L
    rng = np.random.default_rng(5)
    pos = rng.uniform(-1, 1, size=(3000, 3)).astype(np.float32)
    m = d.Mesh.from_arrays(pos, rng.integers(0, 3000, size=(4000, 3)).astype(np.uint32))
    vox.build_bvh(m)
    N = 128
    vox.voxelize(N, d.MODE_PARITY)
    want = vox.fetch_bits()
    assert vox.fetch_sparse().size > want.nbytes // 2
    vox.set_read_back(L.READ_BACK_SPARSE)
    try:
        got = np.full(want.shape, 0xDEADBEEF, np.uint32)
        vox.voxelize_to_host(N, d.MODE_PARITY, 0, N, got.ctypes.data, got.nbytes)
        assert np.array_equal(got, want)
    finally:
        vox.set_read_back(L.READ_BACK_AUTO)

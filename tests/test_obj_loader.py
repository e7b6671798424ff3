"""Mesh-input stage: the product loader must reproduce the
reference's XUSG::ObjLoader::Import(file, true, true) byte for byte.

Pins: (1) CRC-32 of the outputs of the reference's own XUSGObjLoader.cpp compiled on Linux
(SURVEY.md section 8c); (2) a live comparison against that reference loader (oracle/_ref, built from
/root/reference by oracle/Makefile) on the shipped meshes and on synthetic edge-case OBJ texts.
"""
import os
import zlib

import numpy as np
import pytest

import dxrvoxelizer_b200 as d

PINS = {  # name: (numVerts, numIndices, crc(idx), crc(pos), crc(vb), first triangle)
    "dragon.obj": (50000, 300000, 0x91B1C8C5, 0x35D2AE37, 0xC61EE72B, (47437, 42256, 29824)),
    "bunny.obj": (34835, 208998, 0xBB231548, 0x618F00B8, 0x02C3F513, (34834, 33422, 12706)),
    "TuringBowl.obj": (23188, 68232, 0x89E8F1D5, 0xD3E68503, 0x426597DF, (23187, 15358, 23186)),
}
BOUNDS = {  # Voxelizer.cpp:52-57 on the reference loader's AABB
    "dragon.obj": (0.0, 4.96995, 0.0, 7.0467),
    "bunny.obj": (0.0, 4.927, 0.0, 5.0151),
    "TuringBowl.obj": (0.0, -9.3168, 0.0, 168.4234),
}


@pytest.mark.parametrize("name", sorted(PINS))
def test_loader_matches_reference_crc_pins(name):
    m = d.load_obj(d.asset_path(name))
    nv, ni, crc_idx, crc_pos, crc_vb, first = PINS[name]
    assert (m.num_vertices, m.indices.size, m.stride) == (nv, ni, 24)
    assert zlib.crc32(m.indices.tobytes()) == crc_idx
    assert zlib.crc32(np.ascontiguousarray(m.vertices[:, :3]).tobytes()) == crc_pos
    assert zlib.crc32(m.vertex_bytes.tobytes()) == crc_vb
    assert tuple(int(i) for i in m.indices[:3]) == first
    np.testing.assert_allclose(m.bound, BOUNDS[name], rtol=2e-5, atol=1e-6)


def _same_as_reference(path, oracle_mod, nan_bits=True):
    if not oracle_mod.ref_loader_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    vb, ib, st, aabb = oracle_mod.ref_load_obj(path)
    m = d.load_obj(path)
    assert m.stride == st
    assert np.array_equal(m.indices, ib)
    if nan_bits:
        assert m.vertex_bytes.tobytes() == vb.tobytes()   # bit-exact, NaNs included
        assert m.aabb.tobytes() == aabb.tobytes()
    else:
        # every value bit for bit (signed zeros and infinities included), a NaN where the reference has a NaN -- but not the
        # NaN's sign and payload: x86 hands on the FIRST operand's NaN, so they follow the compiler's operand order (they
        # differ between two builds of the reference itself), not the loader's semantics
        for ours, ref in ((m.vertex_bytes, vb), (m.aabb.view(np.uint8), aabb.view(np.uint8))):
            a, b = ours.view(np.uint32), ref.view(np.uint32)
            nan_a, nan_b = np.isnan(ours.view(np.float32)), np.isnan(ref.view(np.float32))
            assert np.array_equal(nan_a, nan_b)
            assert np.array_equal(a[~nan_a], b[~nan_b])


@pytest.mark.parametrize("name", sorted(PINS))
def test_loader_byte_identical_to_reference_loader(name, oracle_mod):
    _same_as_reference(d.asset_path(name), oracle_mod)


EDGE_CASES = {
    "quad_fan_negative": """# comment
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0.5
v 0.5 0.5 1
f 1 2 3 4
f -1 -2 -3
f 1 2 5
""",
    "with_normals_split": """o thing
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vn 0 0 1
vn 0 1 0
vn 1 0 0
f 1//1 2//1 3//1
f 1//2 3//2 4//2
f 2//3 3//1 4//3
s off
""",
    "with_texcoords_and_normals": """mtllib x.mtl
v 0 0 0
v 2 0 0
v 2 2 0
v 0 2 1
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 1
vn 0 0.6 0.8
usemtl m
f 1/1/1 2/2/1 3/3/2 4/4/2
f 3/3/1 2/2/2 1/1/2
""",
    "texcoords_only": """v 0 0 0
v 1 0 0
v 0 1 0
v 0 0 1
vt 0 0
vt 1 1
f 1/1 2/2 3/1
f 1/1 3/2 4/2 2/1
""",
    "crlf_and_exponents": "v 1e-1 -2.5E+0 3\r\nv +4 5. .5\r\nv 7 8 9\r\nv -1 -1 -1\r\nf 1 2 3\r\nf 2 3 4\r\n",
    "degenerate_face_gives_nan_normals": """v 0 0 0
v 1 0 0
v 2 0 0
v 0 1 0
f 1 2 3
f 1 2 4
""",
}


@pytest.mark.parametrize("case", sorted(EDGE_CASES))
def test_loader_edge_cases_match_reference_loader(case, tmp_path, oracle_mod):
    p = tmp_path / (case + ".obj")
    p.write_bytes(EDGE_CASES[case].encode())
    _same_as_reference(str(p), oracle_mod)


def _decimal_torture(seed, count):
    """`count` vertex records whose coordinates are written in every decimal style the fast number path takes or must
    refuse: few and many digits, values next to the midpoint of two adjacent floats (where rounding through a double
    would differ from rounding once), leading zeros, bare '.5' / '5.', signs, exponents, more than 18 digits."""
    rng = np.random.default_rng(seed)
    out = []
    def number():
        kind = int(rng.integers(0, 8))
        if kind == 0:
            return "%.*f" % (int(rng.integers(0, 12)), float(rng.uniform(-2000, 2000)))
        if kind == 1:
            return "%.9g" % float(np.float32(rng.uniform(-10, 10)))
        if kind == 2:                                        # the midpoint of two adjacent floats, cut after k digits
            f = np.float32(rng.uniform(1, 4)); g = np.nextafter(f, np.float32(8))
            return "%.*f" % (int(rng.integers(8, 18)), (float(f) + float(g)) / 2)
        if kind == 3:                                        # ... and written out exactly (a tie: rounds to even)
            f = np.float32(rng.uniform(1, 4)); g = np.nextafter(f, np.float32(8))
            return "%.30f" % ((float(f) + float(g)) / 2)
        if kind == 4:
            return "%s%d.%d" % ("-" if rng.random() < 0.5 else "", int(rng.integers(0, 100000)), int(rng.integers(0, 10 ** 11)))
        if kind == 5:
            return ["0", "-0", "+1.5", ".5", "-.25", "5.", "000012.3400", "16777217", "0.1", "123456789.123456789"][int(rng.integers(0, 10))]
        if kind == 6:
            return "%.6e" % float(rng.uniform(-1e3, 1e3))
        return "%.17g" % float(rng.uniform(-1, 1))
    for _ in range(count):
        out.append("v %s %s %s" % (number(), number(), number()))
    faces = ["f %d %d %d" % (i + 1, i + 2, i + 3) for i in range(0, count - 3, 3)]
    return "\n".join(out + faces) + "\n"


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_decimal_conversion_matches_reference_loader(seed, tmp_path, oracle_mod):
    """Every coordinate bit for bit as the reference's fscanf("%f") reads it (product: one exact double division for
    plain decimals, std::from_chars behind it)."""
    p = tmp_path / ("decimals_%d.obj" % seed)
    p.write_text(_decimal_torture(seed, 6000))
    _same_as_reference(str(p), oracle_mod)
    # (Python's float(text) -> float32 is NOT a checker here: it rounds twice, and the values next to float midpoints in
    # this file are exactly where that differs from fscanf / the product -- e.g. seed 1, vertex 1: ...437 against ...438.)


def test_random_obj_files_match_reference_loader(tmp_path, oracle_mod):
    """A slice of the differential fuzzer (tools/fuzz_obj.py: random records, all four corner syntaxes, polygons,
    negative indices, comments, odd blanks, CRLF, no final newline; 1 650 files compared when it was written)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fuzz_obj", os.path.join(os.path.dirname(__file__), "..", "tools", "fuzz_obj.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    for seed in range(40):
        p = tmp_path / ("fuzz_%d.obj" % seed)
        p.write_bytes(fz.make_obj(seed).encode())
        _same_as_reference(str(p), oracle_mod)
    # files only the reference's TOKEN grammar explains (faces continued on the next line, several faces on one line,
    # unknown records, lines around the reference's 255-byte fgets buffer): the product's second parser
    for seed in range(30):
        p = tmp_path / ("exotic_%d.obj" % seed)
        p.write_bytes(fz.make_exotic(seed).encode())
        _same_as_reference(str(p), oracle_mod)


@pytest.mark.parametrize("threads", ["1", "5"])
def test_multi_chunk_file_and_non_finite_coordinates_match_reference_loader(tmp_path, oracle_mod, meshes_mod, monkeypatch, threads):
    """A 3 MB file (icosphere(6): 81 920 triangles, parsed in several chunks) as written, and with nan / inf / signed
    zeros sprinkled into the coordinates -- vertex 0 included: fscanf reads them, the face normals and the AABB
    (first strict minimum wins, a NaN in vertex 0 poisons its axis) must come out bit for bit, NaNs as NaNs."""
    from bench_configs import write_obj
    monkeypatch.setenv("DXRV_OBJ_THREADS", threads)
    p = tmp_path / "ico6.obj"
    write_obj(str(p), meshes_mod.icosphere(6, seed=3, rotate=True, normals=False))
    _same_as_reference(str(p), oracle_mod)
    txt = p.read_text().splitlines()
    rng = np.random.default_rng(5)
    for i in rng.integers(1, 40000, 300):
        if txt[i].startswith("v "):
            parts = txt[i].split()
            parts[int(rng.integers(1, 4))] = ["nan", "inf", "-inf", "-0.0", "0.0", "-0"][int(rng.integers(0, 6))]
            txt[i] = " ".join(parts)
    txt[1] = "v nan 0.5 -0.0"
    q = tmp_path / "ico6_nonfinite.obj"
    q.write_text("\n".join(txt) + "\n")
    _same_as_reference(str(q), oracle_mod, nan_bits=False)


def test_loader_semantics_without_reference(tmp_path):
    """Same conventions checked directly (runs even when oracle/_ref is absent)."""
    p = tmp_path / "t.obj"
    p.write_text(EDGE_CASES["quad_fan_negative"])
    m = d.load_obj(str(p))
    # 2 (quad fan) + 1 + 1 triangles, index array reversed as a whole (XUSGObjLoader.cpp:227)
    assert m.indices.tolist() == [4, 1, 0, 2, 3, 4, 3, 2, 0, 2, 1, 0]
    # z negated (XUSGObjLoader.cpp:198)
    assert m.vertices[3, :3].tolist() == [0.0, 1.0, -0.5]
    n = np.linalg.norm(m.vertices[:, 3:6], axis=1)
    np.testing.assert_allclose(n, 1.0, rtol=1e-6)


def test_parse_from_memory_equals_load_from_file(tmp_path):
    """dxrv_obj_parse: OBJ text in memory (no terminator, embedded NUL tolerated as an ordinary byte) = dxrv_obj_load."""
    for name in ("bunny.obj", "TuringBowl.obj"):
        path = d.asset_path(name)
        a = d.load_obj(path)
        b = d.load_obj(text=open(path, "rb").read())
        assert a.stride == b.stride and np.array_equal(a.indices, b.indices)
        assert a.vertex_bytes.tobytes() == b.vertex_bytes.tobytes() and a.aabb.tobytes() == b.aabb.tobytes()
    for case, text in EDGE_CASES.items():
        p = tmp_path / (case + ".obj")
        p.write_bytes(text.encode())
        a, b = d.load_obj(str(p)), d.load_obj(text=text.encode())
        assert np.array_equal(a.indices, b.indices) and a.vertex_bytes.tobytes() == b.vertex_bytes.tobytes()
    with pytest.raises(d.DxrvError):
        d.load_obj(text=b"v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n")


def test_loader_missing_file_fails_like_reference():
    with pytest.raises(d.DxrvError) as e:
        d.load_obj("/nonexistent/mesh.obj")
    assert e.value.code == -5   # DXRV_ERR_IO; the reference's Import returns false (XUSGObjLoader.cpp:21-23)


def test_loader_rejects_out_of_range_index(tmp_path):
    p = tmp_path / "bad.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 7\n")
    with pytest.raises(d.DxrvError):
        d.load_obj(str(p))

"""The drop-in boundary: libdxrv.so loads, exports every symbol that
include/dxrv.h declares, and -- on a box without a GPU -- fails loudly instead of falling back."""
import ctypes
import os
import re
import subprocess

import pytest

from dxrvoxelizer_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dxrv.h")).read()
    return sorted(set(re.findall(r"DXRV_API[^;(]*?\b(dxrv_\w+)\s*\(", text)))


def test_header_binding_and_library_agree():
    declared = _declared_symbols()
    assert len(declared) >= 25
    assert sorted(L.SIGNATURES) == declared            # the Python binding covers the whole ABI
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (dxrv_\w+)", out)))
    assert exported == declared                        # nothing undeclared leaks out either


def test_product_does_not_reference_the_oracle():
    """The product path must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "dxrvoxelizer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), f
                assert "libdxrv_oracle" not in text and "dxrv_oracle.h" not in text.replace("oracle/dxrv_oracle.h", ""), f
    deps = subprocess.run(["ldd", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps


def test_error_convention_without_context():
    lib = L.lib()
    assert lib.dxrv_create(None, 0) == L.ERR_INVALID_ARG
    assert b"null" in lib.dxrv_last_error(None)
    h = ctypes.c_void_p()
    assert lib.dxrv_obj_load(b"/nonexistent.obj", ctypes.byref(h)) == L.ERR_IO
    assert b"cannot open" in lib.dxrv_last_error(None)
    assert lib.dxrv_voxelize(None, 64, 1, 0, 64) == L.ERR_INVALID_ARG
    lib.dxrv_destroy(None)   # harmless
    lib.dxrv_obj_free(None)


def test_no_cpu_fallback_when_no_device():
    """Without a usable CUDA device dxrv_create must fail (there is no CPU path to fall back to)."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    code = ("import ctypes,sys; sys.path.insert(0, %r); from dxrvoxelizer_b200 import _lib as L; "
            "h=ctypes.c_void_p(); rc=L.lib().dxrv_create(ctypes.byref(h),0); "
            "print(rc, L.lib().dxrv_last_error(None).decode())" % ROOT)
    out = subprocess.run(["python", "-c", code], capture_output=True, text=True, env=env, check=True).stdout
    assert out.startswith("-2 ") and "no CPU fallback" in out


def test_cli_reports_failure_like_reference_init(tmp_path):
    exe = os.path.join(ROOT, "dxrvoxelizer_b200", "dxrvoxelizer")
    assert os.path.exists(exe)
    r = subprocess.run([exe, "-mesh", str(tmp_path / "missing.obj"), "0.0", "2.8", "0.0", "0.03"], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr   # Init returns false when Import fails


def test_cli_batch_mode_reports_failures(tmp_path):
    """-batch list.txt (dxrv_voxelize_obj_batch): a missing list, an empty list and -- without a GPU -- the loud failure."""
    exe = os.path.join(ROOT, "dxrvoxelizer_b200", "dxrvoxelizer")
    r = subprocess.run([exe, "-batch", str(tmp_path / "nolist.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr
    empty = tmp_path / "empty.txt"
    empty.write_text("# nothing\n\n")
    r = subprocess.run([exe, "-batch", str(empty)], capture_output=True, text=True)
    assert r.returncode == 1 and "lists no meshes" in r.stderr
    lst = tmp_path / "list.txt"
    lst.write_text("# one mesh\n%s\n" % os.path.join(ROOT, "assets", "missing.obj"))
    r = subprocess.run([exe, "-batch", str(lst), "-grid", "64", "-mode", "parity", "-streams", "2"], capture_output=True, text=True,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 1 and "no CUDA device" in r.stderr


def _dry(*args):
    import json
    exe = os.path.join(ROOT, "dxrvoxelizer_b200", "dxrvoxelizer")
    r = subprocess.run([exe, "-dryrun"] + list(args), capture_output=True, text=True, check=True)
    return json.loads(r.stdout)


def test_cli_argument_grammar_follows_the_reference():
    """DXRVoxelizer::ParseCommandLineArgs (DXRVoxelizer.cpp:363-408): '-' or '/' prefix, case-insensitive names, a value may
    start with '-' only when a digit or '.' follows, `-mesh path [x y z scale]`, defaults Assets/bunny.obj and (0,0,0,1)
    (DXRVoxelizer.cpp:36-37); -warp / -uma accepted.  The .bat files: Dragon.bat = `-mesh Assets/dragon.obj`,
    TuringBowl.bat = `-mesh Assets/TuringBowl.obj 0.0 2.8 0.0 0.03`."""
    d0 = _dry()
    assert d0["mesh"] == "Assets/bunny.obj" and d0["posScale"] == [0, 0, 0, 1] and d0["grid"] == 64 and d0["mode"] == "shader"
    t = _dry("-mesh", "Assets/TuringBowl.obj", "0.0", "2.8", "0.0", "0.03")
    assert t["mesh"] == "Assets/TuringBowl.obj" and t["posScale"] == pytest.approx([0.0, 2.8, 0.0, 0.03])
    assert _dry("/MESH", "a.obj", "-WARP", "-uma")["mesh"] == "a.obj"                       # '/' prefix, any case
    n = _dry("-mesh", "a.obj", "-0.5", ".5", "-1", "2", "-grid", "256")
    assert n["posScale"] == pytest.approx([-0.5, 0.5, -1.0, 2.0]) and n["grid"] == 256      # negative numbers are values
    p = _dry("-mesh", "a.obj", "1.5", "-mode", "parity")                                     # fewer than four numbers: the rest keep their defaults
    assert p["posScale"] == pytest.approx([1.5, 0, 0, 1]) and p["mode"] == "parity"
    assert _dry("-mesh", "-grid", "128")["mesh"] == "Assets/bunny.obj"                       # an option is not a value
    assert _dry("-mesh", "/abs/path/mesh.obj")["mesh"] == "/abs/path/mesh.obj"              # (Linux deviation: absolute paths)
    b = _dry("-batch", "list.txt", "-streams", "8", "-gpus", "2", "-slab", "3", "9", "-out", "g.bin", "-view", "v.png")
    assert (b["batch"], b["streams"], b["gpus"], b["slab"], b["out"], b["view"]) == ("list.txt", 8, 2, [3, 9], "g.bin", "v.png")


def test_bat_file_equivalents_pass_the_reference_arguments():
    """Dragon.sh / TuringBowl.sh = Bin/Dragon.bat / Bin/TuringBowl.bat: the mesh and, for the bowl, `0.0 2.8 0.0 0.03`."""
    import json
    for script, mesh, pos_scale in (("Dragon.sh", "dragon.obj", [0, 0, 0, 1]), ("TuringBowl.sh", "TuringBowl.obj", [0.0, 2.8, 0.0, 0.03])):
        r = subprocess.run([os.path.join(ROOT, script), "-dryrun", "-grid", "256"], capture_output=True, text=True, check=True)
        j = json.loads(r.stdout)
        assert j["mesh"].endswith(mesh) and os.path.exists(j["mesh"]) and j["grid"] == 256 and j["mode"] == "shader"
        assert j["posScale"] == pytest.approx(pos_scale)


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/dxrv.h must compile as C99 (what a cgo / JNI / ctypes-free FFI would include) and C++11."""
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    c = tmp_path / "hdr.c"
    c.write_text('#include "dxrv.h"\nint main(void) { return DXRV_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(c)], check=True)
    cpp = tmp_path / "hdr.cpp"
    cpp.write_text('#include "dxrv.h"\nint main() { return DXRV_OK; }\n')
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(cpp)], check=True)


def test_plain_c_example_links_against_the_library(tmp_path):
    """examples/voxelize.c: the ABI bound from C99 -- compiles, links against libdxrv.so, loads a mesh (host code) and, without
    a GPU, stops at dxrv_create with the library's message."""
    import shutil
    import dxrvoxelizer_b200 as d
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    exe = tmp_path / "voxelize"
    libdir = os.path.join(ROOT, "dxrvoxelizer_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "voxelize.c"), "-L", libdir, "-ldxrv", "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
    r = subprocess.run([str(exe), d.asset_path("bunny.obj"), "64"], capture_output=True, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert "34835 vertices, 69666 triangles" in r.stdout
    assert r.returncode == 1 and "no CUDA device" in r.stderr

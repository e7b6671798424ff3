"""dxrv_voxelize_obj_batch (include/dxrv.h; csrc/batch.cpp), host side: the pipeline's threads, ordering, bounded
look-ahead and error paths run here against stubs of the context entry points (tools/batch_mock.cpp: no GPU), and the
argument checks of the real library.  The GPU test is tests/test_gpu_batch.py."""
import os
import shutil
import subprocess

import pytest

import dxrvoxelizer_b200 as d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_pipeline_against_stub_contexts(tmp_path):
    exe = tmp_path / "batch_mock"
    csrc = os.path.join(ROOT, "dxrvoxelizer_b200", "csrc")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + csrc, os.path.join(ROOT, "tools", "batch_mock.cpp"),
                    os.path.join(csrc, "batch.cpp"), os.path.join(csrc, "obj_loader.cpp"), "-o", str(exe), "-lpthread"], check=True)
    work = tmp_path / "objs"
    work.mkdir()
    r = subprocess.run([str(exe), str(work)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout + r.stderr


def test_argument_checks_of_the_library(tmp_path):
    with pytest.raises(d.DxrvError) as e:
        d.voxelize_obj_batch([], [str(tmp_path / "a.obj")], 64)
    assert e.value.code == -1 and "null argument" in str(e.value)

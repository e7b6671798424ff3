"""Regenerates tests/golden/oracle_counts_64.json from the CPU oracle (run from the repo root)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import dxrvoxelizer_b200 as d  # noqa: E402
import oracle  # noqa: E402


def popcount(a):
    return int(np.unpackbits(np.ascontiguousarray(a).view(np.uint8)).sum())


out = {"generated_by": "python tests/golden/make_oracle_counts.py", "grid": 64, "meshes": {}}
for n in ("dragon.obj", "bunny.obj", "TuringBowl.obj"):
    m = d.load_obj(d.asset_path(n))
    s = oracle.voxelize(m.vertices, m.indices, 64, 0)
    p = oracle.voxelize(m.vertices, m.indices, 64, 1)
    out["meshes"][n] = {"inside_shader": popcount(s["bits"]), "inside_parity": popcount(p["bits"]),
                        "crossings_parity": p["crossings"], "shader_xor_parity": popcount(s["bits"] ^ p["bits"])}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_counts_64.json"), "w"), indent=1)

# bit grids (and the dragon's texels) at the reference's GRID_SIZE, as regression fixtures
grids = {}
for n in ("dragon.obj", "bunny.obj", "TuringBowl.obj"):
    m = d.load_obj(d.asset_path(n))
    for mode, name in ((0, "shader"), (1, "parity")):
        r = oracle.voxelize(m.vertices, m.indices, 64, mode, texels=(mode == 0 and n == "dragon.obj"))
        grids["%s_%s" % (n.split(".")[0], name)] = r["bits"]
        if r["texels"] is not None:
            grids["%s_texels" % n.split(".")[0]] = r["texels"]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_grids_64.npz"), **grids)

#!/usr/bin/env python3
"""Import the reference's mesh fixtures (Bin/Assets/*.obj) as xz-compressed copies.

The three OBJ files are the only input fixtures the reference ships
(/root/reference/Bin/Assets, used by Bin/Dragon.bat and Bin/TuringBowl.bat).  They are
DATA, not source: the GPU box has no /root/reference, so the parity tests and bench.py
need them in-tree.  They are stored xz-compressed under assets/ and unpacked on demand by
dxrvoxelizer_b200.assets.asset_path().

Run (in the build container only):  python tools/import_assets.py
"""
import hashlib
import json
import lzma
import os
import sys

SRC = "/root/reference/Bin/Assets"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets")


def main():
    manifest = {}
    for name in ("dragon.obj", "bunny.obj", "TuringBowl.obj"):
        with open(os.path.join(SRC, name), "rb") as f:
            raw = f.read()
        out = os.path.join(DST, name + ".xz")
        with open(out, "wb") as f:
            f.write(lzma.compress(raw, preset=9 | lzma.PRESET_EXTREME))
        manifest[name] = {"bytes": len(raw), "sha256": hashlib.sha256(raw).hexdigest()}
        print(name, len(raw), "->", os.path.getsize(out))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    sys.exit(main())

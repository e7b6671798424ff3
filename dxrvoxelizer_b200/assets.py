"""The reference's mesh fixtures (Bin/Assets/*.obj of StarsX/DXRVoxelizer), shipped xz-compressed
under assets/ (see tools/import_assets.py) and unpacked on demand into a cache directory."""
import hashlib
import json
import lzma
import os
import tempfile

_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets")
NAMES = ("dragon.obj", "bunny.obj", "TuringBowl.obj")


def asset_path(name):
    """Path of the unpacked OBJ file `name` (e.g. "dragon.obj"), verified against MANIFEST.json."""
    if name not in NAMES:
        raise KeyError(name)
    with open(os.path.join(_ROOT, "MANIFEST.json")) as f:
        manifest = json.load(f)[name]
    cache = os.environ.get("DXRV_ASSET_CACHE", os.path.join(tempfile.gettempdir(), "dxrv_assets_%d" % os.getuid()))
    os.makedirs(cache, exist_ok=True)
    out = os.path.join(cache, name)
    if not (os.path.exists(out) and os.path.getsize(out) == manifest["bytes"]):
        with open(os.path.join(_ROOT, name + ".xz"), "rb") as f:
            raw = lzma.decompress(f.read())
        if hashlib.sha256(raw).hexdigest() != manifest["sha256"]:
            raise IOError("asset %s is corrupt" % name)
        tmp = out + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(raw)
        os.replace(tmp, out)
    return out

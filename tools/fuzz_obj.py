"""Differential fuzzing of the product OBJ loader against the reference's own XUSGObjLoader.cpp (oracle/_ref).

  python tools/fuzz_obj.py [--seeds A:B] [--scale S] [--exotic] [--dir /tmp/fuzz_obj]        (DXRV_OBJ_THREADS=k: chunks per file)

Every seed writes one OBJ text of random records -- positions, normals, texture coordinates, faces in one corner syntax
per file, polygons, negative indices, comments, blank lines, leading blanks, tabs, CRLF, no
final newline, groups / materials, numbers in every decimal style -- with all indices in range (the reference reads out of
bounds otherwise), loads it with both loaders in a child process (a crash of the reference must not stop the run) and
compares vertex bytes, indices, stride and AABB bit for bit.  Build container only: oracle/_ref needs /root/reference.
"""
import argparse
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def number(rng):
    kind = int(rng.integers(0, 9))
    if kind == 0:
        return "%.*f" % (int(rng.integers(0, 10)), float(rng.uniform(-50, 50)))
    if kind == 1:
        return "%.9g" % float(np.float32(rng.uniform(-10, 10)))
    if kind == 2:
        return "%d" % int(rng.integers(-20, 20))
    if kind == 3:
        return "%.4e" % float(rng.uniform(-1e3, 1e3))
    if kind == 4:
        return ["0", "-0", "+1.5", ".5", "-.25", "5.", "0001.50", "1E2", "-3.e-1", "+.5e+1"][int(rng.integers(0, 10))]
    if kind == 5:
        return "%.17g" % float(rng.uniform(-1, 1))
    if kind == 6:
        f = np.float32(rng.uniform(1, 4)); g = np.nextafter(f, np.float32(8))
        return "%.*f" % (int(rng.integers(8, 25)), (float(f) + float(g)) / 2)
    return "%.6f" % float(rng.uniform(-2, 2))


def blank(rng, wild):
    if not wild or rng.random() < 0.8:
        return " "
    return [" ", "  ", "\t", " \t ", "   "][int(rng.integers(0, 5))]


SCALE = 1          # --scale: multiplies the record counts (files beyond 128 KB are parsed in several chunks)


def make_obj(seed):
    rng = np.random.default_rng(seed)
    wild = rng.random() < 0.5                      # odd blanks, leading blanks, junk records
    # the corner syntax follows from what the file defines (the reference reads "/vt" from every corner of a file with
    # texture coordinates and "/vn" from every corner of a file with normals, XUSGObjLoader.cpp:243-256)
    style = int(rng.integers(0, 4))                # 0: v   1: v//vn   2: v/vt/vn   3: v/vt
    nv = int(rng.integers(3, 60 * SCALE))
    nn = int(rng.integers(1, 20 * SCALE)) if style in (1, 2) else 0
    nt = int(rng.integers(1, 20 * SCALE)) if style in (2, 3) else 0
    nf = int(rng.integers(1, 80 * SCALE))
    eol = "\r\n" if rng.random() < 0.2 else "\n"
    lines = []

    def lead():
        return blank(rng, True) if wild and rng.random() < 0.1 else ""

    def junk():
        r = rng.random()
        if r < 0.3:
            lines.append("# a comment with f 1 2 3 and v 1 2 3 inside")
        elif r < 0.5:
            lines.append("")
        elif r < 0.6:
            lines.append(["g group1", "o object", "s 1", "s off", "usemtl mat", "mtllib file.mtl"][int(rng.integers(0, 6))])
        elif r < 0.65 and wild:
            lines.append("   ")

    # definitions first (a face may only use what is already defined -- negative indices count from here)
    interleave = rng.random() < 0.3
    defs = [("v", 3)] * nv + [("vn", 3)] * nn + [("vt", 2 if rng.random() < 0.7 else 3)] * nt
    if interleave:
        rng.shuffle(defs)
    for kind, k in defs:
        if rng.random() < 0.1:
            junk()
        extra = ""
        if kind == "v" and wild and rng.random() < 0.1:
            extra = blank(rng, wild) + blank(rng, wild).join(number(rng) for _ in range(int(rng.integers(1, 4))))   # w / colours
        lines.append(lead() + kind + blank(rng, wild) + blank(rng, wild).join(number(rng) for _ in range(k)) + extra
                     + (blank(rng, True) if wild and rng.random() < 0.1 else ""))
    for _ in range(nf):
        if rng.random() < 0.1:
            junk()
        corners = 3 if rng.random() < 0.7 else int(rng.integers(4, 8))
        st = style
        toks = []
        for _ in range(corners):
            neg = rng.random() < 0.15
            vi = -int(rng.integers(1, nv + 1)) if neg else int(rng.integers(1, nv + 1))
            ti = (-int(rng.integers(1, nt + 1)) if neg else int(rng.integers(1, nt + 1))) if nt else 0
            ni = (-int(rng.integers(1, nn + 1)) if neg else int(rng.integers(1, nn + 1))) if nn else 0
            if st == 0:
                toks.append("%d" % vi)
            elif st == 1:
                toks.append("%d//%d" % (vi, ni))
            elif st == 2:
                toks.append("%d/%d/%d" % (vi, ti, ni))
            else:
                toks.append("%d/%d" % (vi, ti))
        lines.append(lead() + "f" + blank(rng, wild) + blank(rng, wild).join(toks) + (blank(rng, True) if wild and rng.random() < 0.1 else ""))
    text = eol.join(lines)
    if rng.random() < 0.8:
        text += eol
    return text


def make_exotic(seed):
    """Files the reference's TOKEN grammar accepts but no exporter writes: faces continued on the next line, several
    records on one line, unknown records (`vp`, `l`, `p`), comment / group lines longer than the reference's 255-byte
    fgets buffer (what is left of such a line is then read as records), vt records of one to three numbers.  These take
    the product's second parser (parseObj, the token-by-token restatement of the reference's grammar)."""
    rng = np.random.default_rng(seed)
    nv = int(rng.integers(4, 40))
    nf = int(rng.integers(2, 30))
    lines = []
    pending = []

    def flush():
        if pending:
            lines.append(" ".join(pending))
            pending.clear()

    def emit(rec):
        pending.append(rec)
        if rng.random() < 0.8:
            flush()

    def noise():
        r = rng.random()
        if r < 0.25:
            flush(); lines.append("# " + "x" * int(rng.integers(230, 300)) + " v 1 2 3")      # around the 255-byte limit
        elif r < 0.4:
            flush(); lines.append("g " + "name" * int(rng.integers(60, 80)))
        elif r < 0.55:
            flush(); lines.append(["vp 0.5 0.5", "l 1 2 3", "p 1", "curv 0 1 1 2", "s 2"][int(rng.integers(0, 5))])
        elif r < 0.7:
            flush(); lines.append("#" + "y" * 254)
        elif r < 0.8:
            flush(); lines.append("#" + "z" * 253 + " ")

    for _ in range(nv):
        if rng.random() < 0.15:
            noise()
        flush(); lines.append("v " + " ".join(number(rng) for _ in range(3)))   # (a second `v` on the line would be dropped by the reference: one per line)
    for _ in range(nf):
        if rng.random() < 0.2:
            noise()
        corners = [str(int(rng.integers(1, nv + 1))) for _ in range(int(rng.integers(3, 7)))]
        if rng.random() < 0.3:                                   # continued on the next line(s)
            flush()
            cut = int(rng.integers(1, len(corners)))
            lines.append("f " + " ".join(corners[:cut]))
            lines.append(" ".join(corners[cut:]))
        else:
            emit("f " + " ".join(corners))
    flush()
    return "\n".join(lines) + ("\n" if rng.random() < 0.8 else "")


def child(paths):
    sys.path.insert(0, ROOT)
    import oracle
    import dxrvoxelizer_b200 as d
    for p in paths:
        try:
            vb, ib, st, aabb = oracle.ref_load_obj(p)
            ref = (vb.tobytes(), ib.tobytes(), st, aabb.tobytes())
        except Exception as e:                                     # noqa: BLE001
            ref = ("ref failed: %s" % e,)
        try:
            m = d.load_obj(p)
            ours = (m.vertex_bytes.tobytes(), m.indices.tobytes(), m.stride, m.aabb.tobytes())
        except Exception as e:                                     # noqa: BLE001
            ours = ("ours failed: %s" % e,)
        if ref != ours:
            what = "DIFF"
            if len(ref) == 4 and len(ours) == 4:
                what += " vb=%s ib=%s stride=%s/%s aabb=%s" % (ref[0] == ours[0], ref[1] == ours[1], ref[2], ours[2], ref[3] == ours[3])
            else:
                what += " %s | %s" % (ref[0] if len(ref) == 1 else "ref ok", ours[0] if len(ours) == 1 else "ours ok")
            print(what, p, flush=True)
        else:
            print("same", p, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="0:500")
    ap.add_argument("--dir", default="/tmp/fuzz_obj")
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--exotic", action="store_true", help="token-grammar files (make_exotic) instead of exporter-style ones")
    ap.add_argument("--child", nargs="*")
    a = ap.parse_args()
    global SCALE
    SCALE = a.scale
    if a.child is not None:
        child(a.child)
        return
    os.makedirs(a.dir, exist_ok=True)
    lo, hi = (int(x) for x in a.seeds.split(":"))
    same = diff = crashed = 0
    batch = 50
    for b in range(lo, hi, batch):
        paths = []
        for s in range(b, min(hi, b + batch)):
            p = os.path.join(a.dir, "fuzz_%d.obj" % s)
            with open(p, "wb") as f:
                f.write((make_exotic(s) if a.exotic else make_obj(s)).encode())
            paths.append(p)
        while paths:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"] + paths, capture_output=True, text=True)
            done = [l for l in r.stdout.splitlines() if l.startswith(("same", "DIFF"))]
            for l in done:
                if l.startswith("same"):
                    same += 1
                else:
                    diff += 1
                    print(l)
            paths = paths[len(done):]
            if r.returncode != 0 and paths:                       # the child died on paths[0]
                crashed += 1
                print("CRASH rc=%d %s" % (r.returncode, paths[0]))
                paths = paths[1:]
    print("same=%d diff=%d crashed=%d" % (same, diff, crashed))
    return 1 if diff or crashed else 0


if __name__ == "__main__":
    sys.exit(main())

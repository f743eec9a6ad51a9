#!/usr/bin/env python
"""Convert the reference's scene inputs (model/*.txt|obj, texture/*) into the compact binary containers
under cgraytracing_b200/assets/ that travel with the repo (the GPU box has no /root/reference).

Run in the build container only:  python tools/make_assets.py [/root/reference]

* meshes   -> .cgrtmesh : b"CGRTMSH1", int32 nverts, int32 nfaces, float64 verts[nverts*3] (file space, before the
              loader's z flip / scale / translate of objects.h:348,365,384), int32 faces[nfaces*3] (0-based)
* textures -> .cgrttex  : b"CGRTTEX1", int32 w, int32 h, uint8 rgb[h*w*3], decoded with the reference's own vendored
              stb_image v2.19 through oracle/_ref/libcgref.so (main.cpp:300) so texel bytes are exactly the reference's.
Values are the data files' own numbers (parsed with correctly rounded strtod, as scanf("%lf") does); no source code is copied.
"""
import ctypes
import os
import struct
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "cgraytracing_b200", "assets")


def parse_type0(path):  # "begin / vertex x y z (x3) / end" triangle soup, objects.h:346
    v = []
    for line in open(path):
        t = line.split()
        if len(t) == 4 and t[0] == "vertex":
            v.append([float(t[1]), float(t[2]), float(t[3])])
    v = np.asarray(v, dtype=np.float64)
    assert len(v) % 3 == 0
    f = np.arange(len(v), dtype=np.int32).reshape(-1, 3)
    return v, f


def parse_indexed(path):  # types 1 and 2: count, "v x y z" lines, count, "f a b c" or "f a/b/c ..." lines (1-based)
    v, f = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            v.append([float(t[1]), float(t[2]), float(t[3])])
        elif t[0] == "f":
            f.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
    return np.asarray(v, dtype=np.float64), np.asarray(f, dtype=np.int32)


def write_mesh(name, v, f):
    with open(os.path.join(OUT, name + ".cgrtmesh"), "wb") as fp:
        fp.write(b"CGRTMSH1")
        fp.write(struct.pack("<ii", len(v), len(f)))
        fp.write(np.ascontiguousarray(v, dtype="<f8").tobytes())
        fp.write(np.ascontiguousarray(f, dtype="<i4").tobytes())
    print(f"{name}: {len(v)} vertices, {len(f)} faces")


def main():
    os.makedirs(OUT, exist_ok=True)
    write_mesh("dragon", *parse_indexed(f"{REF}/model/dragon.txt"))
    write_mesh("lowpolybunny", *parse_type0(f"{REF}/model/lowpolybunny.txt"))
    write_mesh("Mesh000", *parse_indexed(f"{REF}/model/Mesh000.obj"))
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcgref.so"))
    lib.ref_stbi_load.restype = ctypes.POINTER(ctypes.c_uint8)
    lib.ref_stbi_load.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    lib.ref_free.argtypes = [ctypes.c_void_p]
    for fn in ["ChessBoard.png", "stone.jpg", "granite_texture.jpg", "iiis.png"]:
        w, h = ctypes.c_int(), ctypes.c_int()
        p = lib.ref_stbi_load(f"{REF}/texture/{fn}".encode(), ctypes.byref(w), ctypes.byref(h))
        assert p, fn
        rgb = np.ctypeslib.as_array(p, shape=(h.value * w.value * 3,)).copy()
        lib.ref_free(p)
        name = os.path.splitext(fn)[0]
        with open(os.path.join(OUT, name + ".cgrttex"), "wb") as fp:
            fp.write(b"CGRTTEX1")
            fp.write(struct.pack("<ii", w.value, h.value))
            fp.write(rgb.tobytes())
        print(f"{name}: {w.value}x{h.value}, min texel byte {rgb.min()}")


if __name__ == "__main__":
    main()

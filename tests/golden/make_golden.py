#!/usr/bin/env python
"""Generates tests/golden/ref_golden.npz from the UNMODIFIED reference compiled here (oracle/_ref/libcgref.so, built by
`make -C oracle` from /root/reference/main.cpp + headers read in place). Run in the build container only:

    python tests/golden/make_golden.py

The reference repository holds no golden vectors of its own, so these are reference OUTPUTS on seeded
inputs; the inputs are stored next to the outputs so the tests need neither /root/reference nor libcgref.so.
Every array name is `<group>__<field>`. Groups:

  hash_<H>        Hashtable(1000001, 200.0/H) ctor + compute_coord + hash       hash.h:22-42
  isect_<preset>  closest hit of trace() over all objects                      main.cpp:50-76, objects.h, bezier.h
  tri             Triangle::intersect on one triangle                          objects.h:96-111
  tex             Texture::color / height table                                texture.h:19-72
  surf_<preset>   Object::getSurfaceColor                                      objects.h:533-539
  eye_c2          trace(flag=true) on a pixel sub-grid -> hitpoints             main.cpp:42-100,129-157
  photon_c2       trace(flag=false) with a seeded rand() stream -> flux/r2/n    main.cpp:101-128,158-166
  misc            det/inv, gammaCorr, Bezier basis                              vec3.h:95-119, util.h:45-47, bezier.h:127-162
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cgraytracing_b200 import preset  # noqa: E402
from oracle import binding as ob  # noqa: E402
from tests.util import camera_rays, random_rays  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz")

# (preset, max_tris): meshes are truncated so the fixtures stay small; the truncation is part of the stored input
ISECT_CASES = [("c1_spheres", None), ("c2_bunny_chess", None), ("c3_dragon_glass", 4000), ("default_bump", 1500), ("c4_bump_dof", None)]
HEIGHTS = (512, 768, 1024, 1080, 4096)
PIX_STEP = 24
N_PHOTON = 2500
LIBC_SEED = 20261018


def main():
    assert ob.have_ref(), "build oracle/_ref first: make -C oracle"
    g = {}
    rng = np.random.default_rng(20261018)

    # ---- hash keys
    pos = rng.uniform(-40, 60, (1500, 3))
    pos[:300] = np.round(pos[:300] * 8) / 8
    pos[300:306] = [(0, -20, 20), (-20, 0, 40), (19.999, 19.999, -5.5), (-5.3, -12.25, 30.125), (3.14159, 2.71828, 1.41421), (-40, 0, 50)]
    r = ob.Ref()
    g["hash__pos"] = pos
    for h in HEIGHTS:
        key, ixyz, cells, cl = r.hash_keys(pos, 1000001, 200.0 / h)
        g[f"hash_{h}__key"], g[f"hash_{h}__ixyz"] = key, ixyz
        g[f"hash_{h}__cells"], g[f"hash_{h}__celllength"] = np.int32(cells), np.float64(cl)
    cells3 = rng.integers(-2000, 2000, (400, 3)).astype(np.int32)
    g["hash3__ixyz"] = cells3
    g["hash3__key"] = np.array([r.hash3(*c, 1000001) for c in cells3], np.uint32)

    # ---- closest hit, per preset
    for name, max_tris in ISECT_CASES:
        s = preset(name, max_tris=max_tris)
        r = ob.Ref(s)
        o1, d1 = random_rays(700, 11)
        o2, d2 = camera_rays(1024, 768, 40)
        org, dr = np.concatenate([o1, o2]), np.concatenate([d1, d2])
        a = r.intersect_batch(org, dr)
        g[f"isect_{name}__max_tris"] = np.int32(-1 if max_tris is None else max_tris)
        g[f"isect_{name}__org"], g[f"isect_{name}__dir"] = org, dr
        for k in ("t", "nrm", "nrm_raw", "obj", "into"):
            g[f"isect_{name}__{k}"] = a[k]
        # getSurfaceColor of the (textured) floor
        p = np.stack([rng.uniform(-25, 25, 600), np.full(600, -20.0), rng.uniform(-5, 45, 600)], -1)
        g[f"surf_{name}__pos"] = p
        g[f"surf_{name}__col"] = r.surface_color(3 if name == "c1_spheres" else 0, p)

    # ---- one triangle
    tri = np.array([-1, -1, 5, 2, -1, 6, 0, 3, 4], np.float64)
    d = rng.normal(size=(400, 3)) * [0.4, 0.4, 1.0]
    d[:, 2] = np.abs(d[:, 2])
    d /= np.linalg.norm(d, axis=1)[:, None]
    o = rng.uniform(-0.5, 0.5, (400, 3))
    hit, ln, nrm = r.triangle_intersect(tri, o, d)
    g["tri__tri9"], g["tri__org"], g["tri__dir"], g["tri__hit"], g["tri__t"], g["tri__nrm"] = tri, o, d, hit, ln, nrm

    # ---- texture (synthetic 5x7 texels on the floor, and on the two wall orientations)
    tex_rgb = rng.integers(1, 256, (5, 7, 3)).astype(np.uint8)
    r = ob.Ref()
    specs = [((0, 1, 0), (-21, 0, 0), 42, 40, True), ((1, 0, 0), (0, -20, 0), 40, 40, False), ((0, 0, 1), (-20, -20, 0), 40, 40, False)]
    for n, p, lx, ly, bump in specs:
        r.add_texture(tex_rgb, n, p, lx, ly, bump)
    g["tex__rgb"] = tex_rgb
    g["tex__specs"] = np.array([list(n) + list(p) + [lx, ly, float(bump)] for n, p, lx, ly, bump in specs])
    pts = rng.uniform(-25, 45, (900, 3))
    for t, axis in enumerate((1, 0, 2)):
        q = pts[300 * t:300 * (t + 1)].copy()
        q[:, axis] = rng.uniform(-0.02, 0.02, 300) + specs[t][1][axis]
        pts[300 * t:300 * (t + 1)] = q
        hit, col = r.texture_color(t, q)
        g[f"tex__hit{t}"], g[f"tex__col{t}"] = hit, col
    g["tex__pts"] = pts
    g["tex__height"] = np.array([[r.texture_height(0, i, j) for j in range(7)] for i in range(5)])

    # ---- eye + photon trace() on the bunny scene (reference image size is compiled in: 1024x768)
    s = preset("c2_bunny_chess")
    r = ob.Ref(s)
    W, H = r.image_size()
    r.htable_new(1000001)
    o2, d2 = camera_rays(W, H, PIX_STEP)
    hs, ws = np.meshgrid(np.arange(0, H, PIX_STEP), np.arange(0, W, PIX_STEP), indexing="ij")
    hs, ws = hs.ravel(), ws.ravel()
    for i in range(len(o2)):
        r.trace(o2[i], d2[i], (0, 0, 0), (1, 1, 1), True, int(ws[i]), int(hs[i]))
    hp = r.download_hitpoints()
    g["eye_c2__size"] = np.array([W, H, PIX_STEP], np.int32)
    for k in ("pos", "normal", "f", "r2", "hw", "key"):
        g[f"eye_c2__{k}"] = hp[k]
    r.seed(LIBC_SEED)
    po = np.stack([rng.uniform(-2, 2, N_PHOTON), np.full(N_PHOTON, 19.999), 20 + rng.uniform(-2, 2, N_PHOTON)], -1)
    pd = rng.normal(size=(N_PHOTON, 3))
    pd /= np.linalg.norm(pd, axis=1)[:, None]
    flux0 = 700.0 * (3.14159265358979 * 4.0)
    for i in range(N_PHOTON):
        r.trace(po[i], pd[i], (flux0,) * 3, (1, 1, 1), False)
    hp = r.download_hitpoints()
    g["photon_c2__seed"] = np.uint64(LIBC_SEED)
    g["photon_c2__org"], g["photon_c2__dir"], g["photon_c2__flux0"] = po, pd, np.float64(flux0)
    for k in ("flux", "r2", "n"):
        g[f"photon_c2__{k}"] = hp[k]

    # ---- misc scalar helpers
    abc = rng.uniform(-3, 3, (50, 3, 3))
    dets, invs, oks = [], [], []
    for m in abc:
        dd, ok, inv = r.det_inv(m[0], m[1], m[2])
        dets.append(dd); oks.append(ok); invs.append(inv)
    g["misc__abc"], g["misc__det"], g["misc__inv_ok"], g["misc__inv"] = abc, np.array(dets), np.array(oks), np.array(invs)
    x = np.concatenate([[0, 0.01, 0.1, 0.5, 1, 3, 100], rng.uniform(0, 4, 200)])
    g["misc__gamma_x"], g["misc__gamma"] = x, r.gamma_corr(x)
    s = preset("c1_spheres_bezier")
    r = ob.Ref(s)
    bid = len(s.objects) - 1
    us = np.linspace(0, 1, 21)
    g["bez__u"] = us
    g["bez__P"] = np.array([r.bezier_eval(bid, 0, (0, u, 0)) for u in us])
    g["bez__dP"] = np.array([r.bezier_eval(bid, 1, (0, u, 0)) for u in us])
    par = rng.uniform([15, 0, -3], [40, 1, 3], (40, 3))
    bo, bd = np.zeros(3), np.array([0.1, 0.1, 1.0])
    g["bez__par"], g["bez__org"], g["bez__dir"] = par, bo, bd
    g["bez__F"] = np.array([r.bezier_eval(bid, 2, p, bo, bd) for p in par])
    for c, what in enumerate((4, 5, 6)):  # Jacobian columns d/dt, d/du, d/dtheta (bezier.h:150-162)
        g[f"bez__J{c}"] = np.array([r.bezier_eval(bid, what, p, bo, bd) for p in par])
    g["bez__normal"] = np.array([r.bezier_eval(bid, 3, p, bo, bd) for p in par])

    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generates tests/golden/ref_mirror.npz: the mirror branch of trace() (main.cpp:129-134) as the UNMODIFIED reference
computes it (oracle/_ref/libcgref.so, `make -C oracle`). Run in the build container only:

    python tests/golden/make_golden_mirror.py

Scene `c1_mirror` = the spheres of main.cpp:288-290 + one mirror sphere inside the room + the chessboard floor + walls. None of the
reference's own scenes has a REACHABLE mirror made of deterministic primitives (its mirror sphere sits behind the back wall, the vase
is solved by a randomised Newton), so this fixture is what pins eye: adj*f*refl, origin + n*1e-4; photon: flux*f*refl.

  eye_mirror      trace(flag=true) on a pixel sub-grid of the compiled-in 1024x768 image -> hitpoints
  photon_mirror   trace(flag=false), seeded rand() stream, photons aimed at the mirror sphere -> flux / r2 / n per hitpoint
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cgraytracing_b200 import preset  # noqa: E402
from oracle import binding as ob  # noqa: E402
from tests.util import camera_rays  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_mirror.npz")
PIX_STEP, N_PHOTON, LIBC_SEED = 16, 4000, 777
MIRROR_C, MIRROR_R = np.array([-8.0, -13.0, 25.0]), 7.0


def photon_batch():
    """Half of the photons start on the light and are aimed into the mirror sphere's disc, the rest are isotropic."""
    rng = np.random.default_rng(4242)
    po = np.stack([rng.uniform(-2, 2, N_PHOTON), np.full(N_PHOTON, 19.999), 20 + rng.uniform(-2, 2, N_PHOTON)], -1)
    pd = rng.normal(size=(N_PHOTON, 3))
    tgt = MIRROR_C + rng.uniform(-0.65, 0.65, (N_PHOTON // 2, 3)) * MIRROR_R
    pd[: N_PHOTON // 2] = tgt - po[: N_PHOTON // 2]
    pd /= np.linalg.norm(pd, axis=1)[:, None]
    return po, pd


def main():
    assert ob.have_ref(), "build oracle/_ref first: make -C oracle"
    g = {}
    s = preset("c1_mirror")
    r = ob.Ref(s)
    W, H = r.image_size()
    r.htable_new(1000001)
    o2, d2 = camera_rays(W, H, PIX_STEP)
    hs, ws = np.meshgrid(np.arange(0, H, PIX_STEP), np.arange(0, W, PIX_STEP), indexing="ij")
    hs, ws = hs.ravel(), ws.ravel()
    first = r.intersect_batch(o2, d2)
    g["eye_mirror__primary_obj"] = first["obj"]
    for i in range(len(o2)):
        r.trace(o2[i], d2[i], (0, 0, 0), (1, 1, 1), True, int(ws[i]), int(hs[i]))
    hp = r.download_hitpoints()
    g["eye_mirror__size"] = np.array([W, H, PIX_STEP], np.int32)
    for k in ("pos", "normal", "f", "r2", "hw", "key"):
        g[f"eye_mirror__{k}"] = hp[k]
    r.seed(LIBC_SEED)
    po, pd = photon_batch()
    flux0 = 700.0 * (3.14159265358979 * 4.0)
    for i in range(N_PHOTON):
        r.trace(po[i], pd[i], (flux0,) * 3, (1, 1, 1), False)
    hp = r.download_hitpoints()
    g["photon_mirror__seed"] = np.uint64(LIBC_SEED)
    g["photon_mirror__org"], g["photon_mirror__dir"], g["photon_mirror__flux0"] = po, pd, np.float64(flux0)
    for k in ("flux", "r2", "n"):
        g[f"photon_mirror__{k}"] = hp[k]
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", int((first["obj"] == 3).sum()), "primary rays on the mirror;",
          int(hp["n"].sum()), "deposits")


if __name__ == "__main__":
    main()

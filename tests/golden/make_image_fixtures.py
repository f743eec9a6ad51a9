#!/usr/bin/env python
"""Generates tests/golden/ref_images.npz: 8-bit pictures of the REFERENCE ALGORITHM at a fixed photon budget, two seeds per scene —
the yard-stick of the image gate (BASELINE.md section 5 / north star: RMSE(gpu, ref) <= 1.05 x RMSE(ref seed A, ref seed B),
8x8 box-filtered <= 1.5 x, channel means within 1 %). Run in the build container only (about 10 minutes on 4 cores):

    python tests/golden/make_image_fixtures.py

"Reference algorithm" = the oracle in the mode that is pinned BIT-EXACT to the compiled reference (tests/test_oracle_vs_ref.py,
tests/test_oracle_golden.py, tests/test_mirror.py): per-photon update U1 (main.cpp:119-122), rejection samplers on a rand()-style
31-bit stream (sampling.h), hit-count-parity inside/outside rule (objects.h:318-332), Bezier by 10 random Newton restarts
(bezier.h:233-249), one thread (the reference's 8 racing threads are a non-deterministic interleaving of the same per-photon rule),
tone map + gamma + flip of main.cpp:403-411. libcgref.so itself cannot produce them: image size, photon count and scene are compile-time
literals of main.cpp and its seed is time(0).

  c3  glass dragon + chessboard floor at 1024 x 768 (the reference's compiled-in size), 8,000,000 photons
  c1  spheres + Bezier vase + chessboard floor at 512 x 512 (BASELINE config 1), 2,000,000 photons
"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_images.npz")
JOBS = {  # name: (preset, W, H, photons)
    "c3": ("c3_dragon_glass", 1024, 768, 8_000_000),
    "c1": ("c1_spheres_bezier", 512, 512, 2_000_000),
}
SEEDS = {"A": 11, "B": 29}


def run(job):
    name, tag = job
    from cgraytracing_b200 import RenderConfig, preset
    from oracle import binding as ob

    pre, W, H, photons = JOBS[name]
    o = ob.Oracle(preset(pre), RenderConfig(width=W, height=H, update_mode=0, into_rule=0))
    t0 = time.time()
    o.set_libc_rng(1, SEEDS[tag] * 7919)  # the eye pass of the Bezier scene draws from the stream too (bezier.h:236-239)
    o.eye_pass()
    o.set_libc_rng(1, SEEDS[tag])
    o.photon_pass(0, photons)
    img = ob.tonemap_flip(o.gather_image(float(photons)))
    print(name, tag, "done in %.0f s, mean %.3f" % (time.time() - t0, img.mean()), flush=True)
    return name, tag, img


def main():
    jobs = [(n, t) for n in JOBS for t in SEEDS]
    g = {}
    with ProcessPoolExecutor(max_workers=4) as ex:
        for name, tag, img in ex.map(run, jobs):
            g[f"{name}__{tag}"] = img
    for name, (pre, W, H, photons) in JOBS.items():
        g[f"{name}__meta"] = np.array([W, H, photons], np.int64)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()

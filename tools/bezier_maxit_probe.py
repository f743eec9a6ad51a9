"""c1 render with the library named by CGRT_LIB: image saved for comparison, timing of the photon pass (dev tool)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgraytracing_b200 import Context, RenderConfig, preset
tag = sys.argv[1]
s = preset("c1_spheres_bezier")
cfg = RenderConfig(width=512, height=512)
P = 1 << 20
with Context(0) as g:
    g.set_config(cfg, accum_mode=0); s.build_into(g); g.commit()
    g.eye_pass(); g.build_grid()
    hp = g.download_hitpoints()
    g.synchronize(); t0 = time.time()
    for r in range(10):
        g.photon_pass(r * P, P); g.round_update()
    g.synchronize(); t1 = time.time()
    img = g.gather_image(10.0 * P)
    c = g.counters()
np.savez(f"gpurun_out/bez_{tag}.npz", img=img, pos=hp["pos"], key=hp["key"])
print(tag, "hitpoints", c["hitpoints"], "deposits", c["deposits"], "segments", c["photon_segments"], "ms/round", 100 * (t1 - t0), "mean", img.mean())

"""Eye-pass timing per workload, repeated so that the memory pool is warm (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgraytracing_b200 import Context, RenderConfig, preset
for name, W, H, dof, ns in (("c2_bunny_chess", 1024, 1024, 0, 1), ("c3_dragon_glass", 1024, 1024, 0, 1), ("c4_bump_dof", 1920, 1080, 1, 4), ("c1_spheres_bezier", 512, 512, 0, 1)):
    s = preset(name)
    for it in range(3):
        with Context(0, s, RenderConfig(width=W, height=H, use_dof=dof, num_of_samples=ns)) as g:
            t0 = time.time(); g.eye_pass(); t1 = time.time(); g.build_grid(); t2 = time.time()
            c = g.counters(); tm = g.timings()
            print(f"{name:18s} it{it}: eye {1e3*(t1-t0):7.2f} ms wall ({tm['eye']:7.2f} ms device)  grid {1e3*(t2-t1):6.2f} ms  segments {c['eye_segments']}  hitpoints {c['hitpoints']}  -> {c['eye_segments']/(t1-t0)/1e6:7.1f} M eye rays/s")

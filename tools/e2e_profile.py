"""Stage-by-stage wall-clock of one render() through the C ABI (dev tool)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgraytracing_b200 import Context, RenderConfig, preset

scene = preset("c3_dragon_glass")
cfg = RenderConfig(width=1024, height=1024)
P = 16 * 1024 * 1024
for it in range(3):
    t = [time.time()]
    def lap(name):
        t.append(time.time()); print(f"  {name:14s} {1e3*(t[-1]-t[-2]):9.2f} ms")
    print("iteration", it)
    g = Context(0); lap("create")
    g.set_config(cfg); scene.build_into(g); lap("add objects")
    g.commit(); lap("commit")
    g.eye_pass(); lap("eye")
    g.build_grid(); lap("grid")
    g.photon_pass(it * P, P); lap("photon")
    g.round_update(); lap("update")
    img, rgb8 = g.gather_image(float(P), want_rgb8=True); lap("gather")
    print("  timings", {k: round(v, 2) for k, v in g.timings().items()})
    g.close(); lap("destroy")
    print(f"  total {1e3*(t[-1]-t[0]):.1f} ms")

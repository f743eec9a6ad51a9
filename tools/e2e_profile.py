"""Stage-by-stage wall-clock of whole render() calls through the C ABI (dev tool)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgraytracing_b200 import Context, RenderConfig, preset

scene = preset("c3_dragon_glass")
cfg = RenderConfig(width=1024, height=1024)
P = 16 * 1024 * 1024
for it, rounds in enumerate((1, 1, 50, 50)):
    t = [time.time()]
    def lap(name):
        t.append(time.time()); print(f"  {name:14s} {1e3*(t[-1]-t[-2]):9.2f} ms")
    print("iteration", it, "rounds", rounds)
    g = Context(0); lap("create")
    g.set_config(cfg, accum_mode=1); scene.build_into(g); lap("add objects")
    g.commit(); lap("commit")
    g.eye_pass(); lap("eye")
    g.build_grid(); lap("grid")
    for r in range(rounds):
        g.photon_pass(r * P, P); g.round_update()
    lap("enqueue rounds")
    g.synchronize(); lap("rounds done")
    img, rgb8 = g.gather_image(float(P) * rounds, want_rgb8=True); lap("gather")
    g.close(); lap("destroy")
    print(f"  total {1e3*(t[-1]-t[0]):.1f} ms")

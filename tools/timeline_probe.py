import sys, os
sys.path.insert(0, '/root/repo')
from cgraytracing_b200 import Context, RenderConfig, preset
P = 16 << 20
with Context(0, preset("c3_dragon_glass"), RenderConfig(width=1024, height=1024)) as g:
    g.set_config(RenderConfig(width=1024, height=1024), accum_mode=1) if False else None
    g.eye_pass(); g.build_grid()
    for r in range(3): g.photon_pass(r*P, P); g.round_update()
    g.synchronize()
    print("---- measured", file=sys.stderr)
    for r in range(3, 8): g.photon_pass(r*P, P); g.round_update()
    g.synchronize()

"""Per-round device time over a long render (dev tool): do later rounds slow down (clocks, shrinking radii)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgraytracing_b200 import Context, RenderConfig, preset
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
P = 16 << 20
with Context(0, preset("c3_dragon_glass"), RenderConfig(width=1024, height=1024)) as g:
    g.set_config(RenderConfig(width=1024, height=1024), accum_mode=1) if False else None
    g.eye_pass(); g.build_grid()
    st = torch.cuda.ExternalStream(g.stream())
    for blk in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        c0 = g.counters()
        for r in range(10):
            g.photon_pass((blk * 10 + r) * P, P); g.round_update()
        e1.record(st); g.synchronize(); torch.cuda.synchronize()
        c1 = g.counters()
        clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000
        print(f"rounds {blk*10:3d}-{blk*10+9:3d}: {e0.elapsed_time(e1)/10:7.2f} ms/round  deposits/hit {(c1['deposits']-c0['deposits'])/(c1['diffuse_hits']-c0['diffuse_hits']):.2f}  cand/hit {(c1['candidates']-c0['candidates'])/(c1['diffuse_hits']-c0['diffuse_hits']):.1f}  sm clock {clk} MHz  power {pw:.0f} W")

"""Small renders of every BASELINE scene family through the C ABI with fenced device buffers (CGRT_GUARD=1): every buffer of a
context sits between two 4 KiB fences and the probe fails when a kernel has written into one. compute-sanitizer is not available
on the GPU pool; this catches the out-of-bounds writes it would. tests/test_gpu_parity.py runs it.

  CGRT_GUARD=1 python tools/sanitize_probe.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgraytracing_b200 import Context, RenderConfig, preset

CASES = (("c1_spheres_bezier", dict(width=48, height=48), 20000),
         ("c2_bunny_chess", dict(width=64, height=64), 30000),
         ("c3_dragon_glass", dict(width=64, height=64), 30000),
         ("c4_bump_dof", dict(width=64, height=36, use_dof=1, num_of_samples=2), 30000))
damaged = 0
for name, kw, P in CASES:
    s = preset(name)
    for accum, overlap in ((0, 0), (1, 0), (1, 1)):
        with Context(0) as g:
            g.set_config(RenderConfig(**kw), accum_mode=accum)
            g.set_overlap(bool(overlap))
            s.build_into(g); g.commit()
            g.eye_pass(); g.build_grid()
            for r in range(2):
                g.photon_pass(r * P, P); g.round_update()
            img, rgb8 = g.gather_image(2.0 * P, want_rgb8=True)
            c = g.counters()
            rng = np.random.default_rng(1)
            o = rng.uniform(-5, 5, (256, 3)); o[:, 2] -= 20
            d = rng.normal(size=(256, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
            g.intersect_batch(o, d)
            damaged += g.check_guards()
            print(f"{name:18s} accum {accum} overlap {overlap}: hitpoints {c['hitpoints']} deposits {c['deposits']} mean {img.mean():.4f} finite {bool(np.isfinite(img).all())}", flush=True)
    with Context(0) as g:
        imgs = np.random.default_rng(2).integers(0, 256, (9, 16, 16, 3), dtype=np.uint8)
        g.average_u8(imgs)
        damaged += g.check_guards()
if os.environ.get("CGRT_GUARD", "0") != "0":
    # self-test of the fences: a deliberate 8-byte write just below the hitpoint record buffer must be reported
    import torch
    from cgraytracing_b200.distributed import _DevArray
    with Context(0) as g:
        g.set_config(RenderConfig(width=32, height=32)); preset("c1_spheres_bezier").build_into(g); g.commit(); g.eye_pass()
        ptr, n = g.export_hitpoints_dev()
        assert g.check_guards() == 0
        torch.as_tensor(_DevArray(ptr - 8, 1, "<f8"), device="cuda:0").fill_(1.0)
        torch.cuda.synchronize()
        seen = g.check_guards()
        print("fence self-test: deliberate 8-byte underflow reported as", seen, "damaged bytes")
        assert seen == 8, seen
print("probe done; guard mode", os.environ.get("CGRT_GUARD", "0"), "damaged fence bytes", damaged)
sys.exit(1 if damaged else 0)

"""Image parity at scale (documented experiment, not part of the test suite).

  python tools/image_parity_experiment.py ref <seed> <out.npy>      # CPU: the reference algorithm (oracle U1, libc-style stream), one thread
  python tools/image_parity_experiment.py gpu <rounds> <out.npy>    # GPU: libcgrt.so, U2, Philox
  python tools/image_parity_experiment.py report a.npy b.npy gpu.npy [more gpu.npy]

Scene c3_dragon_glass at 1024x768 (the reference's own size), PHOTONS photons in total. Writes the 8-bit tone-mapped picture."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgraytracing_b200 import RenderConfig, preset  # noqa: E402

W, H = 1024, 768
PHOTONS = int(os.environ.get("CGRT_EXP_PHOTONS", 8_000_000))
SCENE = os.environ.get("CGRT_EXP_SCENE", "c3_dragon_glass")


def rmse8(a, b):
    return float(np.sqrt(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))


def box(a, k=8):
    h, w = a.shape[:2]
    return a[: h // k * k, : w // k * k].astype(np.float64).reshape(h // k, k, w // k, k, 3).mean((1, 3))


def main():
    mode = sys.argv[1]
    if mode == "ref":
        from oracle import binding as ob

        o = ob.Oracle(preset(SCENE), RenderConfig(width=W, height=H, update_mode=0, into_rule=0))
        t0 = time.time()
        o.eye_pass()
        o.set_libc_rng(1, int(sys.argv[2]))
        o.photon_pass(0, PHOTONS)
        np.save(sys.argv[3], ob.tonemap_flip(o.gather_image(float(PHOTONS))))
        print("ref seed", sys.argv[2], "done in", time.time() - t0, "s")
    elif mode == "gpu":
        from cgraytracing_b200 import Context

        rounds = int(sys.argv[2])
        per = PHOTONS // rounds
        with Context(0, preset(SCENE), RenderConfig(width=W, height=H)) as g:
            g.eye_pass(); g.build_grid()
            for r in range(rounds):
                g.photon_pass(r * per, per); g.round_update()
            img, rgb8 = g.gather_image(float(per * rounds), want_rgb8=True)
        np.save(sys.argv[3], rgb8)
        print("gpu rounds", rounds, "mean", rgb8.mean())
    else:
        a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
        print(f"reference seed A vs seed B: RMSE {rmse8(a, b):.3f}  (8x8 box {rmse8(box(a), box(b)):.3f})  means {a.mean():.3f} {b.mean():.3f}")
        for f in sys.argv[4:]:
            g = np.load(f)
            print(f"{os.path.basename(f)} vs A: RMSE {rmse8(g, a):.3f} (box {rmse8(box(g), box(a)):.3f}); vs B: {rmse8(g, b):.3f} (box {rmse8(box(g), box(b)):.3f}); "
                  f"mean {g.mean():.3f}; channel means {[round(float(g[..., c].mean() / a[..., c].mean()), 4) for c in range(3)]}")


if __name__ == "__main__":
    main()

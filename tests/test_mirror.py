"""The mirror material of trace() (main.cpp:129-134: d' = d - 2(n.d)n, origin + n*1e-4; eye: adj*f*refl, photon: flux*f*refl),
pinned on a scene whose mirror is REACHABLE and deterministic (`c1_mirror`: the spheres of main.cpp:288-290 plus one mirror sphere
inside the room; the reference's own mirror sphere sits behind the back wall and its vase is a randomised Newton solve).

  CPU        oracle == tests/golden/ref_mirror.npz (outputs of the unmodified reference, tests/golden/make_golden_mirror.py), bit-exact
  CPU, ref   oracle == the compiled reference live, on more rays / photons
  GPU        eye hitpoints == the reference's golden hitpoints bit for bit; eye + photon rounds == the oracle (Philox) bit for bit
"""
import os

import numpy as np
import pytest

from cgraytracing_b200 import RenderConfig, preset
from tests.util import assert_flux_close, camera_rays

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_mirror.npz")
MIRROR_ID = 3  # the fourth sphere of c1_mirror


@pytest.fixture(scope="module")
def G():
    return np.load(GOLDEN)


def _through_mirror(f):
    """Hitpoints on a grey (0.15, 0.15, 0.15) wall seen through exactly one bounce off the mirror: f*adj = 0.15 * (0.9, 0.8, 0.7) * 0.8."""
    want = 0.15 * np.array([0.9, 0.8, 0.7]) * 0.8
    return np.all(np.abs(f - want) <= 1e-12, axis=1)


def _grid(W, H, step):
    hs, ws = np.meshgrid(np.arange(0, H, step), np.arange(0, W, step), indexing="ij")
    return hs.ravel(), ws.ravel()


def test_mirror_scene_has_a_reachable_mirror(G):
    s = preset("c1_mirror")
    o = s.objects[MIRROR_ID]
    assert o["kind"] == "sphere" and o["refl"] >= 1e-4 and o["transp"] < 1e-4  # main.cpp:129: mirror branch
    assert (G["eye_mirror__primary_obj"] == MIRROR_ID).sum() > 100
    # reflected eye paths end on diffuse surfaces with f*adj = f_diffuse * (f_mirror * refl)^k: such hitpoints exist in the golden
    assert _through_mirror(G["eye_mirror__f"]).sum() >= 30


def test_oracle_mirror_trace_vs_reference_golden(oracle_lib, G):
    W, H, step = (int(v) for v in G["eye_mirror__size"])
    o = oracle_lib.Oracle(preset("c1_mirror"), RenderConfig(width=W, height=H, update_mode=0, into_rule=0))
    org, dr = camera_rays(W, H, step)
    assert np.array_equal(o.intersect_batch(org, dr)["obj"], G["eye_mirror__primary_obj"])
    hs, ws = _grid(W, H, step)
    for i, (h, w) in enumerate(zip(hs, ws)):
        o.trace(org[i], dr[i], (0, 0, 0), (1, 1, 1), True, int(w), int(h), path=int(h) * W + int(w))
    hp = o.download_hitpoints()
    assert len(hp["pos"]) == len(G["eye_mirror__pos"]) > 2000
    for k in ("key", "hw", "pos", "normal", "f", "r2"):
        assert np.array_equal(hp[k], G[f"eye_mirror__{k}"]), k
    o.set_libc_rng(1, int(G["photon_mirror__seed"]))
    f0 = float(G["photon_mirror__flux0"])
    for po, pd in zip(G["photon_mirror__org"], G["photon_mirror__dir"]):
        o.trace(po, pd, (f0, f0, f0), (1, 1, 1), False)
    hp = o.download_hitpoints()
    assert hp["n"].sum() == G["photon_mirror__n"].sum() > 500
    for k in ("n", "r2", "flux"):
        assert np.array_equal(hp[k], G[f"photon_mirror__{k}"]), k


@pytest.mark.ref
def test_oracle_mirror_trace_vs_reference_live(oracle_lib):
    """More pixels and more photons than the fixture, against libcgref.so itself."""
    ob = oracle_lib
    s = preset("c1_mirror")
    r = ob.Ref(s)
    W, H = r.image_size()
    o = ob.Oracle(s, RenderConfig(width=W, height=H, update_mode=0, into_rule=0))
    r.htable_new(1000001)
    org, dr = camera_rays(W, H, 8)
    hs, ws = _grid(W, H, 8)
    for i, (h, w) in enumerate(zip(hs, ws)):
        r.trace(org[i], dr[i], (0, 0, 0), (1, 1, 1), True, int(w), int(h))
        o.trace(org[i], dr[i], (0, 0, 0), (1, 1, 1), True, int(w), int(h), path=i)
    r.seed(31)
    o.set_libc_rng(1, 31)
    rng = np.random.default_rng(8)
    n = 6000
    po = np.stack([rng.uniform(-2, 2, n), np.full(n, 19.999), 20 + rng.uniform(-2, 2, n)], -1)
    tgt = np.array([-8.0, -13.0, 25.0]) + rng.uniform(-5, 5, (n, 3))
    pd = tgt - po
    pd /= np.linalg.norm(pd, axis=1)[:, None]
    f0 = 700.0 * (3.14159265358979 * 4.0)
    for i in range(n):
        r.trace(po[i], pd[i], (f0,) * 3, (1, 1, 1), False)
        o.trace(po[i], pd[i], (f0,) * 3, (1, 1, 1), False)
    a, b = r.download_hitpoints(), o.download_hitpoints()
    assert a["n"].sum() > 1000
    for k in a:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.gpu
def test_gpu_mirror_eye_hitpoints_equal_the_reference_golden(gpu, G):
    """The GPU eye pass over the whole 1024x768 image; the hitpoints of the golden's pixel sub-grid are the reference's, bit for bit."""
    W, H, step = (int(v) for v in G["eye_mirror__size"])
    with gpu.Context(0, gpu.preset("c1_mirror"), gpu.RenderConfig(width=W, height=H)) as g:
        g.eye_pass(); g.build_grid()
        a = g.download_hitpoints()
    on_grid = (a["hw"][:, 0] % step == 0) & (a["hw"][:, 1] % step == 0)
    assert on_grid.sum() == len(G["eye_mirror__pos"])
    # canonical order restricted to a pixel subset is still (bucket, creation order): the same sequence as the golden's
    for k in ("key", "hw", "pos", "normal", "f", "r2"):
        assert np.array_equal(a[k][on_grid], G[f"eye_mirror__{k}"]), k


@pytest.mark.gpu
@pytest.mark.parametrize("accum", [0, 1])
def test_gpu_mirror_rounds_equal_the_oracle(gpu, oracle_lib, accum):
    """Eye pass + two photon rounds on the mirror scene: hitpoints bit-exact, accepted-photon counts equal, flux to accumulation order."""
    W, H, NPH = 256, 192, 60000
    s = gpu.preset("c1_mirror")
    cfg = gpu.RenderConfig(width=W, height=H, into_rule=1, update_mode=1)
    o = oracle_lib.Oracle(s, cfg)
    o.eye_pass(nthreads=o.max_threads())
    with gpu.Context(0) as g:
        g.set_config(cfg, accum_mode=accum)
        s.build_into(g); g.commit()
        g.eye_pass(); g.build_grid()
        a, b = g.download_hitpoints(), o.download_hitpoints()
        assert len(a["pos"]) == len(b["pos"]) > W * H
        for k in ("key", "hw", "pos", "normal", "f", "r2"):
            assert np.array_equal(a[k], b[k]), k
        for rnd in range(2):
            g.photon_pass(rnd * NPH, NPH); o.photon_pass(rnd * NPH, NPH, o.max_threads())
            df, m = g.download_accum(); odf, om = o.download_accum()
            assert np.array_equal(m.astype(np.int64), om.astype(np.int64))
            assert_flux_close(df, odf, om, accum)
            g.round_update(); o.round_update()
        gc, oc = g.counters(), o.counters()
        for k in ("photon_segments", "diffuse_hits", "deposits"):
            assert gc[k] == oc[k], k
        a, b = g.download_hitpoints(), o.download_hitpoints()
        assert np.array_equal(a["n"], b["n"])
        # hitpoints seen through the mirror receive photons too
        tinted = _through_mirror(b["f"])
        assert tinted.sum() > 500 and b["n"][tinted].sum() > 100
        assert np.allclose(g.gather_image(2.0 * NPH), o.gather_image(2.0 * NPH), rtol=1e-9 if accum == 0 else 1e-4, atol=1e-12 if accum == 0 else 1e-6)

"""trace() as a callable entry (main.cpp:42) behind the C ABI: cgrt_trace over caller-made rays.
Eye rays must leave exactly the hitpoints cgrt_eye_pass leaves for the same camera rays; photon rays must deposit exactly what the oracle's
trace() deposits for the same rays and random-number indices."""
import numpy as np
import pytest

from tests.util import assert_flux_close, camera_rays

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,max_tris", [("c2_bunny_chess", None), ("c3_dragon_glass", 20000), ("c1_mirror", None)])
def test_eye_rays_through_trace_equal_the_eye_pass(gpu, name, max_tris):
    W, H = 128, 96
    cfg = gpu.RenderConfig(width=W, height=H, into_rule=1, update_mode=1)
    s = gpu.preset(name, max_tris=max_tris)
    org, dir = camera_rays(W, H)
    hs, ws = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    with gpu.Context(0, s, cfg) as a, gpu.Context(0, s, cfg) as b:
        a.eye_pass(); a.build_grid()
        # in two calls and out of order: the canonical order does not depend on who traced what when
        half = len(org) // 2
        for sl in (slice(half, None), slice(0, half)):
            b.trace(org[sl], dir[sl], np.ones((len(org[sl]), 3)), True, 0, ws.reshape(-1)[sl], hs.reshape(-1)[sl])
        b.build_grid()
        ha, hb = a.download_hitpoints(), b.download_hitpoints()
        assert len(ha["pos"]) == len(hb["pos"]) > 0
        for k in ("pos", "normal", "f", "key", "seq", "hw", "r2"):
            assert np.array_equal(ha[k], hb[k]), k
        assert a.counters()["eye_segments"] == b.counters()["eye_segments"]


@pytest.mark.parametrize("name,max_tris", [("c2_bunny_chess", None), ("c3_dragon_glass", 20000)])
def test_photon_rays_through_trace_deposit_like_the_oracle(gpu, oracle_lib, name, max_tris):
    W, H, N, FIRST = 96, 72, 1500, 12345
    cfg = gpu.RenderConfig(width=W, height=H, into_rule=1, update_mode=1)
    s = gpu.preset(name, max_tris=max_tris)
    rng = np.random.default_rng(11)
    d = rng.normal(size=(N, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    org = np.array([0.0, 19.999, 20.0]) + np.stack([rng.uniform(-2, 2, N), np.zeros(N), rng.uniform(-2, 2, N)], -1)
    flux = np.tile(np.array([[700.0, 700.0, 700.0]]) * (3.14159265358979 * 4.0), (N, 1))
    with gpu.Context(0, s, cfg) as g:
        g.eye_pass(); g.build_grid()
        g.trace(org, d, flux, False, 0, first_index=FIRST)
        df, m = g.download_accum()
        kg = g.counters()
    o = oracle_lib.Oracle(s, cfg)
    o.eye_pass()
    for k in range(N):
        o.trace(org[k], d[k], flux[k], (1, 1, 1), False, path=FIRST + k)
    odf, om = o.download_accum()
    ko = o.counters()
    assert kg["photon_segments"] == ko["photon_segments"] and kg["diffuse_hits"] == ko["diffuse_hits"]
    assert om.sum() > 0 and np.array_equal(m.astype(np.int64), om.astype(np.int64))
    assert_flux_close(df, odf, om, 0)
    # depth argument: a ray that starts at MAX_DEPTH traces nothing (main.cpp:46)
    with gpu.Context(0, s, cfg) as g:
        g.eye_pass(); g.build_grid()
        g.trace(org[:10], d[:10], flux[:10], False, cfg.max_depth)
        assert g.download_accum()[1].sum() == 0

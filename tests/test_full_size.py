"""GPU vs CPU oracle at the BENCHMARKED sizes (BASELINE.json configs 3, 4, 5) — not on truncated meshes or thumbnails:

  c3  whole 100,000-triangle glass dragon, 1024 x 1024, 2 Mi photons             (the bench.py headline workload)
  c4  stone displacement floor (146,744 triangles) + Mesh000 water, thin-lens camera, 1920 x 1080 x 4 samples, 1 Mi photons
  c5  the c3 scene at 4096 x 4096 with hashsize 16777259, 1 Mi photons

For each: the eye pass's hitpoints, cell keys and canonical order are the oracle's bit for bit; after the photon pass the accepted-photon
count of EVERY hitpoint equals the oracle's, the flux accumulators agree to the accumulation-order rounding of the accumulator type
(fp64 atomics: 1e-9 relative; float `red` accumulators: the worst-case bound of summing m positive terms in float in any order,
(m + 8) * 2^-24 relative, m = that hitpoint's accepted photons — hitpoints under the light collect 10^4 deposits per round), and the
segment / hit / deposit counters are equal. The oracle runs its eye pass and photon loop on all host threads (Philox streams make it order-independent); it is
pinned bit-exact to the compiled reference in tests/test_oracle_vs_ref.py, tests/test_oracle_golden.py and tests/test_mirror.py."""
import numpy as np
import pytest

from tests.util import assert_flux_close

pytestmark = pytest.mark.gpu

CASES = {
    #            preset             W     H     hashsize  dof samples photons
    "c3": ("c3_dragon_glass", 1024, 1024, 1000001, 0, 1, 1 << 21),
    "c4": ("c4_bump_dof", 1920, 1080, 1000001, 1, 4, 1 << 20),
    "c5": ("c3_dragon_glass", 4096, 4096, 16777259, 0, 1, 1 << 20),
}

_cache = {}


def oracle_side(ob, gpu, case):
    """Eye pass + one photon pass of the oracle, once per case (shared by the two accumulator modes)."""
    if case not in _cache:
        _cache.clear()  # one case in memory at a time (c5: 18 M hitpoints)
        name, W, H, hs, dof, smp, N = CASES[case]
        s = gpu.preset(name)
        cfg = gpu.RenderConfig(width=W, height=H, hashsize=hs, use_dof=dof, num_of_samples=smp, into_rule=1, update_mode=1)
        o = ob.Oracle(s, cfg)
        nt = o.max_threads()
        o.eye_pass(nthreads=nt)
        hp = o.download_hitpoints(fields=("key", "hw", "pos", "normal", "f"))
        seq = o.download_hitpoints(fields=("path", "code"))
        hp["seq"] = (seq["path"] * 16 + seq["code"]).astype(np.uint32)
        o.photon_pass(0, N, nt)
        df, m = o.download_accum()
        ctr = o.counters()
        o.close()
        _cache[case] = (s, cfg, hp, df, m, ctr)
    return _cache[case]


@pytest.mark.parametrize("case,accum", [("c3", 0), ("c3", 1), ("c4", 0), ("c4", 1), ("c5", 0), ("c5", 1)])
def test_full_size_equals_oracle(gpu, oracle_lib, case, accum):
    s, cfg, ohp, odf, om, octr = oracle_side(oracle_lib, gpu, case)
    N = CASES[case][6]
    with gpu.Context(0) as g:
        g.set_config(cfg, accum_mode=accum)
        s.build_into(g); g.commit()
        g.eye_pass(); g.build_grid()
        n = g.num_hitpoints()
        assert n == len(ohp["key"]) >= cfg.width * cfg.height * cfg.num_of_samples * 0.95
        for k in ("key", "seq", "hw", "pos", "normal", "f"):  # one field at a time: c5 holds 18 M hitpoints
            a = g.download_hitpoints(fields=(k,))[k]
            assert np.array_equal(a, ohp[k]), k
            del a
        g.photon_pass(0, N)
        df, m = g.download_accum()
        gc = g.counters()
    assert gc["eye_segments"] == octr["eye_segments"]
    for k in ("photon_segments", "diffuse_hits", "deposits"):
        assert gc[k] == octr[k], k
    assert np.array_equal(m.astype(np.int64), om.astype(np.int64))   # accepted photons per hitpoint: exact in both accumulator types
    assert int(m.sum()) == gc["deposits"] > N
    assert_flux_close(df, odf, om, accum)

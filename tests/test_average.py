"""average.cpp: the reference combines separately rendered 8-bit pictures as sum_k (img_k / 9). The published
result/depth.png is bit-exactly that average of result/t1..t9.png; a crop of those ten pictures is committed as a fixture
(tests/golden/average_crop.npz, cut out with PIL in the build container)."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "average_crop.npz")
REF_RESULT = "/root/reference/result"


def test_oracle_average_reproduces_published_depth_png_crop(oracle_lib):
    G = np.load(GOLDEN)
    out = oracle_lib.average_u8(list(G["runs"]))
    assert np.array_equal(out, G["average"])
    # numpy restatement of the same arithmetic
    assert np.array_equal(out, sum((r // 9).astype(np.uint8) for r in G["runs"]).astype(np.uint8))


def test_oracle_average_full_images():
    PIL = pytest.importorskip("PIL.Image")
    if not os.path.isdir(REF_RESULT):
        pytest.skip("reference checkout not present")
    from oracle import binding as ob

    runs = [np.array(PIL.open(f"{REF_RESULT}/t{i}.png").convert("RGB")) for i in range(1, 10)]
    depth = np.array(PIL.open(f"{REF_RESULT}/depth.png").convert("RGB"))
    assert np.array_equal(ob.average_u8(runs), depth)


@pytest.mark.gpu
def test_gpu_average_modes(gpu, oracle_lib):
    G = np.load(GOLDEN)
    rng = np.random.default_rng(0)
    with gpu.Context(0) as g:
        assert np.array_equal(g.average_u8(list(G["runs"])), G["average"])
        for n in (1, 2, 7, 255):
            imgs = [rng.integers(0, 256, (37, 53, 3)).astype(np.uint8) for _ in range(n)]
            assert np.array_equal(g.average_u8(imgs), oracle_lib.average_u8(imgs))
        rad = [rng.uniform(0, 3, (40, 30, 3)) for _ in range(5)]
        mean, rgb8 = g.average_f64(rad, want_rgb8=True)
        want = (rad[0] + rad[1] + rad[2] + rad[3] + rad[4]) / 5.0
        assert np.array_equal(mean, want)
        assert np.array_equal(rgb8.ravel(), oracle_lib.gamma_corr(want).astype(np.uint8))

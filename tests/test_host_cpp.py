"""The host-side C++ mirror of the reference's scene/render API (cgraytracing_b200/host/cgrt_host.hpp) on top of the C ABI.
CPU part: it builds with plain g++, its mesh loader reproduces the reference's three text formats, and without a GPU the
render fails loudly. GPU part: a render through the C++ classes equals the same render through the Python binding bit for bit."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "cgraytracing_b200", "host")


@pytest.fixture(scope="module")
def host_bins():
    from cgraytracing_b200 import build

    build.build()
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    subprocess.check_call(["make", "-C", HOST, "-s"], env=env)
    return os.path.join(HOST, "example_main"), os.path.join(HOST, "dump_mesh")


def test_cpp_mesh_loader_matches_reference_loader(host_bins, oracle_lib, tmp_path):
    """oracle.load_mesh_text is pinned to the compiled reference's loader (tests/test_oracle_vs_ref.py)."""
    _, dump = host_bins
    files = {
        0: "begin\nvertex 0 0 0\nvertex 1 0 0.5\nvertex 0 1 -2\nend\n\nbegin\nvertex 1.25 1 1\nvertex 2 1e-3 1\nvertex 1 3 1.25\nend\n\n",
        1: "4\nv  0 0 0\nv  1 0 0.1\nv  0 1 0\nv  0 0.3 1\n2\nf 1 2 3 \nf 1 3 4 \n",
        2: "4\nv 0.5 0 0\nv 1 0 0\nv 0 1 0.7\nv 0 0 1\n2\nf 1/1/1 2/2/2 3/3/3 \nf 1/1/1 3/3/3 4/4/4 \n",
    }
    for typ, text in files.items():
        f = tmp_path / f"m{typ}.txt"
        f.write_text(text)
        out = subprocess.check_output([dump, str(f), str(typ), "2.5", "1", "-2", "3"], text=True)
        got = np.array([[float(v) for v in line.split()] for line in out.strip().splitlines()])
        want = oracle_lib.load_mesh_text(str(f), typ, 2.5, (1, -2, 3))
        assert got.shape == want.shape == (2, 9)
        assert np.array_equal(got, want), typ


def test_cpp_render_fails_loudly_without_gpu(host_bins, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    exe, _ = host_bins
    p = subprocess.run([exe, "bunny", "32", "24", "100", "1", str(tmp_path / "o.ppm"), os.path.join(ROOT, "cgraytracing_b200", "assets")],
                       capture_output=True, text=True)
    assert p.returncode == 3 and "no CPU fallback" in p.stderr
    assert not (tmp_path / "o.ppm").exists()


@pytest.mark.gpu
@pytest.mark.parametrize("scene,preset_name", [("bunny", "c2_bunny_chess"), ("spheres", "c1_spheres")])
def test_cpp_render_equals_python_binding(host_bins, gpu, tmp_path, scene, preset_name):
    exe, _ = host_bins
    W, H, NPH, ROUNDS = 96, 64, 9000, 2
    out = subprocess.check_output([exe, scene, str(W), str(H), str(NPH), str(ROUNDS), str(tmp_path / "o.ppm"),
                                   os.path.join(ROOT, "cgraytracing_b200", "assets")], text=True).split()
    with gpu.Context(0, gpu.preset(preset_name), gpu.RenderConfig(width=W, height=H)) as g:
        g.eye_pass(); g.build_grid()
        done = 0
        for r in range(ROUNDS):
            n = NPH // ROUNDS + (1 if r < NPH % ROUNDS else 0)
            g.photon_pass(done, n); g.round_update()
            done += n
        img, rgb8 = g.gather_image(float(NPH), want_rgb8=True)
        k = g.counters()
    assert int(out[0]) == k["hitpoints"] and int(out[1]) == k["deposits"] > 0
    ppm = (tmp_path / "o.ppm").read_bytes()
    header = f"P6\n{W} {H}\n255\n".encode()
    assert ppm.startswith(header)
    got8 = np.frombuffer(ppm[len(header):], np.uint8).reshape(H, W, 3)
    # fp64 atomics commute only up to rounding: the 8-bit pictures agree except for rare +-1 levels
    assert np.abs(got8.astype(int) - rgb8.astype(int)).max() <= 1
    assert abs(float(out[3]) - rgb8.mean()) < 0.01


@pytest.mark.gpu
def test_cpp_two_gpu_render_equals_one_gpu(host_bins, tmp_path):
    """render() with num_gpus = 2 (one host thread and one context per GPU; rows and photon ranges split; hitpoint records all-gathered and
    accumulators all-reduced by NCCL behind the C ABI) == the one-GPU render: same hitpoints, same deposits, same picture."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box (gpurun --gpus 2)")
    exe, _ = host_bins
    assets = os.path.join(ROOT, "cgraytracing_b200", "assets")
    outs = []
    for tag, gpus, how in (("1", 1, "peer"), ("2p", 2, "peer"), ("2n", 2, "nccl")):  # both exchanges: over peer memory (default) and ncclAllReduce
        o = subprocess.check_output([exe, "bunny", "160", "120", "60001", "3", str(tmp_path / f"o{tag}.ppm"), assets, str(gpus), how], text=True)
        outs.append(o.strip().splitlines()[-1].split())  # the last line is the program's (NCCL prints its version banner on stdout first)
    for k in (1, 2):
        assert outs[0][0] == outs[k][0] and outs[0][1] == outs[k][1] and int(outs[0][1]) > 0   # hitpoints, deposits
        assert abs(float(outs[0][3]) - float(outs[k][3])) < 0.01
    a, b, c = (np.frombuffer((tmp_path / f"o{g}.ppm").read_bytes()[-160 * 120 * 3:], np.uint8) for g in ("1", "2p", "2n"))
    assert np.abs(a.astype(int) - b.astype(int)).max() <= 1   # fp64 atomics + reduction order: identical up to rare +-1 levels
    assert np.abs(a.astype(int) - c.astype(int)).max() <= 1

"""CPU: the C-ABI library builds, loads, and exports exactly what include/cgrt.h declares; without a GPU the product fails
loudly (no CPU fallback). No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cgrt.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgrt_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from cgraytracing_b200 import build

    path = build.build()
    return path, C.CDLL(path)


def test_header_and_binding_list_agree():
    from cgraytracing_b200.binding import ABI_SYMBOLS

    assert header_functions() == sorted(ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    path, L = lib
    for name in header_functions():
        assert hasattr(L, name), name
    # and nothing of the test oracle is linked into the product
    syms = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    assert "orc_" not in syms and "cgref_" not in syms
    exported = sorted(set(re.findall(r"\bT (cgrt_[a-z0-9_]+)\b", syms)))
    assert exported == header_functions()


def test_header_compiles_as_c_and_cxx(tmp_path):
    for comp, ext, std in (("gcc", "c", "-std=c99"), ("g++", "cpp", "-std=c++11")):
        f = tmp_path / f"t.{ext}"
        f.write_text('#include "cgrt.h"\nint main(void){ cgrt_config c; cgrt_counters k; (void)c; (void)k; return sizeof(cgrt_config) == 112 ? 0 : 1; }\n')
        exe = tmp_path / f"t_{ext}"
        subprocess.check_call([comp, std, "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(f), "-o", str(exe)], env=_env())
        assert subprocess.call([str(exe)]) == 0


def _env():
    e = dict(os.environ)
    e.pop("CC", None)
    e.pop("CXX", None)
    return e


def test_config_struct_layout_matches_binding(lib):
    from cgraytracing_b200.binding import CgrtConfig, CgrtCounters

    _, L = lib
    assert C.sizeof(CgrtConfig) == 112 and C.sizeof(CgrtCounters) == 104
    k = CgrtConfig()
    L.cgrt_default_config(C.byref(k))
    # the reference's literals: main.cpp:28-29,35-36,177-184
    assert (k.width, k.height, k.max_depth, k.num_of_samples, k.hashsize) == (1024, 768, 5, 1, 1000001)
    assert (k.alpha, k.focus_plane, k.lens_radius) == (0.7, 20.0, 1.5)
    assert tuple(k.lightorg) == (0.0, 19.999, 20.0) and tuple(k.camorg) == (0.0, 0.0, -10.0)
    assert L.cgrt_version() >= 100


def test_no_gpu_means_loud_failure(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the no-device path cannot be exercised")
    _, L = lib
    h = C.c_void_p()
    L.cgrt_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    assert L.cgrt_create(0, C.byref(h)) == -3 and not h.value  # CGRT_ERR_NO_DEVICE
    from cgraytracing_b200 import CgrtError, Context

    with pytest.raises(CgrtError):
        Context(0)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "cgraytracing_b200")
    banned = ("import oracle", "from oracle", "liborc", "ppm_oracle", "oracle_capi", "libcgref", "oracle/")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for b in banned:
                    assert b not in text, (os.path.join(dirpath, f), b)

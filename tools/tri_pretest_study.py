"""Offline study for a conservative fp32 reject in front of the fp64 triangle test: how many of the triangle tests a
traversal performs could a float pretest with a rigorous error bound skip, and does it ever reject a triangle the fp64 test accepts?
CPU only (numpy); the float arithmetic is emulated operation by operation (no fused multiply-add, like the library's -fmad=false build).

  python tools/tri_pretest_study.py [rays]
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgraytracing_b200 import preset

nrays = int(sys.argv[1]) if len(sys.argv) > 1 else 400
rng = np.random.default_rng(7)
mesh = [o for o in preset("c3_dragon_glass").objects if o["kind"] == "mesh"][0]
T = mesh["tri9"].reshape(-1, 3, 3)
pa, pb, pc = T[:, 0], T[:, 1], T[:, 2]
e1, e2 = pa - pb, pa - pc  # objects.h:98-99
lo, hi = T.min(1) - 8e-5, T.max(1) + 8e-5
blo, bhi = lo.min(0), hi.max(0)

def det3(a, b, c):  # vec3.h determinant as the reference writes it, in the dtype of the inputs
    return (a[..., 0] * b[..., 1] * c[..., 2] + b[..., 0] * c[..., 1] * a[..., 2] + c[..., 0] * a[..., 1] * b[..., 2]
            - a[..., 0] * c[..., 1] * b[..., 2] - b[..., 0] * a[..., 1] * c[..., 2] - c[..., 0] * b[..., 1] * a[..., 2])

def absdet3(a, b, c):  # sum of the absolute values of the six products: the scale of the rounding error
    a, b, c = np.abs(a), np.abs(b), np.abs(c)
    return (a[..., 0] * b[..., 1] * c[..., 2] + b[..., 0] * c[..., 1] * a[..., 2] + c[..., 0] * a[..., 1] * b[..., 2]
            + a[..., 0] * c[..., 1] * b[..., 2] + b[..., 0] * a[..., 1] * c[..., 2] + c[..., 0] * b[..., 1] * a[..., 2])

tests = accept = rej_exact = rej_cheap = false_rej = 0
U = np.float32(2.0 ** -24)
for _ in range(nrays):
    # a ray from somewhere in the room through a random point of the dragon's box (what reaches the tree)
    o = np.array([rng.uniform(-20, 20), rng.uniform(-20, 20), rng.uniform(-10, 40)])
    tgt = rng.uniform(blo, bhi)
    d = tgt - o; d /= np.linalg.norm(d)
    inv = 1.0 / d
    t0, t1 = (lo - o) * inv, (hi - o) * inv
    tn = np.maximum(np.minimum(t0, t1).max(1), 0.0); tf = np.maximum(t0, t1).min(1)
    idx = np.nonzero(tn <= tf)[0]  # triangles whose padded box the ray crosses: a superset of the leaves a BVH walk tests
    if idx.size == 0: continue
    E1, E2, PA = e1[idx], e2[idx], pa[idx]
    s = PA - o
    dd = np.broadcast_to(d, s.shape)
    D1, D3, D4 = det3(dd, E1, E2), det3(dd, s, E2), det3(dd, E1, s)
    pos = D1 > 0
    ok = (D1 != 0) & ((D3 == 0) | ((D3 > 0) == pos)) & ((D4 == 0) | ((D4 > 0) == pos)) & np.where(pos, D3 + D4 <= D1, D3 + D4 >= D1)
    tq = det3(s, E1, E2) / np.where(D1 == 0, 1, D1)
    ok &= tq > 0
    # float pretest: the three determinants in float with |error| <= 16 u * (sum of |products|)
    f = lambda x: x.astype(np.float32)
    fd, fe1, fe2, fs = f(dd), f(E1), f(E2), f(s)
    F1, F3, F4 = det3(fd, fe1, fe2), det3(fd, fs, fe2), det3(fd, fe1, fs)
    B1, B3, B4 = (np.float32(16) * U * absdet3(fd, fe1, fe2), np.float32(16) * U * absdet3(fd, fs, fe2), np.float32(16) * U * absdet3(fd, fe1, fs))
    sure = np.abs(F1) > B1
    sg = np.where(F1 > 0, np.float32(1), np.float32(-1))
    rej = sure & ((F3 * sg < -B3) | (F4 * sg < -B4) | ((F3 + F4 - F1) * sg > B1 + B3 + B4))
    # cheaper bound: one scale per determinant from infinity norms (6 |a||b||c|)
    n = lambda x: np.abs(x).max(-1)
    C1, C3, C4 = (np.float32(96) * U * n(fd) * n(fe1) * n(fe2), np.float32(96) * U * n(fd) * n(fs) * n(fe2), np.float32(96) * U * n(fd) * n(fe1) * n(fs))
    rej2 = (np.abs(F1) > C1) & ((F3 * sg < -C3) | (F4 * sg < -C4) | ((F3 + F4 - F1) * sg > C1 + C3 + C4))
    tests += idx.size; accept += int(ok.sum()); rej_exact += int(rej.sum()); rej_cheap += int(rej2.sum()); false_rej += int((rej & ok).sum()) + int((rej2 & ok).sum())
print(f"rays {nrays}: box-hit triangle tests {tests}, fp64 accepts {accept} ({100*accept/max(tests,1):.2f} %)")
print(f"float pretest rejects {rej_exact} ({100*rej_exact/max(tests-accept,1):.1f} % of the misses) with the per-product bound, "
      f"{rej_cheap} ({100*rej_cheap/max(tests-accept,1):.1f} %) with the infinity-norm bound; false rejects {false_rej}")

"""Scene descriptions for the photon-mapping hot path: plain data, no compute.

A `SceneDesc` is the ordered object list that the reference builds in `main()` (main.cpp:277-378) expressed as
arrays, so that exactly the same inputs can be handed to the GPU path (through the C ABI, include/cgrt.h) and, in
tests, to the CPU oracle. The presets are the reference's scene code and its commented-out variants
(SURVEY.md Appendix C); every literal below is cited to main.cpp.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


# ---------------------------------------------------------------------------------------------------------------
# asset containers (written by tools/make_assets.py)
# ---------------------------------------------------------------------------------------------------------------
def read_mesh_asset(path: str):
    """-> (verts float64 [nv,3] in file space, faces int32 [nf,3] 0-based)."""
    with open(path, "rb") as fp:
        raw = fp.read()
    if raw[:8] != b"CGRTMSH1":
        raise ValueError(f"{path}: not a .cgrtmesh container")
    nv, nf = struct.unpack_from("<ii", raw, 8)
    v = np.frombuffer(raw, dtype="<f8", count=nv * 3, offset=16).reshape(nv, 3)
    f = np.frombuffer(raw, dtype="<i4", count=nf * 3, offset=16 + nv * 24).reshape(nf, 3)
    if nf and (f.min() < 0 or f.max() >= nv):
        raise ValueError(f"{path}: face index out of range")
    return v.copy(), f.copy()


def read_texture_asset(path: str) -> np.ndarray:
    """-> uint8 [h,w,3], the bytes stbi_load(...,3) returns in the reference (main.cpp:300)."""
    with open(path, "rb") as fp:
        raw = fp.read()
    if raw[:8] != b"CGRTTEX1":
        raise ValueError(f"{path}: not a .cgrttex container")
    w, h = struct.unpack_from("<ii", raw, 8)
    return np.frombuffer(raw, dtype=np.uint8, count=w * h * 3, offset=16).reshape(h, w, 3).copy()


def transform_mesh(verts: np.ndarray, faces: np.ndarray, a: float, b) -> np.ndarray:
    """TriangleMesh loader arithmetic (objects.h:348,365,371,384,398): Vec3(x,y,-z) * a + b, per component in fp64.
    -> tri9 float64 [nf, 9] = (pa, pb, pc)."""
    v = np.array(verts, dtype=np.float64, copy=True)
    v[:, 2] = -v[:, 2]
    v = v * np.float64(a) + np.asarray(b, dtype=np.float64)[None, :]
    return np.ascontiguousarray(v[faces].reshape(len(faces), 9))


# ---------------------------------------------------------------------------------------------------------------
@dataclass
class SceneDesc:
    """Ordered textures and objects; `objects[i]` gets object id i, the order of `objs` in main.cpp:355-378."""

    textures: List[Dict[str, Any]] = field(default_factory=list)
    objects: List[Dict[str, Any]] = field(default_factory=list)
    name: str = "custom"

    def add_texture(self, rgb, n, p, lenx, leny, isbump=False) -> int:  # Texture(data,n,p,lx,ly,flag) texture.h:19
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        assert rgb.ndim == 3 and rgb.shape[2] == 3
        self.textures.append(dict(rgb=rgb, n=_v(n), p=_v(p), lenx=float(lenx), leny=float(leny), isbump=bool(isbump)))
        return len(self.textures) - 1

    def add_sphere(self, c, r, col, refl=0.0, transp=0.0) -> int:  # Sphere(c,r,sc,refl,transp) objects.h:28-38
        self.objects.append(dict(kind="sphere", c=_v(c), r=float(r), col=_v(col), refl=float(refl), transp=float(transp)))
        return len(self.objects) - 1

    def add_plane(self, p, n, col, refl=0.0, transp=0.0, tex=-1) -> int:  # Plane(p,n,sc,refl,transp,tx) objects.h:480
        self.objects.append(dict(kind="plane", p=_v(p), n=_v(n), col=_v(col), refl=float(refl), transp=float(transp), tex=int(tex)))
        return len(self.objects) - 1

    def add_mesh(self, tri9, col, refl=0.0, transp=0.0, objtype=0) -> int:  # TriangleMesh(...) objects.h:338-340
        tri9 = np.ascontiguousarray(tri9, dtype=np.float64).reshape(-1, 9)
        self.objects.append(dict(kind="mesh", tri9=tri9, col=_v(col), refl=float(refl), transp=float(transp), objtype=int(objtype)))
        return len(self.objects) - 1

    def add_bezier(self, cp, pos, col, refl=0.0, transp=0.0) -> int:  # Bezier(points,pos,sc,refl,transp) bezier.h:44-45
        cp = np.ascontiguousarray(cp, dtype=np.float64).reshape(-1, 3)
        self.objects.append(dict(kind="bezier", cp=cp, pos=_v(pos), col=_v(col), refl=float(refl), transp=float(transp)))
        return len(self.objects) - 1

    def build_into(self, backend) -> None:
        """Replay the description into anything exposing add_texture/add_sphere/add_plane/add_mesh/add_bezier."""
        for t in self.textures:
            backend.add_texture(t["rgb"], t["n"], t["p"], t["lenx"], t["leny"], t["isbump"])
        for o in self.objects:
            k = o["kind"]
            if k == "sphere":
                backend.add_sphere(o["c"], o["r"], o["col"], o["refl"], o["transp"])
            elif k == "plane":
                backend.add_plane(o["p"], o["n"], o["col"], o["refl"], o["transp"], o["tex"])
            elif k == "mesh":
                backend.add_mesh(o["tri9"], o["col"], o["refl"], o["transp"], o["objtype"])
            elif k == "bezier":
                backend.add_bezier(o["cp"], o["pos"], o["col"], o["refl"], o["transp"])
            else:
                raise ValueError(k)

    def num_triangles(self) -> int:
        return int(sum(len(o["tri9"]) for o in self.objects if o["kind"] == "mesh"))


def _v(x) -> np.ndarray:
    a = np.asarray(x, dtype=np.float64).reshape(3)
    return a.copy()


# ---------------------------------------------------------------------------------------------------------------
# presets
# ---------------------------------------------------------------------------------------------------------------
def _walls(s: SceneDesc, floor_tex: int = -1) -> None:
    """The five planes of main.cpp:348-353 (floor, right, left, back, ceiling), in that order."""
    s.add_plane((0.0, -20, 0), (0, 1, 0), (0.15, 0.15, 0.15), 0.0, 0.0, floor_tex)
    s.add_plane((20, 0.0, 0), (-1, 0, 0), (0.15, 0.50, 0.15), 0.0, 0.0)
    s.add_plane((-20, 0.0, 0), (1, 0, 0), (0.50, 0.15, 0.15), 0.0, 0.0)
    s.add_plane((0.0, 0.0, 40), (0, 0, -1), (0.15, 0.15, 0.15), 0.0, 0.0)
    s.add_plane((0.0, 20, 0), (0, -1, 0), (0.15, 0.15, 0.15), 0.0, 0.0)


def _floor_texture(s: SceneDesc, name: str, bump: bool) -> int:
    """Texture(tdata, Vec3(0,1,0), Vec3(-21,0,0), 42, 40, bump) — main.cpp:320."""
    rgb = read_texture_asset(os.path.join(ASSET_DIR, name + ".cgrttex"))
    return s.add_texture(rgb, (0, 1, 0), (-21, 0, 0), 42, 40, bump)


def _mesh(name: str, a: float, b):
    v, f = read_mesh_asset(os.path.join(ASSET_DIR, name + ".cgrtmesh"))
    return transform_mesh(v, f, a, b)


def preset(name: str, max_tris: Optional[int] = None) -> SceneDesc:
    """Named scenes. `max_tris` truncates meshes (tests only; never used by bench.py)."""
    s = SceneDesc(name=name)

    def cut(t):
        return t if max_tris is None else t[:max_tris]

    if name == "default_bump":
        # main.cpp as checked in (:292, :300, :320, :348-366): diffuse dragon + stone displacement floor
        tex = _floor_texture(s, "stone", True)
        _walls(s, tex)
        s.add_mesh(cut(_mesh("dragon", 1.5, (-5, -20, 30))), (0.25, 0.25, 0.5), 0.0, 0.0, 1)
    elif name == "c1_spheres_bezier":
        # main.cpp:288-290 spheres (objs order: spheres first, :356-359), chessboard floor, Bezier vase :371-378
        tex = _floor_texture(s, "ChessBoard", False)
        s.add_sphere((-15.0, -20.0, 60), 10, (0.3, 0.3, 0.3), 0.0, 0.0)
        s.add_sphere((10.0, -20.0, 60), 7, (1.0, 1.0, 1.0), 0.8, 0.0)
        s.add_sphere((10.0, -20.0, 30), 7, (1.0, 1.0, 1.0), 0.8, 0.5)
        _walls(s, tex)
        s.add_bezier([(0, -10, 4), (0, 2, 4), (0, -2, 0), (0, 10, 2)], (15, -10.1, 35), (1.0, 1.0, 1.0), 0.5, 0.0)
    elif name == "c1_spheres":
        # c1 without the Bezier vase (deterministic primitives only)
        tex = _floor_texture(s, "ChessBoard", False)
        s.add_sphere((-15.0, -20.0, 60), 10, (0.3, 0.3, 0.3), 0.0, 0.0)
        s.add_sphere((10.0, -20.0, 60), 7, (1.0, 1.0, 1.0), 0.8, 0.0)
        s.add_sphere((10.0, -20.0, 30), 7, (1.0, 1.0, 1.0), 0.8, 0.5)
        _walls(s, tex)
    elif name == "c1_mirror":
        # c1_spheres plus a mirror sphere INSIDE the room, so that the mirror branch of trace() (main.cpp:129-134) is reached by eye
        # rays and photons: the only mirror of main.cpp:288-290, Sphere((10,-20,60),7,...,0.8,0), sits behind the back wall (z = 40)
        tex = _floor_texture(s, "ChessBoard", False)
        s.add_sphere((-15.0, -20.0, 60), 10, (0.3, 0.3, 0.3), 0.0, 0.0)
        s.add_sphere((10.0, -20.0, 60), 7, (1.0, 1.0, 1.0), 0.8, 0.0)
        s.add_sphere((10.0, -20.0, 30), 7, (1.0, 1.0, 1.0), 0.8, 0.5)
        s.add_sphere((-8.0, -13.0, 25), 7, (0.9, 0.8, 0.7), 0.8, 0.0)
        _walls(s, tex)
    elif name == "c2_bunny_chess":
        # main.cpp:293 glass bunny (type 0) + chessboard floor
        tex = _floor_texture(s, "ChessBoard", False)
        _walls(s, tex)
        s.add_mesh(cut(_mesh("lowpolybunny", 10, (0, -15, 40))), (1.0, 1.0, 1.0), 0.8, 0.5, 0)
    elif name == "c3_dragon_glass":
        # main.cpp:292 geometry with the glass material of :293-294 + chessboard floor
        tex = _floor_texture(s, "ChessBoard", False)
        _walls(s, tex)
        s.add_mesh(cut(_mesh("dragon", 1.5, (-5, -20, 30))), (1.0, 1.0, 1.0), 0.8, 0.5, 1)
    elif name == "c4_bump_dof":
        # stone displacement floor + Mesh000.obj "water" (main.cpp:294 with the mesh that exists) ; DOF camera is a Config flag
        tex = _floor_texture(s, "stone", True)
        _walls(s, tex)
        s.add_mesh(cut(_mesh("Mesh000", 20, (0, -15, 30))), (1.0, 1.0, 1.0), 0.8, 0.5, 2)
    elif name == "walls_only":
        _walls(s)
    else:
        raise ValueError(f"unknown preset {name!r}")
    return s


PRESETS = ["default_bump", "c1_spheres_bezier", "c1_spheres", "c1_mirror", "c2_bunny_chess", "c3_dragon_glass", "c4_bump_dof", "walls_only"]


@dataclass
class RenderConfig:
    """Every compile-time constant of the reference's render()/trace() as a parameter (main.cpp:28-36,177-184,222-224)."""

    width: int = 1024
    height: int = 768
    max_depth: int = 5
    num_of_samples: int = 1
    use_dof: int = 0
    consume_dof_rng: int = 1
    hashsize: int = 1000001
    update_mode: int = 1  # 0 = U1 per-photon (reference), 1 = U2 per-round
    into_rule: int = 1  # 0 = reference hit-count parity heuristic, 1 = winding sign (GPU rule)
    alpha: float = 0.7
    focus_plane: float = 20.0
    lens_radius: float = 1.5
    lightorg: tuple = (0.0, 19.999, 20.0)
    camorg: tuple = (0.0, 0.0, -10.0)
    seed: int = 20261018

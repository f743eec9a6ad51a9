// ppm_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A UB-free, parameterised restatement of CGRayTracing's progressive photon mapper hot path,
// following the reference's arithmetic *operation by operation* so that it is bit-identical to
// the reference on identical inputs and identical random numbers (pinned by tests/test_oracle_vs_ref.py
// against oracle/_ref, which compiles the reference's own headers from /root/reference).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
// anything under oracle/. The product (cgraytracing_b200/) never links or calls it.
//
// Compile with -ffp-contract=off (no FMA contraction) so that fp64 results equal the GPU's
// (-fmad=false) and the reference's (g++ -O2 without -mfma never contracts).
//
// Every block cites the reference file:line it restates (paths relative to /root/reference).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------------------------
// headers/vec3.h:11-92 — double3 value type. operator* is scalar or element-wise; normalize()
// multiplies by (1/len) (vec3.h:35-43), it does not divide.
// ------------------------------------------------------------------------------------------------
struct Vec3 {
    double x, y, z;
    Vec3(double x_ = 0, double y_ = 0, double z_ = 0) : x(x_), y(y_), z(z_) {}
    double norm() const { return std::sqrt(x * x + y * y + z * z); }       // vec3.h:31-33
    Vec3 normalize() {                                                      // vec3.h:35-43
        double len = norm();
        if (len > 0) {
            x *= 1 / len;
            y *= 1 / len;
            z *= 1 / len;
        }
        return *this;
    }
    Vec3 operator*(double f) const { return Vec3(x * f, y * f, z * f); }    // vec3.h:50-52
    Vec3 operator*(const Vec3 &v) const { return Vec3(x * v.x, y * v.y, z * v.z); }  // :54-56
    Vec3 mul(const Vec3 &v) const { return Vec3(x * v.x, y * v.y, z * v.z); }        // :57-59
    double dot(const Vec3 &v) const { return x * v.x + y * v.y + z * v.z; }          // :61-63
    Vec3 operator+(const Vec3 &v) const { return Vec3(x + v.x, y + v.y, z + v.z); }
    Vec3 operator+(double b) const { return Vec3(x + b, y + b, z + b); }
    Vec3 operator-(const Vec3 &v) const { return Vec3(x - v.x, y - v.y, z - v.z); }
    Vec3 operator-(double b) const { return Vec3(x - b, y - b, z - b); }
    Vec3 operator-() const { return Vec3(-x, -y, -z); }
    Vec3 cross(const Vec3 &b) const {                                       // vec3.h:81-83
        return Vec3(y * b.z - z * b.y, z * b.x - x * b.z, x * b.y - y * b.x);
    }
};

// vec3.h:95-97 — Sarrus expansion, evaluated strictly left to right.
inline double det(const Vec3 &a, const Vec3 &b, const Vec3 &c) {
    return (a.x * b.y * c.z + b.x * c.y * a.z + c.x * a.y * b.z - a.x * c.y * b.z - b.x * a.y * c.z -
            c.x * b.y * a.z);
}
// vec3.h:99-101
inline Vec3 matrixVectorProduct(const Vec3 &a, const Vec3 &b, const Vec3 &c, const Vec3 &d) {
    return a * d.x + b * d.y + c * d.z;
}
// vec3.h:103-119 — singular iff |det| < 1e-4 (doubleeps, vec3.h:9).
inline bool inv(const Vec3 &a, const Vec3 &b, const Vec3 &c, Vec3 &resa, Vec3 &resb, Vec3 &resc) {
    const double doubleeps = 1e-4;
    double d = det(a, b, c);
    if (d < doubleeps && d > -doubleeps) return false;
    resa.x = (b.y * c.z - b.z * c.y) / d;
    resa.y = (c.y * a.z - c.z * a.y) / d;
    resa.z = (a.y * b.z - a.z * b.y) / d;
    resb.x = (c.x * b.z - c.z * b.x) / d;
    resb.y = (a.x * c.z - a.z * c.x) / d;
    resb.z = (b.x * a.z - b.z * a.x) / d;
    resc.x = (b.x * c.y - c.x * b.y) / d;
    resc.y = (c.x * a.y - c.y * a.x) / d;
    resc.z = (a.x * b.y - a.y * b.x) / d;
    return true;
}

// headers/util.h:16-26, 32-42 — 3-argument max/min with the reference's comparison structure.
inline double max3(double a, double b, double c) {
    if (a > b && a > c) return a;
    else if (b > c) return b;
    else return c;
}
inline double min3(double a, double b, double c) {
    if (a < b && a < c) return a;
    else if (b < c) return b;
    else return c;
}
// util.h:45-47
inline int gammaCorr(double x) { return int(std::pow(1 - std::exp(-x), 1 / 2.2) * 255 + .5); }

// ------------------------------------------------------------------------------------------------
// Random numbers. The reference draws (double)rand()/RAND_MAX (sampling.h:13-15,32) with glibc's
// RAND_MAX = 2^31-1. Two interchangeable generators:
//   * LibcCompat: pulls 31-bit integers from a caller-supplied function in exactly the reference's
//     call order (used to pin the oracle bit-for-bit against oracle/_ref, whose rand() is interposed
//     with the same sequence);
//   * Philox4x32-10 counter streams (the definition shared with the GPU): key = {seed_lo, seed_hi + pass},
//     counter = {path_lo, path_hi, dim, block}; word w -> u = (double)(w >> 1) * 2^-31, a 31-bit lattice on
//     [0,1) like rand()/RAND_MAX's on [0,1] (one exact multiply instead of an fp64 division on the GPU).
// ------------------------------------------------------------------------------------------------
inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { PASS_EYE = 0, PASS_PHOTON = 1, PASS_BEZIER = 2 };

struct Rng {
    // mode 0: libc-compatible external source; mode 1: Philox stream
    int mode = 1;
    int (*ext)(void *) = nullptr;  // returns an integer in [0, 2^31-1]
    void *ext_state = nullptr;
    uint32_t key[2] = {0, 0}, ctr[4] = {0, 0, 0, 0}, buf[4] = {0, 0, 0, 0};
    int idx = 4;
    uint64_t draws = 0;

    static Rng external(int (*f)(void *), void *st) {
        Rng r; r.mode = 0; r.ext = f; r.ext_state = st; return r;
    }
    static Rng philox(uint64_t seed, uint32_t pass, uint64_t path, uint32_t dim) {
        Rng r; r.mode = 1; r.reseed(seed, pass, path, dim); return r;
    }
    void reseed(uint64_t seed, uint32_t pass, uint64_t path, uint32_t dim) {
        if (mode == 0) return;  // a libc stream is never re-keyed
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32) + pass;
        ctr[0] = (uint32_t)path;
        ctr[1] = (uint32_t)(path >> 32);
        ctr[2] = dim;
        ctr[3] = 0;
        idx = 4;
    }
    // (double)rand() / RAND_MAX
    double u01() {
        draws++;
        if (mode == 0) return (double)ext(ext_state) / 2147483647.0;
        if (idx == 4) {
            philox4x32_10(ctr, key, buf);
            ctr[3]++;
            idx = 0;
        }
        return (double)(buf[idx++] >> 1) * (1.0 / 2147483648.0);
    }
};

// Rejection loops are bounded so the GPU twin cannot spin forever; P(64 consecutive rejections) < 1e-20.
static const int MAX_REJECT = 64;

// sin and cos of 2*pi*v for v in [0,1), from +,-,* only (no libm, no contraction), so that the CPU oracle and the GPU
// produce identical bits. Quadrant reduction is exact (4v, k = round-to-nearest quadrant, 4v - k are exact in binary
// fp64); the kernels are the classical minimax polynomials on [-pi/4, pi/4] (error < 1e-16).
inline void sincos2pi(double v, double &s_out, double &c_out) {
    double t = 4.0 * v;
    int k = (int)(t + 0.5);
    double r = t - (double)k;
    double x = r * 1.5707963267948966;
    double z = x * x;
    double ps = -1.66666666666666324348e-01 + z * (8.33333333332248946124e-03 + z * (-1.98412698298579493134e-04 + z * (2.75573137070700676789e-06 +
                z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10))));
    double sn = x + x * z * ps;
    double pc = 4.16666666666666019037e-02 + z * (-1.38888888888741095749e-03 + z * (2.48015872894767294178e-05 + z * (-2.75573143513906633035e-07 +
                z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11))));
    double cs = 1.0 - 0.5 * z + z * z * pc;
    switch (k & 3) {
        case 0: s_out = sn; c_out = cs; break;
        case 1: s_out = cs; c_out = -sn; break;
        case 2: s_out = -sn; c_out = -cs; break;
        default: s_out = -cs; c_out = sn; break;
    }
}

// sampling.h:11-20. libc mode: the reference's rejection loop draw for draw (pins the oracle to the compiled reference).
// Philox mode (the definition shared with the GPU): the same uniform distribution on the sphere by Archimedes' map,
// z = 1 - 2 u1, phi = 2 pi u2 — two draws, no rejection, hence no divergent loop on the GPU (any exact sampler
// of the same distribution is admissible; tests/test_oracle_golden.py checks the two samplers agree in distribution).
inline Vec3 uniform_sampling_sphere(Rng &rng) {
    if (rng.mode != 0) {
        double z = 1.0 - 2.0 * rng.u01();
        double sn, cs;
        sincos2pi(rng.u01(), sn, cs);
        double r = std::sqrt(1.0 - z * z);
        return Vec3(r * cs, r * sn, z);
    }
    Vec3 v;
    for (int it = 0; it < MAX_REJECT; it++) {
        double x = rng.u01() * 2.0 - 1;
        double y = rng.u01() * 2.0 - 1;
        double z = rng.u01() * 2.0 - 1;
        v = Vec3(x, y, z);
        if (x * x + y * y + z * z <= 1) break;
    }
    return v.normalize();
}
// sampling.h:22-29. Philox mode: a uniform sphere sample mirrored into the hemisphere (measure preserving).
inline Vec3 uniform_sampling_halfsphere(Rng &rng, const Vec3 &dir) {
    if (rng.mode != 0) {
        Vec3 s = uniform_sampling_sphere(rng);
        if (s.dot(dir) < 0) s = Vec3(-s.x, -s.y, -s.z);
        return s;
    }
    Vec3 s;
    for (int it = 0; it < MAX_REJECT; it++) {
        s = uniform_sampling_sphere(rng);
        if (s.dot(dir) > 0) break;
    }
    return s;
}
// sampling.h:31-33
inline double uniform_sampling_zeroone(Rng &rng) { return rng.u01(); }
// sampling.h:35-43 (strict < 1)
inline Vec3 uniform_sampling_circle(Rng &rng, double radius) {
    double x = 0, y = 0;
    for (int it = 0; it < MAX_REJECT; it++) {
        x = rng.u01() * 2.0 - 1;
        y = rng.u01() * 2.0 - 1;
        if (x * x + y * y < 1) break;
    }
    return Vec3(x, y, 0) * radius;
}

// ------------------------------------------------------------------------------------------------
// headers/texture.h — texels (byte/256, main.cpp:307-311), luminance height table (texture.h:27-37),
// planar-projection nearest-texel lookup (texture.h:39-72).
// ------------------------------------------------------------------------------------------------
struct Texture {
    bool htexture = false;
    bool isbump = false;  // the reference leaves this uninitialised in the default ctor (texture.h:16-18); we define it
    int H = 0, W = 0;     // data.size(), data[0].size()
    std::vector<Vec3> data;       // row-major [H][W]
    std::vector<double> height;   // row-major [H][W]
    Vec3 normal, position;
    double lenx = 0, leny = 0;

    Texture() {}
    // texture.h:19-38
    Texture(const std::vector<Vec3> &d, int H_, int W_, const Vec3 &n, const Vec3 &p, double lx, double ly, bool flag = false)
        : htexture(true), isbump(flag), H(H_), W(W_), data(d), normal(n), position(p), lenx(lx), leny(ly) {
        height.assign((size_t)H * W, 0.0);
        double coeff = 0.5;
        if (isbump) {
            for (int i = 0; i < H; i++)
                for (int j = 0; j < W; j++) {
                    const Vec3 &t = data[(size_t)i * W + j];
                    double h = (0.299 * t.x + 0.587 * t.y + 0.114 * t.z);
                    h = 1 - std::exp(-3.3 * h);
                    h *= coeff;
                    height[(size_t)i * W + j] = h;
                }
        }
    }
    // texture.h:39-72. Returns the texel index through *row,*col as well (for tests).
    bool color(const Vec3 &point, Vec3 &color, int *row = nullptr, int *col = nullptr) const {
        const double texteps = 1e-2;
        if (!htexture) return false;
        Vec3 d = point - position;
        d = d - normal * (d.dot(normal));
        int r = -1, c = -1;
        if (d.x < texteps && d.x > -texteps) {
            if (0 < d.y && d.y < lenx && 0 < d.z && d.z < leny) {
                int id1 = (int)std::floor(d.y / lenx * H);
                int id2 = (int)std::floor(d.z / leny * W);
                r = id1; c = id2;
            } else return false;
        } else if (d.y < texteps && d.y > -texteps) {
            if (0 < d.x && d.x < lenx && 0 < d.z && d.z < leny) {
                int id1 = (int)std::floor(d.x / lenx * W);
                int id2 = (int)std::floor(d.z / leny * H);
                r = id2; c = id1;
            } else return false;
        } else if (d.z < texteps && d.z > -texteps) {
            if (0 < d.x && d.x < lenx && 0 < d.y && d.y < leny) {
                int id1 = (int)std::floor(d.x / lenx * W);
                int id2 = (int)std::floor(d.y / leny * H);
                r = H - 1 - id2; c = id1;
            } else return false;
        } else {
            return false;
        }
        // The reference indexes data[r][c] unchecked; for the first branch (id1 scaled by H, id2 by W) that is
        // in range only for square textures. Clamp defensively (no reference scene uses that branch).
        if (r < 0) r = 0; if (r >= H) r = H - 1;
        if (c < 0) c = 0; if (c >= W) c = W - 1;
        color = data[(size_t)r * W + c];
        if (row) *row = r;
        if (col) *col = c;
        return true;
    }
};

// ------------------------------------------------------------------------------------------------
// headers/objects.h:91-141 — Triangle
// ------------------------------------------------------------------------------------------------
struct Triangle {
    Vec3 pa, pb, pc;
    Triangle() {}
    Triangle(const Vec3 &a, const Vec3 &b, const Vec3 &c) : pa(a), pb(b), pc(c) {}
    // objects.h:96-111 — Cramer's rule with four determinants and four divisions.
    bool intersect(const Vec3 &rayorig, const Vec3 &raydir, double &len, Vec3 &normalvector) const {
        Vec3 e1 = pa - pb;
        Vec3 e2 = pa - pc;
        Vec3 s = pa - rayorig;
        double det1 = det(raydir, e1, e2);
        double det2 = det(s, e1, e2);
        double det3 = det(raydir, s, e2);
        double det4 = det(raydir, e1, s);
        if (det2 / det1 > 0.0 && det3 / det1 >= 0.0 && det4 / det1 >= 0.0 && (det3 + det4) / det1 <= 1.0) {
            len = det2 / det1;
            normalvector = ((pa - pb).cross(pa - pc)).normalize();
            return true;
        }
        return false;
    }
    Vec3 normalvec() const { return ((pa - pb).cross(pa - pc)).normalize(); }
    double max_x() const { return max3(pa.x, pb.x, pc.x); }
    double max_y() const { return max3(pa.y, pb.y, pc.y); }
    double max_z() const { return max3(pa.z, pb.z, pc.z); }
    double min_x() const { return min3(pa.x, pb.x, pc.x); }
    double min_y() const { return min3(pa.y, pb.y, pc.y); }
    double min_z() const { return min3(pa.z, pb.z, pc.z); }
};

static const int Minkdsize = 10;         // objects.h:143
static const double epsdouble = 1e-4;    // objects.h:144
static const double doubleINF = 1e10;    // objects.h:15

// Ray / axis-aligned-box "any face hit at t>0 within +-1e-4" test shared by KDNode::intersect
// (objects.h:166-200) and Bezier::intersect_with_box (bezier.h:72-126).
struct Box6 {
    double xmax, xmin, ymax, ymin, zmax, zmin;
};

// objects.h:166-200
inline bool box_any_face_hit(const Box6 &b, const Vec3 &o, const Vec3 &d) {
    double t;
    Vec3 p;
    t = (b.xmax - o.x) / d.x; p = o + d * t;
    if (t > 0 && p.y >= b.ymin - epsdouble && p.y <= b.ymax + epsdouble && p.z >= b.zmin - epsdouble && p.z <= b.zmax + epsdouble) return true;
    t = (b.xmin - o.x) / d.x; p = o + d * t;
    if (t > 0 && p.y >= b.ymin - epsdouble && p.y <= b.ymax + epsdouble && p.z >= b.zmin - epsdouble && p.z <= b.zmax + epsdouble) return true;
    t = (b.ymax - o.y) / d.y; p = o + d * t;
    if (t > 0 && p.x >= b.xmin - epsdouble && p.x <= b.xmax + epsdouble && p.z >= b.zmin - epsdouble && p.z <= b.zmax + epsdouble) return true;
    t = (b.ymin - o.y) / d.y; p = o + d * t;
    if (t > 0 && p.x >= b.xmin - epsdouble && p.x <= b.xmax + epsdouble && p.z >= b.zmin - epsdouble && p.z <= b.zmax + epsdouble) return true;
    t = (b.zmax - o.z) / d.z; p = o + d * t;
    if (t > 0 && p.x >= b.xmin - epsdouble && p.x <= b.xmax + epsdouble && p.y >= b.ymin - epsdouble && p.y <= b.ymax + epsdouble) return true;
    t = (b.zmin - o.z) / d.z; p = o + d * t;
    if (t > 0 && p.x >= b.xmin - epsdouble && p.x <= b.xmax + epsdouble && p.y >= b.ymin - epsdouble && p.y <= b.ymax + epsdouble) return true;
    return false;
}

// How a mesh decides whether the ray origin is inside (only glass materials read it, main.cpp:140-150).
enum IntoRule {
    INTO_REFERENCE_PARITY = 0,  // objects.h:318-332: parity of running-minimum updates over the whole tree
    INTO_WINDING = 1            // stored winding normal x per-mesh orientation sign (the GPU rule)
};

// ------------------------------------------------------------------------------------------------
// headers/objects.h:147-333 — KDNode / KDTree. A median-split BVH. The reference stores a copy of the
// sub-list in every node; we store the id list (same order), which is all intersect_subtree reads.
// ------------------------------------------------------------------------------------------------
struct KDNode {
    Box6 box;
    int left = -1, right = -1;
    std::vector<int> tri;  // the node's triangleList, as ids into KDTree::tris, in the order received
};

struct KDCounters {
    uint64_t node_visits = 0, tri_tests = 0;
};

struct KDTree {
    std::vector<Triangle> tris;
    std::vector<KDNode> kdnodes;
    double orient_sign = 1.0;  // +1 if winding normals point outward (signed volume < 0 for (pa-pb)x(pa-pc)), see set_orientation

    bool empty() const { return kdnodes.empty(); }

    // objects.h:217-267 (sublist by value, sorted in place by triangle max coordinate on curDIM)
    void buildKdTree(std::vector<int> sublist, int parID, bool isLeft, int curDIM, bool isRoot = false) {
        int curID = (int)kdnodes.size();
        if (!isRoot) {
            if (isLeft) kdnodes[parID].left = curID;
            else kdnodes[parID].right = curID;
        }
        kdnodes.push_back(KDNode());
        kdnodes[curID].tri = sublist;
        Box6 bx;
        bx.xmax = -doubleINF; bx.ymax = -doubleINF; bx.zmax = -doubleINF;
        bx.xmin = doubleINF; bx.ymin = doubleINF; bx.zmin = doubleINF;
        for (size_t i = 0; i < sublist.size(); i++) {
            const Triangle &a = tris[sublist[i]];
            double max_x = a.max_x(), max_y = a.max_y(), max_z = a.max_z();
            double min_x = a.min_x(), min_y = a.min_y(), min_z = a.min_z();
            if (bx.xmax < max_x) bx.xmax = max_x;
            if (bx.ymax < max_y) bx.ymax = max_y;
            if (bx.zmax < max_z) bx.zmax = max_z;
            if (bx.xmin > min_x) bx.xmin = min_x;
            if (bx.ymin > min_y) bx.ymin = min_y;
            if (bx.zmin > min_z) bx.zmin = min_z;
        }
        kdnodes[curID].box = bx;
        if ((int)sublist.size() < Minkdsize) return;
        const std::vector<Triangle> &T = tris;
        if (curDIM == 0) std::sort(sublist.begin(), sublist.end(), [&T](int p, int q) { return T[p].max_x() < T[q].max_x(); });
        else if (curDIM == 1) std::sort(sublist.begin(), sublist.end(), [&T](int p, int q) { return T[p].max_y() < T[q].max_y(); });
        else std::sort(sublist.begin(), sublist.end(), [&T](int p, int q) { return T[p].max_z() < T[q].max_z(); });
        std::vector<int> leftsublist(sublist.begin(), sublist.begin() + sublist.size() / 2);
        std::vector<int> rightsublist(sublist.begin() + sublist.size() / 2, sublist.end());
        int newDIM = (curDIM + 1) % 3;
        buildKdTree(leftsublist, curID, true, newDIM);
        buildKdTree(rightsublist, curID, false, newDIM);
    }

    void build(const std::vector<Triangle> &t) {
        tris = t;
        kdnodes.clear();
        std::vector<int> ids(tris.size());
        for (size_t i = 0; i < ids.size(); i++) ids[i] = (int)i;
        buildKdTree(ids, 0, false, 0, true);  // objects.h:402, :501
        set_orientation();
    }

    // Per-mesh orientation sign for INTO_WINDING: sum of pa . (pb x pc) (six times the signed volume for
    // counter-clockwise-outward winding). The stored normal is (pa-pb)x(pa-pc) = (pb-pa)x(pc-pa), i.e. the
    // usual winding normal, so it points outward iff the signed volume is positive.
    void set_orientation() {
        double vol = 0.0;
        for (size_t i = 0; i < tris.size(); i++) vol += tris[i].pa.dot(tris[i].pb.cross(tris[i].pc));
        orient_sign = (vol >= 0.0) ? 1.0 : -1.0;
    }

    // objects.h:269-316. best_id reports which triangle produced `len` (not in the reference; for tests).
    int intersect_subtree(const Vec3 &o, const Vec3 &d, double &len, Vec3 &normalvector, int curID, int &best_id, KDCounters *kc) const {
        if (kc) kc->node_visits++;
        if (!box_any_face_hit(kdnodes[curID].box, o, d)) return 0;
        const KDNode &nd = kdnodes[curID];
        if ((int)nd.tri.size() < Minkdsize) {
            double len_temp;
            Vec3 normalvector_temp;
            int counter = 0;
            len = doubleINF;
            for (size_t i = 0; i < nd.tri.size(); i++) {
                if (kc) kc->tri_tests++;
                if (tris[nd.tri[i]].intersect(o, d, len_temp, normalvector_temp)) {
                    if (len_temp < len) {
                        len = len_temp;
                        normalvector = normalvector_temp;
                        best_id = nd.tri[i];
                        counter++;
                    }
                }
            }
            return counter;
        } else {
            double lenleft = 0, lenright = 0;
            Vec3 normalvecleft, normalvecright;
            int idl = -1, idr = -1;
            int numleft = intersect_subtree(o, d, lenleft, normalvecleft, nd.left, idl, kc);
            int numright = intersect_subtree(o, d, lenright, normalvecright, nd.right, idr, kc);
            if (numleft > 0) {
                if (numright > 0) {
                    if (lenleft < lenright) { len = lenleft; normalvector = normalvecleft; best_id = idl; }
                    else { len = lenright; normalvector = normalvecright; best_id = idr; }
                } else { len = lenleft; normalvector = normalvecleft; best_id = idl; }
            } else {
                if (numright > 0) { len = lenright; normalvector = normalvecright; best_id = idr; }
            }
            return numleft + numright;
        }
    }

    // objects.h:318-332 (+ the INTO_WINDING alternative)
    bool intersect(const Vec3 &o, const Vec3 &d, double &len, Vec3 &normalvector, IntoRule rule, int *tri_id = nullptr, KDCounters *kc = nullptr) const {
        if (kdnodes.empty()) return false;
        int best = -1;
        int counter = intersect_subtree(o, d, len, normalvector, 0, best, kc);
        if (tri_id) *tri_id = best;
        if (counter > 0) {
            if (rule == INTO_WINDING) {
                normalvector = normalvector * orient_sign;  // outward geometric normal
            } else if (counter % 2 == 0) {
                normalvector = normalvector * ((normalvector.dot(d) < 0) ? 1 : -1);  // origin outside
            } else {
                normalvector = normalvector * ((normalvector.dot(d) < 0) ? -1 : 1);  // origin inside
            }
            return true;
        }
        return false;
    }

    // Brute force over all triangles (the commented-out objects.h:406-432 variant) — property tests only.
    bool intersect_brute(const Vec3 &o, const Vec3 &d, double &len, Vec3 &normalvector, int *tri_id = nullptr) const {
        bool res = false;
        len = doubleINF;
        double lt; Vec3 nt;
        for (size_t i = 0; i < tris.size(); i++)
            if (tris[i].intersect(o, d, lt, nt) && lt < len) { len = lt; normalvector = nt; res = true; if (tri_id) *tri_id = (int)i; }
        return res;
    }
};

// ------------------------------------------------------------------------------------------------
// The Object plugin surface, objects.h:17-24.
// ------------------------------------------------------------------------------------------------
struct TraceCtx;  // forward (RNG for Bezier)
struct Object {
    virtual ~Object() {}
    virtual bool intersect(const Vec3 &rayorig, const Vec3 &raydir, double &len, Vec3 &normalvector, TraceCtx *tc) const = 0;
    virtual double getTransparency() const = 0;
    virtual double getReflection() const = 0;
    virtual Vec3 getSurfaceColor(const Vec3 &point) const = 0;
};

struct TraceCtx {
    Rng *rng = nullptr;          // the stream Bezier::intersect draws from (bezier.h:183,236,239)
    IntoRule into_rule = INTO_REFERENCE_PARITY;
    int last_prim = -1;          // triangle id of the last mesh hit (tests)
    KDCounters *kdc = nullptr;   // optional node-visit / triangle-test counters
};

// objects.h:26-89
struct Sphere : Object {
    Vec3 center; double radius, radius2; Vec3 surfaceColor; double transparency, reflection;
    Sphere(const Vec3 &c, double r, const Vec3 &sc, double refl = 0, double transp = 0)
        : center(c), radius(r), radius2(r * r), surfaceColor(sc), transparency(transp), reflection(refl) {}
    bool intersect(const Vec3 &rayorig, const Vec3 &raydir, double &len, Vec3 &normalvector, TraceCtx *) const override {
        Vec3 l = center - rayorig;
        double tca = l.dot(raydir);
        double l2 = l.dot(l);
        if (tca < 0 && l2 > radius2) return false;
        double d2 = l.dot(l) - tca * tca;
        if (d2 > radius2) return false;
        double thc = std::sqrt(radius2 - d2);
        double t0 = tca - thc;
        double t1 = tca + thc;
        if (t0 < 0) len = t1; else len = t0;
        Vec3 intersection = rayorig + raydir * len;
        Vec3 ret = intersection - center;
        normalvector = ret.normalize();
        return true;
    }
    double getTransparency() const override { return transparency; }
    double getReflection() const override { return reflection; }
    Vec3 getSurfaceColor(const Vec3 &) const override { return surfaceColor; }
};

// objects.h:335-476. Geometry arrives as triangles already transformed by the loader (see loaders below).
struct TriangleMesh : Object {
    Vec3 surfaceColor; double transparency, reflection; int objtype;
    KDTree kdtree;
    TriangleMesh(const std::vector<Triangle> &tris, const Vec3 &sc, double refl, double transp, int typeofdata)
        : surfaceColor(sc), transparency(transp), reflection(refl), objtype(typeofdata) { kdtree.build(tris); }
    bool intersect(const Vec3 &rayorig, const Vec3 &raydir, double &len, Vec3 &normalvector, TraceCtx *tc) const override {
        int id = -1;
        bool res = kdtree.intersect(rayorig, raydir, len, normalvector, tc ? tc->into_rule : INTO_REFERENCE_PARITY, &id, tc ? tc->kdc : nullptr);
        if (tc) tc->last_prim = id;
        if (objtype == 2)  // objects.h:434-436 (applied even when res is false; harmless)
            normalvector = normalvector * ((normalvector.dot(Vec3(0, 1, 0)) > 0) ? 1 : -1);
        return res;
    }
    double getTransparency() const override { return transparency; }
    double getReflection() const override { return reflection; }
    Vec3 getSurfaceColor(const Vec3 &) const override { return surfaceColor; }
};

// objects.h:478-548 (incl. the displaced height-field "bump mapping" mesh, :482-503)
struct Plane : Object {
    Vec3 normal, position, surfaceColor; double transparency, reflection;
    Texture texture;
    KDTree bumpmapping;
    Plane(const Vec3 &p, const Vec3 &n, const Vec3 &sc, double refl = 0, double transp = 0, const Texture &tx = Texture())
        : normal(n), position(p), surfaceColor(sc), transparency(transp), reflection(refl), texture(tx) {
        if (std::fabs(n.y - 1.0) < 1e-5 && texture.isbump && texture.htexture) {
            std::vector<Triangle> tris;
            bump_triangles(texture, position, tris);
            bumpmapping.build(tris);
        }
    }
    // objects.h:485-500. Note the unsigned arithmetic of size()/step-1 in the reference: for H,W >= 3 it is the
    // plain integer H/3-1.
    static void bump_triangles(const Texture &texture, const Vec3 &position, std::vector<Triangle> &tris) {
        int step = 3;
        int H = texture.H, W = texture.W;
        for (int i = 0; i < H / step - 1; i++) {
            for (int j = 0; j < W / step - 1; j++) {
                double x1 = texture.position.x + texture.lenx * j * step / W;
                double x2 = texture.position.x + texture.lenx * (j + 1) * step / W;
                double y1 = texture.position.z + texture.leny * i * step / H;
                double y2 = texture.position.z + texture.leny * (i + 1) * step / H;
                Vec3 a = Vec3(x1, texture.height[(size_t)(i * step) * W + j * step] + position.y, y1);
                Vec3 b = Vec3(x2, texture.height[(size_t)(i * step) * W + (j + 1) * step] + position.y, y1);
                Vec3 c = Vec3(x1, texture.height[(size_t)((i + 1) * step) * W + j * step] + position.y, y2);
                Vec3 d = Vec3(x2, texture.height[(size_t)((i + 1) * step) * W + (j + 1) * step] + position.y, y2);
                tris.push_back(Triangle(a, b, c));
                tris.push_back(Triangle(d, b, c));
            }
        }
    }
    // objects.h:505-524
    bool intersect(const Vec3 &rayorig, const Vec3 &raydir, double &len, Vec3 &normalvector, TraceCtx *tc) const override {
        Vec3 d = position - rayorig;
        len = d.dot(normal) / raydir.dot(normal);
        if (len > 0) {
            normalvector = normal;
            double lenp;
            Vec3 normalp;
            if (tc) tc->last_prim = -1;
            int id = -1;
            if (texture.isbump && std::fabs(normal.y - 1) < 1e-5 &&
                bumpmapping.intersect(rayorig, raydir, lenp, normalp, tc ? tc->into_rule : INTO_REFERENCE_PARITY, &id, tc ? tc->kdc : nullptr)) {
                if (lenp < len && lenp > 0) {
                    len = lenp;
                    normalvector = normalp;
                    if (tc) tc->last_prim = id;
                }
            }
            return true;
        }
        return false;
    }
    double getTransparency() const override { return transparency; }
    double getReflection() const override { return reflection; }
    // objects.h:533-539
    Vec3 getSurfaceColor(const Vec3 &point) const override {
        Vec3 color;
        if (texture.color(point, color)) return color;
        return surfaceColor;
    }
};

// ------------------------------------------------------------------------------------------------
// headers/bezier.h — surface of revolution about y, randomised multi-start Newton.
// ------------------------------------------------------------------------------------------------
static const double Cni[7][7] = {{1, 0, 0, 0, 0, 0, 0}, {1, 1, 0, 0, 0, 0, 0}, {1, 2, 1, 0, 0, 0, 0}, {1, 3, 3, 1, 0, 0, 0},
                                 {1, 4, 6, 4, 1, 0, 0}, {1, 5, 10, 10, 5, 1, 0}, {1, 6, 15, 20, 15, 6, 1}};
static const int NEWTON_MAX_ITER = 100;          // bezier.h:25
static const double NEWTON_STOP_EPS = 1e-6;      // bezier.h:26
static const int num_of_samples_newton = 10;     // bezier.h:27

// bezier.h:30-35
inline double Bern(int n, int i, double t) {
    if (i > n || i < 0) return 0;
    return Cni[n][i] * std::pow(1 - t, n - i) * std::pow(t, i);
}
// bezier.h:37-40
inline double dBern(int n, int i, double t) { return Bern(n - 1, i - 1, t) * (double)i - Bern(n - 1, i, t) * (double)(n - i); }

struct Bezier : Object {
    std::vector<Vec3> cpoints; Vec3 position, surfaceColor; double transparency, reflection;
    Box6 box;
    // bezier.h:44-71
    Bezier(const std::vector<Vec3> &points, const Vec3 &pos, const Vec3 &sc, double refl = 0, double transp = 0)
        : cpoints(points), position(pos), surfaceColor(sc), transparency(transp), reflection(refl) {
        double max_z = -doubleINF, max_y = -doubleINF, min_y = doubleINF;
        for (size_t i = 0; i < cpoints.size(); i++) {
            if (cpoints[i].z > max_z) max_z = cpoints[i].z;
            if (cpoints[i].y > max_y) max_y = cpoints[i].y;
            if (cpoints[i].y < min_y) min_y = cpoints[i].y;
        }
        box.xmax = max_z + position.x; box.xmin = -max_z + position.x;
        box.ymax = max_y + position.y; box.ymin = min_y + position.y;
        box.zmax = max_z + position.z; box.zmin = -max_z + position.z;
    }
    // bezier.h:72-126: same six face tests as the KD node (only the boolean is used by intersect()).
    bool intersect_with_box(const Vec3 &o, const Vec3 &d) const { return box_any_face_hit(box, o, d); }
    // bezier.h:127-134
    Vec3 valueP(double u) const {
        Vec3 res;
        int n = (int)cpoints.size();
        for (int i = 0; i < n; i++) res = res + cpoints[i] * Bern(n - 1, i, u);
        return res;
    }
    // bezier.h:135-142
    Vec3 gradP(double u) const {
        Vec3 res;
        int n = (int)cpoints.size();
        for (int i = 0; i < n; i++) res = res + cpoints[i] * dBern(n - 1, i, u);
        return res;
    }
    // bezier.h:144-149, paras = (t,u,theta)
    Vec3 funcValue(const Vec3 &paras, const Vec3 &rayorig, const Vec3 &raydir) const {
        Vec3 temp = valueP(paras.y);
        temp.x = temp.z * std::sin(paras.z);
        temp.z *= std::cos(paras.z);
        return rayorig + raydir * paras.x - position - temp;
    }
    // bezier.h:150-162
    void gradValue(const Vec3 &paras, const Vec3 &, const Vec3 &raydir, Vec3 &resa, Vec3 &resb, Vec3 &resc) const {
        resa = raydir;
        Vec3 temp1 = gradP(paras.y);
        Vec3 temp2 = valueP(paras.y);
        resb.x = -std::sin(paras.z) * temp1.z;
        resb.y = -temp1.y;
        resb.z = -std::cos(paras.z) * temp1.z;
        resc.x = -std::cos(paras.z) * temp2.z;
        resc.y = 0;
        resc.z = std::sin(paras.z) * temp2.z;
    }
    // bezier.h:163-214
    Vec3 newtonMethod(const Vec3 &initial, const Vec3 &rayorig, const Vec3 &raydir, Rng &rng) const {
        Vec3 res = initial;
        int counter = 0;
        Vec3 a, b, c, d, e, f;
        Vec3 funcval = funcValue(res, rayorig, raydir);
        while (funcval.norm() > NEWTON_STOP_EPS && counter < NEWTON_MAX_ITER) {
            counter++;
            gradValue(res, rayorig, raydir, a, b, c);
            bool flag = inv(a, b, c, d, e, f);
            if (!flag) {
                // bezier.h:183. C++ leaves the evaluation order of the three constructor arguments unspecified;
                // g++ evaluates them right to left, which is what oracle/_ref observes; reproduce that.
                double r3 = uniform_sampling_zeroone(rng);
                double r2 = uniform_sampling_zeroone(rng);
                double r1 = uniform_sampling_zeroone(rng);
                res = res + Vec3(r1, r2, r3) * 0.2 - 0.1;
            }
            res = res - matrixVectorProduct(d, e, f, funcval);
            funcval = funcValue(res, rayorig, raydir);
        }
        return res;
    }
    // bezier.h:215-224
    Vec3 normalvec(const Vec3 &paras) const {
        Vec3 resp = gradP(paras.y).normalize();
        Vec3 res;
        res.x = resp.x;
        res.z = resp.y;
        res.y = -resp.z;
        res.x = res.z * std::sin(paras.z);
        res.z = res.z * std::cos(paras.z);
        return res;
    }
    // bezier.h:225-290
    bool intersect(const Vec3 &rayorig, const Vec3 &raydir, double &len, Vec3 &normalvector, TraceCtx *tc) const override {
        if (!intersect_with_box(rayorig, raydir)) return false;
        Rng &rng = *tc->rng;
        bool flag = false;
        double b, t;
        len = doubleINF;
        for (int i = 0; i < num_of_samples_newton; i++) {
            b = uniform_sampling_zeroone(rng);
            t = 20 + 10 * uniform_sampling_zeroone(rng);
            Vec3 point = rayorig + raydir * t;
            point = point - position;
            double theta;
            if (point.z < 0) theta = 3.14159265 + std::atan(point.x / point.z);
            else theta = std::atan(point.x / point.z);
            Vec3 initial = Vec3(t, b, theta);
            Vec3 res = newtonMethod(initial, rayorig, raydir, rng);
            if ((funcValue(res, rayorig, raydir).norm() < 1e-4) && (res.x > 0) && (res.y <= 1) && (res.y >= 0)) {
                if (res.x < len) {
                    len = res.x;
                    normalvector = normalvec(res);
                    flag = true;
                }
            }
        }
        normalvector = normalvector * ((normalvector.dot(raydir) < 0) ? 1 : -1);
        double newt = (box.ymax - rayorig.y);
        if (newt > 0.1) {
            newt = newt / raydir.y;
            Vec3 newpoint = rayorig + raydir * newt;
            double rz = cpoints[cpoints.size() - 1].z;
            if ((newpoint.x - position.x) * (newpoint.x - position.x) + (newpoint.z - position.z) * (newpoint.z - position.z) <= rz * rz) {
                len = newt;
                normalvector = Vec3(0, 1, 0);
            }
        }
        return flag;
    }
    double getTransparency() const override { return transparency; }
    double getReflection() const override { return reflection; }
    Vec3 getSurfaceColor(const Vec3 &) const override { return surfaceColor; }
};

// ------------------------------------------------------------------------------------------------
// headers/hitpoints.h:6-19, headers/hash.h:22-54
// ------------------------------------------------------------------------------------------------
struct Hitpoint {
    Vec3 f, pos, normal, flux;
    double r2 = 0;
    int n = 0, h = 0, w = 0;
    // oracle-only bookkeeping (not in the reference): creation sequence number and per-round accumulators (U2)
    uint32_t seq = 0;   // creation order within the eye pass (main.cpp:185-187 loop order, DFS inside a pixel)
    uint32_t code = 0;  // DFS split code, 4 bits MSB first
    uint64_t path = 0;  // pixel*samples + sample
    Vec3 dflux;
    int m = 0;
};

struct Hashtable {
    int hashsize; int num_of_cell_per_dim; double celllength;
    std::vector<std::vector<Hitpoint>> hashtable;
    static constexpr double SIZE_OF_SCENE = 70.0, XMIN = -35.0, YMIN = -35.0, ZMIN = -15.0;  // hash.h:11-18
    // hash.h:22-30
    Hashtable(int hashsize_, double celllength_, bool alloc = true) : hashsize(hashsize_), celllength(celllength_) {
        num_of_cell_per_dim = (int)(std::ceil(SIZE_OF_SCENE / celllength));
        celllength = SIZE_OF_SCENE / num_of_cell_per_dim;
        if (alloc) hashtable.assign((size_t)hashsize, std::vector<Hitpoint>());
    }
    // hash.h:35-37 — the int products wrap (two's complement in practice); computed in uint32 to be UB-free.
    unsigned int hash(int ix, int iy, int iz) const {
        uint32_t v = ((uint32_t)ix * 73856093u) ^ ((uint32_t)iy * 19349663u) ^ ((uint32_t)iz * 83492791u);
        return v % (uint32_t)hashsize;
    }
    // hash.h:38-42
    void compute_coord(double x, double y, double z, int &ix, int &iy, int &iz) const {
        ix = (int)std::floor((x - XMIN) / celllength);
        iy = (int)std::floor((y - YMIN) / celllength);
        iz = (int)std::floor((z - ZMIN) / celllength);
    }
    // hash.h:43-54
    void insert(const Hitpoint &hp) {
        int ix, iy, iz;
        compute_coord(hp.pos.x, hp.pos.y, hp.pos.z, ix, iy, iz);
        hashtable[hash(ix, iy, iz)].push_back(hp);
    }
};

// ------------------------------------------------------------------------------------------------
// main.cpp:24-36, 177-184, 222-224 — every compile-time constant of the reference, as parameters.
// ------------------------------------------------------------------------------------------------
enum UpdateMode {
    UPDATE_U1_PER_PHOTON = 0,  // main.cpp:119-122, the reference's rule
    UPDATE_U2_PER_ROUND = 1    // accumulate a round against the round-start radius, then update once
};

struct Config {
    int width = 1024, height = 768;           // main.cpp:28-29
    int max_depth = 5;                        // main.cpp:35
    double alpha = 0.7;                       // main.cpp:36
    int num_of_samples = 1;                   // main.cpp:177
    double focus_plane = 20.0, lens_radius = 1.5;  // main.cpp:178-179
    int use_dof = 0;                          // 1: trace the thin-lens ray of main.cpp:207 instead of :209
    int consume_dof_rng = 1;                  // main.cpp:205 draws the lens sample even when unused (libc mode only)
    Vec3 lightorg = Vec3(0, 19.999, 20);      // main.cpp:180
    Vec3 camorg = Vec3(0, 0, -10);            // main.cpp:181
    int hashsize = 1000001;                   // main.cpp:184
    UpdateMode update = UPDATE_U1_PER_PHOTON;
    IntoRule into_rule = INTO_REFERENCE_PARITY;
    uint64_t seed = 20261018ull;
};

struct Counters {
    uint64_t eye_segments = 0, photon_segments = 0, diffuse_hits = 0, bucket_probes = 0, nonempty_probes = 0, candidates = 0,
             deposits = 0, misses = 0;
};

static const double EPS = 1e-4;               // main.cpp:24
static const double INF = 1e10;               // main.cpp:25
static const double PI = 3.14159265358979;    // main.cpp:26

struct HitRecord {  // result of the closest-hit loop, main.cpp:50-76
    int id = -1; double t = 0; Vec3 n_raw, n_ff; int into = 1; int prim = -1;
};

struct Renderer {
    Config cfg;
    std::vector<Object *> objs;
    Hashtable *htable = nullptr;
    Counters ctr;
    KDCounters kdc;
    uint32_t next_seq = 0;
    bool owns_htable = true;
    // Oracle-only: when set, the eye pass appends its hitpoints here (creation order) instead of inserting them, so that row blocks can be
    // traced by several threads and merged afterwards in the reference's creation order (main.cpp:185-187: h outer, w inner).
    std::vector<Hitpoint> *sink = nullptr;

    ~Renderer() { if (owns_htable) delete htable; }
    // A shallow per-thread view (own counters, shared scene and hitpoints) for the OpenMP photon loop.
    Renderer worker() const {
        Renderer w;
        w.cfg = cfg; w.objs = objs; w.htable = htable; w.owns_htable = false;
        return w;
    }

    // main.cpp:50-76 — linear closest hit, strict <, first object wins ties; face-forward + into.
    bool closest_hit(const Vec3 &org, const Vec3 &dir, HitRecord &hr, TraceCtx &tc) const {
        double len = 0;
        int id = -1;
        Vec3 normalvec, temp;
        double nearest = INF;
        int prim = -1;
        for (size_t i = 0; i < objs.size(); i++) {
            tc.last_prim = -1;
            if (objs[i]->intersect(org, dir, len, temp, &tc)) {
                if (len < nearest) {
                    id = (int)i;
                    nearest = len;
                    normalvec = temp;
                    prim = tc.last_prim;
                }
            }
        }
        hr.id = id;
        if (id == -1) return false;
        hr.t = nearest;
        hr.n_raw = normalvec;
        hr.prim = prim;
        hr.into = 1;
        if (normalvec.dot(dir) > 0) {
            normalvec = -normalvec;
            hr.into = 0;
        }
        hr.n_ff = normalvec;
        return true;
    }

    // main.cpp:42-167. `path` / `rng` carry the random stream: in libc mode one global stream is consumed in
    // call order; in Philox mode the stream is re-keyed at every segment with dim = depth+1.
    void trace(const Vec3 &org, const Vec3 &dir, Vec3 flux, Vec3 adj, bool flag, int depth, int x, int y, Rng &rng,
               uint64_t path, uint32_t dfs_code) {
        if (depth >= cfg.max_depth) return;
        rng.reseed(cfg.seed, flag ? PASS_EYE : PASS_PHOTON, path, (uint32_t)depth + 1);
        Rng bez_rng = rng;  // Bezier::intersect consumes the same libc stream; in Philox mode it gets its own pass key
        TraceCtx tc;
        tc.into_rule = cfg.into_rule;
        tc.kdc = &kdc;
        if (rng.mode == 1) {
            bez_rng.reseed(cfg.seed, PASS_BEZIER, path, (uint32_t)depth * 2 + (flag ? 0 : 1));
            tc.rng = &bez_rng;
        } else {
            tc.rng = &rng;
        }
        if (flag) ctr.eye_segments++; else ctr.photon_segments++;
        HitRecord hr;
        if (!closest_hit(org, dir, hr, tc)) { ctr.misses++; return; }
        const Object *obj = objs[hr.id];
        double nearest = hr.t;
        Vec3 intersection = org + dir * nearest;
        bool into = hr.into != 0;
        Vec3 normalvec_old = hr.n_raw;
        Vec3 normalvec = hr.n_ff;
        Vec3 f = obj->getSurfaceColor(intersection);
        double p = max3(f.x, f.y, f.z);

        if (obj->getReflection() < EPS && obj->getTransparency() < EPS) {
            double r = 200.0 / cfg.height;
            if (flag) {  // main.cpp:85-100
                Hitpoint hp;
                hp.f = f * adj;
                hp.pos = intersection;
                hp.normal = normalvec;
                hp.w = x;
                hp.h = y;
                hp.flux = Vec3();
                hp.r2 = r * r;
                hp.n = 0;
                hp.seq = next_seq++;
                hp.code = dfs_code & 15u;
                hp.path = path;
                if (sink) sink->push_back(hp);
                else htable->insert(hp);
            } else {  // main.cpp:101-128
                ctr.diffuse_hits++;
                int ix, iy, iz;
                htable->compute_coord(intersection.x, intersection.y, intersection.z, ix, iy, iz);
                ix -= 1; iy -= 1; iz -= 1;
                for (int idx = 0; idx < 3; idx++)
                    for (int idy = 0; idy < 3; idy++)
                        for (int idz = 0; idz < 3; idz++) {
                            int hashid = (int)htable->hash(ix + idx, iy + idy, iz + idz);
                            std::vector<Hitpoint> &bucket = htable->hashtable[hashid];
                            ctr.bucket_probes++;
                            if (!bucket.empty()) ctr.nonempty_probes++;
                            for (size_t i = 0; i < bucket.size(); i++) {
                                Hitpoint &hp = bucket[i];
                                ctr.candidates++;
                                Vec3 d = hp.pos - intersection;
                                if ((hp.normal.dot(normalvec) > EPS) && (d.dot(d) <= hp.r2)) {
                                    ctr.deposits++;
                                    if (cfg.update == UPDATE_U1_PER_PHOTON) {  // main.cpp:119-122
                                        double g = (hp.n * cfg.alpha + cfg.alpha) / (hp.n * cfg.alpha + 1.0);
                                        hp.r2 *= g;
                                        hp.n++;
                                        hp.flux = (hp.flux + hp.f.mul(flux) * (1.0 / PI)) * g;
                                    } else {  // U2: r2 unchanged within the round
                                        Vec3 c = hp.f.mul(flux) * (1.0 / PI);
#ifdef _OPENMP
#pragma omp atomic
                                        hp.dflux.x += c.x;
#pragma omp atomic
                                        hp.dflux.y += c.y;
#pragma omp atomic
                                        hp.dflux.z += c.z;
#pragma omp atomic
                                        hp.m++;
#else
                                        hp.dflux = hp.dflux + c;
                                        hp.m++;
#endif
                                    }
                                }
                            }
                        }
                Vec3 newdir = uniform_sampling_halfsphere(rng, normalvec);
                trace(intersection, newdir, f * flux * (1.0 / p), adj, flag, depth + 1, x, y, rng, path, dfs_code);
            }
        } else if (obj->getTransparency() < EPS) {  // main.cpp:129-134 mirror
            Vec3 newdir = dir - normalvec * 2.0 * normalvec.dot(dir);
            double refl = obj->getReflection();
            intersection = intersection + normalvec * EPS;
            trace(intersection, newdir, f * flux * refl, f * adj * refl, flag, depth + 1, x, y, rng, path, dfs_code);
        } else {  // main.cpp:135-166 glass
            double nc = 1.0, nt = 1.33, nnt = into ? nc / nt : nt / nc, ddn = dir.dot(normalvec), cos2t;
            Vec3 refl_dir = dir - normalvec_old * 2.0 * normalvec_old.dot(dir);
            if ((cos2t = 1 - nnt * nnt * (1 - ddn * ddn)) < 0) {
                trace(intersection + normalvec * EPS, refl_dir, flux, adj, flag, depth + 1, x, y, rng, path, dfs_code);
                return;
            }
            Vec3 refr_dir = (dir * nnt - normalvec_old * ((into ? 1 : -1) * (ddn * nnt + std::sqrt(cos2t)))).normalize();
            double a = nt - nc, b = nt + nc, R0 = a * a / (b * b), c = 1 - (into ? -ddn : refr_dir.dot(normalvec_old));
            double Re = R0 + (1 - R0) * c * c * c * c * c;
            Vec3 fa = f.mul(adj);
            if (flag) {
                // DFS code: one bit per split, MSB first, 4 bits: reflect = 0, refract = 1
                int nsplit_shift = 3 - split_count(dfs_code);
                uint32_t code_refl = bump_split(dfs_code, 0, nsplit_shift);
                uint32_t code_refr = bump_split(dfs_code, 1, nsplit_shift);
                trace(intersection + normalvec * EPS, refl_dir, flux, fa * Re, flag, depth + 1, x, y, rng, path, code_refl);
                trace(intersection - normalvec * EPS, refr_dir, flux, fa * (1 - Re), flag, depth + 1, x, y, rng, path, code_refr);
            } else {
                if (uniform_sampling_zeroone(rng) < 0.5) {
                    trace(intersection + normalvec * EPS, refl_dir, flux, fa * Re * 0.3, flag, depth + 1, x, y, rng, path, dfs_code);
                } else {
                    trace(intersection - normalvec * EPS, refr_dir, flux, fa * (1 - Re * 0.3), flag, depth + 1, x, y, rng, path, dfs_code);
                }
            }
        }
    }
    // dfs_code packs (nsplits << 4) | bits
    static int split_count(uint32_t code) { return (int)(code >> 4); }
    static uint32_t bump_split(uint32_t code, int bit, int shift) {
        uint32_t bits = code & 15u;
        int ns = (int)(code >> 4);
        if (shift >= 0) bits |= (uint32_t)bit << shift;
        return ((uint32_t)(ns + 1) << 4) | bits;
    }

    // main.cpp:183-219 — eye pass. Rows [y0,y1) (tile sharding hook; reference = full image).
    void eye_pass(Rng &rng, int y0 = 0, int y1 = -1) {
        if (y1 < 0) y1 = cfg.height;
        double r = 200.0 / cfg.height;
        if (!htable) htable = new Hashtable(cfg.hashsize, r);
        const int width = cfg.width, height = cfg.height;
        for (int h = y0; h < y1; h++) {
            for (int w = 0; w < width; w++) {
                double x = (2.0 * ((double)w / width) - 1) * 10.0;
                double y = (2.0 * ((double)h / height) - 1) * 10.0 * height / width;
                Vec3 dir = (Vec3(x, y, 0) - cfg.camorg).normalize();
                Vec3 point_on_focus = dir * ((cfg.focus_plane - cfg.camorg.z) / dir.z) + cfg.camorg;
                for (int j = 0; j < cfg.num_of_samples; j++) {
                    uint64_t path = ((uint64_t)h * width + w) * (uint64_t)cfg.num_of_samples + j;
                    Vec3 neworg = cfg.camorg, newdir = dir;
                    if (cfg.use_dof || (rng.mode == 0 && cfg.consume_dof_rng)) {
                        rng.reseed(cfg.seed, PASS_EYE, path, 0);
                        neworg = cfg.camorg + uniform_sampling_circle(rng, cfg.lens_radius);
                        newdir = (point_on_focus - neworg).normalize();
                    }
                    if (cfg.use_dof) trace(neworg, newdir, Vec3(), Vec3(1, 1, 1), true, 0, w, h, rng, path, 0);
                    else trace(cfg.camorg, dir, Vec3(), Vec3(1, 1, 1), true, 0, w, h, rng, path, 0);
                }
            }
        }
    }

    // main.cpp:231-247 — one photon with global index `index`.
    void photon(Rng &rng, uint64_t index) {
        rng.reseed(cfg.seed, PASS_PHOTON, index, 0);
        double a = uniform_sampling_zeroone(rng) * 4 - 2;
        double b = uniform_sampling_zeroone(rng) * 4 - 2;
        Vec3 disturbance = Vec3(a, 0, b);
        Vec3 dir = uniform_sampling_sphere(rng);
        trace(cfg.lightorg + disturbance, dir, Vec3(700, 700, 700) * (PI * 4.0), Vec3(1, 1, 1), false, 0, 0, 0, rng, index, 0);
    }

    // r2' = r2 (n a + a M)/(n a + M); flux' = (flux + dflux) r2'/r2; n' = n + M.
    void round_update() {
        for (auto &b : htable->hashtable)
            for (auto &hp : b) {
                if (hp.m > 0) {
                    double na = hp.n * cfg.alpha;
                    double g = (na + cfg.alpha * hp.m) / (na + hp.m);
                    hp.flux = (hp.flux + hp.dflux) * g;
                    hp.r2 *= g;
                    hp.n += hp.m;
                }
                hp.dflux = Vec3();
                hp.m = 0;
            }
    }

    // main.cpp:252-258. n_emitted = num_photon*num_threads*num_of_samples as a double (the reference's int
    // product overflows above 2^31-1).
    void gather_image(double n_emitted, std::vector<Vec3> &image) const {
        image.assign((size_t)cfg.width * cfg.height, Vec3());
        for (size_t j = 0; j < htable->hashtable.size(); j++)
            for (size_t i = 0; i < htable->hashtable[j].size(); i++) {
                const Hitpoint &hp = htable->hashtable[j][i];
                Vec3 &px = image[(size_t)hp.h * cfg.width + hp.w];
                px = px + hp.flux * (1.0 / (PI * hp.r2 * n_emitted));
            }
    }
};

// ------------------------------------------------------------------------------------------------
// objects.h:338-403 — the three text loaders (typeofdata 0/1/2): z negated, then v*a+b.
// Parsed with a tolerant token scanner that accepts exactly what the reference's scanf formats accept
// on the shipped files. Returns triangles in file order.
// ------------------------------------------------------------------------------------------------
inline bool load_mesh_text(const char *filename, int typeofdata, double a, const Vec3 &b, std::vector<Triangle> &out) {
    FILE *fp = std::fopen(filename, "r");
    if (!fp) return false;
    auto xf = [&](const Vec3 &v) { return v * a + b; };
    if (typeofdata == 0) {
        double ax, ay, az, bx, by, bz, cx, cy, cz;
        while (std::fscanf(fp, " begin vertex %lf %lf %lf vertex %lf %lf %lf vertex %lf %lf %lf end", &ax, &ay, &az, &bx, &by, &bz, &cx, &cy, &cz) == 9)
            out.push_back(Triangle(xf(Vec3(ax, ay, -az)), xf(Vec3(bx, by, -bz)), xf(Vec3(cx, cy, -cz))));
    } else {
        int num = 0;
        double x, y, z;
        if (std::fscanf(fp, "%d", &num) != 1) { std::fclose(fp); return false; }
        std::vector<Vec3> vertices;
        for (int i = 0; i < num; i++) {
            if (std::fscanf(fp, " v %lf %lf %lf", &x, &y, &z) != 3) { std::fclose(fp); return false; }
            vertices.push_back(Vec3(x, y, -z));
        }
        if (typeofdata == 2) {  // optional vn / vt lines (objects.h:387-392); Mesh000.obj has none
            for (;;) {
                int ch;
                while ((ch = std::fgetc(fp)) != EOF && (ch == ' ' || ch == '\n' || ch == '\r' || ch == '\t')) {}
                if (ch == EOF) break;
                if (ch != 'v') { std::ungetc(ch, fp); break; }
                while ((ch = std::fgetc(fp)) != EOF && ch != '\n') {}
            }
        }
        if (std::fscanf(fp, "%d", &num) != 1) { std::fclose(fp); return false; }
        for (int i = 0; i < num; i++) {
            int id1, id2, id3, t1, t2, t3, t4, t5, t6;
            if (typeofdata == 1) {
                if (std::fscanf(fp, " f %d %d %d", &id1, &id2, &id3) != 3) break;
            } else {
                if (std::fscanf(fp, " f %d/%d/%d %d/%d/%d %d/%d/%d", &id1, &t1, &t2, &id2, &t3, &t4, &id3, &t5, &t6) != 9) break;
            }
            out.push_back(Triangle(xf(vertices[id1 - 1]), xf(vertices[id2 - 1]), xf(vertices[id3 - 1])));
        }
    }
    std::fclose(fp);
    return true;
}

}  // namespace orc

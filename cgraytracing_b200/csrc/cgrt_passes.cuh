// cgrt_passes.cuh — the wavefront passes: eye pass, hash grid, photon trace + deposit, round update, image gather.
// Replaces trace()/render() (main.cpp:42-266). Recursion becomes per-depth kernels over ray queues that are
// compacted with warp ballots (one atomic per warp); the hash table of vectors becomes a sorted array + cell-start
// table; photon deposits are atomics into per-round accumulators.
#pragma once
#include "cgrt_build.cuh"

namespace cgrt {

// Ray queue, structure of arrays. w = path weight: adj for eye rays (main.cpp:42 `adj`), flux for photons (`flux`).
struct RayQueue {
    double *ox, *oy, *oz, *dx, *dy, *dz, *wx, *wy, *wz;
    uint32_t *id;    // eye: path = (h*W + w)*samples + s ; photon: offset from the call's first photon index
    uint32_t *aux;   // eye: DFS split code (nsplit << 4 | bits)
};
// Diffuse photon hits waiting for the 27-cell gather (main.cpp:103-125)
struct DepositQueue {
    double *px, *py, *pz, *nx, *ny, *nz, *fx, *fy, *fz;
};

// Hitpoint hot record read per candidate in the gather: 64 bytes (two sectors).
struct __align__(16) HpHot {
    double px, py, pz, r2;
    double nx, ny, nz, pad;
};

struct PassParams {
    int width, height, max_depth, samples, use_dof;
    uint32_t hashsize;
    double celllength;     // Hashtable ctor result, hash.h:25-26
    double r2_init;        // (200/height)^2, main.cpp:84,94
    double alpha, focus_plane, lens_radius;
    double cam[3], light[3];
    uint64_t seed;
};

// One slot per active lane of the warp, claimed with a single atomic (ballot + popc prefix).
__device__ __forceinline__ unsigned int warp_claim(bool want, unsigned int *counter, unsigned int n = 1) {
    unsigned int active = __activemask();
    unsigned int mask = __ballot_sync(active, want);
    if (!want) return 0xffffffffu;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    unsigned int base = 0;
    // all wanting lanes request the same n here (n is uniform across the call sites)
    if (lane == leader) base = atomicAdd(counter, n * __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + n * __popc(mask & ((1u << lane) - 1u));
}

__device__ __forceinline__ void push_ray(const RayQueue &q, unsigned int s, d3 o, d3 d, d3 w, uint32_t id, uint32_t aux) {
    q.ox[s] = o.x; q.oy[s] = o.y; q.oz[s] = o.z;
    q.dx[s] = d.x; q.dy[s] = d.y; q.dz[s] = d.z;
    q.wx[s] = w.x; q.wy[s] = w.y; q.wz[s] = w.z;
    q.id[s] = id;
    if (q.aux) q.aux[s] = aux;
}

struct Counters {
    unsigned long long eye_segments, photon_segments, diffuse_hits, candidates, deposits;
};

// =================================================================================================================
// Eye pass, one kernel per depth level. Level 0 generates the camera rays (main.cpp:188-209) in registers.
// Hitpoint records (12 doubles): pos, normal, f*adj, bits(sortkey), bits(h<<32|w), 0.
// sortkey = bucket key << 32 | (path*16 + dfs bits): sorting by it gives the reference's bucket order (hash.h:52,
// main.cpp:252-254) whatever order the wavefront produced the hitpoints in.
// =================================================================================================================
template <bool FIRST>
__global__ void __launch_bounds__(128) eye_bounce_kernel(const __grid_constant__ SceneDev S, const __grid_constant__ PassParams P, int depth,
                                                         RayQueue qin, unsigned int n_in, int y0, RayQueue qout, unsigned int *n_out,
                                                         double *hp_rec, unsigned int *hp_count, unsigned int hp_cap, Counters *ctr) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    d3 o, d, adj;
    uint32_t path, code;
    if (FIRST) {
        uint32_t s = i % (uint32_t)P.samples;
        uint32_t pix = i / (uint32_t)P.samples;
        int w = (int)(pix % (uint32_t)P.width), h = y0 + (int)(pix / (uint32_t)P.width);
        path = ((uint32_t)h * (uint32_t)P.width + (uint32_t)w) * (uint32_t)P.samples + s;
        code = 0;
        d3 cam = mk(P.cam[0], P.cam[1], P.cam[2]);
        double x = (2.0 * ((double)w / P.width) - 1) * 10.0;
        double y = (2.0 * ((double)h / P.height) - 1) * 10.0 * P.height / P.width;
        d = normalize(mk(x, y, 0) - cam);
        o = cam;
        if (P.use_dof) {  // main.cpp:203-207
            d3 pof = d * ((P.focus_plane - cam.z) / d.z) + cam;
            Philox g;
            g.init(P.seed, PASS_EYE, (uint64_t)path, 0);
            o = cam + sample_circle(g, P.lens_radius);
            d = normalize(pof - o);
        }
        adj = mk(1, 1, 1);
    } else {
        o = mk(qin.ox[i], qin.oy[i], qin.oz[i]);
        d = mk(qin.dx[i], qin.dy[i], qin.dz[i]);
        adj = mk(qin.wx[i], qin.wy[i], qin.wz[i]);
        path = qin.id[i];
        code = qin.aux[i];
    }
    Hit hit;
    bool found = closest_hit<false>(S, o, d, hit, nullptr);
    {   // segments counter: one atomic per warp
        unsigned int act = __activemask();
        if ((threadIdx.x & 31) == (__ffs(act) - 1)) atomicAdd(&ctr->eye_segments, (unsigned long long)__popc(act));
    }
    int mat = -1;
    d3 X = mk(0, 0, 0), n_ff = mk(0, 0, 0), n_old = mk(0, 0, 0), f = mk(0, 0, 0);
    bool into = true;
    if (found) {
        const ObjDev &O = S.obj[hit.obj];
        X = o + d * hit.t;  // main.cpp:68
        n_old = hit.n;
        n_ff = hit.n;
        if (dot(n_ff, d) > 0) { n_ff = -n_ff; into = false; }  // main.cpp:73-76
        f = surface_color(S, hit.obj, X);
        mat = O.material;
    }
    bool cont = (depth + 1 < P.max_depth);

    // ---- diffuse: create the hitpoint (main.cpp:85-99)
    bool mk_hp = (mat == MAT_DIFFUSE);
    unsigned int hs = warp_claim(mk_hp, hp_count);
    if (mk_hp && hs < hp_cap) {
        int ix, iy, iz;
        cell_coord(X, P.celllength, ix, iy, iz);
        uint32_t key = cell_hash(ix, iy, iz, P.hashsize);
        uint64_t sortkey = ((uint64_t)key << 32) | (uint64_t)(path * 16u + (code & 15u));
        uint32_t pix = path / (uint32_t)P.samples;
        uint64_t hw = ((uint64_t)(pix / (uint32_t)P.width) << 32) | (uint64_t)(pix % (uint32_t)P.width);
        d3 fa = f * adj;
        double *r = hp_rec + (size_t)hs * 12;
        r[0] = X.x; r[1] = X.y; r[2] = X.z;
        r[3] = n_ff.x; r[4] = n_ff.y; r[5] = n_ff.z;
        r[6] = fa.x; r[7] = fa.y; r[8] = fa.z;
        r[9] = __longlong_as_double((long long)sortkey);
        r[10] = __longlong_as_double((long long)hw);
        r[11] = 0.0;
    }

    // ---- mirror / glass: children (main.cpp:129-157)
    int nchild = 0;
    d3 co[2], cd[2], cw[2];
    uint32_t ccode[2] = {code, code};
    if (cont && mat == MAT_MIRROR) {
        const ObjDev &O = S.obj[hit.obj];
        cd[0] = d - n_ff * 2.0 * dot(n_ff, d);
        co[0] = X + n_ff * CGRT_EPS;
        cw[0] = f * adj * O.refl;
        nchild = 1;
    } else if (cont && mat == MAT_GLASS) {
        double nc = 1.0, nt = 1.33, nnt = into ? nc / nt : nt / nc, ddn = dot(d, n_ff), cos2t;
        d3 refl_dir = d - n_old * 2.0 * dot(n_old, d);
        if ((cos2t = 1 - nnt * nnt * (1 - ddn * ddn)) < 0) {  // total internal reflection, main.cpp:144
            co[0] = X + n_ff * CGRT_EPS; cd[0] = refl_dir; cw[0] = adj;
            nchild = 1;
        } else {
            d3 refr_dir = normalize(d * nnt - n_old * ((into ? 1 : -1) * (ddn * nnt + sqrt(cos2t))));
            double a = nt - nc, b = nt + nc, R0 = a * a / (b * b), c = 1 - (into ? -ddn : dot(refr_dir, n_old));
            double Re = R0 + (1 - R0) * c * c * c * c * c;
            d3 fa = f * adj;
            int ns = (int)(code >> 4);
            int shift = 3 - ns;
            uint32_t bits = code & 15u;
            co[0] = X + n_ff * CGRT_EPS; cd[0] = refl_dir; cw[0] = fa * Re;
            ccode[0] = ((uint32_t)(ns + 1) << 4) | bits;
            co[1] = X - n_ff * CGRT_EPS; cd[1] = refr_dir; cw[1] = fa * (1 - Re);
            ccode[1] = ((uint32_t)(ns + 1) << 4) | (bits | (shift >= 0 ? (1u << shift) : 0u));
            nchild = 2;
        }
    }
    // mirrors/TIR want 1 slot, splits want 2: claim in two rounds so each claim is uniform
    unsigned int s0 = warp_claim(nchild >= 1, n_out);
    unsigned int s1 = warp_claim(nchild >= 2, n_out);
    if (nchild >= 1) push_ray(qout, s0, co[0], cd[0], cw[0], path, ccode[0]);
    if (nchild >= 2) push_ray(qout, s1, co[1], cd[1], cw[1], path, ccode[1]);
}

// =================================================================================================================
// Grid build helpers
// =================================================================================================================
__global__ void hp_extract_keys_kernel(const double *__restrict__ rec, unsigned int n, uint64_t *__restrict__ keys) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = (uint64_t)__double_as_longlong(rec[(size_t)i * 12 + 9]);
}

struct HpArrays {
    HpHot *hot;            // pos, r2, normal
    double *f;             // [n][4] f*adj (+pad)
    double *flux;          // [n][4] tau (+pad)
    int *cnt;              // accepted photon count n
    int *hw;               // [n][2] pixel (h, w)
    uint32_t *key;         // bucket key
    uint32_t *seq;         // path*16 + dfs bits
};

__global__ void hp_gather_sorted_kernel(const double *__restrict__ rec, const uint32_t *__restrict__ perm, unsigned int n, double r2_init,
                                        HpArrays A, uint64_t *__restrict__ pixkeys, int width) {
    unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double *r = rec + (size_t)perm[k] * 12;
    HpHot h;
    h.px = r[0]; h.py = r[1]; h.pz = r[2]; h.r2 = r2_init;
    h.nx = r[3]; h.ny = r[4]; h.nz = r[5]; h.pad = 0.0;
    A.hot[k] = h;
    A.f[4 * (size_t)k] = r[6]; A.f[4 * (size_t)k + 1] = r[7]; A.f[4 * (size_t)k + 2] = r[8]; A.f[4 * (size_t)k + 3] = 0.0;
    A.flux[4 * (size_t)k] = 0.0; A.flux[4 * (size_t)k + 1] = 0.0; A.flux[4 * (size_t)k + 2] = 0.0; A.flux[4 * (size_t)k + 3] = 0.0;
    A.cnt[k] = 0;
    uint64_t sk = (uint64_t)__double_as_longlong(r[9]);
    uint64_t hw = (uint64_t)__double_as_longlong(r[10]);
    A.key[k] = (uint32_t)(sk >> 32);
    A.seq[k] = (uint32_t)sk;
    int hh = (int)(hw >> 32), ww = (int)(uint32_t)hw;
    A.hw[2 * (size_t)k] = hh; A.hw[2 * (size_t)k + 1] = ww;
    pixkeys[k] = (uint64_t)hh * (uint64_t)width + (uint64_t)ww;
}

// start[v] = first sorted index whose key is >= v, for v in [0, nvals]; keys ascending.
__global__ void lower_bound_table_kernel(const uint32_t *__restrict__ keys32, const uint64_t *__restrict__ keys64, unsigned int n, unsigned int nvals,
                                         uint32_t *__restrict__ start) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    long long prev = (i == 0) ? -1 : (keys32 ? (long long)keys32[i - 1] : (long long)keys64[i - 1]);
    long long cur = (i == n) ? (long long)nvals : (keys32 ? (long long)keys32[i] : (long long)keys64[i]);
    for (long long v = prev + 1; v <= cur; v++) start[v] = i;
}

// =================================================================================================================
// Photon pass: trace kernel (emission K10 fused into depth 0) + deposit kernel.
// =================================================================================================================
template <bool FIRST, bool COUNT>
__global__ void __launch_bounds__(128) photon_trace_kernel(const __grid_constant__ SceneDev S, const __grid_constant__ PassParams P, int depth,
                                                           RayQueue qin, unsigned int n_in, uint64_t first_index, RayQueue qout,
                                                           unsigned int *n_out, DepositQueue dq, unsigned int *n_dq, Counters *ctr,
                                                           TravCounters *tcg) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    d3 o, d, flux;
    uint32_t off;
    if (FIRST) {  // main.cpp:240-246
        off = i;
        Philox g;
        g.init(P.seed, PASS_PHOTON, first_index + (uint64_t)i, 0);
        double a = g.u01() * 4 - 2;
        double b = g.u01() * 4 - 2;
        d = sample_sphere(g);
        o = mk(P.light[0], P.light[1], P.light[2]) + mk(a, 0, b);
        flux = mk(700, 700, 700) * (CGRT_PI * 4.0);
    } else {
        o = mk(qin.ox[i], qin.oy[i], qin.oz[i]);
        d = mk(qin.dx[i], qin.dy[i], qin.dz[i]);
        flux = mk(qin.wx[i], qin.wy[i], qin.wz[i]);
        off = qin.id[i];
    }
    Hit hit;
    TravCounters tcl;
    tcl.node_visits = 0; tcl.tri_tests = 0;
    bool found = closest_hit<COUNT>(S, o, d, hit, &tcl);
    if (COUNT) {
        atomicAdd(&tcg->node_visits, tcl.node_visits);
        atomicAdd(&tcg->tri_tests, tcl.tri_tests);
    }
    {
        unsigned int act = __activemask();
        if ((threadIdx.x & 31) == (__ffs(act) - 1)) atomicAdd(&ctr->photon_segments, (unsigned long long)__popc(act));
    }
    int mat = -1;
    d3 X = mk(0, 0, 0), n_ff = mk(0, 0, 0), n_old = mk(0, 0, 0), f = mk(0, 0, 0);
    bool into = true;
    if (found) {
        X = o + d * hit.t;
        n_old = hit.n;
        n_ff = hit.n;
        if (dot(n_ff, d) > 0) { n_ff = -n_ff; into = false; }
        f = surface_color(S, hit.obj, X);
        mat = S.obj[hit.obj].material;
    }
    // ---- diffuse hit: queue the deposit (main.cpp:103-125 happens in photon_deposit_kernel)
    bool dep = (mat == MAT_DIFFUSE);
    unsigned int ds = warp_claim(dep, n_dq);
    if (dep) {
        dq.px[ds] = X.x; dq.py[ds] = X.y; dq.pz[ds] = X.z;
        dq.nx[ds] = n_ff.x; dq.ny[ds] = n_ff.y; dq.nz[ds] = n_ff.z;
        dq.fx[ds] = flux.x; dq.fy[ds] = flux.y; dq.fz[ds] = flux.z;
    }
    // ---- continuation
    bool cont = found && (depth + 1 < P.max_depth);
    d3 no = X, nd = d, nf = flux;
    if (cont) {
        if (mat == MAT_DIFFUSE) {  // main.cpp:126-127: uniform hemisphere, origin NOT offset, flux * f / max(f)
            Philox g;
            g.init(P.seed, PASS_PHOTON, first_index + (uint64_t)off, (uint32_t)depth + 1);
            nd = sample_halfsphere(g, n_ff);
            double p = max3(f.x, f.y, f.z);
            nf = f * flux * (1.0 / p);
        } else if (mat == MAT_MIRROR) {  // main.cpp:131-134
            nd = d - n_ff * 2.0 * dot(n_ff, d);
            no = X + n_ff * CGRT_EPS;
            nf = f * flux * S.obj[hit.obj].refl;
        } else {  // glass, main.cpp:140-164: 50/50 roulette, flux unchanged
            double nc = 1.0, nt = 1.33, nnt = into ? nc / nt : nt / nc, ddn = dot(d, n_ff), cos2t;
            d3 refl_dir = d - n_old * 2.0 * dot(n_old, d);
            if ((cos2t = 1 - nnt * nnt * (1 - ddn * ddn)) < 0) {
                no = X + n_ff * CGRT_EPS; nd = refl_dir;
            } else {
                d3 refr_dir = normalize(d * nnt - n_old * ((into ? 1 : -1) * (ddn * nnt + sqrt(cos2t))));
                Philox g;
                g.init(P.seed, PASS_PHOTON, first_index + (uint64_t)off, (uint32_t)depth + 1);
                if (g.u01() < 0.5) { no = X + n_ff * CGRT_EPS; nd = refl_dir; }
                else { no = X - n_ff * CGRT_EPS; nd = refr_dir; }
            }
        }
    }
    unsigned int s = warp_claim(cont, n_out);
    if (cont) push_ray(qout, s, no, nd, nf, off, 0);
}

// One warp per diffuse photon hit. Lanes 0..26 look up the 3x3x3 cells (main.cpp:105-113); the candidates of all 27
// buckets are then scanned 32 at a time (main.cpp:114-116), one 64-byte hot record per lane. Two of the 27 cells
// hashing to the same bucket scan it twice, exactly like the reference.
// ACC: 0 = fp64 atomics {dflux.xyz, m}; 1 = one red.global.add.v4.f32.
template <int ACC>
__global__ void __launch_bounds__(256) photon_deposit_kernel(const __grid_constant__ PassParams P, DepositQueue dq, unsigned int n_dq,
                                                             const uint32_t *__restrict__ cell_start, const HpHot *__restrict__ hot,
                                                             const double *__restrict__ hp_f, void *__restrict__ acc, Counters *ctr) {
    const int lane = threadIdx.x & 31;
    unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned int nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long cand_total = 0, dep_total = 0;
    for (unsigned int r = warp; r < n_dq; r += nwarps) {
        d3 X = mk(dq.px[r], dq.py[r], dq.pz[r]);
        d3 nrm = mk(dq.nx[r], dq.ny[r], dq.nz[r]);
        d3 flux = mk(dq.fx[r], dq.fy[r], dq.fz[r]);
        int ix, iy, iz;
        cell_coord(X, P.celllength, ix, iy, iz);
        ix -= 1; iy -= 1; iz -= 1;
        uint32_t beg = 0, cnt = 0;
        if (lane < 27) {
            int idx = lane / 9, idy = (lane / 3) % 3, idz = lane % 3;  // idx outermost, idz innermost (main.cpp:110-112)
            uint32_t key = cell_hash(ix + idx, iy + idy, iz + idz, P.hashsize);
            beg = __ldg(cell_start + key);
            cnt = __ldg(cell_start + key + 1) - beg;
        }
        // exclusive prefix over lanes
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        uint32_t excl = incl - cnt;
        uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        cand_total += (lane == 0) ? total : 0;
        for (uint32_t c0 = 0; c0 < total; c0 += 32) {
            uint32_t c = c0 + lane;
            // find the cell whose [excl, excl+cnt) contains c: count lanes with excl <= c, via a ballot-free binary search
            int lo = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                int probe = lo + step;
                uint32_t e = __shfl_sync(0xffffffffu, excl, probe & 31);
                if (probe < 27 && e <= c) lo = probe;
            }
            // lo may point at an empty cell that shares excl with later ones; the last lane with excl <= c is the owner
            uint32_t e_lo = __shfl_sync(0xffffffffu, excl, lo);
            uint32_t b_lo = __shfl_sync(0xffffffffu, beg, lo);
            if (c < total) {
                uint32_t hidx = b_lo + (c - e_lo);
                const double2 *hp = reinterpret_cast<const double2 *>(hot + hidx);
                double2 a0 = __ldg(hp), a1 = __ldg(hp + 1), b0 = __ldg(hp + 2), b1 = __ldg(hp + 3);
                d3 hpos = mk(a0.x, a0.y, a1.x);
                double r2 = a1.y;
                d3 hn = mk(b0.x, b0.y, b1.x);
                d3 dd = hpos - X;
                if ((dot(hn, nrm) > CGRT_EPS) && (dot(dd, dd) <= r2)) {  // main.cpp:116
                    const double2 *fp = reinterpret_cast<const double2 *>(hp_f + 4 * (size_t)hidx);
                    double2 f0 = __ldg(fp), f1 = __ldg(fp + 1);
                    d3 cc = (mk(f0.x, f0.y, f1.x) * flux) * (1.0 / CGRT_PI);  // f.mul(flux) * (1/PI), main.cpp:122
                    if (ACC == 0) {
                        double *ap = reinterpret_cast<double *>(acc) + 4 * (size_t)hidx;
                        atomicAdd(ap, cc.x); atomicAdd(ap + 1, cc.y); atomicAdd(ap + 2, cc.z); atomicAdd(ap + 3, 1.0);
                    } else {
                        float *ap = reinterpret_cast<float *>(acc) + 4 * (size_t)hidx;
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ap), "f"((float)cc.x), "f"((float)cc.y),
                                     "f"((float)cc.z), "f"(1.0f)
                                     : "memory");
                    }
                    dep_total++;
                }
            }
        }
    }
    // counters: warp-reduce then one atomic per warp
    for (int o = 16; o > 0; o >>= 1) {
        cand_total += __shfl_xor_sync(0xffffffffu, cand_total, o);
        dep_total += __shfl_xor_sync(0xffffffffu, dep_total, o);
    }
    if (lane == 0) {
        if (cand_total) atomicAdd(&ctr->candidates, cand_total);
        if (dep_total) atomicAdd(&ctr->deposits, dep_total);
    }
}

// =================================================================================================================
// K12: per-round radius / flux update (main.cpp:119-122 batched over a round, SURVEY Q1 "U2"):
//   g = (n a + a M) / (n a + M);  flux = (flux + dflux) g;  r2 *= g;  n += M.   Clears the accumulators.
// =================================================================================================================
template <int ACC>
__global__ void round_update_kernel(unsigned int n, double alpha, HpArrays A, void *__restrict__ acc) {
    unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double dx, dy, dz, m;
    if (ACC == 0) {
        double4 *ap = reinterpret_cast<double4 *>(acc) + k;
        double4 a = *ap;
        dx = a.x; dy = a.y; dz = a.z; m = a.w;
        *ap = make_double4(0, 0, 0, 0);
    } else {
        float4 *ap = reinterpret_cast<float4 *>(acc) + k;
        float4 a = *ap;
        dx = a.x; dy = a.y; dz = a.z; m = a.w;
        *ap = make_float4(0, 0, 0, 0);
    }
    if (m > 0) {
        int cnt = A.cnt[k];
        double na = cnt * alpha;
        double g = (na + alpha * m) / (na + m);
        double *fl = A.flux + 4 * (size_t)k;
        fl[0] = (fl[0] + dx) * g;
        fl[1] = (fl[1] + dy) * g;
        fl[2] = (fl[2] + dz) * g;
        A.hot[k].r2 *= g;
        A.cnt[k] = cnt + (int)m;
    }
}

// =================================================================================================================
// K13: image gather (main.cpp:252-258). One thread per pixel walks that pixel's hitpoints in canonical order
// (pix_perm is a stable sort of the canonical order by pixel), so the fp64 sums equal the reference loop's.
// =================================================================================================================
__device__ __forceinline__ int gamma_corr(double x) { return int(pow(1 - exp(-x), 1 / 2.2) * 255 + .5); }  // util.h:45-47

__global__ void image_gather_kernel(int width, int height, double n_emitted, const uint32_t *__restrict__ pix_start,
                                    const uint32_t *__restrict__ pix_perm, const HpHot *__restrict__ hot, const double *__restrict__ flux,
                                    double *__restrict__ rgb, uint8_t *__restrict__ rgb8) {
    unsigned int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (unsigned int)(width * height)) return;
    d3 acc = mk(0, 0, 0);
    for (uint32_t j = pix_start[p]; j < pix_start[p + 1]; j++) {
        uint32_t k = pix_perm[j];
        double r2 = hot[k].r2;
        d3 fl = mk(flux[4 * (size_t)k], flux[4 * (size_t)k + 1], flux[4 * (size_t)k + 2]);
        acc = acc + fl * (1.0 / (CGRT_PI * r2 * n_emitted));
    }
    rgb[3 * (size_t)p] = acc.x; rgb[3 * (size_t)p + 1] = acc.y; rgb[3 * (size_t)p + 2] = acc.z;
    if (rgb8) {  // main.cpp:403-411: row i of the PNG is image[height-1-i]
        int h = p / width, w = p % width;
        size_t q = (size_t)(height - 1 - h) * width + w;
        rgb8[3 * q] = (uint8_t)(char)gamma_corr(acc.x);
        rgb8[3 * q + 1] = (uint8_t)(char)gamma_corr(acc.y);
        rgb8[3 * q + 2] = (uint8_t)(char)gamma_corr(acc.z);
    }
}

// =================================================================================================================
// Parity-hook kernels
// =================================================================================================================
template <bool COUNT>
__global__ void __launch_bounds__(128) intersect_batch_kernel(const __grid_constant__ SceneDev S, int64_t n, const double *__restrict__ org,
                                                              const double *__restrict__ dir, double *t, double *nrm, double *nrm_raw, int *obj,
                                                              int *into, int *prim, TravCounters *tc_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d3 o = mk(org[3 * i], org[3 * i + 1], org[3 * i + 2]), d = mk(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
    Hit h;
    TravCounters tc;
    tc.node_visits = 0; tc.tri_tests = 0;
    bool found = closest_hit<COUNT>(S, o, d, h, &tc);
    if (COUNT) {
        atomicAdd(&tc_out->node_visits, tc.node_visits);
        atomicAdd(&tc_out->tri_tests, tc.tri_tests);
    }
    d3 nf = h.n;
    int in = 1;
    if (found && dot(nf, d) > 0) { nf = -nf; in = 0; }
    if (obj) obj[i] = found ? h.obj : -1;
    if (t) t[i] = found ? h.t : 0.0;
    if (nrm) { nrm[3 * i] = found ? nf.x : 0; nrm[3 * i + 1] = found ? nf.y : 0; nrm[3 * i + 2] = found ? nf.z : 0; }
    if (nrm_raw) { nrm_raw[3 * i] = found ? h.n.x : 0; nrm_raw[3 * i + 1] = found ? h.n.y : 0; nrm_raw[3 * i + 2] = found ? h.n.z : 0; }
    if (into) into[i] = found ? in : 0;
    if (prim) prim[i] = found ? h.prim : -1;
}

__global__ void hash_keys_kernel(int64_t n, const double *__restrict__ pos, uint32_t hashsize, double celllength, uint32_t *key, int *ixyz) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ix, iy, iz;
    cell_coord(mk(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), celllength, ix, iy, iz);
    if (ixyz) { ixyz[3 * i] = ix; ixyz[3 * i + 1] = iy; ixyz[3 * i + 2] = iz; }
    if (key) key[i] = cell_hash(ix, iy, iz, hashsize);
}

__global__ void surface_color_kernel(const __grid_constant__ SceneDev S, int obj, int64_t n, const double *__restrict__ pos, double *col) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d3 c = surface_color(S, obj, mk(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
    col[3 * i] = c.x; col[3 * i + 1] = c.y; col[3 * i + 2] = c.z;
}

__global__ void sample_kernel(uint64_t seed, uint32_t pass, uint64_t path, uint32_t dim, int what, double a0, double a1, double a2, double *out) {
    Philox g;
    g.init(seed, pass, path, dim);
    d3 r;
    if (what == 0) r = sample_sphere(g);
    else if (what == 1) r = sample_halfsphere(g, mk(a0, a1, a2));
    else if (what == 2) r = sample_circle(g, a0);
    else { r.x = g.u01(); r.y = g.u01(); r.z = g.u01(); }
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

}  // namespace cgrt

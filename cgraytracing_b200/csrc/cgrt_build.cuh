// cgrt_build.cuh — hand-written device builders: LSD radix sort (keys + permutation), LBVH (Morton codes ->
// sort -> Karras hierarchy -> bottom-up refit -> top-down collapse into 128-byte four-box nodes), and the displaced height-field mesh.
// Replaces std::sort / KDTree::buildKdTree (objects.h:217-267), Plane's bump ctor (objects.h:482-503) and
// Hashtable::insert's bucket vectors (hash.h:43-54).
#pragma once
#include "cgrt_device.cuh"

namespace cgrt {

// =================================================================================================================
// Radix sort: stable LSD, 8-bit digits, 32- or 64-bit keys with a 32-bit payload (the permutation).
// Per pass: (1) per-tile digit histogram, (2) exclusive scan over [digit][tile], (3) stable scatter.
// =================================================================================================================
#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const K *__restrict__ keys, int64_t n, int shift, uint32_t *__restrict__ counts, int ntiles) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ITEMS; r++) {
        int64_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// Single-block exclusive scan of `m` counters (m = 256 * ntiles); one pass with a running carry.
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t *__restrict__ counts, int64_t m) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t base = 0; base < m; base += 1024) {
        int64_t i = base + threadIdx.x;
        uint32_t v = (i < m) ? counts[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t s = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            warp_sums[lane] = s;
        }
        __syncthreads();
        uint32_t prefix = carry + (wid > 0 ? warp_sums[wid - 1] : 0u) + x - v;
        if (i < m) counts[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
}

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const K *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                                K *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n, int shift,
                                                                const uint32_t *__restrict__ offsets, int ntiles, int first_pass) {
    __shared__ uint32_t running[256];          // per digit: global offset of this tile + keys already placed
    __shared__ uint32_t warp_cnt[RS_THREADS / 32][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    running[threadIdx.x] = offsets[(size_t)threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; w++) warp_cnt[w][threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ITEMS; r++) {
        int64_t i = base + r * RS_THREADS + threadIdx.x;
        bool valid = i < n;
        K k = valid ? keys_in[i] : (K)0;
        uint32_t v = valid ? (first_pass ? (uint32_t)i : vals_in[i]) : 0u;
        uint32_t dg = valid ? ((uint32_t)(k >> shift) & 255u) : 256u;  // 256: never matches a real digit
        uint32_t mask = __match_any_sync(0xffffffffu, dg);
        uint32_t rank = __popc(mask & ((1u << lane) - 1u));
        if (valid && rank == 0) warp_cnt[wid][dg] = __popc(mask);
        __syncthreads();
        if (valid) {
            uint32_t pos = running[dg] + rank;
            for (int w = 0; w < wid; w++) pos += warp_cnt[w][dg];
            keys_out[pos] = k;
            vals_out[pos] = v;
        }
        __syncthreads();
        {
            uint32_t s = 0;
#pragma unroll
            for (int w = 0; w < RS_THREADS / 32; w++) { s += warp_cnt[w][threadIdx.x]; warp_cnt[w][threadIdx.x] = 0; }
            running[threadIdx.x] += s;
        }
        __syncthreads();
    }
}

// =================================================================================================================
// LBVH
// =================================================================================================================
__device__ __forceinline__ uint32_t float_flip(float f) {  // order-preserving float -> uint
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_unflip(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// bounds[0..2] = min (flipped uint), bounds[3..5] = max, of triangle centroids
__global__ void lbvh_bounds_kernel(const double *__restrict__ tri9, int n, uint32_t *__restrict__ bounds) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = {3.0e38f, 3.0e38f, 3.0e38f}, C[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    if (i < n) {
        const double *t = tri9 + (size_t)i * 9;
        for (int a = 0; a < 3; a++) {
            float m = (float)((t[a] + t[3 + a] + t[6 + a]) * (1.0 / 3.0));
            c[a] = m; C[a] = m;
        }
    }
    for (int a = 0; a < 3; a++) {
        float lo = c[a], hi = C[a];
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&bounds[a], float_flip(lo));
            atomicMax(&bounds[3 + a], float_flip(hi));
        }
    }
}

__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void lbvh_morton_kernel(const double *__restrict__ tri9, int n, const uint32_t *__restrict__ bounds, uint64_t *__restrict__ keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *t = tri9 + (size_t)i * 9;
    uint32_t q[3];
    for (int a = 0; a < 3; a++) {
        float lo = float_unflip(bounds[a]), hi = float_unflip(bounds[3 + a]);
        float m = (float)((t[a] + t[3 + a] + t[6 + a]) * (1.0 / 3.0));
        float ext = hi - lo;
        float u = ext > 0.f ? (m - lo) / ext : 0.f;
        u = fminf(fmaxf(u * 1024.f, 0.f), 1023.f);
        q[a] = (uint32_t)u;
    }
    keys[i] = (uint64_t)((expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]));
}

// Sorted position k <- original triangle perm[k]: triangle record + padded float leaf box.
__global__ void lbvh_leaves_kernel(const double *__restrict__ tri9, const uint32_t *__restrict__ perm, int n, TriRec *__restrict__ tris,
                                   int *__restrict__ tri_id, float *__restrict__ box /* [2n-1][6], leaves at n-1+k */) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t src = perm[k];
    const double *t = tri9 + (size_t)src * 9;
    d3 pa = mk(t[0], t[1], t[2]), pb = mk(t[3], t[4], t[5]), pc = mk(t[6], t[7], t[8]);
    d3 e1 = pa - pb, e2 = pa - pc;                 // objects.h:98-99
    d3 nn = normalize(cross(pa - pb, pa - pc));    // objects.h:107
    TriRec R;
    R.pa[0] = pa.x; R.pa[1] = pa.y; R.pa[2] = pa.z;
    R.e1[0] = e1.x; R.e1[1] = e1.y; R.e1[2] = e1.z;
    R.e2[0] = e2.x; R.e2[1] = e2.y; R.e2[2] = e2.z;
    R.n[0] = nn.x; R.n[1] = nn.y; R.n[2] = nn.z;
    tris[k] = R;
    tri_id[k] = (int)src;
    float *b = box + (size_t)(n - 1 + k) * 6;
    const double pad = CGRT_BOX_PAD;  // covers the fp32 slab test's rounding (cgrt_device.cuh)
    for (int a = 0; a < 3; a++) {
        double lo = fmin(fmin(t[a], t[3 + a]), t[6 + a]), hi = fmax(fmax(t[a], t[3 + a]), t[6 + a]);
        b[a] = __double2float_rd(lo - pad - 1e-7 * fabs(lo));
        b[3 + a] = __double2float_ru(hi + pad + 1e-7 * fabs(hi));
    }
}

// Karras 2012: delta(i,j) = common prefix length of (key_i, i) and (key_j, j).
__device__ __forceinline__ int lbvh_delta(const uint64_t *__restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

// node ids: internal i in [0,n-1), leaf k -> n-1+k. child arrays hold these "box ids".
__global__ void lbvh_hierarchy_kernel(const uint64_t *__restrict__ keys, int n, int *__restrict__ left, int *__restrict__ right, int *__restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lc = (lo == gamma) ? (n - 1 + gamma) : gamma;
    int rc = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    left[i] = lc;
    right[i] = rc;
    parent[lc] = i;
    parent[rc] = i;
    if (i == 0) parent[0] = -1;
}

__global__ void lbvh_refit_kernel(int n, const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ parent,
                                  float *__restrict__ box, unsigned int *__restrict__ flags) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int node = parent[n - 1 + k];
    while (node >= 0) {
        unsigned int old = atomicAdd(&flags[node], 1u);
        if (old == 0) return;  // first arrival: the sibling subtree is not finished yet
        __threadfence();
        const volatile float *a = box + (size_t)left[node] * 6;
        const volatile float *b = box + (size_t)right[node] * 6;
        float *o = box + (size_t)node * 6;
        for (int c = 0; c < 3; c++) {
            o[c] = fminf(a[c], b[c]);
            o[3 + c] = fmaxf(a[3 + c], b[3 + c]);
        }
        __threadfence();
        node = parent[node];
    }
}

// Top-down collapse of the binary hierarchy into 4-wide nodes, one level per launch. wide node w in [begin, end) stands for the binary
// internal node queue[w]; its children are that node's two children, the one with the largest box area being replaced by its own two
// children until there are four (or no internal child is left). Internal children get the next free wide indices (atomic counter; the
// memory position of a node does not influence any result: slots inside a node are ordered by binary box id, which is deterministic).
__device__ __forceinline__ float lbvh_box_area(const float *__restrict__ b) {
    const float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
    return dx * dy + dy * dz + dz * dx;
}
__global__ void lbvh_collapse_kernel(int n, int begin, int end, const int *__restrict__ left, const int *__restrict__ right, const float *__restrict__ box,
                                     int *__restrict__ queue, int *__restrict__ counter, BvhNode4 *__restrict__ nodes4) {
    int w = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= end) return;
    const int b = queue[w];
    int c[4];
    int nc = 2;
    c[0] = left[b]; c[1] = right[b]; c[2] = -1; c[3] = -1;
    while (nc < 4) {
        int pick = -1;
        float best = -1.f;
        for (int j = 0; j < nc; j++) {
            if (c[j] >= n - 1) continue;  // a leaf
            float a = lbvh_box_area(box + (size_t)c[j] * 6);
            if (a > best) { best = a; pick = j; }
        }
        if (pick < 0) break;
        const int e = c[pick];
        c[pick] = left[e];
        c[nc++] = right[e];
    }
    // slots in ascending box id (insertion sort of <= 4 entries)
    for (int i = 1; i < nc; i++) {
        int v = c[i], j = i - 1;
        while (j >= 0 && c[j] > v) { c[j + 1] = c[j]; j--; }
        c[j + 1] = v;
    }
    BvhNode4 N;
    for (int j = 0; j < 4; j++) {
        if (j < nc) {
            const float *bb = box + (size_t)c[j] * 6;
            N.lox[j] = bb[0]; N.loy[j] = bb[1]; N.loz[j] = bb[2];
            N.hix[j] = bb[3]; N.hiy[j] = bb[4]; N.hiz[j] = bb[5];
            if (c[j] >= n - 1) {
                N.child[j] = ~(c[j] - (n - 1));
            } else {
                const int at = atomicAdd(counter, 1);
                queue[at] = c[j];
                N.child[j] = at;
            }
        } else {
            N.lox[j] = N.loy[j] = N.loz[j] = 3.0e38f;
            N.hix[j] = N.hiy[j] = N.hiz[j] = -3.0e38f;
            N.child[j] = CGRT_NO_CHILD;
        }
        N.pad[j] = 0;
    }
    nodes4[w] = N;
}

// =================================================================================================================
// Height-field triangulation, objects.h:485-499 (step = 3 texels). `height` is the fp64 table of texture.h:27-37
// (computed on the host with libm's exp so that it is bit-identical to the reference's).
// =================================================================================================================
__global__ void bump_triangles_kernel(const double *__restrict__ height, int H, int W, double tex_px, double tex_pz, double lenx, double leny,
                                      double plane_y, double *__restrict__ tri9) {
    const int step = 3;
    int nj = W / step - 1, ni = H / step - 1;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ni * nj) return;
    int i = idx / nj, j = idx % nj;
    double x1 = tex_px + lenx * j * step / W;
    double x2 = tex_px + lenx * (j + 1) * step / W;
    double y1 = tex_pz + leny * i * step / H;
    double y2 = tex_pz + leny * (i + 1) * step / H;
    double ha = height[(size_t)(i * step) * W + j * step] + plane_y;
    double hb = height[(size_t)(i * step) * W + (j + 1) * step] + plane_y;
    double hc = height[(size_t)((i + 1) * step) * W + j * step] + plane_y;
    double hd = height[(size_t)((i + 1) * step) * W + (j + 1) * step] + plane_y;
    double *o = tri9 + (size_t)idx * 18;
    // Triangle(a,b,c)
    o[0] = x1; o[1] = ha; o[2] = y1;  o[3] = x2; o[4] = hb; o[5] = y1;  o[6] = x1; o[7] = hc; o[8] = y2;
    // Triangle(d,b,c)
    o[9] = x2; o[10] = hd; o[11] = y2;  o[12] = x2; o[13] = hb; o[14] = y1;  o[15] = x1; o[16] = hc; o[17] = y2;
}

// Texture staging (K5): packed RGB8 -> uchar4 rows so one texel is one aligned 4-byte load.
__global__ void texture_stage_kernel(const uint8_t *__restrict__ rgb, int64_t ntexel, uchar4 *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntexel) return;
    out[i] = make_uchar4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 255);
}

}  // namespace cgrt

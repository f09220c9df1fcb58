// faces.cu -- surface skeleton and face extraction on the device, and the whole-path entry.
//
// Replaces extract_skeleton (subpoly.py:556-581), extract_faces (subpoly.py:584-652) with
// regions_to_vertices / r_idx_as_tensor (subpoly.py:281-370), mean_points_with_valid
// (subpoly.py:669-678), sort_polygon_vertices_batch (geometry.py:483-525),
// tensor_to_triangle_faces (subpoly.py:700-728) and the driver subpoly() (subpoly.py:23-86).
//
// The reference groups vertices by region with unique(dim=0) + argsort over an expanded
// [sum 2^k, 36] int64 matrix.  Here every surface vertex owns a warp; lane q enumerates
// the vertex's q-th adjacent region, collects the region's vertices from the cell buckets
// and the vertex that is first in the row ("leader") emits it.  Rows of one leader are
// ranked inside the warp, so the global row order (torch.unique(dim=0) = lexicographic, first
// element = leader) falls out of one ordered scan over the vertices -- no global sort.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "complex.cuh"
#include "cells.cuh"
#include "net_eval.cuh"
#include "scan.cuh"
#include "sort.cuh"

struct tnb_mesh {
    int64_t V = 0, E = 0, P = 0, W = 0, T = 0;
    tnb::DevBuf<float> vert;   // [V][3]
    tnb::DevBuf<float> out;    // [V][R]
    tnb::DevBuf<int2> edges;   // [E]
    tnb::DevBuf<int> poly;     // angle-sorted face rows, one after the other (row p = poly[pstart[p] .. + pcnt[p])); padded to
                               // [P][W] with -1 only when somebody reads the polygons (W = 250 / 1170 for the large sphere / torus)
    tnb::DevBuf<int> pstart;   // [P] first element of every row
    tnb::DevBuf<int> pcnt;     // [P]
    int64_t n_elems = 0;
    tnb::DevBuf<int> tri;      // [T][3]
    tnb::DevBuf<unsigned char> tag;  // [V] slab sharding: bit0 / bit1 = on the plane shared with the lower / upper neighbour
    tnb::DevBuf<int> vidx;     // [V] number of the vertex in the complex (extract_skeleton's v_idx, subpoly.py:575)
    // between tnb_extract_mesh_begin and _finish (the slab exchange of vertex liveness sits in between)
    tnb::DevBuf<int> surf, used, counters;
    tnb::DevBuf<int2> tmp_edges;
    float eps = 0.0f;
    bool begun = false;
    int64_t near_plane = 0;  // slab sharding: vertices within eps of a shared plane that only one slab holds
};

namespace tnb {

constexpr int kThreads = 128;
constexpr int kMaxRow = 8192;      // hard limit of vertices per face row (sizes the HBM row scratch)
constexpr int kSmemRowStride = 16;  // rows up to this length are built in shared memory (16 KB per CTA: a dozen CTAs per SM)
constexpr int kMaxZeros = 5;  // 2^5 regions = one per lane
constexpr int kSortLocal = 32;  // face rows up to this length are angle-sorted in registers / local memory
enum { F_SURF = 0, F_EDGES, F_VERTS, F_ROWS, F_WIDTH, F_ERR_ZEROS, F_ERR_ROW, F_ERR_ORIGIN, F_MAXCNT, F_NEAR, F_LONG_TOTAL, F_LONG_CURSOR, F_RECS, F_NLONG, F_NHUGE, F_ERR_CELL, F_NLONGROWS, F_ELEMS, F_NWIDE, F_NUM = 24 };

// ---- surface skeleton -----------------------------------------------------------------------------
__global__ void k_surface_flags(const __grid_constant__ NetMeta n, const float *__restrict__ vert,
                                const float *__restrict__ out, const int *__restrict__ alive, int64_t V, float eps,
                                int *__restrict__ surf, int *__restrict__ counters)
{
    pdl_wait();
    int local = 0;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        float x[3] = {vert[3 * v], vert[3 * v + 1], vert[3 * v + 2]}, xp[3];
        preprocess(n, x, xp);
        // rows of pruned vertices are still in place (complex.cuh): they are not part of the complex
        bool on = alive[v] && fabsf(out[v * n.R + n.R - 1]) < eps;
        for (int d = 0; d < 3; ++d)
            if (xp[d] > 1.0f || xp[d] < 0.0f) on = false;
        surf[v] = on ? 1 : 0;
        local += on ? 1 : 0;
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(counters + F_SURF, local);
}

// Slab sharding is exact as long as every vertex the reference treats as lying ON a shared plane
// (|x - mark| <= eps, tropical.py:227-236) really was created in that plane, because only those
// exist on both slabs.  A vertex strictly inside one slab but within eps of the plane is seen by
// the reference from the cells on both sides; the neighbour slab does not have it.  They are
// counted here so that a run can say whether it was exact.
__global__ void k_count_near_plane(const uint64_t *__restrict__ sig, const unsigned char *__restrict__ tag,
                                   const int *__restrict__ alive, int64_t V, int plane_lo, int plane_hi,
                                   int *__restrict__ counters)
{
    pdl_wait();
    int local = 0;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        if (!alive[v]) continue;
        const uint64_t g = sig[3 * v + 2];
        if (grid_mask(g, 0)) continue;
        const int off = grid_off(g, 0), t = tag[v];
        if ((off == plane_lo && !(t & 1)) || (off == plane_hi && !(t & 2))) ++local;
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(counters + F_NEAR, local);
}

struct SurfEdgeCount {
    const int2 *edges;
    const int *surf;
    __device__ __forceinline__ int operator()(int64_t e) const { return (surf[edges[e].x] && surf[edges[e].y]) ? 1 : 0; }
};
struct SurfEdgeEmit {
    const int2 *edges;
    int2 *dst;
    int *used;
    __device__ __forceinline__ void operator()(int64_t e, int pos, int) const
    {
        const int2 ed = edges[e];
        dst[pos] = ed;
        used[ed.x] = 1;
        used[ed.y] = 1;
    }
};
struct FlagCount {
    const int *flag;
    __device__ __forceinline__ int operator()(int64_t i) const { return flag[i] ? 1 : 0; }
};
struct SurfVertEmit {  // new number of every surface vertex; its rows follow in k_gather_rows
    int *remap;
    const unsigned char *tag;
    unsigned char *ntag;
    int *vidx;
    __device__ __forceinline__ void operator()(int64_t v, int pos, int) const
    {
        remap[v] = pos;
        vidx[pos] = (int)v;
        ntag[pos] = tag[v];
    }
};
// position and output row of every surface vertex, flattened over (vertex, column): a warp reads and writes
// consecutive floats of the 33-float rows (one thread per row made every store instruction touch 32 sectors)
__global__ void __launch_bounds__(256) k_gather_rows(int64_t Vs, int R, const int *__restrict__ vidx, const float *__restrict__ vert,
                                                     const float *__restrict__ out, float *__restrict__ nvert, float *__restrict__ nout)
{
    pdl_wait();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = Vs * R;
    for (int64_t i = t0; i < total; i += stride) {
        const int64_t pos = i / R;
        nout[i] = out[(int64_t)vidx[pos] * R + (i - pos * R)];
    }
    for (int64_t i = t0; i < Vs * 3; i += stride) {
        const int64_t pos = i / 3;
        nvert[i] = vert[(int64_t)vidx[pos] * 3 + (i - pos * 3)];
    }
}
__global__ void k_remap_edges2(int2 *__restrict__ edges, int64_t E, const int *__restrict__ remap)
{
    pdl_wait();
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        int2 ed = edges[e];
        edges[e] = make_int2(remap[ed.x], remap[ed.y]);
    }
}

// ---- region rows ------------------------------------------------------------------------------------
__global__ void k_surface_sig(const __grid_constant__ NetMeta n, const float *__restrict__ vert,
                              const float *__restrict__ out, int64_t V, float eps, uint64_t *__restrict__ sig)
{
    pdl_wait();
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        float x[3] = {vert[3 * v], vert[3 * v + 1], vert[3 * v + 2]}, xp[3];
        preprocess(n, x, xp);
        uint64_t pos, neg;
        pack_signs(out + v * n.R, n.R, eps, pos, neg);
        sig[3 * v] = pos;
        sig[3 * v + 1] = neg;
        sig[3 * v + 2] = pack_grid(n, n.marks, xp, eps);
    }
}

struct Box {
    int lo[3], hi[3];
};
__device__ __forceinline__ Box box_of(uint64_t g)
{
    Box b;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        b.hi[d] = grid_off(g, d);
        b.lo[d] = b.hi[d] - (grid_mask(g, d) ? 0 : 1);
    }
    return b;
}
__device__ __forceinline__ int64_t cell_of(int cx, int cy, int cz, int dim)
{
    return ((int64_t)(cx + 2) * dim + (cy + 2)) * dim + (cz + 2);
}
__device__ __forceinline__ int zero_count(uint64_t pos, uint64_t neg, uint64_t g, uint64_t colmask)
{
    return __popcll(~(pos | neg) & colmask) + (3 - grid_mask(g, 0) - grid_mask(g, 1) - grid_mask(g, 2));
}

// ---- cell segments sorted by (zero count, vertex) ------------------------------------------------------
// A face row is the list of a region's vertices ordered by (zero count, vertex number) and is emitted by its
// first member.  With the records of every cell in THAT order (the key does not depend on the region), a lane
// that is not the leader of its row knows after the first compatible record -- before, a member of one of the
// reference's coincident-vertex clusters (hundreds of vertices in one cell) streamed the whole cluster to
// learn that some other member leads -- and the leader's row arrives sorted.  One warp per touched cell:
// up to 32 records in registers (bitonic over the lanes), up to kCellSortSmem in shared memory, longer
// segments by a CTA each (k_cell_sort_huge).  Records are rebuilt from the vertex number (cells.cuh files the
// vertex's packed signs with it), so only 64-bit keys are sorted.
constexpr int kCellSortSmem = 1024;
constexpr int kCellSortHuge = 8192;
__device__ __forceinline__ tnb_bucket_rec rec_of_vertex(int v, const uint64_t *__restrict__ sig)
{
    tnb_bucket_rec r;
    r.next = v;
    r.v = v;
    r.pos = sig[3 * (int64_t)v];
    r.neg = sig[3 * (int64_t)v + 1];
    r.grd = sig[3 * (int64_t)v + 2];
    return r;
}
__global__ void __launch_bounds__(kThreads) k_cell_sort(int64_t n_slots, const int2 *__restrict__ slots, const int2 *__restrict__ cells,
                                                        const tnb_bucket_rec *__restrict__ recs, tnb_bucket_rec *__restrict__ sorted,
                                                        const uint64_t *__restrict__ sig, uint64_t colmask, int *__restrict__ huge_list,
                                                        int *__restrict__ counters)
{
    pdl_wait();
    __shared__ unsigned long long s_keys[kThreads / 32][kCellSortSmem];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long *keys = s_keys[warp];
    for (int64_t t0 = ((int64_t)blockIdx.x * (kThreads / 32) + warp) * 32; t0 < n_slots; t0 += (int64_t)gridDim.x * kThreads) {  // warp uniform
        const int64_t t = t0 + lane;
        int2 sl = make_int2(-1, 0);
        if (t < n_slots) sl = slots[t];
        // the li == 0 entry stands for its cell.  Nearly all cells hold one to four vertices: such a cell is its own
        // lane's job (four independent loads, a sorting network in registers); the warp works together only through
        // the larger ones
        const bool rep = sl.x >= 0 && sl.y == 0;
        int2 myseg = make_int2(0, 0);
        if (rep) myseg = cells[sl.x];
        if (rep && myseg.x <= 4) {
            tnb_bucket_rec r[4] = {};
            unsigned long long key[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < myseg.x) r[i] = recs[myseg.y + i];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                key[i] = i < myseg.x ? (((unsigned long long)zero_count(r[i].pos, r[i].neg, r[i].grd, colmask) << 32) | (unsigned)r[i].v) : ~0ull;
#define TNB_CSWAP(a_, b_) do { if (key[a_] > key[b_]) { const unsigned long long tk = key[a_]; key[a_] = key[b_]; key[b_] = tk; \
                                                         const tnb_bucket_rec tr = r[a_]; r[a_] = r[b_]; r[b_] = tr; } } while (0)
            TNB_CSWAP(0, 1); TNB_CSWAP(2, 3); TNB_CSWAP(0, 2); TNB_CSWAP(1, 3); TNB_CSWAP(1, 2);
#undef TNB_CSWAP
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < myseg.x) sorted[myseg.y + i] = r[i];
        }
        unsigned todo = __ballot_sync(0xffffffffu, rep && myseg.x > 4);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int cell = __shfl_sync(0xffffffffu, sl.x, src);
            const int2 seg = make_int2(__shfl_sync(0xffffffffu, myseg.x, src), __shfl_sync(0xffffffffu, myseg.y, src));  // {records, first record}
            if (seg.x <= 32) {
                tnb_bucket_rec r = {0, 0, 0ull, 0ull, 0ull};
                unsigned long long key = ~0ull;
                if (lane < seg.x) {
                    r = recs[seg.y + lane];
                    key = ((unsigned long long)zero_count(r.pos, r.neg, r.grd, colmask) << 32) | (unsigned)r.v;
                }
                int from = lane;
                if (seg.x > 1) {  // bitonic network over the lanes; the source lane travels with the key
#pragma unroll
                    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
                        for (int j = k >> 1; j > 0; j >>= 1) {
                            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, j);
                            const int of = __shfl_xor_sync(0xffffffffu, from, j);
                            const bool up = (lane & k) == 0 || k == 32, lower = (lane & j) == 0;
                            const bool take_min = lower == up;
                            if (take_min ? ok < key : ok > key) { key = ok; from = of; }
                        }
                }
                // lane p holds the p-th key and where its record sits
                tnb_bucket_rec o;
                o.next = __shfl_sync(0xffffffffu, r.next, from);
                o.v = __shfl_sync(0xffffffffu, r.v, from);
                o.pos = __shfl_sync(0xffffffffu, r.pos, from);
                o.neg = __shfl_sync(0xffffffffu, r.neg, from);
                o.grd = __shfl_sync(0xffffffffu, r.grd, from);
                if (lane < seg.x) sorted[seg.y + lane] = o;
            } else if (seg.x <= kCellSortSmem) {
                int n = 64;
                while (n < seg.x) n <<= 1;
                for (int i = lane; i < n; i += 32) {
                    unsigned long long key = ~0ull;
                    if (i < seg.x) {
                        const tnb_bucket_rec r = recs[seg.y + i];
                        key = ((unsigned long long)zero_count(r.pos, r.neg, r.grd, colmask) << 32) | (unsigned)r.v;
                    }
                    keys[i] = key;
                }
                __syncwarp();
                for (int k = 2; k <= n; k <<= 1)
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        for (int i = lane; i < n; i += 32) {
                            const int p = i ^ j;
                            if (p > i) {
                                const unsigned long long x = keys[i], y = keys[p];
                                const bool up = (i & k) == 0;
                                if ((x > y) == up) { keys[i] = y; keys[p] = x; }
                            }
                        }
                        __syncwarp();
                    }
                for (int i = lane; i < seg.x; i += 32) sorted[seg.y + i] = rec_of_vertex((int)(uint32_t)keys[i], sig);
                __syncwarp();
            } else if (lane == 0) {
                if (seg.x > kCellSortHuge) atomicOr(counters + F_ERR_CELL, 1);
                else huge_list[atomicAdd(counters + F_NHUGE, 1)] = cell;
            }
        }
    }
}
// segments of more than kCellSortSmem records: a CTA each, keys in dynamic shared memory (never seen for a sphere;
// the large torus has cells with > 1000 coincident vertices)
__global__ void __launch_bounds__(256) k_cell_sort_huge(const int *__restrict__ huge_list, const int *__restrict__ counters,
                                                        const int2 *__restrict__ cells, const tnb_bucket_rec *__restrict__ recs,
                                                        tnb_bucket_rec *__restrict__ sorted, const uint64_t *__restrict__ sig, uint64_t colmask)
{
    pdl_wait();
    extern __shared__ unsigned long long s_huge[];  // [kCellSortHuge]
    const int n_huge = counters[F_NHUGE];
    for (int h = blockIdx.x; h < n_huge; h += gridDim.x) {
        const int2 seg = cells[huge_list[h]];
        int n = kCellSortSmem;
        while (n < seg.x) n <<= 1;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            unsigned long long key = ~0ull;
            if (i < seg.x) {
                const tnb_bucket_rec r = recs[seg.y + i];
                key = ((unsigned long long)zero_count(r.pos, r.neg, r.grd, colmask) << 32) | (unsigned)r.v;
            }
            s_huge[i] = key;
        }
        __syncthreads();
        for (int k = 2; k <= n; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < n; i += blockDim.x) {
                    const int p = i ^ j;
                    if (p > i) {
                        const unsigned long long x = s_huge[i], y = s_huge[p];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) { s_huge[i] = y; s_huge[p] = x; }
                    }
                }
                __syncthreads();
            }
        for (int i = threadIdx.x; i < seg.x; i += blockDim.x) sorted[seg.y + i] = rec_of_vertex((int)(uint32_t)s_huge[i], sig);
        __syncthreads();
    }
}

// One warp per surface vertex a; lane q builds the row of a's q-th adjacent region.
// mode 0: rows_per_vertex[a] = bit mask of the regions (lanes) whose row a leads and keeps (distinct rows of
//         >= 3 vertices; their count is the popcount), max width.  mode 1 walks the buckets again only for
//         those lanes: a row has one leader, so nearly all (vertex, region) pairs have nothing to do there.
// mode 1: write those rows, lexicographically ranked, at row_off[a].
// Row storage: `stride` 64-bit keys ((zero count << 32) | vertex) per lane, in dynamic shared
// memory when `scratch` is null, else in HBM (one slice per thread).  A row longer than
// `stride` raises F_ERR_ROW and the host retries with an HBM scratch of the measured width.
__device__ __forceinline__ int row_compare(const unsigned long long *x, int nx, const unsigned long long *y, int ny)
{
    const int n = min(nx, ny);
    for (int j = 0; j < n; ++j) {
        const unsigned a = (unsigned)x[j], b = (unsigned)y[j];
        if (a != b) return a < b ? -1 : 1;
    }
    return nx == ny ? 0 : (nx < ny ? -1 : 1);  // torch pads with -1, which sorts first
}

// LPV = lanes per vertex.  A surface vertex of a fitted network lies on two other planes as a rule: four
// adjacent regions, i.e. four busy lanes of a warp's 32.  LPV = 8 packs four vertices into a warp (those with
// up to three zero columns: 8 regions); the rare vertices with four or five zero columns take a whole warp
// (LPV = 32) in a launch of their own, as do the vertices of the long pass.  Everything below is group local:
// ballots are masked to the group's lanes, skips are per-lane predicates (no warp-wide `continue`).
template <int LPV>
__global__ void __launch_bounds__(kThreads) k_region_rows(int64_t V, const uint64_t *__restrict__ sig,
                                                          const int2 *__restrict__ cells,
                                                          const tnb_bucket_rec *__restrict__ recs, int dim,
                                                          uint64_t colmask, int mode, int stride,
                                                          unsigned long long *__restrict__ scratch,
                                                          int *__restrict__ rows_per_vertex,
                                                          const int *__restrict__ row_off, int *__restrict__ rows,
                                                          int *__restrict__ row_cnt, int *__restrict__ counters,
                                                          int cell_lo, int cell_hi, const int *__restrict__ list, int n_list,
                                                          unsigned char *__restrict__ is_long, int *__restrict__ long_list,
                                                          int *__restrict__ elems_per_vertex, const int *__restrict__ elem_off,
                                                          int *__restrict__ row_start, int *__restrict__ wide_list, const int *__restrict__ n_list_dev)
{
    pdl_wait();
    // Two passes share this kernel.  The FAST pass (list == nullptr) takes every vertex, builds rows of up to
    // `stride` (= kSmemRowStride) keys per lane in shared memory -- small enough for a dozen CTAs per SM: the
    // kernel is a chain of dependent L2 round trips per vertex, so what it needs is warps in flight -- and files
    // a vertex one of whose rows is longer (the coincident-vertex clusters of the reference's chunk-overlap
    // duplicates: rows of hundreds of vertices) in long_list.  The LONG pass (list != nullptr) takes those
    // vertices only, rows in an HBM scratch of the longest row seen.
    extern __shared__ unsigned long long s_rows[];  // [kThreads][stride] unless scratch is used
    __shared__ int s_cnt[kThreads];
    constexpr int GPW = 32 / LPV;  // vertices (groups) per warp
    constexpr int kMaxK = LPV == 8 ? 3 : kMaxZeros;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & (LPV - 1), g0 = lane & ~(LPV - 1);
    const unsigned gmask = LPV == 32 ? 0xffffffffu : (((1u << LPV) - 1u) << g0);
    unsigned long long *base = scratch ? scratch + ((size_t)blockIdx.x * kThreads + warp * 32) * stride
                                       : s_rows + (size_t)warp * 32 * stride;
    unsigned long long *mine = base + (size_t)lane * stride;
    int *wcnt = s_cnt + warp * 32;
    // three kinds of launch: every vertex (list == nullptr; the packed one files the vertices with more than three zero
    // columns in wide_list), the wide list (LPV = 32, rows still in shared memory), the long list (rows in the HBM scratch)
    const bool long_pass = scratch != nullptr;
    const int64_t n_items = list ? (n_list_dev ? *n_list_dev : n_list) : V;
    const int64_t per_pass = (int64_t)gridDim.x * (kThreads / 32) * GPW;
    for (int64_t item0 = ((int64_t)blockIdx.x * (kThreads / 32) + warp) * GPW; item0 < n_items; item0 += per_pass) {  // warp uniform
        const int64_t item = item0 + lane / LPV;
        bool active = item < n_items;
        const int64_t a = active ? (list ? list[item] : item) : 0;
        if (active && !long_pass && mode == 1 && is_long[a]) active = false;  // the long pass writes this vertex's rows
        const uint64_t pa = sig[3 * a], na = sig[3 * a + 1], ga = sig[3 * a + 2];
        const uint64_t za = ~(pa | na) & colmask;
        const int gz = 3 - grid_mask(ga, 0) - grid_mask(ga, 1) - grid_mask(ga, 2);
        const int ka = __popcll(za) + gz;
        // which launch takes this vertex: the packed one those with up to 3 zero columns, the wide one the others
        if (LPV == 8 && active && ka > 3) {
            if (mode == 0 && sub == 0 && wide_list) wide_list[atomicAdd(counters + F_NWIDE, 1)] = (int)a;
            active = false;
        }
        int cnt = 0;
        bool lead = false;
        const unsigned todo = !active ? 0u : (mode == 1 ? (unsigned)rows_per_vertex[a] : 0xffffffffu);
        if (active && ka > kMaxZeros) {
            if (sub == 0) atomicOr(counters + F_ERR_ZEROS, 1);
        } else if (active && ka <= kMaxK && sub < (1 << ka) && ((todo >> sub) & 1u)) {
            // region q: bit t of q decides the side of a's t-th zero column (grid axes first)
            int cell[3];
            int t = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                cell[d] = grid_off(ga, d);
                if (!grid_mask(ga, d)) { if (!((sub >> t) & 1)) cell[d] -= 1; ++t; }
            }
            uint64_t pat = pa & colmask;
            for (uint64_t m = za; m; m &= m - 1, ++t)
                if ((sub >> t) & 1) pat |= m & (~m + 1);
            // slab sharding: regions of cells that belong to a neighbour slab are emitted there
            const bool mine_cell = cell[0] >= cell_lo && cell[0] <= cell_hi;
            // collect the region's vertices ordered by (zero count, vertex number): the row
            // order r_idx_as_tensor builds from regions_to_vertices' group-by-zero-count output
            const int2 seg = cells[cell_of(cell[0], cell[1], cell[2], dim)];  // {records, first record} of the region's cell
            bool led_by_other = false;
            if (mine_cell) {
                // the cell's segment is contiguous and sorted by key (k_cell_sort): four independent 32-byte loads in
                // flight per lane; the FIRST member of the row leads it, so a lane whose first compatible record is
                // not a itself is done, and the leader's row arrives in row order
                for (int i0 = 0; i0 < seg.x && !led_by_other; i0 += 4) {
                    tnb_bucket_rec r4[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (i0 + k < seg.x) r4[k] = recs[seg.y + i0 + k];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (i0 + k >= seg.x || led_by_other) continue;
                        const int b = r4[k].v;
                        const uint64_t pb = r4[k].pos, nb = r4[k].neg, gb = r4[k].grd;
                        if ((pb & ~pat & colmask) || (nb & pat)) continue;  // a nonzero sign disagrees
                        const Box bb = box_of(gb);
                        bool in = true;
#pragma unroll
                        for (int d = 0; d < 3; ++d) in = in && cell[d] >= bb.lo[d] && cell[d] <= bb.hi[d];
                        if (!in) continue;
                        if (cnt == 0 && b != (int)a) { led_by_other = true; continue; }  // that vertex emits this row, not a
                        if (cnt < stride) mine[cnt] = ((unsigned long long)zero_count(pb, nb, gb, colmask) << 32) | (unsigned)b;
                        ++cnt;
                    }
                }
            }
            if (led_by_other) cnt = 0;
            lead = cnt >= 3;  // every surviving row starts with a itself
        }
        {   // a row that does not fit: the whole vertex goes to the long pass (group uniform)
            const bool over = cnt > stride;
            if (over) atomicMax(counters + F_MAXCNT, cnt);
            if (__ballot_sync(0xffffffffu, over) & gmask) {
                if (sub == 0 && active) {
                    if (long_pass) atomicOr(counters + F_ERR_ROW, 1);  // the long pass is sized to the longest row: cannot happen
                    else if (mode == 0) {
                        is_long[a] = 1;
                        rows_per_vertex[a] = 0;
                        elems_per_vertex[a] = 0;
                        long_list[atomicAdd(counters + F_NLONG, 1)] = (int)a;
                    }
                }
                active = false;
                lead = false;
                cnt = 0;
            }
        }
        wcnt[lane] = cnt;
        __syncwarp();
        const unsigned lead_mask = __ballot_sync(0xffffffffu, lead) & gmask;
        // identical rows collapse onto the lowest lane (torch.unique(dim=0), subpoly.py:620)
        bool keep = lead;
        if (lead)
            for (unsigned mset = lead_mask & ((1u << lane) - 1u); mset && keep; mset &= mset - 1) {
                const int o = __ffs(mset) - 1;
                if (row_compare(base + (size_t)o * stride, wcnt[o], mine, cnt) == 0) keep = false;
            }
        const unsigned keep_mask = __ballot_sync(0xffffffffu, keep) & gmask;
        if (mode == 0) {
            int wmax = keep ? cnt : 0, wsum = wmax;
#pragma unroll
            for (int d = LPV / 2; d > 0; d >>= 1) {
                wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, d));
                wsum += __shfl_xor_sync(0xffffffffu, wsum, d);
            }
            if (sub == 0 && active) {
                rows_per_vertex[a] = (int)(keep_mask >> g0);
                elems_per_vertex[a] = wsum;   // the rows are stored one after the other: no padding to the longest row
            }
            if (sub == 0 && wmax) atomicMax(counters + F_WIDTH, wmax);
            if (keep && cnt > kSortLocal) atomicAdd(counters + F_LONG_TOTAL, cnt);  // sizes k_sort_rows' key scratch
        } else if (keep) {
            int rank = 0, before = 0;  // lexicographic rank among the rows this vertex leads, and their elements
            for (unsigned mset = keep_mask & ~(1u << lane); mset; mset &= mset - 1) {
                const int o = __ffs(mset) - 1;
                if (row_compare(base + (size_t)o * stride, wcnt[o], mine, cnt) < 0) { ++rank; before += wcnt[o]; }
            }
            const int64_t r = (int64_t)row_off[a] + rank;
            const int64_t first = (int64_t)elem_off[a] + before;
            for (int j = 0; j < cnt; ++j) rows[first + j] = (int)(uint32_t)mine[j];
            row_start[r] = (int)first;
            row_cnt[r] = cnt;
        }
        __syncwarp();
    }
}

struct ArrayCount {
    const int *v;
    __device__ __forceinline__ int operator()(int64_t i) const { return v[i]; }
};
struct PopcCount {
    const int *v;
    __device__ __forceinline__ int operator()(int64_t i) const { return __popc((unsigned)v[i]); }
};
struct OffsetEmit {
    int *off;
    __device__ __forceinline__ void operator()(int64_t i, int pos, int) const { off[i] = pos; }
};

// Angular sort of every face row (geometry.py:483-516) around the normal at the face
// centre (subpoly.py:627-642); rewrites the row in sorted order.  Rows up to kSortLocal
// vertices are sorted in registers/local memory; longer ones (duplicated chunk-boundary
// geometry) in place in HBM with their scores in `score_scratch` [P][W].
__device__ __forceinline__ float angle_score(const float a[3], const float ua[3], const float u[3], const float nrm[3])
{
    const float d0 = a[1] * u[2] - a[2] * u[1], d1 = a[2] * u[0] - a[0] * u[2], d2 = a[0] * u[1] - a[1] * u[0];
    const float un = fmaxf(__fsqrt_rn((u[0] * u[0] + u[1] * u[1]) + u[2] * u[2]), 1e-8f);
    const float c = (ua[0] * __fdiv_rn(u[0], un) + ua[1] * __fdiv_rn(u[1], un)) + ua[2] * __fdiv_rn(u[2], un);
    const float dn = (d0 * nrm[0] + d1 * nrm[1]) + d2 * nrm[2];
    return c * (dn >= 0.0f ? 1.0f : -1.0f) + (dn < 0.0f ? 2.0f : 0.0f);
}

template <class C>
__global__ void __launch_bounds__(kThreads) k_sort_rows(const __grid_constant__ NetMeta n, int64_t P, const int *__restrict__ row_start,
                                                        const float *__restrict__ vert, int *__restrict__ rows,
                                                        const int *__restrict__ row_cnt,
                                                        unsigned long long *__restrict__ key_scratch,
                                                        int *__restrict__ counters, int *__restrict__ long_rows)
{
    pdl_wait();
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        const int cnt = row_cnt[p];
        if (cnt > kSortLocal && long_rows) {  // a warp's job (k_sort_rows_long): one thread sorting hundreds of keys held the whole launch up
            long_rows[atomicAdd(counters + F_NLONGROWS, 1)] = (int)p;
            continue;
        }
        int *row = rows + row_start[p];
        float sx = 0.0f, sy = 0.0f, sz = 0.0f;
        bool origin = false;
        for (int j = 0; j < cnt; ++j) {
            const float *q = vert + 3 * (int64_t)row[j];
            sx = sx + q[0];
            sy = sy + q[1];
            sz = sz + q[2];
            const float n2 = (q[0] * q[0] + q[1] * q[1]) + q[2] * q[2];
            if (!(__fsqrt_rn(n2) > 0.0f)) origin = true;
        }
        if (origin) atomicOr(counters + F_ERR_ORIGIN, 1);  // geometry.py:496 would drop this vertex
        const float k = (float)cnt;
        float mean[3] = {__fdiv_rn(sx, k), __fdiv_rn(sy, k), __fdiv_rn(sz, k)};
        float nrm[3];
        sdf_grad<C>(n, mean, nrm, true);
        const float *q0 = vert + 3 * (int64_t)row[0];
        const float a[3] = {q0[0] - mean[0], q0[1] - mean[1], q0[2] - mean[2]};
        const float an = fmaxf(__fsqrt_rn((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]), 1e-8f);
        const float ua[3] = {__fdiv_rn(a[0], an), __fdiv_rn(a[1], an), __fdiv_rn(a[2], an)};
        if (cnt <= kSortLocal) {
            int id[kSortLocal];
            float sc[kSortLocal];
            for (int j = 0; j < cnt; ++j) {
                id[j] = row[j];
                const float *q = vert + 3 * (int64_t)id[j];
                const float u[3] = {q[0] - mean[0], q[1] - mean[1], q[2] - mean[2]};
                sc[j] = angle_score(a, ua, u, nrm);
            }
            for (int i = 1; i < cnt; ++i) {  // stable, descending
                const float ks = sc[i];
                const int ki = id[i];
                int j = i - 1;
                while (j >= 0 && sc[j] < ks) { sc[j + 1] = sc[j]; id[j + 1] = id[j]; --j; }
                sc[j + 1] = ks;
                id[j + 1] = ki;
            }
            for (int j = 0; j < cnt; ++j) row[j] = id[j];
        } else {
            // long row: 64-bit keys (descending score, then original position = stable) in a
            // scratch segment, heap sort, gather
            unsigned long long *keys = key_scratch + atomicAdd(counters + F_LONG_CURSOR, cnt);
            for (int j = 0; j < cnt; ++j) {
                const float *q = vert + 3 * (int64_t)row[j];
                const float u[3] = {q[0] - mean[0], q[1] - mean[1], q[2] - mean[2]};
                const uint32_t b = __float_as_uint(angle_score(a, ua, u, nrm));
                const uint32_t asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // orders like the float
                keys[j] = ((unsigned long long)(0xFFFFFFFFu - asc) << 32) | (unsigned)j;
            }
            thread_sort(keys, cnt);
            for (int j = 0; j < cnt; ++j) keys[j] = (unsigned)row[(uint32_t)keys[j]];
            for (int j = 0; j < cnt; ++j) row[j] = (int)(uint32_t)keys[j];
        }
    }
}

// The rows of more than kSortLocal vertices (the coincident-vertex clusters of the reference's chunk-overlap
// duplicates: hundreds of vertices), one WARP each: positions gathered into shared memory by all lanes, the
// centre summed left to right by one lane (the order the oracle defines), the normal at the centre, scores by
// all lanes, 64-bit keys (descending score, then original position: stable) sorted by a bitonic network in
// shared memory.  Same operations per element as k_sort_rows, so the order is bit-identical.
constexpr int kLongRowSmem = 1024;
constexpr int kLongRowWarps = 2;
template <class C>
__global__ void __launch_bounds__(kLongRowWarps * 32) k_sort_rows_long(const __grid_constant__ NetMeta n, const int *__restrict__ row_start,
                                                                       const float *__restrict__ vert,
                                                                       int *__restrict__ rows, const int *__restrict__ row_cnt,
                                                                       const int *__restrict__ long_rows,
                                                                       unsigned long long *__restrict__ key_scratch, int *__restrict__ counters)
{
    pdl_wait();
    __shared__ float s_pos[kLongRowWarps][3][kLongRowSmem];
    __shared__ unsigned long long s_keys[kLongRowWarps][kLongRowSmem];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_long = counters[F_NLONGROWS];
    float *px = s_pos[warp][0], *py = s_pos[warp][1], *pz = s_pos[warp][2];
    unsigned long long *keys = s_keys[warp];
    for (int li = blockIdx.x * kLongRowWarps + warp; li < n_long; li += gridDim.x * kLongRowWarps) {
        const int64_t p = long_rows[li];
        const int cnt = row_cnt[p];
        int *row = rows + row_start[p];
        if (cnt > kLongRowSmem) {  // longer than the shared buffers: one lane, keys in the HBM scratch (as k_sort_rows did)
            if (lane == 0) {
                float sx = 0.0f, sy = 0.0f, sz = 0.0f;
                bool origin = false;
                for (int j = 0; j < cnt; ++j) {
                    const float *q = vert + 3 * (int64_t)row[j];
                    sx = sx + q[0];
                    sy = sy + q[1];
                    sz = sz + q[2];
                    const float n2 = (q[0] * q[0] + q[1] * q[1]) + q[2] * q[2];
                    if (!(__fsqrt_rn(n2) > 0.0f)) origin = true;
                }
                if (origin) atomicOr(counters + F_ERR_ORIGIN, 1);
                const float k = (float)cnt;
                float mean[3] = {__fdiv_rn(sx, k), __fdiv_rn(sy, k), __fdiv_rn(sz, k)};
                float nrm[3];
                sdf_grad<C>(n, mean, nrm, true);
                const float *q0 = vert + 3 * (int64_t)row[0];
                const float a[3] = {q0[0] - mean[0], q0[1] - mean[1], q0[2] - mean[2]};
                const float an = fmaxf(__fsqrt_rn((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]), 1e-8f);
                const float ua[3] = {__fdiv_rn(a[0], an), __fdiv_rn(a[1], an), __fdiv_rn(a[2], an)};
                unsigned long long *ks = key_scratch + atomicAdd(counters + F_LONG_CURSOR, cnt);
                for (int j = 0; j < cnt; ++j) {
                    const float *q = vert + 3 * (int64_t)row[j];
                    const float u[3] = {q[0] - mean[0], q[1] - mean[1], q[2] - mean[2]};
                    const uint32_t b = __float_as_uint(angle_score(a, ua, u, nrm));
                    const uint32_t asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // orders like the float
                    ks[j] = ((unsigned long long)(0xFFFFFFFFu - asc) << 32) | (unsigned)j;
                }
                thread_sort(ks, cnt);
                for (int j = 0; j < cnt; ++j) ks[j] = (unsigned)row[(uint32_t)ks[j]];
                for (int j = 0; j < cnt; ++j) row[j] = (int)(uint32_t)ks[j];
            }
            __syncwarp();
            continue;
        }
        bool origin = false;
        for (int j = lane; j < cnt; j += 32) {
            const float *q = vert + 3 * (int64_t)row[j];
            const float x = q[0], y = q[1], z = q[2];
            px[j] = x; py[j] = y; pz[j] = z;
            const float n2 = (x * x + y * y) + z * z;
            if (!(__fsqrt_rn(n2) > 0.0f)) origin = true;
        }
        if (__any_sync(0xffffffffu, origin) && lane == 0) atomicOr(counters + F_ERR_ORIGIN, 1);  // geometry.py:496 would drop this vertex
        __syncwarp();
        float sx = 0.0f, sy = 0.0f, sz = 0.0f;
        if (lane == 0)
            for (int j = 0; j < cnt; ++j) { sx = sx + px[j]; sy = sy + py[j]; sz = sz + pz[j]; }
        sx = __shfl_sync(0xffffffffu, sx, 0);
        sy = __shfl_sync(0xffffffffu, sy, 0);
        sz = __shfl_sync(0xffffffffu, sz, 0);
        const float k = (float)cnt;
        float mean[3] = {__fdiv_rn(sx, k), __fdiv_rn(sy, k), __fdiv_rn(sz, k)};
        float nrm[3];
        sdf_grad<C>(n, mean, nrm, true);  // every lane the same point: the same value
        const float a[3] = {px[0] - mean[0], py[0] - mean[1], pz[0] - mean[2]};
        const float an = fmaxf(__fsqrt_rn((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]), 1e-8f);
        const float ua[3] = {__fdiv_rn(a[0], an), __fdiv_rn(a[1], an), __fdiv_rn(a[2], an)};
        int m = 64;
        while (m < cnt) m <<= 1;
        for (int j = lane; j < m; j += 32) {
            unsigned long long key = ~0ull;
            if (j < cnt) {
                const float u[3] = {px[j] - mean[0], py[j] - mean[1], pz[j] - mean[2]};
                const uint32_t b = __float_as_uint(angle_score(a, ua, u, nrm));
                const uint32_t asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
                key = ((unsigned long long)(0xFFFFFFFFu - asc) << 32) | (unsigned)j;
            }
            keys[j] = key;
        }
        __syncwarp();
        for (int kk = 2; kk <= m; kk <<= 1)
            for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                for (int i = lane; i < m; i += 32) {
                    const int q = i ^ jj;
                    if (q > i) {
                        const unsigned long long x = keys[i], y = keys[q];
                        const bool up = (i & kk) == 0;
                        if ((x > y) == up) { keys[i] = y; keys[q] = x; }
                    }
                }
                __syncwarp();
            }
        int *ids = reinterpret_cast<int *>(px);  // the positions are not needed any more
        for (int j = lane; j < cnt; j += 32) ids[j] = row[(uint32_t)keys[j]];
        __syncwarp();
        for (int j = lane; j < cnt; j += 32) row[j] = ids[j];
        __syncwarp();
    }
}

// ---- fan triangulation ---------------------------------------------------------------------------
// tensor_to_triangle_faces (subpoly.py:700-728) emits, for fan step i = 0,1,..., one triangle
// (first, i+1-th, i+2-th) per row that has at least i+3 vertices: ordered by step, then by row.
// Position of (row p, step i) = base[i] + #{rows before p with >= i+3 vertices}.  Three kernels
// over a fixed grid give all steps at once (the reference loops over steps with boolean
// masks):  per-block counts per step -> column scan -> warp-ballot ranks inside each block.
constexpr int kFanThreads = 256;

__device__ __forceinline__ void fan_slice(int64_t P, int64_t &begin, int64_t &end)
{
    int64_t per = (P + gridDim.x - 1) / gridDim.x;
    per = (per + kFanThreads - 1) / kFanThreads * kFanThreads;
    begin = min(per * blockIdx.x, P);
    end = min(begin + per, P);
}

// G[b][i] = rows of block b with at least i+3 vertices (i < W-2)
__global__ void __launch_bounds__(kFanThreads) k_fan_hist(const int *__restrict__ row_cnt, int64_t P, int W,
                                                          int *__restrict__ G)
{
    pdl_wait();
    extern __shared__ int s_hist[];  // [W+1]
    for (int c = threadIdx.x; c <= W; c += kFanThreads) s_hist[c] = 0;
    __syncthreads();
    int64_t begin, end;
    fan_slice(P, begin, end);
    for (int64_t p = begin + threadIdx.x; p < end; p += kFanThreads) atomicAdd(s_hist + row_cnt[p], 1);
    __syncthreads();
    if (threadIdx.x == 0) {  // suffix sums (W is small)
        int run = 0;
        for (int c = W; c >= 3; --c) {
            run += s_hist[c];
            G[(int64_t)blockIdx.x * W + (c - 3)] = run;
        }
    }
}

// off[b][i] = sum_{b' < b} G[b'][i] (one CTA per fan step i: a column of G), col_total[i] = the column's sum
__global__ void __launch_bounds__(kFanThreads) k_fan_cols(const int *__restrict__ G, int blocks, int W, int *__restrict__ off,
                                                          int *__restrict__ col_total)
{
    pdl_wait();
    __shared__ int s_w[kFanThreads / 32];
    const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (blocks + kFanThreads - 1) / kFanThreads, b0 = threadIdx.x * per, b1 = min(b0 + per, blocks);
    int mine = 0;
    for (int b = b0; b < b1; ++b) mine += G[(int64_t)b * W + i];
    const int incl = warp_inclusive_scan(mine);
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kFanThreads / 32; ++w) {
        const int t = s_w[w];
        if (w < warp) before += t;
        all += t;
    }
    int run = before + incl - mine;
    for (int b = b0; b < b1; ++b) {
        off[(int64_t)b * W + i] = run;
        run += G[(int64_t)b * W + i];
    }
    if (threadIdx.x == 0) col_total[i] = all;
}
// base[i] = triangles of the fan steps before i; total triangle count -> *total
__global__ void __launch_bounds__(1024) k_fan_base(const int *__restrict__ col_total, int n, int *__restrict__ base, int *__restrict__ total)
{
    pdl_wait();
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const int v = i < n ? col_total[i] : 0;
        const int incl = warp_inclusive_scan(v);
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int before = s_carry, all = 0;
        for (int w = 0; w < 32; ++w) {
            const int t = s_w[w];
            if (w < warp) before += t;
            all += t;
        }
        if (i < n) base[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += all;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(kFanThreads) k_fan_write(const int *__restrict__ rows, const int *__restrict__ row_start,
                                                           const int *__restrict__ row_cnt,
                                                           int64_t P, int W, const int *__restrict__ off, const int *__restrict__ base,
                                                           int *__restrict__ tri)
{
    pdl_wait();
    extern __shared__ int s_run[];  // [W] triangles already emitted by this block per step
    __shared__ int s_w[kFanThreads / 32];
    __shared__ int s_max;
    for (int c = threadIdx.x; c < W; c += kFanThreads) s_run[c] = 0;
    __syncthreads();
    int64_t begin, end;
    fan_slice(P, begin, end);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int *boff = off + (int64_t)blockIdx.x * W;
    for (int64_t tile = begin; tile < end; tile += kFanThreads) {
        const int64_t p = tile + threadIdx.x;
        const int c = p < end ? row_cnt[p] : 0;
        int mx = c;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        if (threadIdx.x == 0) s_max = 0;
        __syncthreads();
        if (lane == 0) atomicMax(&s_max, mx);
        __syncthreads();
        const int tile_max = s_max;
        const int *row = rows + (p < end ? row_start[p] : 0);
        for (int i = 0; i + 3 <= tile_max; ++i) {
            const bool pred = c >= i + 3;
            const unsigned ball = __ballot_sync(0xffffffffu, pred);
            if (lane == 0) s_w[warp] = __popc(ball);
            __syncthreads();
            int woff = 0, ttot = 0;
#pragma unroll
            for (int w = 0; w < kFanThreads / 32; ++w) {
                const int t = s_w[w];
                if (w < warp) woff += t;
                ttot += t;
            }
            if (pred) {
                const int64_t pos = (int64_t)base[i] + boff[i] + s_run[i] + woff + __popc(ball & ((1u << lane) - 1u));
                tri[3 * pos] = row[0];
                tri[3 * pos + 1] = row[i + 1];
                tri[3 * pos + 2] = row[i + 2];
            }
            __syncthreads();
            if (threadIdx.x == 0) s_run[i] += ttot;
        }
        __syncthreads();
    }
}

static int read_small(const int *d, int *h, int n, cudaStream_t s)
{
    TNB_CUDA(cudaMemcpyAsync(h, d, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    TNB_CUDA(cudaStreamSynchronize(s));
    return TNB_OK;
}

// extract_skeleton up to the liveness flags of the surface vertices (subpoly.py:556-572); a
// slab-sharded run publishes the flags of its shared-plane vertices here
static int extract_begin_impl(const tnb_net *net, tnb_complex *c, float eps, tnb_mesh *m, cudaStream_t s)
{
    const NetMeta &nm = net->meta;
    {
        int rcs = complex_sync_counts(c, s);
        if (rcs) return rcs;
    }
    const int64_t V = c->V, E = c->E;
    m->eps = eps;
    m->begun = true;
    int rc;
    TNB_CUDA(m->surf.reserve((size_t)std::max<int64_t>(V, 1)));
    TNB_CUDA(m->used.reserve((size_t)std::max<int64_t>(V, 1)));
    TNB_CUDA(m->counters.reserve(F_NUM));
    TNB_CUDA(cudaMemsetAsync(m->counters.p, 0, F_NUM * sizeof(int), s));
    TNB_CUDA(cudaMemsetAsync(m->used.p, 0, (size_t)std::max<int64_t>(V, 1) * sizeof(int), s));
    TNB_CUDA(m->tmp_edges.reserve((size_t)std::max<int64_t>(E, 1)));
    if (V > 0) {
        TNB_CUDA(launch_pdl(k_surface_flags, dim3(grid_for(V, 256)), dim3(256), 0, s, nm, c->cvert(), c->cout_(), c->calive(), V, eps, m->surf.p, m->counters.p));
        TNB_LAUNCH_CHECK();
        DevBuf<int> block_sums;
        TNB_CUDA(block_sums.reserve(kScanMaxBlocks));
        if ((rc = compact(E, SurfEdgeCount{c->cedges(), m->surf.p}, SurfEdgeEmit{c->cedges(), m->tmp_edges.p, m->used.p},
                          block_sums.p, m->counters.p + F_EDGES, s)))
            return rc;
    }
    if (c->halo.enabled && V > 0) {
        TNB_CUDA(launch_pdl(k_count_near_plane, dim3(grid_for(V, 256)), dim3(256), 0, s, c->csig(), c->tag[c->vcur].p, c->calive(), V, c->halo.tag_lower ? c->halo.x_lo : -7,
                                                            c->halo.tag_upper ? c->halo.x_hi : -7, m->counters.p));
        TNB_LAUNCH_CHECK();
    }
    if (c->halo.enabled && (rc = halo_publish_used(c, V, m->used.p, s))) return rc;
    return TNB_OK;
}

static int extract_finish_impl(const tnb_net *net, tnb_complex *c, tnb_mesh *m, cudaStream_t s)
{
    const NetMeta &nm = net->meta;
    const int R = nm.R;
    const float eps = m->eps;
    const int64_t V = c->V, E = c->E;
    (void)E;
    int h[F_NUM];
    int rc;
    const bool halo = c->halo.enabled;
    if (halo) {
        if ((rc = halo_merge_used(c, V, m->used.p, s))) return rc;
        c->counts_stale = true;  // re-read the sticky bits the exchange may have raised
        if ((rc = complex_sync_counts(c, s))) return rc;
    }
    if (V == 0) return TNB_OK;
    DevBuf<int> remap, block_sums;
    DevBuf<int> &used = m->used, &counters = m->counters;
    DevBuf<int2> &tmp_edges = m->tmp_edges;
    TNB_CUDA(remap.reserve((size_t)V));
    TNB_CUDA(block_sums.reserve(kScanMaxBlocks));
    if ((rc = read_small(counters.p, h, F_NUM, s))) return rc;
    m->near_plane = h[F_NEAR];
    if (h[F_SURF] < 3 && !halo) return TNB_OK;  // subpoly.py:568-569
    const int64_t Es = h[F_EDGES];
    // vertex compaction: count first to size the mesh (one bit per vertex slot: the write pass takes 32 slots per thread)
    {
        int64_t blocks = std::min<int64_t>((V + kScanThreads - 1) / kScanThreads, kScanMaxBlocks);
        DevBuf<uint32_t> vmask;
        TNB_CUDA(vmask.reserve((size_t)((V + 31) / 32 + kScanMaxBlocks + 64)));
        TNB_CUDA(launch_pdl(k_scan_count_mask<FlagCount>, dim3((unsigned)blocks), dim3(kScanThreads), 0, s, V, FlagCount{used.p}, block_sums.p, vmask.p, (const int *)nullptr));
        TNB_LAUNCH_CHECK();
        std::vector<int> hb(blocks);
        if ((rc = read_small(block_sums.p, hb.data(), (int)blocks, s))) return rc;
        int64_t Vs = 0;
        for (int v : hb) Vs += v;
        m->V = Vs;
        m->E = Es;
        TNB_CUDA(m->vert.reserve((size_t)std::max<int64_t>(Vs, 1) * 3));
        TNB_CUDA(m->out.reserve((size_t)std::max<int64_t>(Vs, 1) * R));
        TNB_CUDA(m->tag.reserve((size_t)std::max<int64_t>(Vs, 1)));
        TNB_CUDA(m->edges.reserve((size_t)std::max<int64_t>(Es, 1)));
        TNB_CUDA(m->vidx.reserve((size_t)std::max<int64_t>(Vs, 1)));
        SurfVertEmit ve{remap.p, c->tag[c->vcur].p, m->tag.p, m->vidx.p};
        TNB_CUDA(launch_pdl(k_scan_write_mask<SurfVertEmit>, dim3((unsigned)blocks), dim3(kScanThreads), 0, s, V, (const uint32_t *)vmask.p, ve, (const int *)block_sums.p, counters.p + F_VERTS, (const int *)nullptr));
        TNB_LAUNCH_CHECK();
        if (Vs > 0) {
            TNB_CUDA(launch_pdl(k_gather_rows, dim3(grid_for(Vs * R, 256)), dim3(256), 0, s, Vs, R, m->vidx.p, c->cvert(), c->cout_(), m->vert.p, m->out.p));
            TNB_LAUNCH_CHECK();
        }
    }
    if (Es > 0) {
        TNB_CUDA(cudaMemcpyAsync(m->edges.p, tmp_edges.p, (size_t)Es * sizeof(int2), cudaMemcpyDeviceToDevice, s));
        TNB_CUDA(launch_pdl(k_remap_edges2, dim3(grid_for(Es, 256)), dim3(256), 0, s, m->edges.p, Es, remap.p));
        TNB_LAUNCH_CHECK();
    }
    // cells of this slab along the first axis (all of them without slab sharding)
    const int cell_lo = c->halo.tag_lower ? c->halo.x_lo : -(1 << 30);
    const int cell_hi = c->halo.tag_upper ? c->halo.x_hi - 1 : (1 << 30);
    const int64_t Vs = m->V;
    if (Vs == 0) return TNB_OK;

    // ---- extract_faces ----
    DevBuf<uint64_t> sig;
    DevBuf<tnb_bucket_rec> next;
    DevBuf<int> rows_per_vertex, row_off, elems_per_vertex, elem_off;
    DevBuf<unsigned long long> head;   // read as int2 {records, first record} per cell (cells.cuh)
    DevBuf<int2> cslot;
    TNB_CUDA(sig.reserve((size_t)Vs * 3));
    TNB_CUDA(next.reserve((size_t)Vs * 8));
    TNB_CUDA(cslot.reserve((size_t)Vs * 8));
    TNB_CUDA(rows_per_vertex.reserve((size_t)Vs));
    TNB_CUDA(row_off.reserve((size_t)Vs));
    TNB_CUDA(elems_per_vertex.reserve((size_t)Vs));
    TNB_CUDA(elem_off.reserve((size_t)Vs));
    const int dim = nm.n_marks + 2;
    const int64_t n_cells = (int64_t)dim * dim * dim;
    TNB_CUDA(head.reserve((size_t)n_cells));
    TNB_CUDA(cudaMemsetAsync(head.p, 0, (size_t)n_cells * sizeof(unsigned long long), s));
    const uint64_t colmask = (1ull << (R - 1)) - 1ull;  // m_rgn[:, :-1], subpoly.py:611
    TNB_CUDA(launch_pdl(k_surface_sig, dim3(grid_for(Vs, kThreads)), dim3(kThreads), 0, s, nm, m->vert.p, m->out.p, Vs, eps, sig.p));
    TNB_LAUNCH_CHECK();
    TNB_CUDA(cudaMemsetAsync(counters.p + F_RECS, 0, sizeof(int), s));
    if ((rc = cells_build(8, nullptr, nullptr, Vs, sig.p, (int2 *)head.p, cslot.p, next.p, counters.p + F_RECS, dim, s))) return rc;
    DevBuf<tnb_bucket_rec> sorted;    // the same segments, every one ordered by (zero count, vertex)
    DevBuf<int> huge_list;
    TNB_CUDA(sorted.reserve((size_t)Vs * 8));
    TNB_CUDA(huge_list.reserve((size_t)std::max<int64_t>(Vs * 8 / kCellSortSmem, 1) + 1));
    TNB_CUDA(launch_pdl(k_cell_sort, dim3(grid_for(Vs * 8, kThreads)), dim3(kThreads), 0, s, Vs * 8, cslot.p, (const int2 *)head.p, next.p, sorted.p, sig.p, colmask, huge_list.p, counters.p));
    TNB_LAUNCH_CHECK();
    {
        static bool attr_set = false;
        if (!attr_set) {
            TNB_CUDA(cudaFuncSetAttribute(k_cell_sort_huge, cudaFuncAttributeMaxDynamicSharedMemorySize, kCellSortHuge * (int)sizeof(unsigned long long)));
            attr_set = true;
        }
        TNB_CUDA(launch_pdl(k_cell_sort_huge, dim3(kSMs), dim3(256), kCellSortHuge * sizeof(unsigned long long), s, huge_list.p, counters.p, (const int2 *)head.p, next.p, sorted.p, sig.p, colmask));
        TNB_LAUNCH_CHECK();
    }
    const unsigned gw8 = grid_for(Vs, kThreads / 8);
    const size_t rows_smem = (size_t)kThreads * kSmemRowStride * sizeof(unsigned long long);
    DevBuf<unsigned long long> scratch;
    DevBuf<unsigned char> is_long;
    DevBuf<int> long_list, wide_list;
    TNB_CUDA(is_long.reserve((size_t)Vs));
    TNB_CUDA(long_list.reserve((size_t)Vs));
    TNB_CUDA(wide_list.reserve((size_t)Vs));
    static const int wide_ctas = std::getenv("TNB_WIDE_ROW_CTAS") ? std::atoi(std::getenv("TNB_WIDE_ROW_CTAS")) : 4;  // CTAs per SM (A/B)
    const unsigned gwide = kSMs * wide_ctas;  // the wide list is short (its length stays on the device): a fixed grid strides over it
    TNB_CUDA(cudaMemsetAsync(is_long.p, 0, (size_t)Vs, s));
    // fast pass: every vertex, rows in shared memory
    prof_begin(TNB_PROF_FACE_ROWS, s);
    // four vertices per warp for those with up to three zero columns (nearly all), a warp each for the others
    TNB_CUDA(launch_pdl(k_region_rows<8>, dim3(gw8), dim3(kThreads), rows_smem, s, Vs, sig.p, (const int2 *)head.p, sorted.p, dim, colmask, 0, kSmemRowStride, nullptr,
                                                      rows_per_vertex.p, nullptr, nullptr, nullptr, counters.p, cell_lo, cell_hi,
                                                      nullptr, 0, is_long.p, long_list.p, elems_per_vertex.p, nullptr, nullptr, wide_list.p, nullptr));
    TNB_LAUNCH_CHECK();
    TNB_CUDA(launch_pdl(k_region_rows<32>, dim3(gwide), dim3(kThreads), rows_smem, s, Vs, sig.p, (const int2 *)head.p, sorted.p, dim, colmask, 0, kSmemRowStride, nullptr,
                                                          rows_per_vertex.p, nullptr, nullptr, nullptr, counters.p, cell_lo, cell_hi,
                                                          wide_list.p, 0, is_long.p, long_list.p, elems_per_vertex.p, nullptr, nullptr, nullptr,
                                                          counters.p + F_NWIDE));
    TNB_LAUNCH_CHECK();
    prof_end(TNB_PROF_FACE_ROWS, s, Vs, Vs * 28);
    if ((rc = read_small(counters.p, h, F_NUM, s))) return rc;
    if (h[F_ERR_ZEROS]) { set_error("a surface vertex lies on more than 5 planes: more than 32 adjacent regions"); return TNB_ERR_UNSUPPORTED; }
    if (h[F_ERR_CELL]) { set_error("a marks-grid cell holds more than " + std::to_string(kCellSortHuge) + " surface vertices"); return TNB_ERR_UNSUPPORTED; }
    const int n_long = h[F_NLONG];
    int stride = 0;
    unsigned gl = 0;
    if (n_long > 0) {  // long pass: the vertices with a row that did not fit, rows in HBM, sized to the longest row seen
        if (h[F_MAXCNT] > kMaxRow) {
            set_error("a face has more than " + std::to_string(kMaxRow) + " vertices (" + std::to_string(h[F_MAXCNT]) + ")");
            return TNB_ERR_UNSUPPORTED;
        }
        stride = (h[F_MAXCNT] + 7) / 8 * 8;
        gl = std::min<unsigned>(grid_for(n_long, kThreads / 32), kSMs * 4);
        TNB_CUDA(scratch.reserve((size_t)gl * kThreads * stride));
        prof_begin(TNB_PROF_FACE_ROWS, s);
        TNB_CUDA(launch_pdl(k_region_rows<32>, dim3(gl), dim3(kThreads), 0, s, Vs, sig.p, (const int2 *)head.p, sorted.p, dim, colmask, 0, stride, scratch.p, rows_per_vertex.p,
                                                  nullptr, nullptr, nullptr, counters.p, cell_lo, cell_hi, long_list.p, n_long, is_long.p, nullptr,
                                                  elems_per_vertex.p, nullptr, nullptr, nullptr, nullptr));
        TNB_LAUNCH_CHECK();
        prof_end(TNB_PROF_FACE_ROWS, s, n_long, (int64_t)n_long * 28);
    }
    if ((rc = compact(Vs, PopcCount{rows_per_vertex.p}, OffsetEmit{row_off.p}, block_sums.p, counters.p + F_ROWS, s))) return rc;
    if ((rc = compact(Vs, ArrayCount{elems_per_vertex.p}, OffsetEmit{elem_off.p}, block_sums.p, counters.p + F_ELEMS, s))) return rc;
    if ((rc = read_small(counters.p, h, F_NUM, s))) return rc;
    if (h[F_ERR_ROW]) { set_error("face rows: a row outgrew the scratch sized for it"); return TNB_ERR_CAPACITY; }
    const int64_t P = h[F_ROWS];
    const int W = h[F_WIDTH];
    m->P = P;
    m->W = W;
    if (P == 0) return TNB_OK;
    m->n_elems = h[F_ELEMS];
    TNB_CUDA(m->poly.reserve((size_t)std::max<int64_t>(m->n_elems, 1)));
    TNB_CUDA(m->pstart.reserve((size_t)P));
    TNB_CUDA(m->pcnt.reserve((size_t)P));
    prof_begin(TNB_PROF_FACE_ROWS, s);
    TNB_CUDA(launch_pdl(k_region_rows<8>, dim3(gw8), dim3(kThreads), rows_smem, s, Vs, sig.p, (const int2 *)head.p, sorted.p, dim, colmask, 1, kSmemRowStride, nullptr,
                                                      rows_per_vertex.p, row_off.p, m->poly.p, m->pcnt.p, counters.p, cell_lo, cell_hi,
                                                      nullptr, 0, is_long.p, nullptr, nullptr, elem_off.p, m->pstart.p, nullptr, nullptr));
    TNB_LAUNCH_CHECK();
    TNB_CUDA(launch_pdl(k_region_rows<32>, dim3(gwide), dim3(kThreads), rows_smem, s, Vs, sig.p, (const int2 *)head.p, sorted.p, dim, colmask, 1, kSmemRowStride, nullptr,
                                                          rows_per_vertex.p, row_off.p, m->poly.p, m->pcnt.p, counters.p, cell_lo, cell_hi,
                                                          wide_list.p, 0, is_long.p, nullptr, nullptr, elem_off.p, m->pstart.p, nullptr,
                                                          counters.p + F_NWIDE));
    TNB_LAUNCH_CHECK();
    if (n_long > 0) {
        TNB_CUDA(launch_pdl(k_region_rows<32>, dim3(gl), dim3(kThreads), 0, s, Vs, sig.p, (const int2 *)head.p, sorted.p, dim, colmask, 1, stride, scratch.p, rows_per_vertex.p,
                                                  row_off.p, m->poly.p, m->pcnt.p, counters.p, cell_lo, cell_hi, long_list.p, n_long, is_long.p, nullptr,
                                                  nullptr, elem_off.p, m->pstart.p, nullptr, nullptr));
        TNB_LAUNCH_CHECK();
    }
    prof_end(TNB_PROF_FACE_ROWS, s, 0, 0);
    {
        unsigned g = grid_for(P, kThreads);
        DevBuf<unsigned long long> score_scratch;  // one segment per long row (their total was counted with the rows)
        if (W > kSortLocal) TNB_CUDA(score_scratch.reserve((size_t)std::max(h[F_LONG_TOTAL], 1)));
        DevBuf<int> long_rows;
        TNB_CUDA(long_rows.reserve((size_t)std::max(h[F_LONG_TOTAL] / (kSortLocal + 1) + 1, 1)));  // every long row counted more than kSortLocal keys
        if (net->fixed_cfg) TNB_CUDA(launch_pdl(k_sort_rows<CfgRef>, dim3(g), dim3(kThreads), 0, s, nm, P, m->pstart.p, m->vert.p, m->poly.p, m->pcnt.p, score_scratch.p, counters.p, long_rows.p));
        else TNB_CUDA(launch_pdl(k_sort_rows<CfgAny>, dim3(g), dim3(kThreads), 0, s, nm, P, m->pstart.p, m->vert.p, m->poly.p, m->pcnt.p, score_scratch.p, counters.p, long_rows.p));
        TNB_LAUNCH_CHECK();
        if (W > kSortLocal) {
            const unsigned gl2 = (unsigned)std::min<int64_t>(h[F_LONG_TOTAL] / (kSortLocal + 1) / kLongRowWarps + 1, kSMs * 8);
            if (net->fixed_cfg) TNB_CUDA(launch_pdl(k_sort_rows_long<CfgRef>, dim3(gl2), dim3(kLongRowWarps * 32), 0, s, nm, m->pstart.p, m->vert.p, m->poly.p, m->pcnt.p, long_rows.p, score_scratch.p, counters.p));
            else TNB_CUDA(launch_pdl(k_sort_rows_long<CfgAny>, dim3(gl2), dim3(kLongRowWarps * 32), 0, s, nm, m->pstart.p, m->vert.p, m->poly.p, m->pcnt.p, long_rows.p, score_scratch.p, counters.p));
            TNB_LAUNCH_CHECK();
        }
    }
    // fan triangles, ordered by fan step then by row
    if ((rc = read_small(counters.p, h, F_NUM, s))) return rc;
    if (h[F_ERR_ORIGIN]) {
        set_error("a face vertex sits exactly at the origin (geometry.py:496 treats it as padding): unsupported");
        return TNB_ERR_UNSUPPORTED;
    }
    if (W >= 3) {
        const int fblocks = (int)std::min<int64_t>((P + kFanThreads - 1) / kFanThreads, kSMs * 4);
        DevBuf<int> G, off, col_total, base;
        TNB_CUDA(G.reserve((size_t)fblocks * W));
        TNB_CUDA(off.reserve((size_t)fblocks * W));
        TNB_CUDA(col_total.reserve((size_t)W));
        TNB_CUDA(base.reserve((size_t)W));
        TNB_CUDA(launch_pdl(k_fan_hist, dim3(fblocks), dim3(kFanThreads), (W + 1) * sizeof(int), s, m->pcnt.p, P, W, G.p));
        TNB_LAUNCH_CHECK();
        TNB_CUDA(launch_pdl(k_fan_cols, dim3(W - 2), dim3(kFanThreads), 0, s, G.p, fblocks, W, off.p, col_total.p));
        TNB_LAUNCH_CHECK();
        TNB_CUDA(launch_pdl(k_fan_base, dim3(1), dim3(1024), 0, s, col_total.p, W - 2, base.p, counters.p + F_VERTS));
        TNB_LAUNCH_CHECK();
        if ((rc = read_small(counters.p, h, F_NUM, s))) return rc;
        const int64_t T = h[F_VERTS];
        m->T = T;
        TNB_CUDA(m->tri.reserve((size_t)std::max<int64_t>(T, 1) * 3));
        if (T > 0) {
            TNB_CUDA(launch_pdl(k_fan_write, dim3(fblocks), dim3(kFanThreads), W * sizeof(int), s, m->poly.p, m->pstart.p, m->pcnt.p, P, W, off.p, base.p, m->tri.p));
            TNB_LAUNCH_CHECK();
        }
    }
    return TNB_OK;  // locals are released in stream order
}

static void extract_release_scratch(tnb_mesh *m)
{
    m->surf.release(); m->used.release(); m->counters.release(); m->tmp_edges.release();
    m->begun = false;
}

// ---- read back -------------------------------------------------------------------------------------
__global__ void k_i32_to_i64(const int *__restrict__ src, int64_t n, int64_t *__restrict__ dst)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// the polygon rows as the reference's padded [P][W] tensor (-1 where blank), int64
__global__ void k_rows_padded(const int *__restrict__ rows, const int *__restrict__ row_start, const int *__restrict__ row_cnt, int64_t P, int W,
                              int64_t *__restrict__ dst)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P * W; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / W;
        const int j = (int)(i - p * W);
        dst[i] = j < row_cnt[p] ? rows[row_start[p] + j] : -1;
    }
}
__global__ void k_tri_positions(const int *__restrict__ tri, int64_t T, const float *__restrict__ vert, float *__restrict__ faces)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T * 3; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = tri[i];
        faces[3 * i] = vert[3 * (int64_t)v];
        faces[3 * i + 1] = vert[3 * (int64_t)v + 1];
        faces[3 * i + 2] = vert[3 * (int64_t)v + 2];
    }
}

}  // namespace tnb

using namespace tnb;

extern "C" {

int tnb_extract_mesh(const tnb_net *net, const tnb_complex *c, float eps, tnb_mesh **out, void *stream)
{
    if (!net || !c || !out) { set_error("tnb_extract_mesh: null argument"); return TNB_ERR_INVALID; }
    current_stream() = (cudaStream_t)stream;
    tnb_mesh *m = new tnb_mesh();
    tnb_complex *cc = const_cast<tnb_complex *>(c);
    int rc = extract_begin_impl(net, cc, eps, m, (cudaStream_t)stream);
    if (rc == TNB_OK) rc = extract_finish_impl(net, cc, m, (cudaStream_t)stream);
    extract_release_scratch(m);
    if (rc != TNB_OK) { delete m; *out = nullptr; return rc; }
    *out = m;
    return TNB_OK;
}

// The two halves separately: a driver that runs several slabs of one object on ONE device
// calls _begin for every slab before _finish for any (the exchange sits in between).
int tnb_extract_mesh_begin(const tnb_net *net, tnb_complex *c, float eps, tnb_mesh **out, void *stream)
{
    if (!net || !c || !out) { set_error("tnb_extract_mesh_begin: null argument"); return TNB_ERR_INVALID; }
    current_stream() = (cudaStream_t)stream;
    tnb_mesh *m = new tnb_mesh();
    int rc = extract_begin_impl(net, c, eps, m, (cudaStream_t)stream);
    if (rc != TNB_OK) { delete m; *out = nullptr; return rc; }
    *out = m;
    return TNB_OK;
}
int tnb_extract_mesh_finish(const tnb_net *net, tnb_complex *c, tnb_mesh *m, void *stream)
{
    if (!net || !c || !m || !m->begun) { set_error("tnb_extract_mesh_finish: bad argument"); return TNB_ERR_INVALID; }
    current_stream() = (cudaStream_t)stream;
    int rc = extract_finish_impl(net, c, m, (cudaStream_t)stream);
    extract_release_scratch(m);
    return rc;
}
int tnb_mesh_read_tags(const tnb_mesh *m, uint8_t *d_tags, void *stream)
{
    if (!m || !d_tags) { set_error("tnb_mesh_read_tags: null argument"); return TNB_ERR_INVALID; }
    if (m->V) TNB_CUDA(cudaMemcpyAsync(d_tags, m->tag.p, (size_t)m->V, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TNB_OK;
}

int tnb_mesh_read_vertex_index(const tnb_mesh *m, int64_t *d_index, void *stream)
{
    if (!m || !d_index) { set_error("tnb_mesh_read_vertex_index: null argument"); return TNB_ERR_INVALID; }
    if (m->V) {
        k_i32_to_i64<<<grid_for(m->V, 256), 256, 0, (cudaStream_t)stream>>>(m->vidx.p, m->V, d_index);
        TNB_LAUNCH_CHECK();
    }
    return TNB_OK;
}

void tnb_mesh_destroy(tnb_mesh *m) { delete m; }
int64_t tnb_mesh_num_vertices(const tnb_mesh *m) { return m ? m->V : 0; }
int64_t tnb_mesh_num_edges(const tnb_mesh *m) { return m ? m->E : 0; }
int64_t tnb_mesh_num_triangles(const tnb_mesh *m) { return m ? m->T : 0; }
int64_t tnb_mesh_num_polygons(const tnb_mesh *m) { return m ? m->P : 0; }
int64_t tnb_mesh_polygon_width(const tnb_mesh *m) { return m ? m->W : 0; }
int64_t tnb_mesh_near_plane(const tnb_mesh *m) { return m ? m->near_plane : 0; }

int tnb_mesh_read(const tnb_mesh *m, float *d_vertices, int64_t *d_edges, int64_t *d_triangles, float *d_faces,
                  int64_t *d_polygons, void *stream)
{
    if (!m) { set_error("tnb_mesh_read: null mesh"); return TNB_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    current_stream() = s;
    if (d_vertices && m->V) TNB_CUDA(cudaMemcpyAsync(d_vertices, m->vert.p, (size_t)m->V * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (d_edges && m->E) {
        k_i32_to_i64<<<grid_for(m->E * 2, 256), 256, 0, s>>>((const int *)m->edges.p, m->E * 2, d_edges);
        TNB_LAUNCH_CHECK();
    }
    if (d_triangles && m->T) {
        k_i32_to_i64<<<grid_for(m->T * 3, 256), 256, 0, s>>>(m->tri.p, m->T * 3, d_triangles);
        TNB_LAUNCH_CHECK();
    }
    if (d_faces && m->T) {
        k_tri_positions<<<grid_for(m->T * 3, 256), 256, 0, s>>>(m->tri.p, m->T, m->vert.p, d_faces);
        TNB_LAUNCH_CHECK();
    }
    if (d_polygons && m->P) {
        k_rows_padded<<<grid_for(m->P * m->W, 256), 256, 0, s>>>(m->poly.p, m->pstart.p, m->pcnt.p, m->P, (int)m->W, d_polygons);
        TNB_LAUNCH_CHECK();
    }
    return TNB_OK;
}

int tnb_mesh_read_host(const tnb_mesh *m, float *h_vertices, int64_t *h_triangles, float *h_faces, int64_t *h_polygons)
{
    if (!m) { set_error("tnb_mesh_read_host: null mesh"); return TNB_ERR_INVALID; }
    TNB_CUDA(cudaDeviceSynchronize());
    current_stream() = nullptr;
    DevBuf<int64_t> t64, p64;
    DevBuf<float> f;
    if (h_vertices && m->V) TNB_CUDA(cudaMemcpy(h_vertices, m->vert.p, (size_t)m->V * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    if (h_triangles && m->T) {
        TNB_CUDA(t64.reserve((size_t)m->T * 3));
        k_i32_to_i64<<<grid_for(m->T * 3, 256), 256>>>(m->tri.p, m->T * 3, t64.p);
        TNB_LAUNCH_CHECK();
        TNB_CUDA(cudaMemcpy(h_triangles, t64.p, (size_t)m->T * 3 * sizeof(int64_t), cudaMemcpyDeviceToHost));
    }
    if (h_faces && m->T) {
        TNB_CUDA(f.reserve((size_t)m->T * 9));
        k_tri_positions<<<grid_for(m->T * 3, 256), 256>>>(m->tri.p, m->T, m->vert.p, f.p);
        TNB_LAUNCH_CHECK();
        TNB_CUDA(cudaMemcpy(h_faces, f.p, (size_t)m->T * 9 * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (h_polygons && m->P) {
        TNB_CUDA(p64.reserve((size_t)m->P * m->W));
        k_rows_padded<<<grid_for(m->P * m->W, 256), 256>>>(m->poly.p, m->pstart.p, m->pcnt.p, m->P, (int)m->W, p64.p);
        TNB_LAUNCH_CHECK();
        TNB_CUDA(cudaMemcpy(h_polygons, p64.p, (size_t)m->P * m->W * sizeof(int64_t), cudaMemcpyDeviceToHost));
    }
    return TNB_OK;
}

int tnb_subpoly(const tnb_net *net, float size, float eps, int32_t force, int32_t unit, tnb_mesh **out, void *stream)
{
    if (!net || !out) { set_error("tnb_subpoly: null argument"); return TNB_ERR_INVALID; }
    *out = nullptr;
    const double scale0 = t_capacity_scale;
    int rc = TNB_OK;
    // The steps run without host syncs; if the device reports that the work arrays were too
    // small (sticky capacity bit), the extraction is simply repeated with twice the head-room.
    for (int attempt = 0; attempt < 4; ++attempt) {
        tnb_complex *c = nullptr;
        rc = tnb_skeleton(net, unit, size, &c, stream);
        if (rc != TNB_OK) break;
        const int H = net->meta.H, NL = net->meta.NLIN;
        std::vector<int32_t> lh;  // every hidden neuron in layer order, then the output neuron (subpoly.py:58-72)
        for (int l = 0; l < NL - 1; ++l)
            for (int h = 0; h < H; ++h) { lh.push_back(l); lh.push_back(h); }
        lh.push_back(NL - 2);
        lh.push_back(H);
        rc = tnb_subpoly_steps(net, c, lh.data(), (int32_t)(lh.size() / 2), eps, force, stream);
        if (rc == TNB_OK) rc = tnb_extract_mesh(net, c, eps, out, stream);
        tnb_complex_destroy(c);
        if (rc != TNB_ERR_CAPACITY) break;
        t_capacity_scale *= 2.0;
    }
    t_capacity_scale = scale0;
    return rc;
}

}  // extern "C"

// ---- many objects in one call ---------------------------------------------------------------------
// A small extraction is a chain of short kernels with a handful of host decisions in between (sizes of the next
// arrays): one object cannot fill the GPU, eight can.  The workers below are host threads that live as long as the
// library, each with a stream of its own; a batch is dealt to them object by object.  Every worker runs the
// ordinary single-object path (same kernels, same results); complexes small enough for it take the
// thread-block-cluster form of the persistent step kernel (16 SMs each), so the step loops of the objects in
// flight are resident side by side.  The caller's stream is a dependency of every worker stream and depends on
// all of them at the end: to the caller the batch is one stream-ordered operation.
namespace tnb {
struct BatchJob {
    const tnb_net *const *nets = nullptr;
    tnb_mesh **out = nullptr;
    int count = 0;
    float size = 1.2f, eps = 1e-4f;
    int force = 1, unit = 128;
    cudaEvent_t ready = nullptr;
    std::atomic<int> next{0};
    std::vector<int> rc;
    std::vector<std::string> err;
};
struct BatchPool {
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> threads;
    std::vector<cudaStream_t> streams;
    std::vector<cudaEvent_t> done_ev;
    BatchJob *job = nullptr;
    uint64_t generation = 0;
    int wanted = 0;     // workers that take part in the current job
    int running = 0;
    int device = 0;
    std::mutex call_mu;  // one batch at a time

    void worker(int w)
    {
        cudaSetDevice(device);
        uint64_t seen = 0;
        for (;;) {
            BatchJob *j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return generation != seen; });
                seen = generation;
                if (w >= wanted) continue;
                j = job;
            }
            cudaStream_t s = streams[w];
            cudaStreamWaitEvent(s, j->ready, 0);
            t_cluster_max_items = 200000;
            t_no_profile = true;
            for (int i = j->next.fetch_add(1); i < j->count; i = j->next.fetch_add(1)) {
                j->rc[i] = tnb_subpoly(j->nets[i], j->size, j->eps, j->force, j->unit, &j->out[i], (void *)s);
                if (j->rc[i] != TNB_OK) j->err[i] = tnb_last_error();
            }
            t_cluster_max_items = -1;
            cudaEventRecord(done_ev[w], s);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--running == 0) cv_done.notify_all();
            }
        }
    }
    int ensure(int n)
    {
        while ((int)threads.size() < n) {
            cudaStream_t s;
            cudaEvent_t e;
            TNB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
            TNB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            streams.push_back(s);
            done_ev.push_back(e);
            const int w = (int)threads.size();
            threads.emplace_back([this, w] { worker(w); });
            threads.back().detach();
        }
        return TNB_OK;
    }
};
static BatchPool &batch_pool()
{
    static BatchPool *p = new BatchPool;  // never destroyed: its threads wait for work until the process ends
    return *p;
}
}  // namespace tnb

extern "C" int tnb_subpoly_batch(const tnb_net *const *nets, int32_t count, float size, float eps, int32_t force, int32_t unit,
                                 int32_t in_flight, tnb_mesh **out, int32_t *rcs, void *stream)
{
    if (count < 0 || (count > 0 && (!nets || !out))) { set_error("tnb_subpoly_batch: null argument"); return TNB_ERR_INVALID; }
    for (int i = 0; i < count; ++i) {
        out[i] = nullptr;
        if (rcs) rcs[i] = TNB_OK;
        if (!nets[i]) { set_error("tnb_subpoly_batch: null network"); return TNB_ERR_INVALID; }
    }
    if (count == 0) return TNB_OK;
    if (in_flight <= 0) in_flight = 8;   // 16-SM clusters: nine fit a B200
    const int workers = std::min<int>(std::min<int>(in_flight, count), 32);
    cudaStream_t s = (cudaStream_t)stream;
    BatchPool &bp = batch_pool();
    std::lock_guard<std::mutex> call(bp.call_mu);
    int dev = 0;
    TNB_CUDA(cudaGetDevice(&dev));
    if (!bp.threads.empty() && dev != bp.device) { set_error("tnb_subpoly_batch: the worker pool belongs to another device"); return TNB_ERR_UNSUPPORTED; }
    bp.device = dev;
    int rc = bp.ensure(workers);
    if (rc != TNB_OK) return rc;
    BatchJob job;
    job.nets = nets; job.out = out; job.count = count;
    job.size = size; job.eps = eps; job.force = force; job.unit = unit;
    job.rc.assign(count, TNB_OK);
    job.err.assign(count, std::string());
    TNB_CUDA(cudaEventCreateWithFlags(&job.ready, cudaEventDisableTiming));
    TNB_CUDA(cudaEventRecord(job.ready, s));
    {
        std::lock_guard<std::mutex> lk(bp.mu);
        bp.job = &job;
        bp.wanted = workers;
        bp.running = workers;
        ++bp.generation;
    }
    bp.cv_work.notify_all();
    {
        std::unique_lock<std::mutex> lk(bp.mu);
        bp.cv_done.wait(lk, [&] { return bp.running == 0; });
        bp.job = nullptr;
    }
    for (int w = 0; w < workers; ++w) cudaStreamWaitEvent(s, bp.done_ev[w], 0);
    cudaEventDestroy(job.ready);
    rc = TNB_OK;
    for (int i = 0; i < count; ++i) {
        if (rcs) rcs[i] = job.rc[i];
        if (job.rc[i] != TNB_OK && rc == TNB_OK) { rc = job.rc[i]; set_error("object " + std::to_string(i) + ": " + job.err[i]); }
    }
    return rc;
}

// complex.cu -- skeleton sweep and per-hyperplane edge subdivision on the device.
//
// Replaces TropicalHashGrid.skeleton (tropical.py:158-225) and subpoly_ (subpoly.py:90-279,
// planar branch) including check_edges_with_new_vertices (subpoly_debug.py:33-51),
// edge_vertices / regions_to_vertices (subpoly.py:281-340, :484-535) and the pruning block
// (subpoly.py:252-277).  Every compaction is order preserving (scan.cuh), so vertex and
// edge numbering equal the reference's.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include <cooperative_groups.h>

#include "complex.cuh"
#include "cells.cuh"
#include "curve.cuh"
#include "repair.cuh"
#include "halo.cuh"
#include "net_eval.cuh"
#include "scan.cuh"
#include "sort.cuh"

tnb_complex::~tnb_complex() {}

namespace tnb {

double g_capacity_factor = 4.0;
thread_local double t_capacity_scale = 1.0;  // the capacity retry of tnb_subpoly (per calling thread)
double capacity_factor() { return g_capacity_factor * t_capacity_scale; }
constexpr int kThreads = 128;
constexpr int kCachedPartners = 32;  // the partner-count pass keeps this many partners per candidate: the write
                                    // pass of such a list does not walk the buckets again
constexpr int kNetworkPartners = 12; // lists up to this size (nearly all of them) are sorted in registers
// device counters: [0, C_V) are cleared at the start of every step, C_V / C_E hold the complex size
//   C_RAW = edges the plane crosses, C_SPLIT = edges actually split (== C_RAW on the planar path,
//   fewer after strict_check on the curve path)
//   C_VPAR / C_EPAR = which half of the ping-pong vertex / edge arrays is current,
//   C_STICKY = error bits that survive steps (sync-free fused path)
//   C_KEPT = edges kept by pruning, parked until the slab exchange decides whether the step counts
//   C_APAR = which half of the liveness array is current.  C_V counts vertex SLOTS (dead rows included)
//   C_LONG = candidates whose partner list is longer than the cache (their connecting edges are a warp's job)
enum { C_SPLIT = 0, C_FLAG, C_HIT, C_PAIRS, C_CAND, C_OVERFLOW, C_RAW, C_ERR, C_V = 8, C_E = 9, C_VPAR = 10, C_EPAR = 11, C_STICKY = 12, C_KEPT = 13, C_APAR = 14, C_LONG = 15, C_CROSS = 16 /* 64-bit crossing mask */, C_RECS = 18 /* records in the contiguous cell segments */,
       C_STEP = 19 /* device-driven step stream: next position in the step list */, C_IDX = 20 /* its current hyperplane column, -1 = none */, C_PRUNE = 21, C_NUM = 32 };
enum { kErrNoPlane = 1, kErrGradientDescent = 2 /* the repair ended off the planes, or too many walks */, kErrRepair = 4 /* walks filed: run the repair */ };
enum { kStickyCapacity = 1, kStickyNoPlane = 32, kStickyGradientDescent = 64 };  // 2..16: halo.cuh

// ---- allocation ---------------------------------------------------------------------------
// one bit per edge and, behind them, one bit per vertex (the split and the hit compaction of a step can run
// side by side); a compaction over n items needs (n + 31) / 32 + kScanMaxBlocks words (warp-aligned slices)
static size_t scan_mask_vertex_offset(size_t Ecap) { return (Ecap + 31) / 32 + kScanMaxBlocks + 64; }
static size_t scan_mask_words(size_t Vcap, size_t Ecap) { return scan_mask_vertex_offset(Ecap) + (Vcap + 31) / 32 + kScanMaxBlocks + 64; }

int complex_alloc(tnb_complex *c, const tnb_net *net, size_t Vcap, size_t Ecap)
{
    c->R = net->meta.R;
    c->Vcap = Vcap;
    c->Ecap = Ecap;
    for (int k = 0; k < 2; ++k) {
        TNB_CUDA(c->vert[k].reserve(Vcap * 3));
        TNB_CUDA(c->out[k].reserve(Vcap * c->R));
        TNB_CUDA(c->sig[k].reserve(Vcap * 3));
        TNB_CUDA(c->edges[k].reserve(Ecap));
        TNB_CUDA(c->tag[k].reserve(Vcap));
        TNB_CUDA(cudaMemsetAsync(c->tag[k].p, 0, Vcap, current_stream()));
        TNB_CUDA(c->used[k].reserve(Vcap));
        TNB_CUDA(cudaMemsetAsync(c->used[k].p, 1, Vcap * sizeof(int), current_stream()));  // nonzero = alive
    }
    TNB_CUDA(c->split_list.reserve(Ecap));
    TNB_CUDA(c->bmask.reserve(Ecap));
    TNB_CUDA(c->cand.reserve(Vcap));
    TNB_CUDA(c->pcount.reserve(Vcap));
    TNB_CUDA(c->poff.reserve(Vcap));
    TNB_CUDA(c->pcache.reserve(Vcap * kCachedPartners));
    TNB_CUDA(c->next.reserve(Vcap * 8));
    TNB_CUDA(c->cslot.reserve(Vcap * 8));
    TNB_CUDA(c->scan_mask.reserve(scan_mask_words(Vcap, Ecap)));
    TNB_CUDA(c->remap.reserve(Vcap));
    TNB_CUDA(c->block_sums.reserve(3 * kScanMaxBlocks));  // the persistent step kernels keep three sets of block sums
    TNB_CUDA(c->counters.reserve(C_NUM));
    TNB_CUDA(c->gd.reserve(kGdInts));
    k_gd_reset<<<1, 32, 0, current_stream()>>>(c->gd.p);
    TNB_CUDA(c->bytes.reserve(4));
    TNB_CUDA(cudaMemsetAsync(c->bytes.p, 0, 4 * sizeof(unsigned long long), current_stream()));
    TNB_CUDA(cudaMemsetAsync(c->counters.p, 0, C_NUM * sizeof(int), current_stream()));
    {   // one pinned mirror per thread, reused by every complex
        static thread_local int *pinned = nullptr;
        if (!pinned) TNB_CUDA(cudaMallocHost((void **)&pinned, C_NUM * sizeof(int)));
        c->h_counters = pinned;
    }
    // cell buckets: offsets live in [-1, M-1], cells in [-2, M-1] -> M+2 per axis
    c->cell_dim = net->meta.n_marks + 2;
    c->n_cells = (int64_t)c->cell_dim * c->cell_dim * c->cell_dim;
    if (c->n_cells > (int64_t)1 << 31) {
        set_error("marks grid too fine for the dense cell buckets (n_marks " + std::to_string(net->meta.n_marks) + ")");
        return TNB_ERR_UNSUPPORTED;
    }
    TNB_CUDA(c->head.reserve((size_t)c->n_cells));
    TNB_CUDA(cudaMemsetAsync(c->head.p, 0, (size_t)c->n_cells * sizeof(unsigned long long), current_stream()));
    c->stamp = 0;
    return TNB_OK;
}

template <class T>
static int grow(DevBuf<T> &b, size_t new_elems, size_t keep_elems, cudaStream_t s)
{
    if (new_elems <= b.cap) return TNB_OK;
    DevBuf<T> nb;
    TNB_CUDA(nb.reserve(new_elems));
    if (keep_elems) TNB_CUDA(cudaMemcpyAsync(nb.p, b.p, keep_elems * sizeof(T), cudaMemcpyDeviceToDevice, s));
    TNB_CUDA(cudaStreamSynchronize(s));
    b.swap(nb);
    return TNB_OK;
}

// make room for Vneed vertices / Eneed edges, keeping the current contents
int complex_reserve(tnb_complex *c, size_t Vneed, size_t Eneed, cudaStream_t s)
{
    if (Vneed > c->Vcap || Eneed > c->Ecap) {
        int rc = grow(c->scan_mask, scan_mask_words(std::max(Vneed, (size_t)(c->Vcap * 2)), std::max(Eneed, (size_t)(c->Ecap * 2))), 0, s);
        if (rc) return rc;
    }
    if (Vneed > c->Vcap) {
        size_t nc = std::max(Vneed, (size_t)(c->Vcap * 2));
        int rc;
        for (int k = 0; k < 2; ++k) {
            size_t keep = (k == c->vcur) ? (size_t)c->V : 0;
            if ((rc = grow(c->vert[k], nc * 3, keep * 3, s))) return rc;
            if ((rc = grow(c->out[k], nc * c->R, keep * c->R, s))) return rc;
            if ((rc = grow(c->sig[k], nc * 3, keep * 3, s))) return rc;
            if ((rc = grow(c->tag[k], nc, keep, s))) return rc;
            if ((rc = grow(c->used[k], nc, (k == c->acur) ? (size_t)c->V : 0, s))) return rc;
        }
        if ((rc = grow(c->cand, nc, 0, s)) || (rc = grow(c->pcount, nc, 0, s)) || (rc = grow(c->poff, nc, 0, s)) || (rc = grow(c->pcache, nc * kCachedPartners, 0, s)) ||
            (rc = grow(c->next, nc * 8, 0, s)) || (rc = grow(c->cslot, nc * 8, 0, s)) || (rc = grow(c->remap, nc, 0, s)))
            return rc;
        c->Vcap = nc;
    }
    if (Eneed > c->Ecap) {
        size_t nc = std::max(Eneed, (size_t)(c->Ecap * 2));
        int rc;
        for (int k = 0; k < 2; ++k)
            if ((rc = grow(c->edges[k], nc, (k == c->ecur) ? (size_t)c->E : 0, s))) return rc;
        if ((rc = grow(c->split_list, nc, 0, s)) || (rc = grow(c->bmask, nc, 0, s))) return rc;
        c->Ecap = nc;
    }
    return TNB_OK;
}

__global__ void k_set_counts(int *__restrict__ cnt, int V, int E, int vpar, int epar, int apar);
__global__ void k_cross_mask(const int2 *__restrict__ edges, int64_t E, const uint64_t *__restrict__ sig, int *__restrict__ cnt);
// which hyperplanes cross an edge of a fresh complex (its packed signs must be in place)
static int initial_cross_mask(tnb_complex *c, cudaStream_t s)
{
    if (c->E > 0) {
        k_cross_mask<<<grid_for(c->E, 256), 256, 0, s>>>(c->cedges(), c->E, c->csig(), c->counters.p);
        TNB_LAUNCH_CHECK();
    }
    c->cross_stale = true;
    return TNB_OK;
}

static int read_counters(tnb_complex *c, cudaStream_t s)
{
    TNB_CUDA(cudaMemcpyAsync(c->h_counters, c->counters.p, C_NUM * sizeof(int), cudaMemcpyDeviceToHost, s));
    TNB_CUDA(cudaStreamSynchronize(s));
    return TNB_OK;
}

// ================================================================================================
// skeleton
// ================================================================================================
struct ChunkSeg {  // one (chunk, axis) block of candidate grid edges, in the reference's order
    int64_t first;  // first slot number
    int s[3], n[3]; // chunk start / size per axis
    int axis, chunk;
    int64_t row_first;  // first ROW of the block: a row = the slots that differ in the last coordinate only
    int64_t bit_word;   // first word of the chunk piece's "close to the surface" bits ([n0][n1][w32] words)
    int w32;            // words per bit row
};
constexpr int kMaxSegs = 3 * 512;

// |tanh(sdf)| and |grad| of every marks-grid vertex of every chunk piece, per-chunk max |grad|: ONE launch.
// A CTA belongs to one piece (cta_first = its first CTA); the pieces get CTAs in proportion to their vertices,
// four trips of two vertices per thread, so that the hardware back-fills CTAs across pieces.  As one launch
// per chunk (8 for the 201^3 grid) every launch ended in its own tail: the 74^3 corner chunk ran at 5.9 G
// vertices/s against 15.7 for the 128^3 one (profiles/r2_ncu_launches_large_sphere.txt).
struct SweepPiece {
    int s[3], n[3];   // first grid vertex / vertices per axis
    int chunk;        // whose max |grad| this piece feeds
    int cta_first, ctas;
    LatticeStride ls; // split of the piece's thread stride (ctas * kThreads)
};
template <class C, int MINB = 4>
__global__ void __launch_bounds__(kThreads, MINB) k_sweep_pieces(const __grid_constant__ NetMeta n, int M, const SweepPiece *__restrict__ pieces,
                                                           int n_pieces, float *__restrict__ dist, unsigned *__restrict__ max_grad, int sync_trips)
{
    __shared__ SweepPiece s_piece;
    if (threadIdx.x == 0) {
        int a = 0, b = n_pieces - 1;  // last piece whose first CTA is <= blockIdx.x
        while (a < b) {
            const int mid = (a + b + 1) >> 1;
            if (pieces[mid].cta_first <= (int)blockIdx.x) a = mid; else b = mid - 1;
        }
        s_piece = pieces[a];
    }
    __syncthreads();
    const SweepPiece pc = s_piece;
    const int sx = pc.s[0], sy = pc.s[1], sz = pc.s[2], nx = pc.n[0], ny = pc.n[1], nz = pc.n[2];
    const int64_t count = (int64_t)nx * ny * nz;
    float local = 0.0f;
    const int64_t first = ((int64_t)blockIdx.x - pc.cta_first) * blockDim.x + threadIdx.x, stride = (int64_t)pc.ctas * blockDim.x;
    Lattice3 at(first < count ? first : 0, pc.ls, nx, ny);
    if constexpr (C::kFixed) {
        // two grid vertices per trip (t and t + stride): both networks in packed FFMA2s (net_eval.cuh).  The trip count
        // is the CTA's (a thread past the end evaluates its last valid vertex again and stores nothing), so that the
        // warps of a CTA can start every trip together (sync_trips): ~100 KB of straight-line code per trip and 16 warps
        // per SM at 16 different places of it made instruction fetch the top stall of this kernel.
        const int64_t cta_first = ((int64_t)blockIdx.x - pc.cta_first) * blockDim.x;
        for (int64_t tb = cta_first; tb < count; tb += 2 * stride) {
            if (sync_trips) __syncthreads();
            const int64_t t = tb + threadIdx.x;
            const bool one = t < count, two = t + stride < count;
            const int gi0 = one ? sx + at.ix : sx, gj0 = one ? sy + at.iy : sy, gk0 = one ? sz + at.iz : sz;  // past the end: the piece's first vertex, not stored
            if (one) at.advance();
            const int gi1 = two ? sx + at.ix : gi0, gj1 = two ? sy + at.iy : gj0, gk1 = two ? sz + at.iz : gk0;
            if (two) at.advance();
            // preprocess_inverse(marks[...]) (tropical.py:186, model.py:81-82)
            float x0[3] = {n.marks[gi0] * n.pre_2s - n.pre_scale, n.marks[gj0] * n.pre_2s - n.pre_scale, n.marks[gk0] * n.pre_2s - n.pre_scale};
            float x1[3] = {n.marks[gi1] * n.pre_2s - n.pre_scale, n.marks[gj1] * n.pre_2s - n.pre_scale, n.marks[gk1] * n.pre_2s - n.pre_scale};
            float sdf[2], g[2][3];
            sdf_grad_pair<C>(n, x0, x1, sdf, g);
            if (one) {
                dist[((int64_t)gi0 * M + gj0) * M + gk0] = fabsf(sdf[0]);
                local = fmaxf(local, grad_norm(g[0]));
            }
            if (two) {
                dist[((int64_t)gi1 * M + gj1) * M + gk1] = fabsf(sdf[1]);
                local = fmaxf(local, grad_norm(g[1]));
            }
        }
    } else {
        for (int64_t t = first; t < count; t += stride, at.advance()) {
            // x is the lane axis: the hash table is x-fastest, so a warp's gathers touch
            // consecutive entries (the |sdf| store is the only strided access)
            const int gi = sx + at.ix, gj = sy + at.iy, gk = sz + at.iz;
            float x[3] = {n.marks[gi] * n.pre_2s - n.pre_scale, n.marks[gj] * n.pre_2s - n.pre_scale,
                          n.marks[gk] * n.pre_2s - n.pre_scale};
            float g[3];
            const float sdf = sdf_grad<C>(n, x, g, true);
            dist[((int64_t)gi * M + gj) * M + gk] = fabsf(sdf);
            local = fmaxf(local, grad_norm(g));
        }
    }
    // block max (non-negative floats order like their bit patterns)
    unsigned v = __float_as_uint(local);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    __shared__ unsigned s[kThreads / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w) v = max(v, s[w]);
        atomicMax(max_grad + pc.chunk, v);
    }
}

// ---- edge selection, row by row -------------------------------------------------------------------------
// A candidate edge is kept iff both its grid vertices are within the chunk's threshold of the surface
// (tropical.py:113-138).  Per SLOT that is a decode (two divisions) and two scattered loads for 24 M slots of
// which 7 % pass.  Per ROW (the slots of a block that differ in the last coordinate only, up to `unit` of them)
// it is a few word operations: k_skel_bits leaves one bit per (chunk piece, grid vertex) -- "close enough" under
// that chunk's threshold, coalesced along the last axis --, and a row's kept slots are  bits(row) & bits(row
// shifted by one vertex along the block's axis).  The ordered compaction then runs over ~0.4 M rows, each
// emitting its edges in ascending last coordinate: the reference's order (chunk, axis, i, j, k).
constexpr int kMaxRowWords = 32;  // unit <= 1024
__global__ void __launch_bounds__(256) k_skel_bits(const ChunkSeg *__restrict__ segs, int n_segs, int M, const float *__restrict__ dist,
                                                   const unsigned *__restrict__ max_grad, float k_len, uint32_t *__restrict__ bits)
{
    // one warp per bit word; blockIdx.y = segment (only a piece's first block writes its bits)
    const ChunkSeg sg = segs[blockIdx.y];
    if (blockIdx.y > 0 && segs[blockIdx.y - 1].bit_word == sg.bit_word) return;  // same chunk piece as the block before
    const float eps = k_len * __uint_as_float(max_grad[sg.chunk]);
    const int lane = threadIdx.x & 31;
    const int64_t n_words = (int64_t)sg.n[0] * sg.n[1] * sg.w32;
    for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_words; w += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int kw = (int)(w % sg.w32);
        const int64_t row = w / sg.w32;
        const int j = (int)(row % sg.n[1]), i = (int)(row / sg.n[1]), k = kw * 32 + lane;
        bool ok = false;
        if (k < sg.n[2]) ok = dist[((int64_t)(sg.s[0] + i) * M + (sg.s[1] + j)) * M + sg.s[2] + k] <= eps;
        const uint32_t word = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) bits[sg.bit_word + w] = word;
    }
}
struct SkelRows {
    const ChunkSeg *segs;
    int n_segs, M;
    const uint32_t *bits;
    mutable int hint = 0;
    __device__ __forceinline__ int seg_of(int64_t r) const
    {
        int a = hint;
        if (segs[a].row_first <= r && (a + 1 == n_segs || r < segs[a + 1].row_first)) return a;
        a = 0;
        int b = n_segs - 1;
        while (a < b) {
            const int mid = (a + b + 1) >> 1;
            if (segs[mid].row_first <= r) a = mid; else b = mid - 1;
        }
        hint = a;
        return a;
    }
    // kept slots of row r as bit words (bit k of word w = slot k + 32 w of the row); returns the number of words
    __device__ __forceinline__ int row_words(int64_t r, uint32_t *out, int &lo0, int &step) const
    {
        const ChunkSeg sg = segs[seg_of(r)];
        int dims1 = sg.n[1] - (sg.axis == 1 ? 1 : 0);
        const int rr = (int)(r - sg.row_first), i = rr / dims1, j = rr - i * dims1;
        const uint32_t *a = bits + sg.bit_word + ((int64_t)i * sg.n[1] + j) * sg.w32;
        const int len = sg.n[2] - (sg.axis == 2 ? 1 : 0);  // slots of the row
        if (sg.axis == 2) {
            for (int w = 0; w < sg.w32; ++w) {
                const uint32_t x = a[w], nx = w + 1 < sg.w32 ? a[w + 1] : 0u;
                out[w] = x & ((x >> 1) | (nx << 31));
            }
        } else {
            const uint32_t *b = a + (sg.axis == 0 ? (int64_t)sg.n[1] * sg.w32 : sg.w32);
            for (int w = 0; w < sg.w32; ++w) out[w] = a[w] & b[w];
        }
        if (len & 31) out[len >> 5] &= (1u << (len & 31)) - 1u;  // slots past the end of the row
        for (int w = (len + 31) >> 5; w < sg.w32; ++w) out[w] = 0u;
        lo0 = ((sg.s[0] + i) * M + (sg.s[1] + j)) * M + sg.s[2];
        step = sg.axis == 0 ? M * M : (sg.axis == 1 ? M : 1);
        return sg.w32;
    }
};
struct SkelRowCount {
    SkelRows q;
    __device__ __forceinline__ int operator()(int64_t r) const
    {
        uint32_t w[kMaxRowWords];
        int lo0, step;
        const int n = q.row_words(r, w, lo0, step);
        int c = 0;
        for (int i = 0; i < n; ++i) c += __popc(w[i]);
        return c;
    }
};
struct SkelRowEmit {
    SkelRows q;
    int2 *edges;
    int *used;
    __device__ __forceinline__ void operator()(int64_t r, int pos, int) const
    {
        uint32_t w[kMaxRowWords];
        int lo0, step;
        const int n = q.row_words(r, w, lo0, step);
        for (int i = 0; i < n; ++i)
            for (uint32_t m = w[i]; m; m &= m - 1) {
                const int lo = lo0 + 32 * i + __ffs(m) - 1, hi = lo + step;
                edges[pos++] = make_int2(hi, lo);  // (indices[1:], indices[:-1]) column order, tropical.py:130
                used[hi] = 1;
                used[lo] = 1;
            }
    }
};

struct FlagCount {
    const int *flag;
    __device__ __forceinline__ int operator()(int64_t i) const { return flag[i] ? 1 : 0; }
};
struct SkelVertEmit {
    NetMeta const *unused;
    const float *marks;
    float pre_2s, pre_scale;
    int M;
    int *remap;
    float *vert;
    unsigned char *tag;
    int x0;              // first plane of the slab: item v is grid vertex v + x0*M*M
    int plane_lo, plane_hi;  // planes shared with a neighbour rank (-1: none)
    __device__ __forceinline__ void operator()(int64_t v, int pos, int) const
    {
        remap[v] = pos;
        const int k = (int)(v % M), j = (int)((v / M) % M), i = (int)(v / ((int64_t)M * M)) + x0;
        vert[3 * pos] = marks[i] * pre_2s - pre_scale;
        vert[3 * pos + 1] = marks[j] * pre_2s - pre_scale;
        vert[3 * pos + 2] = marks[k] * pre_2s - pre_scale;
        tag[pos] = (unsigned char)((i == plane_lo ? 1 : 0) | (i == plane_hi ? 2 : 0));
    }
};

__global__ void k_remap_edges(int2 *__restrict__ edges, int64_t E, const int *__restrict__ remap)
{
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        int2 ed = edges[e];
        edges[e] = make_int2(remap[ed.x], remap[ed.y]);
    }
}

// A warp's 32 rows are consecutive in `out` (32 * R floats): each lane builds its row in shared memory (stride R
// floats: R = 33 is odd, no bank conflicts), then the warp writes the block with coalesced stores.  One thread
// per 132-byte row cost 32 sectors per store instruction (k_vertex_outputs 0.28 TB/s, k_new_vertices lg_throttle
// 6.7 warps per issue in profiles/r2_ncu_summary.txt).
__device__ __forceinline__ void warp_store_rows(float *__restrict__ out, int64_t first_row, int rows, int R, const float *tile)
{
    const int lane = threadIdx.x & 31;
    float *dst = out + first_row * R;
    const int n = rows * R;
    for (int e = lane; e < n; e += 32) dst[e] = tile[e];
}

// outputs row + packed signs of vertices [first, first+count)
template <class C>
__global__ void __launch_bounds__(kThreads) k_vertex_outputs(const __grid_constant__ NetMeta n,
                                                             const float *__restrict__ vert, int64_t first,
                                                             int64_t count, float *__restrict__ out,
                                                             uint64_t *__restrict__ sig)
{
    extern __shared__ float s_tile[];  // [kThreads][R]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, R = n.R;
    float *wtile = s_tile + (size_t)warp * 32 * R, *row = wtile + lane * R;
    for (int64_t t0 = blockIdx.x * (int64_t)blockDim.x + warp * 32; t0 < count; t0 += (int64_t)gridDim.x * blockDim.x) {  // warp uniform
        const int64_t t = t0 + lane, v = first + t;
        if (t < count) {
            float x[3] = {vert[3 * v], vert[3 * v + 1], vert[3 * v + 2]}, xp[3];
            uint64_t pos, neg, big;
            outputs_row_packed<C>(n, x, row, n.eps, n.eps, xp, pos, neg, big);
            sig[3 * v] = pos;
            sig[3 * v + 1] = neg;
            sig[3 * v + 2] = pack_grid(n, n.marks, xp, n.eps);
        }
        __syncwarp();
        warp_store_rows(out, first + t0, (int)min((int64_t)32, count - t0), R, wtile);
        __syncwarp();
    }
}

static int eval_vertices(const tnb_net *net, tnb_complex *c, int64_t first, int64_t count, cudaStream_t s)
{
    if (count <= 0) return TNB_OK;
    unsigned g = grid_for(count, kThreads);
    prof_begin(TNB_PROF_VERTEX_ROWS, s);
    const size_t tile = (size_t)kThreads * net->meta.R * sizeof(float);
    if (net->fixed_cfg) k_vertex_outputs<CfgRef><<<g, kThreads, tile, s>>>(net->meta, c->cvert(), first, count, c->cout_(), c->csig());
    else k_vertex_outputs<CfgAny><<<g, kThreads, tile, s>>>(net->meta, c->cvert(), first, count, c->cout_(), c->csig());
    TNB_LAUNCH_CHECK();
    prof_end(TNB_PROF_VERTEX_ROWS, s, count, count * (12 + 4 * net->meta.R + 24) + (int64_t)net->table.cap * 8);
    return TNB_OK;
}

}  // namespace tnb

// State between the two halves of the skeleton: the sweep (|sdf|, per-chunk max |grad|) and the
// edge selection.  A slab-sharded run reduces max_grad over the ranks in between, because the
// reference's threshold is per CHUNK (tropical.py:189-197), not per slab.
struct tnb_sweep {
    int M = 0, unit = 0, n_chunks = 0;
    int x_lo = 0, x_hi = 0;  // planes of the slab along the first axis (the whole grid: 0, M-1)
    bool tag_lower = false, tag_upper = false;
    float k_len = 0.0f;
    int64_t slots = 0, rows = 0, bit_words = 0;
    std::vector<tnb::ChunkSeg> segs;
    std::vector<tnb::SweepPiece> pieces;   // host copy (source of the async upload)
    tnb::DevBuf<tnb::SweepPiece> d_pieces;
    tnb::DevBuf<float> dist;         // [(x_hi-x_lo+1) * M * M], plane x_lo first
    tnb::DevBuf<unsigned> max_grad;  // [n_chunks] bit patterns of non-negative floats
};

namespace tnb {

// launch = false: only the layout (segments, buffers); |sdf| and max |grad| are written by the caller
// (tnb_sweep_write_dist / tnb_sweep_write_max_grad: a sweep whose planes were evaluated on several GPUs)
static const int g_sweep_sync = std::getenv("TNB_SWEEP_SYNC") ? std::atoi(std::getenv("TNB_SWEEP_SYNC")) : 1;  // A/B switch (see k_sweep_pieces)
static int sweep_impl(const tnb_net *net, int unit, int x_lo, int x_hi, bool tag_lower, bool tag_upper, tnb_sweep *sw,
                      cudaStream_t s, bool launch = true)
{
    const int M = net->meta.n_marks;
    if (unit < 2) { set_error("tnb_skeleton: unit must be >= 2"); return TNB_ERR_INVALID; }
    if (x_lo < 0 || x_hi >= M || x_lo > x_hi) { set_error("tnb_skeleton_sweep: slab out of range"); return TNB_ERR_INVALID; }
    const int64_t M3 = (int64_t)M * M * M;
    if (M3 > (int64_t)1 << 31) { set_error("tnb_skeleton: marks grid too large"); return TNB_ERR_UNSUPPORTED; }
    // chunks exactly as range(0, L, unit - 1) enumerates them (tropical.py:176-181)
    std::vector<int> starts;
    for (int a = 0; a < M; a += unit - 1) starts.push_back(a);
    const int nc = (int)starts.size();
    if ((int64_t)nc * nc * nc * 3 > kMaxSegs) { set_error("tnb_skeleton: too many chunks"); return TNB_ERR_UNSUPPORTED; }
    float len_max = 0.0f;
    for (int i = 0; i + 1 < M; ++i) len_max = std::max(len_max, net->h_marks[i + 1] - net->h_marks[i]);
    sw->M = M; sw->unit = unit; sw->n_chunks = nc * nc * nc;
    sw->x_lo = x_lo; sw->x_hi = x_hi; sw->tag_lower = tag_lower; sw->tag_upper = tag_upper;
    sw->k_len = (std::sqrt(3.0f) * 2.0f) * len_max;
    const int64_t planes = x_hi - x_lo + 1;
    TNB_CUDA(sw->dist.reserve((size_t)(planes * M * M)));
    TNB_CUDA(sw->max_grad.reserve((size_t)sw->n_chunks));
    TNB_CUDA(cudaMemsetAsync(sw->max_grad.p, 0, (size_t)sw->n_chunks * sizeof(unsigned), s));
    float *dist0 = sw->dist.p - (int64_t)x_lo * M * M;  // indexed by the global vertex number
    int64_t slots = 0, rows = 0, words = 0, total_vertices = 0;
    int chunk = 0, total_ctas = 0;
    for (int a = 0; a < nc; ++a)
        for (int b = 0; b < nc; ++b)
            for (int cc = 0; cc < nc; ++cc, ++chunk) {
                int st[3] = {starts[a], starts[b], starts[cc]};
                int nn[3];
                for (int d = 0; d < 3; ++d) nn[d] = std::min(M, st[d] + unit) - st[d];
                // the part of the chunk inside the slab (a chunk piece emits every edge whose
                // two ends lie in the slab, in the chunk's own order)
                const int lo = std::max(st[0], x_lo), hi = std::min(st[0] + nn[0] - 1, x_hi);
                if (lo > hi) continue;
                st[0] = lo;
                nn[0] = hi - lo + 1;
                const int64_t count = (int64_t)nn[0] * nn[1] * nn[2];
                {
                    SweepPiece pc;
                    for (int d = 0; d < 3; ++d) { pc.s[d] = st[d]; pc.n[d] = nn[d]; }
                    pc.chunk = chunk;
                    pc.cta_first = total_ctas;
                    // four trips of two vertices per thread (one vertex per trip for other network shapes)
                    pc.ctas = (int)std::max<int64_t>(1, (count + kThreads * 8 - 1) / (kThreads * 8));
                    pc.ls = lattice_stride((int64_t)pc.ctas * kThreads, nn[0], nn[1]);
                    total_ctas += pc.ctas;
                    total_vertices += count;
                    sw->pieces.push_back(pc);
                }
                const int w32 = (nn[2] + 31) / 32;
                for (int axis = 0; axis < 3; ++axis) {
                    int dims[3] = {nn[0], nn[1], nn[2]};
                    dims[axis] -= 1;
                    const int64_t cnt = (int64_t)dims[0] * dims[1] * dims[2];
                    if (cnt <= 0) continue;
                    ChunkSeg sg;
                    sg.first = slots;
                    for (int d = 0; d < 3; ++d) { sg.s[d] = st[d]; sg.n[d] = nn[d]; }
                    sg.axis = axis;
                    sg.chunk = chunk;
                    sg.row_first = rows;
                    sg.bit_word = words;
                    sg.w32 = w32;
                    sw->segs.push_back(sg);
                    slots += cnt;
                    rows += (int64_t)dims[0] * dims[1];
                }
                words += (int64_t)nn[0] * nn[1] * w32;
            }
    sw->slots = slots;
    sw->rows = rows;
    sw->bit_words = words;
    if (launch && !sw->pieces.empty()) {
        TNB_CUDA(sw->d_pieces.reserve(sw->pieces.size()));
        TNB_CUDA(cudaMemcpyAsync(sw->d_pieces.p, sw->pieces.data(), sw->pieces.size() * sizeof(SweepPiece), cudaMemcpyHostToDevice, s));
        prof_begin(TNB_PROF_SWEEP, s);
        // registers vs warps per SM, measured (201^3 grid): 3 CTAs/SM = 168 registers, no spills: 0.80 ms; 4 = 128 registers,
        // 200 B of spills: 0.85 ms; 5 = 96 registers, 304 B: 0.87 ms
        static const int minb = std::getenv("TNB_SWEEP_MINB") ? std::atoi(std::getenv("TNB_SWEEP_MINB")) : 3;
        if (net->fixed_cfg && minb == 4)
            k_sweep_pieces<CfgRef, 4><<<(unsigned)total_ctas, kThreads, 0, s>>>(net->meta, M, sw->d_pieces.p, (int)sw->pieces.size(), dist0, sw->max_grad.p, g_sweep_sync);
        else if (net->fixed_cfg && minb == 5)
            k_sweep_pieces<CfgRef, 5><<<(unsigned)total_ctas, kThreads, 0, s>>>(net->meta, M, sw->d_pieces.p, (int)sw->pieces.size(), dist0, sw->max_grad.p, g_sweep_sync);
        else if (net->fixed_cfg)
            k_sweep_pieces<CfgRef, 3><<<(unsigned)total_ctas, kThreads, 0, s>>>(net->meta, M, sw->d_pieces.p, (int)sw->pieces.size(), dist0, sw->max_grad.p, g_sweep_sync);
        else
            k_sweep_pieces<CfgAny><<<(unsigned)total_ctas, kThreads, 0, s>>>(net->meta, M, sw->d_pieces.p, (int)sw->pieces.size(), dist0, sw->max_grad.p, 0);
        TNB_LAUNCH_CHECK();
        prof_end(TNB_PROF_SWEEP, s, total_vertices, total_vertices * 4 + (int64_t)net->table.cap * 8);
    }
    return TNB_OK;
}

static int skeleton_finish_impl(const tnb_net *net, tnb_sweep *sw, tnb_complex **out, cudaStream_t s)
{
    const int M = sw->M;
    const std::vector<ChunkSeg> &segs = sw->segs;
    const int64_t slots = sw->slots;
    tnb_complex *c = new tnb_complex();
    *out = c;
    c->halo.x_lo = sw->x_lo; c->halo.x_hi = sw->x_hi;
    c->halo.tag_lower = sw->tag_lower; c->halo.tag_upper = sw->tag_upper;
    if (segs.empty() || slots == 0) {  // degenerate grid: no edges at all
        int rc = complex_alloc(c, net, 64, 64);
        return rc;
    }
    DevBuf<ChunkSeg> d_segs;
    DevBuf<int> used, remap, block_sums, total;
    const int64_t base = (int64_t)sw->x_lo * M * M;              // first global vertex number of the slab
    const int64_t MS = (int64_t)(sw->x_hi - sw->x_lo + 1) * M * M;  // grid vertices of the slab
    float *dist0 = sw->dist.p - base;
    TNB_CUDA(d_segs.reserve(segs.size()));
    TNB_CUDA(cudaMemcpyAsync(d_segs.p, segs.data(), segs.size() * sizeof(ChunkSeg), cudaMemcpyHostToDevice, s));
    TNB_CUDA(used.reserve((size_t)MS));
    TNB_CUDA(remap.reserve((size_t)MS));
    TNB_CUDA(block_sums.reserve(kScanMaxBlocks));
    TNB_CUDA(total.reserve(2));
    TNB_CUDA(cudaMemsetAsync(used.p, 0, (size_t)MS * sizeof(int), s));

    // pass 1: count surviving edges so the complex can be sized (row by row, see k_skel_bits)
    if (sw->unit > 32 * kMaxRowWords) { set_error("tnb_skeleton: unit above " + std::to_string(32 * kMaxRowWords)); return TNB_ERR_UNSUPPORTED; }
    {
        DevBuf<uint32_t> bits;
        TNB_CUDA(bits.reserve((size_t)sw->bit_words + 1));
        k_skel_bits<<<dim3(kSMs * 2, (unsigned)segs.size()), 256, 0, s>>>(d_segs.p, (int)segs.size(), M, dist0, sw->max_grad.p, sw->k_len, bits.p);
        TNB_LAUNCH_CHECK();
        const SkelRows rows_q{d_segs.p, (int)segs.size(), M, bits.p};
        const int64_t n_rows = sw->rows;
        int64_t blocks = std::min<int64_t>((n_rows + kScanThreads - 1) / kScanThreads, kScanMaxBlocks);
        k_scan_count<<<(unsigned)blocks, kScanThreads, 0, s>>>(n_rows, nullptr, SkelRowCount{rows_q}, block_sums.p);
        TNB_LAUNCH_CHECK();
        std::vector<int> h(blocks);
        TNB_CUDA(cudaMemcpyAsync(h.data(), block_sums.p, blocks * sizeof(int), cudaMemcpyDeviceToHost, s));
        TNB_CUDA(cudaStreamSynchronize(s));
        int64_t E = 0;
        for (int v : h) E += v;
        if (E == 0) return complex_alloc(c, net, 64, 64);
        // vertices are bounded by 2E; real sizing happens after the vertex pass
        DevBuf<int2> raw;
        TNB_CUDA(raw.reserve((size_t)E));
        k_scan_write<<<(unsigned)blocks, kScanThreads, 0, s>>>(n_rows, nullptr, SkelRowCount{rows_q}, SkelRowEmit{rows_q, raw.p, used.p - base}, block_sums.p, total.p);
        TNB_LAUNCH_CHECK();
        // vertex pass: count, size, then place (one bit per grid vertex: ~90 % of them are on no kept edge, and the
        // generic write pass spent its time in the CTA barriers of their tiles)
        FlagCount fc{used.p};
        int64_t vblocks = std::min<int64_t>((MS + kScanThreads - 1) / kScanThreads, kScanMaxBlocks);
        DevBuf<uint32_t> vmask;
        TNB_CUDA(vmask.reserve((size_t)((MS + 31) / 32 + kScanMaxBlocks + 64)));
        k_scan_count_mask<<<(unsigned)vblocks, kScanThreads, 0, s>>>(MS, fc, block_sums.p, vmask.p);
        TNB_LAUNCH_CHECK();
        std::vector<int> hv(vblocks);
        TNB_CUDA(cudaMemcpyAsync(hv.data(), block_sums.p, vblocks * sizeof(int), cudaMemcpyDeviceToHost, s));
        TNB_CUDA(cudaStreamSynchronize(s));
        int64_t V = 0;
        for (int v : hv) V += v;
        // the additive head-room scales with the factor too: a retry with twice the factor doubles a tiny complex's room as well
        size_t Vcap = (size_t)(std::max<int64_t>(V, 4096) * capacity_factor()), Ecap = (size_t)(std::max<int64_t>(E, 4096) * capacity_factor());
        int rc = complex_alloc(c, net, Vcap, Ecap);
        if (rc) return rc;
        SkelVertEmit vemit{nullptr, net->meta.marks, net->meta.pre_2s, net->meta.pre_scale, M, remap.p, c->cvert(),
                           c->tag[c->vcur].p, sw->x_lo, sw->tag_lower ? sw->x_lo : -1, sw->tag_upper ? sw->x_hi : -1};
        k_scan_write_mask<<<(unsigned)vblocks, kScanThreads, 0, s>>>(MS, vmask.p, vemit, block_sums.p, total.p + 1);
        TNB_LAUNCH_CHECK();
        TNB_CUDA(cudaMemcpyAsync(c->cedges(), raw.p, (size_t)E * sizeof(int2), cudaMemcpyDeviceToDevice, s));
        k_remap_edges<<<grid_for(E, 256), 256, 0, s>>>(c->cedges(), E, remap.p - base);
        TNB_LAUNCH_CHECK();
        c->V = V;
        c->E = E;
        k_set_counts<<<1, 1, 0, s>>>(c->counters.p, (int)V, (int)E, c->vcur, c->ecur, c->acur);
        TNB_LAUNCH_CHECK();
        rc = eval_vertices(net, c, 0, V, s);
        if (rc) return rc;
        if ((rc = initial_cross_mask(c, s))) return rc;
    }
    return TNB_OK;
}

static int skeleton_impl(const tnb_net *net, int unit, tnb_complex **out, cudaStream_t s)
{
    tnb_sweep sw;
    int rc = sweep_impl(net, unit, 0, net->meta.n_marks - 1, false, false, &sw, s);
    if (rc) return rc;
    return skeleton_finish_impl(net, &sw, out, s);
}

// ================================================================================================
// one hyperplane (subpoly_)
// ================================================================================================
// All sizes a step produces (splits, hits, pairs, post-prune V/E) stay in device memory:
//   cnt[C_SPLIT..]  transient per step,  cnt[C_V], cnt[C_E] the current complex size.
// Kernels read them there; the host only supplies upper bounds to size grids, and syncs
// ONCE per step (after the connecting-edge count, where it must size the edge array).
struct SplitCount {
    const int2 *edges;
    const float *out;
    int R, idx;
    float eps;
    __device__ __forceinline__ int operator()(int64_t e) const
    {
        const int2 ed = edges[e];
        const float d0 = out[(int64_t)ed.x * R + idx], d1 = out[(int64_t)ed.y * R + idx];
        return ((d0 * d1) < 0.0f && fabsf(d0) > eps && fabsf(d1) > eps) ? 1 : 0;  // subpoly.py:104-105
    }
};
struct ListEmit {
    int *list;
    __device__ __forceinline__ void operator()(int64_t i, int pos, int) const { list[pos] = (int)i; }
};

// new vertex of every split edge: position (subpoly.py:113-117, :180), network row, the
// failover mask of subpoly_debug.py:37-43, edge rewiring (subpoly.py:210-215)
// k-th split edge of the step -> vertex V + k; returns whether the failover override fires
// kPack: also bit-pack the new vertex's region indicator right away, as if the failover override will not
// fire (it almost never does; when it does, k_finalize_new redoes the packing from the overridden row)
template <class C, bool kPack = false>
__device__ __forceinline__ int new_vertex_item(const NetMeta &n, int idx, float eps, int k, int V, int E, const int *split_list,
                                               int2 *edges, float *vert, float *out, uint64_t *sig, uint64_t *bmask,
                                               unsigned char *tag, float *row_dst = nullptr)
{
    const int R = n.R;
    int any = 0;
    {
        const int e = split_list[k];
        const int2 ed = edges[e];
        const float d0 = __fdiv_rn(out[(int64_t)ed.x * R + idx], eps), d1 = __fdiv_rn(out[(int64_t)ed.y * R + idx], eps);
        const float w = __fdiv_rn(fabsf(d0), fabsf(d1 - d0));
        const float omw = 1.0f - w;
        float x[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) x[d] = vert[3 * (int64_t)ed.x + d] * omw + vert[3 * (int64_t)ed.y + d] * w;
        const int64_t nv = (int64_t)V + k;
#pragma unroll
        for (int d = 0; d < 3; ++d) vert[3 * nv + d] = x[d];
        float *row = row_dst ? row_dst : out + nv * R;  // row_dst: the caller stores the warp's rows together
        const uint64_t za = ~(sig[3 * (int64_t)ed.x] | sig[3 * (int64_t)ed.x + 1]);
        const uint64_t zb = ~(sig[3 * (int64_t)ed.y] | sig[3 * (int64_t)ed.y + 1]);
        const uint64_t bm = (za & zb & ((1ull << idx) - 1ull)) | (1ull << idx);
        bmask[k] = bm;
        if constexpr (kPack) {
            uint64_t pos, neg, big;  // big: |value| > eps (the step's eps), for the failover test below
            float xp[3];
            outputs_row_packed<C>(n, x, row, n.eps, eps, xp, pos, neg, big);
            if (big & bm) any = 1;
            sig[3 * nv] = pos;
            sig[3 * nv + 1] = neg;
            sig[3 * nv + 2] = pack_grid(n, n.marks, xp, n.eps);
        } else {
            outputs_row<C>(n, x, row);
            for (uint64_t m = bm; m; m &= m - 1) {
                const int col = __ffsll((long long)m) - 1;
                if (fabsf(row[col]) > eps) any = 1;
            }
        }
        tag[nv] = tag[ed.x] & tag[ed.y];         // on a shared slab plane iff both parents are
        edges[e].y = (int)nv;                    // left part keeps the first endpoint
        edges[E + k] = make_int2(ed.y, (int)nv); // right part
    }
    return any;
}

template <class C, bool kPack = false>
__device__ __forceinline__ void body_new_vertices(const NetMeta &n, int idx, float eps,
                                                           int Vcap, int Ecap, const int *split_list,
                                                           int2 *edges, float *vert,
                                                           float *out, uint64_t *sig,
                                                           uint64_t *bmask, int *cnt, unsigned char *tag, float *tile = nullptr)
{
    const int S = cnt[C_RAW], V = cnt[C_V], E = cnt[C_E];
    if ((int64_t)V + S > Vcap || (int64_t)E + S > Ecap) {  // host grows the arrays and re-runs
        if (blockIdx.x == 0 && threadIdx.x == 0) cnt[C_OVERFLOW] = 1;
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt[C_SPLIT] = S;  // planar path: every crossed edge is split
    int any = 0;
    if (tile) {  // [blockDim.x][R] floats of shared memory: the rows of a warp leave together (warp_store_rows)
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, R = n.R;
        float *wtile = tile + (size_t)warp * 32 * R;
        for (int k0 = blockIdx.x * blockDim.x + warp * 32; k0 < S; k0 += gridDim.x * blockDim.x) {  // warp uniform
            const int k = k0 + lane;
            if (k < S) any |= new_vertex_item<C, kPack>(n, idx, eps, k, V, E, split_list, edges, vert, out, sig, bmask, tag, wtile + lane * R);
            __syncwarp();
            warp_store_rows(out, (int64_t)V + k0, min(32, S - k0), R, wtile);
            __syncwarp();
        }
    } else {
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < S; k += gridDim.x * blockDim.x)
            any |= new_vertex_item<C, kPack>(n, idx, eps, k, V, E, split_list, edges, vert, out, sig, bmask, tag);
    }
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) atomicOr(cnt + C_FLAG, 1);
}

// (sig is read at the old vertices and written at the new ones: no __restrict__)
template <class C>
__global__ void __launch_bounds__(kThreads) k_new_vertices(const __grid_constant__ NetMeta n, int idx, float eps,
                                                           int Vcap, int Ecap, const int *__restrict__ split_list,
                                                           int2 *__restrict__ edges, float *__restrict__ vert,
                                                           float *__restrict__ out, uint64_t *sig,
                                                           uint64_t *__restrict__ bmask, int *__restrict__ cnt,
                                                           unsigned char *__restrict__ tag)
{
    extern __shared__ float s_tile[];  // [kThreads][R]
    body_new_vertices<C, true>(n, idx, eps, Vcap, Ecap, split_list, edges, vert, out, sig, bmask, cnt, tag, s_tile);
}

// apply the failover override when any new vertex violated it, then bit-pack the region
// indicator of the new vertices
__device__ __forceinline__ void finalize_item(const NetMeta &n, const float *marks, const float *vert, float *out, uint64_t *sig,
                                              const uint64_t *bmask, int flag, int V, int k)
{
    const int R = n.R;
    {
        const int64_t v = (int64_t)V + k;
        float *row = out + v * R;
        if (flag)
            for (uint64_t m = bmask[k]; m; m &= m - 1) row[__ffsll((long long)m) - 1] = 0.0f;
        float x[3] = {vert[3 * v], vert[3 * v + 1], vert[3 * v + 2]}, xp[3];
        preprocess(n, x, xp);
        uint64_t pos, neg;
        pack_signs(row, R, n.eps, pos, neg);
        sig[3 * v] = pos;
        sig[3 * v + 1] = neg;
        sig[3 * v + 2] = pack_grid(n, marks, xp, n.eps);
    }
}
__device__ __forceinline__ void body_finalize_new(const NetMeta &n,
                                                           const float *vert, float *out,
                                                           uint64_t *sig, const uint64_t *bmask,
                                                           const int *cnt)
{
    if (cnt[C_OVERFLOW]) return;
    const int flag = cnt[C_FLAG], S = cnt[C_SPLIT], V = cnt[C_V];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < S; k += gridDim.x * blockDim.x) finalize_item(n, n.marks, vert, out, sig, bmask, flag, V, k);
}

// multi-launch path: k_new_vertices packed the region indicators already; only a step whose failover
// override fired (subpoly_debug.py:41-49) has rows to zero and indicators to pack again
__global__ void __launch_bounds__(kThreads) k_finalize_new(const __grid_constant__ NetMeta n,
                                                           const float *__restrict__ vert, float *__restrict__ out,
                                                           uint64_t *__restrict__ sig, const uint64_t *__restrict__ bmask,
                                                           const int *__restrict__ cnt)
{
    if (!cnt[C_FLAG]) return;
    body_finalize_new(n, vert, out, sig, bmask, cnt);
}


// ---- curve-approximation path (force=False, subpoly.py:120-183 + strict_check) ----------------------
// Pass 1: candidate new vertex of every crossed edge into TEMP slots (the idle half of the
// ping-pong vertex arrays), no rewiring yet.  sflag bit0 = edge is not axis aligned (c),
// bit1 = no admissible intersection (gg).
// Warp-cooperative (curve.cuh): every lane of the warp calls it, `active` says whether the lane has
// a crossed edge k of its own.
template <class C>
__device__ __forceinline__ int curve_candidate_item(const NetMeta &n, int idx, float eps, bool active, int k, const int *split_list,
                                                    const int2 *edges, const float *vert, const float *out,
                                                    const uint64_t *sig, float *tvert, float *tout, uint64_t *bmask,
                                                    int *sflag, int *cnt, int *gd)
{
    const int R = n.R;
    float e0[3] = {0.0f, 0.0f, 0.0f}, e1[3] = {0.0f, 0.0f, 0.0f}, x[3] = {0.0f, 0.0f, 0.0f};
    uint64_t common = 0;
    int flags = 0, plane = 0;
    bool curved = false;
    if (active) {
        const int e = split_list[k];
        const int2 ed = edges[e];
#pragma unroll
        for (int d = 0; d < 3; ++d) { e0[d] = vert[3 * (int64_t)ed.x + d]; e1[d] = vert[3 * (int64_t)ed.y + d]; }
        const float d0 = __fdiv_rn(out[(int64_t)ed.x * R + idx], eps), d1 = __fdiv_rn(out[(int64_t)ed.y * R + idx], eps);
        const float w = __fdiv_rn(fabsf(d0), fabsf(d1 - d0));
        const float omw = 1.0f - w;
#pragma unroll
        for (int d = 0; d < 3; ++d) x[d] = e0[d] * omw + e1[d] * w;
        const uint64_t za = ~(sig[3 * (int64_t)ed.x] | sig[3 * (int64_t)ed.x + 1]);
        const uint64_t zb = ~(sig[3 * (int64_t)ed.y] | sig[3 * (int64_t)ed.y + 1]);
        common = za & zb & ((1ull << idx) - 1ull);
        int moved = 0;
#pragma unroll
        for (int d = 0; d < 3; ++d) moved += fabsf(e1[d] - e0[d]) > eps ? 1 : 0;
        if (moved > 1) {  // bi-/tri-linear edge (subpoly.py:122)
            flags = 1;
            if (!common) atomicOr(cnt + C_ERR, kErrNoPlane);  // the reference exits here (subpoly.py:140-148)
            else { curved = true; plane = 63 - __clzll((long long)common); }  // nonzero_last
        }
    }
    float p8[8], q8[8], ints[3] = {0.0f, 0.0f, 0.0f};
    warp_group8_columns<C>(n, curved, e0, e1, n.eps, plane, idx, p8, q8);
    warp_curve_intersection(curved, p8, q8, ints);
    int any = 0;
    bool walks = false;
    if (active) {
        if (curved) {
            bool gg = false;
#pragma unroll
            for (int d = 0; d < 3; ++d) gg = gg || ints[d] < 0.0f || ints[d] > 1.0f;
            if (gg) {
                flags |= 2;
            } else {
                float xg[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) xg[d] = e0[d] * (1.0f - ints[d]) + e1[d] * ints[d];
                float *row = tout + (int64_t)k * R;
                outputs_row<C>(n, xg, row);
                if (fabsf(row[plane]) > eps || fabsf(row[idx]) > eps) {
                    // off one of its two planes: a walk of the gradient-descent repair (subpoly_debug.py:121-165, repair.cuh),
                    // which writes this candidate's vertex and row once the shared step count is known
                    const int r = atomicAdd(gd + GD_COUNT, 1);
                    if (r < kGdCap) {
                        int *w = gd + kGdHead + 5 * r;
                        w[0] = k; w[1] = plane;
                        w[2] = __float_as_int(ints[0]); w[3] = __float_as_int(ints[1]); w[4] = __float_as_int(ints[2]);
                        atomicOr(cnt + C_ERR, kErrRepair);
                    } else atomicOr(cnt + C_ERR, kErrGradientDescent);
                    walks = true;
                }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) x[d] = e0[d] + ints[d] * (e1[d] - e0[d]);  // subpoly.py:183
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) tvert[3 * (int64_t)k + d] = x[d];
        float *row = tout + (int64_t)k * R;
        outputs_row<C>(n, x, row);
        const uint64_t bm = common | (1ull << idx);
        bmask[k] = bm;
        sflag[k] = flags;
        for (uint64_t m = bm; m; m &= m - 1) {
            const int col = __ffsll((long long)m) - 1;
            if (fabsf(row[col]) > eps) any = 1;
        }
        if (walks) any = 0;  // decided by the repaired vertex (gd_complex_walk)
    }
    return any;
}

// The repair's two passes over the walks curve_candidate_item filed (repair.cuh); tid / nthreads = the caller's
// place among the threads that share the walks.  The second pass leaves the repaired vertex and its network row in
// the candidate's temp slot and returns the failover flag of subpoly_debug.py:33-51 for them; if the shared loop
// ended with a walk still off its planes the reference ends the extraction (subpoly.py:172-174): kErrGradientDescent.
template <class C>
static __device__ void gd_complex_note(const NetMeta &n, int idx, float eps, const int *split_list, const int2 *edges,
                                       const float *vert, int *gd, int tid, int nthreads)
{
    const int G = min(gd[GD_COUNT], kGdCap);
    for (int r = tid; r < G; r += nthreads) {
        const int *w = gd + kGdHead + 5 * r;
        const int2 ed = edges[split_list[w[0]]];
        float e0[3], e1[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) { e0[d] = vert[3 * (int64_t)ed.x + d]; e1[d] = vert[3 * (int64_t)ed.y + d]; }
        const float x0[3] = {__int_as_float(w[2]), __int_as_float(w[3]), __int_as_float(w[4])};
        gd_note_steps<C>(n, e0, e1, w[1], idx, eps, x0, gd + GD_MASK);
    }
}
template <class C>
static __device__ int gd_complex_walk(const NetMeta &n, int idx, float eps, const int *split_list, const int2 *edges,
                                      const float *vert, float *tvert, float *tout, const uint64_t *bmask, int *cnt, int *gd,
                                      int tid, int nthreads)
{
    bool ok;
    const int bodies = gd_bodies(gd + GD_MASK, ok);
    if (!ok) {
        if (tid == 0) atomicOr(cnt + C_ERR, kErrGradientDescent);
        return 0;
    }
    const int G = min(gd[GD_COUNT], kGdCap), R = n.R;
    int any = 0;
    for (int r = tid; r < G; r += nthreads) {
        const int *w = gd + kGdHead + 5 * r;
        const int k = w[0];
        const int2 ed = edges[split_list[k]];
        float e0[3], e1[3], d[2];
#pragma unroll
        for (int dd = 0; dd < 3; ++dd) { e0[dd] = vert[3 * (int64_t)ed.x + dd]; e1[dd] = vert[3 * (int64_t)ed.y + dd]; }
        float x[3] = {__int_as_float(w[2]), __int_as_float(w[3]), __int_as_float(w[4])};
        gd_walk<C>(n, e0, e1, w[1], idx, bodies, x, d);
        float xv[3];
#pragma unroll
        for (int dd = 0; dd < 3; ++dd) xv[dd] = e0[dd] + x[dd] * (e1[dd] - e0[dd]);  // subpoly.py:183
#pragma unroll
        for (int dd = 0; dd < 3; ++dd) tvert[3 * (int64_t)k + dd] = xv[dd];
        float *row = tout + (int64_t)k * R;
        outputs_row<C>(n, xv, row);
        for (uint64_t m = bmask[k]; m; m &= m - 1) {
            const int col = __ffsll((long long)m) - 1;
            if (fabsf(row[col]) > eps) any = 1;
        }
    }
    return any;
}
template <class C>
__global__ void __launch_bounds__(64) k_gd_note(const __grid_constant__ NetMeta n, int idx, float eps, const int *__restrict__ split_list,
                                                const int2 *__restrict__ edges, const float *__restrict__ vert, const int *__restrict__ cnt, int *__restrict__ gd)
{
    if (!(cnt[C_ERR] & kErrRepair)) return;
    gd_complex_note<C>(n, idx, eps, split_list, edges, vert, gd, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
template <class C>
__global__ void __launch_bounds__(64) k_gd_walk(const __grid_constant__ NetMeta n, int idx, float eps, const int *__restrict__ split_list,
                                                const int2 *__restrict__ edges, const float *__restrict__ vert, float *__restrict__ tvert,
                                                float *__restrict__ tout, const uint64_t *__restrict__ bmask, int *__restrict__ cnt, int *__restrict__ gd)
{
    if (!(cnt[C_ERR] & kErrRepair)) return;
    if (gd_complex_walk<C>(n, idx, eps, split_list, edges, vert, tvert, tout, bmask, cnt, gd, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x))
        atomicOr(cnt + C_FLAG, 1);
}

template <class C>
__global__ void __launch_bounds__(kThreads) k_new_vertices_curve(const __grid_constant__ NetMeta n, int idx, float eps,
                                                                 int Vcap, int Ecap, const int *__restrict__ split_list,
                                                                 const int2 *__restrict__ edges,
                                                                 const float *__restrict__ vert, const float *__restrict__ out,
                                                                 const uint64_t *__restrict__ sig, float *__restrict__ tvert,
                                                                 float *__restrict__ tout, uint64_t *__restrict__ bmask,
                                                                 int *__restrict__ sflag, int *__restrict__ cnt, int *__restrict__ gd)
{
    const int S = cnt[C_RAW], V = cnt[C_V], E = cnt[C_E];
    if ((int64_t)V + S > Vcap || (int64_t)E + S > Ecap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) cnt[C_OVERFLOW] = 1;
        return;
    }
    int any = 0;
    const int lane = threadIdx.x & 31;
    for (int k0 = blockIdx.x * blockDim.x + (threadIdx.x - lane); k0 < S; k0 += gridDim.x * blockDim.x)  // warp-uniform trip count
        any |= curve_candidate_item<C>(n, idx, eps, k0 + lane < S, k0 + lane, split_list, edges, vert, out, sig, tvert, tout, bmask, sflag, cnt, gd);
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) atomicOr(cnt + C_FLAG, 1);
}

// Pass 2: failover override on the temp rows, then strict_check's keep decision
// (subpoly_debug.py:234-271): on the plane within eps, and not a curved edge without a root.
__device__ __forceinline__ void strict_keep_item(int R, int idx, float eps, float *tout, const uint64_t *bmask, int *sflag, int flag, int k)
{
    float *row = tout + (int64_t)k * R;
    if (flag)
        for (uint64_t m = bmask[k]; m; m &= m - 1) row[__ffsll((long long)m) - 1] = 0.0f;
    const int f = sflag[k];
    const bool keep = fabsf(row[idx]) < eps && !((f & 1) && (f & 2));
    sflag[k] = keep ? 1 : 0;
}
__global__ void __launch_bounds__(kThreads) k_strict_keep(int R, int idx, float eps, float *__restrict__ tout,
                                                          const uint64_t *__restrict__ bmask, int *__restrict__ sflag,
                                                          const int *__restrict__ cnt)
{
    if (cnt[C_OVERFLOW]) return;
    const int flag = cnt[C_FLAG], S = cnt[C_RAW];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < S; k += gridDim.x * blockDim.x) strict_keep_item(R, idx, eps, tout, bmask, sflag, flag, k);
}

struct KeepFlagCount {
    const int *flag;
    const int *overflow;
    __device__ __forceinline__ int operator()(int64_t k) const { return (*overflow == 0 && flag[k]) ? 1 : 0; }
};
// Pass 3: surviving new vertices move to their final slots, edges are rewired (subpoly.py:210-215)
struct CurveCommitEmit {
    NetMeta const *meta;  // device copy not needed: fields below
    const int *split_list;
    int2 *edges;
    const float *tvert, *tout;
    float *vert, *out;
    uint64_t *sig;
    const float *marks;
    const int *cnt;
    int R, n_marks;
    float eps, pre_scale, pre_2s, pre_inv;
    int pre_pow2;
    unsigned char *tag;
    __device__ __forceinline__ void operator()(int64_t k, int rank, int) const
    {
        const int V = cnt[C_V], E = cnt[C_E];
        const int64_t nv = (int64_t)V + rank;
        float x[3], xp[3];
        for (int d = 0; d < 3; ++d) {
            x[d] = tvert[3 * k + d];
            vert[3 * nv + d] = x[d];
            const float t = x[d] + pre_scale;
            xp[d] = pre_pow2 ? t * pre_inv : __fdiv_rn(t, pre_2s);
        }
        float *row = out + nv * R;
        uint64_t pos = 0, neg = 0;
        for (int c0 = 0; c0 < R; c0 += 16) {  // 16 loads in flight (see pack_signs)
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = c0 + i < R ? tout[k * R + c0 + i] : 0.0f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (c0 + i < R) {
                    row[c0 + i] = v[i];
                    if (!(fabsf(v[i]) <= eps)) { if (v[i] > 0.0f) pos |= 1ull << (c0 + i); else neg |= 1ull << (c0 + i); }
                }
            }
        }
        uint64_t g = 0;
        for (int d = 0; d < 3; ++d) {
            int off = lower_bound(marks, n_marks, xp[d] + eps) - 1;
            const float mk = marks[off < 0 ? off + n_marks : off];
            g |= (uint64_t)(uint32_t)(off + 1) << (20 * d);
            g |= (fabsf(mk - xp[d]) > eps ? 1ull : 0ull) << (60 + d);
        }
        sig[3 * nv] = pos;
        sig[3 * nv + 1] = neg;
        sig[3 * nv + 2] = g;
        const int e = split_list[k];
        const int old = edges[e].y;
        tag[nv] = tag[edges[e].x] & tag[old];
        edges[e].y = (int)nv;
        edges[E + rank] = make_int2(old, (int)nv);
    }
};

struct HitCount {
    const float *out;
    const int *alive;  // rows of pruned vertices stay in place (complex.cuh): they are not part of the complex
    int R, idx;
    float eps;
    __device__ __forceinline__ int operator()(int64_t v) const
    {
        const int al = alive[v];
        const float o = out[v * R + idx];  // both loads in flight (a dead vertex's row is still there)
        return (al && fabsf(o) < eps) ? 1 : 0;  // subpoly.py:233
    }
};

// candidates = hit old vertices (already in cand[0..H)), then the new ones; publishes the
// candidate count
__device__ __forceinline__ void body_fill_new_cands(int *cand, int *cnt)
{
    const int H = cnt[C_HIT], S = cnt[C_SPLIT], V = cnt[C_V];
    if (cnt[C_OVERFLOW]) return;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < S; k += gridDim.x * blockDim.x) cand[H + k] = V + k;
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt[C_CAND] = cnt[C_RAW] ? H + S : 0;  // subpoly.py:110 tests the crossed edges
}

__global__ void k_fill_new_cands(int *__restrict__ cand, int *__restrict__ cnt)
{
    body_fill_new_cands(cand, cnt);
}

// ---- cell buckets (linked-list form: the persistent step kernels of small complexes) ---------------
// (CellBox / cell_box / cell_id: cells.cuh)
__device__ __forceinline__ void bucket_insert_item(int c, int v, const uint64_t *sig, unsigned long long *head,
                                                   tnb_bucket_rec *next, int dim, uint32_t stamp)
{
    tnb_bucket_rec r;
    r.v = v;
    r.pos = sig[3 * (int64_t)v];
    r.neg = sig[3 * (int64_t)v + 1];
    r.grd = sig[3 * (int64_t)v + 2];
    const CellBox b = cell_box(r.grd);
    // a vertex lies in 1, 2, 4 or 8 cells (one or two per axis).  All exchanges are issued before the
    // first record is written: their round trips overlap instead of adding up.  Record c*8 + slot,
    // slot = the cell's (dx,dy,dz) bits; which slot a cell gets does not matter.
    unsigned long long old[8];
#pragma unroll
    for (int slot = 0; slot < 8; ++slot) {
        const int cx = b.lo[0] + (slot >> 2), cy = b.lo[1] + ((slot >> 1) & 1), cz = b.lo[2] + (slot & 1);
        const bool in = cx <= b.hi[0] && cy <= b.hi[1] && cz <= b.hi[2];
        const unsigned long long mine = ((unsigned long long)stamp << 32) | (unsigned)(c * 8 + slot);
        old[slot] = in ? atomicExch(head + cell_id(cx, cy, cz, dim), mine) : 0ull;
    }
#pragma unroll
    for (int slot = 0; slot < 8; ++slot) {
        const int cx = b.lo[0] + (slot >> 2), cy = b.lo[1] + ((slot >> 1) & 1), cz = b.lo[2] + (slot & 1);
        if (cx <= b.hi[0] && cy <= b.hi[1] && cz <= b.hi[2]) {
            r.next = ((uint32_t)(old[slot] >> 32) == stamp) ? (int)(uint32_t)old[slot] : -1;
            next[c * 8 + slot] = r;
        }
    }
}
__device__ __forceinline__ void body_bucket_insert(const int *cand, const int *cnt,
                                const uint64_t *sig, unsigned long long *head,
                                tnb_bucket_rec *next, int dim, uint32_t stamp)
{
    const int n_cand = cnt[C_CAND];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cand; c += gridDim.x * blockDim.x)
        bucket_insert_item(c, cand[c], sig, head, next, dim, stamp);
}

__global__ void __launch_bounds__(kThreads) k_bucket_insert(const int *__restrict__ cand, const int *__restrict__ cnt,
                                const uint64_t *__restrict__ sig, unsigned long long *__restrict__ head,
                                tnb_bucket_rec *__restrict__ next, int dim, uint32_t stamp)
{
    body_bucket_insert(cand, cnt, sig, head, next, dim, stamp);
}

constexpr int kLocalPartners = 128;  // partner lists up to this size are sorted in registers/local memory

// Partners of candidate a: candidates b with a larger vertex number that share an expanded
// region with a and at least one plane (subpoly.py:484-535).  Each pair is found in exactly
// one cell (the smallest common one).
struct PartnerQuery {
    int va;
    uint64_t pa, na, ga, za;
    CellBox ba;
};
__device__ __forceinline__ PartnerQuery partner_query(int va, const uint64_t *sig)
{
    PartnerQuery q;
    q.va = va;
    q.pa = sig[3 * (int64_t)va];
    q.na = sig[3 * (int64_t)va + 1];
    q.ga = sig[3 * (int64_t)va + 2];
    q.za = ~(q.pa | q.na);
    q.ba = cell_box(q.ga);
    return q;
}
// is the candidate of record r a partner of q, found in cell (cx,cy,cz)?  (each pair is accepted in
// exactly one cell: the smallest common one)
__device__ __forceinline__ bool partner_test(const PartnerQuery &q, const tnb_bucket_rec &r, int cx, int cy, int cz, uint64_t colmask)
{
    if (r.v <= q.va) return false;
    const uint64_t pb = r.pos, nb = r.neg, gb = r.grd;
    if (((q.pa & nb) | (q.na & pb)) & colmask) return false;  // opposite signs: no common region
    const CellBox bb = cell_box(gb);
    const int cur[3] = {cx, cy, cz};
    bool ok = true;
    int shared = __popcll(q.za & ~(pb | nb) & colmask);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int lo = max(q.ba.lo[d], bb.lo[d]), hi = min(q.ba.hi[d], bb.hi[d]);
        if (lo > hi || cur[d] != lo) ok = false;  // not a common cell / not the canonical one
        if (!grid_mask(q.ga, d) && !grid_mask(gb, d) && q.ba.hi[d] == bb.hi[d]) ++shared;
    }
    return ok && shared >= 1;
}
// walks the bucket of ONE cell; hit(vb) is called for every partner found there
template <class Hit>
__device__ __forceinline__ void walk_cell(const PartnerQuery &q, int cx, int cy, int cz, const unsigned long long *head,
                                          const tnb_bucket_rec *next, int dim, uint32_t stamp, uint64_t colmask, Hit hit)
{
    const unsigned long long h = head[cell_id(cx, cy, cz, dim)];
    if ((uint32_t)(h >> 32) != stamp) return;
    for (int rec = (int)(uint32_t)h; rec >= 0;) {
        const tnb_bucket_rec r = next[rec];
        rec = r.next;
        if (partner_test(q, r, cx, cy, cz, colmask)) hit(r.v);
    }
}
// One thread, all cells of the candidate's box.  Returns the count; when `list` is non-null the
// first `cap` partner vertex numbers are stored there (unsorted).
__device__ __forceinline__ int find_partners(int a, const int *cand, const uint64_t *sig,
                                             const unsigned long long *head, const tnb_bucket_rec *next,
                                             int dim, uint32_t stamp, uint64_t colmask, int *list, int cap, int stride)
{
    const PartnerQuery q = partner_query(cand[a], sig);
    int count = 0;
    for (int cx = q.ba.lo[0]; cx <= q.ba.hi[0]; ++cx)
        for (int cy = q.ba.lo[1]; cy <= q.ba.hi[1]; ++cy)
            for (int cz = q.ba.lo[2]; cz <= q.ba.hi[2]; ++cz)
                walk_cell(q, cx, cy, cz, head, next, dim, stamp, colmask, [&](int vb) {
                    if (list && count < cap) list[(int64_t)count * stride] = vb;
                    ++count;
                });
    return count;
}

// Eight lanes per candidate, one cell of its box each (a vertex on a grid plane / line / point
// lies in 2 / 4 / 8 cells): the dependent-load chains of the cells run side by side.  Handles the
// candidates [begin, end) with the calling CTA (blockDim.x threads, a multiple of 32); leaves the
// count in pcount[a] and the first kCachedPartners partners (unsorted) in pcache.
__device__ __forceinline__ void pair_count_groups(int begin, int end, const int *cand, const uint64_t *sig,
                                                  const unsigned long long *head, const tnb_bucket_rec *next, int dim,
                                                  uint32_t stamp, uint64_t colmask, int *pcount, int *pcache, int *s_cnt,
                                                  int *long_list = nullptr, int *long_cnt = nullptr)
{
    const int g = threadIdx.x & 7, grp = threadIdx.x >> 3, per_pass = blockDim.x >> 3;
    for (int base = begin; base < end; base += per_pass) {
        const int a = base + grp;
        if (g == 0) s_cnt[grp] = 0;
        __syncwarp();
        if (a < end) {
            const PartnerQuery q = partner_query(cand[a], sig);
            const int nx = q.ba.hi[0] - q.ba.lo[0] + 1, ny = q.ba.hi[1] - q.ba.lo[1] + 1, nz = q.ba.hi[2] - q.ba.lo[2] + 1;
            if (g < nx * ny * nz) {
                const int cz = q.ba.lo[2] + g % nz, cy = q.ba.lo[1] + (g / nz) % ny, cx = q.ba.lo[0] + g / (nz * ny);
                walk_cell(q, cx, cy, cz, head, next, dim, stamp, colmask, [&](int vb) {
                    const int pos = atomicAdd(s_cnt + grp, 1);
                    if (pos < kCachedPartners) pcache[(int64_t)a * kCachedPartners + pos] = vb;
                });
            }
        }
        __syncwarp();
        if (g == 0 && a < end) {
            pcount[a] = s_cnt[grp];
            if (long_list && s_cnt[grp] > kCachedPartners) long_list[atomicAdd(long_cnt, 1)] = a;  // any order: each list has its own range
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void body_pair_count(const int *cand, int *cnt,
                                                         const uint64_t *sig,
                                                         const unsigned long long *head,
                                                         const tnb_bucket_rec *next, int dim, uint32_t stamp,
                                                         uint64_t colmask, int *pcount, int *pcache, int *long_list)
{
    __shared__ int s_cnt[kThreads / 8];
    const int n_cand = cnt[C_CAND], per_block = blockDim.x >> 3;
    for (int base = blockIdx.x * per_block; base < n_cand; base += gridDim.x * per_block)
        pair_count_groups(base, min(base + per_block, n_cand), cand, sig, head, next, dim, stamp, colmask, pcount, pcache, s_cnt,
                          long_list, cnt + C_LONG);
}

__global__ void __launch_bounds__(kThreads) k_pair_count(const int *__restrict__ cand, int *__restrict__ cnt,
                                                         const uint64_t *__restrict__ sig,
                                                         const unsigned long long *__restrict__ head,
                                                         const tnb_bucket_rec *__restrict__ next, int dim, uint32_t stamp,
                                                         uint64_t colmask, int *__restrict__ pcount, int *__restrict__ pcache,
                                                         int *__restrict__ long_list)
{
    body_pair_count(cand, cnt, sig, head, next, dim, stamp, colmask, pcount, pcache, long_list);
}

struct ArrayCount {
    const int *v;
    __device__ __forceinline__ int operator()(int64_t i) const { return v[i]; }
};
struct OffsetEmit {
    int *off;
    __device__ __forceinline__ void operator()(int64_t i, int pos, int) const { off[i] = pos; }
};

__device__ __forceinline__ void pair_write_item(int a, const int *cand, const uint64_t *sig,
                                                const unsigned long long *head, const tnb_bucket_rec *next, int dim,
                                                uint32_t stamp, uint64_t colmask, const int *pcount, const int *poff,
                                                int2 *edges_out, const int *pcache, bool defer_long = false)
{
    {
        const int c = pcount[a];
        if (c == 0) return;
        if (defer_long && c > kCachedPartners) return;  // pair_write_long: a warp per long list
        const int va = cand[a];
        int2 *dst = edges_out + poff[a];
        // ascending partner number = the order unique(dim=0) leaves (subpoly.py:243-244)
        if (c <= kNetworkPartners) {  // the count pass left the whole list behind
            int list[kNetworkPartners];
#pragma unroll
            for (int i = 0; i < kNetworkPartners; ++i) list[i] = i < c ? pcache[(int64_t)a * kCachedPartners + i] : 0x7fffffff;
#pragma unroll
            for (int i = 1; i < kNetworkPartners; ++i)
#pragma unroll
                for (int j = kNetworkPartners - 1; j >= i; --j)
                    if (list[j - 1] > list[j]) { const int t = list[j]; list[j] = list[j - 1]; list[j - 1] = t; }
#pragma unroll
            for (int i = 0; i < kNetworkPartners; ++i)
                if (i < c) dst[i] = make_int2(va, list[i]);
        } else if (c <= kCachedPartners) {  // still no second walk: sort the cached list
            int list[kCachedPartners];
            for (int i = 0; i < c; ++i) list[i] = pcache[(int64_t)a * kCachedPartners + i];
            thread_sort(list, c);
            for (int i = 0; i < c; ++i) dst[i] = make_int2(va, list[i]);
        } else if (c <= kLocalPartners) {
            int list[kLocalPartners];
            find_partners(a, cand, sig, head, next, dim, stamp, colmask, list, kLocalPartners, 1);
            thread_sort(list, c);
            for (int i = 0; i < c; ++i) dst[i] = make_int2(va, list[i]);
        } else {  // long list (degenerate, very large region): gather and sort in place in HBM
            int *keys = &dst[0].y;  // stride 2 ints
            find_partners(a, cand, sig, head, next, dim, stamp, colmask, keys, c, 2);
            thread_sort(keys, c, 2);
            for (int i = 0; i < c; ++i) dst[i].x = va;
        }
    }
}
// connecting edges of the candidates [begin, end) by the calling CTA (NT threads), one thread per
// candidate.  (Measured alternative: 8-lane / 32-lane groups walking the cells of one candidate side
// by side and rank-sorting its list: 2.4x SLOWER on the large model, 348 vs 145 us per launch: the
// coincident-vertex clusters of the reference's chunk-overlap duplicates make long lists the common
// case there, and one list per thread keeps far more dependent-load chains in flight.)
template <int NT>
__device__ __forceinline__ void pair_write_range(int begin, int end, const int *cand, const uint64_t *sig,
                                                 const unsigned long long *head, const tnb_bucket_rec *next, int dim,
                                                 uint32_t stamp, uint64_t colmask, const int *pcount, const int *poff,
                                                 int2 *edges_out, const int *pcache, bool defer_long = false)
{
    for (int a = begin + (int)threadIdx.x; a < end; a += NT)
        pair_write_item(a, cand, sig, head, next, dim, stamp, colmask, pcount, poff, edges_out, pcache, defer_long);
}

// Connecting edges of ONE candidate with a long partner list, by a warp.  The long lists are the
// clusters of coincident vertices the reference's chunk-overlap duplicates leave behind (rows of up to
// 250 vertices in the large model): with one thread per candidate the write pass lasted as long as
// its longest list (a single wave of threads, 145 us for the large sphere).  Eight lanes walk one cell
// of the candidate's box each and append to the list's own output range, then the warp sorts the keys
// (bitonic, in shared memory, up to kLongSortMax) and writes (va, key) pairs in ascending order.
constexpr int kLongSortMax = 1024;
constexpr int kSortWarps = 8;
__device__ __forceinline__ void pair_write_long_warp(int a, const int *cand, const uint64_t *sig, const unsigned long long *head,
                                                     const tnb_bucket_rec *next, int dim, uint32_t stamp, uint64_t colmask,
                                                     const int *pcount, const int *poff, int2 *edges_out, int *keys_smem, int *s_pos)
{
    const int lane = threadIdx.x & 31;
    const int c = pcount[a];
    const int va = cand[a];
    int2 *dst = edges_out + poff[a];
    if (lane == 0) *s_pos = 0;
    __syncwarp();
    const bool in_smem = c <= kLongSortMax;
    {
        const PartnerQuery q = partner_query(va, sig);
        const int nx = q.ba.hi[0] - q.ba.lo[0] + 1, ny = q.ba.hi[1] - q.ba.lo[1] + 1, nz = q.ba.hi[2] - q.ba.lo[2] + 1;
        if (lane < nx * ny * nz) {
            const int cz = q.ba.lo[2] + lane % nz, cy = q.ba.lo[1] + (lane / nz) % ny, cx = q.ba.lo[0] + lane / (nz * ny);
            walk_cell(q, cx, cy, cz, head, next, dim, stamp, colmask, [&](int vb) {
                const int pos = atomicAdd(s_pos, 1);
                if (in_smem) keys_smem[pos] = vb; else dst[pos].y = vb;
            });
        }
    }
    __syncwarp();
    if (in_smem) {
        int n = 32;
        while (n < c) n <<= 1;
        for (int i = c + lane; i < n; i += 32) keys_smem[i] = 0x7fffffff;
        __syncwarp();
        for (int k = 2; k <= n; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = lane; i < n; i += 32) {
                    const int p = i ^ j;
                    if (p > i) {
                        const int x = keys_smem[i], y = keys_smem[p];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) { keys_smem[i] = y; keys_smem[p] = x; }
                    }
                }
                __syncwarp();
            }
        for (int i = lane; i < c; i += 32) dst[i] = make_int2(va, keys_smem[i]);
    } else {  // longer than the shared buffer (never seen): one lane sorts in place in HBM
        if (lane == 0) {
            thread_sort(&dst[0].y, c, 2);
            for (int i = 0; i < c; ++i) dst[i].x = va;
        }
    }
    __syncwarp();
}
// all long lists, by the warps of the calling grid (the first kSortWarps warps of every CTA)
__device__ __forceinline__ void pair_write_long(const int *long_list, int n_long, const int *cand, const uint64_t *sig,
                                                const unsigned long long *head, const tnb_bucket_rec *next, int dim, uint32_t stamp,
                                                uint64_t colmask, const int *pcount, const int *poff, int2 *edges_out)
{
    __shared__ int s_keys[kSortWarps][kLongSortMax];
    __shared__ int s_pos[kSortWarps];
    const int warp = threadIdx.x >> 5;
    if (warp >= kSortWarps) return;
    const int warps_per_cta = min((int)(blockDim.x >> 5), kSortWarps);
    for (int li = blockIdx.x * warps_per_cta + warp; li < n_long; li += gridDim.x * warps_per_cta)
        pair_write_long_warp(long_list[li], cand, sig, head, next, dim, stamp, colmask, pcount, poff, edges_out, s_keys[warp], s_pos + warp);
}
__global__ void __launch_bounds__(kSortWarps * 32) k_pair_write_long(const int *__restrict__ long_list, const int *__restrict__ cnt,
                                                                     const int *__restrict__ cand, const uint64_t *__restrict__ sig,
                                                                     const unsigned long long *__restrict__ head,
                                                                     const tnb_bucket_rec *__restrict__ next, int dim, uint32_t stamp,
                                                                     uint64_t colmask, const int *__restrict__ pcount,
                                                                     const int *__restrict__ poff, int2 *__restrict__ edges_out)
{
    pair_write_long(long_list, cnt[C_LONG], cand, sig, head, next, dim, stamp, colmask, pcount, poff, edges_out);
}
template <int NT>
__device__ __forceinline__ void body_pair_write(const int *cand, int n_cand,
                                                         const uint64_t *sig,
                                                         const unsigned long long *head,
                                                         const tnb_bucket_rec *next, int dim, uint32_t stamp,
                                                         uint64_t colmask, const int *pcount,
                                                         const int *poff, int2 *edges_out, const int *pcache, bool defer_long = false)
{
    for (int base = blockIdx.x * NT; base < n_cand; base += gridDim.x * NT)
        pair_write_range<NT>(base, min(base + NT, n_cand), cand, sig, head, next, dim, stamp, colmask, pcount, poff, edges_out, pcache, defer_long);
}

__global__ void __launch_bounds__(kThreads) k_pair_write(const int *__restrict__ cand, int n_cand,
                                                         const uint64_t *__restrict__ sig,
                                                         const unsigned long long *__restrict__ head,
                                                         const tnb_bucket_rec *__restrict__ next, int dim, uint32_t stamp,
                                                         uint64_t colmask, const int *__restrict__ pcount,
                                                         const int *__restrict__ poff, int2 *__restrict__ edges_out,
                                                         const int *__restrict__ pcache, int defer_long)
{
    body_pair_write<kThreads>(cand, n_cand, sig, head, next, dim, stamp, colmask, pcount, poff, edges_out, pcache, defer_long != 0);
}

// ---- connecting edges over contiguous cell segments (cells.cuh): large complexes ----------------------
// Every candidate is filed ONCE, in the lowest cell of its box.  One WARP per candidate a: the boxes of a
// and b overlap iff b's lowest cell lies in [a.lo - 1, a.hi] per axis, i.e. in one of up to 27 cells
// around a.  Their segments are laid end to end; lane l tests records l, l+32, ... of that
// concatenation, four independent 32-byte loads in flight per lane, hits compacted with ballots.  The
// first kCachedPartners partners go to shared memory; a list that fits is sorted right here (bitonic over
// the lanes) and the write pass is a flat copy.  Longer lists (the coincident-vertex clusters the
// reference's chunk-overlap duplicates leave: one cluster is now streamed once per member, not once per
// cell of its box) are filed in long_list for k_pair_write_long_seg.
__device__ __forceinline__ bool partner_test_once(const PartnerQuery &q, const tnb_bucket_rec &r, uint64_t colmask)
{
    if (r.v <= q.va) return false;
    const uint64_t pb = r.pos, nb = r.neg, gb = r.grd;
    if (((q.pa & nb) | (q.na & pb)) & colmask) return false;  // opposite signs: no common region
    const CellBox bb = cell_box(gb);
    bool ok = true;
    int shared = __popcll(q.za & ~(pb | nb) & colmask);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (max(q.ba.lo[d], bb.lo[d]) > min(q.ba.hi[d], bb.hi[d])) ok = false;  // no common cell
        if (!grid_mask(q.ga, d) && !grid_mask(gb, d) && q.ba.hi[d] == bb.hi[d]) ++shared;
    }
    return ok && shared >= 1;
}
// Streams the neighbourhood of q with the calling warp; emit(hit, vb) is called by ALL lanes once per
// round of 32 records (hit = this lane's record is a partner).  s_incl / s_base: 32 ints each, per warp.
template <class Emit>
__device__ __forceinline__ void stream_partners(const PartnerQuery &q, const int2 *__restrict__ cells,
                                                const tnb_bucket_rec *__restrict__ recs, int dim, uint64_t colmask,
                                                int *s_incl, int *s_base, Emit emit)
{
    const int lane = threadIdx.x & 31;
    const int ny = q.ba.hi[1] - q.ba.lo[1] + 2, nz = q.ba.hi[2] - q.ba.lo[2] + 2;
    const int ncell = (q.ba.hi[0] - q.ba.lo[0] + 2) * ny * nz;
    int cnt = 0, base = 0;
    if (lane < ncell) {
        const int cz = q.ba.lo[2] - 1 + lane % nz, cy = q.ba.lo[1] - 1 + (lane / nz) % ny, cx = q.ba.lo[0] - 1 + lane / (nz * ny);
        if (cx >= -2 && cy >= -2 && cz >= -2) {  // no vertex has a lowest cell below -2 (offsets start at -1)
            const int2 cb = cells[cell_id(cx, cy, cz, dim)];
            cnt = cb.x; base = cb.y;
        }
    }
    const int incl = warp_inclusive_scan(cnt);
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    s_incl[lane] = incl;
    s_base[lane] = base - (incl - cnt);  // record i of the concatenation, if it falls into this cell, is recs[s_base + i]
    __syncwarp();
    // (measured: a one-record-per-lane path for total <= 32 -- nearly every neighbourhood -- made the kernel SLOWER,
    // 70.7 -> 79.0 us per launch, like the uniform breaks below: the predicated four-slot form issues its loads earlier)
    for (int i0 = 0; i0 < total; i0 += 128) {
        tnb_bucket_rec r[4];
        bool have[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u + lane;
            have[u] = i < total;
            if (have[u]) {
                int k = 0;  // first cell whose inclusive prefix exceeds i
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                    if (s_incl[k + step - 1] <= i) k += step;
                r[u] = recs[s_base[k] + i];
            }
        }
        // (measured: leaving the unrolled slots past `total` as a warp -- uniform breaks in both loops -- made the
        // kernel 20 % SLOWER, 71 -> 87 us per launch: the slots' loads no longer issue back to back)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + 32 * u >= total) break;  // warp uniform
            emit(have[u] && partner_test_once(q, r[u], colmask), have[u] ? r[u].v : 0);
        }
    }
}
__global__ void __launch_bounds__(256) k_pair_count_seg(const int *__restrict__ cand, int *__restrict__ cnt,
                                                        const uint64_t *__restrict__ sig, const int2 *__restrict__ cells,
                                                        const tnb_bucket_rec *__restrict__ recs, int dim, uint64_t colmask,
                                                        int *__restrict__ pcount, int *__restrict__ pcache, int *__restrict__ long_list,
                                                        const int *__restrict__ idx_dev = nullptr)
{
    pdl_wait();
    if (idx_dev) {  // device-driven step stream: the hyperplane of this step is chosen on the device
        const int idx = *idx_dev;
        if (idx < 0) return;
        colmask = (1ull << idx) - 1ull;
    }
    __shared__ int s_keys[8][kCachedPartners], s_incl[8][32], s_base[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // The candidates are taken in RECORD order (K = 1: every candidate has exactly one record, and records of one cell
    // and of neighbouring cells are neighbours in memory): the record already holds the candidate's number and packed
    // signs, so the chain candidate -> vertex -> signature (two dependent loads, the top stall of this kernel) is one
    // load, and consecutive warps read the same cell headers and records.
    const int n_recs = cnt[C_RECS], warps = gridDim.x * (blockDim.x >> 5);
    for (int ri = blockIdx.x * (blockDim.x >> 5) + warp; ri < n_recs; ri += warps) {
        const tnb_bucket_rec me = recs[ri];
        const int a = me.next;  // the item number (cells.cuh)
        PartnerQuery q;
        q.va = me.v; q.pa = me.pos; q.na = me.neg; q.ga = me.grd;
        q.za = ~(q.pa | q.na);
        q.ba = cell_box(q.ga);
        int found = 0;
        stream_partners(q, cells, recs, dim, colmask, s_incl[warp], s_base[warp], [&](bool hit, int vb) {
            const unsigned ball = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const int pos = found + __popc(ball & ((1u << lane) - 1u));
                if (pos < kCachedPartners) s_keys[warp][pos] = vb;
            }
            found += __popc(ball);
        });
        __syncwarp();
        if (found > 0 && found <= kCachedPartners) {
            // ascending partner number = the order unique(dim=0) leaves (subpoly.py:243-244)
            const int key = warp_sort_asc(lane < found ? s_keys[warp][lane] : 0x7fffffff, found);
            if (lane < found) pcache[(int64_t)a * kCachedPartners + lane] = key;
        }
        if (lane == 0) {
            pcount[a] = found;
            if (found > kCachedPartners) long_list[atomicAdd(cnt + C_LONG, 1)] = a;  // any order: each list has its own range
        }
        __syncwarp();
    }
}
// connecting edges of the lists that fit the cache: a flat copy, one thread per (candidate, entry)
__global__ void __launch_bounds__(256) k_pair_copy(const int *__restrict__ cand, const int *__restrict__ cnt, const int *__restrict__ pcount,
                                                   const int *__restrict__ poff, const int *__restrict__ pcache, int2 *__restrict__ edges_out)
{
    const int64_t n = (int64_t)cnt[C_CAND] * kCachedPartners;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int a = (int)(t / kCachedPartners), i = (int)(t % kCachedPartners);
        const int c = pcount[a];
        if (i < c && c <= kCachedPartners) edges_out[poff[a] + i] = make_int2(cand[a], pcache[t]);
    }
}
// a warp per long list: stream the neighbourhood again, keys into shared memory, bitonic sort, write.  A launch lasts
// as long as its longest list (there are fewer lists than warps: more CTAs change nothing, profiles/round2_notes.md), and
// for a list of several hundred keys that is the sort: lists of more than kWarpSortMax keys are sorted by the whole CTA
// once its warps have gathered theirs (a stage costs n / 256 compare-exchanges per thread instead of n / 32).
constexpr int kWarpSortMax = 256;
__device__ __forceinline__ void pair_write_long_seg_body(const int *__restrict__ long_list, int n_long, const int *__restrict__ cand,
                                                         const uint64_t *__restrict__ sig, const int2 *__restrict__ cells,
                                                         const tnb_bucket_rec *__restrict__ recs, int dim, uint64_t colmask,
                                                         const int *__restrict__ pcount, const int *__restrict__ poff,
                                                         int2 *__restrict__ edges_out)
{
    __shared__ int s_keys[kSortWarps][kLongSortMax];
    __shared__ int s_incl[kSortWarps][32], s_base[kSortWarps][32];
    __shared__ int s_cnt[kSortWarps], s_va[kSortWarps], s_off[kSortWarps];  // lists left to the CTA (count 0: none)
    constexpr int NT = kSortWarps * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int first = blockIdx.x * kSortWarps; first < n_long; first += gridDim.x * kSortWarps) {  // CTA-uniform trips
        const int li = first + warp;
        int left = 0;
        if (li < n_long) {
            const int a = long_list[li], c = pcount[a], va = cand[a];
            int2 *dst = edges_out + poff[a];
            int *keys = s_keys[warp];
            const bool in_smem = c <= kLongSortMax;
            const PartnerQuery q = partner_query(va, sig);
            int found = 0;
            stream_partners(q, cells, recs, dim, colmask, s_incl[warp], s_base[warp], [&](bool hit, int vb) {
                const unsigned ball = __ballot_sync(0xffffffffu, hit);
                if (hit) {
                    const int pos = found + __popc(ball & ((1u << lane) - 1u));
                    if (in_smem) keys[pos] = vb; else dst[pos].y = vb;
                }
                found += __popc(ball);
            });
            __syncwarp();
            if (in_smem) {
                int n = 32;
                while (n < c) n <<= 1;
                for (int i = c + lane; i < n; i += 32) keys[i] = 0x7fffffff;
                __syncwarp();
                if (n <= kWarpSortMax) {
                    for (int k = 2; k <= n; k <<= 1)
                        for (int j = k >> 1; j > 0; j >>= 1) {
                            for (int i = lane; i < n; i += 32) {
                                const int p = i ^ j;
                                if (p > i) {
                                    const int x = keys[i], y = keys[p];
                                    const bool up = (i & k) == 0;
                                    if ((x > y) == up) { keys[i] = y; keys[p] = x; }
                                }
                            }
                            __syncwarp();
                        }
                    for (int i = lane; i < c; i += 32) dst[i] = make_int2(va, keys[i]);
                } else {
                    left = c;
                    if (lane == 0) { s_va[warp] = va; s_off[warp] = poff[a]; }
                }
            } else {  // longer than the shared buffer (never seen): one lane sorts in place in HBM
                if (lane == 0) {
                    thread_sort(&dst[0].y, c, 2);
                    for (int i = 0; i < c; ++i) dst[i].x = va;
                }
            }
        }
        if (lane == 0) s_cnt[warp] = left;
        __syncthreads();
        for (int w = 0; w < kSortWarps; ++w) {
            const int c = s_cnt[w];
            if (c == 0) continue;  // CTA uniform
            int n = 32;
            while (n < c) n <<= 1;
            int *keys = s_keys[w];
            for (int k = 2; k <= n; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = threadIdx.x; i < n; i += NT) {
                        const int p = i ^ j;
                        if (p > i) {
                            const int x = keys[i], y = keys[p];
                            const bool up = (i & k) == 0;
                            if ((x > y) == up) { keys[i] = y; keys[p] = x; }
                        }
                    }
                    __syncthreads();
                }
            int2 *dst = edges_out + s_off[w];
            const int va = s_va[w];
            for (int i = threadIdx.x; i < c; i += NT) dst[i] = make_int2(va, keys[i]);
        }
        __syncthreads();  // the next trip's gathers overwrite the keys
    }
}
__global__ void __launch_bounds__(kSortWarps * 32) k_pair_write_long_seg(const int *__restrict__ long_list, const int *__restrict__ cnt,
                                                                         const int *__restrict__ cand, const uint64_t *__restrict__ sig,
                                                                         const int2 *__restrict__ cells, const tnb_bucket_rec *__restrict__ recs,
                                                                         int dim, uint64_t colmask, const int *__restrict__ pcount,
                                                                         const int *__restrict__ poff, int2 *__restrict__ edges_out)
{
    pair_write_long_seg_body(long_list, cnt[C_LONG], cand, sig, cells, recs, dim, colmask, pcount, poff, edges_out);
}

// ---- pruning -------------------------------------------------------------------------------------
// hyperplanes that separate the two ends of an edge (beyond eps on both sides), from the packed signs
__device__ __forceinline__ uint64_t edge_cross_bits(const uint64_t *sig, int2 ed)
{
    const uint64_t pa = sig[3 * (int64_t)ed.x], na = sig[3 * (int64_t)ed.x + 1];
    const uint64_t pb = sig[3 * (int64_t)ed.y], nb = sig[3 * (int64_t)ed.y + 1];
    return (pa & nb) | (na & pb);
}
__device__ __forceinline__ void cross_publish(uint64_t m, int *cnt)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, d);
    if ((threadIdx.x & 31) == 0 && m) atomicOr((unsigned long long *)(cnt + C_CROSS), (unsigned long long)m);
}
// crossing mask of a fresh complex (skeleton / caller arrays)
__global__ void k_cross_mask(const int2 *__restrict__ edges, int64_t E, const uint64_t *__restrict__ sig, int *__restrict__ cnt)
{
    uint64_t m = 0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) m |= edge_cross_bits(sig, edges[e]);
    cross_publish(m, cnt);
}
// The count pass of the pruning compaction (KeepCount below) that also ORs the crossing bits of the
// edges.  A pruned edge has equal signs in every later column, so the OR over all edges equals the OR
// over the kept ones in the columns that matter.
template <int NT>
__device__ __forceinline__ void keep_count_cross(int64_t n, const int2 *edges, const uint64_t *sig, uint64_t futmask, int *block_sums,
                                                 int *cnt)
{
    constexpr int NW = NT / 32;
    int64_t begin, end;
    scan_slice(n, begin, end);
    int acc = 0;
    uint64_t m = 0;
    for (int64_t e = begin + threadIdx.x; e < end; e += NT) {
        const int2 ed = edges[e];
        const uint64_t pa = sig[3 * (int64_t)ed.x], na = sig[3 * (int64_t)ed.x + 1];
        const uint64_t pb = sig[3 * (int64_t)ed.y], nb = sig[3 * (int64_t)ed.y + 1];
        acc += (((pa ^ pb) | (na ^ nb)) & futmask) ? 1 : 0;
        m |= (pa & nb) | (na & pb);
    }
    cross_publish(m, cnt);
    acc = warp_sum(acc);
    __shared__ int s[NW];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = (int)threadIdx.x < NW ? s[threadIdx.x] : 0;
        t = warp_sum(t);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
    }
    __syncthreads();
}
__global__ void __launch_bounds__(kScanThreads) k_keep_count_cross(int64_t n, const int *__restrict__ n_dev, const int2 *__restrict__ edges,
                                                                   const uint64_t *__restrict__ sig, uint64_t futmask,
                                                                   int *__restrict__ block_sums, int *__restrict__ cnt)
{
    if (n_dev) n = *n_dev;
    keep_count_cross<kScanThreads>(n, edges, sig, futmask, block_sums, cnt);
}
struct KeepCount {  // subpoly.py:262-264: keep an edge iff its ends differ in a future indicator
    const int2 *edges;
    const uint64_t *sig;
    uint64_t futmask;
    __device__ __forceinline__ int operator()(int64_t e) const
    {
        const int2 ed = edges[e];
        const uint64_t dp = sig[3 * (int64_t)ed.x] ^ sig[3 * (int64_t)ed.y];
        const uint64_t dn = sig[3 * (int64_t)ed.x + 1] ^ sig[3 * (int64_t)ed.y + 1];
        return ((dp | dn) & futmask) ? 1 : 0;
    }
};
struct KeepEmit {
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
struct VertexMoveEmit {  // subpoly.py:268-277: new number of every surviving vertex
    int *remap;
    __device__ __forceinline__ void operator()(int64_t v, int pos, int) const { remap[v] = pos; }
};
// ... and the move itself, flattened over (vertex, column) so that a warp reads and writes
// consecutive floats of the 33-float rows (one thread per row costs 32 sectors per store)
struct VertexArrays {
    const float *vert, *out;
    const uint64_t *sig;
    const unsigned char *tag;
    float *nvert, *nout;
    uint64_t *nsig;
    unsigned char *ntag;
};
__device__ __forceinline__ void body_move_rows(int V, int R, const int *used, const int *remap, const VertexArrays a)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    {   // four independent elements in flight per thread: on small complexes this loop is latency bound
        const float *__restrict__ src = a.out;
        float *__restrict__ dst = a.nout;
        const int64_t total = (int64_t)V * R;
        for (int64_t i = t0; i < total; i += 4 * stride) {
            float val[4];
            int64_t to[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t ii = i + k * stride;
                to[k] = -1;
                if (ii < total) {
                    const int v = (int)(ii / R);
                    if (used[v]) {
                        to[k] = (int64_t)remap[v] * R + (ii - (int64_t)v * R);
                        val[k] = src[ii];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (to[k] >= 0) dst[to[k]] = val[k];
        }
    }
    for (int64_t i = t0; i < (int64_t)V * 3; i += stride) {
        const int v = (int)(i / 3);
        if (used[v]) {
            const int64_t j = (int64_t)remap[v] * 3 + (i - (int64_t)v * 3);
            a.nvert[j] = a.vert[i];
            a.nsig[j] = a.sig[i];
        }
    }
    for (int64_t v = t0; v < V; v += stride)
        if (used[v]) a.ntag[remap[v]] = a.tag[v];
}
__global__ void __launch_bounds__(256) k_move_rows(const int *__restrict__ n_dev, int R, const int *__restrict__ used,
                                                   const int *__restrict__ remap, const VertexArrays a)
{
    body_move_rows(*n_dev, R, used, remap, a);
}

__device__ __forceinline__ void body_remap_edges_dev(int2 *edges, const int *n_dev, const int *remap)
{
    const int64_t E = *n_dev;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        int2 ed = edges[e];
        edges[e] = make_int2(remap[ed.x], remap[ed.y]);
    }
}

__global__ void k_remap_edges_dev(int2 *__restrict__ edges, const int *__restrict__ n_dev, const int *__restrict__ remap)
{
    body_remap_edges_dev(edges, n_dev, remap);
}
__global__ void k_set_counts(int *__restrict__ cnt, int V, int E, int vpar, int epar, int apar)
{
    cnt[C_V] = V;
    if (E >= 0) cnt[C_E] = E;  // E < 0: a compaction just left the edge count there
    cnt[C_VPAR] = vpar;
    cnt[C_EPAR] = epar;
    cnt[C_APAR] = apar;
}
__global__ void k_set_scratch_count(int *__restrict__ p, int v) { *p = v; }
__global__ void k_set_parity(int *__restrict__ cnt, int vpar, int epar, int apar)
{
    cnt[C_VPAR] = vpar;
    cnt[C_EPAR] = epar;
    cnt[C_APAR] = apar;
}
__global__ void k_clear_step_counters(int *__restrict__ cnt)
{
    if (threadIdx.x < C_V) cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) { cnt[C_LONG] = 0; cnt[C_RECS] = 0; }
}


// ---- fused hyperplane step (planar path) ------------------------------------------------------------
// The front half of a step is eleven dependent phases over a few thousand items each; as
// separate launches their cost is launch latency.  Both halves are therefore also available as
// ONE cooperative kernel each: a persistent grid (a multiple of the 148 SMs, all CTAs
// co-resident) walks the phases separated by grid.sync().  Same device functions, same results.
namespace cg = cooperative_groups;

// Everything a step needs lives on the device: sizes, which half of the ping-pong arrays is
// current (C_VPAR / C_EPAR), capacity checks and sticky error bits (C_STICKY).  The host can
// therefore enqueue all 33 steps back to back and sync ONCE at the end.
struct StepArgs {
    int R;
    float eps;
    int Vcap, Ecap, dim;
    int2 *edges[2];
    float *vert[2], *out[2];
    uint64_t *sig[2], *bmask;
    int *split_list, *cand, *pcount, *poff, *used[2], *remap, *block_sums, *cnt, *pcache;
    int *gd;                        // curve path: walks of the gradient-descent repair (repair.cuh)
    tnb_bucket_rec *next;
    unsigned long long *head, *bytes;  // bytes[0/1]: algorithmic bytes of the front / back halves, [2/3]: their units
    unsigned char *tag[2];
    int2 *cslot;                    // contiguous cell segments (cells.cuh): slot of every candidate
    uint32_t *mask_e, *mask_v;      // one bit per edge / per vertex (masked compactions)
    // slab sharding (halo.cuh): the back half runs in two launches around the exchange
    int use_cross;                  // the packed signs were made with this step's eps: the crossing mask decides no-op steps
    int halo, part;                 // part 0: whole back half, 1: up to the exchange, 2: after it
    int has_lower, has_upper;
    int *hslot, *stage_count, stage_cap;
    unsigned char *stage[2];
    const unsigned char *in[2];     // liveness bytes received from the lower / upper neighbour
    long long *dbg;                 // TNB_PHASE_TRACE: globaltimer stamps after every phase
};
// what changes from one hyperplane to the next
struct StepVar {
    int idx, do_prune;
    uint32_t stamp;       // bucket generation of this step
    uint64_t colmask, futmask;
};
__host__ __device__ __forceinline__ StepVar step_var(int idx, int R, int do_prune, uint32_t stamp)
{
    StepVar v;
    v.idx = idx;
    v.do_prune = do_prune;
    v.stamp = stamp;
    v.colmask = (1ull << idx) - 1ull;
    v.futmask = ~v.colmask & (R >= 64 ? ~0ull : ((1ull << R) - 1ull));
    return v;
}
// TNB_PHASE_TRACE: time spent up to each phase mark, accumulated over the steps of one launch
// (dbg[63] = time of the previous mark; marks 0 and 16 open the front / back half)
#define TNB_PHASE_MARK(k) do { if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) { const long long now_ = global_ns(); \
        if ((k) != 0 && (k) != 16) a.dbg[k] += now_ - a.dbg[63]; a.dbg[63] = now_; } } while (0)

// How the CTAs of a persistent step kernel wait for each other between phases: the whole
// cooperative grid (one CTA per SM), or ONE thread-block cluster (barrier.cluster: ~0.2 us instead
// of ~2.5 us, which is what a small complex needs: its phases are a few microseconds each).
struct GridSync {
    cg::grid_group g;
    __device__ __forceinline__ void operator()() { g.sync(); }
};
struct ClusterSync {
    __device__ __forceinline__ void operator()() { cg::this_cluster().sync(); }
};

struct TagCount {  // vertices on one shared slab plane
    const unsigned char *tag;
    int bit;
    __device__ __forceinline__ int operator()(int64_t v) const { return (tag[v] & bit) ? 1 : 0; }
};
struct StageEmit {  // k-th vertex of the plane: remember k, publish its liveness
    int *slot;
    unsigned char *stage;
    const int *used;
    int cap;  // a longer plane list is reported by the count (k_halo_send poisons the message)
    __device__ __forceinline__ void operator()(int64_t v, int pos, int) const
    {
        slot[v] = pos;
        if (pos < cap) stage[pos] = used[v] ? 1 : 0;
    }
};

template <class C, int NT, class Sync>
__device__ __forceinline__ void step_front(const NetMeta &n, const StepArgs &a, const StepVar sv, Sync sync)
{
    TNB_PHASE_MARK(0);
    int *cnt = a.cnt;
    if (cnt[C_STICKY]) return;  // an earlier step failed: uniform exit, nobody reaches a grid sync
    const int E = cnt[C_E], V = cnt[C_V], pv = cnt[C_VPAR], pe = cnt[C_EPAR];
    int2 *edges = a.edges[pe];
    float *vert = a.vert[pv], *out = a.out[pv];
    uint64_t *sig = a.sig[pv];
    // the words read above (>= C_V) are not among the transient counters block 0 clears now
    if (blockIdx.x == 0 && threadIdx.x < C_V) cnt[threadIdx.x] = 0;
    const SplitCount sc{edges, out, n.R, sv.idx, a.eps};
    scan_count_body_t<NT>(E, sc, a.block_sums);
    sync();
    TNB_PHASE_MARK(1);
    scan_write_body_t<NT>(E, sc, ListEmit{a.split_list}, a.block_sums, cnt + C_RAW);
    sync();
    TNB_PHASE_MARK(2);
    // subpoly.py:110-111: the plane crosses no edge -> the step changes nothing (uniform exit; a
    // slab must go on, another slab may have crossed: halo.cuh)
    if (cnt[C_RAW] == 0 && !a.halo) return;
    body_new_vertices<C>(n, sv.idx, a.eps, a.Vcap, a.Ecap, a.split_list, edges, vert, out, sig, a.bmask, cnt, a.tag[pv]);
    const HitCount hc{out, a.used[cnt[C_APAR]], n.R, sv.idx, a.eps};  // old vertices only: independent of the new rows
    scan_count_body_t<NT>(V, hc, a.block_sums);
    sync();
    TNB_PHASE_MARK(3);
    body_finalize_new(n, vert, out, sig, a.bmask, cnt);
    {   // every CTA knows the hit count from the block sums: the new vertices' candidate slots
        // [H, H+S) can be filled in the same phase as the hit list [0, H)
        __shared__ int s_h[NT / 32];
        int acc = 0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += NT) acc += a.block_sums[b];
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) s_h[threadIdx.x >> 5] = acc;
        __syncthreads();
        int Hn = 0;
        for (int w = 0; w < NT / 32; ++w) Hn += s_h[w];
        __syncthreads();
        if (!cnt[C_OVERFLOW]) {
            const int S = cnt[C_SPLIT];
            for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < S; k += gridDim.x * blockDim.x) a.cand[Hn + k] = V + k;
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                // a slab cannot know yet whether another slab crossed the plane: it prepares the
                // hit vertices' connecting edges anyway and the exchange decides (halo.cuh)
                cnt[C_CAND] = (cnt[C_RAW] || a.halo) ? Hn + S : 0;
                // split scan (edge + two cached outputs, two passes), hit scan, new vertices, buckets + partner count
                a.bytes[0] += 2ull * 16 * E + 2ull * 4 * V + (unsigned long long)S * (8 + 4 + 2 * (12 + 4 + 16) + 12 + 4 * n.R + 8 + 16) +
                              (unsigned long long)(Hn + S) * (24 + 8 + 4 + 24);
            }
        }
    }
    scan_write_body_t<NT>(V, hc, ListEmit{a.cand}, a.block_sums, cnt + C_HIT);
    sync();
    TNB_PHASE_MARK(4);
    if (sv.do_prune && !cnt[C_OVERFLOW]) {  // the back half marks the vertices that keep an edge in the idle half of the liveness array
        int *used = a.used[cnt[C_APAR] ^ 1];
        const int Vn = V + cnt[C_SPLIT];
        for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < Vn; v += gridDim.x * blockDim.x) used[v] = 0;
    }
    body_bucket_insert(a.cand, cnt, sig, a.head, a.next, a.dim, sv.stamp);
    sync();
    TNB_PHASE_MARK(5);
    const int n_cand = cnt[C_CAND];
    {   // partner search on this CTA's slice of the candidates (8 lanes each), then the slice's sum
        __shared__ int s_grp[NT / 8];
        int64_t begin, end;
        scan_slice(n_cand, begin, end);
        pair_count_groups((int)begin, (int)end, a.cand, sig, a.head, a.next, a.dim, sv.stamp, sv.colmask, a.pcount, a.pcache, s_grp);
        __syncthreads();
    }
    scan_count_body_t<NT>(n_cand, ArrayCount{a.pcount}, a.block_sums);
    sync();
    TNB_PHASE_MARK(6);
    scan_write_body_t<NT>(n_cand, ArrayCount{a.pcount}, OffsetEmit{a.poff}, a.block_sums, cnt + C_PAIRS);
    TNB_PHASE_MARK(7);
}

template <class C>
__global__ void __launch_bounds__(kScanThreads, 2) k_step_front(const __grid_constant__ NetMeta n, const StepArgs a, const StepVar sv)
{
    step_front<C, kScanThreads>(n, a, sv, GridSync{cg::this_grid()});
}

template <int NT, class Sync>
__device__ __forceinline__ void step_back(const StepArgs &a, const StepVar sv, Sync sync)
{
    TNB_PHASE_MARK(16);
    int *cnt = a.cnt;
    // all decisions are uniform over the grid (same device words read by everybody before any write)
    const int sticky = cnt[C_STICKY], raw = cnt[C_RAW], overflow = cnt[C_OVERFLOW];
    const int S = cnt[C_SPLIT], P = cnt[C_PAIRS], V0 = cnt[C_V], E0 = cnt[C_E], n_cand = cnt[C_CAND];
    const int pv = cnt[C_VPAR], pe = cnt[C_EPAR], pa = cnt[C_APAR];
    const int local_flag = cnt[C_FLAG];
    const int word = a.part == 2 ? a.stage_count[2] : 0;  // OR of every slab's status word
    if (sticky) return;
    if (!a.halo && raw == 0) return;  // subpoly.py:110-111: nothing crossed, nothing changes
    const int64_t En = (int64_t)E0 + S + P;
    const int Vn = V0 + S;
    int2 *edges = a.edges[pe], *edges_dst = a.edges[pe ^ 1];
    float *out = a.out[pv];
    uint64_t *sig = a.sig[pv];
    int *alive = a.used[pa], *used = a.used[pa ^ 1];  // `used` was cleared by the front half
    int *kept = a.halo ? cnt + C_KEPT : cnt + C_E;
    if (a.part != 2) {
        if (overflow || En > a.Ecap) {  // the host re-runs the extraction with larger arrays
            if (blockIdx.x == 0 && threadIdx.x == 0) cnt[C_STICKY] = kStickyCapacity;
            return;
        }
        if (P > 0)
            body_pair_write<NT>(a.cand, n_cand, sig, a.head, a.next, a.dim, sv.stamp, sv.colmask, a.pcount, a.poff, edges + E0 + S, a.pcache);
        if (!sv.do_prune) {  // the output neuron (subpoly.py:253): sizes only
            if (a.halo) {   // no liveness to exchange, only the status word
                if (blockIdx.x == 0 && threadIdx.x == 0) a.stage_count[0] = a.stage_count[1] = 0;
                return;
            }
            for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < S; k += gridDim.x * blockDim.x) alive[V0 + k] = 1;
            sync();
            TNB_PHASE_MARK(17);
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                cnt[C_V] = Vn;
                cnt[C_E] = (int)En;
                a.bytes[1] += (unsigned long long)n_cand * 28 + (unsigned long long)P * 8;
            }
            return;
        }
        sync();
        TNB_PHASE_MARK(18);
        const KeepCount kc{edges, sig, sv.futmask};
        scan_count_body_t<NT>(En, kc, a.block_sums);
        sync();
        TNB_PHASE_MARK(19);
        scan_write_body_t<NT>(En, kc, KeepEmit{edges, edges_dst, used}, a.block_sums, kept);
        sync();
        TNB_PHASE_MARK(20);
        if (a.halo) {  // ordered lists of the two shared planes' vertices and their liveness
            for (int side = 0; side < 2; ++side) {
                const TagCount tc{a.tag[pv], 1 << side};
                scan_count_body_t<NT>(Vn, tc, a.block_sums);
                sync();
                TNB_PHASE_MARK(21);
                scan_write_body_t<NT>(Vn, tc, StageEmit{a.hslot, a.stage[side], used, a.stage_cap}, a.block_sums, a.stage_count + side);
                sync();
                TNB_PHASE_MARK(22);
            }
            return;  // k_halo_send / k_halo_recv run between the two launches
        }
    } else {
        if (!(word & kWordRaw)) return;  // no slab crossed the plane: the step changes nothing
        if ((word & kWordFlag) && !local_flag) {
            // another slab raised the failover override (subpoly_debug.py:41-49): it applies to the
            // new vertices here too.  Their masked entries are within eps (else the flag would be
            // up here as well), so the packed signs stay as they are.
            for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < S; k += gridDim.x * blockDim.x)
                for (uint64_t m = a.bmask[k]; m; m &= m - 1) out[((int64_t)V0 + k) * a.R + __ffsll((long long)m) - 1] = 0.0f;
        }
        if (!sv.do_prune) {
            for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < S; k += gridDim.x * blockDim.x) alive[V0 + k] = 1;
            sync();  // everybody has read the counter block
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                cnt[C_V] = Vn;
                cnt[C_E] = (int)En;
            }
            return;
        }
        // a vertex on a shared plane lives if an edge on EITHER side of the plane keeps it; a dead
        // vertex leaves the plane lists for good (both sides agree on who died)
        unsigned char *tag = a.tag[pv];
        for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < Vn; v += gridDim.x * blockDim.x) {
            const int t = tag[v];
            int u = used[v];
            if ((t & 1) && a.has_lower && a.hslot[v] < a.stage_cap && a.in[0][a.hslot[v]]) u = 1;
            if ((t & 2) && a.has_upper && a.hslot[v] < a.stage_cap && a.in[1][a.hslot[v]]) u = 1;
            used[v] = u;
            if (!u && t) tag[v] = 0;
        }
        sync();  // everybody has read the counter block (the parked edge count among it)
        TNB_PHASE_MARK(23);
    }
    // Commit: the freshly marked half of the liveness array and the compacted half of the edge
    // array become current.  The rows of dead vertices stay where they are (complex.cuh).
    TNB_PHASE_MARK(31);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (a.halo) cnt[C_E] = cnt[C_KEPT];
        cnt[C_V] = Vn;
        cnt[C_EPAR] = pe ^ 1;
        cnt[C_APAR] = pa ^ 1;
        a.bytes[1] += (unsigned long long)n_cand * 28 + (unsigned long long)P * 8 + (unsigned long long)En * (8 + 2 * 48) +
                      (unsigned long long)Vn * 8;
    }
}

__global__ void __launch_bounds__(kScanThreads, 2) k_step_back(const StepArgs a, const StepVar sv)
{
    step_back<kScanThreads>(a, sv, GridSync{cg::this_grid()});
}

// ---- persistent step loop ---------------------------------------------------------------------------
// One hyperplane, start to finish, inside a kernel that stays resident over all hyperplanes of
// the extraction.  Compared with the front/back pair above, the phases are arranged so that every
// CTA keeps working on what it produced itself: the CTA that finds a crossed edge also makes its
// new vertex, finalises it, files it into the cell buckets; the CTA that counts a candidate's
// partners also writes its connecting edges.  Offsets and totals come from the per-CTA block
// sums (every CTA adds them up itself), never from a word another CTA publishes, so a compaction
// costs ONE barrier: 7 barriers per hyperplane that crosses something, 1 for one that does not
// (subpoly.py:110-111), and nothing returns to the host in between.
//   returns 0 = done, 2 = nothing to do, 1 = work arrays too small (uniform over the CTAs)
//   kCurve: the curve-approximation path (force=False).  The candidates of a CTA's crossed edges go to
//   temporary rows (the idle half of the vertex arrays); strict_check's keep decision needs the
//   failover flag of ALL candidates and the survivors need their ordered rank: two more barriers.
template <class C, int NT, bool kCurve, class Sync>
__device__ __forceinline__ int step_fused(const NetMeta &n, const StepArgs &a, const StepVar sv, int parity, const float *marks, Sync sync)
{
    int *cnt = a.cnt;
    const int E = cnt[C_E], V = cnt[C_V], pv = cnt[C_VPAR], pe = cnt[C_EPAR], pa = cnt[C_APAR];
    int2 *edges = a.edges[pe], *edges_dst = a.edges[pe ^ 1];
    float *vert = a.vert[pv], *out = a.out[pv];
    uint64_t *sig = a.sig[pv];
    unsigned char *tag = a.tag[pv];
    int *alive = a.used[pa], *used = a.used[pa ^ 1];
    const int nb = (int)gridDim.x, R = n.R;
    int *sums_x = a.block_sums + (parity ? nb : 0), *sums_y = a.block_sums + 2 * nb;
    // Nothing crosses this plane (the last pruning pass looked): subpoly.py:110-111 without touching an
    // edge.  The mask was published before the barrier that ended the step which wrote it.
    unsigned long long *cross = (unsigned long long *)(cnt + C_CROSS);
    if (a.use_cross && !((*cross >> sv.idx) & 1ull)) return 2;
    TNB_PHASE_MARK(0);
    if (blockIdx.x == 0 && threadIdx.x == 0) { cnt[C_FLAG] = 0; cnt[C_ERR] = 0; cnt[C_LONG] = 0; }  // last read at least one barrier ago
    if constexpr (kCurve) {
        if (blockIdx.x == 0 && threadIdx.x < 32) {  // an earlier hyperplane needed the repair: clear its walks
            const int had = a.gd[GD_COUNT];
            __syncwarp();
            if (had && threadIdx.x < GD_MASK + kGdWords) a.gd[threadIdx.x] = threadIdx.x < GD_MASK ? 0 : -1;
        }
    }
    // P0: edges the plane crosses, per CTA slice
    const SplitCount sc{edges, out, R, sv.idx, a.eps};
    scan_count_body_t<NT>(E, sc, sums_x);
    sync();
    TNB_PHASE_MARK(1);
    // P1: split list of this CTA's slice, and straight away the new vertices of exactly those edges
    int s_base, S;
    block_sums_reduce<NT>(sums_x, s_base, S);
    if (S == 0) return 2;  // the plane crosses no edge: the step changes nothing, and published nothing
    if ((int64_t)V + S > a.Vcap || (int64_t)E + S > a.Ecap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) cnt[C_STICKY] = kStickyCapacity;
        return 1;
    }
    int s_end = scan_write_from<NT>(E, sc, ListEmit{a.split_list}, s_base);
    float *tvert = a.vert[pv ^ 1], *tout = a.out[pv ^ 1];  // curve path: temporary rows of the candidates
    int *sflag = a.pcount;                                 // free until P3
    if constexpr (kCurve) {
        // The crossed edges cluster in a few CTAs' slices, and a curved candidate is a warp's work, not a
        // thread's (curve.cuh): once the whole split list is in place the candidates are shared out evenly.
        // From here on the CTA owns the crossed edges [s_base, s_end) of that share.
        sync();
        TNB_PHASE_MARK(9);
        const int per = (S + nb - 1) / nb;
        s_base = min(S, (int)blockIdx.x * per);
        s_end = min(S, s_base + per);
    }
    {
        int any = 0;
        if constexpr (kCurve) {
            // the CTA's j-th candidate goes to warp j % NW, lane (j / NW) % 32: the warps share the curved
            // edges, which a warp works through one after the other (curve.cuh)
            constexpr int NW = NT / 32;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            for (int j0 = 0; s_base + j0 < s_end; j0 += NT) {
                const int k = s_base + j0 + lane * NW + warp;
                any |= curve_candidate_item<C>(n, sv.idx, a.eps, k < s_end, k, a.split_list, edges, vert, out, sig, tvert, tout, a.bmask, sflag, cnt, a.gd);
            }
        } else {
            for (int k = s_base + (int)threadIdx.x; k < s_end; k += NT)
                any |= new_vertex_item<C>(n, sv.idx, a.eps, k, V, E, a.split_list, edges, vert, out, sig, a.bmask, tag);
        }
        if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) atomicOr(cnt + C_FLAG, 1);
    }
    const HitCount hc{out, alive, R, sv.idx, a.eps};  // old vertices only: independent of the new rows
    scan_count_body_t<NT>(V, hc, sums_y);
    sync();
    TNB_PHASE_MARK(2);
    int flag = cnt[C_FLAG];
    const int S_raw = S;
    if constexpr (kCurve) {
        // P1b: an intersection the path cannot place ends the extraction (uniform: C_ERR was last
        // written before the barrier); else strict_check (subpoly_debug.py:234-271) on the own candidates
        int err = cnt[C_ERR];
        if (err == kErrRepair) {
            // intersections off their planes walk down the gradient first (subpoly_debug.py:121-165, repair.cuh):
            // a failover, two more barriers, one thread per walk
            const int tid = (int)blockIdx.x * NT + (int)threadIdx.x, nthreads = nb * NT;
            gd_complex_note<C>(n, sv.idx, a.eps, a.split_list, edges, vert, a.gd, tid, nthreads);
            sync();
            if (gd_complex_walk<C>(n, sv.idx, a.eps, a.split_list, edges, vert, tvert, tout, a.bmask, cnt, a.gd, tid, nthreads))
                atomicOr(cnt + C_FLAG, 1);
            sync();
            err = cnt[C_ERR] & ~kErrRepair;
            flag = cnt[C_FLAG];
        }
        if (err) {
            if (blockIdx.x == 0 && threadIdx.x == 0)
                atomicOr(cnt + C_STICKY, ((err & kErrNoPlane) ? kStickyNoPlane : 0) | ((err & kErrGradientDescent) ? kStickyGradientDescent : 0));
            return 1;
        }
        for (int k = s_base + (int)threadIdx.x; k < s_end; k += NT) strict_keep_item(R, sv.idx, a.eps, tout, a.bmask, sflag, flag, k);
        __syncthreads();
        scan_count_range<NT>(s_base, s_end, FlagCount{sflag}, sums_x);  // everybody finished reading sums_x before the barrier above
        sync();
        TNB_PHASE_MARK(8);
        // P1c: the survivors move to their final slots V + rank, in crossed-edge order; edges are rewired
        int r_base;
        block_sums_reduce<NT>(sums_x, r_base, S);
        const CurveCommitEmit ce{nullptr, a.split_list, edges, tvert, tout, vert, out, sig, marks, cnt, R, n.n_marks, n.eps,
                                 n.pre_scale, n.pre_2s, n.pre_inv, n.pre_pow2, tag};
        const int r_end = scan_write_range<NT>(s_base, s_end, FlagCount{sflag}, ce, r_base);
        s_base = r_base;  // from here on: the CTA's own NEW VERTICES are the ranks [s_base, s_end)
        s_end = r_end;
    }
    // P2: failover override + packed signs of the own new vertices; candidate list (hit old vertices,
    // then the new ones); the own candidates go into the cell buckets
    if (blockIdx.x == 0 && threadIdx.x == 0) *cross = sv.do_prune ? 0ull : ~0ull;  // everybody read it before the first barrier; P5 rebuilds it
    TNB_PHASE_MARK(10);
    int h_base, Hn;
    block_sums_reduce<NT>(sums_y, h_base, Hn);
    for (int k = s_base + (int)threadIdx.x; k < s_end; k += NT) {
        if constexpr (!kCurve) finalize_item(n, marks, vert, out, sig, a.bmask, flag, V, k);
        a.cand[Hn + k] = V + k;
    }
    TNB_PHASE_MARK(11);
    const int h_end = scan_write_from<NT>(V, hc, ListEmit{a.cand}, h_base);  // ends in a CTA barrier
    TNB_PHASE_MARK(12);
    for (int c = h_base + (int)threadIdx.x; c < h_end; c += NT) bucket_insert_item(c, a.cand[c], sig, a.head, a.next, a.dim, sv.stamp);
    for (int k = s_base + (int)threadIdx.x; k < s_end; k += NT) bucket_insert_item(Hn + k, V + k, sig, a.head, a.next, a.dim, sv.stamp);
    TNB_PHASE_MARK(13);
    const int Vn = V + S, n_cand = Hn + S;
    if (sv.do_prune) {  // P6 marks the vertices that keep an edge in the idle half of the liveness array
        for (int v = blockIdx.x * NT + threadIdx.x; v < Vn; v += nb * NT) used[v] = 0;
    } else {
        for (int k = s_base + (int)threadIdx.x; k < s_end; k += NT) alive[V + k] = 1;
    }
    sync();
    TNB_PHASE_MARK(3);
    // P3: partners of this CTA's slice of the candidates (8 lanes each)
    int64_t c_begin, c_end;
    scan_slice(n_cand, c_begin, c_end);
    {
        __shared__ int s_grp[NT / 8];
        pair_count_groups((int)c_begin, (int)c_end, a.cand, sig, a.head, a.next, a.dim, sv.stamp, sv.colmask, a.pcount, a.pcache, s_grp,
                          a.remap, cnt + C_LONG);  // remap is idle during the steps: the list of long partner lists
        __syncthreads();
    }
    scan_count_body_t<NT>(n_cand, ArrayCount{a.pcount}, sums_x);
    sync();
    TNB_PHASE_MARK(4);
    // P4: offsets of the slice, then its connecting edges
    int p_base, P;
    block_sums_reduce<NT>(sums_x, p_base, P);
    const int64_t En = (int64_t)E + S + P;
    if (En > a.Ecap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) cnt[C_STICKY] = kStickyCapacity;
        return 1;
    }
    if (P > 0) {
        scan_write_from<NT>(n_cand, ArrayCount{a.pcount}, OffsetEmit{a.poff}, p_base);
        pair_write_range<NT>((int)c_begin, (int)c_end, a.cand, sig, a.head, a.next, a.dim, sv.stamp, sv.colmask, a.pcount, a.poff, edges + E + S, a.pcache, true);
        const int n_long = cnt[C_LONG];  // complete since the barrier before this phase
        if (n_long > 0) {
            // the offsets of ALL candidates must be in place before a warp takes somebody else's list
            sync();
            pair_write_long(a.remap, n_long, a.cand, sig, a.head, a.next, a.dim, sv.stamp, sv.colmask, a.pcount, a.poff, edges + E + S);
        }
    }
    if (!sv.do_prune) {  // the output neuron (subpoly.py:253): sizes only
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            cnt[C_V] = Vn;
            cnt[C_E] = (int)En;
            a.bytes[0] += 2ull * 16 * E + 2ull * 4 * V + (unsigned long long)S_raw * (8 + 4 + 2 * (12 + 4 + 16) + 12 + 4 * R + 8 + 16) +
                          (unsigned long long)n_cand * (24 + 8 + 4 + 24);
            a.bytes[1] += (unsigned long long)n_cand * 28 + (unsigned long long)P * 8;
        }
        return 0;
    }
    sync();
    TNB_PHASE_MARK(5);
    // P5 / P6: pruning (subpoly.py:252-277): keep the edges whose ends differ in a future indicator
    const KeepCount kc{edges, sig, sv.futmask};
    keep_count_cross<NT>(En, edges, sig, sv.futmask, sums_y, cnt);
    sync();
    TNB_PHASE_MARK(6);
    int k_base, kept;
    block_sums_reduce<NT>(sums_y, k_base, kept);
    scan_write_from<NT>(En, kc, KeepEmit{edges, edges_dst, used}, k_base);
    // Commit: the freshly marked half of the liveness array and the compacted half of the edge array
    // become current; the rows of dead vertices stay where they are (complex.cuh).  The caller's
    // barrier after the step orders these words before the next step reads them.
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cnt[C_V] = Vn;
        cnt[C_E] = kept;
        cnt[C_EPAR] = pe ^ 1;
        cnt[C_APAR] = pa ^ 1;
        a.bytes[0] += 2ull * 16 * E + 2ull * 4 * V + (unsigned long long)S_raw * (8 + 4 + 2 * (12 + 4 + 16) + 12 + 4 * R + 8 + 16) +
                      (unsigned long long)n_cand * (24 + 8 + 4 + 24);
        a.bytes[1] += (unsigned long long)n_cand * 28 + (unsigned long long)P * 8 + (unsigned long long)En * (8 + 2 * 48) +
                      (unsigned long long)Vn * 8;
    }
    TNB_PHASE_MARK(7);
    return 0;
}

constexpr int kClusterThreads = 512;
constexpr int kMarksInSmem = 1312;  // every marks grid the dense cell-head array admits (<= 1288 per axis)
constexpr int kMaxStepList = 128;
struct StepList {
    int n;
    uint32_t stamp0;                  // step i uses bucket generation stamp0 + i
    unsigned char idx[kMaxStepList];  // output column of the hyperplane
    unsigned char prune[kMaxStepList];
};
template <class C, int NT, bool kCurve, class Sync>
__device__ __forceinline__ void steps_loop(const NetMeta &n, const StepArgs &a, const StepList &list, Sync sync)
{
    if (a.cnt[C_STICKY]) return;  // an earlier launch failed (nobody writes this word before the first barrier)
    // The marks stay in shared memory for the whole launch: a grid barrier's fence leaves L1 cold, so the
    // three binary searches of a new vertex's grid word were 24 dependent L2 round trips (~6 us per step).
    __shared__ float s_marks[kMarksInSmem];
    const float *marks = n.marks;
    if (n.n_marks <= kMarksInSmem) {
        for (int i = threadIdx.x; i < n.n_marks; i += NT) s_marks[i] = n.marks[i];
        __syncthreads();
        marks = s_marks;
    }
    for (int i = 0; i < list.n; ++i) {
        const StepVar sv = step_var(list.idx[i], a.R, list.prune[i], list.stamp0 + (uint32_t)i);
        const int r = step_fused<C, NT, kCurve>(n, a, sv, i & 1, marks, sync);
        if (r == 1) return;
        if (r == 0) sync();  // the next step reads the sizes and buffer parities this one published
    }
}
// ... by the cooperative grid (one CTA per SM): the default
template <class C>
__global__ void __launch_bounds__(kScanThreads, 1) k_steps_grid(const __grid_constant__ NetMeta n, const __grid_constant__ StepArgs a,
                                                                const __grid_constant__ StepList list)
{
    steps_loop<C, kScanThreads, false>(n, a, list, GridSync{cg::this_grid()});
}
// ... the curve-approximation path (force=False) the same way
template <class C>
__global__ void __launch_bounds__(kScanThreads, 1) k_steps_grid_curve(const __grid_constant__ NetMeta n, const __grid_constant__ StepArgs a,
                                                                      const __grid_constant__ StepList list)
{
    steps_loop<C, kScanThreads, true>(n, a, list, GridSync{cg::this_grid()});
}
// ... by two CTAs per SM (<= 128 registers): twice the threads in flight for a complex whose phases are
// throughput bound, at the price of a slightly dearer barrier
template <class C>
__global__ void __launch_bounds__(kScanThreads, 2) k_steps_grid2(const __grid_constant__ NetMeta n, const __grid_constant__ StepArgs a,
                                                                 const __grid_constant__ StepList list)
{
    steps_loop<C, kScanThreads, false>(n, a, list, GridSync{cg::this_grid()});
}
// ... by ONE thread-block cluster of 16 CTAs: a barrier costs a tenth, but 16 SMs do the work of
// 148.  Slower for one object; it leaves 132 SMs to other objects' clusters.
template <class C>
__global__ void __launch_bounds__(kClusterThreads, 1) k_steps_cluster(const __grid_constant__ NetMeta n, const __grid_constant__ StepArgs a,
                                                                      const __grid_constant__ StepList list)
{
    steps_loop<C, kClusterThreads, false>(n, a, list, ClusterSync{});
}

// co-resident grid size of a cooperative kernel: blocks/SM x SMs, capped by the scan tables
template <class K>
static int coop_blocks(K kernel)
{
    int dev = 0, sms = kSMs, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kScanThreads, 0) != cudaSuccess || per_sm < 1) return 0;
    return std::min(per_sm * sms, kScanMaxBlocks);
}
static int64_t g_fused_max_items = std::getenv("TNB_FUSED_MAX_ITEMS") ? std::atoll(std::getenv("TNB_FUSED_MAX_ITEMS")) : 700000;  // larger complexes: multi-launch path (measured: the large model's 2.3 M items take 6.2 ms in the persistent kernel, 3.7 ms as separate full-size launches)
static bool g_fused_steps = std::getenv("TNB_NO_FUSED_STEPS") == nullptr;  // A/B switch for profiling

// Refresh the host's view of the complex size (one small D2H + sync).
int complex_sync_counts(tnb_complex *c, cudaStream_t s)
{
    if (c->sticky_rc) { set_error(c->sticky_msg); return c->sticky_rc; }
    if (!c->counts_stale && !c->cross_stale) return TNB_OK;
    int rc = read_counters(c, s);
    if (rc) return rc;
    c->V = c->h_counters[C_V];
    c->E = c->h_counters[C_E];
    c->vcur = c->h_counters[C_VPAR];
    c->ecur = c->h_counters[C_EPAR];
    c->acur = c->h_counters[C_APAR];
    memcpy(&c->cross, c->h_counters + C_CROSS, sizeof(uint64_t));
    c->cross_stale = false;
    c->counts_stale = false;
    if (c->bytes.p) {  // algorithmic bytes the fused kernels accumulated on the device
        unsigned long long hb[4] = {0, 0, 0, 0};
        TNB_CUDA(cudaMemcpyAsync(hb, c->bytes.p, sizeof(hb), cudaMemcpyDeviceToHost, s));
        TNB_CUDA(cudaMemsetAsync(c->bytes.p, 0, sizeof(hb), s));
        TNB_CUDA(cudaStreamSynchronize(s));
        if (c->halo.enabled || c->bytes_by_half) {
            prof_add(TNB_PROF_NEW_VERTICES, (int64_t)hb[2], (int64_t)hb[0]);
            prof_add(TNB_PROF_PAIRS, (int64_t)hb[3], (int64_t)hb[1]);
            c->bytes_by_half = false;
        } else {
            prof_add(TNB_PROF_STEPS, 0, (int64_t)(hb[0] + hb[1]));
        }
    }
    // A sticky device error freezes the complex at the failing step (the step kernels exit early once
    // a bit is set): latch it, so that EVERY later call on this complex reports it, not only this one.
    const int sticky = c->h_counters[C_STICKY];
    auto latch = [&](int rc, const std::string &msg) {
        c->sticky_rc = rc;
        c->sticky_msg = msg;
        set_error(msg);
        return rc;
    };
    if (sticky & kStickyCapacity)
        return latch(TNB_ERR_CAPACITY, "work buffers too small for this complex (capacity factor " + std::to_string(capacity_factor()) + ")");
    if (sticky & kStickyNoPlane)
        return latch(TNB_ERR_INVALID, "curve path: a non-axis-aligned edge lies on no earlier plane (the reference exits here, subpoly.py:140-148)");
    if (sticky & kStickyGradientDescent)
        return latch(TNB_ERR_UNSUPPORTED, "curve path: an intersection is still off its planes after the gradient-descent repair (subpoly_debug.py:121-165); the reference ends the extraction here (subpoly.py:172-174)");
    if (sticky & kStickyHaloPayload) return latch(TNB_ERR_CAPACITY, "slab exchange: a shared plane holds more vertices than the mailbox payload");
    if (sticky & kStickyHaloTimeout) return latch(TNB_ERR_CUDA, "slab exchange: a peer did not answer within the timeout");
    if (sticky & kStickyHaloMismatch) return latch(TNB_ERR_INVALID, "slab exchange: the two sides of a shared plane disagree on its vertex count");
    if (sticky & kStickyHaloPeer) return latch(TNB_ERR_CAPACITY, "slab exchange: another slab reported an error");
    return TNB_OK;
}

// ---- slab exchange (halo.cuh) ------------------------------------------------------------------
static long long g_halo_timeout_ns = 2000000000ll;

static HaloArgs halo_args(tnb_complex *c, int word_raw_index, int word_flag_index)
{
    tnb_halo &h = c->halo;
    HaloArgs a;
    memset(&a, 0, sizeof(a));
    a.rank = h.rank; a.world = h.world; a.seq = h.seq; a.parity = (int)(h.seq & 1u);
    a.payload = h.payload;
    for (int r = 0; r < h.world; ++r) a.boxes[r] = h.boxes[r];
    a.stage[0] = h.stage[0].p; a.stage[1] = h.stage[1].p;
    a.stage_count = h.stage_count.p;
    a.cnt = c->counters.p;
    a.has[0] = h.tag_lower ? 1 : 0; a.has[1] = h.tag_upper ? 1 : 0;
    a.word_raw_index = word_raw_index; a.word_flag_index = word_flag_index; a.sticky_index = C_STICKY;
    a.timeout_ns = g_halo_timeout_ns;
    return a;
}

// liveness bytes + status word to the peers (starts a new exchange)
int halo_send(tnb_complex *c, int word_raw_index, int word_flag_index, cudaStream_t s)
{
    c->halo.seq += 1;
    const HaloArgs a = halo_args(c, word_raw_index, word_flag_index);
    k_halo_send<<<3, 256, 0, s>>>(a);
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}
// waits (on the device) for the peers' messages of the current exchange
int halo_recv(tnb_complex *c, cudaStream_t s)
{
    const HaloArgs a = halo_args(c, C_RAW, C_FLAG);
    k_halo_recv<<<1, 256, 0, s>>>(a);
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}
const unsigned char *halo_in(const tnb_complex *c, int side)
{
    const tnb_halo &h = c->halo;
    return halo_inbox(h.boxes[h.rank], h.payload, side, (int)(h.seq & 1u)) + kHaloHeader;
}

__global__ void k_halo_apply(const unsigned char *__restrict__ tag, const int *__restrict__ slot, int V, int cap, int has_lower,
                             int has_upper, const unsigned char *__restrict__ in0, const unsigned char *__restrict__ in1,
                             int *__restrict__ used)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        const int t = tag[v];
        if ((t & 1) && has_lower && slot[v] < cap && in0[slot[v]]) used[v] = 1;
        if ((t & 2) && has_upper && slot[v] < cap && in1[slot[v]]) used[v] = 1;
    }
}

// The same exchange for the surface skeleton (faces.cu): liveness of the current complex's
// shared-plane vertices out, then (halo_merge_used) the neighbours' flags in.
int halo_publish_used(tnb_complex *c, int64_t V, const int *used, cudaStream_t s)
{
    tnb_halo &h = c->halo;
    int rc;
    for (int side = 0; side < 2; ++side) {
        const TagCount tc{c->tag[c->vcur].p, 1 << side};
        if ((rc = compact(V, tc, StageEmit{h.slot.p, h.stage[side].p, used, (int)h.payload}, c->block_sums.p, h.stage_count.p + side, s)))
            return rc;
    }
    return halo_send(c, C_RAW, C_FLAG, s);
}
int halo_merge_used(tnb_complex *c, int64_t V, int *used, cudaStream_t s)
{
    int rc = halo_recv(c, s);
    if (rc) return rc;
    if (V > 0) {
        k_halo_apply<<<grid_for(V, 256), 256, 0, s>>>(c->tag[c->vcur].p, c->halo.slot.p, (int)V, (int)c->halo.payload, c->halo.tag_lower,
                                                      c->halo.tag_upper, halo_in(c, 0), halo_in(c, 1), used);
        TNB_LAUNCH_CHECK();
    }
    return TNB_OK;
}

static void fill_step_args(const tnb_net *net, tnb_complex *c, float eps, StepArgs &sa)
{
    memset(&sa, 0, sizeof(sa));
    sa.R = net->meta.R; sa.eps = eps;
    sa.use_cross = (eps == net->meta.eps && net->meta.R <= 64) ? 1 : 0;
    sa.Vcap = (int)c->Vcap; sa.Ecap = (int)c->Ecap; sa.dim = c->cell_dim;
    for (int k = 0; k < 2; ++k) {
        sa.edges[k] = c->edges[k].p; sa.vert[k] = c->vert[k].p; sa.out[k] = c->out[k].p; sa.sig[k] = c->sig[k].p;
        sa.used[k] = c->used[k].p; sa.tag[k] = c->tag[k].p;
    }
    sa.bmask = c->bmask.p; sa.split_list = c->split_list.p; sa.cand = c->cand.p; sa.pcount = c->pcount.p;
    sa.poff = c->poff.p; sa.pcache = c->pcache.p; sa.gd = c->gd.p; sa.next = c->next.p; sa.remap = c->remap.p;
    sa.block_sums = c->block_sums.p; sa.cnt = c->counters.p; sa.head = c->head.p; sa.bytes = c->bytes.p;
    sa.cslot = c->cslot.p; sa.mask_e = c->scan_mask.p; sa.mask_v = c->scan_mask.p + scan_mask_vertex_offset(c->Ecap);
}

static const bool g_phase_trace = std::getenv("TNB_PHASE_TRACE") != nullptr;
static DevBuf<long long> g_phase_dbg;
static int phase_trace_begin(StepArgs &sa, cudaStream_t s)
{
    if (!g_phase_trace) return TNB_OK;
    TNB_CUDA(g_phase_dbg.reserve(64));
    TNB_CUDA(cudaMemsetAsync(g_phase_dbg.p, 0, 64 * sizeof(long long), s));
    sa.dbg = g_phase_dbg.p;
    return TNB_OK;
}
static int phase_trace_print(const char *what, int idx, cudaStream_t s)
{
    if (!g_phase_trace) return TNB_OK;
    long long h[64];
    TNB_CUDA(cudaMemcpyAsync(h, g_phase_dbg.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    TNB_CUDA(cudaStreamSynchronize(s));
    fprintf(stderr, "phase-trace %s %d front:", what, idx);
    for (int k = 1; k < 16; ++k) if (h[k]) fprintf(stderr, " %d:%.1f", k, h[k] * 1e-3);
    fprintf(stderr, " | back:");
    for (int k = 17; k < 32; ++k) if (h[k]) fprintf(stderr, " %d:%.1f", k, h[k] * 1e-3);
    fprintf(stderr, "\n");
    return TNB_OK;
}

// ---- all hyperplanes in one cluster launch ---------------------------------------------------------
// largest cluster (CTAs) k_steps_cluster can run with on this device: 16 (non-portable), else 8; 0 = none
template <class K>
static int cluster_ctas(K kernel)
{
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) cudaGetLastError();
    static const int want = std::getenv("TNB_CLUSTER_CTAS") ? std::atoi(std::getenv("TNB_CLUSTER_CTAS")) : 16;  // 8: twice as many objects side by side
    for (int cs : {16, 8}) {
        if (cs > want) continue;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs);
        cfg.blockDim = dim3(kClusterThreads);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) == cudaSuccess && n >= 1) return cs;
        cudaGetLastError();
    }
    return 0;
}
static bool g_long_lists = std::getenv("TNB_NO_LONG_LISTS") == nullptr;  // A/B switch: long partner lists by warps
static bool g_fused_curve = std::getenv("TNB_NO_FUSED_CURVE") == nullptr;  // A/B switch: curve path step by step
thread_local int64_t t_cluster_max_items = -1;  // tnb_subpoly_batch: its workers' complexes take the cluster form (>= 0 overrides the global)
static int64_t g_cluster_max_items = std::getenv("TNB_CLUSTER_MAX_ITEMS") ? std::atoll(std::getenv("TNB_CLUSTER_MAX_ITEMS")) : 0;

// How can the hyperplanes of this complex run?  0 = one step at a time (multi-launch kernels with
// full-size grids: large complexes; the curve path; slab sharding has its own kernels),
// 1 = persistent cooperative grid, 2 = persistent single cluster (opt-in: tnb_set_cluster_max_items)
int steps_mode(const tnb_net *net, const tnb_complex *c, bool planar)
{
    (void)net;
    if (c->halo.enabled || !g_fused_steps || c->E <= 0) return 0;
    if (!planar && !g_fused_curve) return 0;
    const int64_t items = c->E + c->V;  // may be stale upper bounds: good enough for this choice
    const int64_t cluster_max = t_cluster_max_items >= 0 ? t_cluster_max_items : g_cluster_max_items;
    if (planar && cluster_max > 0 && items <= cluster_max) return 2;
    return items <= g_fused_max_items ? 1 : 0;
}

// lh[2*i], lh[2*i+1] = (layer, neuron) of step i, as tnb_subpoly_step takes them.  One launch; the
// host does not wait: sizes, buffer parity, capacity and error bits stay in the counter block.
// TNB_ERR_UNSUPPORTED = this device / step list cannot run that way (the caller goes step by step).
int steps_persistent_impl(const tnb_net *net, tnb_complex *c, const int32_t *lh, int n_steps, float eps, int mode, bool planar, cudaStream_t s)
{
    if (!planar && mode != 1) return TNB_ERR_UNSUPPORTED;
    const NetMeta &m = net->meta;
    const int H = m.H, R = m.R;
    if (n_steps <= 0) return TNB_OK;
    if (n_steps > kMaxStepList || R > 256) return TNB_ERR_UNSUPPORTED;
    static int cs_ref = -1, cs_any = -1, gb_ref = -1, gb_any = -1, g2_ref = -1, g2_any = -1, gc_ref = -1, gc_any = -1;
    if (cs_ref < 0) {
        gc_ref = coop_blocks(k_steps_grid_curve<CfgRef>);
        gc_any = coop_blocks(k_steps_grid_curve<CfgAny>);
        cs_ref = cluster_ctas(k_steps_cluster<CfgRef>);
        cs_any = cluster_ctas(k_steps_cluster<CfgAny>);
        gb_ref = coop_blocks(k_steps_grid<CfgRef>);
        gb_any = coop_blocks(k_steps_grid<CfgAny>);
        g2_ref = coop_blocks(k_steps_grid2<CfgRef>);
        g2_any = coop_blocks(k_steps_grid2<CfgAny>);
    }
    static const int64_t wide_from = std::getenv("TNB_WIDE_FROM_ITEMS") ? std::atoll(std::getenv("TNB_WIDE_FROM_ITEMS")) : 150000;  // medium model: steps 0.68 -> 0.58 ms with two CTAs per SM
    const bool wide = planar && mode == 1 && c->E + c->V > wide_from && (net->fixed_cfg ? g2_ref : g2_any) >= 2 * kSMs;
    static const int env_blocks = std::getenv("TNB_STEP_BLOCKS") ? std::atoi(std::getenv("TNB_STEP_BLOCKS")) : 0;  // tuning knob
    // one CTA per SM keeps the grid barrier cheap
    const int blocks = mode == 2 ? (net->fixed_cfg ? cs_ref : cs_any)
                       : wide    ? 2 * kSMs
                       : !planar ? std::min(env_blocks > 0 ? env_blocks : kSMs, net->fixed_cfg ? gc_ref : gc_any)
                                 : std::min(env_blocks > 0 ? env_blocks : kSMs, net->fixed_cfg ? gb_ref : gb_any);
    if (blocks <= 0 || blocks > kScanMaxBlocks) return TNB_ERR_UNSUPPORTED;
    StepList list;
    memset(&list, 0, sizeof(list));
    list.n = n_steps;
    bool prunes = false;
    for (int i = 0; i < n_steps; ++i) {
        const int l = lh[2 * i], h = lh[2 * i + 1], idx = l * H + h;
        if (l < 0 || h < 0 || h > H || idx >= R) { set_error("tnb_subpoly_steps: (l,h) out of range"); return TNB_ERR_INVALID; }
        list.idx[i] = (unsigned char)idx;
        list.prune[i] = h < H ? 1 : 0;
        prunes = prunes || h < H;
    }
    if (c->bucket_mode == 2) {  // counts of an abandoned multi-launch step
        TNB_CUDA(cudaMemsetAsync(c->head.p, 0, (size_t)c->n_cells * sizeof(unsigned long long), s));
        c->stamp = 0;
    }
    c->bucket_mode = 1;
    if (c->stamp > 0xffffffffu - (uint32_t)n_steps - 1u) {  // the generation stamps of this launch would wrap
        TNB_CUDA(cudaMemsetAsync(c->head.p, 0, (size_t)c->n_cells * sizeof(unsigned long long), s));
        c->stamp = 0;
    }
    list.stamp0 = c->stamp + 1;
    c->stamp += (uint32_t)n_steps;
    StepArgs sa;
    fill_step_args(net, c, eps, sa);
    int rc;
    if ((rc = phase_trace_begin(sa, s))) return rc;
    prof_begin(TNB_PROF_STEPS, s);
    if (mode == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(blocks);
        cfg.blockDim = dim3(kClusterThreads);
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = blocks; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (net->fixed_cfg) TNB_CUDA(cudaLaunchKernelEx(&cfg, k_steps_cluster<CfgRef>, m, sa, list));
        else TNB_CUDA(cudaLaunchKernelEx(&cfg, k_steps_cluster<CfgAny>, m, sa, list));
    } else {
        void *params[] = {(void *)&m, (void *)&sa, (void *)&list};
        const void *kern = !planar ? (net->fixed_cfg ? (const void *)k_steps_grid_curve<CfgRef> : (const void *)k_steps_grid_curve<CfgAny>)
                           : wide  ? (net->fixed_cfg ? (const void *)k_steps_grid2<CfgRef> : (const void *)k_steps_grid2<CfgAny>)
                                : (net->fixed_cfg ? (const void *)k_steps_grid<CfgRef> : (const void *)k_steps_grid<CfgAny>);
        TNB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(kScanThreads), params, 0, s));
    }
    count_launch();
    prof_end(TNB_PROF_STEPS, s, 0);
    c->counts_stale = true;
    c->cross_stale = true;
    if (prunes) c->maybe_dead = true;
    return phase_trace_print(mode == 2 ? "cluster steps" : "grid steps", n_steps, s);
}

// part 0: the whole step; 1: up to and including the send of the slab exchange; 2: from its receive on
static int step_impl(const tnb_net *net, tnb_complex *c, int l, int h, float eps, bool planar, cudaStream_t s, int part = 0)
{
    const NetMeta &m = net->meta;
    const int H = m.H, R = m.R;
    const int idx = l * H + h;
    if (l < 0 || h < 0 || h > H || idx >= R) { set_error("tnb_subpoly_step: (l,h) out of range"); return TNB_ERR_INVALID; }
    int *cnt = c->counters.p;
    int rc;
    // c->V / c->E are upper bounds here when counts_stale (the exact values are in cnt[C_V], cnt[C_E])
    const bool halo = c->halo.enabled;
    if (c->E == 0 && !halo) return TNB_OK;
    if (halo && !planar) { set_error("slab-sharded extraction supports the planar path only"); return TNB_ERR_UNSUPPORTED; }
    if (!halo && part != 0) { set_error("tnb_subpoly_step_part: the complex has no slab exchange configured"); return TNB_ERR_INVALID; }
    const uint64_t colmask = (1ull << idx) - 1ull;

    // planar path on small complexes: both halves as one cooperative kernel each, fully
    // device-driven (sizes, buffer parity, capacity and error flags live in the counter block):
    // nothing to wait for here, the host syncs when somebody asks for the sizes or the data.
    static int front_blocks_ref = -1, front_blocks_any = -1, back_blocks = -1;
    if (front_blocks_ref < 0) {
        front_blocks_ref = coop_blocks(k_step_front<CfgRef>);
        front_blocks_any = coop_blocks(k_step_front<CfgAny>);
        back_blocks = coop_blocks(k_step_back);
    }
    // One CTA per SM keeps grid.sync() cheap.  Large complexes keep the multi-launch path with
    // full-size grids (c->V / c->E may be stale upper bounds here: good enough for this choice).
    static const int env_blocks = std::getenv("TNB_STEP_BLOCKS") ? std::atoi(std::getenv("TNB_STEP_BLOCKS")) : 0;  // tuning knob
    const int sm_blocks = std::min(env_blocks > 0 ? env_blocks : kSMs, std::min(front_blocks_ref, back_blocks));
    const int front_blocks = std::min(net->fixed_cfg ? front_blocks_ref : front_blocks_any, sm_blocks);
    const bool fused = planar && front_blocks > 0 && back_blocks > 0 && (halo || (g_fused_steps && c->E + c->V <= g_fused_max_items));
    if (halo && !fused) { set_error("slab-sharded extraction needs the cooperative step kernels"); return TNB_ERR_UNSUPPORTED; }
    if (!halo) {  // a small complex: the persistent step kernel, with a list of one
        const int mode = steps_mode(net, c, planar);
        if (mode) {
            const int32_t one[2] = {l, h};
            rc = steps_persistent_impl(net, c, one, 1, eps, mode, planar, s);
            if (rc != TNB_ERR_UNSUPPORTED) return rc;
        }
    }
    if (fused && halo) {
        if (part != 2) {
            if (c->bucket_mode == 2) {
                TNB_CUDA(cudaMemsetAsync(c->head.p, 0, (size_t)c->n_cells * sizeof(unsigned long long), s));
                c->stamp = 0;
            }
            c->bucket_mode = 1;
            c->stamp += 1;
            if (c->stamp == 0) {
                TNB_CUDA(cudaMemsetAsync(c->head.p, 0, (size_t)c->n_cells * sizeof(unsigned long long), s));
                c->stamp = 1;
            }
        }
        StepArgs sa;
        fill_step_args(net, c, eps, sa);
        const StepVar sv = step_var(idx, R, h < H ? 1 : 0, c->stamp);
        if ((rc = phase_trace_begin(sa, s))) return rc;
        sa.halo = halo ? 1 : 0;
        void *fparams[] = {(void *)&m, (void *)&sa, (void *)&sv};
        void *bparams[] = {(void *)&sa, (void *)&sv};
        const dim3 bgrid(std::min(back_blocks, sm_blocks));
        if (halo) {
            sa.has_lower = c->halo.tag_lower; sa.has_upper = c->halo.tag_upper;
            sa.hslot = c->halo.slot.p; sa.stage_count = c->halo.stage_count.p;
            sa.stage[0] = c->halo.stage[0].p; sa.stage[1] = c->halo.stage[1].p;
            sa.stage_cap = (int)c->halo.payload;
        }
        if (part != 2) {
            prof_begin(TNB_PROF_NEW_VERTICES, s);
            if (net->fixed_cfg)
                TNB_CUDA(cudaLaunchCooperativeKernel((void *)k_step_front<CfgRef>, dim3(front_blocks), dim3(kScanThreads), fparams, 0, s));
            else
                TNB_CUDA(cudaLaunchCooperativeKernel((void *)k_step_front<CfgAny>, dim3(front_blocks), dim3(kScanThreads), fparams, 0, s));
            count_launch();
            prof_end(TNB_PROF_NEW_VERTICES, s, 0);
            sa.part = halo ? 1 : 0;
            prof_begin(TNB_PROF_PAIRS, s);
            TNB_CUDA(cudaLaunchCooperativeKernel((void *)k_step_back, bgrid, dim3(kScanThreads), bparams, 0, s));
            count_launch();
            prof_end(TNB_PROF_PAIRS, s, 0);
            if (halo && (rc = halo_send(c, C_RAW, C_FLAG, s))) return rc;
        }
        if (halo && part != 1) {
            if ((rc = halo_recv(c, s))) return rc;
            sa.part = 2;
            sa.in[0] = halo_in(c, 0); sa.in[1] = halo_in(c, 1);
            prof_begin(TNB_PROF_PAIRS, s);
            TNB_CUDA(cudaLaunchCooperativeKernel((void *)k_step_back, bgrid, dim3(kScanThreads), bparams, 0, s));
            count_launch();
            prof_end(TNB_PROF_PAIRS, s, 0);
        }
        c->counts_stale = true;  // sizes and buffer parity are on the device until the next sync
        if (sv.do_prune) c->maybe_dead = true;
        return phase_trace_print("step", idx, s);
    }
    // multi-launch path: needs the host's view of sizes and buffer parity to be current
    if ((rc = complex_sync_counts(c, s))) return rc;
    if (c->E == 0) return TNB_OK;
    // dead rows cost a little in every per-vertex pass: squeeze them out once the arrays are three quarters full
    // the pruning pass of the last step that changed the complex saw no edge across this plane
    if (eps == m.eps && idx < 64 && !((c->cross >> idx) & 1ull)) return TNB_OK;
    if (c->maybe_dead && (size_t)c->V * 4 > c->Vcap * 3 && (rc = complex_compact(c, s))) return rc;
    cnt = c->counters.p;

    for (int attempt = 0;; ++attempt) {
      {
        k_clear_step_counters<<<1, 32, 0, s>>>(cnt);
        TNB_LAUNCH_CHECK();
        // 1. edges the hyperplane crosses
        SplitCount sc{c->cedges(), c->cout_(), R, idx, eps};
        if ((rc = compact_masked(c->E, sc, ListEmit{c->split_list.p}, c->block_sums.p, c->scan_mask.p, cnt + C_RAW, s, cnt + C_E))) return rc;
        if (attempt == 0) {  // most hyperplanes of a fitted network cross nothing: learn it now (subpoly.py:110-111)
            if ((rc = read_counters(c, s))) return rc;
            if (c->h_counters[C_RAW] == 0) return TNB_OK;
        }
        // 2. new vertices, their network rows, rewired edges (all sized on the device)
        if (!planar) {
            const int o = c->vcur ^ 1;  // idle half of the ping-pong arrays = temp slots
            unsigned g = grid_for(c->E, kThreads);
            k_gd_reset<<<1, 32, 0, s>>>(c->gd.p);
            prof_begin(TNB_PROF_NEW_VERTICES, s);
            if (net->fixed_cfg)
                k_new_vertices_curve<CfgRef><<<g, kThreads, 0, s>>>(m, idx, eps, (int)c->Vcap, (int)c->Ecap, c->split_list.p, c->cedges(), c->cvert(), c->cout_(), c->csig(), c->vert[o].p, c->out[o].p, c->bmask.p, c->pcount.p, cnt, c->gd.p);
            else
                k_new_vertices_curve<CfgAny><<<g, kThreads, 0, s>>>(m, idx, eps, (int)c->Vcap, (int)c->Ecap, c->split_list.p, c->cedges(), c->cvert(), c->cout_(), c->csig(), c->vert[o].p, c->out[o].p, c->bmask.p, c->pcount.p, cnt, c->gd.p);
            TNB_LAUNCH_CHECK();
            // the gradient-descent repair of the intersections that are off their planes (both return at once if there are none)
            if (net->fixed_cfg) {
                k_gd_note<CfgRef><<<kSMs, 64, 0, s>>>(m, idx, eps, c->split_list.p, c->cedges(), c->cvert(), cnt, c->gd.p);
                k_gd_walk<CfgRef><<<kSMs, 64, 0, s>>>(m, idx, eps, c->split_list.p, c->cedges(), c->cvert(), c->vert[o].p, c->out[o].p, c->bmask.p, cnt, c->gd.p);
            } else {
                k_gd_note<CfgAny><<<kSMs, 64, 0, s>>>(m, idx, eps, c->split_list.p, c->cedges(), c->cvert(), cnt, c->gd.p);
                k_gd_walk<CfgAny><<<kSMs, 64, 0, s>>>(m, idx, eps, c->split_list.p, c->cedges(), c->cvert(), c->vert[o].p, c->out[o].p, c->bmask.p, cnt, c->gd.p);
            }
            TNB_LAUNCH_CHECK();
            prof_end(TNB_PROF_NEW_VERTICES, s, 0);
            k_strict_keep<<<g, kThreads, 0, s>>>(R, idx, eps, c->out[o].p, c->bmask.p, c->pcount.p, cnt);
            TNB_LAUNCH_CHECK();
            CurveCommitEmit ce{nullptr, c->split_list.p, c->cedges(), c->vert[o].p, c->out[o].p, c->cvert(), c->cout_(), c->csig(),
                               m.marks, cnt, R, m.n_marks, m.eps, m.pre_scale, m.pre_2s, m.pre_inv, m.pre_pow2, c->tag[c->vcur].p};
            if ((rc = compact(c->E, KeepFlagCount{c->pcount.p, cnt + C_OVERFLOW}, ce, c->block_sums.p, cnt + C_SPLIT, s, cnt + C_RAW))) return rc;
        } else {
            unsigned g = grid_for(c->E, kThreads);
            prof_begin(TNB_PROF_NEW_VERTICES, s);
            const size_t row_tile = (size_t)kThreads * R * sizeof(float);
            if (net->fixed_cfg)
                k_new_vertices<CfgRef><<<g, kThreads, row_tile, s>>>(m, idx, eps, (int)c->Vcap, (int)c->Ecap, c->split_list.p, c->cedges(), c->cvert(), c->cout_(), c->csig(), c->bmask.p, cnt, c->tag[c->vcur].p);
            else
                k_new_vertices<CfgAny><<<g, kThreads, row_tile, s>>>(m, idx, eps, (int)c->Vcap, (int)c->Ecap, c->split_list.p, c->cedges(), c->cvert(), c->cout_(), c->csig(), c->bmask.p, cnt, c->tag[c->vcur].p);
            TNB_LAUNCH_CHECK();
            prof_end(TNB_PROF_NEW_VERTICES, s, 0);
            k_finalize_new<<<g, kThreads, 0, s>>>(m, c->cvert(), c->cout_(), c->csig(), c->bmask.p, cnt);
            TNB_LAUNCH_CHECK();
        }
        // 3. candidates for connecting edges: old vertices on the plane, then the new ones
        HitCount hc{c->cout_(), c->calive(), R, idx, eps};
        if ((rc = compact_masked(c->V, hc, ListEmit{c->cand.p}, c->block_sums.p, c->scan_mask.p, cnt + C_HIT, s, cnt + C_V))) return rc;
        const int64_t cand_ub = std::min<int64_t>(c->V + c->E, (int64_t)c->Vcap);
        k_fill_new_cands<<<grid_for(c->E, 256), 256, 0, s>>>(c->cand.p, cnt);
        TNB_LAUNCH_CHECK();
        // contiguous cell segments (cells.cuh): the cell grid is all-zero between uses
        if (c->bucket_mode != 0) {  // the linked-list form left its generation stamps there, or a failed step its counts
            TNB_CUDA(cudaMemsetAsync(c->head.p, 0, (size_t)c->n_cells * sizeof(unsigned long long), s));
            c->stamp = 0;
        }
        c->bucket_mode = 2;
        if ((rc = cells_build(1, c->cand.p, cnt + C_CAND, cand_ub, c->csig(), (int2 *)c->head.p, c->cslot.p, c->next.p, cnt + C_RECS, c->cell_dim, s))) return rc;
        prof_begin(TNB_PROF_PAIRS, s);
        k_pair_count_seg<<<kSMs * 8, 256, 0, s>>>(c->cand.p, cnt, c->csig(), (const int2 *)c->head.p, c->next.p, c->cell_dim, colmask, c->pcount.p, c->pcache.p, c->remap.p);
        TNB_LAUNCH_CHECK();
        prof_end(TNB_PROF_PAIRS, s, 0);
        if ((rc = compact(cand_ub, ArrayCount{c->pcount.p}, OffsetEmit{c->poff.p}, c->block_sums.p, cnt + C_PAIRS, s, cnt + C_CAND))) return rc;
      }
        // ---- the one host sync of the step ----
        if ((rc = read_counters(c, s))) return rc;
        c->V = c->h_counters[C_V];
        c->E = c->h_counters[C_E];
        c->counts_stale = false;
        if (!c->h_counters[C_OVERFLOW]) break;
        if (attempt > 0) { set_error("work buffers did not grow"); return TNB_ERR_CAPACITY; }
        const int S_need = c->h_counters[C_RAW];
        if ((rc = complex_reserve(c, c->V + S_need, c->E + S_need, s))) return rc;
    }
    if (c->h_counters[C_ERR] & kErrNoPlane) {
        set_error("curve path: a non-axis-aligned edge lies on no earlier plane (the reference exits here, subpoly.py:140-148)");
        return TNB_ERR_INVALID;
    }
    if (c->h_counters[C_ERR] & kErrGradientDescent) {
        set_error("curve path: an intersection is still off its planes after the gradient-descent repair (subpoly_debug.py:121-165); the reference ends the extraction here (subpoly.py:172-174)");
        return TNB_ERR_UNSUPPORTED;
    }
    if (c->h_counters[C_RAW] == 0) return TNB_OK;  // subpoly.py:110-111
    const int S = c->h_counters[C_SPLIT];
    const int V0 = (int)c->V, E0 = (int)c->E;
    {   // algorithmic bytes of what ran before the sync (sizes are known only now)
        const int64_t Hc = c->h_counters[C_HIT], newv = (int64_t)S * (8 + 4 + 2 * (12 + 4 + 16) + 12 + 4 * R + 8 + 16);
        prof_add(TNB_PROF_NEW_VERTICES, S, newv + (int64_t)net->table.cap * 8);
        prof_add(TNB_PROF_PAIRS, Hc + S, (Hc + S) * (24 + 4));
    }
    const int Hn = c->h_counters[C_HIT], P = c->h_counters[C_PAIRS];
    const int n_cand = Hn + S;
    if ((rc = complex_reserve(c, (size_t)V0 + S, (size_t)E0 + S + P, s))) return rc;
    cnt = c->counters.p;
    if (P > 0) {
        prof_begin(TNB_PROF_PAIRS, s);
        k_pair_copy<<<grid_for((int64_t)n_cand * kCachedPartners, 256), 256, 0, s>>>(c->cand.p, cnt, c->pcount.p, c->poff.p, c->pcache.p, c->cedges() + E0 + S);
        TNB_LAUNCH_CHECK();
        const int n_long = c->h_counters[C_LONG];
        if (n_long > 0) {
            k_pair_write_long_seg<<<(unsigned)std::min<int64_t>((n_long + kSortWarps - 1) / kSortWarps, kSMs * 4), kSortWarps * 32, 0, s>>>(
                c->remap.p, cnt, c->cand.p, c->csig(), (const int2 *)c->head.p, c->next.p, c->cell_dim, colmask, c->pcount.p, c->poff.p, c->cedges() + E0 + S);
            TNB_LAUNCH_CHECK();
        }
        prof_end(TNB_PROF_PAIRS, s, n_cand, (int64_t)n_cand * 28 + (int64_t)P * 8);
    }
    if ((rc = cells_clear(1, nullptr, n_cand, (int2 *)c->head.p, c->cslot.p, s))) return rc;
    c->bucket_mode = 0;
    c->V = V0 + S;
    c->E = (int64_t)E0 + S + P;

    // 4. pruning (not for the output neuron, subpoly.py:253): edges are compacted, vertices only
    //    marked (complex.cuh); the edge count stays on the device
    if (h < H) {
        const uint64_t futmask = ~colmask & (R >= 64 ? ~0ull : ((1ull << R) - 1ull));
        int *used = c->used[c->acur ^ 1].p;
        TNB_CUDA(cudaMemsetAsync(used, 0, (size_t)c->V * sizeof(int), s));
        KeepCount kc{c->cedges(), c->csig(), futmask};
        int2 *dst = c->edges[c->ecur ^ 1].p;
        {   // the compaction of scan.cuh's compact(), its count pass also collecting the crossing mask
            const int64_t blocks = std::min<int64_t>((c->E + kScanThreads - 1) / kScanThreads, kScanMaxBlocks);
            TNB_CUDA(cudaMemsetAsync(cnt + C_CROSS, 0, sizeof(uint64_t), s));
            k_keep_count_cross<<<(unsigned)blocks, kScanThreads, 0, s>>>(c->E, nullptr, c->cedges(), c->csig(), futmask, c->block_sums.p, cnt);
            TNB_LAUNCH_CHECK();
            k_scan_write<<<(unsigned)blocks, kScanThreads, 0, s>>>(c->E, nullptr, kc, KeepEmit{c->cedges(), dst, used}, c->block_sums.p, cnt + C_E);
            TNB_LAUNCH_CHECK();
            c->cross_stale = true;
        }
        c->ecur ^= 1;
        c->acur ^= 1;
        k_set_counts<<<1, 1, 0, s>>>(cnt, (int)c->V, -1, c->vcur, c->ecur, c->acur);  // the fused kernels read sizes and parity on the device
        TNB_LAUNCH_CHECK();
        c->maybe_dead = true;
        c->counts_stale = true;  // c->E is an upper bound until the next sync
    } else {
        TNB_CUDA(cudaMemsetAsync(c->calive() + V0, 1, (size_t)S * sizeof(int), s));  // the new vertices are alive
        TNB_CUDA(cudaMemsetAsync(cnt + C_CROSS, 0xff, sizeof(uint64_t), s));          // no pruning pass: crossings unknown
        c->cross = ~0ull;
        k_set_counts<<<1, 1, 0, s>>>(cnt, (int)c->V, (int)c->E, c->vcur, c->ecur, c->acur);
        TNB_LAUNCH_CHECK();
    }
    return TNB_OK;
}

// ================================================================================================
// Device-driven step stream: the hyperplanes of a LARGE complex (planar path)
// ================================================================================================
// A large complex keeps full-size grids per phase (a persistent 148-CTA grid is too thin for millions of
// edges), but nothing a phase needs comes from the host any more: the hyperplane of the step (C_IDX), every
// size, which half of the ping-pong arrays is current and the capacity checks live in the counter block,
// and the step list itself is a kernel parameter.  One SEQUENCE of 16 launches runs one hyperplane that
// crosses something; its last kernel (k_sd_finish) commits the step, picks the next hyperplane whose bit
// is set in the crossing mask (the 25 of 33 planes of a fitted network that cross nothing cost nothing) and
// reports through mapped pinned memory.  The host enqueues one sequence ahead of the one it has a report
// for, so the stream never drains and the host never synchronises: before, every crossing hyperplane cost
// three cudaStreamSynchronize round trips (crossing mask, crossed-edge count, sizing of the edge array).
// Work arrays do not grow on this path: an overflow raises the sticky capacity bit and tnb_subpoly
// repeats the extraction with twice the head-room, exactly as for the persistent kernels.
struct PreMeta {  // what finalising a new vertex needs of NetMeta (5 KB as a kernel parameter: not for a housekeeping kernel)
    int R, n_marks, pre_pow2;
    float eps, pre_scale, pre_2s, pre_inv;
    const float *marks;
};
static PreMeta pre_meta(const NetMeta &m)
{
    PreMeta p;
    p.R = m.R; p.n_marks = m.n_marks; p.pre_pow2 = m.pre_pow2;
    p.eps = m.eps; p.pre_scale = m.pre_scale; p.pre_2s = m.pre_2s; p.pre_inv = m.pre_inv;
    p.marks = m.marks;
    return p;
}
// finalize_item without NetMeta (same operations)
__device__ __forceinline__ void finalize_item_slim(const PreMeta &n, const float *vert, float *out, uint64_t *sig, const uint64_t *bmask, int V, int k)
{
    const int64_t v = (int64_t)V + k;
    float *row = out + v * n.R;
    for (uint64_t m = bmask[k]; m; m &= m - 1) row[__ffsll((long long)m) - 1] = 0.0f;
    uint64_t pos, neg;
    pack_signs(row, n.R, n.eps, pos, neg);
    uint64_t g = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float t = vert[3 * v + d] + n.pre_scale;
        const float xp = n.pre_pow2 ? t * n.pre_inv : __fdiv_rn(t, n.pre_2s);
        const int off = lower_bound(n.marks, n.n_marks, xp + n.eps) - 1;
        const float mk = n.marks[off < 0 ? off + n.n_marks : off];
        g |= (uint64_t)(uint32_t)(off + 1) << (20 * d);
        g |= (fabsf(mk - xp) > n.eps ? 1ull : 0ull) << (60 + d);
    }
    sig[3 * v] = pos;
    sig[3 * v + 1] = neg;
    sig[3 * v + 2] = g;
}

// is there a step to run, and may it go on?  (uniform over the grid: these words are written by single-CTA kernels
// of earlier launches only)
__device__ __forceinline__ bool sd_active(const int *cnt) { return cnt[C_IDX] >= 0 && !cnt[C_STICKY]; }
// ... after the new vertices: did the plane cross something, and did everything fit?
__device__ __forceinline__ bool sd_crossed(const int *cnt) { return sd_active(cnt) && cnt[C_RAW] > 0 && !cnt[C_OVERFLOW]; }
// ... after the connecting-edge count: room for the edges?
__device__ __forceinline__ bool sd_fits(const StepArgs &a) { return (int64_t)a.cnt[C_E] + a.cnt[C_SPLIT] + a.cnt[C_PAIRS] <= a.Ecap; }

// The split and hit tests from the PACKED signs (the stream only runs with the eps the signs were packed with):
//   crossed  <=>  d0 * d1 < 0, |d0| > eps, |d1| > eps  <=>  one end's positive bit and the other end's negative bit
//   (v > eps and w < -eps: the product is below -eps^2, no underflow to -0), subpoly.py:104-105;
//   hit      <=>  |v| < eps: only a vertex with neither bit set (|v| <= eps) can pass, and only for those is the
//   float read (the boundary case |v| == eps is decided by the float, as in subpoly.py:233).
// The sign words of the whole complex are 17 MB and stay in L2 across the kernels of a step; the 132-byte rows the
// float tests gathered from are 92 MB (ncu: 115-177 MB of DRAM reads per launch of this kernel before).
struct SplitCountSig {
    const int2 *edges;
    const uint64_t *sig;
    int idx;
    __device__ __forceinline__ int operator()(int64_t e) const
    {
        const int2 ed = edges[e];
        const uint64_t pa = sig[3 * (int64_t)ed.x], na = sig[3 * (int64_t)ed.x + 1];
        const uint64_t pb = sig[3 * (int64_t)ed.y], nb = sig[3 * (int64_t)ed.y + 1];
        return (int)((((pa & nb) | (na & pb)) >> idx) & 1ull);
    }
};
struct HitCountSig {
    const float *out;
    const uint64_t *sig;
    const int *alive;
    int R, idx;
    float eps;
    __device__ __forceinline__ int operator()(int64_t v) const
    {
        const int al = alive[v];
        const uint64_t nz = sig[3 * v] | sig[3 * v + 1];
        if (!al || ((nz >> idx) & 1ull)) return 0;
        return fabsf(out[v * R + idx]) < eps ? 1 : 0;
    }
};
// split compaction over the edges (first half of the grid) and hit compaction over the vertices (second half)
__global__ void __launch_bounds__(kScanThreads) k_sd_count(const __grid_constant__ StepArgs a)
{
    pdl_wait();
    const int *cnt = a.cnt;
    if (!sd_active(cnt)) return;
    const int idx = cnt[C_IDX], nb = (int)gridDim.x >> 1, pv = cnt[C_VPAR];
    if ((int)blockIdx.x < nb)
        scan_count_mask_part(cnt[C_E], SplitCountSig{a.edges[cnt[C_EPAR]], a.sig[pv], idx}, a.block_sums, a.mask_e, (int)blockIdx.x, nb);
    else
        scan_count_mask_part(cnt[C_V], HitCountSig{a.out[pv], a.sig[pv], a.used[cnt[C_APAR]], a.R, idx, a.eps}, a.block_sums + nb, a.mask_v,
                             (int)blockIdx.x - nb, nb);
}
__global__ void __launch_bounds__(kScanThreads) k_sd_write(const __grid_constant__ StepArgs a)
{
    pdl_wait();
    int *cnt = a.cnt;
    if (!sd_active(cnt)) return;
    const int nb = (int)gridDim.x >> 1;
    if ((int)blockIdx.x < nb) scan_write_mask_part(cnt[C_E], a.mask_e, ListEmit{a.split_list}, a.block_sums, cnt + C_RAW, (int)blockIdx.x, nb);
    else scan_write_mask_part(cnt[C_V], a.mask_v, ListEmit{a.cand}, a.block_sums + nb, cnt + C_HIT, (int)blockIdx.x - nb, nb);
}
template <class C>
__global__ void __launch_bounds__(kThreads) k_sd_new_vertices(const __grid_constant__ NetMeta n, const __grid_constant__ StepArgs a)
{
    pdl_wait();
    int *cnt = a.cnt;
    if (!sd_active(cnt) || cnt[C_RAW] == 0) return;
    const int pv = cnt[C_VPAR];
    extern __shared__ float s_tile[];  // [kThreads][R]
    body_new_vertices<C, true>(n, cnt[C_IDX], a.eps, a.Vcap, a.Ecap, a.split_list, a.edges[cnt[C_EPAR]], a.vert[pv], a.out[pv], a.sig[pv], a.bmask, cnt, a.tag[pv], s_tile);
}
// candidate list (hit old vertices, then the new ones), the failover override when it fired, liveness array of the
// pruning pass cleared, crossing mask reset (k_sd_keep_count rebuilds it)
__global__ void __launch_bounds__(256) k_sd_cands(const __grid_constant__ StepArgs a, const PreMeta pm)
{
    pdl_wait();
    int *cnt = a.cnt;
    if (!sd_crossed(cnt)) return;
    const int Hn = cnt[C_HIT], S = cnt[C_SPLIT], V = cnt[C_V], flag = cnt[C_FLAG], prune = cnt[C_PRUNE], pv = cnt[C_VPAR], pa = cnt[C_APAR];
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int k = t0; k < S; k += stride) {
        a.cand[Hn + k] = V + k;
        if (flag) finalize_item_slim(pm, a.vert[pv], a.out[pv], a.sig[pv], a.bmask, V, k);
        if (!prune) a.used[pa][V + k] = 1;  // the output neuron (subpoly.py:253): the new vertices simply join
    }
    if (prune) {
        int *used = a.used[pa ^ 1];
        for (int v = t0; v < V + S; v += stride) used[v] = 0;
    }
    if (t0 == 0) {
        cnt[C_CAND] = Hn + S;
        *(unsigned long long *)(cnt + C_CROSS) = prune ? 0ull : ~0ull;
    }
}
__global__ void __launch_bounds__(256) k_sd_pair_copy(const __grid_constant__ StepArgs a)
{
    pdl_wait();
    const int *cnt = a.cnt;
    if (!sd_crossed(cnt) || !sd_fits(a) || cnt[C_PAIRS] == 0) return;
    int2 *edges_out = a.edges[cnt[C_EPAR]] + cnt[C_E] + cnt[C_SPLIT];
    // one thread per candidate: nearly every list has one to four entries (a thread per (candidate, slot) left
    // 30 of 32 threads with nothing to copy)
    const int n_cand = cnt[C_CAND];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cand; c += gridDim.x * blockDim.x) {
        const int pc = a.pcount[c];
        if (pc == 0 || pc > kCachedPartners) continue;
        const int va = a.cand[c], off = a.poff[c];
        const int *src = a.pcache + (int64_t)c * kCachedPartners;
        for (int i = 0; i < pc; ++i) edges_out[off + i] = make_int2(va, src[i]);
    }
}
__global__ void __launch_bounds__(kSortWarps * 32) k_sd_pair_long(const __grid_constant__ StepArgs a)
{
    pdl_wait();
    const int *cnt = a.cnt;
    if (!sd_crossed(cnt) || !sd_fits(a) || cnt[C_LONG] == 0) return;
    pair_write_long_seg_body(a.remap, cnt[C_LONG], a.cand, a.sig[cnt[C_VPAR]], (const int2 *)a.head, a.next, a.dim, (1ull << cnt[C_IDX]) - 1ull,
                             a.pcount, a.poff, a.edges[cnt[C_EPAR]] + cnt[C_E] + cnt[C_SPLIT]);
}
__device__ __forceinline__ uint64_t sd_futmask(int idx, int R) { return ~((1ull << idx) - 1ull) & (R >= 64 ? ~0ull : ((1ull << R) - 1ull)); }
__global__ void __launch_bounds__(kScanThreads) k_sd_keep_count(const __grid_constant__ StepArgs a)
{
    pdl_wait();
    int *cnt = a.cnt;
    if (!sd_crossed(cnt) || !sd_fits(a) || !cnt[C_PRUNE]) return;
    const int64_t En = (int64_t)cnt[C_E] + cnt[C_SPLIT] + cnt[C_PAIRS];
    keep_count_cross<kScanThreads>(En, a.edges[cnt[C_EPAR]], a.sig[cnt[C_VPAR]], sd_futmask(cnt[C_IDX], a.R), a.block_sums, cnt);
}
__global__ void __launch_bounds__(kScanThreads) k_sd_keep_write(const __grid_constant__ StepArgs a)
{
    pdl_wait();
    int *cnt = a.cnt;
    if (!sd_crossed(cnt) || !sd_fits(a) || !cnt[C_PRUNE]) return;
    const int64_t En = (int64_t)cnt[C_E] + cnt[C_SPLIT] + cnt[C_PAIRS];
    const int pe = cnt[C_EPAR];
    const int2 *edges = a.edges[pe];
    // (four consecutive items per thread and a 1024-item tile measured SLOWER here: 19.3 -> 23.7 us per launch)
    scan_write_body(En, KeepCount{edges, a.sig[cnt[C_VPAR]], sd_futmask(cnt[C_IDX], a.R)}, KeepEmit{edges, a.edges[pe ^ 1], a.used[cnt[C_APAR] ^ 1]},
                    a.block_sums, cnt + C_KEPT);
}
// Commits the step that just ran (first = 0), chooses the next hyperplane and reports to the host.
//   slot = report + 8 * (seq % kReportSlots): slot[1..5] = {another step follows, vertex slots, edges, position of the
//   chosen step in the list, sticky bits}, then slot[0] = seq (the host waits for that word).  A ring, because the
//   device may be one report ahead of the one the host is reading.
constexpr int kReportSlots = 4;
__global__ void k_sd_finish(const __grid_constant__ StepArgs a, const __grid_constant__ StepList list, int first, int seq, volatile int *report)
{
    pdl_wait();
    if (threadIdx.x != 0) return;
    report += 8 * (seq & (kReportSlots - 1));
    int *cnt = a.cnt;
    if (!first && sd_active(cnt)) {
        if (cnt[C_OVERFLOW]) cnt[C_STICKY] = kStickyCapacity;
        else if (cnt[C_RAW] > 0) {
            const int E0 = cnt[C_E], S = cnt[C_SPLIT], P = cnt[C_PAIRS], V0 = cnt[C_V], n_cand = cnt[C_CAND];
            const int64_t En = (int64_t)E0 + S + P;
            if (En > a.Ecap) cnt[C_STICKY] = kStickyCapacity;
            else {
                cnt[C_V] = V0 + S;
                if (cnt[C_PRUNE]) {
                    cnt[C_E] = cnt[C_KEPT];
                    cnt[C_EPAR] ^= 1;
                    cnt[C_APAR] ^= 1;
                } else {
                    cnt[C_E] = (int)En;
                }
                // compulsory bytes of the two TIMED kernel classes only (the scans and the pruning pass around them are
                // not inside the event pairs): a new vertex reads its edge, both ends' position / value / signs and
                // writes position, row, signs and two edges; a candidate reads its record and the records around it
                // (counted once: 32 B) and writes its count, offset and pairs
                a.bytes[0] += (unsigned long long)S * (8 + 4 + 2 * (12 + 4 + 16) + 12 + 4 * a.R + 24 + 16);
                a.bytes[1] += (unsigned long long)n_cand * (32 + 32 + 8) + (unsigned long long)P * 8;
                a.bytes[2] += (unsigned long long)S;
                a.bytes[3] += (unsigned long long)n_cand;
            }
        }
    }
    // the next hyperplane that crosses an edge (subpoly.py:110-111 for the others, decided from the crossing mask the
    // last pruning pass left: a clear bit means the float test fails for every edge)
    int i = first ? 0 : cnt[C_STEP];
    const unsigned long long cross = *(unsigned long long *)(cnt + C_CROSS);
    if (cnt[C_STICKY]) i = list.n;
    while (i < list.n && !((cross >> list.idx[i]) & 1ull)) ++i;
    const int more = i < list.n ? 1 : 0;
    cnt[C_IDX] = more ? (int)list.idx[i] : -1;
    cnt[C_PRUNE] = more ? (int)list.prune[i] : 0;
    cnt[C_STEP] = more ? i + 1 : list.n;
    for (int k = 0; k < C_V; ++k) cnt[k] = 0;
    cnt[C_LONG] = 0;
    cnt[C_RECS] = 0;
    report[1] = more;
    report[2] = cnt[C_V];
    report[3] = cnt[C_E];
    report[4] = i;
    report[5] = cnt[C_STICKY];
    __threadfence_system();
    report[0] = seq;
}

static bool g_stream_steps = std::getenv("TNB_NO_STREAM_STEPS") == nullptr;  // A/B switch: one hyperplane at a time with host syncs

// can the hyperplanes of this complex run as a device-driven stream?
static bool stream_ok(const tnb_net *net, const tnb_complex *c, float eps, bool planar)
{
    // eps >= 1e-15: the packed-sign form of the split test needs eps^2 to be a normal float (no underflow of d0 * d1)
    return g_stream_steps && planar && !c->halo.enabled && c->E > 0 && eps == net->meta.eps && eps >= 1e-15f && net->meta.R <= 64;
}

// Runs the steps lh[0 .. n_steps) (or, when the complex becomes small enough for the persistent kernel, a prefix
// of them) and returns how many list entries are done in *consumed.  The host does not wait for the last sequence.
static int steps_stream_impl(const tnb_net *net, tnb_complex *c, const int32_t *lh, int n_steps, float eps, cudaStream_t s, int *consumed)
{
    const NetMeta &m = net->meta;
    const int H = m.H, R = m.R;
    *consumed = 0;
    if (n_steps <= 0) return TNB_OK;
    if (n_steps > kMaxStepList) n_steps = kMaxStepList;  // the caller comes back for the rest
    StepList list;
    memset(&list, 0, sizeof(list));
    list.n = n_steps;
    for (int i = 0; i < n_steps; ++i) {
        const int l = lh[2 * i], h = lh[2 * i + 1], idx = l * H + h;
        if (l < 0 || h < 0 || h > H || idx >= R) { set_error("tnb_subpoly_steps: (l,h) out of range"); return TNB_ERR_INVALID; }
        list.idx[i] = (unsigned char)idx;
        list.prune[i] = h < H ? 1 : 0;
    }
    // mapped pinned report block, one per host thread; sequence numbers never repeat, so a late write of an
    // earlier call cannot be mistaken for a report of this one
    static thread_local int *report = nullptr, *report_dev = nullptr;
    static thread_local int seq_base = 0;
    if (!report) {
        TNB_CUDA(cudaHostAlloc((void **)&report, kReportSlots * 8 * sizeof(int), cudaHostAllocMapped));
        TNB_CUDA(cudaHostGetDevicePointer((void **)&report_dev, report, 0));
        memset(report, 0, kReportSlots * 8 * sizeof(int));
    }
    volatile int *rep = report;  // set by wait_for to the slot of the report it waited for
    if (c->bucket_mode != 0) {  // the cell grid must be all-zero for the contiguous segments (cells.cuh)
        TNB_CUDA(cudaMemsetAsync(c->head.p, 0, (size_t)c->n_cells * sizeof(unsigned long long), s));
        c->stamp = 0;
        c->bucket_mode = 0;
    }
    StepArgs sa;
    fill_step_args(net, c, eps, sa);
    const PreMeta pm = pre_meta(m);
    const size_t row_tile = (size_t)kThreads * R * sizeof(float);
    const int64_t cand_ub = (int64_t)c->Vcap;
    const unsigned nb2 = 2 * kScanMaxBlocks;
    auto wait_for = [&](int seq) -> int {
        rep = report + 8 * (seq & (kReportSlots - 1));
        for (long spins = 0; rep[0] - seq < 0; ++spins) {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
            if ((spins & 0x3ff) == 0x3ff) std::this_thread::yield();  // several extractions in flight on one box share its cores
            if ((spins & 0xfffff) == 0xfffff) {  // now and then: did the stream die?
                cudaError_t e = cudaStreamQuery(s);
                if (e != cudaSuccess && e != cudaErrorNotReady) return cuda_fail(e, "device-driven step stream", __FILE__, __LINE__);
                if (e == cudaSuccess && rep[0] - seq < 0) { set_error("device-driven step stream: the stream drained without a report"); return TNB_ERR_CUDA; }
            }
        }
        return TNB_OK;
    };
    auto enqueue_sequence = [&](int seq) -> int {
        int rc;
        // every kernel of the sequence is launched as a programmatic dependent of the one before it (common.cuh)
        TNB_CUDA(launch_pdl(k_sd_count, dim3(nb2), dim3(kScanThreads), 0, s, sa));
        TNB_LAUNCH_CHECK();
        TNB_CUDA(launch_pdl(k_sd_write, dim3(nb2), dim3(kScanThreads), 0, s, sa));
        TNB_LAUNCH_CHECK();
        prof_begin(TNB_PROF_NEW_VERTICES, s);
        if (net->fixed_cfg) TNB_CUDA(launch_pdl(k_sd_new_vertices<CfgRef>, dim3(kSMs * 8), dim3(kThreads), row_tile, s, m, sa));
        else TNB_CUDA(launch_pdl(k_sd_new_vertices<CfgAny>, dim3(kSMs * 8), dim3(kThreads), row_tile, s, m, sa));
        TNB_LAUNCH_CHECK();
        prof_end(TNB_PROF_NEW_VERTICES, s, 0);
        TNB_CUDA(launch_pdl(k_sd_cands, dim3(kSMs * 8), dim3(256), 0, s, sa, pm));
        TNB_LAUNCH_CHECK();
        if ((rc = cells_build(1, sa.cand, sa.cnt + C_CAND, cand_ub, sa.sig[c->vcur], (int2 *)sa.head, sa.cslot, sa.next, sa.cnt + C_RECS, sa.dim, s))) return rc;
        prof_begin(TNB_PROF_PAIRS, s);
        TNB_CUDA(launch_pdl(k_pair_count_seg, dim3(kSMs * 8), dim3(256), 0, s, (const int *)sa.cand, sa.cnt, (const uint64_t *)sa.sig[c->vcur], (const int2 *)sa.head,
                            (const tnb_bucket_rec *)sa.next, sa.dim, (uint64_t)0, sa.pcount, sa.pcache, sa.remap, (const int *)(sa.cnt + C_IDX)));
        TNB_LAUNCH_CHECK();
        prof_end(TNB_PROF_PAIRS, s, 0);
        if ((rc = compact(cand_ub, ArrayCount{sa.pcount}, OffsetEmit{sa.poff}, sa.block_sums, sa.cnt + C_PAIRS, s, sa.cnt + C_CAND))) return rc;
        prof_begin(TNB_PROF_PAIRS, s);
        TNB_CUDA(launch_pdl(k_sd_pair_copy, dim3(kSMs * 16), dim3(256), 0, s, sa));
        TNB_LAUNCH_CHECK();
        static const int long_ctas = std::getenv("TNB_PAIR_LONG_CTAS") ? std::atoi(std::getenv("TNB_PAIR_LONG_CTAS")) : 2;  // CTAs per SM (A/B)
        TNB_CUDA(launch_pdl(k_sd_pair_long, dim3(kSMs * long_ctas), dim3(kSortWarps * 32), 0, s, sa));
        TNB_LAUNCH_CHECK();
        prof_end(TNB_PROF_PAIRS, s, 0);
        if ((rc = cells_clear(1, sa.cnt + C_CAND, cand_ub, (int2 *)sa.head, sa.cslot, s))) return rc;
        TNB_CUDA(launch_pdl(k_sd_keep_count, dim3(kScanMaxBlocks), dim3(kScanThreads), 0, s, sa));
        TNB_LAUNCH_CHECK();
        TNB_CUDA(launch_pdl(k_sd_keep_write, dim3(kScanMaxBlocks), dim3(kScanThreads), 0, s, sa));
        TNB_LAUNCH_CHECK();
        TNB_CUDA(launch_pdl(k_sd_finish, dim3(1), dim3(32), 0, s, sa, list, 0, seq, (volatile int *)report_dev));
        TNB_LAUNCH_CHECK();
        return TNB_OK;
    };
    c->bytes_by_half = true;
    c->counts_stale = true;
    c->cross_stale = true;
    c->maybe_dead = true;
    int rc;
    // Reports: seq0 = the initial choice; sequence k (k = 0, 1, ...) ends with report seq0 + 1 + k, which holds the
    // sizes after it and the choice for sequence k + 1.  Sequence numbers are reserved up front.
    const int seq0 = seq_base + 1;
    seq_base += n_steps + 2;
    k_sd_finish<<<1, 32, 0, s>>>(sa, list, 1, seq0, report_dev);
    TNB_LAUNCH_CHECK();
    if ((rc = enqueue_sequence(seq0 + 1))) return rc;   // run-ahead: enqueued before its own choice is known
    int enq = 1;
    *consumed = n_steps;
    for (;;) {
        // the report that precedes the sequence enqueued last: is that sequence a real step, and of which list entry?
        if ((rc = wait_for(seq0 + enq - 1))) return rc;
        const int more = rep[1], pos = rep[4];
        const int64_t items = (int64_t)rep[2] + rep[3];   // sizes BEFORE that sequence
        if (rep[5]) return TNB_OK;           // sticky error: what is in flight does nothing; complex_sync_counts reports it
        if (!more) return TNB_OK;            // the sequence in flight finds no step: every plane left crosses nothing
        if (pos + 1 >= n_steps) return TNB_OK;   // it runs the last list entry: nothing can follow
        if (g_fused_steps && items <= g_fused_max_items) {
            // small enough for the persistent kernel by now: let the sequence in flight finish; its report says where
            // the list goes on
            if ((rc = wait_for(seq0 + enq))) return rc;
            if (!rep[5] && rep[1]) {
                *consumed = rep[4];
                c->V = rep[2];
                c->E = rep[3];
            }
            return TNB_OK;
        }
        if ((rc = enqueue_sequence(seq0 + 1 + enq))) return rc;
        ++enq;
    }
}

// Drop the rows of dead vertices (order preserving) and renumber the edges: what the reference
// does after every hyperplane (subpoly.py:268-277), done here once, when the complex is read.
int complex_compact(tnb_complex *c, cudaStream_t s)
{
    int rc;
    if ((rc = complex_sync_counts(c, s))) return rc;
    if (!c->maybe_dead) return TNB_OK;
    c->maybe_dead = false;
    if (c->V == 0) return TNB_OK;
    int *cnt = c->counters.p;
    const int o = c->vcur ^ 1, R = c->R;
    k_set_scratch_count<<<1, 1, 0, s>>>(cnt + C_KEPT, (int)c->V);  // slots before the compaction, for k_move_rows
    TNB_LAUNCH_CHECK();
    if ((rc = compact(c->V, FlagCount{c->calive()}, VertexMoveEmit{c->remap.p}, c->block_sums.p, cnt + C_V, s))) return rc;
    k_move_rows<<<grid_for(c->V * R, 256, kSMs * 8), 256, 0, s>>>(cnt + C_KEPT, R, c->calive(), c->remap.p,
        VertexArrays{c->cvert(), c->cout_(), c->csig(), c->tag[c->vcur].p, c->vert[o].p, c->out[o].p, c->sig[o].p, c->tag[o].p});
    TNB_LAUNCH_CHECK();
    c->vcur = o;
    if (c->E > 0) {
        k_remap_edges_dev<<<grid_for(c->E, 256), 256, 0, s>>>(c->cedges(), cnt + C_E, c->remap.p);
        TNB_LAUNCH_CHECK();
    }
    TNB_CUDA(cudaMemsetAsync(c->calive(), 1, (size_t)c->V * sizeof(int), s));  // what is left is alive
    k_set_parity<<<1, 1, 0, s>>>(cnt, c->vcur, c->ecur, c->acur);
    TNB_LAUNCH_CHECK();
    c->counts_stale = true;
    return complex_sync_counts(c, s);
}

// ---- hypercube fallback (subpoly.py:51-52, :731-750) -----------------------------------------------
static void hypercube(float size, std::vector<float> &v, std::vector<int64_t> &e)
{
    const float c[2] = {-size, size};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) { v.push_back(c[i]); v.push_back(c[j]); v.push_back(c[k]); }
    for (int a = 0; a < 8; ++a)
        for (int b = a + 1; b < 8; ++b) {
            int diff = 0;
            for (int d = 0; d < 3; ++d) diff += (v[3 * a + d] * v[3 * b + d] < 0.0f) ? 1 : 0;
            if (diff == 1) { e.push_back(a); e.push_back(b); }
        }
}

__global__ void k_edges_from_i64(const int64_t *__restrict__ src, int64_t E, int2 *__restrict__ dst)
{
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x)
        dst[e] = make_int2((int)src[2 * e], (int)src[2 * e + 1]);
}
__global__ void k_edges_to_i64(const int2 *__restrict__ src, int64_t E, int64_t *__restrict__ dst)
{
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        dst[2 * e] = src[e].x;
        dst[2 * e + 1] = src[e].y;
    }
}

static int from_arrays_impl(const tnb_net *net, const float *d_vertices, int64_t V, const int64_t *d_edges, int64_t E,
                            tnb_complex **out, cudaStream_t s)
{
    tnb_complex *c = new tnb_complex();
    *out = c;
    int rc = complex_alloc(c, net, (size_t)(std::max<int64_t>(V, 4096) * capacity_factor()), (size_t)(std::max<int64_t>(E, 4096) * capacity_factor()));
    if (rc) return rc;
    if (V > 0) TNB_CUDA(cudaMemcpyAsync(c->cvert(), d_vertices, (size_t)V * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (E > 0) {
        k_edges_from_i64<<<grid_for(E, 256), 256, 0, s>>>(d_edges, E, c->cedges());
        TNB_LAUNCH_CHECK();
    }
    c->V = V;
    c->E = E;
    k_set_counts<<<1, 1, 0, s>>>(c->counters.p, (int)V, (int)E, c->vcur, c->ecur, c->acur);
    TNB_LAUNCH_CHECK();
    if ((rc = eval_vertices(net, c, 0, V, s))) return rc;
    return initial_cross_mask(c, s);
}

}  // namespace tnb

using namespace tnb;

extern "C" {

int tnb_set_capacity_factor(double f)
{
    if (!(f >= 1.0)) { set_error("capacity factor must be >= 1"); return TNB_ERR_INVALID; }
    g_capacity_factor = f;
    return TNB_OK;
}

int64_t tnb_set_fused_max_items(int64_t items)
{
    const int64_t before = g_fused_max_items;
    if (items >= 0) g_fused_max_items = items;
    return before;
}

int64_t tnb_set_cluster_max_items(int64_t items)
{
    const int64_t before = g_cluster_max_items;
    if (items >= 0) g_cluster_max_items = items;
    return before;
}

int tnb_skeleton(const tnb_net *net, int32_t unit, float size, tnb_complex **out, void *stream)
{
    if (!net || !out) { set_error("tnb_skeleton: null argument"); return TNB_ERR_INVALID; }
    *out = nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    current_stream() = s;
    tnb_complex *c = nullptr;
    int rc = skeleton_impl(net, unit, &c, s);
    if (rc == TNB_OK && c->E == 0 && size > 0.0f) {  // we start with a hypercube (subpoly.py:51-52)
        delete c;
        c = nullptr;
        std::vector<float> v;
        std::vector<int64_t> e;
        hypercube(size, v, e);
        DevBuf<float> dv;
        DevBuf<int64_t> de;
        TNB_CUDA(dv.reserve(v.size()));
        TNB_CUDA(de.reserve(e.size()));
        TNB_CUDA(cudaMemcpyAsync(dv.p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice, s));
        TNB_CUDA(cudaMemcpyAsync(de.p, e.data(), e.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        rc = from_arrays_impl(net, dv.p, 8, de.p, (int64_t)e.size() / 2, &c, s);
        cudaStreamSynchronize(s);  // v, e (host vectors) are read by the async copies
    }
    if (rc != TNB_OK) { delete c; return rc; }
    c->stream = s;
    *out = c;
    return TNB_OK;
}

// ---- slab-sharded skeleton: sweep | (max_grad reduced over the ranks by the caller) | edges ----------
int tnb_skeleton_sweep(const tnb_net *net, int32_t unit, int32_t x_lo, int32_t x_hi, int32_t shared_lower,
                       int32_t shared_upper, tnb_sweep **out, void *stream)
{
    if (!net || !out) { set_error("tnb_skeleton_sweep: null argument"); return TNB_ERR_INVALID; }
    *out = nullptr;
    if ((shared_lower || shared_upper) && x_hi <= x_lo) { set_error("tnb_skeleton_sweep: a shared slab needs at least one cell"); return TNB_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    current_stream() = s;
    tnb_sweep *sw = new tnb_sweep();
    int rc = sweep_impl(net, unit, x_lo, x_hi, shared_lower != 0, shared_upper != 0, sw, s);
    if (rc != TNB_OK) { delete sw; return rc; }
    *out = sw;
    return TNB_OK;
}
int tnb_skeleton_sweep_alloc(const tnb_net *net, int32_t unit, tnb_sweep **out, void *stream)
{
    if (!net || !out) { set_error("tnb_skeleton_sweep_alloc: null argument"); return TNB_ERR_INVALID; }
    *out = nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    current_stream() = s;
    tnb_sweep *sw = new tnb_sweep();
    int rc = sweep_impl(net, unit, 0, net->meta.n_marks - 1, false, false, sw, s, false);
    if (rc != TNB_OK) { delete sw; return rc; }
    *out = sw;
    return TNB_OK;
}
int64_t tnb_sweep_num_planes(const tnb_sweep *sw) { return sw ? sw->x_hi - sw->x_lo + 1 : 0; }
int tnb_sweep_read_dist(const tnb_sweep *sw, float *d_out, void *stream)
{
    if (!sw || !d_out) { set_error("tnb_sweep_read_dist: null argument"); return TNB_ERR_INVALID; }
    const size_t n = (size_t)(sw->x_hi - sw->x_lo + 1) * sw->M * sw->M;
    TNB_CUDA(cudaMemcpyAsync(d_out, sw->dist.p, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TNB_OK;
}
int tnb_sweep_write_dist(tnb_sweep *sw, const float *d_in, int32_t x_lo, int32_t x_hi, void *stream)
{
    if (!sw || !d_in || x_lo < sw->x_lo || x_hi > sw->x_hi || x_lo > x_hi) { set_error("tnb_sweep_write_dist: bad argument"); return TNB_ERR_INVALID; }
    const size_t plane = (size_t)sw->M * sw->M;
    TNB_CUDA(cudaMemcpyAsync(sw->dist.p + (size_t)(x_lo - sw->x_lo) * plane, d_in, (size_t)(x_hi - x_lo + 1) * plane * sizeof(float),
                             cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TNB_OK;
}
void tnb_sweep_destroy(tnb_sweep *sw) { delete sw; }
int32_t tnb_sweep_num_chunks(const tnb_sweep *sw) { return sw ? sw->n_chunks : 0; }
int tnb_sweep_read_max_grad(const tnb_sweep *sw, float *d_out, void *stream)
{
    if (!sw || !d_out) { set_error("tnb_sweep_read_max_grad: null argument"); return TNB_ERR_INVALID; }
    TNB_CUDA(cudaMemcpyAsync(d_out, sw->max_grad.p, (size_t)sw->n_chunks * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TNB_OK;
}
int tnb_sweep_write_max_grad(tnb_sweep *sw, const float *d_in, void *stream)
{
    if (!sw || !d_in) { set_error("tnb_sweep_write_max_grad: null argument"); return TNB_ERR_INVALID; }
    TNB_CUDA(cudaMemcpyAsync(sw->max_grad.p, d_in, (size_t)sw->n_chunks * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TNB_OK;
}
int tnb_skeleton_finish(const tnb_net *net, tnb_sweep *sw, tnb_complex **out, void *stream)
{
    if (!net || !sw || !out) { set_error("tnb_skeleton_finish: null argument"); return TNB_ERR_INVALID; }
    *out = nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    current_stream() = s;
    tnb_complex *c = nullptr;
    int rc = skeleton_finish_impl(net, sw, &c, s);
    if (rc != TNB_OK) { delete c; return rc; }
    c->stream = s;
    *out = c;
    return TNB_OK;
}

// ---- mailboxes of the slab exchange (plain cudaMalloc: exportable over CUDA IPC) ------------------------
int64_t tnb_mailbox_bytes(int64_t payload) { return (int64_t)halo_box_bytes((size_t)payload); }
int tnb_mailbox_create(int64_t payload, void **out)
{
    if (!out || payload < 16) { set_error("tnb_mailbox_create: bad argument"); return TNB_ERR_INVALID; }
    void *p = nullptr;
    TNB_CUDA(cudaMalloc(&p, halo_box_bytes((size_t)payload)));
    TNB_CUDA(cudaMemset(p, 0, halo_box_bytes((size_t)payload)));
    *out = p;
    return TNB_OK;
}
int tnb_mailbox_destroy(void *box) { TNB_CUDA(cudaFree(box)); return TNB_OK; }
int tnb_mailbox_export(void *box, void *handle64)
{
    if (!box || !handle64) { set_error("tnb_mailbox_export: null argument"); return TNB_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    TNB_CUDA(cudaIpcGetMemHandle(&h, box));
    memcpy(handle64, &h, 64);
    return TNB_OK;
}
int tnb_mailbox_import(const void *handle64, void **out)
{
    if (!handle64 || !out) { set_error("tnb_mailbox_import: null argument"); return TNB_ERR_INVALID; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    TNB_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return TNB_OK;
}
int tnb_mailbox_release(void *imported) { TNB_CUDA(cudaIpcCloseMemHandle(imported)); return TNB_OK; }

int tnb_complex_set_halo(tnb_complex *c, int32_t rank, int32_t world, void *const *boxes, int64_t payload, int32_t timeout_ms,
                         uint32_t seq0)
{
    if (!c || !boxes || world < 1 || world > kHaloMaxWorld || rank < 0 || rank >= world || payload < 16) {
        set_error("tnb_complex_set_halo: bad argument");
        return TNB_ERR_INVALID;
    }
    tnb_halo &h = c->halo;
    if (h.tag_lower && rank == 0) { set_error("tnb_complex_set_halo: rank 0 has no lower neighbour"); return TNB_ERR_INVALID; }
    if (h.tag_upper && rank == world - 1) { set_error("tnb_complex_set_halo: the last rank has no upper neighbour"); return TNB_ERR_INVALID; }
    current_stream() = c->stream;
    h.rank = rank; h.world = world; h.payload = (size_t)payload; h.seq = seq0;
    for (int r = 0; r < world; ++r) {
        if (!boxes[r]) { set_error("tnb_complex_set_halo: null mailbox"); return TNB_ERR_INVALID; }
        h.boxes[r] = (unsigned char *)boxes[r];
    }
    TNB_CUDA(h.slot.reserve(c->Vcap));
    TNB_CUDA(h.stage[0].reserve((size_t)payload));
    TNB_CUDA(h.stage[1].reserve((size_t)payload));
    TNB_CUDA(h.stage_count.reserve(4));
    TNB_CUDA(cudaMemsetAsync(h.stage_count.p, 0, 4 * sizeof(int), c->stream));
    if (timeout_ms > 0) g_halo_timeout_ns = (long long)timeout_ms * 1000000ll;
    h.enabled = true;
    return TNB_OK;
}

int tnb_subpoly_step_part(const tnb_net *net, tnb_complex *c, int32_t l, int32_t h, float eps, int32_t force, int32_t part,
                          void *stream)
{
    if (!net || !c || part < 0 || part > 2) { set_error("tnb_subpoly_step_part: bad argument"); return TNB_ERR_INVALID; }
    current_stream() = (cudaStream_t)stream;
    c->stream = (cudaStream_t)stream;
    return step_impl(net, c, l, h, eps, force != 0, (cudaStream_t)stream, part);
}

int tnb_complex_from_arrays(const tnb_net *net, const float *d_vertices, int64_t V, const int64_t *d_edges, int64_t E,
                            tnb_complex **out, void *stream)
{
    if (!net || !out || V < 0 || E < 0 || (V > 0 && !d_vertices) || (E > 0 && !d_edges)) {
        set_error("tnb_complex_from_arrays: bad argument");
        return TNB_ERR_INVALID;
    }
    *out = nullptr;
    current_stream() = (cudaStream_t)stream;
    tnb_complex *c = nullptr;
    int rc = from_arrays_impl(net, d_vertices, V, d_edges, E, &c, (cudaStream_t)stream);
    if (rc != TNB_OK) { delete c; return rc; }
    c->stream = (cudaStream_t)stream;
    *out = c;
    return TNB_OK;
}

// the caller's cached rows (subpoly.py:92-95: `outputs_` may carry the zeros the failover override wrote) replace the
// evaluated ones; the packed signs follow them
static __global__ void k_pack_rows(const float *__restrict__ out, int64_t V, int R, float eps, uint64_t *__restrict__ sig)
{
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        uint64_t pos, neg;
        pack_signs(out + v * R, R, eps, pos, neg);
        sig[3 * v] = pos;
        sig[3 * v + 1] = neg;
    }
}
int tnb_complex_write_outputs(const tnb_net *net, tnb_complex *c, const float *d_outputs, void *stream)
{
    if (!net || !c || !d_outputs) { set_error("tnb_complex_write_outputs: null argument"); return TNB_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    current_stream() = s;
    c->stream = s;
    int rc = complex_compact(c, s);   // the rows are the caller's: one per vertex of the complex as tnb_complex_read returns it
    if (rc) return rc;
    if (c->V == 0) return TNB_OK;
    TNB_CUDA(cudaMemcpyAsync(c->cout_(), d_outputs, (size_t)c->V * c->R * sizeof(float), cudaMemcpyDeviceToDevice, s));
    k_pack_rows<<<grid_for(c->V, 256), 256, 0, s>>>(c->cout_(), c->V, c->R, net->meta.eps, c->csig());
    TNB_LAUNCH_CHECK();
    TNB_CUDA(cudaMemsetAsync(c->counters.p + C_CROSS, 0, sizeof(uint64_t), s));
    return initial_cross_mask(c, s);
}

void tnb_complex_destroy(tnb_complex *c) { delete c; }
int64_t tnb_complex_num_vertices(const tnb_complex *c)
{
    if (!c) return 0;
    current_stream() = c->stream;
    if (complex_compact(const_cast<tnb_complex *>(c), c->stream) != TNB_OK) return -1;  // tnb_last_error() says why
    return c->V;
}
int64_t tnb_complex_num_edges(const tnb_complex *c)
{
    if (!c) return 0;
    if (complex_sync_counts(const_cast<tnb_complex *>(c), c->stream) != TNB_OK) return -1;
    return c->E;
}

int tnb_complex_read(const tnb_complex *c, float *d_vertices, int64_t *d_edges, float *d_outputs, void *stream)
{
    if (!c) { set_error("tnb_complex_read: null complex"); return TNB_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    current_stream() = s;
    int rcs = complex_compact(const_cast<tnb_complex *>(c), s);
    if (rcs) return rcs;
    if (d_vertices && c->V) TNB_CUDA(cudaMemcpyAsync(d_vertices, c->cvert(), (size_t)c->V * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (d_outputs && c->V) TNB_CUDA(cudaMemcpyAsync(d_outputs, c->cout_(), (size_t)c->V * c->R * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (d_edges && c->E) {
        k_edges_to_i64<<<grid_for(c->E, 256), 256, 0, s>>>(c->cedges(), c->E, d_edges);
        TNB_LAUNCH_CHECK();
    }
    return TNB_OK;
}

int tnb_subpoly_steps(const tnb_net *net, tnb_complex *c, const int32_t *lh, int32_t n_steps, float eps, int32_t force,
                      void *stream)
{
    if (!net || !c || (n_steps > 0 && !lh) || n_steps < 0) { set_error("tnb_subpoly_steps: bad argument"); return TNB_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    current_stream() = s;
    c->stream = s;
    for (int i = 0; i < n_steps;) {
        // as soon as the complex is small enough, everything that is left runs as ONE persistent launch
        const int mode = steps_mode(net, c, force != 0);
        if (mode) {
            const int rc = steps_persistent_impl(net, c, lh + 2 * i, n_steps - i, eps, mode, force != 0, s);
            if (rc != TNB_ERR_UNSUPPORTED) return rc;
        }
        if (stream_ok(net, c, eps, force != 0)) {   // a large complex: device-driven stream of launches, no host syncs
            int consumed = 0;
            const int rc = steps_stream_impl(net, c, lh + 2 * i, n_steps - i, eps, s, &consumed);
            if (rc) return rc;
            if (consumed > 0) { i += consumed; continue; }
        }
        const int rc = step_impl(net, c, lh[2 * i], lh[2 * i + 1], eps, force != 0, s);
        if (rc) return rc;
        ++i;
    }
    return TNB_OK;
}

int tnb_subpoly_step(const tnb_net *net, tnb_complex *c, int32_t l, int32_t h, float eps, int32_t force, void *stream)
{
    if (!net || !c) { set_error("tnb_subpoly_step: null argument"); return TNB_ERR_INVALID; }
    current_stream() = (cudaStream_t)stream;
    c->stream = (cudaStream_t)stream;
    return step_impl(net, c, l, h, eps, force != 0, (cudaStream_t)stream);
}

}  // extern "C"

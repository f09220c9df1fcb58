// cells.cuh -- contiguous per-cell segments of vertex records (count -> allocate -> fill).
//
// The connecting-edge search (subpoly.py:484-535) and the face rows (subpoly.py:281-370) both ask
// "which candidates lie in this marks-grid cell".  Linked lists per cell answer that with one
// DEPENDENT load per record: a cell that holds one of the reference's coincident-vertex clusters
// (chunk-overlap duplicates, up to 250 vertices) costs 250 L2 round trips per lane.  Here every
// touched cell owns a contiguous segment of the record array instead, so a lane (or a whole warp)
// STREAMS a cell: independent, coalesced 32-byte loads.
//
//   cells[cell] = {count, base}   int2 over the dense (M+2)^3 cell grid, all-zero between uses
//   slots[K*i+s] = {cell, li}     K = 8: item i is filed in every cell it lies in (up to 8: the face rows ask "who
//                                 lies in THIS cell"); K = 1: only in its lowest cell (the connecting-edge search
//                                 looks at the up-to-27 cells around a candidate instead, and a cluster of
//                                 coincident vertices is streamed once, not eight times).  li = index inside the segment
//   recs[base + li]               the item's record (vertex number + packed sign vector)
//
// Three small kernels: k_cell_count (one atomicAdd per (item, cell) returns li), k_cell_alloc (the
// li == 0 item of every touched cell carves count records out of the array: one atomicAdd per CTA),
// k_cell_fill; k_cell_clear zeroes the touched cells again (the grid is 8 M cells for the large
// model: never memset per step).  The order inside a segment is whatever the atomics gave; every
// consumer sorts what it collects, so results do not depend on it.
#pragma once
#include "complex.cuh"
#include "scan.cuh"

namespace tnb {

// A vertex lies in the cells [lo_d, hi_d] per axis: hi = offset, lo = offset - 1 when it sits on the
// grid plane (mask 0) -- the (m-1)//2 + offset expansion of subpoly.py:332.
struct CellBox {
    int lo[3], hi[3];
};
__device__ __forceinline__ CellBox cell_box(uint64_t g)
{
    CellBox b;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        b.hi[d] = grid_off(g, d);
        b.lo[d] = b.hi[d] - (grid_mask(g, d) ? 0 : 1);
    }
    return b;
}
__device__ __forceinline__ int64_t cell_id(int cx, int cy, int cz, int dim)
{
    return ((int64_t)(cx + 2) * dim + (cy + 2)) * dim + (cz + 2);
}

// items: n (or *n_dev when given) entries; item i is vertex cand[i] (cand == nullptr: vertex i)
template <int K>
static __global__ void __launch_bounds__(256) k_cell_count(const int *__restrict__ cand, const int *__restrict__ n_dev, int64_t n,
                                                    const uint64_t *__restrict__ sig, int2 *__restrict__ cells,
                                                    int2 *__restrict__ slots, int dim)
{
    pdl_wait();
    if (n_dev) n = *n_dev;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = cand ? cand[i] : i;
        const CellBox b = cell_box(sig[3 * v + 2]);
        int cell[K], li[K];
        // all atomics are issued before the first result is used: their round trips overlap
#pragma unroll
        for (int s = 0; s < K; ++s) {
            const int cx = b.lo[0] + (s >> 2), cy = b.lo[1] + ((s >> 1) & 1), cz = b.lo[2] + (s & 1);
            const bool in = cx <= b.hi[0] && cy <= b.hi[1] && cz <= b.hi[2];
            cell[s] = in ? (int)cell_id(cx, cy, cz, dim) : -1;
            li[s] = in ? atomicAdd(&cells[cell[s]].x, 1) : 0;
        }
#pragma unroll
        for (int s = 0; s < K; ++s) slots[K * i + s] = make_int2(cell[s], li[s]);
    }
}

// first record of every touched cell: the item that got li == 0 allocates for its cell
static __global__ void __launch_bounds__(256) k_cell_alloc(const int *__restrict__ n_dev, int64_t n, int K, int2 *__restrict__ cells,
                                                    const int2 *__restrict__ slots, int *__restrict__ total)
{
    pdl_wait();
    if (n_dev) n = *n_dev;
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n8 = n * K, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t0 = blockIdx.x * (int64_t)blockDim.x; t0 < n8; t0 += stride) {  // CTA-uniform trip count
        const int64_t t = t0 + threadIdx.x;
        int want = 0, cell = -1;
        if (t < n8) {
            const int2 s = slots[t];
            if (s.x >= 0 && s.y == 0) { cell = s.x; want = cells[cell].x; }
        }
        const int incl = warp_inclusive_scan(want);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int w = 0; w < 8; ++w) { const int x = s_warp[w]; s_warp[w] = run; run += x; }
            s_base = run ? atomicAdd(total, run) : 0;
        }
        __syncthreads();
        if (cell >= 0) cells[cell].y = s_base + s_warp[warp] + incl - want;
        __syncthreads();
    }
}

static __global__ void __launch_bounds__(256) k_cell_fill(const int *__restrict__ cand, const int *__restrict__ n_dev, int64_t n, int K,
                                                   const uint64_t *__restrict__ sig, const int2 *__restrict__ cells,
                                                   const int2 *__restrict__ slots, tnb_bucket_rec *__restrict__ recs)
{
    pdl_wait();
    if (n_dev) n = *n_dev;
    const int64_t n8 = n * K;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n8; t += (int64_t)gridDim.x * blockDim.x) {
        const int2 s = slots[t];
        if (s.x < 0) continue;
        const int64_t i = K == 8 ? t >> 3 : t / K, v = cand ? cand[i] : i;
        tnb_bucket_rec r;
        r.next = (int)i;  // the item number (the linked-list form keeps its link here)
        r.v = (int)v;
        r.pos = sig[3 * v];
        r.neg = sig[3 * v + 1];
        r.grd = sig[3 * v + 2];
        recs[cells[s.x].y + s.y] = r;
    }
}

// back to all-zero: only the touched cells are visited
static __global__ void __launch_bounds__(256) k_cell_clear(const int *__restrict__ n_dev, int64_t n, int K, int2 *__restrict__ cells,
                                                    const int2 *__restrict__ slots)
{
    pdl_wait();
    if (n_dev) n = *n_dev;
    const int64_t n8 = n * K;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n8; t += (int64_t)gridDim.x * blockDim.x) {
        const int2 s = slots[t];
        if (s.x >= 0 && s.y == 0) cells[s.x] = make_int2(0, 0);
    }
}

// count -> allocate -> fill for the items [0, n) (n_dev: the count lives on the device, n is the host's
// upper bound).  `total` (device int) must be zero on entry and holds the number of records afterwards.
// K = slots per item (8: every cell of the item's box, 1: its lowest cell only).
inline int cells_build(int K, const int *cand, const int *n_dev, int64_t n, const uint64_t *sig, int2 *cells, int2 *slots,
                       tnb_bucket_rec *recs, int *total, int dim, cudaStream_t s)
{
    if (n <= 0) return TNB_OK;
    const unsigned g1 = grid_for(n, 256), gk = grid_for(n * K, 256);
    if (K == 8) TNB_CUDA(launch_pdl(k_cell_count<8>, dim3(g1), dim3(256), 0, s, cand, n_dev, n, sig, cells, slots, dim));
    else TNB_CUDA(launch_pdl(k_cell_count<1>, dim3(g1), dim3(256), 0, s, cand, n_dev, n, sig, cells, slots, dim));
    TNB_LAUNCH_CHECK();
    TNB_CUDA(launch_pdl(k_cell_alloc, dim3(gk), dim3(256), 0, s, n_dev, n, K, cells, slots, total));
    TNB_LAUNCH_CHECK();
    TNB_CUDA(launch_pdl(k_cell_fill, dim3(gk), dim3(256), 0, s, cand, n_dev, n, K, sig, cells, slots, recs));
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}
inline int cells_clear(int K, const int *n_dev, int64_t n, int2 *cells, const int2 *slots, cudaStream_t s)
{
    if (n <= 0) return TNB_OK;
    TNB_CUDA(launch_pdl(k_cell_clear, dim3(grid_for(n * K, 256)), dim3(256), 0, s, n_dev, n, K, cells, slots));
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}

// ascending sort of one key per lane (bitonic network over the lanes, register shuffles).  Only the first `count`
// lanes hold keys (the others pass INT_MAX): the network stops at the first power of two >= count -- most
// partner lists have one to four entries, for which the full 15-stage network was most of the kernel's
// shuffle traffic.
__device__ __forceinline__ int warp_sort_asc(int key, int count = 32)
{
    const int lane = threadIdx.x & 31;
    int n = 2;
    while (n < count) n <<= 1;  // warp uniform
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int o = __shfl_xor_sync(0xffffffffu, key, j);
            const bool up = (lane & k) == 0 || k == n, lower = (lane & j) == 0;
            key = (lower == up) ? min(key, o) : max(key, o);
        }
    return key;
}

}  // namespace tnb

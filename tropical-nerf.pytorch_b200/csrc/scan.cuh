// scan.cuh -- order-preserving stream compaction / exclusive scan built on warp-level
// prefix scans.  This replaces the torch boolean indexing (`x[mask]`), `nonzero`,
// `masked_scatter_` and `unique(return_inverse=True)` calls the reference uses to
// compact vertices and edges (e.g. subpoly.py:113-114, :210-218, :265-272;
// tropical.py:93-102, :211-220): the output order is exactly the input order, which is
// what keeps vertex and edge numbering identical to the reference.
//
// Two launches per compaction over a fixed grid:
//   count:  every block owns one contiguous slice of the index space and adds up
//           count(i) of its slice (warp shuffle reduction) -> block_sums[b]
//   write:  every block sums block_sums[0..b) for its base offset, then walks its
//           slice tile by tile; inside a tile an inclusive warp scan (shfl_up) plus a
//           scan over the warp totals gives each item its exclusive position, and
//           emit(i, position, count) places it.
#pragma once
#include "common.cuh"

namespace tnb {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanMaxBlocks = kSMs * 4;

__device__ __forceinline__ int warp_inclusive_scan(int v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// slice of block b: [begin, end)
__device__ __forceinline__ void scan_slice(int64_t n, int64_t &begin, int64_t &end)
{
    int64_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + 31) / 32 * 32;  // warp-aligned slices: a short list still spreads over all CTAs
    begin = per * blockIdx.x;
    end = begin + per;
    if (begin > n) begin = n;
    if (end > n) end = n;
}

// the same for block `bid` of `nb` (kernels that run two compactions side by side in one grid)
__device__ __forceinline__ void scan_slice_part(int64_t n, int bid, int nb, int64_t &begin, int64_t &end)
{
    int64_t per = (n + nb - 1) / nb;
    per = (per + 31) / 32 * 32;
    begin = per * bid;
    end = begin + per;
    if (begin > n) begin = n;
    if (end > n) end = n;
}

// The two phases as device functions, so that fused persistent kernels can run them between
// grid / cluster syncs; NT = threads per CTA (blockDim.x).  The __global__ wrappers follow.
template <int NT, class Count>
__device__ __forceinline__ void scan_count_body_t(int64_t n, Count count, int *block_sums)
{
    constexpr int NW = NT / 32;
    int64_t begin, end;
    scan_slice(n, begin, end);
    int acc = 0;
    for (int64_t i = begin + threadIdx.x; i < end; i += NT) acc += count(i);
    acc = warp_sum(acc);
    __shared__ int s[NW];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = (int)threadIdx.x < NW ? s[threadIdx.x] : 0;
        t = warp_sum(t);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
    }
    __syncthreads();  // s[] may be reused by the caller's next phase
}

template <int NT, class Count, class Emit>
__device__ __forceinline__ void scan_write_body_t(int64_t n, Count count, Emit emit, const int *block_sums, int *total)
{
    constexpr int NW = NT / 32;
    __shared__ int s_warp[NW];
    __shared__ int s_excl[32];
    __shared__ int s_base, s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // base offset of this block = sum of the sums before it; block 0 also publishes the total
    {
        int acc = 0;
        const int upto = (blockIdx.x == 0) ? (int)gridDim.x : (int)blockIdx.x;
        for (int b = threadIdx.x; b < upto; b += NT) acc += block_sums[b];
        acc = warp_sum(acc);
        if (lane == 0) s_warp[warp] = acc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int t = lane < NW ? s_warp[lane] : 0;
            t = warp_sum(t);
            if (threadIdx.x == 0) {
                if (blockIdx.x == 0) { if (total) *total = t; s_base = 0; }
                else s_base = t;
            }
        }
        __syncthreads();
    }
    int64_t begin, end;
    scan_slice(n, begin, end);
    int running = s_base;
    for (int64_t tile = begin; tile < end; tile += NT) {
        const int64_t i = tile + threadIdx.x;
        const int c = (i < end) ? count(i) : 0;
        const int incl = warp_inclusive_scan(c);
        __syncthreads();  // s_warp / s_excl reuse
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {  // scan of the warp totals
            const int t = lane < NW ? s_warp[lane] : 0;
            const int ti = warp_inclusive_scan(t);
            s_excl[lane] = ti - t;
            if (lane == 31) s_tile = ti;
        }
        __syncthreads();
        if (c) emit(i, running + s_excl[warp] + incl - c, c);
        running += s_tile;
    }
    __syncthreads();
}

// Pieces for kernels that keep going after a compaction: every CTA derives its own base offset
// AND the grand total from the block sums (no second pass over a published total), and the write
// pass returns where the CTA's items went, so that the CTA can process exactly what it emitted.
template <int NT>
__device__ __forceinline__ void block_sums_reduce(const int *block_sums, int &base, int &total)
{
    constexpr int NW = NT / 32;
    __shared__ int s_b[NW], s_t[NW];
    __shared__ int s_out[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int before = 0, all = 0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += NT) {
        const int v = block_sums[b];
        all += v;
        if (b < (int)blockIdx.x) before += v;
    }
    before = warp_sum(before);
    all = warp_sum(all);
    if (lane == 0) { s_b[warp] = before; s_t[warp] = all; }
    __syncthreads();
    if (threadIdx.x < 32) {
        int x = lane < NW ? s_b[lane] : 0, y = lane < NW ? s_t[lane] : 0;
        x = warp_sum(x);
        y = warp_sum(y);
        if (lane == 0) { s_out[0] = x; s_out[1] = y; }
    }
    __syncthreads();
    base = s_out[0];
    total = s_out[1];
    __syncthreads();  // s_out may be rewritten by the next call
}

// write pass of this CTA's slice starting at position `base`; returns the position after its last item
template <int NT, class Count, class Emit>
__device__ __forceinline__ int scan_write_from(int64_t n, Count count, Emit emit, int base)
{
    constexpr int NW = NT / 32;
    __shared__ int s_warp[NW];
    __shared__ int s_excl[32];
    __shared__ int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t begin, end;
    scan_slice(n, begin, end);
    int running = base;
    for (int64_t tile = begin; tile < end; tile += NT) {
        const int64_t i = tile + threadIdx.x;
        const int c = (i < end) ? count(i) : 0;
        const int incl = warp_inclusive_scan(c);
        __syncthreads();  // s_warp / s_excl reuse
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int t = lane < NW ? s_warp[lane] : 0;
            const int ti = warp_inclusive_scan(t);
            s_excl[lane] = ti - t;
            if (lane == 31) s_tile = ti;
        }
        __syncthreads();
        if (c) emit(i, running + s_excl[warp] + incl - c, c);
        running += s_tile;
    }
    __syncthreads();
    return running;
}

// The same two passes over a range [begin, end) the CTA owns for another reason than scan_slice
// (e.g. the items it emitted itself in an earlier phase: ranges of consecutive CTAs are consecutive).
template <int NT, class Count>
__device__ __forceinline__ void scan_count_range(int64_t begin, int64_t end, Count count, int *block_sums)
{
    constexpr int NW = NT / 32;
    int acc = 0;
    for (int64_t i = begin + threadIdx.x; i < end; i += NT) acc += count(i);
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
template <int NT, class Count, class Emit>
__device__ __forceinline__ int scan_write_range(int64_t begin, int64_t end, Count count, Emit emit, int base)
{
    constexpr int NW = NT / 32;
    __shared__ int s_warp[NW];
    __shared__ int s_excl[32];
    __shared__ int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int running = base;
    for (int64_t tile = begin; tile < end; tile += NT) {
        const int64_t i = tile + threadIdx.x;
        const int c = (i < end) ? count(i) : 0;
        const int incl = warp_inclusive_scan(c);
        __syncthreads();  // s_warp / s_excl reuse
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int t = lane < NW ? s_warp[lane] : 0;
            const int ti = warp_inclusive_scan(t);
            s_excl[lane] = ti - t;
            if (lane == 31) s_tile = ti;
        }
        __syncthreads();
        if (c) emit(i, running + s_excl[warp] + incl - c, c);
        running += s_tile;
    }
    __syncthreads();
    return running;
}

template <class Count>
__device__ __forceinline__ void scan_count_body(int64_t n, Count count, int *block_sums)
{
    scan_count_body_t<kScanThreads>(n, count, block_sums);
}
template <class Count, class Emit>
__device__ __forceinline__ void scan_write_body(int64_t n, Count count, Emit emit, const int *block_sums, int *total)
{
    scan_write_body_t<kScanThreads>(n, count, emit, block_sums, total);
}

// n_dev (optional): the item count lives in device memory (a previous kernel produced it);
// n is then only the host's upper bound used to size the grid.
template <class Count>
__global__ void __launch_bounds__(kScanThreads) k_scan_count(int64_t n, const int *__restrict__ n_dev, Count count,
                                                             int *__restrict__ block_sums)
{
    pdl_wait();
    if (n_dev) n = *n_dev;
    scan_count_body(n, count, block_sums);
}

// Count pass that also leaves one bit per item (bit i & 31 of mask[i >> 5]): when count(i) is
// expensive and few items pass, the write pass reads the bit (MaskCount) instead of evaluating
// count(i) again.  Slices are warp aligned, so a warp's ballot is exactly one mask word.
template <class Count>
__device__ __forceinline__ void scan_count_mask_part(int64_t n, Count count, int *__restrict__ block_sums, uint32_t *__restrict__ mask,
                                                     int bid, int nb)
{
    constexpr int NW = kScanThreads / 32;
    int64_t begin, end;
    scan_slice_part(n, bid, nb, begin, end);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int acc = 0;
    for (int64_t i0 = begin + warp * 32; i0 < end; i0 += kScanThreads) {
        const int64_t i = i0 + lane;
        const int c = (i < end && count(i)) ? 1 : 0;
        const uint32_t word = __ballot_sync(0xffffffffu, c);
        if (lane == 0) { mask[i0 >> 5] = word; acc += __popc(word); }
    }
    __shared__ int s[NW];
    if (lane == 0) s[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < NW; ++w) t += s[w];
        block_sums[bid] = t;
    }
}
template <class Count>
__global__ void __launch_bounds__(kScanThreads) k_scan_count_mask(int64_t n, Count count, int *__restrict__ block_sums,
                                                                  uint32_t *__restrict__ mask, const int *__restrict__ n_dev = nullptr)
{
    pdl_wait();
    if (n_dev) n = *n_dev;
    scan_count_mask_part(n, count, block_sums, mask, (int)blockIdx.x, (int)gridDim.x);
}
// Write pass over the bits k_scan_count_mask left: one 32-bit word (32 items) per thread and tile instead
// of one item, so a tile is 8 192 items (the generic pass spends its time in the three CTA barriers of
// each 256-item tile when nearly all items fail: 152 us for the 24 M skeleton slots of the large model).
// Same slices, same order: emit(i, position, 1) for every set bit, ascending i.
template <class Emit>
__device__ __forceinline__ void scan_write_mask_part(int64_t n, const uint32_t *__restrict__ mask, Emit emit,
                                                     const int *__restrict__ block_sums, int *__restrict__ total, int bid, int nb)
{
    constexpr int NT = kScanThreads, NW = NT / 32;
    __shared__ int s_warp[NW];
    __shared__ int s_excl[32];
    __shared__ int s_base, s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        int acc = 0;
        const int upto = (bid == 0) ? nb : bid;
        for (int b = threadIdx.x; b < upto; b += NT) acc += block_sums[b];
        acc = warp_sum(acc);
        if (lane == 0) s_warp[warp] = acc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int t = lane < NW ? s_warp[lane] : 0;
            t = warp_sum(t);
            if (threadIdx.x == 0) {
                if (bid == 0) { if (total) *total = t; s_base = 0; }
                else s_base = t;
            }
        }
        __syncthreads();
    }
    int64_t begin, end;
    scan_slice_part(n, bid, nb, begin, end);
    const int64_t wb = begin >> 5, we = begin < end ? (end + 31) >> 5 : wb;  // slices are warp aligned: a word belongs to one CTA
    int running = s_base;
    for (int64_t tile = wb; tile < we; tile += NT) {
        const int64_t w = tile + threadIdx.x;
        uint32_t word = w < we ? mask[w] : 0u;  // bits past the end of the slice were written as 0
        const int c = __popc(word);
        const int incl = warp_inclusive_scan(c);
        __syncthreads();
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int t = lane < NW ? s_warp[lane] : 0;
            const int ti = warp_inclusive_scan(t);
            s_excl[lane] = ti - t;
            if (lane == 31) s_tile = ti;
        }
        __syncthreads();
        int pos = running + s_excl[warp] + incl - c;
        while (word) {
            const int b = __ffs(word) - 1;
            word &= word - 1;
            emit(w * 32 + b, pos++, 1);
        }
        running += s_tile;
    }
}
template <class Emit>
__global__ void __launch_bounds__(kScanThreads) k_scan_write_mask(int64_t n, const uint32_t *__restrict__ mask, Emit emit,
                                                                  const int *__restrict__ block_sums, int *__restrict__ total,
                                                                  const int *__restrict__ n_dev = nullptr)
{
    pdl_wait();
    if (n_dev) n = *n_dev;
    scan_write_mask_part(n, mask, emit, block_sums, total, (int)blockIdx.x, (int)gridDim.x);
}
struct MaskCount {
    const uint32_t *mask;
    __device__ __forceinline__ int operator()(int64_t i) const { return (int)((mask[i >> 5] >> (i & 31)) & 1u); }
};

template <class Count, class Emit>
__global__ void __launch_bounds__(kScanThreads) k_scan_write(int64_t n, const int *__restrict__ n_dev, Count count,
                                                             Emit emit, const int *__restrict__ block_sums,
                                                             int *__restrict__ total)
{
    pdl_wait();
    if (n_dev) n = *n_dev;
    scan_write_body(n, count, emit, block_sums, total);
}

// Host helper: runs both phases.  `block_sums` must hold kScanMaxBlocks ints; the total
// lands in *d_total (device).
template <class Count, class Emit>
inline int compact(int64_t n, Count count, Emit emit, int *block_sums, int *d_total, cudaStream_t s,
                   const int *n_dev = nullptr)
{
    if (n <= 0) {
        TNB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int), s));
        return TNB_OK;
    }
    int64_t blocks = (n + kScanThreads - 1) / kScanThreads;
    if (blocks > kScanMaxBlocks) blocks = kScanMaxBlocks;
    TNB_CUDA(launch_pdl(k_scan_count<Count>, dim3((unsigned)blocks), dim3(kScanThreads), 0, s, n, n_dev, count, block_sums));
    TNB_LAUNCH_CHECK();
    TNB_CUDA(launch_pdl(k_scan_write<Count, Emit>, dim3((unsigned)blocks), dim3(kScanThreads), 0, s, n, n_dev, count, emit, (const int *)block_sums, d_total));
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}


// The same compaction through a bit mask: the count pass evaluates count(i) ONCE and leaves one bit per
// item, the write pass takes 32 items per thread from the bits.  For sparse selections whose test is a
// gather (edges crossed by a plane, vertices hit by it: a few percent pass) the second evaluation and
// the per-256-item barriers of the generic write pass were most of the cost.
// `mask` must hold (n + 31) / 32 + kScanMaxBlocks words.
template <class Count, class Emit>
inline int compact_masked(int64_t n, Count count, Emit emit, int *block_sums, uint32_t *mask, int *d_total, cudaStream_t s,
                          const int *n_dev = nullptr)
{
    if (n <= 0) {
        TNB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int), s));
        return TNB_OK;
    }
    int64_t blocks = (n + kScanThreads - 1) / kScanThreads;
    if (blocks > kScanMaxBlocks) blocks = kScanMaxBlocks;
    TNB_CUDA(launch_pdl(k_scan_count_mask<Count>, dim3((unsigned)blocks), dim3(kScanThreads), 0, s, n, count, block_sums, mask, n_dev));
    TNB_LAUNCH_CHECK();
    TNB_CUDA(launch_pdl(k_scan_write_mask<Emit>, dim3((unsigned)blocks), dim3(kScanThreads), 0, s, n, (const uint32_t *)mask, emit, (const int *)block_sums, d_total, n_dev));
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}

}  // namespace tnb

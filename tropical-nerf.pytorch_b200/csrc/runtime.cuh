// runtime.cuh -- host-side objects behind the C ABI handles.
#pragma once
#include <vector>

#include "common.cuh"

namespace tnb {

// Stream the library is currently working on (set at every C ABI entry).  Device arrays
// are carved from the CUDA stream-ordered memory pool of the device, whose release
// threshold is raised once so that repeated extractions never go back to cudaMalloc.
cudaStream_t &current_stream();
void init_pool_once();

// Released blocks are filed under (stream, size class) and handed out again to the next request of that class on
// that stream -- safe by stream order, and no CUDA call at all: a small extraction makes ~100 allocations, and with
// several extractions in flight on one GPU cudaMallocAsync / cudaFreeAsync were what their host threads queued for
// (profiles/r2_batch_phases.txt: no overlap at all before, 4x after).  Size classes are 1/8 octave apart
// (<= 12.5 % slack).  A block filed under a stream that has been destroyed is only ever handed back to the pool
// (tnb_release_cached_blocks).
// TNB_NO_BLOCK_CACHE=1 turns it off (every request goes to cudaMallocAsync / cudaFreeAsync).
void *block_acquire(size_t bytes, cudaStream_t s, size_t *got, cudaError_t *err);
void block_release(void *p, size_t bytes, cudaStream_t s);

// RAII device array (stream-ordered allocation)
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;    // elements asked for
    size_t bytes = 0;  // size of the block behind it
    cudaStream_t stream = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) block_release(p, bytes, stream);
        p = nullptr;
        cap = 0;
        bytes = 0;
    }
    // (re)allocate without preserving contents
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        release();
        init_pool_once();
        stream = current_stream();
        cudaError_t e = cudaSuccess;
        p = (T *)block_acquire((n ? n : 1) * sizeof(T), stream, &bytes, &e);
        if (e == cudaSuccess) cap = n; else { p = nullptr; bytes = 0; }
        return e;
    }
    void swap(DevBuf &o)
    {
        T *tp = p; p = o.p; o.p = tp;
        size_t tc = cap; cap = o.cap; o.cap = tc;
        size_t tb = bytes; bytes = o.bytes; o.bytes = tb;
        cudaStream_t ts = stream; stream = o.stream; o.stream = ts;
    }
};

inline unsigned grid_for(int64_t n, int threads, int max_blocks = kSMs * 16)
{
    int64_t b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (unsigned)b;
}

}  // namespace tnb

struct tnb_net {
    tnb::NetMeta meta;
    bool fixed_cfg = false;  // CfgRef applies
    tnb::DevBuf<float2> table;
    tnb::DevBuf<float> mlp;
    tnb::DevBuf<float> marks;
    std::vector<float> h_marks;
    std::vector<float> h_scale;
    std::vector<uint32_t> h_res, h_size, h_off;
    mutable int face_row_hint = 0;  // longest face row of the previous extraction (faces.cu)
};

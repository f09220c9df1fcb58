// common.cuh -- shared definitions of the B200 mesh-extraction kernels (sm_100a).
//
// Float contract: the whole library is compiled with -fmad=false, IEEE division and
// square root, no flush-to-zero.  Every fused multiply-add is written explicitly as
// __fmaf_rn; everything else is a single rounded operation in source order.  This is
// the operation order oracle/trinet_ref.c documents, so results are bit-identical to
// the CPU checker.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/tropical_b200.h"

namespace tnb {

constexpr int kMaxLevels = TNB_MAX_LEVELS;
constexpr int kMaxLinear = TNB_MAX_LINEAR;
constexpr int kMaxHidden = TNB_MAX_HIDDEN;
constexpr int kMlpParamMax = 1024;  // floats kept in kernel-parameter (constant bank) space
constexpr int kSMs = 148;           // B200

// mode: how the 8 corner indices of a cell are formed (all three give exactly
// tiny-cuda-nn's grid_index() result; the first two avoid its per-corner multiplies and
// the runtime modulo):
//   0  dense level   index = cx + cy*res + cz*res^2  (wraps only for out-of-grid points)
//   1  hashed level with a power-of-two table: (cx ^ cy*P1 ^ cz*P2) & (size-1)
//   2  anything else: the generic routine
enum { kLevelDense = 0, kLevelHashPow2 = 1, kLevelGeneric = 2 };
struct LevelMeta {
    float scale;
    uint32_t res, size, off;
    uint32_t mode, res2;
};

// Passed BY VALUE to kernels (__grid_constant__): lives in the constant bank, so the
// MLP weights of the reference-sized networks are FFMA constant operands.
struct NetMeta {
    int L, H, NLIN, R;
    float pre_scale, pre_2s, eps;
    float pre_inv;   // 1 / pre_2s
    int pre_pow2;    // pre_2s is a power of two: multiplying by pre_inv IS the IEEE division
    int n_marks;
    int mlp_in_param;  // 1 when mlp_c holds the weights
    LevelMeta lvl[kMaxLevels];
    const float2 *table;  // [entries] (F = 2)
    const float *mlp;     // global copy of the packed MLP
    const float *marks;   // global copy of the marks
    float mlp_c[kMlpParamMax];
};

// ---- error plumbing ---------------------------------------------------------------
void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
void count_launch(int n = 1);
// optional CUDA-event timers around the heavy kernels (tnb_profile_*)
void prof_begin(int cls, cudaStream_t s);
void prof_end(int cls, cudaStream_t s, int64_t units, int64_t bytes = 0);
void prof_add(int cls, int64_t units, int64_t bytes);  // sizes that are only known after a later sync

#define TNB_CUDA(expr)                                                         \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return tnb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define TNB_LAUNCH_CHECK()                                                     \
    do {                                                                       \
        tnb::count_launch();                                                   \
        cudaError_t _e = cudaGetLastError();                                   \
        if (_e != cudaSuccess) return tnb::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

// ---- device helpers -----------------------------------------------------------------
#ifdef __CUDACC__
// Programmatic dependent launch: a kernel launched with launch_pdl() may be scheduled while the kernel before it
// in the stream drains; it must call pdl_wait() before it reads anything that kernel wrote (griddepcontrol.wait
// returns at once when the launch carried no programmatic dependency).  The device-driven step stream is 16 short
// kernels per hyperplane: this hides their launch latency behind the predecessor's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
extern bool g_pdl;   // A/B switch (TNB_NO_PDL)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float det_expf(float y)
{
    float n = rintf(y * 1.44269504088896341f);
    float r = __fmaf_rn(n, -0.693359375f, y);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    float e = __fmaf_rn(p, r * r, r) + 1.0f;
    float s = __uint_as_float((uint32_t)((int)n + 127) << 23);
    return e * s;
}

// tanh as oracle/trinet_ref.c defines it (libm's and CUDA's differ in the last ulp)
__device__ __forceinline__ float det_tanhf(float x)
{
    float ax = fabsf(x);
    float t;
    if (ax < 0.25f) {
        float x2 = ax * ax;
        float p = 0.021869488536155203f;
        p = __fmaf_rn(p, x2, -0.053968253968253971f);
        p = __fmaf_rn(p, x2, 0.13333333333333333f);
        p = __fmaf_rn(p, x2, -0.33333333333333331f);
        t = __fmaf_rn(ax * x2, p, ax);
    } else if (ax > 9.0f) {
        t = 1.0f;
    } else {
        float e = det_expf(2.0f * ax);
        t = 1.0f - __fdiv_rn(2.0f, e + 1.0f);
    }
    return copysignf(t, x);
}

__device__ __forceinline__ uint32_t grid_index(uint32_t size, uint32_t res, uint32_t cx,
                                               uint32_t cy, uint32_t cz)
{
    // tiny-cuda-nn grid_index(): dense while the running stride fits the table, prime
    // XOR hash otherwise
    uint32_t stride = 1, index = 0;
    if (stride <= size) { index += cx * stride; stride *= res; }
    else return (cx ^ (cy * 2654435761u) ^ (cz * 805459861u)) % size;
    if (stride <= size) { index += cy * stride; stride *= res; }
    else return (cx ^ (cy * 2654435761u) ^ (cz * 805459861u)) % size;
    if (stride <= size) { index += cz * stride; stride *= res; }
    else return (cx ^ (cy * 2654435761u) ^ (cz * 805459861u)) % size;
    if (size < stride) index = cx ^ (cy * 2654435761u) ^ (cz * 805459861u);
    return index % size;
}

// x / pre_2s, bit-identical to the IEEE division (exact scaling when pre_2s is a power of two)
__device__ __forceinline__ float div_2s(const NetMeta &n, float v)
{
    return n.pre_pow2 ? v * n.pre_inv : __fdiv_rn(v, n.pre_2s);
}

__device__ __forceinline__ void preprocess(const NetMeta &n, const float x[3], float xp[3])
{
#pragma unroll
    for (int d = 0; d < 3; ++d) xp[d] = div_2s(n, x[d] + n.pre_scale);
}

// the 8 corner indices of cell (cx,cy,cz): corner bit d set = +1 along axis d
__device__ __forceinline__ void corner_indices(const LevelMeta &lv, uint32_t cx, uint32_t cy, uint32_t cz,
                                               uint32_t idx[8])
{
    if (lv.mode == kLevelDense) {
        const uint32_t base = cx + cy * lv.res + cz * lv.res2;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t i = base + (c & 1) + ((c >> 1) & 1) * lv.res + ((c >> 2) & 1) * lv.res2;
            if (i >= lv.size) i %= lv.size;  // only for points outside the grid
            idx[c] = i;
        }
    } else if (lv.mode == kLevelHashPow2) {
        const uint32_t hx[2] = {cx, cx + 1u};
        const uint32_t hy[2] = {cy * 2654435761u, (cy + 1u) * 2654435761u};
        const uint32_t hz[2] = {cz * 805459861u, (cz + 1u) * 805459861u};
        const uint32_t mask = lv.size - 1u;
#pragma unroll
        for (int c = 0; c < 8; ++c) idx[c] = (hx[c & 1] ^ hy[(c >> 1) & 1] ^ hz[(c >> 2) & 1]) & mask;
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c)
            idx[c] = grid_index(lv.size, lv.res, cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1));
    }
}

// The same 8 indices for the cells every in-grid point lies in, without modulo: a dense cell (corners
// below the table size, or wrapping around it once), or a hashed level with a power-of-two table.
// false = this cell needs the general routine (a point far outside the grid, or a table size that is
// not a power of two): the callers then take their out-of-line slow path.  The split exists
// for code size: the evaluation kernels are fully unrolled straight-line code of several thousand
// instructions whose issue rate is bounded by instruction fetch (ncu: stall no_instruction is the
// largest stall of k_sweep_chunk), so inlined copies of paths that in-grid points never take cost time.
__device__ __forceinline__ bool corner_indices_fast(const LevelMeta &lv, uint32_t cx, uint32_t cy, uint32_t cz,
                                                    uint32_t idx[8])
{
    if (lv.mode == kLevelDense) {
        const uint32_t base = cx + cy * lv.res + cz * lv.res2;
        if (!(base < lv.size)) return false;
        const uint32_t top = base + 1u + lv.res + lv.res2;  // the largest of the 8 (size <= 2^31: no overflow)
        if (top < lv.size) {
#pragma unroll
            for (int c = 0; c < 8; ++c) idx[c] = base + (c & 1) + ((c >> 1) & 1) * lv.res + ((c >> 2) & 1) * lv.res2;
        } else {
            // the last layer of cells of a level (up to half of the grid at the coarsest one): the upper
            // corners wrap around the table once (1 + res + res^2 <= res^3 <= size), so i % size = i - size
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t i = base + (c & 1) + ((c >> 1) & 1) * lv.res + ((c >> 2) & 1) * lv.res2;
                idx[c] = i >= lv.size ? i - lv.size : i;
            }
        }
        return true;
    }
    if (lv.mode == kLevelHashPow2) {
        const uint32_t hx[2] = {cx, cx + 1u};
        const uint32_t hy[2] = {cy * 2654435761u, (cy + 1u) * 2654435761u};
        const uint32_t hz[2] = {cz * 805459861u, (cz + 1u) * 805459861u};
        const uint32_t mask = lv.size - 1u;
#pragma unroll
        for (int c = 0; c < 8; ++c) idx[c] = (hx[c & 1] ^ hy[(c >> 1) & 1] ^ hz[(c >> 2) & 1]) & mask;
        return true;
    }
    return false;
}

// trilinear interpolation of the 8 corner rows (tiny-cuda-nn kernel_grid, linear interpolation).
// The two features of a row advance together in ONE packed FFMA2 (Blackwell fma.rn.f32x2: two
// independent IEEE binary32 FMAs per issue slot, the weight broadcast to both halves), so each half
// is exactly the __fmaf_rn of the scalar definition in oracle/trinet_ref.c.
__device__ __forceinline__ float2 interpolate8(const float2 v[8], const float frac[3])
{
    float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
        float w = 1.0f;
#pragma unroll
        for (int d = 0; d < 3; ++d) w = w * ((corner >> d) & 1 ? frac[d] : 1.0f - frac[d]);
        acc = __ffma2_rn(make_float2(w, w), v[corner], acc);
    }
    return acc;
}
// a level through the general index routine, out of line (same arithmetic as the fast path)
static __device__ __noinline__ float2 encode_level_general(const float2 *tab, uint32_t size, uint32_t res, uint32_t cx, uint32_t cy,
                                                           uint32_t cz, float f0, float f1, float f2)
{
    const float frac[3] = {f0, f1, f2};
    float2 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = __ldg(tab + grid_index(size, res, cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1)));
    return interpolate8(v, frac);
}

// Level l of the hash encoding at xp.  Returns the two features; optionally the cell
// and the fractional position for the backward pass.
__device__ __forceinline__ float2 encode_level(const NetMeta &n, int l, const float xp[3],
                                               uint32_t cell[3], float frac[3])
{
    const LevelMeta lv = n.lvl[l];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float pos = __fmaf_rn(lv.scale, xp[d], 0.5f);
        float fl = floorf(pos);
        cell[d] = (uint32_t)(int)fl;
        frac[d] = pos - fl;
    }
    const float2 *tab = n.table + lv.off;
    uint32_t idx[8];
    if (!corner_indices_fast(lv, cell[0], cell[1], cell[2], idx))
        return encode_level_general(tab, lv.size, lv.res, cell[0], cell[1], cell[2], frac[0], frac[1], frac[2]);
    float2 v[8];
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) v[corner] = __ldg(tab + idx[corner]);
    return interpolate8(v, frac);
}

// (ix, iy, iz) of the flat lattice index i = ix + nx * (iy + ny * iz) along a grid-stride loop: split
// once per thread, then advanced by carries.  (Three 64-bit divisions per point were ~15 % of the
// instructions of the sweep kernels.)
struct LatticeStride {  // split of the grid stride, made on the host
    int sx, sy, sz;
};
inline LatticeStride lattice_stride(int64_t stride, int nx, int ny)
{
    const int64_t plane = (int64_t)nx * ny;
    LatticeStride s;
    s.sz = (int)(stride / plane);
    const int64_t r = stride - (int64_t)s.sz * plane;
    s.sy = (int)(r / nx);
    s.sx = (int)(r - (int64_t)s.sy * nx);
    return s;
}
#ifdef __CUDACC__
struct Lattice3 {
    int ix, iy, iz;      // current point
    int sx, sy, sz;      // the stride, split the same way
    int nx, ny;
    __device__ __forceinline__ Lattice3(int64_t first, LatticeStride st, int nx_, int ny_) : sx(st.sx), sy(st.sy), sz(st.sz), nx(nx_), ny(ny_)
    {
        if (first < 0x7fffffff) {  // every launch in practice: 32-bit divisions
            const unsigned f = (unsigned)first, plane = (unsigned)nx_ * (unsigned)ny_;
            iz = (int)(f / plane);
            const unsigned r = f - (unsigned)iz * plane;
            iy = (int)(r / (unsigned)nx_);
            ix = (int)(r - (unsigned)iy * (unsigned)nx_);
        } else {
            const int64_t plane = (int64_t)nx_ * ny_;
            iz = (int)(first / plane);
            const int r = (int)(first - (int64_t)iz * plane);
            iy = r / nx_;
            ix = r - iy * nx_;
        }
    }
    __device__ __forceinline__ void advance()
    {
        ix += sx;
        const int cx = ix >= nx ? 1 : 0;
        ix -= cx ? nx : 0;
        iy += sy + cx;
        const int cy = iy >= ny ? 1 : 0;
        iy -= cy ? ny : 0;
        iz += sz + cy;
    }
};
#endif

// torch.searchsorted(marks, v, right=False)
__device__ __forceinline__ int lower_bound(const float *__restrict__ marks, int m, float v)
{
    int lo = 0, hi = m;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (marks[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- packed sign vectors -------------------------------------------------------------
// Per vertex three 64-bit words:
//   pos : bit c set  <=>  outputs[c] >  eps            (sign +1)
//   neg : bit c set  <=>  outputs[c] < -eps            (sign -1)   (neither = sign 0)
//   grd : bits [0,20) [20,40) [40,60) = offset+1 per axis (tropical.py:230-231),
//         bits 60,61,62 = grid mask per axis (1 = strictly inside a cell, tropical.py:234)
__device__ __forceinline__ uint64_t pack_grid(const NetMeta &n, const float *__restrict__ marks,
                                              const float xp[3], float eps)
{
    uint64_t g = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int off = lower_bound(marks, n.n_marks, xp[d] + eps) - 1;
        float mk = marks[off < 0 ? off + n.n_marks : off];
        uint64_t inside = fabsf(mk - xp[d]) > eps ? 1ull : 0ull;
        g |= (uint64_t)(uint32_t)(off + 1) << (20 * d);
        g |= inside << (60 + d);
    }
    return g;
}
__device__ __forceinline__ int grid_off(uint64_t g, int d) { return (int)((g >> (20 * d)) & 0xFFFFF) - 1; }
__device__ __forceinline__ int grid_mask(uint64_t g, int d) { return (int)((g >> (60 + d)) & 1); }

#endif  // __CUDACC__

}  // namespace tnb

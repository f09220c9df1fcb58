// repair.cuh -- the gradient-descent repair of the curve path (subpoly_debug.deal_with_gradient_descent,
// subpoly_debug.py:121-165): intersections of two curved planes that the closed form left farther than eps from
// one of the planes walk down the gradient of d0^2 + d1^2 in steps of 1e-2 (in edge parameters, clamped to
// [0,1]) -- ALL of them another step while ANY of them is still off, at most 500 steps.  The reference then
// ends the extraction if one is still off (subpoly.py:172-174).
//
// The walks are independent of one another except for the step count they share, so the device does it in two
// passes: (A) every walk runs all 500 steps and notes in a 500-bit word at which steps it was within eps; the
// AND of those words over the walks names the first step T at which all of them were; (B) every walk runs
// again, T + 1 steps (500 if there is no such T).  Same operations in the same order as
// oracle/trinet_ref.c (gd_body, curve_gradient_descent): bit-identical.  A failover, not a hot path: one thread
// per walk.
#pragma once
#include "net_eval.cuh"

namespace tnb {

constexpr int kGdMaxSteps = 500;                       // subpoly_debug.py:144
constexpr int kGdWords = (kGdMaxSteps + 31) / 32;      // 16
constexpr int kGdCap = 4096;                           // walks per hyperplane (more: reported as a capacity error)
// layout of the int scratch block: [0] walks filed, [1] bodies executed (pass B), [2] 1 = every walk ended within eps,
// [4 .. 4+16) the AND words, [kGdHead + 5 i ...) walk i = (candidate, plane column, x0 bits * 3)
enum { GD_COUNT = 0, GD_BODIES = 1, GD_OK = 2, GD_MASK = 4, kGdHead = 32 };
constexpr int kGdInts = kGdHead + 5 * kGdCap;

// one body of the loop (subpoly_debug.py:145-151): d = the two distances at the x it started from
template <class C>
static __device__ __noinline__ void gd_body(const NetMeta &n, const float e0[3], const float e1[3], int colA, int colB, float x[3], float d[2])
{
    float xw[3], g[3], gx[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float t = x[k] * (e1[k] - e0[k]);
        xw[k] = e0[k] + t;
    }
    pair_grad<C>(n, xw, colA, colB, d, g);
#pragma unroll
    for (int k = 0; k < 3; ++k) gx[k] = g[k] * (e1[k] - e0[k]);
    float s = gx[0] * gx[0];
    s = __fmaf_rn(gx[1], gx[1], s);
    s = __fmaf_rn(gx[2], gx[2], s);
    float nrm = __fsqrt_rn(s);
    if (!(nrm > 1e-12f)) nrm = 1e-12f;  // F.normalize: v / max(|v|, 1e-12)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float q = __fdiv_rn(gx[k], nrm);
        float v = x[k] - 0.01f * q;
        v = v < 0.0f ? 0.0f : v;
        v = v > 1.0f ? 1.0f : v;
        x[k] = v;
    }
}

// pass A for one walk: AND its "within eps at step i" bits into the shared words
template <class C>
static __device__ void gd_note_steps(const NetMeta &n, const float e0[3], const float e1[3], int plane, int idx, float eps,
                                     const float x0[3], int *__restrict__ words)
{
    float x[3] = {x0[0], x0[1], x0[2]}, d[2];
    uint32_t w = 0;
#pragma unroll 1
    for (int i = 0; i < kGdMaxSteps; ++i) {
        gd_body<C>(n, e0, e1, plane, idx, x, d);
        if (!(fabsf(d[0]) > eps || fabsf(d[1]) > eps)) w |= 1u << (i & 31);
        if ((i & 31) == 31 || i == kGdMaxSteps - 1) {
            atomicAnd(words + (i >> 5), (int)w);
            w = 0;
        }
    }
}
// number of bodies the shared loop executes, from the AND words; ok = it ended with every walk within eps
__device__ __forceinline__ int gd_bodies(const int *__restrict__ words, bool &ok)
{
    for (int k = 0; k < kGdWords; ++k) {
        const uint32_t w = (uint32_t)words[k];
        if (w) {
            const int t = 32 * k + __ffs(w) - 1;
            if (t < kGdMaxSteps) { ok = true; return t + 1; }
        }
    }
    ok = false;
    return kGdMaxSteps;
}
// pass B for one walk
template <class C>
static __device__ void gd_walk(const NetMeta &n, const float e0[3], const float e1[3], int plane, int idx, int bodies, float x[3], float d[2])
{
#pragma unroll 1
    for (int i = 0; i < bodies; ++i) gd_body<C>(n, e0, e1, plane, idx, x, d);
}

// ---- stage-level form (tnb_curve_gradient_descent): walks given as arrays --------------------------------
template <class C>
__global__ void __launch_bounds__(64) k_gd_stage_note(const __grid_constant__ NetMeta n, const float *__restrict__ e, const float *__restrict__ x0,
                                                      const int *__restrict__ plane, int idx, float eps, int64_t G, int *__restrict__ gd)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < G; r += (int64_t)gridDim.x * blockDim.x) {
        const float e0[3] = {e[6 * r], e[6 * r + 1], e[6 * r + 2]}, e1[3] = {e[6 * r + 3], e[6 * r + 4], e[6 * r + 5]};
        const float xs[3] = {x0[3 * r], x0[3 * r + 1], x0[3 * r + 2]};
        gd_note_steps<C>(n, e0, e1, plane[r], idx, eps, xs, gd + GD_MASK);
    }
}
template <class C>
__global__ void __launch_bounds__(64) k_gd_stage_walk(const __grid_constant__ NetMeta n, const float *__restrict__ e, float *__restrict__ x,
                                                      const int *__restrict__ plane, int idx, int64_t G, float *__restrict__ d_out, int *__restrict__ gd)
{
    bool ok;
    const int bodies = gd_bodies(gd + GD_MASK, ok);
    if (blockIdx.x == 0 && threadIdx.x == 0) { gd[GD_BODIES] = bodies; gd[GD_OK] = ok ? 1 : 0; }
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < G; r += (int64_t)gridDim.x * blockDim.x) {
        const float e0[3] = {e[6 * r], e[6 * r + 1], e[6 * r + 2]}, e1[3] = {e[6 * r + 3], e[6 * r + 4], e[6 * r + 5]};
        float xs[3] = {x[3 * r], x[3 * r + 1], x[3 * r + 2]}, d[2] = {0.0f, 0.0f};
        gd_walk<C>(n, e0, e1, plane[r], idx, bodies, xs, d);
#pragma unroll
        for (int k = 0; k < 3; ++k) x[3 * r + k] = xs[k];
        d_out[2 * r] = d[0];
        d_out[2 * r + 1] = d[1];
    }
}
static __global__ void k_gd_reset(int *__restrict__ gd)
{
    const int t = threadIdx.x;
    if (t < GD_MASK) gd[t] = 0;
    else if (t < GD_MASK + kGdWords) gd[t] = -1;
}

}  // namespace tnb

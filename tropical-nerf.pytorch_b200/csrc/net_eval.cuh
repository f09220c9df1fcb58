// net_eval.cuh -- per-point evaluation of the trilinear network (hash grid fused with
// the ReLU MLP), forward and input-gradient.  Replaces, for the mesh-extraction path,
// TropicalHashGrid.forward (tropical.py:46) + Net.forward(gather=True) (model.py:52-76)
// + the autograd x-gradient of Net.sdf (tropical.py:190-195, model.py:105-123).
//
// Cfg<L,H,NLIN> fixes the sizes at compile time (all loops unroll, the MLP weights are
// read as constant-bank FFMA operands out of the kernel parameter block); Cfg<0,0,0> is
// the runtime-sized fallback for other Net(...) shapes.
#pragma once
#include "common.cuh"

namespace tnb {

template <int kL, int kH, int kNLIN>
struct Cfg {
    static constexpr bool kFixed = kL > 0;
    static constexpr int kMaxL = kFixed ? kL : kMaxLevels;
    static constexpr int kMaxH = kFixed ? kH : kMaxHidden;
    static constexpr int kMaxLin = kFixed ? kNLIN : kMaxLinear;
    static constexpr int kMaxW = (2 * kMaxL > kMaxH) ? 2 * kMaxL : kMaxH;  // widest activation
    static constexpr int kUnroll = kFixed ? 64 : 1;  // full unroll only for compile-time sizes
    __device__ __forceinline__ static int L(const NetMeta &n) { return kFixed ? kL : n.L; }
    __device__ __forceinline__ static int H(const NetMeta &n) { return kFixed ? kH : n.H; }
    __device__ __forceinline__ static int NLIN(const NetMeta &n) { return kFixed ? kNLIN : n.NLIN; }
    __device__ __forceinline__ static float w(const NetMeta &n, int i)
    {
        if (kFixed) return n.mlp_c[i];
        return __ldg(n.mlp + i);
    }
    __device__ __forceinline__ static int nin(const NetMeta &n, int i) { return i == 0 ? 2 * L(n) : H(n); }
    __device__ __forceinline__ static int nout(const NetMeta &n, int i) { return i == NLIN(n) - 1 ? 2 : H(n); }
};

using CfgRef = Cfg<4, 16, 3>;  // every network the reference builds (train.py:82)
using CfgAny = Cfg<0, 0, 0>;

// Forward pass from grid coordinates.  pre[(i*maxH)+j] = pre-activation j of hidden
// layer i; o[2] = last layer.
template <class C>
__device__ __forceinline__ void forward(const NetMeta &n, const float xp[3],
                                        float *__restrict__ pre, float o[2])
{
    float act[C::kMaxW];
#pragma unroll(C::kUnroll)
    for (int l = 0; l < C::kMaxL; ++l) {
        if (l < C::L(n)) {
            uint32_t cell[3];
            float frac[3];
            float2 f = encode_level(n, l, xp, cell, frac);
            act[2 * l] = f.x;
            act[2 * l + 1] = f.y;
        }
    }
    int base = 0;
#pragma unroll(C::kUnroll)
    for (int i = 0; i < C::kMaxLin; ++i) {
        if (i < C::NLIN(n)) {
            const int ni = C::nin(n, i), no = C::nout(n, i);
            const bool last = i == C::NLIN(n) - 1;
            float nxt[C::kMaxH];
#pragma unroll(C::kUnroll)
            for (int j = 0; j < C::kMaxH; ++j) {
                if (j < no) {
                    float acc = C::w(n, base + no * ni + j);
#pragma unroll(C::kUnroll)
                    for (int c = 0; c < C::kMaxW; ++c)
                        if (c < ni) acc = __fmaf_rn(act[c], C::w(n, base + j * ni + c), acc);
                    nxt[j] = acc;
                }
            }
            if (last) {
                o[0] = nxt[0];
                o[1] = nxt[1];
            } else {
#pragma unroll(C::kUnroll)
                for (int j = 0; j < C::kMaxH; ++j)
                    if (j < no) {
                        pre[i * C::kMaxH + j] = nxt[j];
                        act[j] = nxt[j] > 0.0f ? nxt[j] : 0.0f;
                    }
            }
            base += no * ni + no;
        }
    }
}

// Forward pass that keeps only what the backward pass needs: the ReLU masks as bits
// (mask[i] bit j = pre-activation j of hidden layer i is > 0) instead of the rows.
template <class C>
__device__ __forceinline__ void forward_masks(const NetMeta &n, const float xp[3], uint64_t mask[C::kMaxLin],
                                              float o[2])
{
    float act[C::kMaxW];
#pragma unroll(C::kUnroll)
    for (int l = 0; l < C::kMaxL; ++l) {
        if (l < C::L(n)) {
            uint32_t cell[3];
            float frac[3];
            float2 f = encode_level(n, l, xp, cell, frac);
            act[2 * l] = f.x;
            act[2 * l + 1] = f.y;
        }
    }
    int base = 0;
#pragma unroll(C::kUnroll)
    for (int i = 0; i < C::kMaxLin; ++i) {
        if (i < C::NLIN(n)) {
            const int ni = C::nin(n, i), no = C::nout(n, i);
            const bool last = i == C::NLIN(n) - 1;
            float nxt[C::kMaxH];
#pragma unroll(C::kUnroll)
            for (int j = 0; j < C::kMaxH; ++j) {
                if (j < no) {
                    float acc = C::w(n, base + no * ni + j);
#pragma unroll(C::kUnroll)
                    for (int c = 0; c < C::kMaxW; ++c)
                        if (c < ni) acc = __fmaf_rn(act[c], C::w(n, base + j * ni + c), acc);
                    nxt[j] = acc;
                }
            }
            if (last) {
                o[0] = nxt[0];
                o[1] = nxt[1];
            } else {
                uint64_t mk = 0;
#pragma unroll(C::kUnroll)
                for (int j = 0; j < C::kMaxH; ++j)
                    if (j < no) {
                        const bool on = nxt[j] > 0.0f;
                        mk |= (uint64_t)on << j;
                        act[j] = on ? nxt[j] : 0.0f;
                    }
                mask[i] = mk;
            }
            base += no * ni + no;
        }
    }
}

// d(level features)/d xp for all three axes from ONE fetch of the cell's 8 corners
// (tiny-cuda-nn kernel_grid_backward_input: for axis d, sum over the 4 corner pairs that
// differ in d of scale * prod(other weights) * (right - left)).  Accumulates
// g0 * d f0 + g1 * d f1 into acc[d] in the oracle's order.
// sum over the 4 corner pairs that differ in axis d of scale * prod(other weights) * (right - left)
__device__ __forceinline__ float2 level_dx(const float2 v[8], const float frac[3], float scale, int d)
{
    float2 dl = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int idx = 0; idx < 4; ++idx) {
        float w = scale;
        int corner = 0;
#pragma unroll
        for (int nd = 0; nd < 2; ++nd) {
            const int dim = nd >= d ? nd + 1 : nd;
            const int bit = (idx >> nd) & 1;
            w = w * (bit ? frac[dim] : 1.0f - frac[dim]);
            corner |= bit << dim;
        }
        const float2 vl = v[corner], vr = v[corner | (1 << d)];
        // both features at once: packed subtract (a + (-b) IS a - b in IEEE arithmetic) and packed FMA
        dl = __ffma2_rn(make_float2(w, w), __fadd2_rn(vr, make_float2(-vl.x, -vl.y)), dl);
    }
    return dl;
}
// one axis of a level through the general index routine, out of line (see corner_indices_fast)
static __device__ __noinline__ float2 level_dx_general(const float2 *tab, uint32_t size, uint32_t res, float scale, int d, uint32_t cx,
                                                       uint32_t cy, uint32_t cz, float f0, float f1, float f2)
{
    const float frac[3] = {f0, f1, f2};
    float2 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = __ldg(tab + grid_index(size, res, cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1)));
    if (d == 0) return level_dx(v, frac, scale, 0);
    if (d == 1) return level_dx(v, frac, scale, 1);
    return level_dx(v, frac, scale, 2);
}
__device__ __forceinline__ void encode_level_grad(const NetMeta &n, int l, const float xp[3], float g0, float g1,
                                                  float acc[3])
{
    const LevelMeta lv = n.lvl[l];
    uint32_t cell[3];
    float frac[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float pos = __fmaf_rn(lv.scale, xp[d], 0.5f);
        float fl = floorf(pos);
        cell[d] = (uint32_t)(int)fl;
        frac[d] = pos - fl;
    }
    const float2 *tab = n.table + lv.off;
    uint32_t cidx[8];
    if (!corner_indices_fast(lv, cell[0], cell[1], cell[2], cidx)) {
#pragma unroll 1
        for (int d = 0; d < 3; ++d) {
            const float2 dl = level_dx_general(tab, lv.size, lv.res, lv.scale, d, cell[0], cell[1], cell[2], frac[0], frac[1], frac[2]);
            const float a = __fmaf_rn(g0, dl.x, d == 0 ? acc[0] : d == 1 ? acc[1] : acc[2]);
            const float b = __fmaf_rn(g1, dl.y, a);
            if (d == 0) acc[0] = b; else if (d == 1) acc[1] = b; else acc[2] = b;
        }
        return;
    }
    float2 v[8];
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) v[corner] = __ldg(tab + cidx[corner]);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float2 dl = level_dx(v, frac, lv.scale, d);
        acc[d] = __fmaf_rn(g0, dl.x, acc[d]);
        acc[d] = __fmaf_rn(g1, dl.y, acc[d]);
    }
}

// ---- two points per thread (the lattice sweeps) ---------------------------------------------------
// Blackwell's packed FFMA2 (fma.rn.f32x2) does two independent IEEE binary32 FMAs per issue slot, and it
// takes a SCALAR second operand that is broadcast to both halves.  With two lattice points per thread,
// a pair = (value at point 0, value at point 1), and the MLP weight -- a uniform-register / constant
// operand -- as the broadcast scalar, every FMA of the network advances both points at once, each half
// bit-identical to the __fmaf_rn chain of the one-point code (and of oracle/trinet_ref.c).  The sweeps
// are fp32-issue bound: this halves their dominant instruction class.
template <class C>
__device__ __forceinline__ void encode_pair(const NetMeta &n, const float xp0[3], const float xp1[3], float2 act[C::kMaxW])
{
#pragma unroll(C::kUnroll)
    for (int l = 0; l < C::kMaxL; ++l) {
        if (l < C::L(n)) {
            uint32_t cell[3];
            float frac[3];
            const float2 f0 = encode_level(n, l, xp0, cell, frac);
            const float2 f1 = encode_level(n, l, xp1, cell, frac);
            act[2 * l] = make_float2(f0.x, f1.x);
            act[2 * l + 1] = make_float2(f0.y, f1.y);
        }
    }
}
// one nn.Linear for both points: nxt[j] = b_j + sum_c act[c] * W[j][c], c ascending (the oracle's order)
template <class C>
__device__ __forceinline__ void layer_pair(const NetMeta &n, int base, int ni, int no, const float2 act[C::kMaxW], float2 nxt[C::kMaxH])
{
#pragma unroll(C::kUnroll)
    for (int j = 0; j < C::kMaxH; ++j) {
        if (j < no) {
            const float b = C::w(n, base + no * ni + j);
            float2 acc = make_float2(b, b);
#pragma unroll(C::kUnroll)
            for (int c = 0; c < C::kMaxW; ++c)
                if (c < ni) {
                    const float w = C::w(n, base + j * ni + c);
                    acc = __ffma2_rn(act[c], make_float2(w, w), acc);
                }
            nxt[j] = acc;
        }
    }
}
// forward<C> for two points: pre[(i*maxH)+j] = pre-activation j of hidden layer i, o[k] = last layer
template <class C>
__device__ __forceinline__ void forward_pair(const NetMeta &n, const float xp0[3], const float xp1[3], float2 *__restrict__ pre, float2 o[2])
{
    float2 act[C::kMaxW];
    encode_pair<C>(n, xp0, xp1, act);
    int base = 0;
#pragma unroll(C::kUnroll)
    for (int i = 0; i < C::kMaxLin; ++i) {
        if (i < C::NLIN(n)) {
            const int ni = C::nin(n, i), no = C::nout(n, i);
            float2 nxt[C::kMaxH];
            layer_pair<C>(n, base, ni, no, act, nxt);
            if (i == C::NLIN(n) - 1) {
                o[0] = nxt[0];
                o[1] = nxt[1];
            } else {
#pragma unroll(C::kUnroll)
                for (int j = 0; j < C::kMaxH; ++j)
                    if (j < no) {
                        pre[i * C::kMaxH + j] = nxt[j];
                        act[j] = make_float2(nxt[j].x > 0.0f ? nxt[j].x : 0.0f, nxt[j].y > 0.0f ? nxt[j].y : 0.0f);
                    }
            }
            base += no * ni + no;
        }
    }
}
// sdf_grad<C> for two points: t[k] = tanh(o1 - o0), grad[k][3] = d t / d x of point k
template <class C>
__device__ __forceinline__ void sdf_grad_pair(const NetMeta &n, const float x0[3], const float x1[3], float t[2], float grad[2][3])
{
    float xp0[3], xp1[3];
    preprocess(n, x0, xp0);
    preprocess(n, x1, xp1);
    uint64_t mask0[C::kMaxLin], mask1[C::kMaxLin];
    float2 o[2];
    {
        float2 act[C::kMaxW];
        encode_pair<C>(n, xp0, xp1, act);
        int base = 0;
#pragma unroll(C::kUnroll)
        for (int i = 0; i < C::kMaxLin; ++i) {
            if (i < C::NLIN(n)) {
                const int ni = C::nin(n, i), no = C::nout(n, i);
                float2 nxt[C::kMaxH];
                layer_pair<C>(n, base, ni, no, act, nxt);
                if (i == C::NLIN(n) - 1) {
                    o[0] = nxt[0];
                    o[1] = nxt[1];
                } else {
                    uint64_t m0 = 0, m1 = 0;
#pragma unroll(C::kUnroll)
                    for (int j = 0; j < C::kMaxH; ++j)
                        if (j < no) {
                            const bool on0 = nxt[j].x > 0.0f, on1 = nxt[j].y > 0.0f;
                            m0 |= (uint64_t)on0 << j;
                            m1 |= (uint64_t)on1 << j;
                            act[j] = make_float2(on0 ? nxt[j].x : 0.0f, on1 ? nxt[j].y : 0.0f);
                        }
                    mask0[i] = m0;
                    mask1[i] = m1;
                }
                base += no * ni + no;
            }
        }
    }
    t[0] = det_tanhf(o[1].x - o[0].x);
    t[1] = det_tanhf(o[1].y - o[0].y);
    const float gs0 = 1.0f - t[0] * t[0], gs1 = 1.0f - t[1] * t[1];
    float2 g_out[C::kMaxW], g_in[C::kMaxW];
    g_out[0] = make_float2(-gs0, -gs1);
    g_out[1] = make_float2(gs0, gs1);
    int base[C::kMaxLin];
    {
        int b = 0;
#pragma unroll(C::kUnroll)
        for (int i = 0; i < C::kMaxLin; ++i)
            if (i < C::NLIN(n)) { base[i] = b; b += C::nout(n, i) * C::nin(n, i) + C::nout(n, i); }
    }
#pragma unroll(C::kUnroll)
    for (int k = C::kMaxLin - 1; k >= 0; --k) {
        if (k < C::NLIN(n)) {
            const int ni = C::nin(n, k), no = C::nout(n, k);
#pragma unroll(C::kUnroll)
            for (int c = 0; c < C::kMaxW; ++c) {
                if (c < ni) {
                    float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll(C::kUnroll)
                    for (int j = 0; j < C::kMaxH; ++j)
                        if (j < no) {
                            const float w = C::w(n, base[k] + j * ni + c);
                            acc = __ffma2_rn(make_float2(w, w), g_out[j], acc);
                        }
                    g_in[c] = acc;
                }
            }
            if (k > 0) {
#pragma unroll(C::kUnroll)
                for (int c = 0; c < C::kMaxH; ++c)
                    if (c < ni) g_out[c] = make_float2(((mask0[k - 1] >> c) & 1) ? g_in[c].x : 0.0f, ((mask1[k - 1] >> c) & 1) ? g_in[c].y : 0.0f);
            }
        }
    }
    float acc0[3] = {0.0f, 0.0f, 0.0f}, acc1[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll(C::kUnroll)
    for (int l = 0; l < C::kMaxL; ++l)
        if (l < C::L(n)) {
            encode_level_grad(n, l, xp0, g_in[2 * l].x, g_in[2 * l + 1].x, acc0);
            encode_level_grad(n, l, xp1, g_in[2 * l].y, g_in[2 * l + 1].y, acc1);
        }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        grad[0][d] = div_2s(n, acc0[d]);
        grad[1][d] = div_2s(n, acc1[d]);
    }
}

// tanh(o1-o0) and its gradient w.r.t. the world-space input.
template <class C>
__device__ __forceinline__ float sdf_grad(const NetMeta &n, const float x[3], float grad[3],
                                          bool want_grad)
{
    float xp[3];
    preprocess(n, x, xp);
    uint64_t mask[C::kMaxLin];
    float o[2];
    forward_masks<C>(n, xp, mask, o);
    const float t = det_tanhf(o[1] - o[0]);
    if (!want_grad) return t;
    const float gs = 1.0f - t * t;
    float g_out[C::kMaxW], g_in[C::kMaxW];
    g_out[0] = -gs;
    g_out[1] = gs;
    // offsets of the packed layers
    int base[C::kMaxLin];
    {
        int b = 0;
#pragma unroll(C::kUnroll)
        for (int i = 0; i < C::kMaxLin; ++i)
            if (i < C::NLIN(n)) { base[i] = b; b += C::nout(n, i) * C::nin(n, i) + C::nout(n, i); }
    }
#pragma unroll(C::kUnroll)
    for (int k = C::kMaxLin - 1; k >= 0; --k) {
        if (k < C::NLIN(n)) {
            const int ni = C::nin(n, k), no = C::nout(n, k);
#pragma unroll(C::kUnroll)
            for (int c = 0; c < C::kMaxW; ++c) {
                if (c < ni) {
                    float acc = 0.0f;
#pragma unroll(C::kUnroll)
                    for (int j = 0; j < C::kMaxH; ++j)
                        if (j < no) acc = __fmaf_rn(C::w(n, base[k] + j * ni + c), g_out[j], acc);
                    g_in[c] = acc;
                }
            }
            if (k > 0) {
#pragma unroll(C::kUnroll)
                for (int c = 0; c < C::kMaxH; ++c)
                    if (c < ni) g_out[c] = ((mask[k - 1] >> c) & 1) ? g_in[c] : 0.0f;
            }
        }
    }
    float acc[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll(C::kUnroll)
    for (int l = 0; l < C::kMaxL; ++l)
        if (l < C::L(n)) encode_level_grad(n, l, xp, g_in[2 * l], g_in[2 * l + 1], acc);
#pragma unroll(C::kUnroll)
    for (int d = 0; d < 3; ++d) grad[d] = div_2s(n, acc[d]);
    return t;
}

// y = dA^2 + dB^2 with dA, dB = output columns colA, colB at x; grad = dy/dx in world space (what autograd
// hands the gradient-descent repair of the curve path, subpoly_debug.py:147-149).  Same reverse sweep and
// operation order as oracle/trinet_ref.c pair_grad: the seeds 2 dA, 2 dB enter at their pre-activations (the
// last column o1 - o0 seeds the last layer with -s, +s).  A rare path: out of line, arrays in local memory.
template <class C>
static __device__ __noinline__ void pair_grad(const NetMeta &n, const float x[3], int colA, int colB, float d[2], float grad[3])
{
    float xp[3];
    preprocess(n, x, xp);
    float pre[(C::kMaxLin - 1) * C::kMaxH];
    float o[2];
    forward<C>(n, xp, pre, o);
    const int H = C::H(n), R = n.R, last = C::NLIN(n) - 1;
    const int col[2] = {colA, colB};
    float seed[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        d[t] = col[t] == R - 1 ? o[1] - o[0] : pre[(col[t] / H) * C::kMaxH + col[t] % H];
        seed[t] = 2.0f * d[t];
    }
    float g_out[C::kMaxW], g_in[C::kMaxW];
#pragma unroll(C::kUnroll)
    for (int j = 0; j < C::kMaxW; ++j) g_out[j] = 0.0f;
    int base[C::kMaxLin];
    {
        int b = 0;
#pragma unroll(C::kUnroll)
        for (int i = 0; i < C::kMaxLin; ++i)
            if (i < C::NLIN(n)) { base[i] = b; b += C::nout(n, i) * C::nin(n, i) + C::nout(n, i); }
    }
#pragma unroll 1
    for (int k = last; k >= 0; --k) {
        const int ni = C::nin(n, k), no = C::nout(n, k);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (col[t] == R - 1) {
                if (k == last) { g_out[0] = g_out[0] - seed[t]; g_out[1] = g_out[1] + seed[t]; }
            } else if (col[t] / H == k) {
                g_out[col[t] % H] = g_out[col[t] % H] + seed[t];
            }
        }
        for (int c = 0; c < ni; ++c) {
            float acc = 0.0f;
            for (int j = 0; j < no; ++j) acc = __fmaf_rn(C::w(n, base[k] + j * ni + c), g_out[j], acc);
            g_in[c] = acc;
        }
        if (k > 0)
            for (int c = 0; c < ni; ++c) g_out[c] = pre[(k - 1) * C::kMaxH + c] > 0.0f ? g_in[c] : 0.0f;
    }
    float acc[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll 1
    for (int l = 0; l < C::L(n); ++l) encode_level_grad(n, l, xp, g_in[2 * l], g_in[2 * l + 1], acc);
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) grad[dd] = div_2s(n, acc[dd]);
}

__device__ __forceinline__ float grad_norm(const float g[3])
{
    float s = g[0] * g[0];
    s = __fmaf_rn(g[1], g[1], s);
    s = __fmaf_rn(g[2], g[2], s);
    return __fsqrt_rn(s);
}

// One row of torch.cat(Net.forward(x, gather=True)[1], -1): R floats to `row`
// (stride 1), from world coordinates.
template <class C>
__device__ __forceinline__ void outputs_row(const NetMeta &n, const float x[3], float *__restrict__ row)
{
    float xp[3];
    preprocess(n, x, xp);
    float pre[(C::kMaxLin - 1) * C::kMaxH];
    float o[2];
    forward<C>(n, xp, pre, o);
    const int H = C::H(n), NL = C::NLIN(n);
#pragma unroll(C::kUnroll)
    for (int i = 0; i < C::kMaxLin - 1; ++i)
        if (i < NL - 1) {
#pragma unroll(C::kUnroll)
            for (int j = 0; j < C::kMaxH; ++j)
                if (j < H) row[i * H + j] = pre[i * C::kMaxH + j];
        }
    row[(NL - 1) * H] = o[1] - o[0];
}

// sign of one output with tolerance eps >= 0 (model.py:97-98) into bit c of pos / neg:
//   pos <=> |v| > eps and v > 0 <=> v > eps;   neg <=> not(|v| <= eps) and not(v > 0) <=> not(v >= -eps)
// (the second form keeps a NaN on the negative side, like the reference's comparison chain).  One
// compare and one predicated OR on a 32-bit half each: as nested ifs on 64-bit words this compiled
// to a BSSY / BRA / BSYNC triple per neuron (a tenth of the instructions of the dense sign sweep).
struct SignWords {
    uint32_t plo = 0, phi = 0, nlo = 0, nhi = 0;
    __device__ __forceinline__ void add(float v, float eps, int c)
    {
        const uint32_t bit = 1u << (c & 31);
        const bool p = v > eps, q = !(v >= -eps);
        if (c < 32) {
            if (p) plo |= bit;
            if (q) nlo |= bit;
        } else {
            if (p) phi |= bit;
            if (q) nhi |= bit;
        }
    }
    __device__ __forceinline__ uint64_t pos() const { return ((uint64_t)phi << 32) | plo; }
    __device__ __forceinline__ uint64_t neg() const { return ((uint64_t)nhi << 32) | nlo; }
};

// outputs_row + the packed signs of the row straight from the registers (tolerance eps_sign), the grid
// coordinates, and `big`: bit c set <=> |row[c]| > eps_big
template <class C>
__device__ __forceinline__ void outputs_row_packed(const NetMeta &n, const float x[3], float *__restrict__ row, float eps_sign,
                                                   float eps_big, float xp[3], uint64_t &pos, uint64_t &neg, uint64_t &big)
{
    preprocess(n, x, xp);
    float pre[(C::kMaxLin - 1) * C::kMaxH];
    float o[2];
    forward<C>(n, xp, pre, o);
    const int H = C::H(n), NL = C::NLIN(n);
    SignWords w;
    uint32_t blo = 0, bhi = 0;
#pragma unroll(C::kUnroll)
    for (int i = 0; i < C::kMaxLin - 1; ++i)
        if (i < NL - 1) {
#pragma unroll(C::kUnroll)
            for (int j = 0; j < C::kMaxH; ++j)
                if (j < H) {
                    const float v = pre[i * C::kMaxH + j];
                    const int c = i * H + j;
                    row[c] = v;
                    w.add(v, eps_sign, c);
                    if (fabsf(v) > eps_big) { if (c < 32) blo |= 1u << (c & 31); else bhi |= 1u << (c & 31); }
                }
        }
    {
        const float v = o[1] - o[0];
        const int c = (NL - 1) * H;
        row[c] = v;
        w.add(v, eps_sign, c);
        if (fabsf(v) > eps_big) { if (c < 32) blo |= 1u << (c & 31); else bhi |= 1u << (c & 31); }
    }
    pos = w.pos();
    neg = w.neg();
    big = ((uint64_t)bhi << 32) | blo;
}

// sign bits of a row of outputs (model.py:97-98)
__device__ __forceinline__ void pack_signs(const float *__restrict__ row, int R, float eps,
                                           uint64_t &pos, uint64_t &neg)
{
    SignWords w;
    // 16 loads in flight at a time: the row usually sits in L2 (written before a grid barrier, or by
    // another kernel), and one load per loop trip made this 33 dependent round trips
    for (int c0 = 0; c0 < R; c0 += 16) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = c0 + i < R ? row[c0 + i] : 0.0f;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c0 + i < R) w.add(v[i], eps, c0 + i);
    }
    pos = w.pos();
    neg = w.neg();
}

}  // namespace tnb

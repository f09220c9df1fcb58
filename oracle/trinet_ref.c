/*
 * oracle/trinet_ref.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the piecewise-trilinear network evaluation on the
 * reference's mesh-extraction path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * shipped path (tropical-nerf.pytorch_b200/) never does.
 *
 * What it follows in the reference (/root/reference):
 *   - Net.forward(gather=True)        tropical/stanford/model.py:52-76
 *   - Net.preprocess / sdf            tropical/stanford/model.py:78-88
 *   - Net.region (sign vectors)       tropical/stanford/model.py:90-103
 *   - TropicalHashGrid.region         tropical/tropical.py:227-236
 *   - the x-gradient autograd takes in TropicalHashGrid.skeleton
 *     (tropical/tropical.py:190-195) and Net.normal (model.py:105-123)
 *   - Net.forward(gather=True, group=8) (model.py:65-70) and
 *     geometry.intersection_of_two_planes / batched_polynomial_roots
 *     (geometry.py:24-138, :259-299) for the curve-approximation path, and its
 *     gradient-descent repair (subpoly_debug.py:121-165)
 *   - TropicalHashGrid.forward -> tcnn.Encoding (tropical/tropical.py:32-47).
 *     tiny-cuda-nn is a third-party dependency that is NOT in the reference
 *     tree and NOT pinned by its requirements.txt; the multiresolution hash
 *     encoding is restated here from its published algorithm (Mueller et al.
 *     2022, Sec. 3 + App. A; tiny-cuda-nn GridEncoding: pos = x*scale+0.5,
 *     dense index below the table size, prime-XOR hash above it).
 *     => the encoding itself is "parity unpinned" against real tiny-cuda-nn;
 *     everything downstream is pinned by running the reference's own Python
 *     on top of the same encoding (tests/golden/make_golden.py).
 *
 * Floating point contract (what "bit-exact" means for the CUDA path): every
 * operation below is a single IEEE-754 binary32 operation (+, -, *, /, sqrt,
 * fma) in exactly the written order; fmaf() marks the fused ones and the file
 * is compiled with -ffp-contract=off so nothing else fuses.  tanh is NOT taken
 * from libm (glibc and CUDA differ in the last ulp): det_tanhf() below is the
 * definition, built from the same primitive operations.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define TN_MAX_LEVELS 16
#define TN_MAX_WIDTH 64
#define TN_MAX_LINEAR 8

typedef struct {
    int32_t n_levels;   /* L */
    int32_t n_feat;     /* F (features per level) */
    int32_t n_linear;   /* number of nn.Linear layers (= Net.num_layers) */
    int32_t n_hidden;   /* H */
    float pre_scale;    /* Net.scale */
    const float *lvl_scale;    /* [L] */
    const uint32_t *lvl_res;   /* [L] */
    const uint32_t *lvl_size;  /* [L] table entries of the level */
    const uint32_t *lvl_off;   /* [L] first entry of the level */
    const float *table;        /* [sum(size) * F] */
    const float *mlp;          /* per layer: W[out][in] row-major, then b[out] */
} trinet_t;

static int layer_in(const trinet_t *n, int i) { return i == 0 ? n->n_levels * n->n_feat : n->n_hidden; }
static int layer_out(const trinet_t *n, int i) { return i == n->n_linear - 1 ? 2 : n->n_hidden; }

int trinet_n_outputs(const trinet_t *n) { return (n->n_linear - 1) * n->n_hidden + 1; }

/* ---- deterministic transcendental ------------------------------------------------ */
static float det_expf(float y)
{
    float n = rintf(y * 1.44269504088896341f);
    float r = fmaf(n, -0.693359375f, y);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float e = fmaf(p, r * r, r) + 1.0f;
    union { float f; uint32_t u; } s;
    s.u = (uint32_t)((int32_t)n + 127) << 23;
    return e * s.f;
}

float det_tanhf(float x)
{
    float ax = fabsf(x);
    float t;
    if (ax < 0.25f) {
        float x2 = ax * ax;
        float p = 0.021869488536155203f;            /* 62/2835 */
        p = fmaf(p, x2, -0.053968253968253971f);    /* -17/315 */
        p = fmaf(p, x2, 0.13333333333333333f);      /* 2/15 */
        p = fmaf(p, x2, -0.33333333333333331f);     /* -1/3 */
        t = fmaf(ax * x2, p, ax);
    } else if (ax > 9.0f) {
        t = 1.0f;
    } else {
        float e = det_expf(2.0f * ax);
        t = 1.0f - 2.0f / (e + 1.0f);
    }
    return copysignf(t, x);
}

/* ---- hash grid ------------------------------------------------------------------- */
static uint32_t grid_index(uint32_t size, uint32_t res, const uint32_t c[3])
{
    uint32_t stride = 1, index = 0;
    for (int d = 0; d < 3 && stride <= size; ++d) {
        index += c[d] * stride;
        stride *= res;
    }
    if (size < stride)
        index = (c[0] * 1u) ^ (c[1] * 2654435761u) ^ (c[2] * 805459861u);
    return index % size;
}

typedef struct {
    uint32_t cell[TN_MAX_LEVELS][3];
    float frac[TN_MAX_LEVELS][3];
} enc_ctx_t;

/* xp: preprocessed point in hash-grid coordinates */
static void encode(const trinet_t *n, const float xp[3], float *enc, enc_ctx_t *ctx)
{
    const int F = n->n_feat;
    for (int l = 0; l < n->n_levels; ++l) {
        const float scale = n->lvl_scale[l];
        uint32_t cell[3];
        float frac[3];
        for (int d = 0; d < 3; ++d) {
            float pos = fmaf(scale, xp[d], 0.5f);
            float fl = floorf(pos);
            cell[d] = (uint32_t)(int32_t)fl;
            frac[d] = pos - fl;
            if (ctx) { ctx->cell[l][d] = cell[d]; ctx->frac[l][d] = frac[d]; }
        }
        const float *tab = n->table + (size_t)n->lvl_off[l] * F;
        for (int f = 0; f < F; ++f) enc[l * F + f] = 0.0f;
        for (int corner = 0; corner < 8; ++corner) {
            float w = 1.0f;
            uint32_t c[3];
            for (int d = 0; d < 3; ++d) {
                if (corner & (1 << d)) { w = w * frac[d]; c[d] = cell[d] + 1u; }
                else                   { w = w * (1.0f - frac[d]); c[d] = cell[d]; }
            }
            uint32_t idx = grid_index(n->lvl_size[l], n->lvl_res[l], c);
            for (int f = 0; f < F; ++f)
                enc[l * F + f] = fmaf(w, tab[(size_t)idx * F + f], enc[l * F + f]);
        }
    }
}

/* d enc[l*F+f] / d xp[d]  (tiny-cuda-nn kernel_grid_backward_input, linear interpolation) */
static void encode_dx(const trinet_t *n, const enc_ctx_t *ctx, int l, int d, float *dl)
{
    const int F = n->n_feat;
    const float *tab = n->table + (size_t)n->lvl_off[l] * F;
    for (int f = 0; f < F; ++f) dl[f] = 0.0f;
    for (int idx = 0; idx < 4; ++idx) {
        float w = n->lvl_scale[l];
        uint32_t c[3];
        for (int nd = 0; nd < 2; ++nd) {
            int dim = nd >= d ? nd + 1 : nd;
            if (idx & (1 << nd)) { w = w * ctx->frac[l][dim]; c[dim] = ctx->cell[l][dim] + 1u; }
            else                 { w = w * (1.0f - ctx->frac[l][dim]); c[dim] = ctx->cell[l][dim]; }
        }
        c[d] = ctx->cell[l][d];
        uint32_t il = grid_index(n->lvl_size[l], n->lvl_res[l], c);
        c[d] = ctx->cell[l][d] + 1u;
        uint32_t ir = grid_index(n->lvl_size[l], n->lvl_res[l], c);
        for (int f = 0; f < F; ++f)
            dl[f] = fmaf(w, tab[(size_t)ir * F + f] - tab[(size_t)il * F + f], dl[f]);
    }
}

void trinet_encode(const trinet_t *n, const float *xp, int64_t count, float *enc)
{
    const int W = n->n_levels * n->n_feat;
    for (int64_t i = 0; i < count; ++i) encode(n, xp + 3 * i, enc + (size_t)W * i, 0);
}

/* ---- MLP ------------------------------------------------------------------------- */
/* pre[i][j] keeps the pre-activation of layer i; returns o[2] */
static void mlp_forward(const trinet_t *n, const float *enc, float pre[TN_MAX_LINEAR][TN_MAX_WIDTH])
{
    float act[TN_MAX_WIDTH];
    const float *p = n->mlp;
    int nin = layer_in(n, 0);
    for (int c = 0; c < nin; ++c) act[c] = enc[c];
    for (int i = 0; i < n->n_linear; ++i) {
        nin = layer_in(n, i);
        const int nout = layer_out(n, i);
        const float *W = p, *b = p + (size_t)nout * nin;
        for (int j = 0; j < nout; ++j) {
            float acc = b[j];
            for (int c = 0; c < nin; ++c) acc = fmaf(act[c], W[j * nin + c], acc);
            pre[i][j] = acc;
        }
        if (i != n->n_linear - 1)
            for (int j = 0; j < nout; ++j) act[j] = pre[i][j] > 0.0f ? pre[i][j] : 0.0f;
        p = b + nout;
    }
}

static void preprocess(const trinet_t *n, const float *x, float xp[3])
{
    for (int d = 0; d < 3; ++d) xp[d] = (x[d] + n->pre_scale) / (n->pre_scale * 2.0f);
}

/* Net.forward(gather=True)[1] concatenated: hidden pre-activations, then o1 - o0 */
void trinet_outputs(const trinet_t *n, const float *x, int64_t count, float *out)
{
    const int R = trinet_n_outputs(n), H = n->n_hidden;
    for (int64_t i = 0; i < count; ++i) {
        float xp[3], enc[TN_MAX_LEVELS * 8], pre[TN_MAX_LINEAR][TN_MAX_WIDTH];
        preprocess(n, x + 3 * i, xp);
        encode(n, xp, enc, 0);
        mlp_forward(n, enc, pre);
        float *o = out + (size_t)R * i;
        for (int k = 0; k < n->n_linear - 1; ++k)
            for (int j = 0; j < H; ++j) o[k * H + j] = pre[k][j];
        o[R - 1] = pre[n->n_linear - 1][1] - pre[n->n_linear - 1][0];
    }
}

/* Net.sdf(x)[:,0] = tanh(o1-o0) and its gradient w.r.t. the world-space input */
void trinet_sdf_grad(const trinet_t *n, const float *x, int64_t count, float *sdf, float *grad)
{
    const int F = n->n_feat;
    for (int64_t i = 0; i < count; ++i) {
        float xp[3], enc[TN_MAX_LEVELS * 8], pre[TN_MAX_LINEAR][TN_MAX_WIDTH];
        enc_ctx_t ctx;
        preprocess(n, x + 3 * i, xp);
        encode(n, xp, enc, &ctx);
        mlp_forward(n, enc, pre);
        const int last = n->n_linear - 1;
        float t = det_tanhf(pre[last][1] - pre[last][0]);
        sdf[i] = t;
        if (!grad) continue;
        float gs = 1.0f - t * t;
        float g_out[TN_MAX_WIDTH], g_in[TN_MAX_WIDTH];
        g_out[0] = -gs; g_out[1] = gs;
        /* locate the packed weights of each layer */
        const float *Wp[TN_MAX_LINEAR];
        const float *p = n->mlp;
        for (int k = 0; k < n->n_linear; ++k) {
            Wp[k] = p;
            p += (size_t)layer_out(n, k) * layer_in(n, k) + layer_out(n, k);
        }
        for (int k = last; k >= 0; --k) {
            const int nin = layer_in(n, k), nout = layer_out(n, k);
            for (int c = 0; c < nin; ++c) {
                float acc = 0.0f;
                for (int j = 0; j < nout; ++j) acc = fmaf(Wp[k][j * nin + c], g_out[j], acc);
                g_in[c] = acc;
            }
            if (k > 0)
                for (int c = 0; c < nin; ++c) g_out[c] = pre[k - 1][c] > 0.0f ? g_in[c] : 0.0f;
        }
        /* g_in now holds d sdf / d enc */
        for (int d = 0; d < 3; ++d) {
            float acc = 0.0f;
            for (int l = 0; l < n->n_levels; ++l) {
                float dl[8];
                encode_dx(n, &ctx, l, d, dl);
                for (int f = 0; f < F; ++f) acc = fmaf(g_in[l * F + f], dl[f], acc);
            }
            grad[3 * i + d] = acc / (n->pre_scale * 2.0f);
        }
    }
}

/* |grad| exactly as the skeleton sweep defines it (max_grad = max over a chunk of this) */
void trinet_grad_norm(const float *grad, int64_t count, float *norm)
{
    for (int64_t i = 0; i < count; ++i) {
        const float *g = grad + 3 * i;
        float s = g[0] * g[0];
        s = fmaf(g[1], g[1], s);
        s = fmaf(g[2], g[2], s);
        norm[i] = sqrtf(s);
    }
}

/* ---- gradient-descent repair of the curve path (subpoly_debug.py:121-165) -------------- */
/* y = dA^2 + dB^2 with dA, dB = columns colA, colB of trinet_outputs at x; grad = dy/dx (world
 * space), by the reverse sweep autograd takes: the seeds 2*dA, 2*dB enter at their
 * pre-activations (the last column o1 - o0 seeds the last layer with -s, +s), every layer's
 * input gradient is the fma chain over its output neurons in ascending order from 0. */
static void pair_grad(const trinet_t *n, const float *x, int colA, int colB, float *dA, float *dB, float grad[3])
{
    const int F = n->n_feat, H = n->n_hidden, R = trinet_n_outputs(n), last = n->n_linear - 1;
    float xp[3], enc[TN_MAX_LEVELS * 8], pre[TN_MAX_LINEAR][TN_MAX_WIDTH];
    enc_ctx_t ctx;
    preprocess(n, x, xp);
    encode(n, xp, enc, &ctx);
    mlp_forward(n, enc, pre);
    const float oA = colA == R - 1 ? pre[last][1] - pre[last][0] : pre[colA / H][colA % H];
    const float oB = colB == R - 1 ? pre[last][1] - pre[last][0] : pre[colB / H][colB % H];
    *dA = oA;
    *dB = oB;
    const float seed[2] = {2.0f * oA, 2.0f * oB};
    const int col[2] = {colA, colB};
    const float *Wp[TN_MAX_LINEAR];
    const float *p = n->mlp;
    for (int k = 0; k < n->n_linear; ++k) {
        Wp[k] = p;
        p += (size_t)layer_out(n, k) * layer_in(n, k) + layer_out(n, k);
    }
    float g_out[TN_MAX_WIDTH], g_in[TN_MAX_WIDTH];
    for (int j = 0; j < TN_MAX_WIDTH; ++j) g_out[j] = 0.0f;
    for (int k = last; k >= 0; --k) {
        const int nin = layer_in(n, k), nout = layer_out(n, k);
        /* seeds that enter at this layer's pre-activations (A first, then B) */
        for (int t = 0; t < 2; ++t) {
            if (col[t] == R - 1) {
                if (k == last) { g_out[0] = g_out[0] - seed[t]; g_out[1] = g_out[1] + seed[t]; }
            } else if (col[t] / H == k) {
                g_out[col[t] % H] = g_out[col[t] % H] + seed[t];
            }
        }
        for (int c = 0; c < nin; ++c) {
            float acc = 0.0f;
            for (int j = 0; j < nout; ++j) acc = fmaf(Wp[k][j * nin + c], g_out[j], acc);
            g_in[c] = acc;
        }
        if (k > 0)
            for (int c = 0; c < nin; ++c) g_out[c] = pre[k - 1][c] > 0.0f ? g_in[c] : 0.0f;
    }
    for (int d = 0; d < 3; ++d) {
        float acc = 0.0f;
        for (int l = 0; l < n->n_levels; ++l) {
            float dl[8];
            encode_dx(n, &ctx, l, d, dl);
            for (int f = 0; f < F; ++f) acc = fmaf(g_in[l * F + f], dl[f], acc);
        }
        grad[d] = acc / (n->pre_scale * 2.0f);
    }
}

/* One body of the loop at subpoly_debug.py:144-151 for one edge: x (edge parameters, one per
 * axis) -> x - 1e-2 * normalize(dy/dx), clamped to [0,1]; d[2] = the two distances at the x the
 * body STARTED from.  F.normalize divides by max(|g|, 1e-12). */
static void gd_body(const trinet_t *n, const float e0[3], const float e1[3], int colA, int colB, float x[3], float d[2])
{
    float xw[3], g[3], gx[3];
    for (int k = 0; k < 3; ++k) {
        const float t = x[k] * (e1[k] - e0[k]);
        xw[k] = e0[k] + t;
    }
    pair_grad(n, xw, colA, colB, &d[0], &d[1], g);
    for (int k = 0; k < 3; ++k) gx[k] = g[k] * (e1[k] - e0[k]);
    float s = gx[0] * gx[0];
    s = fmaf(gx[1], gx[1], s);
    s = fmaf(gx[2], gx[2], s);
    float nrm = sqrtf(s);
    if (!(nrm > 1e-12f)) nrm = 1e-12f;
    for (int k = 0; k < 3; ++k) {
        const float q = gx[k] / nrm;
        float v = x[k] - 0.01f * q;
        v = v < 0.0f ? 0.0f : v;
        v = v > 1.0f ? 1.0f : v;
        x[k] = v;
    }
}

/* deal_with_gradient_descent (subpoly_debug.py:121-165) for the `count` edges that need it:
 * ALL of them take another step while ANY of them is farther than eps from one of its two planes,
 * at most 500 steps.  x: [count][3] in/out; plane: column of the earlier plane each edge lies in;
 * d: [count][2] = distances at the x the last body started from.  Returns the number of bodies. */
int curve_gradient_descent(const trinet_t *n, const float *e0, const float *e1, float *x, const int32_t *plane,
                           int32_t idx, float eps, int64_t count, float *d)
{
    int i = 0, more = 1;
    while (more && i < 500) {
        more = 0;
        for (int64_t r = 0; r < count; ++r) {
            gd_body(n, e0 + 3 * r, e1 + 3 * r, plane[r], idx, x + 3 * r, d + 2 * r);
            if (fabsf(d[2 * r]) > eps || fabsf(d[2 * r + 1]) > eps) more = 1;
        }
        ++i;
    }
    return i;
}

/* ---- region indicators ------------------------------------------------------------ */
/* torch.searchsorted(marks, v) with right=False: first i with marks[i] >= v */
static int32_t lower_bound(const float *marks, int32_t m, float v)
{
    int32_t lo = 0, hi = m;
    while (lo < hi) {
        int32_t mid = (lo + hi) >> 1;
        if (marks[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* signs: [count][3+R] int8 = (grid mask 0/1 per axis, then -1/0/+1 per neuron);
 * offset: [count][3] int32 (can be -1, tropical.py:231) */
void trinet_region(const trinet_t *n, const float *marks, int32_t n_marks, float eps,
                   const float *x, const float *outputs, int64_t count,
                   int8_t *signs, int32_t *offset)
{
    const int R = trinet_n_outputs(n);
    for (int64_t i = 0; i < count; ++i) {
        float xp[3];
        preprocess(n, x + 3 * i, xp);
        int8_t *s = signs + (size_t)(3 + R) * i;
        for (int d = 0; d < 3; ++d) {
            int32_t off = lower_bound(marks, n_marks, xp[d] + eps) - 1;
            float mk = marks[off < 0 ? off + n_marks : off];
            s[d] = fabsf(mk - xp[d]) > eps ? 1 : 0;
            offset[3 * i + d] = off;
        }
        const float *o = outputs + (size_t)R * i;
        for (int c = 0; c < R; ++c)
            s[3 + c] = fabsf(o[c]) <= eps ? 0 : (o[c] > 0.0f ? 1 : -1);
    }
}

/* ---- curve-approximation path (force=False) ---------------------------------------- */
/* Net.forward(x, gather=True, group=8) (model.py:65-70): x holds groups of 8 corner points;
 * inside a group every hidden neuron is kept (linear) iff it is > eps at the first OR the last
 * corner, else multiplied by 0 -- "infer within a common linear space".  out: [groups*8][R]. */
void trinet_outputs_group8(const trinet_t *n, const float *x, int64_t groups, float eps, float *out)
{
    const int R = trinet_n_outputs(n), H = n->n_hidden;
    for (int64_t g = 0; g < groups; ++g) {
        float act[8][TN_MAX_WIDTH], pre[8][TN_MAX_WIDTH];
        for (int k = 0; k < 8; ++k) {
            float xp[3];
            preprocess(n, x + 3 * (8 * g + k), xp);
            encode(n, xp, act[k], 0);
        }
        const float *p = n->mlp;
        for (int i = 0; i < n->n_linear; ++i) {
            const int nin = layer_in(n, i), nout = layer_out(n, i);
            const float *W = p, *b = p + (size_t)nout * nin;
            for (int k = 0; k < 8; ++k)
                for (int j = 0; j < nout; ++j) {
                    float acc = b[j];
                    for (int c = 0; c < nin; ++c) acc = fmaf(act[k][c], W[j * nin + c], acc);
                    pre[k][j] = acc;
                }
            if (i != n->n_linear - 1) {
                for (int k = 0; k < 8; ++k)
                    for (int j = 0; j < H; ++j) out[(size_t)R * (8 * g + k) + i * H + j] = pre[k][j];
                for (int j = 0; j < nout; ++j) {
                    const float m = (pre[0][j] > eps || pre[7][j] > eps) ? 1.0f : 0.0f;
                    for (int k = 0; k < 8; ++k) act[k][j] = pre[k][j] * m;
                }
            } else {
                for (int k = 0; k < 8; ++k) out[(size_t)R * (8 * g + k) + R - 1] = pre[k][1] - pre[k][0];
            }
            p = b + nout;
        }
    }
}

/* Which root of c[0] t^deg + ... + c[deg] (deg <= 4) does the reference take?
 * It forms the companion matrix, calls torch.linalg.eigvals (LAPACK sgeev) and keeps the LAST
 * eigenvalue, in LAPACK's output order, that is real (|imag| <= 1e-9) and lies in [0,1]
 * (geometry.py:271-299, nonzero_last at :296).  For a quadratic with both roots in [0,1] that order
 * is fixed by slanv2's standardisation of the 2x2 block: ascending whenever the root sum is positive,
 * so the reference takes the LARGER root (187 of the 189 multi-root intersections of the medium-torus
 * fixture are quadratics, every one of them ascending: tests/golden/investigate_roots.py).  For
 * cubics and quartics with several admissible roots the order depends on the rounding history of the
 * float32 QR iteration (the two quartic cases of that run went one each way): implementation defined.
 * DEFINED here, and followed by the device (csrc/curve.cuh), as: the LARGEST real root in [0,1],
 * in double precision:
 *   deg 1: -c1/c0;
 *   deg 2: discriminant D = b*b - 4ac (D < 0: none); q = -(b + sign(b) sqrt(D))/2; roots q/a and c/q;
 *   deg 3, 4: the critical points of p in (0,1) (roots of p', found the same way one degree down)
 *            cut [0,1] into intervals on which p is monotone; from the right, the first interval
 *            whose end values differ in sign (or whose right end is an exact zero) holds the root:
 *            64 bisection steps.
 * Unlike a fixed-step sign scan this finds both members of a close root pair. */
static double poly_eval(const double *c, int deg, double t)
{
    double v = c[0];
    for (int i = 1; i <= deg; ++i) v = v * t + c[i];
    return v;
}
static double bisect_root(const double *c, int deg, double lo, double hi, double flo)
{
    for (int it = 0; it < 64; ++it) {
        const double mid = 0.5 * (lo + hi), fm = poly_eval(c, deg, mid);
        if (fm == 0.0) return mid;
        if ((fm < 0.0) == (flo < 0.0)) { lo = mid; flo = fm; } else hi = mid;
    }
    return 0.5 * (lo + hi);
}
/* all real roots of a quadratic a t^2 + b t + c (a != 0) in ascending order; returns their number */
static int quadratic_roots(double a, double b, double c, double *r)
{
    const double D = b * b - 4.0 * a * c;
    if (D < 0.0) return 0;
    const double sq = sqrt(D);
    const double q = -0.5 * (b + (b < 0.0 ? -sq : sq));
    double r0 = q / a, r1 = (q != 0.0) ? c / q : r0;
    if (r0 > r1) { const double t = r0; r0 = r1; r1 = t; }
    r[0] = r0;
    r[1] = r1;
    return 2;
}
/* interior points of (0,1), ascending, where a polynomial of degree deg (<= 3, leading coefficient
 * possibly zero) changes sign or touches zero: the cut points for the next degree up */
static int roots_inside01(const double *c, int deg, double *r)
{
    while (deg > 0 && c[0] == 0.0) { ++c; --deg; }
    int n = 0;
    if (deg <= 0) return 0;
    if (deg == 1) {
        const double t = -c[1] / c[0];
        if (t > 0.0 && t < 1.0) r[n++] = t;
        return n;
    }
    if (deg == 2) {
        double q[2];
        const int k = quadratic_roots(c[0], c[1], c[2], q);
        for (int i = 0; i < k; ++i)
            if (q[i] > 0.0 && q[i] < 1.0 && (n == 0 || q[i] > r[n - 1])) r[n++] = q[i];
        return n;
    }
    /* deg == 3: monotone pieces between the critical points */
    double d[3] = {3.0 * c[0], 2.0 * c[1], c[2]}, cut[4];
    int nc = 0;
    cut[nc++] = 0.0;
    {
        double q[2];
        const int k = quadratic_roots(d[0], d[1], d[2], q);
        for (int i = 0; i < k; ++i)
            if (q[i] > 0.0 && q[i] < 1.0 && q[i] > cut[nc - 1]) cut[nc++] = q[i];
    }
    cut[nc++] = 1.0;
    for (int i = 0; i + 1 < nc; ++i) {
        const double lo = cut[i], hi = cut[i + 1], flo = poly_eval(c, 3, lo), fhi = poly_eval(c, 3, hi);
        double t = -1.0;
        if (flo == 0.0) t = lo;
        else if (fhi != 0.0 && (flo < 0.0) != (fhi < 0.0)) t = bisect_root(c, 3, lo, hi, flo);
        if (t > 0.0 && t < 1.0 && (n == 0 || t > r[n - 1])) r[n++] = t;
    }
    return n;
}
static double last_root01(const double *c, int deg)
{
    if (deg == 1) {
        const double t = -c[1] / c[0];
        return (t >= 0.0 && t <= 1.0) ? t : -1.0;
    }
    if (deg == 2) {
        double q[2];
        const int k = quadratic_roots(c[0], c[1], c[2], q);
        for (int i = k - 1; i >= 0; --i)
            if (q[i] >= 0.0 && q[i] <= 1.0) return q[i];
        return -1.0;
    }
    /* deg 3 or 4: cut [0,1] at the critical points, look at the pieces from the right */
    double d[4], cut[5];
    for (int i = 0; i < deg; ++i) d[i] = (double)(deg - i) * c[i];
    int nc = 0;
    cut[nc++] = 0.0;
    nc += roots_inside01(d, deg - 1, cut + nc);
    cut[nc++] = 1.0;
    for (int i = nc - 2; i >= 0; --i) {
        const double lo = cut[i], hi = cut[i + 1], flo = poly_eval(c, deg, lo), fhi = poly_eval(c, deg, hi);
        if (fhi == 0.0) return hi;
        if (flo != 0.0 && (flo < 0.0) != (fhi < 0.0)) return bisect_root(c, deg, lo, hi, flo);
        if (i == 0 && flo == 0.0) return lo;
    }
    return -1.0;
}

/* geometry.intersection_of_two_planes for ONE edge: p, q = the two planes' values at the 8
 * corners (index 4*iz + 2*iy + ix).  out = (x, y, z) trilinear coordinates, -1 where the
 * reference leaves "no root".  Planar (bilinear) configurations return (-1,-1,-1) exactly as
 * the reference does with its failover switched off (geometry.py:90-108). */
void curve_intersection(const float *p, const float *q, float *out)
{
    static const int T_[3][4] = {{0, 1, 4, 5}, {0, 1, 2, 3}, {0, 4, 2, 6}};
    static const int U_[3][4] = {{2, 3, 6, 7}, {4, 5, 6, 7}, {1, 5, 3, 7}};
    for (int pl = 0; pl < 3; ++pl) {
        int same = 1;
        for (int k = 0; k < 4; ++k)
            same = same && p[T_[pl][k]] == p[U_[pl][k]] && q[T_[pl][k]] == q[U_[pl][k]];
        if (same) { out[0] = out[1] = out[2] = -1.0f; return; }
    }
    static const int r[4] = {0, 1, 4, 5}, s[4] = {2, 3, 6, 7};
    /* z(v) = [v0, v1+v2, v3] in float32 like the reference, then everything in double */
    double a[3], b[3], c[3], d[3];
    a[0] = q[r[0]]; a[1] = (float)(q[r[1]] + q[r[2]]); a[2] = q[r[3]];   /* z(q[:, r]) */
    b[0] = p[s[0]]; b[1] = (float)(p[s[1]] + p[s[2]]); b[2] = p[s[3]];   /* z(p[:, s]) */
    c[0] = q[s[0]]; c[1] = (float)(q[s[1]] + q[s[2]]); c[2] = q[s[3]];   /* z(q[:, s]) */
    d[0] = p[r[0]]; d[1] = (float)(p[r[1]] + p[r[2]]); d[2] = p[r[3]];   /* z(p[:, r]) */
    double A[3][3], B[3][3], TA[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = a[i] * b[j] - c[i] * d[j];
    static const double T[3][3] = {{1, -2, 1}, {-1, 1, 0}, {1, 0, 0}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double v = 0.0;
            for (int k = 0; k < 3; ++k) v += T[k][i] * A[k][j];   /* T^T A */
            TA[i][j] = v;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double v = 0.0;
            for (int k = 0; k < 3; ++k) v += TA[i][k] * T[k][j];  /* (T^T A) T */
            B[i][j] = v;
        }
    double co[5] = {B[0][0], B[1][0] + B[0][1], B[2][0] + B[1][1] + B[0][2], B[1][2] + B[2][1], B[2][2]};
    float x = -1.0f;
    {
        float cf[5];
        for (int i = 0; i < 5; ++i) { cf[i] = (float)co[i]; if (fabsf(cf[i]) < 1e-9f) { cf[i] = 0.0f; co[i] = 0.0; } }
        int lead = 0;
        while (lead < 4 && !(fabsf(cf[lead]) > 1e-9f)) ++lead;   /* first coefficient that counts */
        if (lead < 4) {
            float mean = 0.0f;
            for (int i = lead; i < 5; ++i) mean += fabsf(cf[i]);
            mean /= (float)(5 - lead);
            if (mean > 1e-9f) {
                const double rt = last_root01(co + lead, 4 - lead);
                if (rt >= 0.0) x = (float)rt;
            }
        }
    }
    /* quad_y (geometry.py:61-67), float32, left-to-right sums */
    const float w0 = (1.0f - x) * (1.0f - x), w1 = x * (1.0f - x), w3 = x * x;
    const float AX = ((q[r[0]] * w0 + q[r[1]] * w1) + q[r[2]] * w1) + q[r[3]] * w3;
    const float BX = ((q[s[0]] * w0 + q[s[1]] * w1) + q[s[2]] * w1) + q[s[3]] * w3;
    out[0] = x;
    out[1] = AX / (AX - BX);
    out[2] = x;
}

void curve_intersections(const float *p, const float *q, int64_t count, float *out)
{
    for (int64_t i = 0; i < count; ++i) curve_intersection(p + 8 * i, q + 8 * i, out + 3 * i);
}

// curve.cuh -- device pieces of the curve-approximation path (force=False):
// Net.forward(gather=True, group=8) (model.py:65-70) restricted to the two columns the
// step needs, and geometry.intersection_of_two_planes (geometry.py:24-138) for one edge.
// Same operation order as oracle/trinet_ref.c (trinet_outputs_group8, curve_intersection).
#pragma once
#include "net_eval.cuh"

namespace tnb {

// All helpers here are WARP-COOPERATIVE: every lane of the warp calls them (full mask), each
// lane brings its own candidate edge (or none), and the warp works through the candidates that
// need the curve treatment together.  A candidate costs one network evaluation and at most 8
// rounds of polynomial samples in lock-step instead of 8 evaluations and 1024 dependent samples
// in one thread, which is what a persistent step kernel with a few candidates per CTA needs.
// Per corner / per sample the operations and their order are those of oracle/trinet_ref.c, so
// the results are bit-identical to the one-thread formulation.
constexpr unsigned kFullWarp = 0xffffffffu;

__device__ __forceinline__ double shfl_double(double v, int src)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(kFullWarp, lo, src);
    hi = __shfl_sync(kFullWarp, hi, src);
    return __hiloint2double(hi, lo);
}

// Values of output columns colA and colB at the 8 corners of the box spanned by edge (e0,e1)
// (corner index 4*iz + 2*iy + ix, coordinate from endpoint 0 or 1 per axis, geometry.py:350-372),
// evaluated "within a common linear space": a hidden neuron stays linear iff it is > eps at the
// first or the last corner, else it is multiplied by 0.
// Four candidates at a time, 8 lanes (one corner each) per candidate.
template <class C>
static __device__ void warp_group8_columns(const NetMeta &n, bool need, const float e0[3], const float e1[3], float eps, int colA,
                                    int colB, float pA[8], float pB[8])
{
    const int lane = threadIdx.x & 31, grp = lane >> 3, corner = lane & 7;
    const int L = C::L(n), H = C::H(n), NL = C::NLIN(n), R = n.R;
    unsigned todo = __ballot_sync(kFullWarp, need);
    while (todo) {
        // the grp-th pending candidate is this lane group's job (-1: none left for the group)
        int src = -1;
        {
            unsigned t = todo;
            for (int g = 0; g < 4; ++g) {
                const int s = t ? __ffs(t) - 1 : -1;
                if (g == grp) src = s;
                if (t) t &= t - 1;
            }
            todo = t;
        }
        const int from = src < 0 ? lane : src;
        float x[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float a = __shfl_sync(kFullWarp, e0[d], from), b = __shfl_sync(kFullWarp, e1[d], from);
            x[d] = ((corner >> d) & 1) ? b : a;
        }
        const int cA = __shfl_sync(kFullWarp, colA, from), cB = __shfl_sync(kFullWarp, colB, from);
        float act[C::kMaxW], pre[C::kMaxH];
        float myA = 0.0f, myB = 0.0f;
        float xp[3];
        preprocess(n, x, xp);
        for (int l = 0; l < L; ++l) {
            uint32_t cell[3];
            float frac[3];
            const float2 f = encode_level(n, l, xp, cell, frac);
            act[2 * l] = f.x;
            act[2 * l + 1] = f.y;
        }
        int base = 0;
        for (int i = 0; i < NL; ++i) {
            const int ni = C::nin(n, i), no = C::nout(n, i);
#pragma unroll(C::kUnroll)
            for (int j = 0; j < C::kMaxH; ++j) {
                if (j < no) {
                    float acc = C::w(n, base + no * ni + j);
#pragma unroll(C::kUnroll)
                    for (int c = 0; c < C::kMaxW; ++c)
                        if (c < ni) acc = __fmaf_rn(act[c], C::w(n, base + j * ni + c), acc);
                    pre[j] = acc;
                }
            }
            if (i != NL - 1) {
#pragma unroll(C::kUnroll)
                for (int j = 0; j < C::kMaxH; ++j) {
                    if (j < no) {
                        const int col = i * H + j;
                        if (col == cA) myA = pre[j];
                        if (col == cB) myB = pre[j];
                        const float first = __shfl_sync(kFullWarp, pre[j], grp * 8), last = __shfl_sync(kFullWarp, pre[j], grp * 8 + 7);
                        const float m = (first > eps || last > eps) ? 1.0f : 0.0f;
                        act[j] = pre[j] * m;
                    }
                }
            } else {
                const float v = pre[1] - pre[0];
                if (cA == R - 1) myA = v;
                if (cB == R - 1) myB = v;
            }
            base += no * ni + no;
        }
        // hand the 8 corner values to the lane that owns the candidate
        int my_grp = -1;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int s = __shfl_sync(kFullWarp, src, g * 8);
            if (s == lane) my_grp = g;
        }
        const int take = my_grp < 0 ? 0 : my_grp * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float a = __shfl_sync(kFullWarp, myA, take + k), b = __shfl_sync(kFullWarp, myB, take + k);
            if (my_grp >= 0) { pA[k] = a; pB[k] = b; }
        }
    }
}

// ---- which root (geometry.py:259-300) ---------------------------------------------------------------
// The reference keeps the LAST admissible eigenvalue of the companion matrix in LAPACK's order; for a
// quadratic with both roots in [0,1] that is the larger root, for cubics / quartics it is implementation
// defined.  Defined in oracle/trinet_ref.c (last_root01) as the LARGEST real root in [0,1], found exactly:
// closed form up to degree 2, monotone pieces between the critical points + 64 bisections above.  The
// functions below are that file's, operation for operation, in double precision (no contraction: the
// library is built with -fmad=false), so the root is bit-identical to the oracle's.  One lane per
// candidate: admissible candidates are rare (a few per thousand crossed edges), the warp does not share them.
__device__ __forceinline__ double poly_eval(const double *c, int deg, double t)
{
    double v = c[0];
    for (int i = 1; i <= deg; ++i) v = v * t + c[i];
    return v;
}
__device__ __forceinline__ double bisect_root(const double *c, int deg, double lo, double hi, double flo)
{
    for (int it = 0; it < 64; ++it) {
        const double mid = 0.5 * (lo + hi), fm = poly_eval(c, deg, mid);
        if (fm == 0.0) return mid;
        if ((fm < 0.0) == (flo < 0.0)) { lo = mid; flo = fm; } else hi = mid;
    }
    return 0.5 * (lo + hi);
}
// all real roots of a t^2 + b t + c (a != 0), ascending; returns their number
__device__ __forceinline__ int quadratic_roots(double a, double b, double c, double *r)
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
// interior points of (0,1), ascending, where a polynomial of degree <= 3 has a root: the cut points one degree up
static __device__ __noinline__ int roots_inside01(const double *c, int deg, double *r)
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
// largest real root in [0,1] of c[0] t^deg + ... + c[deg] (1 <= deg <= 4, c[0] != 0), or -1
static __device__ __noinline__ double last_root01(const double *c, int deg)
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

// p, q: the two planes' values at the 8 corners (valid where need).  out = (x, y, z) trilinear
// coordinates.  geometry.intersection_of_two_planes (geometry.py:24-138) for one edge per lane.
static __device__ void warp_curve_intersection(bool need, const float *p, const float *q, float out[3])
{
    const int r[4] = {0, 1, 4, 5}, s[4] = {2, 3, 6, 7};
    double co[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    int lead = 0;
    bool bilinear = false, want_root = false;
    if (need) {
        const int T_[3][4] = {{0, 1, 4, 5}, {0, 1, 2, 3}, {0, 4, 2, 6}};
        const int U_[3][4] = {{2, 3, 6, 7}, {4, 5, 6, 7}, {1, 5, 3, 7}};
        for (int pl = 0; pl < 3; ++pl) {
            bool same = true;
            for (int k = 0; k < 4; ++k) same = same && p[T_[pl][k]] == p[U_[pl][k]] && q[T_[pl][k]] == q[U_[pl][k]];
            bilinear = bilinear || same;  // bilinear cases: geometry.py:108
        }
        if (!bilinear) {
            double a[3], b[3], c[3], d[3];
            a[0] = q[r[0]]; a[1] = (double)(q[r[1]] + q[r[2]]); a[2] = q[r[3]];
            b[0] = p[s[0]]; b[1] = (double)(p[s[1]] + p[s[2]]); b[2] = p[s[3]];
            c[0] = q[s[0]]; c[1] = (double)(q[s[1]] + q[s[2]]); c[2] = q[s[3]];
            d[0] = p[r[0]]; d[1] = (double)(p[r[1]] + p[r[2]]); d[2] = p[r[3]];
            double A[3][3], B[3][3], TA[3][3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) A[i][j] = a[i] * b[j] - c[i] * d[j];
            const double T[3][3] = {{1, -2, 1}, {-1, 1, 0}, {1, 0, 0}};
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    double v = 0.0;
                    for (int k = 0; k < 3; ++k) v += T[k][i] * A[k][j];
                    TA[i][j] = v;
                }
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    double v = 0.0;
                    for (int k = 0; k < 3; ++k) v += TA[i][k] * T[k][j];
                    B[i][j] = v;
                }
            co[0] = B[0][0]; co[1] = B[1][0] + B[0][1]; co[2] = B[2][0] + B[1][1] + B[0][2]; co[3] = B[1][2] + B[2][1]; co[4] = B[2][2];
            float cf[5];
            for (int i = 0; i < 5; ++i) {
                cf[i] = (float)co[i];
                if (fabsf(cf[i]) < 1e-9f) { cf[i] = 0.0f; co[i] = 0.0; }
            }
            while (lead < 4 && !(fabsf(cf[lead]) > 1e-9f)) ++lead;
            if (lead < 4) {
                float mean = 0.0f;
                for (int i = lead; i < 5; ++i) mean += fabsf(cf[i]);
                mean = __fdiv_rn(mean, (float)(5 - lead));
                want_root = mean > 1e-9f;
            }
        }
    }
    const double rt = want_root ? last_root01(co + lead, 4 - lead) : -1.0;
    if (!need) return;
    if (bilinear) { out[0] = out[1] = out[2] = -1.0f; return; }
    float x = -1.0f;
    if (want_root && rt >= 0.0) x = (float)rt;
    const float w0 = (1.0f - x) * (1.0f - x), w1 = x * (1.0f - x), w3 = x * x;
    const float AX = ((q[r[0]] * w0 + q[r[1]] * w1) + q[r[2]] * w1) + q[r[3]] * w3;
    const float BX = ((q[s[0]] * w0 + q[s[1]] * w1) + q[s[2]] * w1) + q[s[3]] * w3;
    out[0] = x;
    out[1] = __fdiv_rn(AX, AX - BX);
    out[2] = x;
}

}  // namespace tnb

// curve.cuh -- device pieces of the curve-approximation path (force=False):
// Net.forward(gather=True, group=8) (model.py:65-70) restricted to the two columns the
// step needs, and geometry.intersection_of_two_planes (geometry.py:24-138) for one edge.
// Same operation order as oracle/trinet_ref.c (trinet_outputs_group8, curve_intersection).
#pragma once
#include "net_eval.cuh"

namespace tnb {

// Values of output columns colA and colB at the 8 corners of the box spanned by edge (e0,e1)
// (corner index 4*iz + 2*iy + ix, coordinate from endpoint 0 or 1 per axis, geometry.py:350-372),
// evaluated "within a common linear space": a hidden neuron stays linear iff it is > eps at the
// first or the last corner, else it is multiplied by 0.
template <class C>
__device__ void group8_columns(const NetMeta &n, const float e0[3], const float e1[3], float eps, int colA,
                               int colB, float pA[8], float pB[8])
{
    float act[8][C::kMaxW], pre[8][C::kMaxH];
    const int L = C::L(n), H = C::H(n), NL = C::NLIN(n), R = n.R;
    for (int k = 0; k < 8; ++k) {
        const float x[3] = {(k & 1) ? e1[0] : e0[0], (k & 2) ? e1[1] : e0[1], (k & 4) ? e1[2] : e0[2]};
        float xp[3];
        preprocess(n, x, xp);
        for (int l = 0; l < L; ++l) {
            uint32_t cell[3];
            float frac[3];
            const float2 f = encode_level(n, l, xp, cell, frac);
            act[k][2 * l] = f.x;
            act[k][2 * l + 1] = f.y;
        }
    }
    int base = 0;
    for (int i = 0; i < NL; ++i) {
        const int ni = C::nin(n, i), no = C::nout(n, i);
        for (int k = 0; k < 8; ++k)
            for (int j = 0; j < no; ++j) {
                float acc = C::w(n, base + no * ni + j);
                for (int c = 0; c < ni; ++c) acc = __fmaf_rn(act[k][c], C::w(n, base + j * ni + c), acc);
                pre[k][j] = acc;
            }
        if (i != NL - 1) {
            for (int j = 0; j < no; ++j) {
                const int col = i * H + j;
                if (col == colA)
                    for (int k = 0; k < 8; ++k) pA[k] = pre[k][j];
                if (col == colB)
                    for (int k = 0; k < 8; ++k) pB[k] = pre[k][j];
                const float m = (pre[0][j] > eps || pre[7][j] > eps) ? 1.0f : 0.0f;
                for (int k = 0; k < 8; ++k) act[k][j] = pre[k][j] * m;
            }
        } else {
            for (int k = 0; k < 8; ++k) {
                const float v = pre[k][1] - pre[k][0];
                if (colA == R - 1) pA[k] = v;
                if (colB == R - 1) pB[k] = v;
            }
        }
        base += no * ni + no;
    }
}

__device__ __forceinline__ double poly_eval(const double *c, int deg, double t)
{
    double v = c[0];
    for (int i = 1; i <= deg; ++i) v = v * t + c[i];
    return v;
}

// smallest real root in [0,1] (the one the reference's eigenvalue filter keeps), or -1
__device__ double smallest_root01(const double *c, int deg)
{
    const int N = 1024;
    double t0 = 0.0, f0 = poly_eval(c, deg, 0.0);
    if (f0 == 0.0) return 0.0;
    for (int k = 1; k <= N; ++k) {
        const double t1 = (double)k / (double)N, f1 = poly_eval(c, deg, t1);
        if (f1 == 0.0) return t1;
        if ((f0 < 0.0) != (f1 < 0.0)) {
            double lo = t0, hi = t1, flo = f0;
            for (int it = 0; it < 60; ++it) {
                const double mid = 0.5 * (lo + hi), fm = poly_eval(c, deg, mid);
                if (fm == 0.0) return mid;
                if ((fm < 0.0) == (flo < 0.0)) { lo = mid; flo = fm; } else hi = mid;
            }
            return 0.5 * (lo + hi);
        }
        t0 = t1;
        f0 = f1;
    }
    return -1.0;
}

// p, q: the two planes' values at the 8 corners.  out = (x, y, z) trilinear coordinates.
__device__ void curve_intersection(const float *p, const float *q, float out[3])
{
    const int T_[3][4] = {{0, 1, 4, 5}, {0, 1, 2, 3}, {0, 4, 2, 6}};
    const int U_[3][4] = {{2, 3, 6, 7}, {4, 5, 6, 7}, {1, 5, 3, 7}};
    for (int pl = 0; pl < 3; ++pl) {
        bool same = true;
        for (int k = 0; k < 4; ++k) same = same && p[T_[pl][k]] == p[U_[pl][k]] && q[T_[pl][k]] == q[U_[pl][k]];
        if (same) { out[0] = out[1] = out[2] = -1.0f; return; }  // bilinear cases: geometry.py:108
    }
    const int r[4] = {0, 1, 4, 5}, s[4] = {2, 3, 6, 7};
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
    double co[5] = {B[0][0], B[1][0] + B[0][1], B[2][0] + B[1][1] + B[0][2], B[1][2] + B[2][1], B[2][2]};
    float x = -1.0f;
    {
        float cf[5];
        for (int i = 0; i < 5; ++i) {
            cf[i] = (float)co[i];
            if (fabsf(cf[i]) < 1e-9f) { cf[i] = 0.0f; co[i] = 0.0; }
        }
        int lead = 0;
        while (lead < 4 && !(fabsf(cf[lead]) > 1e-9f)) ++lead;
        if (lead < 4) {
            float mean = 0.0f;
            for (int i = lead; i < 5; ++i) mean += fabsf(cf[i]);
            mean = __fdiv_rn(mean, (float)(5 - lead));
            if (mean > 1e-9f) {
                const double rt = smallest_root01(co + lead, 4 - lead);
                if (rt >= 0.0) x = (float)rt;
            }
        }
    }
    const float w0 = (1.0f - x) * (1.0f - x), w1 = x * (1.0f - x), w3 = x * x;
    const float AX = ((q[r[0]] * w0 + q[r[1]] * w1) + q[r[2]] * w1) + q[r[3]] * w3;
    const float BX = ((q[s[0]] * w0 + q[s[1]] * w1) + q[s[2]] * w1) + q[s[3]] * w3;
    out[0] = x;
    out[1] = __fdiv_rn(AX, AX - BX);
    out[2] = x;
}

}  // namespace tnb

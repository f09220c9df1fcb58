// compat.cu -- stage-level entry points of the reference's `tropical.geometry` / `Net.forward(group=8)`
// for callers that drive the stages of the curve-approximation path themselves.  The extraction path
// (tnb_subpoly / tnb_subpoly_steps) runs the same device functions fused inside its step kernels
// (curve.cuh, faces.cu); here each one is a kernel of its own over the caller's arrays:
//   tnb_net_outputs_group8   Net.forward(x, gather=True, group=8)            model.py:52-76 (:65-70)
//   tnb_curve_intersections  geometry.intersection_of_two_planes             geometry.py:24-138
//   tnb_polygon_order        geometry.sort_polygon_vertices_batch's ordering geometry.py:483-516
//   tnb_curve_gradient_descent  subpoly_debug.deal_with_gradient_descent     subpoly_debug.py:121-165
#include "curve.cuh"
#include "repair.cuh"
#include "net_eval.cuh"
#include "runtime.cuh"
#include "sort.cuh"

namespace tnb {

constexpr int kCompatThreads = 128;

// Eight lanes per group of eight points (the corners of an edge's box): every lane evaluates its own
// point; after each hidden layer a neuron stays linear for the whole group iff it is > eps at the first
// or the last point of the group, else it is multiplied by 0 ("infer within a common linear space").
// out: [groups * 8][R] pre-activation rows, the last column o1 - o0; raw: [groups * 8][2] the last layer.
template <class C>
__global__ void __launch_bounds__(kCompatThreads) k_outputs_group8(const __grid_constant__ NetMeta n, const float *__restrict__ x,
                                                                   int64_t groups, float eps, float *__restrict__ out,
                                                                   float *__restrict__ raw)
{
    const int lane = threadIdx.x & 31, first = lane & ~7, last = first + 7;
    const int L = C::L(n), H = C::H(n), NL = C::NLIN(n), R = n.R;
    const int64_t total = groups * 8, rounds = (total + (int64_t)gridDim.x * blockDim.x - 1) / ((int64_t)gridDim.x * blockDim.x);
    for (int64_t r = 0; r < rounds; ++r) {  // warp-uniform trip count: the shuffles below need every lane
        const int64_t t = (r * gridDim.x + blockIdx.x) * (int64_t)blockDim.x + threadIdx.x;
        const bool active = t < total;
        const int64_t tt = active ? t : 0;
        float xw[3] = {x[3 * tt], x[3 * tt + 1], x[3 * tt + 2]}, xp[3];
        preprocess(n, xw, xp);
        float act[C::kMaxW], pre[C::kMaxH];
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
                        if (active) out[tt * R + i * H + j] = pre[j];
                        const float a = __shfl_sync(0xffffffffu, pre[j], first), b = __shfl_sync(0xffffffffu, pre[j], last);
                        act[j] = pre[j] * ((a > eps || b > eps) ? 1.0f : 0.0f);
                    }
                }
            } else if (active) {
                out[tt * R + R - 1] = pre[1] - pre[0];
                if (raw) { raw[2 * tt] = pre[0]; raw[2 * tt + 1] = pre[1]; }
            }
            base += no * ni + no;
        }
    }
}

// Net.forward(x, gather=True) with the last layer's two outputs as well (model.py:52-76 returns both)
template <class C>
__global__ void __launch_bounds__(kCompatThreads) k_forward_raw(const __grid_constant__ NetMeta n, const float *__restrict__ x, int64_t count,
                                                                float *__restrict__ out, float *__restrict__ raw)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const float p[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
        float xp[3];
        preprocess(n, p, xp);
        float pre[(C::kMaxLin - 1) * C::kMaxH];
        float o[2];
        forward<C>(n, xp, pre, o);
        const int H = C::H(n), NL = C::NLIN(n);
        if (out) {
            float *row = out + i * n.R;
#pragma unroll(C::kUnroll)
            for (int l = 0; l < C::kMaxLin - 1; ++l)
                if (l < NL - 1) {
#pragma unroll(C::kUnroll)
                    for (int j = 0; j < C::kMaxH; ++j)
                        if (j < H) row[l * H + j] = pre[l * C::kMaxH + j];
                }
            row[(NL - 1) * H] = o[1] - o[0];
        }
        raw[2 * i] = o[0];
        raw[2 * i + 1] = o[1];
    }
}

__global__ void __launch_bounds__(kCompatThreads) k_curve_intersections(const float *__restrict__ p, const float *__restrict__ q, int64_t count,
                                                                        float *__restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        float pp[8], qq[8], o[3];
#pragma unroll
        for (int k = 0; k < 8; ++k) { pp[k] = p[8 * i + k]; qq[k] = q[8 * i + k]; }
        warp_curve_intersection(true, pp, qq, o);
        out[3 * i] = o[0];
        out[3 * i + 1] = o[1];
        out[3 * i + 2] = o[2];
    }
}

// One thread per padded face row v[b][0..M): entries whose position has norm 0 are padding.  The score of
// every entry (padding included, as in the reference) around the row's centre against the normal n[b], and
// the stable descending order of the scores.  Operation order: oracle/subpoly_ref.py:polygon_order.
__global__ void __launch_bounds__(kCompatThreads) k_polygon_order(const float *__restrict__ v, const float *__restrict__ nrm, int64_t B, int M,
                                                                  int base, unsigned long long *__restrict__ keys, int64_t *__restrict__ order,
                                                                  unsigned char *__restrict__ valid_sorted)
{
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        const float *row = v + b * M * 3;
        float sx = 0.0f, sy = 0.0f, sz = 0.0f;
        int k = 0;
        for (int j = 0; j < M; ++j) {
            const float x = row[3 * j], y = row[3 * j + 1], z = row[3 * j + 2];
            sx = sx + x;
            sy = sy + y;
            sz = sz + z;
            if (__fsqrt_rn((x * x + y * y) + z * z) > 0.0f) ++k;
        }
        const float kf = (float)(k == 0 ? 1 : k);
        const float mean[3] = {__fdiv_rn(sx, kf), __fdiv_rn(sy, kf), __fdiv_rn(sz, kf)};
        const float a[3] = {row[3 * base] - mean[0], row[3 * base + 1] - mean[1], row[3 * base + 2] - mean[2]};
        const float an = fmaxf(__fsqrt_rn((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]), 1e-8f);
        const float ua[3] = {__fdiv_rn(a[0], an), __fdiv_rn(a[1], an), __fdiv_rn(a[2], an)};
        const float n3[3] = {nrm[3 * b], nrm[3 * b + 1], nrm[3 * b + 2]};
        unsigned long long *ks = keys + b * M;
        for (int j = 0; j < M; ++j) {
            const float u[3] = {row[3 * j] - mean[0], row[3 * j + 1] - mean[1], row[3 * j + 2] - mean[2]};
            const float d0 = a[1] * u[2] - a[2] * u[1], d1 = a[2] * u[0] - a[0] * u[2], d2 = a[0] * u[1] - a[1] * u[0];
            const float un = fmaxf(__fsqrt_rn((u[0] * u[0] + u[1] * u[1]) + u[2] * u[2]), 1e-8f);
            const float c = (ua[0] * __fdiv_rn(u[0], un) + ua[1] * __fdiv_rn(u[1], un)) + ua[2] * __fdiv_rn(u[2], un);
            const float dn = (d0 * n3[0] + d1 * n3[1]) + d2 * n3[2];
            const float s = c * (dn >= 0.0f ? 1.0f : -1.0f) + (dn < 0.0f ? 2.0f : 0.0f);
            const uint32_t bits = __float_as_uint(s);
            const uint32_t asc = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);  // orders like the float
            ks[j] = ((unsigned long long)(0xFFFFFFFFu - asc) << 32) | (unsigned)j;  // descending score, then position: stable
        }
        thread_sort(ks, M);
        for (int j = 0; j < M; ++j) {
            const int src = (int)(uint32_t)ks[j];
            order[b * M + j] = src;
            const float x = row[3 * src], y = row[3 * src + 1], z = row[3 * src + 2];
            valid_sorted[b * M + j] = __fsqrt_rn((x * x + y * y) + z * z) > 0.0f ? 1 : 0;
        }
    }
}

}  // namespace tnb

using namespace tnb;

extern "C" {

int tnb_net_outputs_group8(const tnb_net *net, const float *d_x, int64_t groups, float eps, float *d_out, float *d_raw, void *stream)
{
    if (!net || groups < 0 || (groups > 0 && (!d_x || !d_out))) { set_error("tnb_net_outputs_group8: bad argument"); return TNB_ERR_INVALID; }
    if (groups == 0) return TNB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(groups * 8, kCompatThreads);
    if (net->fixed_cfg) k_outputs_group8<CfgRef><<<g, kCompatThreads, 0, s>>>(net->meta, d_x, groups, eps, d_out, d_raw);
    else k_outputs_group8<CfgAny><<<g, kCompatThreads, 0, s>>>(net->meta, d_x, groups, eps, d_out, d_raw);
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}

int tnb_net_forward(const tnb_net *net, const float *d_x, int64_t n, float *d_out, float *d_raw, void *stream)
{
    if (!net || n < 0 || (n > 0 && (!d_x || !d_raw))) { set_error("tnb_net_forward: bad argument"); return TNB_ERR_INVALID; }
    if (n == 0) return TNB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(n, kCompatThreads);
    if (net->fixed_cfg) k_forward_raw<CfgRef><<<g, kCompatThreads, 0, s>>>(net->meta, d_x, n, d_out, d_raw);
    else k_forward_raw<CfgAny><<<g, kCompatThreads, 0, s>>>(net->meta, d_x, n, d_out, d_raw);
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}

int tnb_curve_intersections(const float *d_p, const float *d_q, int64_t count, float *d_out, void *stream)
{
    if (count < 0 || (count > 0 && (!d_p || !d_q || !d_out))) { set_error("tnb_curve_intersections: bad argument"); return TNB_ERR_INVALID; }
    if (count == 0) return TNB_OK;
    k_curve_intersections<<<grid_for(count, kCompatThreads), kCompatThreads, 0, (cudaStream_t)stream>>>(d_p, d_q, count, d_out);
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}

int tnb_curve_gradient_descent(const tnb_net *net, const float *d_edges, float *d_ints, const int32_t *d_plane, int32_t idx, float eps,
                               int64_t count, float *d_dnew, int32_t *bodies, int32_t *within_eps, void *stream)
{
    if (!net || count < 0 || idx < 1 || idx >= net->meta.R || (count > 0 && (!d_edges || !d_ints || !d_plane || !d_dnew))) {
        set_error("tnb_curve_gradient_descent: bad argument");
        return TNB_ERR_INVALID;
    }
    if (bodies) *bodies = 0;
    if (within_eps) *within_eps = 1;
    if (count == 0) return TNB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    current_stream() = s;
    DevBuf<int> gd;
    TNB_CUDA(gd.reserve(kGdHead));
    k_gd_reset<<<1, 32, 0, s>>>(gd.p);
    const unsigned g = grid_for(count, 64);
    if (net->fixed_cfg) {
        k_gd_stage_note<CfgRef><<<g, 64, 0, s>>>(net->meta, d_edges, d_ints, d_plane, idx, eps, count, gd.p);
        k_gd_stage_walk<CfgRef><<<g, 64, 0, s>>>(net->meta, d_edges, d_ints, d_plane, idx, count, d_dnew, gd.p);
    } else {
        k_gd_stage_note<CfgAny><<<g, 64, 0, s>>>(net->meta, d_edges, d_ints, d_plane, idx, eps, count, gd.p);
        k_gd_stage_walk<CfgAny><<<g, 64, 0, s>>>(net->meta, d_edges, d_ints, d_plane, idx, count, d_dnew, gd.p);
    }
    TNB_LAUNCH_CHECK();
    int h[4];
    TNB_CUDA(cudaMemcpyAsync(h, gd.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    TNB_CUDA(cudaStreamSynchronize(s));
    if (bodies) *bodies = h[GD_BODIES];
    if (within_eps) *within_eps = h[GD_OK];
    return TNB_OK;
}

int tnb_polygon_order(const float *d_v, const float *d_normals, int64_t B, int32_t M, int32_t base, int64_t *d_order,
                      uint8_t *d_valid_sorted, void *stream)
{
    if (B < 0 || M < 1 || base < 0 || base >= M || (B > 0 && (!d_v || !d_normals || !d_order || !d_valid_sorted))) {
        set_error("tnb_polygon_order: bad argument");
        return TNB_ERR_INVALID;
    }
    if (B == 0) return TNB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    current_stream() = s;
    DevBuf<unsigned long long> keys;
    TNB_CUDA(keys.reserve((size_t)B * M));
    k_polygon_order<<<grid_for(B, kCompatThreads), kCompatThreads, 0, s>>>(d_v, d_normals, B, M, base, keys.p, d_order, d_valid_sorted);
    TNB_LAUNCH_CHECK();
    return TNB_OK;  // the key scratch is released in stream order
}

}  // extern "C"

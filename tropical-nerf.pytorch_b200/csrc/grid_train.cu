// grid_train.cu -- the hash-grid encoding under autograd: forward, backward and the backward of
// the backward, for the training loop of stanford/train.py:180-201 (L1 + eikonal loss: the
// eikonal term differentiates d sdf / d x once more, so the input gradient itself must be
// differentiable).  Replaces what tiny-cuda-nn's GridEncoding provides to the reference through
// tcnn.Encoding (tropical.py:32-47): kernel_grid (forward), kernel_grid_backward +
// kernel_grid_backward_input (backward), kernel_grid_backward_input_backward_grid /
// _backward_input (double backward), restated from the published algorithm for F = 2, D = 3,
// linear interpolation.
//
// The table is the CALLER's parameter storage (the torch Parameter `enc.module.params`): no copy
// of the weights is made, gradients are accumulated into the caller's gradient buffer with
// float2 atomics (sm_90+ has a native 8-byte float2 atomicAdd: one red per corner).
//
// One thread per (point, level); the 8 corner rows of a cell are fetched once and reused by every
// term.  Training is launch bound at the reference's batch size (1000 points), so the kernels aim
// at few launches (3 per iteration instead of ~2000 torch ops), not at bandwidth.
#include "common.cuh"
#include "runtime.cuh"

namespace tnb {

struct GridLayout {
    int L;
    LevelMeta lvl[kMaxLevels];
};

static int make_layout(const tnb_grid_desc *d, GridLayout &g, int64_t *total_out)
{
    if (!d || d->n_levels < 1 || d->n_levels > kMaxLevels || d->log2_hashmap < 1 || d->log2_hashmap > 30 || d->base_resolution < 1) {
        set_error("tnb_grid_train: bad grid description");
        return TNB_ERR_INVALID;
    }
    memset(&g, 0, sizeof(g));
    g.L = d->n_levels;
    // same derivation as tnb_net_create (tiny-cuda-nn's GridEncoding constructor)
    const float log2_pls = std::log2((float)d->per_level_scale);
    uint64_t total = 0;
    for (int l = 0; l < g.L; ++l) {
        const float scale = std::exp2((float)l * log2_pls) * (float)d->base_resolution - 1.0f;
        const uint32_t res = (uint32_t)std::ceil(scale) + 1u;
        const uint32_t max_params = 0xFFFFFFFFu / 2;
        uint32_t n = std::pow((float)res, 3.0f) > (float)max_params ? max_params : res * res * res;
        n = (n + 7u) / 8u * 8u;
        const uint32_t cap = 1u << d->log2_hashmap;
        if (n > cap) n = cap;
        const uint64_t r1 = res, r2 = r1 * r1, r3 = r2 * r1;
        uint32_t mode = kLevelGeneric;
        if (r1 <= n && r2 <= n && r3 <= n && r3 < (1ull << 32)) mode = kLevelDense;
        else if ((n & (n - 1u)) == 0u) mode = kLevelHashPow2;
        g.lvl[l] = LevelMeta{scale, res, n, (uint32_t)total, mode, (uint32_t)(r2 & 0xFFFFFFFFull)};
        total += n;
    }
    if (total_out) *total_out = (int64_t)total;
    return TNB_OK;
}

constexpr int kTrainThreads = 128;

// cell, fractional position and the 8 corner rows of (point i, level l)
struct Cell {
    float frac[3];
    uint32_t idx[8];
    float scale;
};
__device__ __forceinline__ void locate(const LevelMeta &lv, const float *__restrict__ x, Cell &c)
{
    uint32_t cell[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float pos = __fmaf_rn(lv.scale, x[d], 0.5f);
        const float fl = floorf(pos);
        cell[d] = (uint32_t)(int)fl;
        c.frac[d] = pos - fl;
    }
    corner_indices(lv, cell[0], cell[1], cell[2], c.idx);
    c.scale = lv.scale;
}
__device__ __forceinline__ float wgt(const Cell &c, int d, int bit) { return bit ? c.frac[d] : 1.0f - c.frac[d]; }

// enc[i][2l..2l+1] = sum_c w_c * table[idx_c]
__global__ void __launch_bounds__(kTrainThreads) k_grid_train_fwd(const __grid_constant__ GridLayout g, const float2 *__restrict__ table,
                                                                  const float *__restrict__ x, int64_t n, float *__restrict__ enc)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * g.L) return;
    const int64_t i = t / g.L;
    const int l = (int)(t - i * g.L);
    const LevelMeta lv = g.lvl[l];
    Cell c;
    locate(lv, x + 3 * i, c);
    const float2 *tab = table + lv.off;
    float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float w = wgt(c, 0, k & 1) * wgt(c, 1, (k >> 1) & 1) * wgt(c, 2, (k >> 2) & 1);
        const float2 v = __ldg(tab + c.idx[k]);
        acc.x = __fmaf_rn(w, v.x, acc.x);
        acc.y = __fmaf_rn(w, v.y, acc.y);
    }
    reinterpret_cast<float2 *>(enc)[i * g.L + l] = acc;
}

// dtable[idx_c] += w_c * denc ;  dx[d] += scale * sum_pairs w_other * <denc, v_hi - v_lo>
// dx is accumulated over the levels of a point with atomics on [n,3] floats (L threads per point).
__global__ void __launch_bounds__(kTrainThreads) k_grid_train_bwd(const __grid_constant__ GridLayout g, const float2 *__restrict__ table,
                                                                  const float *__restrict__ x, int64_t n, const float *__restrict__ denc,
                                                                  float2 *__restrict__ dtable, float *__restrict__ dx)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * g.L) return;
    const int64_t i = t / g.L;
    const int l = (int)(t - i * g.L);
    const LevelMeta lv = g.lvl[l];
    Cell c;
    locate(lv, x + 3 * i, c);
    const float2 dy = reinterpret_cast<const float2 *>(denc)[i * g.L + l];
    if (dtable) {
        float2 *dt = dtable + lv.off;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float w = wgt(c, 0, k & 1) * wgt(c, 1, (k >> 1) & 1) * wgt(c, 2, (k >> 2) & 1);
            atomicAdd(dt + c.idx[k], make_float2(w * dy.x, w * dy.y));
        }
    }
    if (dx) {
        const float2 *tab = table + lv.off;
        float2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(tab + c.idx[k]);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
            float acc = 0.0f;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int b1 = p & 1, b2 = p >> 1;
                const int lo = (b1 << d1) | (b2 << d2), hi = lo | (1 << d);
                const float w = wgt(c, d1, b1) * wgt(c, d2, b2);
                acc += w * (dy.x * (v[hi].x - v[lo].x) + dy.y * (v[hi].y - v[lo].y));
            }
            atomicAdd(dx + 3 * i + d, c.scale * acc);
        }
    }
}

// Backward of  dx = J(x, table) . denc  given ddx = dLoss / d(dx):
//   ddenc[f]      = sum_d ddx[d] * scale * sum_pairs w_other * (v_hi - v_lo)[f]
//   dtable[idx_k] += sum_d ddx[d] * scale * sign_d(k) * w_other(k) * denc           (sign: +1 upper, -1 lower corner along d)
//   dx2[e]        = sum_{d != e} ddx[d] * scale^2 * sum_{bit of the third axis} w_third * <denc, v11 - v10 - v01 + v00>
//                   (the interpolation is linear in each axis: the pure second derivatives vanish)
__global__ void __launch_bounds__(kTrainThreads) k_grid_train_bwd_bwd(const __grid_constant__ GridLayout g, const float2 *__restrict__ table,
                                                                      const float *__restrict__ x, int64_t n, const float *__restrict__ denc,
                                                                      const float *__restrict__ ddx, float2 *__restrict__ dtable,
                                                                      float *__restrict__ ddenc, float *__restrict__ dx2)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * g.L) return;
    const int64_t i = t / g.L;
    const int l = (int)(t - i * g.L);
    const LevelMeta lv = g.lvl[l];
    Cell c;
    locate(lv, x + 3 * i, c);
    const float2 dy = reinterpret_cast<const float2 *>(denc)[i * g.L + l];
    const float gx[3] = {ddx[3 * i] * c.scale, ddx[3 * i + 1] * c.scale, ddx[3 * i + 2] * c.scale};
    const float2 *tab = table + lv.off;
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(tab + c.idx[k]);
    if (ddenc) {
        float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int b1 = p & 1, b2 = p >> 1;
                const int lo = (b1 << d1) | (b2 << d2), hi = lo | (1 << d);
                const float w = gx[d] * wgt(c, d1, b1) * wgt(c, d2, b2);
                acc.x += w * (v[hi].x - v[lo].x);
                acc.y += w * (v[hi].y - v[lo].y);
            }
        }
        reinterpret_cast<float2 *>(ddenc)[i * g.L + l] = acc;
    }
    if (dtable) {
        float2 *dt = dtable + lv.off;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float coef = 0.0f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
                const float w = wgt(c, d1, (k >> d1) & 1) * wgt(c, d2, (k >> d2) & 1);
                coef += ((k >> d) & 1) ? gx[d] * w : -(gx[d] * w);
            }
            atomicAdd(dt + c.idx[k], make_float2(coef * dy.x, coef * dy.y));
        }
    }
    if (dx2) {
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            float acc = 0.0f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (d == e) continue;
                const int o = 3 - d - e;  // the third axis
                float m = 0.0f;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int k00 = b << o, k10 = k00 | (1 << d), k01 = k00 | (1 << e), k11 = k10 | (1 << e);
                    const float sx = v[k11].x - v[k10].x - v[k01].x + v[k00].x;
                    const float sy = v[k11].y - v[k10].y - v[k01].y + v[k00].y;
                    m += wgt(c, o, b) * (dy.x * sx + dy.y * sy);
                }
                acc += gx[d] * m;
            }
            atomicAdd(dx2 + 3 * i + e, c.scale * acc);
        }
    }
}

static unsigned train_grid(int64_t items) { return (unsigned)((items + kTrainThreads - 1) / kTrainThreads); }

}  // namespace tnb

using namespace tnb;

extern "C" {

int64_t tnb_grid_train_table_len(const tnb_grid_desc *desc)
{
    GridLayout g;
    int64_t total = 0;
    if (make_layout(desc, g, &total)) return -1;
    return total * 2;
}

int tnb_grid_train_forward(const tnb_grid_desc *desc, const float *d_table, const float *d_x, int64_t n, float *d_enc, void *stream)
{
    GridLayout g;
    int rc;
    if ((rc = make_layout(desc, g, nullptr))) return rc;
    if (n < 0 || (n > 0 && (!d_table || !d_x || !d_enc))) { set_error("tnb_grid_train_forward: null argument"); return TNB_ERR_INVALID; }
    if (n == 0) return TNB_OK;
    if (tnb_device_count() == 0) { set_error("no CUDA device: this library has no CPU path"); return TNB_ERR_CUDA; }
    cudaStream_t s = (cudaStream_t)stream;
    k_grid_train_fwd<<<train_grid(n * g.L), kTrainThreads, 0, s>>>(g, reinterpret_cast<const float2 *>(d_table), d_x, n, d_enc);
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}

int tnb_grid_train_backward(const tnb_grid_desc *desc, const float *d_table, const float *d_x, int64_t n, const float *d_denc,
                            float *d_dtable, float *d_dx, void *stream)
{
    GridLayout g;
    int rc;
    if ((rc = make_layout(desc, g, nullptr))) return rc;
    if (n < 0 || (n > 0 && (!d_table || !d_x || !d_denc))) { set_error("tnb_grid_train_backward: null argument"); return TNB_ERR_INVALID; }
    if (n == 0 || (!d_dtable && !d_dx)) return TNB_OK;
    if (tnb_device_count() == 0) { set_error("no CUDA device: this library has no CPU path"); return TNB_ERR_CUDA; }
    cudaStream_t s = (cudaStream_t)stream;
    if (d_dx) TNB_CUDA(cudaMemsetAsync(d_dx, 0, (size_t)n * 3 * sizeof(float), s));
    k_grid_train_bwd<<<train_grid(n * g.L), kTrainThreads, 0, s>>>(g, reinterpret_cast<const float2 *>(d_table), d_x, n, d_denc,
                                                                   reinterpret_cast<float2 *>(d_dtable), d_dx);
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}

int tnb_grid_train_backward_backward(const tnb_grid_desc *desc, const float *d_table, const float *d_x, int64_t n, const float *d_denc,
                                     const float *d_ddx, float *d_dtable, float *d_ddenc, float *d_dx2, void *stream)
{
    GridLayout g;
    int rc;
    if ((rc = make_layout(desc, g, nullptr))) return rc;
    if (n < 0 || (n > 0 && (!d_table || !d_x || !d_denc || !d_ddx))) { set_error("tnb_grid_train_backward_backward: null argument"); return TNB_ERR_INVALID; }
    if (n == 0 || (!d_dtable && !d_ddenc && !d_dx2)) return TNB_OK;
    if (tnb_device_count() == 0) { set_error("no CUDA device: this library has no CPU path"); return TNB_ERR_CUDA; }
    cudaStream_t s = (cudaStream_t)stream;
    if (d_dx2) TNB_CUDA(cudaMemsetAsync(d_dx2, 0, (size_t)n * 3 * sizeof(float), s));
    k_grid_train_bwd_bwd<<<train_grid(n * g.L), kTrainThreads, 0, s>>>(g, reinterpret_cast<const float2 *>(d_table), d_x, n, d_denc, d_ddx,
                                                                       reinterpret_cast<float2 *>(d_dtable), d_ddenc, d_dx2);
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}

}  // extern "C"

// net_kernels.cu -- network-level entry points of the C ABI: hash-grid encoding, fused
// trilinear evaluation, SDF + input gradient, region indicators, dense sign sweep.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "net_eval.cuh"
#include "runtime.cuh"

namespace tnb {

// ---- error state ---------------------------------------------------------------------
static thread_local std::string g_error;
static thread_local int64_t g_launches = 0;
void set_error(const std::string &msg) { g_error = msg; }
int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    g_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" +
              std::to_string(line) + ")";
    return TNB_ERR_CUDA;
}
void count_launch(int n) { g_launches += n; }

cudaStream_t &current_stream()
{
    static thread_local cudaStream_t s = nullptr;
    return s;
}
void init_pool_once()
{
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = ~0ull;  // never trim: the work buffers are reused by the next extraction
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    });
}

// ---- block cache (runtime.cuh) --------------------------------------------------------------------
static const bool g_block_cache = std::getenv("TNB_NO_BLOCK_CACHE") == nullptr;
void block_cache_trim();
static size_t size_class(size_t bytes)
{
    if (bytes <= 512) return 512;
    size_t top = (size_t)1 << (63 - __builtin_clzll((unsigned long long)bytes));  // largest power of two <= bytes
    const size_t step = top >> 3;
    return (bytes + step - 1) / step * step;
}
struct BlockCache {   // one per process; the lock covers a vector push / pop, never a CUDA call of the steady state
    struct Key {
        cudaStream_t s;
        size_t bytes;
        bool operator==(const Key &o) const { return s == o.s && bytes == o.bytes; }
    };
    struct Hash {
        size_t operator()(const Key &k) const { return std::hash<const void *>()((const void *)k.s) * 1000003u ^ std::hash<size_t>()(k.bytes); }
    };
    std::mutex mu;
    std::unordered_map<Key, std::vector<void *>, Hash> free_;
    size_t cached_bytes = 0;   // what sits in free_
    size_t max_bytes = std::getenv("TNB_BLOCK_CACHE_MAX_GB") ? (size_t)(std::atof(std::getenv("TNB_BLOCK_CACHE_MAX_GB")) * (1ull << 30))
                                                             : (size_t)64 << 30;   // beyond this a released block goes back to the pool
};
static BlockCache &block_cache()
{
    static BlockCache *c = new BlockCache;  // never destroyed: blocks may be released while the process shuts down
    return *c;
}
void *block_acquire(size_t bytes, cudaStream_t s, size_t *got, cudaError_t *err)
{
    *err = cudaSuccess;
    if (!g_block_cache) {
        void *p = nullptr;
        *got = bytes;
        *err = cudaMallocAsync(&p, bytes, s);
        return *err == cudaSuccess ? p : nullptr;
    }
    const size_t cls = size_class(bytes);
    *got = cls;
    BlockCache &bc = block_cache();
    {
        std::lock_guard<std::mutex> g(bc.mu);
        auto it = bc.free_.find(BlockCache::Key{s, cls});
        if (it != bc.free_.end() && !it->second.empty()) {
            void *p = it->second.back();
            it->second.pop_back();
            bc.cached_bytes -= cls;
            return p;
        }
    }
    void *p = nullptr;
    *err = cudaMallocAsync(&p, cls, s);
    if (*err != cudaSuccess) {   // out of memory with blocks sitting in the cache: give them back and try once more
        cudaGetLastError();
        block_cache_trim();
        *err = cudaMallocAsync(&p, cls, s);
    }
    return *err == cudaSuccess ? p : nullptr;
}
void block_cache_trim()
{
    BlockCache &bc = block_cache();
    std::lock_guard<std::mutex> g(bc.mu);
    for (auto &kv : bc.free_) {
        for (void *q : kv.second) cudaFreeAsync(q, kv.first.s);
        kv.second.clear();
    }
    bc.cached_bytes = 0;
}
void block_release(void *p, size_t bytes, cudaStream_t s)
{
    if (!g_block_cache) { cudaFreeAsync(p, s); return; }
    BlockCache &bc = block_cache();
    {
        std::lock_guard<std::mutex> g(bc.mu);
        if (bc.cached_bytes + bytes <= bc.max_bytes) {
            bc.free_[BlockCache::Key{s, bytes}].push_back(p);
            bc.cached_bytes += bytes;
            return;
        }
    }
    cudaFreeAsync(p, s);
}

struct ProfClass {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    cudaEvent_t open = nullptr;
    int64_t units = 0, bytes = 0;
    double ms = 0.0;
    int64_t launches = 0;
};
static bool g_prof_on = false;
thread_local bool t_no_profile = false;  // the workers of tnb_subpoly_batch: the event timers belong to the calling thread
static ProfClass g_prof[TNB_PROF_CLASSES];
bool g_pdl = std::getenv("TNB_NO_PDL") == nullptr;
void prof_begin(int cls, cudaStream_t s)
{
    if (!g_prof_on || t_no_profile) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    g_prof[cls].open = e;
}
void prof_add(int cls, int64_t units, int64_t bytes)
{
    if (!g_prof_on || t_no_profile) return;
    g_prof[cls].units += units;
    g_prof[cls].bytes += bytes;
}
void prof_end(int cls, cudaStream_t s, int64_t units, int64_t bytes)
{
    if (!g_prof_on || t_no_profile || !g_prof[cls].open) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    g_prof[cls].ev.emplace_back(g_prof[cls].open, e);
    g_prof[cls].open = nullptr;
    g_prof[cls].units += units;
    g_prof[cls].bytes += bytes;
}
static void prof_collect(int cls)
{
    for (auto &pr : g_prof[cls].ev) {
        cudaEventSynchronize(pr.second);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, pr.first, pr.second);
        g_prof[cls].ms += ms;
        g_prof[cls].launches += 1;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    g_prof[cls].ev.clear();
}

constexpr int kThreads = 128;

// ---- kernels ---------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(kThreads) k_encode(const __grid_constant__ NetMeta n,
                                                     const float *__restrict__ xp, int64_t count,
                                                     float *__restrict__ enc)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        float p[3] = {xp[3 * i], xp[3 * i + 1], xp[3 * i + 2]};
        const int L = C::L(n);
#pragma unroll(C::kUnroll)
        for (int l = 0; l < C::kMaxL; ++l)
            if (l < L) {
                uint32_t cell[3];
                float frac[3];
                float2 f = encode_level(n, l, p, cell, frac);
                enc[i * (2 * L) + 2 * l] = f.x;
                enc[i * (2 * L) + 2 * l + 1] = f.y;
            }
    }
}

template <class C>
__global__ void __launch_bounds__(kThreads) k_outputs(const __grid_constant__ NetMeta n,
                                                      const float *__restrict__ x, int64_t count,
                                                      float *__restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        float p[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
        outputs_row<C>(n, p, out + i * n.R);
    }
}

template <class C>
__global__ void __launch_bounds__(kThreads) k_sdf_grad(const __grid_constant__ NetMeta n,
                                                       const float *__restrict__ x, int64_t count,
                                                       float *__restrict__ sdf, float *__restrict__ grad)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        float p[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
        float g[3];
        float t = sdf_grad<C>(n, p, g, grad != nullptr);
        sdf[i] = t;
        if (grad) {
            grad[3 * i] = g[0];
            grad[3 * i + 1] = g[1];
            grad[3 * i + 2] = g[2];
        }
    }
}

// outputs must be given (evaluated by k_outputs beforehand when the caller has none)
__global__ void __launch_bounds__(kThreads) k_region(const __grid_constant__ NetMeta n,
                                                     const float *__restrict__ x,
                                                     const float *__restrict__ outputs, int64_t count,
                                                     float eps, int8_t *__restrict__ signs,
                                                     int32_t *__restrict__ offset,
                                                     uint64_t *__restrict__ packed)
{
    const int R = n.R;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        float p[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
        float xp[3];
        preprocess(n, p, xp);
        uint64_t g = pack_grid(n, n.marks, xp, eps);
        uint64_t pos, neg;
        pack_signs(outputs + i * R, R, eps, pos, neg);
        if (packed) {
            packed[3 * i] = pos;
            packed[3 * i + 1] = neg;
            packed[3 * i + 2] = g;
        }
        if (offset)
            for (int d = 0; d < 3; ++d) offset[3 * i + d] = grid_off(g, d);
        if (signs) {
            int8_t *s = signs + i * (3 + R);
            for (int d = 0; d < 3; ++d) s[d] = (int8_t)grid_mask(g, d);
            for (int c = 0; c < R; ++c)
                s[3 + c] = (int8_t)(((pos >> c) & 1) ? 1 : (((neg >> c) & 1) ? -1 : 0));
        }
    }
}

// Dense lattice sweep: evaluate + bit-pack, nothing but 16 B per point leaves the SM.
template <class C, int MINB = 4>
__global__ void __launch_bounds__(kThreads, MINB) k_sweep_signs(const __grid_constant__ NetMeta n, float3 lo,
                                                          float3 step, int nx, int ny, int nz, LatticeStride ls,
                                                          float eps, ulonglong2 *__restrict__ packed)
{
    const int64_t count = (int64_t)nx * ny * nz;
    const int64_t first = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (first >= count) return;
    Lattice3 at(first, ls, nx, ny);
    if constexpr (C::kFixed) {
        // two lattice points per trip (i and i + stride): the MLP of both in packed FFMA2s (net_eval.cuh)
        for (int64_t i = first; i < count; i += 2 * stride) {
            const bool two = i + stride < count;
            float p0[3] = {__fmaf_rn((float)at.ix, step.x, lo.x), __fmaf_rn((float)at.iy, step.y, lo.y), __fmaf_rn((float)at.iz, step.z, lo.z)};
            at.advance();
            float p1[3] = {p0[0], p0[1], p0[2]};  // no second point: the first one twice, not stored
            if (two) { p1[0] = __fmaf_rn((float)at.ix, step.x, lo.x); p1[1] = __fmaf_rn((float)at.iy, step.y, lo.y); p1[2] = __fmaf_rn((float)at.iz, step.z, lo.z); }
            at.advance();
            float xp0[3], xp1[3];
            preprocess(n, p0, xp0);
            preprocess(n, p1, xp1);
            float2 pre[(C::kMaxLin - 1) * C::kMaxH];
            float2 o[2];
            forward_pair<C>(n, xp0, xp1, pre, o);
            SignWords w0, w1;
            const int H = C::H(n), NL = C::NLIN(n);
#pragma unroll(C::kUnroll)
            for (int l = 0; l < C::kMaxLin - 1; ++l)
                if (l < NL - 1) {
#pragma unroll(C::kUnroll)
                    for (int j = 0; j < C::kMaxH; ++j)
                        if (j < H) {
                            w0.add(pre[l * C::kMaxH + j].x, eps, l * H + j);
                            w1.add(pre[l * C::kMaxH + j].y, eps, l * H + j);
                        }
                }
            w0.add(o[1].x - o[0].x, eps, (NL - 1) * H);
            w1.add(o[1].y - o[0].y, eps, (NL - 1) * H);
            packed[i] = make_ulonglong2(w0.pos(), w0.neg());  // x fastest: coalesced 16 B stores
            if (two) packed[i + stride] = make_ulonglong2(w1.pos(), w1.neg());
        }
        return;
    }
    for (int64_t i = first; i < count; i += stride, at.advance()) {
        // x is the lane axis (table is x-fastest); the output index stays z-fastest
        const int ix = at.ix, iy = at.iy, iz = at.iz;
        float p[3] = {__fmaf_rn((float)ix, step.x, lo.x), __fmaf_rn((float)iy, step.y, lo.y),
                      __fmaf_rn((float)iz, step.z, lo.z)};
        float xp[3];
        preprocess(n, p, xp);
        float pre[(C::kMaxLin - 1) * C::kMaxH];
        float o[2];
        forward<C>(n, xp, pre, o);
        SignWords w;
        const int H = C::H(n), NL = C::NLIN(n);
#pragma unroll(C::kUnroll)
        for (int l = 0; l < C::kMaxLin - 1; ++l)
            if (l < NL - 1) {
#pragma unroll(C::kUnroll)
                for (int j = 0; j < C::kMaxH; ++j)
                    if (j < H) w.add(pre[l * C::kMaxH + j], eps, l * H + j);
            }
        w.add(o[1] - o[0], eps, (NL - 1) * H);
        packed[i] = make_ulonglong2(w.pos(), w.neg());  // x fastest: coalesced 16 B stores
    }
}

// ---- launchers shared with the complex code ----------------------------------------------
int launch_outputs(const tnb_net *net, const float *d_x, int64_t n, float *d_out, cudaStream_t s)
{
    if (n <= 0) return TNB_OK;
    unsigned g = grid_for(n, kThreads);
    if (net->fixed_cfg) k_outputs<CfgRef><<<g, kThreads, 0, s>>>(net->meta, d_x, n, d_out);
    else k_outputs<CfgAny><<<g, kThreads, 0, s>>>(net->meta, d_x, n, d_out);
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}

int launch_sdf_grad(const tnb_net *net, const float *d_x, int64_t n, float *d_sdf, float *d_grad,
                    cudaStream_t s)
{
    if (n <= 0) return TNB_OK;
    unsigned g = grid_for(n, kThreads);
    if (net->fixed_cfg) k_sdf_grad<CfgRef><<<g, kThreads, 0, s>>>(net->meta, d_x, n, d_sdf, d_grad);
    else k_sdf_grad<CfgAny><<<g, kThreads, 0, s>>>(net->meta, d_x, n, d_sdf, d_grad);
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}

int launch_region(const tnb_net *net, const float *d_x, const float *d_outputs, int64_t n, float eps,
                  int8_t *d_signs, int32_t *d_offset, uint64_t *d_packed, cudaStream_t s)
{
    if (n <= 0) return TNB_OK;
    k_region<<<grid_for(n, kThreads), kThreads, 0, s>>>(net->meta, d_x, d_outputs, n, eps, d_signs,
                                                       d_offset, d_packed);
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}

}  // namespace tnb

using namespace tnb;

// ---- C ABI ------------------------------------------------------------------------------------
extern "C" {

const char *tnb_last_error(void) { return g_error.c_str(); }
int tnb_version(void) { return 100; }
int tnb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int64_t tnb_launch_count(void) { return g_launches; }
void tnb_launch_count_reset(void) { g_launches = 0; }
void tnb_release_cached_blocks(void) { tnb::block_cache_trim(); }
int tnb_profile_enable(int on) { g_prof_on = on != 0; return TNB_OK; }
void tnb_profile_reset(void)
{
    for (int c = 0; c < TNB_PROF_CLASSES; ++c) {
        prof_collect(c);
        g_prof[c].ms = 0.0;
        g_prof[c].launches = 0;
        g_prof[c].units = 0;
        g_prof[c].bytes = 0;
    }
}
int tnb_profile_read(int cls, double *ms, int64_t *launches, int64_t *units, int64_t *bytes)
{
    if (cls < 0 || cls >= TNB_PROF_CLASSES) { set_error("tnb_profile_read: bad class"); return TNB_ERR_INVALID; }
    prof_collect(cls);
    if (ms) *ms = g_prof[cls].ms;
    if (launches) *launches = g_prof[cls].launches;
    if (units) *units = g_prof[cls].units;
    if (bytes) *bytes = g_prof[cls].bytes;
    return TNB_OK;
}

int tnb_net_create(const tnb_net_desc *d, tnb_net **out)
{
    if (!d || !out) { set_error("tnb_net_create: null argument"); return TNB_ERR_INVALID; }
    *out = nullptr;
    if (d->n_features != 2) { set_error("only n_features_per_level == 2 is supported (model.py:32)"); return TNB_ERR_UNSUPPORTED; }
    if (d->n_levels < 1 || d->n_levels > kMaxLevels || d->num_layers < 2 || d->num_layers > kMaxLinear ||
        d->num_hidden < 1 || d->num_hidden > kMaxHidden) {
        set_error("tnb_net_create: network shape out of range");
        return TNB_ERR_INVALID;
    }
    const int R = (d->num_layers - 1) * d->num_hidden + 1;
    if (R > 64) { set_error("more than 64 neurons: sign vectors do not fit the packed words"); return TNB_ERR_UNSUPPORTED; }
    if (d->n_marks < 2 || d->n_marks >= (1 << 20) - 2) { set_error("tnb_net_create: bad n_marks"); return TNB_ERR_INVALID; }
    if (tnb_device_count() == 0) { set_error("no CUDA device: this library has no CPU path"); return TNB_ERR_CUDA; }

    current_stream() = nullptr;  // created with blocking copies on the default stream
    tnb_net *net = new tnb_net();
    NetMeta &m = net->meta;
    memset(&m, 0, sizeof(m));
    m.L = d->n_levels; m.H = d->num_hidden; m.NLIN = d->num_layers; m.R = R;
    m.pre_scale = d->scale; m.pre_2s = d->scale * 2.0f; m.eps = d->eps; m.n_marks = d->n_marks;
    {
        int ex = 0;
        const float mant = std::frexp(m.pre_2s, &ex);
        m.pre_pow2 = (mant == 0.5f && ex > -100 && ex < 100) ? 1 : 0;
        m.pre_inv = 1.0f / m.pre_2s;
    }
    // level layout exactly as tiny-cuda-nn's GridEncoding constructor derives it
    const float log2_pls = std::log2((float)d->per_level_scale);
    uint64_t total = 0;
    for (int l = 0; l < m.L; ++l) {
        float scale = std::exp2((float)l * log2_pls) * (float)d->base_resolution - 1.0f;
        uint32_t res = (uint32_t)std::ceil(scale) + 1u;
        const uint32_t max_params = 0xFFFFFFFFu / 2;
        uint32_t n = std::pow((float)res, 3.0f) > (float)max_params ? max_params : res * res * res;
        n = (n + 7u) / 8u * 8u;
        uint32_t cap = 1u << d->log2_hashmap;
        if (n > cap) n = cap;
        // which index path reproduces grid_index() for this level (common.cuh)
        const uint64_t r1 = res, r2 = r1 * r1, r3 = r2 * r1;
        uint32_t mode = kLevelGeneric;
        if (r1 <= n && r2 <= n && r3 <= n && r3 < (1ull << 32)) mode = kLevelDense;
        else if ((n & (n - 1u)) == 0u) mode = kLevelHashPow2;
        m.lvl[l] = LevelMeta{scale, res, n, (uint32_t)total, mode, (uint32_t)(r2 & 0xFFFFFFFFull)};
        net->h_scale.push_back(scale); net->h_res.push_back(res); net->h_size.push_back(n);
        net->h_off.push_back((uint32_t)total);
        total += n;
    }
    if ((int64_t)total * 2 != d->table_len) {
        set_error("tnb_net_create: table_len " + std::to_string(d->table_len) + " != 2 * " + std::to_string(total));
        delete net;
        return TNB_ERR_INVALID;
    }
    int64_t mlp_len = 0;
    for (int i = 0; i < m.NLIN; ++i) {
        int ni = i == 0 ? 2 * m.L : m.H, no = i == m.NLIN - 1 ? 2 : m.H;
        mlp_len += (int64_t)no * ni + no;
    }
    if (mlp_len != d->mlp_len) { set_error("tnb_net_create: mlp_len mismatch"); delete net; return TNB_ERR_INVALID; }

    cudaError_t e;
    if ((e = net->table.reserve(total)) != cudaSuccess || (e = net->mlp.reserve(mlp_len)) != cudaSuccess ||
        (e = net->marks.reserve(d->n_marks)) != cudaSuccess) {
        delete net;
        return cuda_fail(e, "cudaMalloc(net)", __FILE__, __LINE__);
    }
    if ((e = cudaMemcpy(net->table.p, d->table, total * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(net->mlp.p, d->mlp, mlp_len * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(net->marks.p, d->marks, d->n_marks * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
        delete net;
        return cuda_fail(e, "cudaMemcpy(net)", __FILE__, __LINE__);
    }
    m.table = net->table.p; m.mlp = net->mlp.p; m.marks = net->marks.p;
    net->h_marks.assign(d->marks, d->marks + d->n_marks);
    if (mlp_len <= kMlpParamMax) {
        memcpy(m.mlp_c, d->mlp, mlp_len * sizeof(float));
        m.mlp_in_param = 1;
    }
    net->fixed_cfg = m.mlp_in_param && m.L == 4 && m.H == 16 && m.NLIN == 3;
    *out = net;
    return TNB_OK;
}

void tnb_net_destroy(tnb_net *net) { delete net; }
int tnb_net_num_outputs(const tnb_net *net) { return net ? net->meta.R : 0; }

int tnb_net_level_layout(const tnb_net *net, float *scale, uint32_t *res, uint32_t *size, uint32_t *offset)
{
    if (!net) { set_error("null net"); return TNB_ERR_INVALID; }
    for (int l = 0; l < net->meta.L; ++l) {
        if (scale) scale[l] = net->h_scale[l];
        if (res) res[l] = net->h_res[l];
        if (size) size[l] = net->h_size[l];
        if (offset) offset[l] = net->h_off[l];
    }
    return TNB_OK;
}

int tnb_grid_encode(const tnb_net *net, const float *d_xp, int64_t n, float *d_enc, void *stream)
{
    if (!net || (n > 0 && (!d_xp || !d_enc))) { set_error("tnb_grid_encode: null argument"); return TNB_ERR_INVALID; }
    if (n <= 0) return TNB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned g = grid_for(n, kThreads);
    if (net->fixed_cfg) k_encode<CfgRef><<<g, kThreads, 0, s>>>(net->meta, d_xp, n, d_enc);
    else k_encode<CfgAny><<<g, kThreads, 0, s>>>(net->meta, d_xp, n, d_enc);
    TNB_LAUNCH_CHECK();
    return TNB_OK;
}

int tnb_net_outputs(const tnb_net *net, const float *d_x, int64_t n, float *d_out, void *stream)
{
    if (!net || (n > 0 && (!d_x || !d_out))) { set_error("tnb_net_outputs: null argument"); return TNB_ERR_INVALID; }
    return launch_outputs(net, d_x, n, d_out, (cudaStream_t)stream);
}

int tnb_net_sdf_grad(const tnb_net *net, const float *d_x, int64_t n, float *d_sdf, float *d_grad, void *stream)
{
    if (!net || (n > 0 && (!d_x || !d_sdf))) { set_error("tnb_net_sdf_grad: null argument"); return TNB_ERR_INVALID; }
    return launch_sdf_grad(net, d_x, n, d_sdf, d_grad, (cudaStream_t)stream);
}

int tnb_net_region(const tnb_net *net, const float *d_x, const float *d_outputs, int64_t n, float eps,
                   int8_t *d_signs, int32_t *d_offset, uint64_t *d_packed, void *stream)
{
    if (!net || (n > 0 && !d_x)) { set_error("tnb_net_region: null argument"); return TNB_ERR_INVALID; }
    if (n <= 0) return TNB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    current_stream() = s;
    DevBuf<float> tmp;
    if (!d_outputs) {
        TNB_CUDA(tmp.reserve((size_t)n * net->meta.R));
        int rc = launch_outputs(net, d_x, n, tmp.p, s);
        if (rc) return rc;
        d_outputs = tmp.p;
    }
    int rc = launch_region(net, d_x, d_outputs, n, eps, d_signs, d_offset, d_packed, s);
    if (rc) return rc;
    return TNB_OK;
}

int tnb_sweep_signs(const tnb_net *net, const float lo[3], const float hi[3], const int32_t nn[3], float eps,
                    uint64_t *d_packed, void *stream)
{
    if (!net || !lo || !hi || !nn || !d_packed) { set_error("tnb_sweep_signs: null argument"); return TNB_ERR_INVALID; }
    if (nn[0] < 1 || nn[1] < 1 || nn[2] < 1) { set_error("tnb_sweep_signs: empty lattice"); return TNB_ERR_INVALID; }
    if ((int64_t)nn[0] * nn[1] >= 0x7fffffff) { set_error("tnb_sweep_signs: a lattice plane of 2^31 points or more"); return TNB_ERR_UNSUPPORTED; }
    float3 l = make_float3(lo[0], lo[1], lo[2]);
    float3 st = make_float3(nn[0] > 1 ? (hi[0] - lo[0]) / (float)(nn[0] - 1) : 0.0f,
                            nn[1] > 1 ? (hi[1] - lo[1]) / (float)(nn[1] - 1) : 0.0f,
                            nn[2] > 1 ? (hi[2] - lo[2]) / (float)(nn[2] - 1) : 0.0f);
    int64_t count = (int64_t)nn[0] * nn[1] * nn[2];
    cudaStream_t s = (cudaStream_t)stream;
    unsigned g = grid_for(count, kThreads, kSMs * 32);
    const LatticeStride ls = lattice_stride((int64_t)g * kThreads, nn[0], nn[1]);
    prof_begin(TNB_PROF_SIGN_SWEEP, s);
    static const int minb = std::getenv("TNB_SIGNS_MINB") ? std::atoi(std::getenv("TNB_SIGNS_MINB")) : 4;  // A/B: registers vs warps per SM
    if (net->fixed_cfg && minb == 3)
        k_sweep_signs<CfgRef, 3><<<g, kThreads, 0, s>>>(net->meta, l, st, nn[0], nn[1], nn[2], ls, eps, (ulonglong2 *)d_packed);
    else if (net->fixed_cfg && minb == 5)
        k_sweep_signs<CfgRef, 5><<<g, kThreads, 0, s>>>(net->meta, l, st, nn[0], nn[1], nn[2], ls, eps, (ulonglong2 *)d_packed);
    else if (net->fixed_cfg)
        k_sweep_signs<CfgRef><<<g, kThreads, 0, s>>>(net->meta, l, st, nn[0], nn[1], nn[2], ls, eps, (ulonglong2 *)d_packed);
    else
        k_sweep_signs<CfgAny><<<g, kThreads, 0, s>>>(net->meta, l, st, nn[0], nn[1], nn[2], ls, eps, (ulonglong2 *)d_packed);
    TNB_LAUNCH_CHECK();
    prof_end(TNB_PROF_SIGN_SWEEP, s, count, count * 16);
    return TNB_OK;
}

}  // extern "C"

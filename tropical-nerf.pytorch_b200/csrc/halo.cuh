// halo.cuh -- what crosses a slab boundary when ONE object is sharded over several GPUs.
//
// The reference's subpoly_ is cell-local except for three decisions:
//   * "nothing crossed -> skip the step"  (subpoly.py:110-111)      is taken on ALL edges,
//   * the failover override                (subpoly_debug.py:41-49)  is taken on ALL new vertices,
//   * a vertex survives pruning when ANY of its edges survives (subpoly.py:268-272), and a vertex
//     on a shared slab plane has edges on both sides of it.
// So every hyperplane step ends in one exchange: a status word to every rank and one liveness
// byte per shared-plane vertex to the two neighbours.  The messages are written straight into the
// receiver's memory (peer-mapped mailboxes over NVLink, or plain device memory when several slabs
// run on one GPU) by a one-CTA-per-destination kernel; the receiver spins on a sequence word.
// No host involvement, no NCCL call on the data path.
//
// Shared-plane vertices are matched by ORDER: both neighbours hold the same plane vertices in the
// same relative order (vertex and edge numbering on a rank is the restriction of the single-GPU
// numbering), so the k-th tagged vertex here is the k-th tagged vertex there.  The
// count travels with the message and a mismatch raises a sticky error.
#pragma once
#include "common.cuh"

namespace tnb {

constexpr int kHaloMaxWorld = 64;
constexpr size_t kHaloHeader = 16;  // [0] sequence word, [4] count, [8..16) spare
enum { kWordRaw = 1, kWordFlag = 2, kWordSticky = 4 };
enum { kStickyHaloTimeout = 2, kStickyHaloMismatch = 4, kStickyHaloPeer = 8, kStickyHaloPayload = 16 };

// Mailbox of one rank:
//   status  [2 parities][kHaloMaxWorld] u32   (sequence << 8 | word) written by every rank
//   inbox   [2 sides][2 parities] { header, payload }   side 0: from the lower neighbour, 1: upper
__host__ __device__ inline size_t halo_msg_bytes(size_t payload) { return kHaloHeader + ((payload + 15) / 16) * 16; }
__host__ __device__ inline size_t halo_status_bytes() { return 2 * kHaloMaxWorld * sizeof(uint32_t); }
__host__ __device__ inline size_t halo_box_bytes(size_t payload) { return halo_status_bytes() + 4 * halo_msg_bytes(payload); }
__host__ __device__ inline uint32_t *halo_status(unsigned char *box, int parity, int src)
{
    return reinterpret_cast<uint32_t *>(box) + parity * kHaloMaxWorld + src;
}
__host__ __device__ inline unsigned char *halo_inbox(unsigned char *box, size_t payload, int side, int parity)
{
    return box + halo_status_bytes() + (size_t)(side * 2 + parity) * halo_msg_bytes(payload);
}

struct HaloArgs {
    int rank, world, parity;
    uint32_t seq;
    size_t payload;
    unsigned char *boxes[kHaloMaxWorld];
    const unsigned char *stage[2];  // outgoing liveness bytes: lower, upper
    int *stage_count;               // [0] lower count, [1] upper count, [2] global word (out), [3] spare
    int *cnt;                       // the complex's counter block (sticky bits)
    int has[2];                     // neighbour below / above
    int word_raw_index, word_flag_index, sticky_index;
    long long timeout_ns;
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long global_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// grid = 2 + 1 CTAs: CTA 0 / 1 write the liveness bytes into the lower / upper neighbour's inbox,
// CTA 2 writes the status word into every rank's mailbox.
__global__ void __launch_bounds__(256) k_halo_send(const HaloArgs a)
{
    const int sticky = a.cnt[a.sticky_index];
    if (blockIdx.x < 2) {
        const int side = blockIdx.x;
        if (!a.has[side]) return;
        const int peer = a.rank + (side == 0 ? -1 : 1);
        // my lower neighbour receives this in its "from the upper neighbour" inbox and vice versa
        unsigned char *dst = halo_inbox(a.boxes[peer], a.payload, side == 0 ? 1 : 0, a.parity);
        int n = a.stage_count[side];
        if (!sticky && n >= 0 && (size_t)n > a.payload && threadIdx.x == 0) atomicOr(a.cnt + a.sticky_index, kStickyHaloPayload);
        if (sticky || n < 0 || (size_t)n > a.payload) n = -1;  // poison: the receiver raises its own sticky bit
        if (n > 0) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(a.stage[side]);
            uint32_t *d = reinterpret_cast<uint32_t *>(dst + kHaloHeader);
            const int words = (n + 3) / 4;
            for (int i = threadIdx.x; i < words; i += blockDim.x) d[i] = src[i];
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            reinterpret_cast<int *>(dst)[1] = n;
            __threadfence_system();
            st_release_sys(reinterpret_cast<uint32_t *>(dst), a.seq);
        }
    } else {
        uint32_t w = 0;
        if (a.cnt[a.word_raw_index] > 0) w |= kWordRaw;
        if (a.cnt[a.word_flag_index]) w |= kWordFlag;
        if (sticky) w |= kWordSticky;
        for (int r = threadIdx.x; r < a.world; r += blockDim.x)
            st_release_sys(halo_status(a.boxes[r], a.parity, a.rank), (a.seq << 8) | w);
    }
}

// One CTA.  Waits for the status words of all ranks and the neighbour messages of this exchange;
// leaves the OR of the status words in stage_count[2].  A peer that never answers costs one
// timeout and a sticky bit, not a hung GPU.
__global__ void __launch_bounds__(256) k_halo_recv(const HaloArgs a)
{
    __shared__ uint32_t s_word;
    __shared__ int s_fail;
    if (threadIdx.x == 0) { s_word = 0; s_fail = 0; }
    __syncthreads();
    unsigned char *box = a.boxes[a.rank];
    const long long t0 = global_ns();
    const bool already = a.cnt[a.sticky_index] & kStickyHaloTimeout;
    // status words
    for (int r = threadIdx.x; r < a.world; r += blockDim.x) {
        const uint32_t *p = halo_status(box, a.parity, r);
        uint32_t v = ld_acquire_sys(p);
        while ((v >> 8) != (a.seq & 0xFFFFFFu)) {
            if (already || global_ns() - t0 > a.timeout_ns) { atomicOr(&s_fail, kStickyHaloTimeout); break; }
            __nanosleep(64);
            v = ld_acquire_sys(p);
        }
        if ((v >> 8) == (a.seq & 0xFFFFFFu)) atomicOr(&s_word, v & 0xFFu);
    }
    // neighbour messages
    if (threadIdx.x < 2 && a.has[threadIdx.x]) {
        const int side = threadIdx.x;
        const unsigned char *in = halo_inbox(box, a.payload, side, a.parity);
        const uint32_t *p = reinterpret_cast<const uint32_t *>(in);
        uint32_t v = ld_acquire_sys(p);
        while (v != a.seq) {
            if (already || global_ns() - t0 > a.timeout_ns) { atomicOr(&s_fail, kStickyHaloTimeout); break; }
            __nanosleep(64);
            v = ld_acquire_sys(p);
        }
        if (v == a.seq) {
            const int n = reinterpret_cast<const volatile int *>(in)[1];
            if (n < 0) atomicOr(&s_fail, kStickyHaloPeer);
            else if (n != a.stage_count[side]) atomicOr(&s_fail, kStickyHaloMismatch);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_fail) atomicOr(a.cnt + a.sticky_index, s_fail);
        if ((s_word & kWordSticky) && !a.cnt[a.sticky_index]) atomicOr(a.cnt + a.sticky_index, kStickyHaloPeer);
        a.stage_count[2] = (int)s_word;
    }
}
#endif  // __CUDACC__

}  // namespace tnb

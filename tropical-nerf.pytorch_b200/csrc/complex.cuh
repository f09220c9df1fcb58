// complex.cuh -- the device-resident polyhedral complex behind tnb_complex.
#pragma once
#include "runtime.cuh"

// Layout in HBM (structure of arrays, all rows indexed by vertex / edge number):
//   vert  [V][3] f32   world-space position
//   out   [V][R] f32   cached Net.forward(gather=True) row (subpoly.py:92-95, :275)
//   sig   [V][3] u64   packed region indicator {pos bits, neg bits, grid word}
//   edges [E]    int2  vertex pair (first, second) in the reference's column order
//   used  [V]    int   liveness: pruning (subpoly.py:268-272) only MARKS a vertex dead; its row stays
//                      where it is and no edge refers to it any more.  Every order the reference
//                      derives from vertex numbers (ascending partner lists, face-row leaders, the
//                      final numbering) depends on relative order only, which lazy deletion keeps, so
//                      the rows are compacted once, when somebody reads the complex
//                      (complex_compact), not after each of the 33 hyperplanes.
// The edge array is double-buffered: pruning compacts it from one half into the other (order
// preserving) and flips.  The vertex arrays are double-buffered for complex_compact and for the
// curve path's temporary rows; the liveness array is double-buffered so that a step that is
// abandoned half-way (slab exchange: no slab crossed the plane) leaves it intact.
// One entry of a cell bucket (generation-stamped linked lists over the marks-grid cells).  The
// candidate's packed sign vector travels with the link, so walking a list costs ONE dependent
// 32-byte load per hop instead of link -> candidate -> signature.
struct alignas(32) tnb_bucket_rec {
    int next;  // next record of the same cell, -1 = end
    int v;     // vertex number
    uint64_t pos, neg, grd;
};

// Slab sharding (one object split over several GPUs along the first grid axis): the complex of
// one rank holds the cells [x_lo, x_hi-1]; the planes x_lo / x_hi are shared with the neighbours.
// What crosses a shared plane is exchanged once per hyperplane step through peer-mapped
// mailboxes (see halo.cuh): the liveness of the plane's vertices (neighbours) and one status
// word (all ranks).
struct tnb_halo {
    bool enabled = false;
    int rank = 0, world = 1;
    int x_lo = -1, x_hi = -1;        // marks-grid plane numbers bounding the slab (inclusive)
    bool tag_lower = false, tag_upper = false;  // a neighbour exists below / above
    unsigned char *boxes[64] = {};   // mailbox base pointer of every rank (own one included)
    size_t payload = 0;              // bytes of one neighbour message
    uint32_t seq = 0;                // exchange number (both parities of the mailboxes alternate)
    tnb::DevBuf<int> slot;           // [Vcap] position of a tagged vertex in its plane list
    tnb::DevBuf<unsigned char> stage[2];  // [payload] outgoing liveness bytes (lower, upper)
    tnb::DevBuf<int> stage_count;    // [4] plane vertices (lower, upper), exchange status word, spare
};

struct tnb_complex {
    int R = 0;
    int64_t V = 0, E = 0;
    size_t Vcap = 0, Ecap = 0;
    int vcur = 0, ecur = 0, acur = 0;
    // bit j set = some edge of the current complex has ends on opposite sides of hyperplane j (from the
    // packed signs, OR-ed over the edges by the pruning pass): a clear bit means step j is a no-op
    // (subpoly.py:110-111) and is skipped without looking at the edges again
    uint64_t cross = ~0ull;
    bool cross_stale = false;       // the device holds a newer mask than `cross`
    bool maybe_dead = false;        // a prune ran since the last compaction: rows of dead vertices may exist
    bool bytes_by_half = false;     // `bytes` was filled by the device-driven step stream: front half -> new vertices, back half -> pairs
    tnb::DevBuf<float> vert[2], out[2];
    tnb::DevBuf<uint64_t> sig[2];
    tnb::DevBuf<int2> edges[2];
    tnb::DevBuf<unsigned char> tag[2];  // [Vcap] lineage on a shared slab plane: bit0 lower, bit1 upper
    tnb_halo halo;
    // scratch
    tnb::DevBuf<int> split_list;    // [Ecap]   edge numbers being split
    tnb::DevBuf<uint64_t> bmask;    // [Ecap]   override mask of each new vertex
    tnb::DevBuf<int> cand;          // [Vcap]   candidate vertex numbers (hits, then new)
    tnb::DevBuf<int> pcount;        // [Vcap]   partners per candidate
    tnb::DevBuf<int> poff;          // [Vcap]   exclusive scan of pcount
    tnb::DevBuf<int> pcache;        // [32*Vcap] first partners of each candidate, left by the count pass
    tnb::DevBuf<tnb_bucket_rec> next;  // [8*Vcap] bucket chains
    tnb::DevBuf<unsigned long long> head;  // [n_cells] (stamp << 32 | record); the contiguous form (cells.cuh) reads it as int2 {count, base}
    tnb::DevBuf<int2> cslot;        // [8*Vcap] contiguous form: {cell, index in the cell's segment} of every (candidate, cell)
    int bucket_mode = 0;            // what `head` holds: 0 all-zero, 1 generation stamps (linked lists), 2 counts of a step in flight
    tnb::DevBuf<int> used[2];       // [Vcap]   vertex referenced by a kept edge; used[acur] = liveness of the current complex
    tnb::DevBuf<int> remap;         // [Vcap]
    tnb::DevBuf<int> block_sums;    // [kScanMaxBlocks]
    tnb::DevBuf<uint32_t> scan_mask;  // one bit per edge / vertex: masked compactions (scan.cuh)
    tnb::DevBuf<int> counters;      // [16] device counters
    tnb::DevBuf<int> gd;            // curve path: walks of the gradient-descent repair (repair.cuh)
    tnb::DevBuf<unsigned long long> bytes;  // [4] algorithmic bytes (front / back half of a step) and units accumulated by the fused kernels
    int *h_counters = nullptr;      // pinned mirror (per thread, not owned)
    bool counts_stale = false;      // V/E are upper bounds; exact sizes are in counters[C_V], [C_E]
    int sticky_rc = 0;              // latched device error (capacity, curve path, slab exchange): every later call returns it
    std::string sticky_msg;
    cudaStream_t stream = nullptr;  // stream of the last call
    uint32_t stamp = 0;             // bucket generation
    int64_t n_cells = 0;
    int cell_dim = 0;

    float *cvert() { return vert[vcur].p; }
    float *cout_() { return out[vcur].p; }
    uint64_t *csig() { return sig[vcur].p; }
    int2 *cedges() { return edges[ecur].p; }
    const float *cvert() const { return vert[vcur].p; }
    const float *cout_() const { return out[vcur].p; }
    const uint64_t *csig() const { return sig[vcur].p; }
    const int2 *cedges() const { return edges[ecur].p; }
    int *calive() { return used[acur].p; }
    const int *calive() const { return used[acur].p; }
    ~tnb_complex();
};

namespace tnb {
int complex_alloc(tnb_complex *c, const tnb_net *net, size_t Vcap, size_t Ecap);
int complex_reserve(tnb_complex *c, size_t Vneed, size_t Eneed, cudaStream_t s);
int complex_sync_counts(tnb_complex *c, cudaStream_t s);
int complex_compact(tnb_complex *c, cudaStream_t s);
int launch_outputs(const tnb_net *net, const float *d_x, int64_t n, float *d_out, cudaStream_t s);
int launch_sdf_grad(const tnb_net *net, const float *d_x, int64_t n, float *d_sdf, float *d_grad, cudaStream_t s);
int launch_region(const tnb_net *net, const float *d_x, const float *d_outputs, int64_t n, float eps,
                  int8_t *d_signs, int32_t *d_offset, uint64_t *d_packed, cudaStream_t s);
// slab exchange of vertex liveness (complex.cu, halo.cuh)
int halo_publish_used(tnb_complex *c, int64_t V, const int *used, cudaStream_t s);
int halo_merge_used(tnb_complex *c, int64_t V, int *used, cudaStream_t s);
extern double g_capacity_factor;
extern thread_local double t_capacity_scale;
extern thread_local bool t_no_profile;   // tnb_profile_* timers are skipped on this thread
double capacity_factor();
extern thread_local int64_t t_cluster_max_items;  // >= 0: overrides tnb_set_cluster_max_items for the calling thread
}  // namespace tnb

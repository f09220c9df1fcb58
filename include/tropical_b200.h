/*
 * tropical_b200.h -- C ABI of the B200-native polyhedral-complex mesh extractor.
 *
 * Drop-in boundary for the reference's mesh-extraction path.  The reference is pure
 * Python over torch + tiny-cuda-nn; each entry point below names the reference
 * function (file:line under the reference tree) whose work it replaces.  Plain
 * pointers and sizes only; no torch types.  Pointers prefixed d_ are CUDA device
 * pointers, h_ are host pointers.  `stream` is a cudaStream_t passed as void*
 * (NULL = the legacy default stream).  Every function returns TNB_OK (0) or a
 * negative TNB_ERR_* code; tnb_last_error() gives the message of the last failure on
 * the calling thread.
 *
 * There is no CPU implementation behind this interface: without a CUDA device every
 * compute entry point returns TNB_ERR_CUDA.
 */
#ifndef TROPICAL_B200_H
#define TROPICAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNB_OK 0
#define TNB_ERR_INVALID (-1)   /* bad argument / unsupported configuration */
#define TNB_ERR_CUDA (-2)      /* CUDA runtime error (no device, launch failure, ...) */
#define TNB_ERR_CAPACITY (-3)  /* device work buffers too small; raise capacity and retry */
#define TNB_ERR_UNSUPPORTED (-4)

#define TNB_MAX_LEVELS 16
#define TNB_MAX_LINEAR 8
#define TNB_MAX_HIDDEN 64

/* Host-side description of one trilinear network = the state of the reference's
 * `Net` (tropical/stanford/model.py:18-50) and its `TropicalHashGrid`
 * (tropical/tropical.py:20-44).  All arrays are host pointers, copied at create. */
typedef struct tnb_net_desc {
    int32_t n_levels;         /* L  (tropical.py:21)                                   */
    int32_t n_features;       /* F, must be 2 (model.py:32)                            */
    int32_t log2_hashmap;     /* T                                                     */
    int32_t base_resolution;  /* N_min                                                 */
    double per_level_scale;   /* b  (tropical.py:31)                                   */
    int32_t num_layers;       /* number of nn.Linear layers (model.py:39-50)           */
    int32_t num_hidden;       /* H                                                     */
    float scale;              /* Net.scale (model.py:34)                               */
    float eps;                /* Net.eps                                               */
    const float *table;       /* tcnn params, [sum(level sizes) * F]                   */
    int64_t table_len;        /* number of floats in `table` (checked)                 */
    const float *mlp;         /* per layer: weight [out][in] row-major, then bias      */
    int64_t mlp_len;
    const float *marks;       /* TropicalHashGrid.marks (tropical.py:49-79)            */
    int32_t n_marks;
} tnb_net_desc;

typedef struct tnb_net tnb_net;          /* device-resident network                    */
typedef struct tnb_complex tnb_complex;  /* device-resident vertices/edges/outputs     */
typedef struct tnb_mesh tnb_mesh;        /* device-resident extracted surface mesh     */
typedef struct tnb_sweep tnb_sweep;      /* |sdf| / max |grad| of a marks-grid slab    */

const char *tnb_last_error(void);
int tnb_version(void);
int tnb_device_count(void);

/* ---- network ------------------------------------------------------------------ */
int tnb_net_create(const tnb_net_desc *desc, tnb_net **out);
void tnb_net_destroy(tnb_net *net);
int tnb_net_num_outputs(const tnb_net *net); /* R = (num_layers-1)*H + 1 */
/* per-level layout tiny-cuda-nn derives; arrays of n_levels entries (host) */
int tnb_net_level_layout(const tnb_net *net, float *scale, uint32_t *res, uint32_t *size,
                         uint32_t *offset);

/* TropicalHashGrid.forward (tropical.py:46-47 -> tcnn.Encoding): d_xp [n,3] in grid
 * coordinates -> d_enc [n, L*F] */
int tnb_grid_encode(const tnb_net *net, const float *d_xp, int64_t n, float *d_enc, void *stream);

/* torch.cat(Net.forward(x, gather=True)[1], -1) (model.py:52-76): d_x [n,3] world
 * coordinates -> d_out [n,R] hidden pre-activations then (o1 - o0) */
int tnb_net_outputs(const tnb_net *net, const float *d_x, int64_t n, float *d_out, void *stream);

/* Net.sdf(x)[:,0] (model.py:84-88) and d sdf / d x as autograd returns it in
 * TropicalHashGrid.skeleton (tropical.py:190-195) / Net.normal (model.py:105-123).
 * d_grad may be NULL. */
int tnb_net_sdf_grad(const tnb_net *net, const float *d_x, int64_t n, float *d_sdf,
                     float *d_grad, void *stream);

/* Net.region (model.py:90-103) + TropicalHashGrid.region (tropical.py:227-236).
 * d_outputs may be NULL (then evaluated).  d_signs [n, 3+R] int8: grid masks 0/1 then
 * neuron signs -1/0/+1.  d_offset [n,3] int32.  Either output may be NULL.
 * d_packed (optional) [n,3] uint64: {positive-sign bits, negative-sign bits, grid word}
 * -- the bit-packed form the subdivision kernels use. */
int tnb_net_region(const tnb_net *net, const float *d_x, const float *d_outputs, int64_t n,
                   float eps, int8_t *d_signs, int32_t *d_offset, uint64_t *d_packed,
                   void *stream);

/* Evaluation sweep over a dense lattice (BASELINE config "batched trilinear network
 * eval + sign-vector sweep"): points lo + (hi-lo)*i/(n-1) per axis; writes the packed
 * sign vector of every lattice point, d_packed [nz][ny][nx][2] uint64 {positive bits,
 * negative bits} (zero = neither) -- x is the fastest index, like the hash table. */
int tnb_sweep_signs(const tnb_net *net, const float lo[3], const float hi[3], const int32_t n[3],
                    float eps, uint64_t *d_packed, void *stream);

/* ---- polyhedral complex -------------------------------------------------------- */
/* TropicalHashGrid.skeleton(net, unit) (tropical.py:158-225, distance pruning).  When
 * no grid edge survives, the hypercube of subpoly.py:51-52 / :731-750 with half-size
 * `size` is returned instead (what subpoly() does next); size <= 0 disables the fallback
 * and returns the empty complex (what skeleton() itself returns, tropical.py:208-209). */
int tnb_skeleton(const tnb_net *net, int32_t unit, float size, tnb_complex **out, void *stream);
/* Build a complex from caller arrays (device pointers): vertices [V,3] f32, edges
 * [E,2] i64.  Outputs are evaluated. */
int tnb_complex_from_arrays(const tnb_net *net, const float *d_vertices, int64_t V,
                            const int64_t *d_edges, int64_t E, tnb_complex **out, void *stream);
/* The `outputs_` argument the reference threads through subpoly_ / extract_skeleton / extract_faces
 * (tropical/subpoly.py:92-95, :556-560, :607): replace the complex's evaluated network rows by the
 * caller's [V][R] rows (they may carry the exact zeros of the failover override,
 * subpoly_debug.py:41-49); the packed sign vectors are rebuilt from them. */
int tnb_complex_write_outputs(const tnb_net *net, tnb_complex *c, const float *d_outputs, void *stream);
void tnb_complex_destroy(tnb_complex *c);
/* -1: the complex carries a device-side error (capacity, curve path, slab exchange); the same
 * TNB_ERR_* code is returned by every later call on it and tnb_last_error() says why */
int64_t tnb_complex_num_vertices(const tnb_complex *c);
int64_t tnb_complex_num_edges(const tnb_complex *c);
/* copy out to device buffers (any may be NULL): vertices [V,3] f32, edges [E,2] i64,
 * outputs [V,R] f32 */
int tnb_complex_read(const tnb_complex *c, float *d_vertices, int64_t *d_edges, float *d_outputs,
                     void *stream);

/* subpoly_(vertices, edges, net, l, h, eps, outputs, force=...) (subpoly.py:90-279):
 * subdivide every edge the hyperplane of neuron (l,h) crosses, connect the new
 * vertices, prune.  force != 0 is the planar path (reference default). */
int tnb_subpoly_step(const tnb_net *net, tnb_complex *c, int32_t l, int32_t h, float eps,
                     int32_t force, void *stream);

/* The loop over hyperplanes of subpoly() (subpoly.py:58-72) in one call: lh[2*i], lh[2*i+1] =
 * (l, h) of step i.  Same result as n_steps calls of tnb_subpoly_step; a small complex (planar
 * or curve-approximation path) runs all of them in ONE persistent cooperative launch, with no
 * host round trip between steps (optionally, planar path, inside one thread-block cluster:
 * tnb_set_cluster_max_items); a large complex (planar path) runs them as a device-driven stream
 * of launches: the device picks the next hyperplane that crosses an edge, the host never waits
 * (tnb_set_fused_max_items). */
int tnb_subpoly_steps(const tnb_net *net, tnb_complex *c, const int32_t *lh, int32_t n_steps, float eps,
                      int32_t force, void *stream);

/* extract_skeleton + extract_faces (subpoly.py:556-652). */
int tnb_extract_mesh(const tnb_net *net, const tnb_complex *c, float eps, tnb_mesh **out,
                     void *stream);
void tnb_mesh_destroy(tnb_mesh *m);
int64_t tnb_mesh_num_vertices(const tnb_mesh *m);
int64_t tnb_mesh_num_edges(const tnb_mesh *m);
int64_t tnb_mesh_num_triangles(const tnb_mesh *m);   /* rows of faces_with_indices */
int64_t tnb_mesh_num_polygons(const tnb_mesh *m);
int64_t tnb_mesh_polygon_width(const tnb_mesh *m);
/* copy out (device buffers, any may be NULL): vertices [V,3] f32, edges [E,2] i64,
 * triangles [T,3] i64 (faces_with_indices), faces [T,3,3] f32 (triangle corner
 * positions, first return value of subpoly()), polygons [P,W] i64 (-1 padded, the
 * angle-sorted face rows the triangles fan out of). */
int tnb_mesh_read(const tnb_mesh *m, float *d_vertices, int64_t *d_edges, int64_t *d_triangles,
                  float *d_faces, int64_t *d_polygons, void *stream);

/* subpoly(net, d, size, eps, force) (subpoly.py:23-86): the whole path. */
int tnb_subpoly(const tnb_net *net, float size, float eps, int32_t force, int32_t unit,
                tnb_mesh **out, void *stream);
/* same, end to end with HOST buffers: runs the path and copies the mesh to the host.
 * Call once with all buffers NULL to get the sizes in n_out[4] = {V, T, P, W}, then
 * tnb_mesh_read_host. */
int tnb_mesh_read_host(const tnb_mesh *m, float *h_vertices, int64_t *h_triangles,
                       float *h_faces, int64_t *h_polygons);

/* ---- slab sharding: ONE object split over several GPUs along the first grid axis ----
 * Rank r holds the marks-grid planes [x_lo, x_hi] (cells x_lo .. x_hi-1); neighbouring
 * slabs share one plane.  The reference's path is cell-local except for three decisions
 * that look at all edges / vertices at once: "nothing crossed -> skip the step"
 * (subpoly.py:110-111), the failover override (subpoly_debug.py:41-49), and vertex
 * survival after pruning (subpoly.py:268-272: a vertex on a shared plane has edges on
 * both sides).  Every hyperplane step and the surface-skeleton extraction therefore end in
 * one exchange, written by the step's own kernels straight into the peers' mailboxes
 * (device memory mapped with CUDA IPC over NVLink; plain pointers when the slabs of one
 * object run on one device).  Planar path only.
 *
 * Skeleton of one slab, in two halves: the reference's distance threshold is per chunk
 * (tropical.py:189-197), so the caller reduces max_grad (MAX over ranks) in between. */
int tnb_skeleton_sweep(const tnb_net *net, int32_t unit, int32_t x_lo, int32_t x_hi,
                       int32_t shared_lower, int32_t shared_upper, tnb_sweep **out, void *stream);
/* Plane sharding of the sweep only (exact by construction: the marks-grid planes are split over the GPUs,
 * every rank evaluates |sdf| and the per-chunk max |grad| of its planes, the planes travel by ONE all-gather
 * and the maxima by one MAX all-reduce, then every rank holds the sweep of the whole grid):
 * tnb_skeleton_sweep_alloc = the layout of the whole grid's sweep with nothing evaluated;
 * tnb_sweep_read_dist copies a sweep's planes out ([planes][M][M] floats, plane x_lo first),
 * tnb_sweep_write_dist copies the planes [x_lo, x_hi] in. */
int tnb_skeleton_sweep_alloc(const tnb_net *net, int32_t unit, tnb_sweep **out, void *stream);
int64_t tnb_sweep_num_planes(const tnb_sweep *sw);
int tnb_sweep_read_dist(const tnb_sweep *sw, float *d_out, void *stream);
int tnb_sweep_write_dist(tnb_sweep *sw, const float *d_in, int32_t x_lo, int32_t x_hi, void *stream);
void tnb_sweep_destroy(tnb_sweep *sw);
int32_t tnb_sweep_num_chunks(const tnb_sweep *sw);
int tnb_sweep_read_max_grad(const tnb_sweep *sw, float *d_out, void *stream);   /* [n_chunks] */
int tnb_sweep_write_max_grad(tnb_sweep *sw, const float *d_in, void *stream);
int tnb_skeleton_finish(const tnb_net *net, tnb_sweep *sw, tnb_complex **out, void *stream);
/* Mailboxes: `payload` = capacity in shared-plane vertices of one neighbour message. */
int64_t tnb_mailbox_bytes(int64_t payload);
int tnb_mailbox_create(int64_t payload, void **out);
int tnb_mailbox_destroy(void *box);
int tnb_mailbox_export(void *box, void *handle64);           /* cudaIpcMemHandle_t, 64 bytes */
int tnb_mailbox_import(const void *handle64, void **out);    /* in another process           */
int tnb_mailbox_release(void *imported);
/* boxes[world]: the mailbox of every rank as seen from this process (own one included).
 * timeout_ms bounds the device-side wait for a peer (<= 0: keep the default, 2 s).  seq0:
 * exchange numbers of this complex start above seq0; all ranks pass the same value and a
 * larger one for every new extraction that reuses the mailboxes. */
int tnb_complex_set_halo(tnb_complex *c, int32_t rank, int32_t world, void *const *boxes,
                         int64_t payload, int32_t timeout_ms, uint32_t seq0);
/* tnb_subpoly_step in two halves around the exchange: part 1 = up to and including the send,
 * part 2 = from the receive on, part 0 = both (one slab per device).  Several slabs on ONE
 * device: part 1 for every slab, then part 2 for every slab, on one stream. */
int tnb_subpoly_step_part(const tnb_net *net, tnb_complex *c, int32_t l, int32_t h, float eps,
                          int32_t force, int32_t part, void *stream);
/* tnb_extract_mesh in two halves around the exchange of surface-vertex liveness. */
int tnb_extract_mesh_begin(const tnb_net *net, tnb_complex *c, float eps, tnb_mesh **out, void *stream);
int tnb_extract_mesh_finish(const tnb_net *net, tnb_complex *c, tnb_mesh *m, void *stream);
/* per mesh vertex: bit0 / bit1 = lies on the plane shared with the lower / upper neighbour
 * (what the merge de-duplicates) */
int tnb_mesh_read_tags(const tnb_mesh *m, uint8_t *d_tags, void *stream);
/* Exactness indicator of a slab run: vertices of the final complex that lie within eps of a
 * shared plane (so the reference sees them from both sides) but exist on one slab only.
 * 0 = the slab mesh equals the single-GPU mesh; otherwise the faces touching those vertices
 * from the other side can differ. */
int64_t tnb_mesh_near_plane(const tnb_mesh *m);

/* ---- hash-grid encoding under autograd (the training loop) -------------------------
 * What the reference gets from tiny-cuda-nn through tcnn.Encoding when autograd is on
 * (tropical.py:32-47; stanford/train.py:180-201: L1 + eikonal loss, i.e. the input gradient is
 * differentiated once more).  d_table is the CALLER's parameter storage (`enc.module.params`,
 * [sum(level sizes) * 2] floats on the device); gradients w.r.t. the table are ACCUMULATED
 * into d_dtable (same layout; zero it first when that is wanted).  d_x [n,3] grid coordinates. */
typedef struct tnb_grid_desc {
    int32_t n_levels;         /* L                  */
    int32_t log2_hashmap;     /* T                  */
    int32_t base_resolution;  /* N_min              */
    double per_level_scale;   /* b (tropical.py:31) */
} tnb_grid_desc;
/* floats in the table of this layout (-1: bad description) */
int64_t tnb_grid_train_table_len(const tnb_grid_desc *desc);
/* d_enc [n, 2L] = encoding(x) */
int tnb_grid_train_forward(const tnb_grid_desc *desc, const float *d_table, const float *d_x,
                           int64_t n, float *d_enc, void *stream);
/* given d_denc = dLoss/d enc [n,2L]: d_dtable += dLoss/d table, d_dx [n,3] = dLoss/d x
 * (either may be NULL) */
int tnb_grid_train_backward(const tnb_grid_desc *desc, const float *d_table, const float *d_x,
                            int64_t n, const float *d_denc, float *d_dtable, float *d_dx,
                            void *stream);
/* backward of the map (table, x, denc) -> dx above, given d_ddx = dLoss/d(dx) [n,3]:
 * d_dtable += dLoss/d table, d_ddenc [n,2L] = dLoss/d denc, d_dx2 [n,3] = dLoss/d x
 * (any may be NULL) */
int tnb_grid_train_backward_backward(const tnb_grid_desc *desc, const float *d_table,
                                     const float *d_x, int64_t n, const float *d_denc,
                                     const float *d_ddx, float *d_dtable, float *d_ddenc,
                                     float *d_dx2, void *stream);

/* ---- stage-level pieces of the curve-approximation path ------------------------------ */
/* Net.forward(x, gather=True, group=8) (tropical/stanford/model.py:52-76, the masking of :65-70):
 * d_x holds groups of 8 points (the corners of an edge's box, geometry.corner_points); inside a
 * group a hidden neuron stays linear iff it is > eps at the first or the last point, else it is
 * multiplied by 0.  d_out [groups*8][R] = the gathered pre-activation rows (last column o1 - o0),
 * d_raw [groups*8][2] = the last layer's output (may be NULL). */
int tnb_net_outputs_group8(const tnb_net *net, const float *d_x, int64_t groups, float eps,
                           float *d_out, float *d_raw, void *stream);
/* Net.forward(x, gather=True) (tropical/stanford/model.py:52-76): both return values in one pass,
 * d_raw [n][2] = the last layer's output (o0, o1), d_out [n][R] = the gathered pre-activation rows
 * (as tnb_net_outputs; may be NULL). */
int tnb_net_forward(const tnb_net *net, const float *d_x, int64_t n, float *d_out, float *d_raw,
                    void *stream);
/* geometry.intersection_of_two_planes(p, q, plane="xz") (tropical/geometry.py:24-138) with
 * batched_polynomial_roots / _batched_polynomial_roots (:259-300): d_p, d_q [count][8] corner
 * values of the two planes, d_out [count][3] trilinear coordinates (-1 where the reference
 * leaves "no root"; bilinear configurations -1 as with the reference's failover off). */
int tnb_curve_intersections(const float *d_p, const float *d_q, int64_t count, float *d_out,
                            void *stream);
/* subpoly_debug.deal_with_gradient_descent(c, d_new, e, eps, gg, idx, inds, ints, net)
 * (tropical/subpoly_debug.py:121-165) for the `count` edges it selects (its mask `gd`): d_edges
 * [count][2][3] end points, d_ints [count][3] trilinear coordinates (in: the closed form's, out: after the
 * loop), d_plane [count] output column of the earlier plane each edge lies in (inds[:, 1]), idx the
 * current hyperplane's column; d_dnew [count][2] = the distances (d0, d1) the reference stores back into
 * d_new.  All edges take another step of 1e-2 down the gradient of d0^2 + d1^2 while any of them is farther
 * than eps from one of its planes, at most 500 steps: *bodies = steps taken, *within_eps = 0 if the loop
 * ran out (the reference then ends the extraction, subpoly.py:172-174).  Synchronises the stream. */
int tnb_curve_gradient_descent(const tnb_net *net, const float *d_edges, float *d_ints,
                               const int32_t *d_plane, int32_t idx, float eps, int64_t count,
                               float *d_dnew, int32_t *bodies, int32_t *within_eps, void *stream);
/* The ordering of geometry.sort_polygon_vertices_batch(v, n, idx) (tropical/geometry.py:483-516):
 * d_v [B][M][3] padded face rows (norm 0 = padding), d_normals [B][3]; d_order [B][M] = the
 * permutation that sorts every row by angle around its centre (stable, descending score),
 * d_valid_sorted [B][M] = the padding mask in that order. */
int tnb_polygon_order(const float *d_v, const float *d_normals, int64_t B, int32_t M, int32_t base,
                      int64_t *d_order, uint8_t *d_valid_sorted, void *stream);
/* extract_skeleton's third return value (tropical/subpoly.py:556-581): for every mesh vertex its
 * row in the vertex arrays of the complex the mesh was extracted from (for a complex built by
 * tnb_complex_from_arrays: the caller's numbering). */
int tnb_mesh_read_vertex_index(const tnb_mesh *m, int64_t *d_index, void *stream);

/* subpoly(net, ...) (tropical/subpoly.py:23-86) for `count` networks in ONE call: the reference extracts one
 * object per process run (train.py:127); a service that extracts many small objects calls this.  nets[count]
 * in, out[count] meshes (as tnb_subpoly; NULL where an object failed), rcs[count] per-object return codes (may
 * be NULL).  Up to `in_flight` objects (<= 0: 8) are worked on at the same time by host threads of the library,
 * each on a stream of its own that is ordered after `stream` at entry; `stream` is ordered after all of them at
 * return.  Every object takes the single-object path (same kernels, same results); small complexes run their
 * hyperplanes in one thread-block cluster each, so up to nine step loops are resident side by side.
 * Returns the first failure's code (its message names the object) or TNB_OK. */
int tnb_subpoly_batch(const tnb_net *const *nets, int32_t count, float size, float eps, int32_t force,
                      int32_t unit, int32_t in_flight, tnb_mesh **out, int32_t *rcs, void *stream);

/* ---- knobs / introspection ------------------------------------------------------ */
/* work-buffer growth factor for the complex (default 4.0) */
int tnb_set_capacity_factor(double f);
/* tnb_subpoly_steps / tnb_subpoly run the hyperplanes of a complex with at most `items`
 * vertices + edges as one thread-block-cluster launch (default 200000, env
 * TNB_CLUSTER_MAX_ITEMS; 0 = never).  Returns the previous value; items < 0 only queries. */
int64_t tnb_set_cluster_max_items(int64_t items);
/* tnb_subpoly_steps / tnb_subpoly run the hyperplanes of a complex with at most `items` vertices + edges
 * inside the persistent step kernel (default 700000, env TNB_FUSED_MAX_ITEMS); a larger complex runs them
 * as a device-driven stream of full-size launches (no host synchronisation between hyperplanes, planar
 * path).  0 = always the stream.  Returns the previous value; items < 0 only queries. */
int64_t tnb_set_fused_max_items(int64_t items);
/* number of CUDA kernels this library launched on the calling thread since the last
 * reset (bench.py's gpu_launches) */
int64_t tnb_launch_count(void);
void tnb_launch_count_reset(void);
/* Work buffers released by an extraction are kept by the calling thread for its next extraction on the same
 * stream (no allocator call in the steady state); this hands them back to the device's memory pool. */
void tnb_release_cached_blocks(void);
/* Per-kernel timers: when enabled, the library brackets its heavy kernels with CUDA
 * events on the launching stream.  Classes: 0 = marks-grid sweep (sdf + gradient),
 * 1 = vertex network rows (outputs + packed signs), 2 = new-vertex subdivision kernel,
 * 3 = connecting-edge search, 4 = face rows, 5 = dense sign sweep.
 * On small complexes all hyperplanes run inside ONE persistent cooperative kernel (class 6).
 * A slab-sharded complex runs each step as two cooperative kernels around the exchange: class 2
 * then times the front half (split scan + new vertices + hit scan + buckets + partner count) and
 * class 3 the back half (connecting-edge write + pruning compaction).
 * tnb_profile_read synchronises the recorded events and returns the summed milliseconds,
 * the launch count, the units (vertices / points / edges / candidates) processed and the
 * ALGORITHMIC bytes of those launches (compulsory HBM traffic). */
#define TNB_PROF_SWEEP 0
#define TNB_PROF_VERTEX_ROWS 1
#define TNB_PROF_NEW_VERTICES 2
#define TNB_PROF_PAIRS 3
#define TNB_PROF_FACE_ROWS 4
#define TNB_PROF_SIGN_SWEEP 5
#define TNB_PROF_STEPS 6       /* persistent step kernel: every hyperplane of a small complex in one launch */
#define TNB_PROF_CLASSES 7
int tnb_profile_enable(int on);
int tnb_profile_read(int cls, double *ms, int64_t *launches, int64_t *units, int64_t *bytes);
void tnb_profile_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* TROPICAL_B200_H */

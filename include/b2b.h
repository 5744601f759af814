/*
 * b2b.h -- C-ABI of the B200-native batched rigid-body step (libb2b.so).
 *
 * The reference (pypybox2d r331) has NO native boundary for this path: World.step is a chain of
 * Python method calls (pypybox2d/world.py:307-354).  The drop-in surface is therefore the Python
 * class API (pypybox2d_b200.World/Body/Fixture) and this header is the boundary underneath it --
 * what a maintainer of the reference would bind with ctypes to replace the body of World.step
 * (binding stub: INTEGRATION.md).  Each entry point names the reference code it replaces.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a negative
 * b2b_error on failure, with a message available from b2b_last_error().  All device arrays are
 * owned by the caller (torch tensors in the Python host layer); the library borrows them between
 * b2b_bind and b2b_destroy and never allocates inside b2b_step.  One context is bound to one CUDA
 * device and is not thread-safe; calls are asynchronous on the caller's stream unless stated.
 *
 * Array layout: world-major structure-of-arrays, `[n_worlds][cap][k]`; the field list, element
 * types and flag bits are documented once in pypybox2d_b200/schema.py (FIELDS) and DESIGN.md.
 * `real` is float (real_bytes = 4) or double (real_bytes = 8) per context.
 */
#ifndef B2B_H
#define B2B_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum b2b_error {
    B2B_OK = 0,
    B2B_ERR_ARG = -1,        /* bad argument / capacity */
    B2B_ERR_CUDA = -2,       /* CUDA runtime error (message in b2b_last_error) */
    B2B_ERR_UNBOUND = -3,    /* b2b_step before b2b_bind */
    B2B_ERR_OVERFLOW = -4    /* a world overflowed a capacity (see b2b_status) */
} b2b_error;

#define B2B_STEP_PROFILE 1u   /* b2b_step flags: accumulate per-phase cycle counters (see b2b_profile) */
#define B2B_STEP_TIME 2u      /* b2b_step flags: bracket every kernel with CUDA events (see b2b_kernel_ms) */
#define B2B_STEP_NO_FLAT 16u  /* b2b_step flags: run the narrowphase inside the per-world plan kernel (debugging) */
#define B2B_STEP_GROUP_KERNEL 32u /* b2b_step flags: run the group-capable position kernel even if no world has a joint island schedule (testing) */
#define B2B_STEP_CONTINUOUS 128u /* b2b_step_host flags: some world of the batch has continuous collision on (w_flags WORLD_CONTINUOUS), so body
                                   * state is final only after the last kernel of the step: read it back after World.solve_toi */
/* bits 8..15: reserved */

typedef struct b2b_caps {
    int32_t n_worlds;        /* worlds in this context (this GPU's shard) */
    int32_t nb;              /* bodies per world (capacity) */
    int32_t nf;              /* fixtures == broadphase proxies per world */
    int32_t nc;              /* contacts per world */
    int32_t npair;           /* pairs one BroadPhase.update_pairs may emit per world */
    int32_t nl;              /* island body-list slots per world (>= nb + static re-visits) */
    int32_t n_shapes;        /* rows of the shared shape table */
    int32_t real_bytes;      /* 4: fp32 kernels, 8: fp64 kernels (parity build, no FMA contraction) */
    int32_t device;          /* CUDA device ordinal */
    int32_t threads;         /* lanes of a warp that step one world in the solve kernels (0: default 4) */
    int32_t ctas_per_sm;     /* resident CTAs per SM the persistent grid is sized for (0: from occupancy) */
    int32_t reserved;        /* lanes per world in the plan and pair kernels (0: default, 8 or 32 by contact capacity) */
    int32_t nj;              /* joints per world (World.joints); 0: the batch has no joints -- the step then runs the kernels without
                              * the joint hooks, and the j_* arrays keep one unused row per world */
    int32_t joint_types;     /* bit t set: some joint of the batch has type t (1 revolute, 2 distance, 3 weld, 4 prismatic, 5 rope, 6 friction,
                              * 7 wheel, 8 pulley, 9 mouse); 0: unknown.  A batch whose joints are all revolute joints (1 << 1) runs kernels without the
                              * type dispatch.  The caller keeps j_def consistent with it */
} b2b_caps;

/* Device pointers, one per schema field, same names and order as pypybox2d_b200/schema.py FIELDS.
 * Replaces the attribute storage of Body (body.py:65-103), Sweep (common.py:870-889), Fixture /
 * FixtureProxy (fixture.py:31-69), Manifold (contact_util.py:50-166), Contact (contact.py:57-81). */
typedef struct b2b_buffers {
    void *w_count, *w_flags, *w_status, *w_gravity, *w_island;
    void *b_xf, *b_pos, *b_vel, *b_pos0, *b_force, *b_mass, *b_damp, *b_flags, *b_island;
    void *f_body, *f_shape, *f_mat, *f_filter, *f_fat, *f_tight, *f_moved;
    void *c_fix, *c_flags, *c_mat, *c_mtype, *c_mf, *c_pts, *c_imp, *c_id;
    void *p_pairs, *l_body, *l_contact, *l_start;
    void *s_meta, *s_radius, *s_data;
    /* joints of the island solver: replaces the attribute storage of RevoluteJoint / DistanceJoint / WeldJoint
     * (pypybox2d/joints/revolute.py:58-93, distance.py:57-93, weld.py:56-81), World.joints and Body.joints (world.py:268-273) */
    void *w_joint, *j_def, *j_par, *j_imp, *j_tmp, *j_adj, *b_jadj, *l_joint, *l_jstart;
} b2b_buffers;

typedef struct b2b_ctx b2b_ctx;

/* Create / destroy a context.  Allocates only the per-CTA scratch (b2b_scratch_bytes). */
int b2b_create(const b2b_caps* caps, b2b_ctx** out);
void b2b_destroy(b2b_ctx* ctx);

/* Bytes of device scratch a context with these capacities allocates (0 on bad caps). */
size_t b2b_scratch_bytes(const b2b_caps* caps);

/* Borrow the caller's device arrays. */
int b2b_bind(b2b_ctx* ctx, const b2b_buffers* buffers);

/* Fixture._create_proxies + DynamicTree.create_proxy for every fixture of every world
 * (fixture.py:250-261, dynamictree.py:115-134): tight AABB from the body transform, fat AABB =
 * tight +- AABB_EXTENSION, proxy put in the move buffer, world flagged `_new_fixture`. */
int b2b_init_proxies(b2b_ctx* ctx, void* stream);

/* World.step(dt, vel_iters, pos_iters), n_steps times, for every world of the batch
 * (world.py:307-354): find_new_contacts if new fixtures,
 * ContactManager.collide (contact_manager.py:174-223), World._solve (world.py:475-592: island DFS,
 * Island.solve island.py:65-179, ContactSolver contact.py:375-864, fixture synchronisation,
 * find_new_contacts contact_manager.py:102-158), and for the worlds whose w_flags carry WORLD_CONTINUOUS
 * (World(continuous_physics=True), the reference's default) World.solve_toi (world.py:594-821: time_of_impact toi.py:223-391
 * on shape_distance distance.py:419-561, Island.solve_toi island.py:181-243, solve_toi_position_constraints
 * contact.py:866-933), clear forces.  Batches created with a joint capacity (b2b_caps.nj > 0) also walk Body.joints in the island
 * DFS (world.py:548-572) and run Joint._init_velocity_constraints / _solve_velocity_constraints / _solve_position_constraints of
 * their revolute, distance, weld, prismatic, rope, friction, wheel, pulley and mouse joints inside Island.solve (island.py:107-113,
 * 145-148; joints/revolute.py:208-453, distance.py:155-311, weld.py:106-257, prismatic.py:232-513, rope.py:106-248,
 * friction.py:105-233, wheel.py:202-416, pulley.py:156-318, mouse.py:138-229), and Body._should_collide's joint test
 * (body.py:920-926).
 * inv_dt0 is World._inv_dt0 before the first of the n_steps (0 for a fresh world); the first step
 * uses dt_ratio = inv_dt0 * dt, later ones (1/dt) * dt, both evaluated in double as world.py:323-328.
 * Asynchronous on `stream` (a cudaStream_t, NULL = default stream). */
int b2b_step(b2b_ctx* ctx, double dt, double inv_dt0, int vel_iters, int pos_iters, int n_steps,
             uint32_t flags, void* stream);

/* Same step driven from HOST buffers (the end-to-end call an environment loop makes):
 * h_force_torque [n_worlds][nb][3] (fx, fy, torque; may be NULL) goes host->device and is applied as Body.apply_force /
 * apply_torque before the step would (body.py:477-545: dynamic bodies only; a sleeping body that receives a non-zero
 * force or torque is woken and its sleep timer restarts); body state [n_worlds][nb][6] (c.x, c.y, a, v.x, v.y, w) of the
 * last step comes back into h_state_out (may be NULL).  Copies and kernels are pipelined over chunks of worlds on two
 * internal copy streams and the caller's stream; returns after everything has completed.  Host buffers must be
 * pinned for the copies to overlap (and for full PCIe bandwidth). */
int b2b_step_host(b2b_ctx* ctx, const void* h_force_torque, void* h_state_out, double dt, double inv_dt0,
                  int vel_iters, int pos_iters, int n_steps, uint32_t flags, void* stream);

/* Test hook (T1): run only the narrowphase dispatch (collision.py collide_*) on n explicit pairs.
 * d_shape_a/b: int32 shape rows; d_xf_a/b: real[n][4] (x, y, cos, sin); outputs as c_mtype/c_mf/c_pts/c_id. */
int b2b_collide_pairs(b2b_ctx* ctx, int n, const void* d_shape_a, const void* d_shape_b, const void* d_xf_a,
                      const void* d_xf_b, void* d_mtype, void* d_mf, void* d_pts, void* d_id, void* stream);

/* Test hook (T1): only the solver stage of World.step -- World._solve's island construction (world.py:475-580) and
 * Island.solve for every island (island.py:65-179, ContactSolver contact.py:375-864) on the contacts as they are: no
 * narrowphase before it (run b2b_step with dt = 0 for that, world.py:338-349) and no fixture synchronisation or
 * find_new_contacts after it.  inv_dt0 as for b2b_step; dt > 0.  Asynchronous on `stream`. */
int b2b_solve_islands(b2b_ctx* ctx, double dt, double inv_dt0, int vel_iters, int pos_iters, uint32_t flags, void* stream);

/* Queries, the callers either side of the step.  All arrays are device arrays of the caller; one warp per
 * query scans the world's proxies (the reference's tree only prunes: DynamicTree.query / ray_cast, dynamictree.py:190-262, report
 * a proxy iff its fat box passes the tests restated here), so results come in ascending fixture order where the reference
 * reports them in tree order -- the same SET.  Asynchronous on `stream`.
 *
 * World.query_aabb (world.py:371-385): for query q, world d_world[q] (int32) and box d_box[q] = (lx, ly, ux, uy) (real):
 * d_count[q] = number of fixtures whose fat AABB overlaps the box, the first max_hits of them in d_fix[q][max_hits] (int32). */
int b2b_query_aabb(b2b_ctx* ctx, int n_queries, const void* d_world, const void* d_box, int max_hits, void* d_count, void* d_fix,
                   void* stream);

/* World.ray_cast (world.py:387-440) of the segment d_seg[q] = (x1, y1, x2, y2) in world d_world[q], shape by shape as
 * Shape.ray_cast (shapes.py:147-176 circle, 266-314 polygon, 655-700 edge: the port returns an edge's normal in the edge's
 * local frame).  mode 0: every fixture hit (a callback that never clips), d_count[q] of them, the first max_hits in
 * d_fix[q][.] with d_hit[q][.][5] = {fraction, normal.x, normal.y, point.x, point.y} (real); mode 1: the closest hit only
 * (a callback that clips to each reported fraction; ties: the lowest fixture index) in slot 0, d_count[q] = 0 or 1.
 * A zero-length segment (the reference raises ValueError, dynamictree.py:222-223) gets d_count[q] = -1. */
int b2b_ray_cast(b2b_ctx* ctx, int n_rays, const void* d_world, const void* d_seg, int mode, int max_hits, void* d_count,
                 void* d_fix, void* d_hit, void* stream);

/* Fixture.test_point (fixture.py:228-233, shapes.py:141-145, 256-263) for n triples (d_world[i], d_fixture[i], d_point[i][2]):
 * d_inside[i] (uint8) = 1 when the point is inside the fixture's shape (an edge contains no point). */
int b2b_test_points(b2b_ctx* ctx, int n, const void* d_world, const void* d_fixture, const void* d_point, void* d_inside,
                    void* stream);

/* OR of every world's status word (STATUS_* bits in schema.py); synchronises the stream. */
int b2b_status(b2b_ctx* ctx, uint32_t* host_out, void* stream);

/* Per-world state checksum and energy for the optional end-of-run gather:
 * d_out real64[n_worlds][4] = {low 52 bits of the FNV-1a hash of the packed body state, as an exactly representable
 * integer (comparable with ==, never NaN), kinetic+potential energy, n_contacts, n_awake}.  masses are passed as
 * inverse mass/inertia already in b_mass. */
int b2b_checksum(b2b_ctx* ctx, void* d_out, void* stream);

/* Per-phase cycle counters of the step kernel, summed over CTAs (only accumulated by steps run with
 * flags & B2B_STEP_PROFILE): out[32] int64, slots named in pypybox2d_b200/csrc/b2b_step.cuh (PROF_*).
 * reset != 0 clears them afterwards.  Synchronises the device. */
int b2b_profile(b2b_ctx* ctx, int64_t* host_out32, int reset);

/* Device time of the kernels launched by b2b_step calls that had B2B_STEP_TIME set, from CUDA events recorded on
 * the launching stream between the kernels: out[8] = total milliseconds of {collide, plan, prepare, solve_velocity,
 * solve_position (both instantiations), sync, broadphase (pairs + new contacts), all of them}; n_steps = steps covered.
 * Synchronises the stream's events; clears the accumulators. */
int b2b_kernel_ms(b2b_ctx* ctx, float* out8, int32_t* n_steps);

/* Launch geometry actually used and kernel launches issued so far (for bench.py's gpu_launches). */
int b2b_info(b2b_ctx* ctx, int32_t* grid, int32_t* threads, int32_t* smem_bytes, int64_t* launches);

const char* b2b_last_error(void);
const char* b2b_field_names(void);   /* comma-separated b2b_buffers member names */
const char* b2b_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B2B_H */

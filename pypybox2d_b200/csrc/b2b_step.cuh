// b2b_step.cuh -- one World.step for one world, written as a "team program".
//
// A world is stepped by a TEAM of threads (a group of lanes of one warp; several worlds per warp).  The step is a
// sequence of phases separated by team barriers; inside a phase, work items (contacts, bodies, proxies, pairs,
// constraints of one dependency level) are strided over the team's lanes.  The kernels (b2b_kernels.inl) call
//   collide_one                      from the flat narrowphase kernel, one thread per contact of the whole batch
//   run_plan / run_solve / run_broadphase   from the three persistent kernels, one team per world
//
// Reference call tree being replaced:
//   World.step                      pypybox2d/world.py:307-354
//   ContactManager.collide          contact_manager.py:174-223   -> collide_one, phase_collide   (K1)
//   Contact.update                  contact.py:210-268           -> contact_update
//   World._solve (island DFS)       world.py:475-580             -> phase_islands, phase_schedule (K2)
//   Island.solve / ContactSolver    island.py:65-179, contact.py:375-864 -> solve_islands         (K3)
//   Body._synchronize_fixtures ...  body.py:898-904, fixture.py:287-298, dynamictree.py:152-188 -> phase_sync (K4)
//   BroadPhase.update_pairs         broadphase.py:83-109         -> phase_pairs          (K5)
//   ContactManager.find_new_contacts contact_manager.py:102-158  -> phase_pairs          (K6)
//
// The parallel-loop macros have a sequential expansion when compiled without CUDA so that
// tests/hostcheck can run this exact source under g++ for debugging; that build is test-only.
#pragma once
#include <cstring>
#include <type_traits>
#include "b2b_collide.cuh"
#include "b2b_toi.cuh"

// A world is stepped by a TEAM of threads: the whole CTA in the plan and broadphase kernels, a group of
// `t_n` lanes of one warp in the solve kernel (several worlds per warp).  Work items are strided over the team
// and phases end in a team barrier.  Without CUDA the team is one sequential thread (tests/hostcheck).
#if defined(__CUDA_ARCH__)
#define B2B_PFOR(i, n) for (int i = t_id; i < (n); i += t_n)
#define B2B_SYNC()                               \
    do {                                         \
        if (t_cta) __syncthreads();              \
        else __syncwarp(t_mask);                 \
    } while (0)
#define B2B_IS_T0 (t_id == 0)
#define B2B_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define B2B_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define B2B_ATOMIC_OR(p, v) atomicOr((p), (v))
#define B2B_ATOMIC_AND(p, v) atomicAnd((p), (v))
#define B2B_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define B2B_CLOCK() clock64()
#define B2B_POPC(m) __popc(m)
#define B2B_CTZ(m) (__ffs((int)(m)) - 1)
#define B2B_PREFETCH(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#else
#define B2B_PFOR(i, n) for (int i = 0; i < (n); ++i)
#define B2B_SYNC() ((void)0)
#define B2B_IS_T0 true
#define B2B_CLOCK() 0LL
#define B2B_POPC(m) __builtin_popcount(m)
#define B2B_CTZ(m) __builtin_ctz(m)
#define B2B_PREFETCH(p) ((void)0)
namespace b2b {
template <class T> inline T host_atomic_min(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T host_atomic_add(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T host_atomic_or(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T host_atomic_and(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T host_atomic_max(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
}
#define B2B_ATOMIC_MIN(p, v) b2b::host_atomic_min((p), (v))
#define B2B_ATOMIC_ADD(p, v) b2b::host_atomic_add((p), (v))
#define B2B_ATOMIC_OR(p, v) b2b::host_atomic_or((p), (v))
#define B2B_ATOMIC_AND(p, v) b2b::host_atomic_and((p), (v))
#define B2B_ATOMIC_MAX(p, v) b2b::host_atomic_max((p), (v))
#endif

// cycle accounting by thread 0 (wall-clock cycles of the CTA, since phases end in CTA barriers)
#define B2B_PROF(slot, t)                                                      \
    do {                                                                       \
        if (profiling) {                                                       \
            const long long t1_ = B2B_CLOCK();                                 \
            if (B2B_IS_T0) prof[slot] += t1_ - (t);                            \
            (t) = t1_;                                                         \
        }                                                                      \
    } while (0)


namespace b2b {

struct caps_t {
    int n_worlds, nb, nf, nc, npair, nl, n_shapes;
    int nj = 0;              // joints per world (0: the batch has none and runs the kernels without the joint hooks); the J arrays have max(nj, 1) rows
    int joint_types = 0;     // bit t set: some joint of the batch has type t (b2b_joints.inl); 0: unknown (any type)
};

// Pointers to the batch arrays (world-major).  Field order == schema.FIELDS == b2b_buffers.
template <class real> struct batch_ptrs {
    int32_t* w_count; uint32_t* w_flags; uint32_t* w_status; real* w_gravity; int32_t* w_island;
    real *b_xf, *b_pos, *b_vel, *b_pos0, *b_force, *b_mass, *b_damp; uint32_t* b_flags; int32_t* b_island;
    int32_t *f_body, *f_shape; real* f_mat; uint32_t* f_filter; real *f_fat, *f_tight; uint8_t* f_moved;
    int32_t* c_fix; uint32_t* c_flags; real* c_mat; int32_t* c_mtype; real *c_mf, *c_pts, *c_imp; uint32_t* c_id;
    int32_t *p_pairs, *l_body, *l_contact, *l_start;
    int32_t* s_meta; real* s_radius; real* s_data;
    int32_t *w_joint, *j_def; real *j_par, *j_imp, *j_tmp; int32_t *j_adj, *b_jadj, *l_joint, *l_jstart;   // joints (b2b_joints.inl)
};

// number of real-valued SoA rows per constraint slot (see cc_* accessors below)
enum { VREC_QUADS = 8, PREC_QUADS = 6 };   // velocity / position record of one constraint, in 4-vectors of reals ("quads")
// per-CTA cycle counters (step_params.flags & STEP_PROFILE): see PROF_* below
enum { PROF_SLOTS = 32, STEP_PROFILE = 1, STEP_FLAT_COLLIDE = 4, STEP_GROUP_KERNEL = 32, STEP_SOLVE_ONLY = 64,
       STEP_FIRST = 256 };      // (set by the launcher) the first of a call's n_steps: joints scale their impulses by dt_ratio_first
// per-world summary word: what the flat kernels found, so that the per-world kernels can skip a settled world in O(1)
enum { SUM_PENDING = 1u, SUM_DESTROY = 2u, SUM_CHANGED = 4u, SUM_MOVED = 8u,
       SUM_SENSOR = 16u,          // an ACTIVE sensor contact was left to the plan kernel: its pending contacts have to be walked
       SUM_COLLIDE = 23u };
enum { PROF_PAIRS0 = 0, PROF_COLLIDE, PROF_ADJACENCY, PROF_ISLANDS, PROF_SOLVE, PROF_SYNC, PROF_PAIRS, PROF_CLEAR,
       PROF_S_STAGE = 8, PROF_S_INTEGRATE, PROF_S_SCHEDULE, PROF_S_INIT, PROF_S_WARM, PROF_S_VELOCITY, PROF_S_STORE,
       PROF_S_POSITION, PROF_S_WRITEBACK, PROF_S_SLEEP, PROF_WORLD_STEPS = 31 };

// ---- memory plan --------------------------------------------------------------------------------
// (1) caller-owned batch arrays (b2b_buffers);  (2) context-owned per-WORLD plan arrays written by the plan
// kernel and read by the solve / broadphase kernels;  (3) context-owned per-CTA scratch (persistent grids reuse
// it for every world they own, so it stays L1/L2 resident);  (4) shared memory, one layout per kernel.
struct plan_layout {          // per world, in uint16_t elements
    size_t off_cbody, off_sched, off_slot_isl, off_joint, off_lv, off_adj, off_adj_start, off_cstate, off_summary, stride;
};
B2B_HD size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
// Level-step table of the solve kernel: one 4-bit count per level-step, eight per word.  Island `isl`, whose
// constraint slots start at c0, owns the words from (c0 >> 3) + 2 * isl on: n slots make at most n level-steps =
// ceil(n / 8) words plus a closing zero word (the reader runs one word ahead), and the next island starts at
// least floor(n / 8) + 2 words later, so islands never share a word and can be scheduled concurrently.
B2B_HD int lv_words(const caps_t& c) { return (c.nc + 7) / 8 + 2 * c.nb + 4; }
B2B_HD int lv_offset(int c0, int isl) { return (c0 >> 3) + 2 * isl; }
enum { LV_MAX_LANES = 8 };      // a level-step holds at most min(solve lanes, 8) slots (wider levels are split)
B2B_HD plan_layout plan_plan(const caps_t& c)
{
    plan_layout P;
    size_t o = 0;
    P.off_cbody = o; o = align_up(o + 2 * (size_t)c.nc, 8);
    P.off_sched = o; o = align_up(o + (size_t)c.nc + 2, 8);
    P.off_slot_isl = o; o = align_up(o + (size_t)c.nc + 2, 8);               // joint schedule: contact island of every slot
    P.off_joint = o; o = align_up(o + 4, 8);                                 // [0] joint schedule in use, [1] contact islands
    P.off_lv = o; o = align_up(o + 2 * (size_t)lv_words(c), 8);             // uint32 words, counted in uint16 units
    P.off_adj = o; o = align_up(o + 2 * (size_t)c.nc + 2, 8);
    P.off_adj_start = o; o = align_up(o + (size_t)c.nb + 2, 8);
    P.off_cstate = o; o = align_up(o + ((size_t)c.nc + 1) / 2 + 1, 8);      // one byte per contact
    P.off_summary = o; o = align_up(o + 2, 8);                               // one uint32 (SUM_* bits)
    P.stride = o;
    return P;
}

// continuous collision (phase_toi): per-contact time of impact + state word, a batch of sweep snapshots, the team's
// reduction words, and the sub-island World.solve_toi solves (at most 2 * MAX_TOI_CONTACTS bodies, MAX_TOI_CONTACTS contacts)
enum { TOI_BATCH = 32, TOI_SNAP_REALS = 13, TOI_CC_REALS = 56 };
struct toi_layout { size_t off_alpha, off_state, off_list, off_snap, off_snap_ci, off_red_a, off_red_i, off_ib, off_ic, off_P, off_V, off_cc, total; };
B2B_HD toi_layout toi_plan(const caps_t& c, int real_bytes)
{
    toi_layout T;
    size_t o = 0;
    T.off_alpha = o; o = align_up(o + (size_t)real_bytes * ((size_t)c.nc + 2), 32);
    T.off_state = o; o = align_up(o + sizeof(int32_t) * ((size_t)c.nc + 2), 32);
    T.off_list = o; o = align_up(o + sizeof(int32_t) * 3 * ((size_t)c.nc + 2), 32);     // compacted index lists: candidates, contacts of seed a / b
    T.off_snap = o; o = align_up(o + (size_t)real_bytes * TOI_SNAP_REALS * TOI_BATCH, 32);
    T.off_snap_ci = o; o = align_up(o + sizeof(int32_t) * TOI_BATCH, 32);
    T.off_red_a = o; o = align_up(o + (size_t)real_bytes * 32, 32);
    T.off_red_i = o; o = align_up(o + sizeof(int32_t) * 32, 32);
    T.off_ib = o; o = align_up(o + sizeof(int32_t) * (2 * MAX_TOI_CONTACTS + 2), 32);
    T.off_ic = o; o = align_up(o + sizeof(int32_t) * (MAX_TOI_CONTACTS + 2), 32);
    T.off_P = o; o = align_up(o + (size_t)real_bytes * 3 * (2 * MAX_TOI_CONTACTS + 2), 32);
    T.off_V = o; o = align_up(o + (size_t)real_bytes * 3 * (2 * MAX_TOI_CONTACTS + 2), 32);
    T.off_cc = o; o = align_up(o + (size_t)real_bytes * TOI_CC_REALS * (MAX_TOI_CONTACTS + 1), 32);
    T.total = o;
    return T;
}

struct scratch_layout {       // bytes; `aux` slabs are per CTA of the plan/broadphase kernels, `rec` slabs per world of a solve group
    size_t off_wake, off_scan, off_tmp_contact, off_pair_keep, off_prof, off_toi, total_aux;
    size_t off_cc_vel, off_cc_pos, off_cc_imp, off_cc_state, total_rec;
    size_t total_sprof;       // per solve team: cycle counters
};
inline scratch_layout scratch_plan(const caps_t& c, int real_bytes)
{
    scratch_layout L;
    size_t o = 0;
    const size_t nmax = (size_t)((c.nc > c.nf ? c.nc : c.nf) > c.npair ? (c.nc > c.nf ? c.nc : c.nf) : c.npair) + 2;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 128); return at; };
    L.off_wake = take(sizeof(int32_t) * (size_t)c.nb);
    L.off_scan = take(sizeof(int32_t) * (nmax > (size_t)c.nb + 2 ? nmax : (size_t)c.nb + 2));
    // staging copy of the contact rows for order-preserving compaction: 2+1+1+2 words + 14 reals
    L.off_tmp_contact = take((sizeof(int32_t) * 6 + (size_t)real_bytes * 14) * ((size_t)c.nc + 2));
    L.off_pair_keep = take(sizeof(int32_t) * ((size_t)c.npair + 2));
    L.off_prof = take(sizeof(long long) * PROF_SLOTS);
    L.off_toi = take(toi_plan(c, real_bytes).total);
    L.total_aux = align_up(o, 256);
    o = 0;
    L.off_cc_vel = take((size_t)real_bytes * 4 * VREC_QUADS * ((size_t)c.nc + 2));
    L.off_cc_pos = take((size_t)real_bytes * 4 * PREC_QUADS * ((size_t)c.nc + 2));
    L.off_cc_imp = take((size_t)real_bytes * 4 * ((size_t)c.nc + 2));
    L.off_cc_state = take((size_t)real_bytes * 3 * (size_t)c.nb);      // the velocities the solver starts from, per body
    L.total_rec = align_up(o, 256);
    L.total_sprof = align_up(sizeof(long long) * PROF_SLOTS, 256);
    return L;
}

// shared memory.  plan kernel (per CTA): [ctl | flags | ls | cbody | adj | adj_start | stack/last | lev | cflag];
// solve kernel (per team): [three reals per body]; broadphase kernel (per CTA): [ctl].
struct smem_layout {
    size_t off_fat, off_moved;                                // (pair finding; part of the broadphase prefix)
    size_t off_ctl, off_flags, off_ls, off_cbody, off_adj, off_adj_start, off_stack, off_lev, off_cflag, total_plan;
    size_t off_state, total_solve;                            // relative to the team's base
    size_t total_broad;
    size_t total_broad_nobits;                                // the pair-finding layout without the candidate bitmap (TOI kernel)
    size_t off_pairbits;                                      // broadphase kernel: candidate bitmap [nf][ceil(nf / 32)] (0: none, too big)
};
B2B_HD smem_layout smem_plan_layout(const caps_t& c, int real_bytes)
{
    smem_layout S;
    size_t o = 0;
    S.off_ctl = o; o = align_up(o + sizeof(int32_t) * 16, 32);
    // pair finding stages every proxy's fat box and move flag: right after the control words in the broadphase kernel;
    // in the plan kernel (first step only, before the contact graph exists) on top of the graph arrays from off_ls on
    const size_t fat_bytes = align_up((size_t)real_bytes * 4 * (size_t)c.nf, 32), moved_bytes = align_up((size_t)c.nf, 32);
    S.off_fat = o; S.off_moved = o + fat_bytes;
    S.total_broad = o + fat_bytes + moved_bytes;
    S.total_broad_nobits = S.total_broad;
    {   // one bit per (mover, proxy): the overlap tests of pass 1 are not repeated by pass 2 (when it fits 24 KB)
        const size_t bits_bytes = sizeof(uint32_t) * (((size_t)c.nf + 31) / 32) * (size_t)c.nf;
        S.off_pairbits = 0;
        if (bits_bytes <= 24 * 1024) { S.off_pairbits = S.total_broad; S.total_broad += align_up(bits_bytes, 32); }
    }
    S.off_flags = o; o = align_up(o + sizeof(uint32_t) * (size_t)c.nb, 32);
    S.off_ls = o; o = align_up(o + sizeof(uint16_t) * ((size_t)c.nc + 3 * (size_t)c.nb + 4), 32);
    S.off_cbody = o; o = align_up(o + sizeof(uint16_t) * 2 * (size_t)c.nc, 32);
    S.off_adj = o; o = align_up(o + sizeof(uint16_t) * (2 * (size_t)c.nc + 2), 32);
    S.off_adj_start = o; o = align_up(o + sizeof(uint16_t) * ((size_t)c.nb + 2), 32);
    S.off_stack = o; o = align_up(o + sizeof(uint16_t) * ((size_t)c.nb + 2), 32);
    {   // level of every island slot (phase_schedule); before that the adjacency builder's per-body counters (int32)
        const size_t lev_bytes = sizeof(uint16_t) * ((size_t)c.nc + 2), cnt_bytes = sizeof(int32_t) * ((size_t)c.nb + 2);
        S.off_lev = o; o = align_up(o + (lev_bytes > cnt_bytes ? lev_bytes : cnt_bytes), 32);
    }
    S.off_cflag = o; o = align_up(o + (size_t)c.nc, 32);
    S.total_plan = o;
    if (S.total_plan < S.off_ls + fat_bytes + moved_bytes) S.total_plan = S.off_ls + fat_bytes + moved_bytes;
    o = 0;
    S.off_state = o; o = align_up(o + (size_t)real_bytes * 3 * (size_t)c.nb, 32);
    S.total_solve = o;
    return S;
}

struct step_params {
    double dt, dt_ratio_first, dt_ratio_rest;
    int vel_iters, pos_iters, n_steps;
    uint32_t flags;
    int solve_lanes;             // lanes per world of the solve kernel: the level-step table is cut to this width
    uint32_t zero;               // always 0: a value the compiler cannot fold (orders the record loads, see level_sweep)
    unsigned int* solve_ctr;     // three device words per launch_solve_group call (zeroed by it): the next world of the group for the
                                 // velocity / position / joint position kernel's teams (null: worlds strided over the teams)
    unsigned int* joint_any;     // device word, cleared every step: raised by the prepare kernel when some world of the batch
                                 // has a joint island schedule (the position kernel then runs its group-capable instantiation)
};

// ================================================================================================
template <class real> struct world_step {
    typedef vec2<real> V2;
    typedef xform<real> XF;
    struct alignas(4 * sizeof(real)) quad { real x, y, z, w; };

    caps_t cap;
    // this world's rows of the batch arrays
    int32_t* w_count; uint32_t* w_flags; uint32_t* w_status; real* w_gravity; int32_t* w_island;
    real *b_xf, *b_pos, *b_vel, *b_pos0, *b_force; const real *b_mass, *b_damp; uint32_t* b_flags; int32_t* b_island;
    const int32_t *f_body, *f_shape; const real* f_mat; const uint32_t* f_filter; real *f_fat, *f_tight; uint8_t* f_moved;
    int32_t* c_fix; uint32_t* c_flags; real* c_mat; int32_t* c_mtype; real *c_mf, *c_pts, *c_imp; uint32_t* c_id;
    int32_t *p_pairs, *l_body, *l_contact, *l_start;
    const int32_t* s_meta; const real* s_radius; const real* s_data;
    int32_t *w_joint, *j_def; const real* j_par; real *j_imp, *j_tmp; const int32_t *j_adj, *b_jadj; int32_t *l_joint, *l_jstart;
    // this world's rows of the plan arrays (global)
    uint16_t *g_cbody, *g_sched, *g_slot_isl, *g_joint, *g_adj, *g_adj_start;
    real* s_fat; uint8_t* s_moved;   // pair finding: fat boxes and move flags staged in shared memory
    uint32_t* s_pairbits = nullptr;  // pair finding: candidate bitmap (broadphase kernel, when it fits)
    uint32_t* g_lv;          // level-step table (see lv_words)
    uint32_t* g_summary;     // SUM_* bits raised by the flat kernels, consumed (and cleared) by the per-world kernels
    uint8_t* g_cstate;       // per contact, written by the flat narrowphase kernel: 0 pending, 1 updated, 2 destroy, 3 updated + touching changed
    // scratch (global): per CTA (plan / broadphase kernels) or per team (solve kernel)
    int32_t *wake_at, *scan;
    real *cc_vel, *cc_pos, *cc_imp, *cc_state;
    int32_t* tmp_ci; real* tmp_cr;
    int32_t* pair_keep;
    long long* prof;
    // continuous collision (phase_toi): see toi_layout
    struct toi_cc;
    real *toi_alpha, *toi_snap, *toi_red_a, *toi_P, *toi_V;
    int32_t *toi_state, *toi_list, *toi_snap_ci, *toi_red_i, *toi_ib, *toi_ic;
    toi_cc* toi_C;
    // shared memory
#if defined(B2B_TRACE) && defined(__CUDA_ARCH__)
    // development trace (tools/trace_probe.py): cycles of lane 0 per segment of a level-step, summed over the run
    unsigned tr[7], tlast;
#define B2B_TSTAMP(i, dep) do { asm volatile("" :: "r"(dep) : "memory"); const unsigned now_ = (unsigned)clock(); tr[i] += now_ - tlast; tlast = now_; } while (0)
#define B2B_TSTAMP_F(i, a, b, c, d) do { asm volatile("" :: "f"(a), "f"(b), "f"(c), "f"(d) : "memory"); const unsigned now_ = (unsigned)clock(); tr[i] += now_ - tlast; tlast = now_; } while (0)
#else
#define B2B_TSTAMP(i, dep) ((void)0)
#define B2B_TSTAMP_F(i, a, b, c, d) ((void)0)
#endif
    real* s_third;           // solve kernel: third component of every body's state (after the (x, y) pairs)
    real* s_state;           // solve kernel: per body (v.x, v.y, w) during the velocity phase, then (c.x, c.y, a)
    uint32_t* s_flags;       // plan kernel: body flags
    int32_t* s_ctl;          // [16] control words
    uint16_t *s_ls, *s_cbody, *s_adj, *s_adj_start, *s_stack, *s_last, *s_lev;
    uint8_t* s_cflag;
    // the team of threads stepping this world
    int t_id, t_n;
    unsigned t_mask;
    unsigned w_mask = 0u;    // solve kernels: lanes of the warp whose team has a world this round (0: not bound)
    bool t_cta;

    B2B_D void bind_team_cta()
    {
#if defined(__CUDA_ARCH__)
        t_id = (int)threadIdx.x; t_n = (int)blockDim.x; t_mask = 0xffffffffu; t_cta = true;
#else
        t_id = 0; t_n = 1; t_mask = 1u; t_cta = true;
#endif
    }
    // `lanes` consecutive lanes of the warp (a power of two); returns the team's index inside the CTA
    B2B_D int bind_team_lanes(int lanes)
    {
#if defined(__CUDA_ARCH__)
        const int lane = (int)(threadIdx.x & 31);
        t_n = lanes; t_id = lane % lanes; t_cta = false;
        const int team = lane / lanes;
        t_mask = (lanes == 32) ? 0xffffffffu : (((1u << lanes) - 1u) << (team * lanes));
        return (int)(threadIdx.x >> 5) * (32 / lanes) + team;
#else
        (void)lanes;
        t_id = 0; t_n = 1; t_mask = 1u; t_cta = false;
        return 0;
#endif
    }

    B2B_D void bind_rows(const batch_ptrs<real>& B, const caps_t& c, int w, uint16_t* plan, const plan_layout& P)
    {
        cap = c;
        const size_t W = (size_t)w;
        w_count = B.w_count + 4 * W; w_flags = B.w_flags + W; w_status = B.w_status + W;
        w_gravity = B.w_gravity + 2 * W; w_island = B.w_island + 4 * W;
        const size_t ob = W * c.nb, of = W * c.nf, oc = W * c.nc;
        b_xf = B.b_xf + 4 * ob; b_pos = B.b_pos + 4 * ob; b_vel = B.b_vel + 4 * ob; b_pos0 = B.b_pos0 + 4 * ob;
        b_force = B.b_force + 2 * ob; b_mass = B.b_mass + 4 * ob; b_damp = B.b_damp + 4 * ob;
        b_flags = B.b_flags + ob; b_island = B.b_island + ob;
        f_body = B.f_body + of; f_shape = B.f_shape + of; f_mat = B.f_mat + 2 * of; f_filter = B.f_filter + 2 * of;
        f_fat = B.f_fat + 4 * of; f_tight = B.f_tight + 4 * of; f_moved = B.f_moved + of;
        c_fix = B.c_fix + 2 * oc; c_flags = B.c_flags + oc; c_mat = B.c_mat + 2 * oc; c_mtype = B.c_mtype + oc;
        c_mf = B.c_mf + 4 * oc; c_pts = B.c_pts + 4 * oc; c_imp = B.c_imp + 4 * oc; c_id = B.c_id + 2 * oc;
        p_pairs = B.p_pairs + 2 * W * c.npair; l_body = B.l_body + W * c.nl; l_contact = B.l_contact + oc;
        l_start = B.l_start + 2 * ob;
        s_meta = B.s_meta; s_radius = B.s_radius; s_data = B.s_data;
        const size_t oj = W * (size_t)(c.nj > 0 ? c.nj : 1);
        w_joint = B.w_joint + 4 * W; j_def = B.j_def + 4 * oj; j_par = B.j_par + 12 * oj; j_imp = B.j_imp + 4 * oj; j_tmp = B.j_tmp + JOINT_TMP * oj;
        j_adj = B.j_adj + 2 * oj; b_jadj = B.b_jadj + 2 * ob; l_joint = B.l_joint + oj; l_jstart = B.l_jstart + ob;
        uint16_t* pw = plan + W * P.stride;
        g_cbody = pw + P.off_cbody; g_sched = pw + P.off_sched; g_slot_isl = pw + P.off_slot_isl; g_joint = pw + P.off_joint; g_lv = reinterpret_cast<uint32_t*>(pw + P.off_lv);
        g_adj = pw + P.off_adj; g_adj_start = pw + P.off_adj_start;
        g_cstate = reinterpret_cast<uint8_t*>(pw + P.off_cstate);
        g_summary = reinterpret_cast<uint32_t*>(pw + P.off_summary);
    }
    B2B_D void bind_scratch_aux(unsigned char* scratch, const scratch_layout& L)
    {
        wake_at = (int32_t*)(scratch + L.off_wake); scan = (int32_t*)(scratch + L.off_scan);
        tmp_cr = (real*)(scratch + L.off_tmp_contact);
        tmp_ci = (int32_t*)(scratch + L.off_tmp_contact + sizeof(real) * 14 * ((size_t)cap.nc + 2));
        pair_keep = (int32_t*)(scratch + L.off_pair_keep);
        prof = (long long*)(scratch + L.off_prof);
        const toi_layout T = toi_plan(cap, (int)sizeof(real));
        unsigned char* tb = scratch + L.off_toi;
        toi_alpha = (real*)(tb + T.off_alpha); toi_state = (int32_t*)(tb + T.off_state);
        toi_list = (int32_t*)(tb + T.off_list);
        toi_snap = (real*)(tb + T.off_snap); toi_snap_ci = (int32_t*)(tb + T.off_snap_ci);
        toi_red_a = (real*)(tb + T.off_red_a); toi_red_i = (int32_t*)(tb + T.off_red_i);
        toi_ib = (int32_t*)(tb + T.off_ib); toi_ic = (int32_t*)(tb + T.off_ic);
        toi_P = (real*)(tb + T.off_P); toi_V = (real*)(tb + T.off_V);
        toi_C = (toi_cc*)(tb + T.off_cc);
    }
    // `slabs` + `off`: this world's constraint-record slab (written by the prepare kernel, streamed by the solve kernels);
    // `team_prof`: the team's cycle counters (may be null outside the solve kernels)
    B2B_D void bind_scratch_solve(unsigned char* slabs, size_t off, const scratch_layout& L, unsigned char* team_prof)
    {
        size_t o_vel = off + L.off_cc_vel, o_pos = off + L.off_cc_pos, o_imp = off + L.off_cc_imp;
#if defined(__CUDA_ARCH__)
        // keep the offsets of the record streams in registers: the compiler would otherwise re-derive them from the kernel's
        // parameters (64-bit multiplies and constant-bank loads) in every level-step when registers are tight.  (Offsets,
        // not pointers: a pointer that went through an asm loses its address space and is loaded from with generic LDs.)
        asm volatile("" : "+l"(o_vel), "+l"(o_pos), "+l"(o_imp));
#endif
        cc_vel = (real*)(slabs + o_vel);
        cc_pos = (real*)(slabs + o_pos);
        cc_imp = (real*)(slabs + o_imp);
        cc_state = (real*)(slabs + off + L.off_cc_state);
        prof = (long long*)team_prof;
    }
    B2B_D void bind_smem_aux(unsigned char* smem, bool broadphase_kernel = false, bool with_pairbits = true)
    {
        const smem_layout S = smem_plan_layout(cap, (int)sizeof(real));
        s_ctl = (int32_t*)(smem + S.off_ctl);
        const size_t fat_at = broadphase_kernel ? S.off_fat : S.off_ls;
        s_fat = (real*)(smem + fat_at); s_moved = (uint8_t*)(smem + fat_at + (S.off_moved - S.off_fat));
        s_pairbits = (broadphase_kernel && with_pairbits && S.off_pairbits != 0) ? (uint32_t*)(smem + S.off_pairbits) : nullptr;
        s_flags = (uint32_t*)(smem + S.off_flags);
        s_ls = (uint16_t*)(smem + S.off_ls);
        s_cbody = (uint16_t*)(smem + S.off_cbody); s_adj = (uint16_t*)(smem + S.off_adj);
        s_adj_start = (uint16_t*)(smem + S.off_adj_start); s_stack = (uint16_t*)(smem + S.off_stack);
        s_last = s_stack;
        s_lev = (uint16_t*)(smem + S.off_lev); s_cflag = (uint8_t*)(smem + S.off_cflag);
    }
    B2B_D void bind_smem_solve(unsigned char* team_smem)
    {
        const smem_layout S = smem_plan_layout(cap, (int)sizeof(real));
        s_state = (real*)(team_smem + S.off_state);
        s_third = s_state + 2 * (size_t)cap.nb;
    }

    B2B_D shape_view<real> shape_of_fixture(int f) const { return shape_load<real>(s_meta, s_radius, s_data, f_shape[f]); }
    B2B_D int body_type(int b) const { return (int)(b_flags[b] & BODY_TYPE_MASK); }

    // Exclusive prefix sum over scan[0..n) in place; returns the total in s_ctl[slot] (valid after the call).
    // v1: one thread; the arrays are a few hundred entries and the call sites are off the critical path.
    B2B_D void exclusive_scan(int32_t* a, int n, int slot)
    {
        B2B_SYNC();
#if defined(__CUDA_ARCH__)
        if (!t_cta) {
            // lane teams: every lane sums a contiguous chunk (four loads in flight), the chunk totals are scanned with
            // shuffles, then every lane rewrites its chunk
            const int chunk = (n + t_n - 1) / t_n, lo = t_id * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
            int sum = 0;
            for (int i = lo; i < hi; i += 4) {
                int v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = (i + k < hi) ? a[i + k] : 0;
                sum += (v[0] + v[1]) + (v[2] + v[3]);
            }
            int incl = sum;
            for (int d = 1; d < t_n; d <<= 1) {
                const int up = __shfl_up_sync(t_mask, incl, d, t_n);
                if (t_id >= d) incl += up;
            }
            int run = incl - sum;
            const int total = __shfl_sync(t_mask, incl, t_n - 1, t_n);
            for (int i = lo; i < hi; i += 4) {
                int v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = (i + k < hi) ? a[i + k] : 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) if (i + k < hi) { a[i + k] = run; run += v[k]; }
            }
            if (t_id == 0) s_ctl[slot] = total;
            B2B_SYNC();
            return;
        }
#endif
        if (B2B_IS_T0) {
            int run = 0;
            for (int i = 0; i < n; ++i) { int v = a[i]; a[i] = run; run += v; }
            s_ctl[slot] = run;
        }
        B2B_SYNC();
    }

    // ------------------------------------------------------------------------------------ K1
    // Contact.update (contact.py:210-268): evaluate, carry impulses of matching ids, touching flag.
    // Returns true when the touching state changed (both bodies must be woken).
    // 128-bit (fp32) / 256-bit (fp64) row access: every [.,4] field row is one aligned quad
    B2B_D static quad ldq(const real* p) { return *reinterpret_cast<const quad*>(p); }
    B2B_D static void stq(real* p, quad q) { *reinterpret_cast<quad*>(p) = q; }

    // Everything Contact.update reads, fetched in three rounds of independent loads (contact rows; fixture rows;
    // body and shape rows) so that the flat kernel's threads wait for three memory round trips, not one per field.
    struct contact_in {
        quad qmf, qpts, qimp, qxa, qxb;
        int mt, fa, fb, ba, bb;
        uint32_t id0, id1, fl;
        shape_view<real> sa, sb;
    };
    B2B_D void contact_load(int ci, contact_in& in)
    {
        in.fa = c_fix[2 * ci]; in.fb = c_fix[2 * ci + 1];
        in.qmf = ldq(c_mf + 4 * (size_t)ci); in.qpts = ldq(c_pts + 4 * (size_t)ci); in.qimp = ldq(c_imp + 4 * (size_t)ci);
        in.mt = c_mtype[ci];
        in.id0 = c_id[2 * ci]; in.id1 = c_id[2 * ci + 1];
        in.fl = c_flags[ci];
        in.ba = f_body[in.fa]; in.bb = f_body[in.fb];
        const int sha = f_shape[in.fa], shb = f_shape[in.fb];
        in.qxa = ldq(b_xf + 4 * (size_t)in.ba); in.qxb = ldq(b_xf + 4 * (size_t)in.bb);
        in.sa = shape_load<real>(s_meta, s_radius, s_data, sha);
        in.sb = shape_load<real>(s_meta, s_radius, s_data, shb);
    }
    B2B_D bool contact_update(int ci)
    {
        // (the sensor branch lives here, not in the overload the flat narrowphase kernel calls: a call in that kernel costs it its
        // register allocation -- measured 7.0 -> 10.4 ms per 131,072 pyramid worlds; collide_one defers sensor contacts to the
        // plan kernel, which comes through here)
        const uint32_t fl = c_flags[ci];
        if (fl & CONTACT_SENSOR) { contact_update_sensor(ci, fl, c_mtype[ci]); return false; }
        contact_in in;
        contact_load(ci, in);
        return contact_update(ci, in);
    }
    // A sensor contact (contact.py:232-239): no manifold points; touching = the shapes overlap (GJK with radii,
    // distance.test_overlap); a change wakes nobody and is not a change of the contact graph.
    B2B_D_NOINLINE void contact_update_sensor(int ci, uint32_t fl, int mt)
    {
        // (scalars in, rows fetched again here: handing the caller's row struct to an out-of-line function would put it --
        // and with it the hot narrowphase path -- into local memory)
        const int fa = c_fix[2 * ci], fb = c_fix[2 * ci + 1];
        const XF xa = xf_load<real>(b_xf + 4 * (size_t)f_body[fa]), xb = xf_load<real>(b_xf + 4 * (size_t)f_body[fb]);
        const bool touching = shapes_overlap<real>(shape_of_fixture(fa), shape_of_fixture(fb), xa, xb);
        c_mtype[ci] = mt & 0xff;
        c_flags[ci] = ((fl | CONTACT_ENABLED) & ~CONTACT_TOUCHING) | (touching ? CONTACT_TOUCHING : 0u);
    }
    B2B_D bool contact_update(int ci, const contact_in& in)
    {
        manifold<real> m, old;
        const quad qimp = in.qimp;
        {
            m.type = in.mt & 0xff; m.count = (in.mt >> 8) & 0xff;
            m.local_normal = mk<real>(in.qmf.x, in.qmf.y); m.local_point = mk<real>(in.qmf.z, in.qmf.w);
            m.pt[0] = mk<real>(in.qpts.x, in.qpts.y); m.pt[1] = mk<real>(in.qpts.z, in.qpts.w);
            m.id[0] = in.id0; m.id[1] = in.id1;
        }
        old = m;
        const uint32_t fl = in.fl;
        const bool was_touching = (fl & CONTACT_TOUCHING) != 0;
        const quad qxa = in.qxa, qxb = in.qxb;
        const XF xa = xf_make<real>(mk<real>(qxa.x, qxa.y), qxa.z, qxa.w), xb = xf_make<real>(mk<real>(qxb.x, qxb.y), qxb.z, qxb.w);
        const shape_view<real> sa = in.sa, sb = in.sb;      // (copies: the out-of-line colliders take their address)
        collide_shapes(m, sa, xa, sb, xb);
        const bool touching = m.count > 0;
        // carry the impulses of matching ids; entries beyond the new point count keep their stale values
        quad nimp = qimp;
        for (int i = 0; i < m.count; ++i) {
            real ni = real(0), ti = real(0);
            if (old.count > 0 && old.id[0] == m.id[i]) { ni = qimp.x; ti = qimp.y; }
            if (old.count > 1 && old.id[1] == m.id[i]) { ni = qimp.z; ti = qimp.w; }   // last duplicate wins
            if (i == 0) { nimp.x = ni; nimp.y = ti; } else { nimp.z = ni; nimp.w = ti; }
        }
        if (m.count > 0) stq(c_imp + 4 * (size_t)ci, nimp);
        c_mtype[ci] = m.type | (m.count << 8);
        quad nmf; nmf.x = m.local_normal.x; nmf.y = m.local_normal.y; nmf.z = m.local_point.x; nmf.w = m.local_point.y;
        quad npts; npts.x = m.pt[0].x; npts.y = m.pt[0].y; npts.z = m.pt[1].x; npts.w = m.pt[1].y;
        stq(c_mf + 4 * (size_t)ci, nmf);
        stq(c_pts + 4 * (size_t)ci, npts);
        c_id[2 * ci] = m.id[0]; c_id[2 * ci + 1] = m.id[1];
        c_flags[ci] = ((fl | CONTACT_ENABLED) & ~CONTACT_TOUCHING) | (touching ? CONTACT_TOUCHING : 0u);
        return touching != was_touching;
    }

    // Narrowphase for ONE contact of this world, called by the flat kernel (one thread per contact of the whole
    // batch).  A contact one of whose bodies is awake and not static on entry is active whatever the list order
    // (contact_manager.py:193-197; bodies are only ever WOKEN during the walk), so it is evaluated here.  A contact
    // whose bodies are both inactive on entry stays pending: the reference skips it unless an EARLIER contact's
    // touching change woke one of its bodies -- which needs a touching change somewhere in the world (SUM_CHANGED);
    // only then does phase_collide walk the pending ones in list order.
    B2B_D void collide_one(int c)
    {
        // all rows are fetched before the first decision (every index below w_count[2] is a live contact)
        contact_in in;
        contact_load(c, in);
        const quad fat_a = ldq(f_fat + 4 * (size_t)in.fa), fat_b = ldq(f_fat + 4 * (size_t)in.fb);
        const uint32_t fla = b_flags[in.ba], flb = b_flags[in.bb];
        const bool ns_a = (fla & BODY_TYPE_MASK) != BODY_STATIC, ns_b = (flb & BODY_TYPE_MASK) != BODY_STATIC;
        const bool active = (ns_a && (fla & BODY_AWAKE)) || (ns_b && (flb & BODY_AWAKE));
        uint8_t state = 0;
        if (in.fl & CONTACT_SENSOR) {
            // a sensor contact is evaluated by the plan kernel (a GJK overlap test: a cold path, and a call in this kernel costs it
            // its register allocation): left pending, and the world's pending contacts have to be walked (SUM_SENSOR)
            g_cstate[c] = 0;
            if ((*g_summary & (SUM_PENDING | SUM_SENSOR)) != (SUM_PENDING | SUM_SENSOR)) B2B_ATOMIC_OR(g_summary, (uint32_t)(SUM_PENDING | SUM_SENSOR));
            return;
        }
        if (active) {
            const real fa4[4] = {fat_a.x, fat_a.y, fat_a.z, fat_a.w}, fb4[4] = {fat_b.x, fat_b.y, fat_b.z, fat_b.w};
            if (!aabb_overlap<real>(fa4, fb4)) state = 2;
            else state = contact_update(c, in) ? 3 : 1;
        }
        g_cstate[c] = state;
        if (state != 1) {
            const uint32_t bit = (state == 0) ? SUM_PENDING : (state == 2) ? SUM_DESTROY : SUM_CHANGED;
            if (!(*g_summary & bit)) B2B_ATOMIC_OR(g_summary, bit);
        }
    }

    // ContactManager.collide (contact_manager.py:174-223).  The reference walks the contact list in
    // order and a touching change wakes both bodies, which can activate LATER contacts of the same
    // step.  wake_at[b] = index of the first contact that wakes body b (-1: awake on entry) makes the
    // parallel evaluation order-exact: contact c sees body b awake iff wake_at[b] < c.  Iterate to the
    // fixed point (one extra round only when a sleeping body is woken).
    B2B_D void phase_collide(bool after_flat)
    {
        const int nb = w_count[0], nc = w_count[2];
        const int NEVER = 0x7fffffff;
        if (after_flat) {
            // every contact came back "updated, touching unchanged" or "both bodies inactive" (the settled regime): no
            // touching change means nobody was woken, so the pending contacts are exactly the ones the reference skips
            const uint32_t sum = *g_summary & SUM_COLLIDE;
            B2B_SYNC();
            if (sum != 0 && B2B_IS_T0) B2B_ATOMIC_AND(g_summary, ~(uint32_t)SUM_COLLIDE);
            if ((sum & (SUM_DESTROY | SUM_CHANGED | SUM_SENSOR)) == 0) { B2B_SYNC(); return; }
        }
        if (B2B_IS_T0) { s_ctl[0] = 0; s_ctl[1] = 0; s_ctl[10] = 0; s_ctl[12] = 0; }
        B2B_SYNC();
        // 0: pending, 1: keep (processed), 2: destroy.  After the flat kernel most contacts are already processed.
        B2B_PFOR(c, nc) {
            const int st = after_flat ? (int)g_cstate[c] : 0;
            scan[c] = (st == 3) ? 1 : st;
            if (st == 3) s_ctl[10] = 1;
            if (st == 2) s_ctl[1] = 1;
            if (st == 0) s_ctl[12] = 1;
        }
        B2B_SYNC();
        const bool any_pending = s_ctl[12] != 0;
        if (any_pending) {
            B2B_PFOR(b, nb) wake_at[b] = (b_flags[b] & BODY_AWAKE) ? -1 : NEVER;
            B2B_SYNC();
            // touching changes the flat kernel found wake their bodies at that point of the list (contact.py:257-259)
            if (after_flat) B2B_PFOR(c, nc) {
                if (g_cstate[c] != 3) continue;
                const int ba = f_body[c_fix[2 * c]], bb = f_body[c_fix[2 * c + 1]];
                B2B_ATOMIC_MIN(&wake_at[ba], c);
                B2B_ATOMIC_MIN(&wake_at[bb], c);
            }
        }
        B2B_SYNC();
        for (; any_pending;) {
            B2B_PFOR(c, nc) {
                if (scan[c] != 0) continue;
                const int fa = c_fix[2 * c], fb = c_fix[2 * c + 1];
                const int ba = f_body[fa], bb = f_body[fb];
                const bool act_a = (wake_at[ba] < c) && body_type(ba) != BODY_STATIC;
                const bool act_b = (wake_at[bb] < c) && body_type(bb) != BODY_STATIC;
                if (!act_a && !act_b) continue;
                if (!aabb_overlap<real>(f_fat + 4 * (size_t)fa, f_fat + 4 * (size_t)fb)) {
                    scan[c] = 2;
                    s_ctl[1] = 1;              // something to destroy (benign same-value race)
                    continue;
                }
                scan[c] = 1;
                if (contact_update(c)) {
                    s_ctl[10] = 1;             // a touching flag changed: the contact graph is different
                    if (B2B_ATOMIC_MIN(&wake_at[ba], c) == NEVER) s_ctl[0] = 1;
                    if (B2B_ATOMIC_MIN(&wake_at[bb], c) == NEVER) s_ctl[0] = 1;
                }
            }
            B2B_SYNC();
            const int again = s_ctl[0];
            B2B_SYNC();
            if (!again) break;
            if (B2B_IS_T0) s_ctl[0] = 0;
            B2B_SYNC();
        }
        // Body.awake = True for bodies woken by a touching change (body.py:800-804)
        if (any_pending) B2B_PFOR(b, nb) {
            if (!(b_flags[b] & BODY_AWAKE) && wake_at[b] != NEVER) {
                b_flags[b] |= BODY_AWAKE;
                b_pos[4 * (size_t)b + 3] = real(0);
            }
        }
        // ... also when no contact was pending: a sleeping body whose ONLY contact is with an awake body (it hangs on a joint, or
        // floats without gravity: a body asleep on something has a pending contact with it) is woken by that contact's touching change
        if (!any_pending && after_flat && s_ctl[10]) B2B_PFOR(c, nc) {
            if (g_cstate[c] != 3) continue;
            const int bs[2] = {f_body[c_fix[2 * c]], f_body[c_fix[2 * c + 1]]};
            for (int k = 0; k < 2; ++k) {
                if (b_flags[bs[k]] & BODY_AWAKE) continue;
                B2B_ATOMIC_OR(&b_flags[bs[k]], (uint32_t)BODY_AWAKE);
                b_pos[4 * (size_t)bs[k] + 3] = real(0);
            }
        }
        // ContactManager.destroy (contact_manager.py:160-172): order-preserving removal, nobody is woken (A.9 #12)
        if (s_ctl[1]) {
            B2B_PFOR(c, nc) scan[c] = (scan[c] == 2) ? 0 : 1;
            exclusive_scan(scan, nc, 2);
            const int kept = s_ctl[2];
            B2B_PFOR(c, nc) {
                const bool keep = (c + 1 < nc) ? (scan[c + 1] != scan[c]) : (kept != scan[c]);
                if (!keep) continue;
                const int d = scan[c];
                real* tr = tmp_cr + 14 * (size_t)d; int32_t* ti = tmp_ci + 6 * (size_t)d;
                ti[0] = c_fix[2 * c]; ti[1] = c_fix[2 * c + 1]; ti[2] = (int32_t)c_flags[c]; ti[3] = c_mtype[c];
                ti[4] = (int32_t)c_id[2 * c]; ti[5] = (int32_t)c_id[2 * c + 1];
                tr[0] = c_mat[2 * c]; tr[1] = c_mat[2 * c + 1];
                for (int k = 0; k < 4; ++k) { tr[2 + k] = c_mf[4 * (size_t)c + k]; tr[6 + k] = c_pts[4 * (size_t)c + k]; tr[10 + k] = c_imp[4 * (size_t)c + k]; }
            }
            B2B_SYNC();
            B2B_PFOR(d, kept) {
                const real* tr = tmp_cr + 14 * (size_t)d; const int32_t* ti = tmp_ci + 6 * (size_t)d;
                c_fix[2 * d] = ti[0]; c_fix[2 * d + 1] = ti[1]; c_flags[d] = (uint32_t)ti[2]; c_mtype[d] = ti[3];
                c_id[2 * d] = (uint32_t)ti[4]; c_id[2 * d + 1] = (uint32_t)ti[5];
                c_mat[2 * d] = tr[0]; c_mat[2 * d + 1] = tr[1];
                for (int k = 0; k < 4; ++k) { c_mf[4 * (size_t)d + k] = tr[2 + k]; c_pts[4 * (size_t)d + k] = tr[6 + k]; c_imp[4 * (size_t)d + k] = tr[10 + k]; }
            }
            if (B2B_IS_T0) w_count[2] = kept;
        }
        if (B2B_IS_T0 && (s_ctl[1] || s_ctl[10])) w_flags[0] &= ~WORLD_GRAPH_VALID;
        B2B_SYNC();
    }

    // ------------------------------------------------------------------------------------ K2 (plan kernel)
    B2B_D void stage_flags()
    {
        const int nb = w_count[0];
        B2B_PFOR(b, nb) { s_flags[b] = b_flags[b] & ~BODY_ISLAND; b_island[b] = -1; }
        B2B_SYNC();
    }

    // Body._contacts of every non-static body as CSR rows in contact-creation order
    // (contact_manager.py:157-158 appends, :170-172 removes in place), built in shared memory together with
    // each contact's body pair and island-eligibility bit; a copy goes to the world's plan rows for the
    // solve and broadphase kernels.  Static bodies never expand in the island DFS (world.py:524-525), so
    // their rows are not built.
    B2B_D void build_adjacency()
    {
        const int nb = w_count[0], nc = w_count[2];
        int32_t* cnt = reinterpret_cast<int32_t*>(s_lev);        // per-body counters / row cursors: shared-memory atomics

        B2B_PFOR(b, nb + 1) cnt[b] = 0;
        B2B_SYNC();
        for (int c0 = t_id; c0 < nc; c0 += 4 * t_n) {
            int fa[4], fb[4], ba[4], bb[4];
            uint32_t cf[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = c0 + k * t_n;
                fa[k] = fb[k] = 0; cf[k] = 0u;
                if (c < nc) { fa[k] = c_fix[2 * c]; fb[k] = c_fix[2 * c + 1]; cf[k] = c_flags[c]; }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { ba[k] = f_body[fa[k]]; bb[k] = f_body[fb[k]]; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = c0 + k * t_n;
                if (c >= nc) continue;
                s_cbody[2 * c] = (uint16_t)ba[k]; s_cbody[2 * c + 1] = (uint16_t)bb[k];
                g_cbody[2 * c] = (uint16_t)ba[k]; g_cbody[2 * c + 1] = (uint16_t)bb[k];
                s_cflag[c] = (uint8_t)(((cf[k] & CONTACT_ENABLED) && (cf[k] & CONTACT_TOUCHING) && !(cf[k] & CONTACT_SENSOR)) ? 1 : 0);   // solid + touching (world.py:528-533)
                if ((s_flags[ba[k]] & BODY_TYPE_MASK) != BODY_STATIC) B2B_ATOMIC_ADD(&cnt[ba[k]], 1);
                if ((s_flags[bb[k]] & BODY_TYPE_MASK) != BODY_STATIC) B2B_ATOMIC_ADD(&cnt[bb[k]], 1);
            }
        }
        exclusive_scan(cnt, nb + 1, 3);
        B2B_PFOR(b, nb + 1) { s_adj_start[b] = (uint16_t)cnt[b]; g_adj_start[b] = (uint16_t)cnt[b]; }
        B2B_SYNC();
        B2B_PFOR(c, nc) {
            const int ba = s_cbody[2 * c], bb = s_cbody[2 * c + 1];
            if ((s_flags[ba] & BODY_TYPE_MASK) != BODY_STATIC) s_adj[B2B_ATOMIC_ADD(&cnt[ba], 1)] = (uint16_t)c;
            if ((s_flags[bb] & BODY_TYPE_MASK) != BODY_STATIC) s_adj[B2B_ATOMIC_ADD(&cnt[bb], 1)] = (uint16_t)c;
        }
        B2B_SYNC();
        // restore creation order inside each row (rows are short: <= contacts per body)
        B2B_PFOR(b, nb) {
            const int s = s_adj_start[b], e = s_adj_start[b + 1];
            for (int i = s + 1; i < e; ++i) {
                const uint16_t v = s_adj[i];
                int j = i - 1;
                while (j >= s && s_adj[j] > v) { s_adj[j + 1] = s_adj[j]; --j; }
                s_adj[j + 1] = v;
            }
            for (int i = s; i < e; ++i) g_adj[i] = s_adj[i];
        }
        B2B_SYNC();
    }

    // World._solve island construction (world.py:484-580): seeds in body order, LIFO DFS over
    // Body._contacts in creation order, statics join but do not propagate.  Strictly sequential per
    // world; the lists it emits (l_body / l_contact / l_start) are graded integer outputs.  All the state the
    // walk touches is in shared memory; flags and sleep timers are written back afterwards.
    B2B_D void phase_islands()
    {
        const int nb = w_count[0], nc = w_count[2];
        // Joint._island_flag of every joint cleared (world.py:491-492); only batches with a joint capacity have joints
        int nj = 0;
        if (cap.nj > 0) {
            nj = w_joint[0];
            if (nj > cap.nj) { nj = cap.nj; if (B2B_IS_T0) w_status[0] |= STATUS_JOINT_OVERFLOW; }
            B2B_PFOR(j, nj) j_def[4 * j + 3] &= ~(int32_t)JF_ISLAND;
            B2B_SYNC();
        }
        if (B2B_IS_T0) {
            int n_islands = 0, lb = 0, lc = 0, lj = 0;
            bool overflow = false;
            for (int seed = 0; seed < nb && !overflow; ++seed) {
                const uint32_t fl = s_flags[seed];
                if (fl & BODY_ISLAND) continue;
                if (!(fl & BODY_AWAKE) || !(fl & BODY_ACTIVE)) continue;
                if ((fl & BODY_TYPE_MASK) == BODY_STATIC) continue;
                const int body_start = lb, contact_start = lc, joint_start = lj;
                int sp = 0, n_static = 0;
                int statics[4];                                        // statics that joined this island
                s_stack[sp++] = (uint16_t)seed;
                s_flags[seed] = fl | BODY_ISLAND;
                while (sp > 0) {
                    const int b = s_stack[--sp];
                    if (lb >= cap.nl) { overflow = true; break; }
                    b_island[b] = lb - body_start;
                    l_body[lb++] = b;
                    // Body.awake = True (body.py:800-804): a sleeping body that is reached restarts its sleep timer
                    uint32_t bf = s_flags[b];
                    if (!(bf & BODY_AWAKE)) { bf |= BODY_AWAKE | BODY_WOKEN; s_flags[b] = bf; }
                    if ((bf & BODY_TYPE_MASK) == BODY_STATIC) {
                        if (n_static < 4) statics[n_static] = b;
                        ++n_static;
                        continue;
                    }
                    // the reference appends all of b's free contacts first and then walks the appended ones
                    // (world.py:531-546); a contact is in b's list once and the second pass only touches flags the
                    // first never reads, so one pass gives the same lists without reading l_contact back from memory
                    // (four entries of the row at a time, their contact words fetched before any of them is looked at: a row
                    // names a contact once, so nothing this round writes can change what was fetched, and the walk -- a chain
                    // of dependent shared-memory loads, ~110 cycles per entry -- waits once per round instead of per entry)
                    const int row_end = s_adj_start[b + 1];
                    for (int k0 = s_adj_start[b]; k0 < row_end; k0 += 4) {
                        int cs[4];
                        uint8_t fs[4];
                        uint32_t pair[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) cs[j] = s_adj[(k0 + j < row_end) ? k0 + j : k0];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            fs[j] = s_cflag[cs[j]];
                            pair[j] = (uint32_t)s_cbody[2 * cs[j]] | ((uint32_t)s_cbody[2 * cs[j] + 1] << 16);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (k0 + j >= row_end || fs[j] != 1) continue;   // already in an island, disabled, a sensor or not touching
                            const int c = cs[j];
                            s_lev[lc] = (uint16_t)c;                   // (a shared-memory copy of the list for phase_schedule: its sequential
                            l_contact[lc++] = c;                       //  loop then waits on no global load; the per-body counters are done with)
                            s_cflag[c] = 3;                            // island flag
                            const int ba = (int)(pair[j] & 0xffffu), bb = (int)(pair[j] >> 16);
                            const int other = (b == ba) ? bb : ba;
                            if (s_flags[other] & BODY_ISLAND) continue;
                            s_stack[sp++] = (uint16_t)other;
                            s_flags[other] |= BODY_ISLAND;
                        }
                    }
                    // Body.joints in append order (world.py:548-572): a joint joins the island of the first of its bodies the
                    // walk expands, unless its other body is inactive; the other body is taken along
                    if (nj > 0) {
                        for (int k = b_jadj[2 * b], ke = k + b_jadj[2 * b + 1]; k < ke; ++k) {
                            const int j = j_adj[k];
                            int32_t* def = j_def + 4 * (size_t)j;
                            if (def[3] & JF_ISLAND) continue;
                            const int other = (def[1] == b) ? def[2] : def[1];
                            if (other < 0) { l_joint[lj++] = j; def[3] |= JF_ISLAND; continue; }   // a joint on one body (MouseJoint): world.py:556-560
                            if (!(s_flags[other] & BODY_ACTIVE)) continue;
                            l_joint[lj++] = j;
                            def[3] |= JF_ISLAND;
                            if (s_flags[other] & BODY_ISLAND) continue;
                            s_stack[sp++] = (uint16_t)other;
                            s_flags[other] |= BODY_ISLAND;
                        }
                    }
                }
                if (overflow) break;
                l_start[2 * n_islands] = body_start; l_start[2 * n_islands + 1] = contact_start;
                if (cap.nj > 0) l_jstart[n_islands] = joint_start;
                // statics may join other islands (world.py:577-580); the few that joined were noted on the way (the
                // island's body list is in global memory: reading it back would cost a round trip per island)
                if (n_static <= 4) {
                    for (int k = 0; k < n_static; ++k) s_flags[statics[k]] &= ~BODY_ISLAND;
                } else {
                    for (int k = body_start; k < lb; ++k) {
                        const int b = l_body[k];
                        if ((s_flags[b] & BODY_TYPE_MASK) == BODY_STATIC) s_flags[b] &= ~BODY_ISLAND;
                    }
                }
                ++n_islands;
            }
            if (overflow) w_status[0] |= STATUS_ISLAND_OVERFLOW;
            s_ctl[11] = 0;
            w_island[0] = n_islands; w_island[1] = lb; w_island[2] = lc; w_island[3] = 0;
            if (cap.nj > 0) w_joint[1] = lj;
            s_ctl[4] = n_islands; s_ctl[5] = lb; s_ctl[6] = lc;
        }
        B2B_SYNC();
        B2B_PFOR(b, nb) {
            uint32_t f = s_flags[b];
            if (f & BODY_WOKEN) { b_pos[4 * (size_t)b + 3] = real(0); f &= ~BODY_WOKEN; s_flags[b] = f; s_ctl[11] = 1; }
            b_flags[b] = f;
            s_last[b] = 0;                                 // the DFS stack's storage becomes the scheduler's `last level`
        }
        // (four contacts per lane and round, loads first: one memory round trip per round, not per contact)
        for (int c0 = t_id; c0 < nc; c0 += 4 * t_n) {
            uint32_t cf[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int c = c0 + k * t_n; if (c < nc) cf[k] = c_flags[c]; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = c0 + k * t_n;
                if (c < nc) c_flags[c] = (s_cflag[c] == 3) ? (cf[k] | CONTACT_ISLAND) : (cf[k] & ~CONTACT_ISLAND);
            }
        }
        B2B_SYNC();
    }

    // Dependency levels of every island's contact list: level(k) = 1 + max level of an earlier contact that
    // shares a non-static body.  Contacts of one level touch disjoint non-static bodies, so solving a level's
    // contacts concurrently gives bit-identical results to the reference's sequential Gauss-Seidel sweep,
    // and every body still sees its contacts in list order.  A stable counting sort by level gives `sched`
    // (slot -> position in l_contact) and the per-island level table `ls`; both go to the world's plan rows.
    // Islands are independent, so each one is scheduled by its own thread.
    B2B_D void phase_schedule(int solve_lanes)
    {
        const int n_islands = s_ctl[4], n_lb = s_ctl[5], n_lc = s_ctl[6];
        (void)n_lb;
        const int chunk = (solve_lanes < (int)LV_MAX_LANES) ? (solve_lanes < 1 ? 1 : solve_lanes) : (int)LV_MAX_LANES;
        // One island's slots [c0, c1) -- or, for the joint schedule, the slots of ALL islands taken as one list: islands
        // share no non-static body, so the level of a contact is the same either way and level l of the joint list is
        // the union of every island's level l.
        auto schedule_range = [&](int c0, int c1, int isl) {
            const int ncon = c1 - c0;
            uint16_t* ls = s_ls + c0 + 3 * isl;           // level l (1-based) occupies slots [c0 + ls[l], c0 + ls[l+1])
            int nlev = 0;
            if (ncon > 0) {
                for (int k = 0; k < ncon + 3; ++k) ls[k] = 0;
                // (four slots per round: the island's contact list is in global memory, so its entries -- and with them the
                // body pairs and body types, which this loop never writes -- are fetched before the round's sequential part)
                for (int k0 = c0; k0 < c1; k0 += 4) {
                    int cs[4], ba[4], bb[4];
                    bool dyn_a[4], dyn_b[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) cs[j] = s_lev[(k0 + j < c1) ? k0 + j : k0];   // the list as phase_islands left it; overwritten by the levels below
#pragma unroll
                    for (int j = 0; j < 4; ++j) { ba[j] = s_cbody[2 * cs[j]]; bb[j] = s_cbody[2 * cs[j] + 1]; }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        dyn_a[j] = (s_flags[ba[j]] & BODY_TYPE_MASK) != BODY_STATIC;
                        dyn_b[j] = (s_flags[bb[j]] & BODY_TYPE_MASK) != BODY_STATIC;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = k0 + j;
                        if (k >= c1) continue;
                        int l = 0;
                        if (dyn_a[j]) l = s_last[ba[j]];
                        if (dyn_b[j] && s_last[bb[j]] > l) l = s_last[bb[j]];
                        ++l;
                        if (dyn_a[j]) s_last[ba[j]] = (uint16_t)l;
                        if (dyn_b[j]) s_last[bb[j]] = (uint16_t)l;
                        s_lev[k] = (uint16_t)l;
                        ++ls[l + 1];                      // count of level l parked one slot up
                        if (l > nlev) nlev = l;
                    }
                }
                // ls[l + 1] = count(l)  ->  ls[l] = start(l) for l = 1..nlev+1  (ls[nlev + 1] = ncon)
                int run = 0;
                for (int l = 1; l <= nlev; ++l) { const int cnt = ls[l + 1]; ls[l] = (uint16_t)run; run += cnt; }
                ls[nlev + 1] = (uint16_t)run;
                // stable placement; ls[l] is advanced while filling and restored afterwards
                for (int k = c0; k < c1; ++k) { const int l = s_lev[k]; g_sched[c0 + ls[l]] = (uint16_t)k; ++ls[l]; }
                for (int l = nlev; l >= 2; --l) ls[l] = ls[l - 1];
                ls[1] = 0;
                // level-step table: every level in pieces of at most `chunk` slots, one 4-bit count per piece
                uint32_t* lv = g_lv + lv_offset(c0, isl);
                uint32_t word = 0;
                int nib = 0;
                for (int l = 1; l <= nlev; ++l) {
                    int left = (int)ls[l + 1] - (int)ls[l];
                    while (left > 0) {
                        const int take = left < chunk ? left : chunk;
                        word |= (uint32_t)take << (4 * nib);
                        left -= take;
                        if (++nib == 8) { *lv++ = word; word = 0; nib = 0; }
                    }
                }
                *lv++ = word;                             // the partial word, or a zero word: the reader runs one word ahead
                if (nib != 0) *lv = 0;
            }
        };
        // Joint schedule (B2B_JOINT_ISLANDS): worlds with several islands that have contacts (stacks side by side) are
        // solved level-step by level-step across ALL their islands at once instead of island after island; velocity
        // sweeps are the same operations in the same per-body order, position sweeps skip the slots of islands that are
        // already solved (island.py:141-153 stops iterating per island).  Up to 32 such islands (one mask word).
        int n_ci = 0, largest = 0;
        for (int isl = 0; isl < n_islands; ++isl) {
            const int c0 = l_start[2 * isl + 1];
            const int c1 = (isl + 1 < n_islands) ? l_start[2 * isl + 3] : n_lc;
            n_ci += (c1 > c0) ? 1 : 0;
            largest = (c1 - c0 > largest) ? c1 - c0 : largest;
        }
#if defined(B2B_JOINT_ISLANDS) && !B2B_JOINT_ISLANDS
        const bool joint = false;
#else
        // Worth it only when no island dominates: the joint list is at least as long as the largest island's, and one world
        // with a joint schedule moves the whole batch to the group-capable position kernel (measured 2 % of the solve time
        // slower for the worlds that stay one island per group), so a stray second island next to a big one (a box beside a
        // pyramid: ~1 % of the contacts) does not qualify.
        // (batches with a joint capacity solve island after island: an island's joints are visited between its contact sweeps)
        const bool joint = (n_ci >= 2) && (n_ci <= 32) && (3 * (size_t)cap.nb >= 32) && (10 * largest <= 9 * n_lc) && cap.nj == 0;
#endif
        if (joint) {
            if (B2B_IS_T0) schedule_range(0, n_lc, 0);
            B2B_SYNC();
            // contact island (ordinal among the islands with contacts) of every list position, then of every slot
            if (B2B_IS_T0) {
                int ord = 0;
                for (int isl = 0; isl < n_islands; ++isl) {
                    const int c0 = l_start[2 * isl + 1];
                    const int c1 = (isl + 1 < n_islands) ? l_start[2 * isl + 3] : n_lc;
                    for (int k = c0; k < c1; ++k) s_lev[k] = (uint16_t)ord;
                    ord += (c1 > c0) ? 1 : 0;
                }
            }
            B2B_SYNC();
            B2B_PFOR(slot, n_lc) g_slot_isl[slot] = s_lev[g_sched[slot]];
        } else {
            B2B_PFOR(isl, n_islands) {
                const int c0 = l_start[2 * isl + 1];
                const int c1 = (isl + 1 < n_islands) ? l_start[2 * isl + 3] : n_lc;
                schedule_range(c0, c1, isl);
            }
        }
        if (B2B_IS_T0) { g_joint[0] = (uint16_t)(joint ? 1 : 0); g_joint[1] = (uint16_t)n_ci; }
        B2B_SYNC();
    }

#include "b2b_joints.inl"

    // ------------------------------------------------------------------------------------ K3 (prepare + solve kernels)
    // One constraint = a read-only VELOCITY record of 8 four-vectors ("quads": exactly one 128-byte line in fp32 -- what a
    // velocity sweep streams per constraint), a read-only POSITION record of 6 quads, and one read-write quad of accumulated
    // impulses in its own array (updating it never dirties the read-only lines).  All are in schedule order in the world's
    // record slab, written by the flat prepare kernel (one thread per constraint of the whole group of worlds) and streamed
    // by the solve kernels, whose lanes fetch the record of the NEXT level-step into registers while they solve the current
    // one.  Records are read and written two quads ("oct", 32 bytes in fp32) at a time: every lane of a level-step reads its
    // own line, so a load instruction costs the L1 one wavefront per active lane whatever its width, and the 256-bit form
    // (LDG.E.256) needs four per velocity record where 128-bit loads needed eight.  Integer fields ride in spare lanes of
    // the real records as raw bits.  NOTHING else is read from global memory while a slot is solved: every load of the level
    // loop shares one hardware scoreboard with the record stream, so a load inside the solve -- even a predicated-off one --
    // makes its first use wait for the next record as well (measured: +20 % on the velocity kernel).
    //   v0  normal.x normal.y friction [body_a | body_b << 16]     v1  rA0.x rA0.y rB0.x rB0.y
    //   v2  nmass0   tmass0   nmass1   tmass1                      v3  rA1.x rA1.y rB1.x rB1.y
    //   v4  k11      k12      k22      [vinfo]                     v5  nm11 nm12 nm22 -
    //   v6  inv_mass_a inv_I_a inv_mass_b inv_I_b                  v7  bias0 bias1 - -
    //   p0  inv_mass_a inv_I_a inv_mass_b inv_I_b                  p1  ln.x ln.y lp.x lp.y
    //   p2  p0.x p0.y p1.x p1.y                                    p3  radius_a radius_b [bodies] [info]
    //   p4  lc_a.x lc_a.y lc_b.x lc_b.y                            p5  -
    //   imp ni0 ti0 ni1 ti1
    //   vinfo = n_vel | static_a << 2 | static_b << 3 | contact index << 8
    //   info  = n_vel | type << 8 | n_pos << 16 | static_a << 24 | static_b << 25
    // Shared memory holds three reals per body: its velocity while velocities are being solved, its sweep position
    // afterwards.  Static bodies never change and are shared between islands: each island stages their velocity
    // and later their position for itself (islands are solved one after the other).
    enum { INFO_STATIC_A = 1 << 24, INFO_STATIC_B = 1 << 25, INFO_ISLAND_SHIFT = 26 };   // bits 26-30 (position records): island ordinal of a joint schedule
    enum { VINFO_NVEL_MASK = 3, VINFO_STATIC_A = 1 << 2, VINFO_STATIC_B = 1 << 3, VINFO_CI_SHIFT = 8 };
    struct alignas(8 * sizeof(real)) oct { quad lo, hi; };
    enum { BODY_BATCH = 4 };     // (8 measured slower: register spills reach the level loop)
    B2B_D static quad mkq(real x, real y, real z, real w) { quad q; q.x = x; q.y = y; q.z = z; q.w = w; return q; }
    B2B_D static real real_from_bits(uint32_t u)
    {
        union { real r; uint32_t w[sizeof(real) / 4]; } x;
        x.r = real(0); x.w[0] = u;
        return x.r;
    }
    B2B_D static uint32_t bits_of_real(real r)
    {
        union { real r; uint32_t w[sizeof(real) / 4]; } x;
        x.r = r;
        return x.w[0];
    }
    B2B_D oct* vel_rec(int slot) { return reinterpret_cast<oct*>(cc_vel) + (size_t)(VREC_QUADS / 2) * slot; }
    B2B_D oct* pos_rec(int slot) { return reinterpret_cast<oct*>(cc_pos) + (size_t)(PREC_QUADS / 2) * slot; }
    B2B_D quad* imp_rec(int slot) { return reinterpret_cast<quad*>(cc_imp) + slot; }
    // (x, y) pairs first, then the third component: one two-wide and one scalar shared-memory access per body
    struct alignas(2 * sizeof(real)) pair2 { real x, y; };
    B2B_D quad state_get(int b) const
    {
        const pair2 xy = reinterpret_cast<const pair2*>(s_state)[b];
        return mkq(xy.x, xy.y, s_third[b], real(0));
    }
    B2B_D void state_set(int b, real x, real y, real z)
    {
        pair2 xy; xy.x = x; xy.y = y;
        reinterpret_cast<pair2*>(s_state)[b] = xy;
        s_third[b] = z;
    }
    B2B_D bool is_static_body(int b) const { return (b_flags[b] & BODY_TYPE_MASK) == BODY_STATIC; }

    B2B_D static V2 rel_velocity(V2 vb, real wb, V2 rb, V2 va, real wa, V2 ra)
    {
        return vb + cross_sv(wb, rb) - va - cross_sv(wa, ra);
    }
    // A lane's position in the record streams of the world: the velocity record + impulse row, or the position record, of
    // one slot.  Records are consumed in slot order, so the position only ever moves forward by a few slots (pointer
    // arithmetic, no index -> address conversion on the level loop's dependent chain); `prefetch` asks the L2 for the slot
    // PREFETCH_AHEAD further on (an immediate offset of the same pointers; a request past the island's last slot is
    // harmless, it stays inside the world's slab), `prefetch_first` for the first slots of the island: the next sweep.
#ifndef B2B_PREFETCH_AHEAD
#define B2B_PREFETCH_AHEAD 6
#endif
    enum { PREFETCH_AHEAD = B2B_PREFETCH_AHEAD, VREC_OCTS = VREC_QUADS / 2, PREC_OCTS = PREC_QUADS / 2 };
    struct vrec { quad q0, q1, q2, q3, q4, q5, q6, q7, qi; };      // the quads v0..v7 + imp of the layout above, in registers
    struct prec { quad q7, q8, q9, q10, q11; };                    // p0..p4
    struct vel_stream {
        const oct* rec; quad* imp;
        B2B_D vel_stream ahead(int n) const { vel_stream t; t.rec = rec + (int)VREC_OCTS * n; t.imp = imp + n; return t; }
        B2B_D void prefetch() const
        {
            B2B_PREFETCH(rec + (int)VREC_OCTS * (int)PREFETCH_AHEAD);
            if (sizeof(real) > 4) B2B_PREFETCH(reinterpret_cast<const char*>(rec + (int)VREC_OCTS * (int)PREFETCH_AHEAD) + 128);
            B2B_PREFETCH(imp + (int)PREFETCH_AHEAD);
        }
        B2B_D void prefetch_first(int t_n) const
        {
            for (int k = 0; k < (int)PREFETCH_AHEAD; k += t_n) { B2B_PREFETCH(rec + (int)VREC_OCTS * k); B2B_PREFETCH(imp + k); }
        }
        B2B_D void load(vrec& r) const
        {
            const oct o0 = rec[0], o1 = rec[1], o2 = rec[2], o3 = rec[3];
            r.q0 = o0.lo; r.q1 = o0.hi; r.q2 = o1.lo; r.q3 = o1.hi; r.q4 = o2.lo; r.q5 = o2.hi; r.q6 = o3.lo; r.q7 = o3.hi;
            r.qi = *imp;
        }
    };
    struct pos_stream {
        const oct* rec;
        B2B_D pos_stream ahead(int n) const { pos_stream t; t.rec = rec + (int)PREC_OCTS * n; return t; }
        B2B_D void prefetch() const
        {
            B2B_PREFETCH(rec + (int)PREC_OCTS * (int)PREFETCH_AHEAD);
            B2B_PREFETCH(rec + (int)PREC_OCTS * (int)PREFETCH_AHEAD + ((int)PREC_OCTS - 1));   // the record's last oct: its second line, if it has one
        }
        B2B_D void prefetch_first(int t_n) const
        {
            for (int k = 0; k < (int)PREFETCH_AHEAD; k += t_n) { B2B_PREFETCH(rec + (int)PREC_OCTS * k); B2B_PREFETCH(rec + (int)PREC_OCTS * k + ((int)PREC_OCTS - 1)); }
        }
        B2B_D void load(prec& r) const
        {
            const oct o0 = rec[0], o1 = rec[1], o2 = rec[2];
            r.q7 = o0.lo; r.q8 = o0.hi; r.q9 = o1.lo; r.q10 = o1.hi; r.q11 = o2.lo;
        }
    };
    B2B_D vel_stream vel_stream_at(int slot) { vel_stream t; t.rec = vel_rec(slot); t.imp = imp_rec(slot); return t; }
    B2B_D pos_stream pos_stream_at(int slot) { pos_stream t; t.rec = pos_rec(slot); return t; }

    // Reader of an island's level-step table (phase_schedule): one 4-bit slot count per level-step, eight per
    // word.  The following word is requested when the current one is half used and only moved in when it runs out,
    // so its load never sits on the level loop's critical path.
    struct lv_reader {
        const uint32_t* p;
        uint32_t cur, next;
        int left;
        B2B_D void start(const uint32_t* lv) { cur = lv[0]; next = 0u; p = lv + 1; left = 8; }
        B2B_D int take()
        {
            const int c = (int)(cur & 15u);
            cur >>= 4;
            --left;
            // (branches, not selects: a select would read `next` in every level-step, i.e. wait for its load -- and with it for
            // the record loads just issued, which share its scoreboard -- right after it was requested)
            if (left == 0) { cur = next; left = 8; }
            else if (left == 4) next = *p++;
            return c;
        }
    };

    // One Gauss-Seidel sweep over an island's constraint slots [c0, c1), level-step by level-step.  The slots of a
    // level-step touch disjoint non-static bodies, so the team's lanes solve them concurrently and the result is
    // bit-identical to the reference's sequential sweep.  `at` is the lane's stream position at slot c0 + t_id,
    // `solve(rec, at)` applies a record to the bodies in shared memory, `ready(rec)` returns a word of the record.
    // On the GPU every lane fetches the record of its NEXT level-step before it solves the current one, so the
    // L2/HBM latency of the record stream is off the dependent chain.  The advance of the stream positions is made to
    // depend on a word of the CURRENT record (`& zero`, a run-time zero -- it has to be a real instruction: the hardware
    // scheduler's view is the machine code): all loads of the loop share one hardware scoreboard, so without that order
    // the first use of the current record would also wait for the loads just issued.  The table reader moves its
    // look-ahead word in at the same point, for the same reason.
#if defined(B2B_VAR_NOSYNC)
#define B2B_LEVEL_SYNC() ((void)0)
#else
#define B2B_LEVEL_SYNC() __syncwarp(t_mask)
#endif
    template <class Rec, class Stream, class Solve, class Ready>
    B2B_D void level_sweep(int c0, int c1, const uint32_t* lv, uint32_t zero, Stream at, Solve solve, Ready ready)
    {
#if defined(__CUDA_ARCH__)
        asm volatile("" : "+r"(zero));                              // (kept in a register: not re-read from the parameters every level-step)
        lv_reader rd;
        rd.start(lv);
        // two record buffers used alternately (the loop body is written out twice), so nothing is copied
        int begin = c0, cnt = rd.take();
        bool act = t_id < cnt;
        Stream sa = at, sb = at;
        Rec ra, rb;
        if (act) sa.load(ra);
#define B2B_LEVEL_STEP(CUR, SCUR, NXT, SNXT)                                                                          \
            {                                                                                                         \
                /* (an idle lane's record is whatever its registers held: only its arrival matters) */                \
                const int adv = cnt + (int)(ready(CUR) & zero);                                                       \
                const int nbegin = begin + adv;                                                                       \
                B2B_TSTAMP(0, nbegin);                                                                                \
                const int ncnt = rd.take();                                                                           \
                const bool more = (nbegin < c1) & (ncnt != 0);      /* a zero count never occurs in a valid table */  \
                const bool nact = more & (t_id < ncnt);                                                               \
                SNXT = SCUR.ahead(adv);                                                                               \
                if (nact) { SNXT.prefetch(); SNXT.load(NXT); }                                                        \
                B2B_TSTAMP(1, ncnt);                                                                                  \
                if (act) solve(CUR, SCUR);                                                                            \
                B2B_TSTAMP(4, 0);                                                                                     \
                B2B_LEVEL_SYNC();                                                                                     \
                B2B_TSTAMP(5, 0);                                                                                     \
                B2B_TSTAMP(6, 0);                                                                                     \
                if (!more) break;                                                                                     \
                act = nact; begin = nbegin; cnt = ncnt;                                                               \
            }
        for (;;) {
            B2B_LEVEL_STEP(ra, sa, rb, sb)
            B2B_LEVEL_STEP(rb, sb, ra, sa)
        }
#undef B2B_LEVEL_STEP
        at.prefetch_first(t_n);                                     // the sweep that follows, if any
#else
        // (sequential build: slot order is the level-step order; the table is the GPU schedule's business)
        (void)ready; (void)lv; (void)zero;
        Stream s = at;
        for (int slot = c0; slot < c1; ++slot) { Rec r; s.load(r); solve(r, s); s = s.ahead(1); }
#endif
    }

    // island.py:84-95: the velocity an island body carries into the contact solver (gravity and forces applied,
    // then damped); static and kinematic bodies keep theirs.  Evaluated by the prepare kernel (restitution bias)
    // and by the solve kernel (staging) with the same expression.
    B2B_D void integrated_velocity(int b, real dt, V2 gravity, V2& v, real& w) const
    {
        const quad bv = ldq(b_vel + 4 * (size_t)b);
        v = mk<real>(bv.x, bv.y);
        w = bv.z;
        if ((b_flags[b] & BODY_TYPE_MASK) == BODY_DYNAMIC) {
            const quad damp = ldq(b_damp + 4 * (size_t)b), mass = ldq(b_mass + 4 * (size_t)b);
            const V2 force = mk<real>(b_force[2 * (size_t)b], b_force[2 * (size_t)b + 1]);
            const real torque = bv.w;
            v = v + dt * (damp.z * gravity + mass.x * force);
            w = w + dt * mass.y * torque;
            v = clamp_r(real(1) - dt * damp.x, real(0), real(1)) * v;
            w = w * clamp_r(real(1) - dt * damp.y, real(0), real(1));
        }
    }

    // WorldManifold (contact_util.py:195-245): the contact normal and the world points of both manifold points
    B2B_D static void world_manifold(int type, V2 local_normal, V2 local_point, const V2* lp, const XF& xa, const XF& xb,
                                     real rad_a, real rad_b, V2& normal, V2* wp)
    {
        if (type == MANIFOLD_CIRCLES) {
            normal = mk<real>(real(1), real(0));
            const V2 pa = xf_mul(xa, local_point), pb = xf_mul(xb, lp[0]);
            const V2 d = pa - pb;
            if (dot(d, d) > K<real>::epsilon_sqr) normal = normalize(pb - pa);
            const V2 c_a = pa + rad_a * normal, c_b = pb - rad_b * normal;
            wp[0] = real(0.5) * (c_a + c_b);
            wp[1] = wp[0];
        } else {
            // FACE_A and FACE_B are the same arithmetic with the bodies' roles exchanged (and the normal flipped at
            // the end for FACE_B); selecting the operands keeps the lanes of a warp on one code path
            const bool face_a = (type == MANIFOLD_FACE_A);
            const XF xr = face_a ? xa : xb, xo = face_a ? xb : xa;
            const real rad_r = face_a ? rad_a : rad_b, rad_o = face_a ? rad_b : rad_a;
            normal = rot_mul(xr, local_normal);
            const V2 plane = xf_mul(xr, local_point);
            for (int j = 0; j < 2; ++j) {
                const V2 clip = xf_mul(xo, lp[j]);
                const V2 c_r = clip + (rad_r - dot(clip - plane, normal)) * normal;
                const V2 c_o = clip - rad_o * normal;
                wp[j] = real(0.5) * (c_r + c_o);        // IEEE addition commutes: same bits as 0.5 * (c_a + c_b)
            }
            if (!face_a) normal = -normal;
        }
    }

    // ContactSolver.__init__ + initialize_velocity_constraints for one slot
    // (contact.py:376-496, contact_util.py:195-245 WorldManifold, :253-277 VelocityConstraintPoint).
    // Runs in the flat prepare kernel: reads global state only.
    B2B_D void constraint_init(int slot, real dt, real dt_ratio)
    {
        const bool warm = (w_flags[0] & WORLD_WARM_STARTING) != 0;
        const V2 gravity = mk<real>(w_gravity[0], w_gravity[1]);
        const int ci = l_contact[g_sched[slot]];
        const int fa = c_fix[2 * ci], fb = c_fix[2 * ci + 1];
        const int ba = g_cbody[2 * ci], bb = g_cbody[2 * ci + 1];
        const int mt = c_mtype[ci];
        const int type = mt & 0xff, count = (mt >> 8) & 0xff;
        const quad mf = ldq(c_mf + 4 * (size_t)ci), pts = ldq(c_pts + 4 * (size_t)ci), imp = ldq(c_imp + 4 * (size_t)ci);
        const V2 local_normal = mk<real>(mf.x, mf.y), local_point = mk<real>(mf.z, mf.w);
        V2 lp[2]; lp[0] = mk<real>(pts.x, pts.y); lp[1] = mk<real>(pts.z, pts.w);
        const real imp_n[2] = {imp.x, imp.z}, imp_t[2] = {imp.y, imp.w};
        const real rad_a = s_radius[f_shape[fa]], rad_b = s_radius[f_shape[fb]];
        const real friction = c_mat[2 * ci], restitution = c_mat[2 * ci + 1];
        const quad mass_a = ldq(b_mass + 4 * (size_t)ba), mass_b = ldq(b_mass + 4 * (size_t)bb);
        const real ma = mass_a.x, mb = mass_b.x, ia = mass_a.y, ib = mass_b.y;
        const V2 lca = mk<real>(mass_a.z, mass_a.w), lcb = mk<real>(mass_b.z, mass_b.w);
        const bool stat_a = is_static_body(ba), stat_b = is_static_body(bb);
        // positions are the step's start positions; velocities are the integrated ones
        const quad pa_ = ldq(b_pos + 4 * (size_t)ba), pb_ = ldq(b_pos + 4 * (size_t)bb);
        const V2 ca = mk<real>(pa_.x, pa_.y), cb = mk<real>(pb_.x, pb_.y);
        V2 va, vb;
        real wa, wb;
        integrated_velocity(ba, dt, gravity, va, wa);
        integrated_velocity(bb, dt, gravity, vb, wb);
        const XF xa = xf_from_sweep<real>(ca, pa_.z, lca);
        const XF xb = xf_from_sweep<real>(cb, pb_.z, lcb);

        // world manifold
        V2 normal, wp[2];
        world_manifold(type, local_normal, local_point, lp, xa, xb, rad_a, rad_b, normal, wp);
        const V2 tangent = right_perp(normal);
        int n_vel = count;
        V2 ra[2], rb[2];
        real nmass[2], tmass[2], bias[2], ni[2], ti[2];
        for (int j = 0; j < 2; ++j) {
            ra[j] = rb[j] = mk<real>(real(0), real(0));
            nmass[j] = tmass[j] = bias[j] = ni[j] = ti[j] = real(0);
            if (j >= count) continue;
            ra[j] = wp[j] - ca; rb[j] = wp[j] - cb;
            const real rn_a = cross(ra[j], normal), rn_b = cross(rb[j], normal);
            const real k_normal = ma + mb + ia * sq_r(rn_a) + ib * sq_r(rn_b);      // contact.py:444-445: ** 2
            const real rt_a = cross(ra[j], tangent), rt_b = cross(rb[j], tangent);
            const real k_tangent = ma + mb + ia * sq_r(rt_a) + ib * sq_r(rt_b);     // contact.py:456-457
            const real v_rel = dot(normal, rel_velocity(vb, wb, rb[j], va, wa, ra[j]));
            if (v_rel < -K<real>::velocity_threshold) bias[j] = -restitution * v_rel;
            nmass[j] = (k_normal > real(0)) ? real(1) / k_normal : real(0);
            tmass[j] = (k_tangent > real(0)) ? real(1) / k_tangent : real(0);
            ni[j] = warm ? dt_ratio * imp_n[j] : imp_n[j];
            ti[j] = warm ? dt_ratio * imp_t[j] : imp_t[j];
        }
        real k11 = real(0), k12 = real(0), k22 = real(0), nm11 = real(0), nm12 = real(0), nm22 = real(0);
        if (count == 2) {
            const real rn1a = cross(ra[0], normal), rn1b = cross(rb[0], normal);
            const real rn2a = cross(ra[1], normal), rn2b = cross(rb[1], normal);
            k11 = ma + mb + ia * rn1a * rn1a + ib * rn1b * rn1b;
            k22 = ma + mb + ia * rn2a * rn2a + ib * rn2b * rn2b;
            k12 = ma + mb + ia * rn1a * rn2a + ib * rn1b * rn2b;
            if (sq_r(k11) < K<real>::max_condition_number * (k11 * k22 - sq_r(k12))) {      // contact.py:489
                real det = k11 * k22 - k12 * k12;            // Mat22.inverse (common.py:505-519)
                if (det != real(0)) det = real(1) / det;
                nm11 = det * k22; nm12 = -det * k12; nm22 = det * k11;
            } else {
                n_vel = 1;                                   // redundant point dropped from the velocity rows only
            }
        }
        const uint32_t info = (uint32_t)(n_vel | (type << 8) | (count << 16) | (stat_a ? INFO_STATIC_A : 0) | (stat_b ? INFO_STATIC_B : 0));
        const real bodies_r = real_from_bits((uint32_t)ba | ((uint32_t)bb << 16)), info_r = real_from_bits(info);
        const uint32_t vinfo = (uint32_t)n_vel | (stat_a ? VINFO_STATIC_A : 0u) | (stat_b ? VINFO_STATIC_B : 0u) | ((uint32_t)ci << VINFO_CI_SHIFT);
        oct* q = vel_rec(slot);
        oct o;
        o.lo = mkq(normal.x, normal.y, friction, bodies_r);      o.hi = mkq(ra[0].x, ra[0].y, rb[0].x, rb[0].y);          q[0] = o;
        o.lo = mkq(nmass[0], tmass[0], nmass[1], tmass[1]);      o.hi = mkq(ra[1].x, ra[1].y, rb[1].x, rb[1].y);          q[1] = o;
        o.lo = mkq(k11, k12, k22, real_from_bits(vinfo));        o.hi = mkq(nm11, nm12, nm22, real(0));                   q[2] = o;
        o.lo = mkq(ma, ia, mb, ib);                              o.hi = mkq(bias[0], bias[1], real(0), real(0));          q[3] = o;
        *imp_rec(slot) = mkq(ni[0], ti[0], ni[1], ti[1]);
        oct* p = pos_rec(slot);
        o.lo = mkq(ma, ia, mb, ib);
        o.hi = mkq(local_normal.x, local_normal.y, local_point.x, local_point.y);
        p[0] = o;
        o.lo = mkq(lp[0].x, lp[0].y, lp[1].x, lp[1].y);
        o.hi = mkq(rad_a, rad_b, bodies_r, (g_joint[0] != 0) ? real_from_bits(info | ((uint32_t)g_slot_isl[slot] << INFO_ISLAND_SHIFT)) : info_r);
        p[1] = o;
        o.lo = mkq(lca.x, lca.y, lcb.x, lcb.y);                  o.hi = mkq(real(0), real(0), real(0), real(0));
        p[2] = o;
    }

    // ContactSolver.warm_start for one slot (contact.py:498-525)
    B2B_D void constraint_warm_start(const vrec& r)
    {
        const quad q0 = r.q0, q1 = r.q1, q3 = r.q3, q7 = r.q6, qi = r.qi;
        const uint32_t bodies = bits_of_real(q0.w), vinfo = bits_of_real(r.q4.w);
        const int ba = (int)(bodies & 0xffffu), bb = (int)(bodies >> 16), n_vel = (int)(vinfo & VINFO_NVEL_MASK);
        const bool stat_a = (vinfo & VINFO_STATIC_A) != 0, stat_b = (vinfo & VINFO_STATIC_B) != 0;
        const V2 normal = mk<real>(q0.x, q0.y), tangent = right_perp(normal);
        const real ma = q7.x, ia = q7.y, mb = q7.z, ib = q7.w;
        const quad sa = state_get(ba), sb = state_get(bb);
        V2 va = mk<real>(sa.x, sa.y), vb = mk<real>(sb.x, sb.y);
        real wa = sa.z, wb = sb.z;
        for (int j = 0; j < n_vel; ++j) {
            const quad r = (j == 0) ? q1 : q3;
            const V2 ra = mk<real>(r.x, r.y), rb = mk<real>(r.z, r.w);
            const V2 p = ((j == 0) ? qi.x : qi.z) * normal + ((j == 0) ? qi.y : qi.w) * tangent;
            wa = wa - ia * cross(ra, p);
            va = va - ma * p;
            wb = wb + ib * cross(rb, p);
            vb = vb + mb * p;
        }
        if (!stat_a) state_set(ba, va.x, va.y, wa);
        if (!stat_b) state_set(bb, vb.x, vb.y, wb);
    }

    // ContactSolver.solve_velocity_constraints for one slot (contact.py:527-799).  The accumulated impulses go
    // back to the slot's impulse row, or -- on the last velocity iteration -- straight to the contact's manifold
    // (store_impulses, contact.py:801-806: only as many points as the velocity rows kept).
    B2B_D void constraint_solve_velocity(const vrec& r, quad* imp_row, bool last)
    {
        const quad q0 = r.q0, q1 = r.q1, q3 = r.q3, q7 = r.q6;
        quad qi = r.qi;
        const uint32_t bodies = bits_of_real(q0.w), vinfo = bits_of_real(r.q4.w);
        const int ba = (int)(bodies & 0xffffu), bb = (int)(bodies >> 16), n_vel = (int)(vinfo & VINFO_NVEL_MASK);
        const bool stat_a = (vinfo & VINFO_STATIC_A) != 0, stat_b = (vinfo & VINFO_STATIC_B) != 0;
        const V2 normal = mk<real>(q0.x, q0.y), tangent = right_perp(normal);
        const real friction = q0.z;
        const real ma = q7.x, ia = q7.y, mb = q7.z, ib = q7.w;
        const real nmass0 = r.q2.x, tmass0 = r.q2.y, nmass1 = r.q2.z, tmass1 = r.q2.w;
        const real bias0 = r.q7.x, bias1 = r.q7.y;
        const quad sa = state_get(ba), sb = state_get(bb);
        V2 va = mk<real>(sa.x, sa.y), vb = mk<real>(sb.x, sb.y);
        real wa = sa.z, wb = sb.z;
        B2B_TSTAMP_F(2, (float)sa.x, (float)sa.z, (float)sb.x, (float)sb.z);
        const V2 ra0 = mk<real>(q1.x, q1.y), rb0 = mk<real>(q1.z, q1.w), ra1 = mk<real>(q3.x, q3.y), rb1 = mk<real>(q3.z, q3.w);
        // friction rows first, point by point
        {
            const V2 dv = rel_velocity(vb, wb, rb0, va, wa, ra0);
            const real vt = dot(dv, tangent);
            real lambda = tmass0 * (-vt);
            const real max_friction = friction * qi.x;
            const real new_imp = clamp_r(qi.y + lambda, -max_friction, max_friction);
            lambda = new_imp - qi.y;
            const V2 P = lambda * tangent;
            va = va - ma * P; wa = wa - ia * cross(ra0, P);
            vb = vb + mb * P; wb = wb + ib * cross(rb0, P);
            qi.y = new_imp;
        }
        if (n_vel == 2) {
            const V2 dv = rel_velocity(vb, wb, rb1, va, wa, ra1);
            const real vt = dot(dv, tangent);
            real lambda = tmass1 * (-vt);
            const real max_friction = friction * qi.z;
            const real new_imp = clamp_r(qi.w + lambda, -max_friction, max_friction);
            lambda = new_imp - qi.w;
            const V2 P = lambda * tangent;
            va = va - ma * P; wa = wa - ia * cross(ra1, P);
            vb = vb + mb * P; wb = wb + ib * cross(rb1, P);
            qi.w = new_imp;
        }
        if (n_vel == 1) {
            const V2 dv = rel_velocity(vb, wb, rb0, va, wa, ra0);
            const real vn = dot(dv, normal);
            real lambda = -nmass0 * (vn - bias0);
            const real new_imp = pymax(qi.x + lambda, real(0));
            lambda = new_imp - qi.x;
            const V2 P = lambda * normal;
            va = va - ma * P; wa = wa - ia * cross(ra0, P);
            vb = vb + mb * P; wb = wb + ib * cross(rb0, P);
            qi.x = new_imp;
        } else {
            // 2x2 LCP by total enumeration; first feasible case wins (contact.py:598-796)
            const V2 a = mk<real>(qi.x, qi.z);
            const V2 dv1 = rel_velocity(vb, wb, rb0, va, wa, ra0);
            const V2 dv2 = rel_velocity(vb, wb, rb1, va, wa, ra1);
            const real vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
            const real k11 = r.q4.x, k12 = r.q4.y, k22 = r.q4.z;
            V2 b = mk<real>(vn1 - bias0, vn2 - bias1);
            b = b - mk<real>(k11 * a.x + k12 * a.y, k12 * a.x + k22 * a.y);
            const real nm11 = r.q5.x, nm12 = r.q5.y, nm22 = r.q5.z;
            V2 x = -mk<real>(nm11 * b.x + nm12 * b.y, nm12 * b.x + nm22 * b.y);
            bool solved = (x.x >= real(0) && x.y >= real(0));
            if (!solved) {
                x = mk<real>(-nmass0 * b.x, real(0));
                solved = (x.x >= real(0)) && (k12 * x.x + b.y >= real(0));
            }
            if (!solved) {
                x = mk<real>(real(0), -nmass1 * b.y);
                solved = (x.y >= real(0)) && (k12 * x.y + b.x >= real(0));
            }
            if (!solved) {
                x = mk<real>(real(0), real(0));
                solved = (b.x >= real(0)) && (b.y >= real(0));
            }
            if (solved) {
                const V2 d = x - a;
                const V2 p1 = d.x * normal, p2 = d.y * normal;
                const V2 p12 = p1 + p2;
                va = va - ma * p12; wa = wa - ia * (cross(ra0, p1) + cross(ra1, p2));
                vb = vb + mb * p12; wb = wb + ib * (cross(rb0, p1) + cross(rb1, p2));
                qi.x = x.x; qi.z = x.y;
            }
        }
        B2B_TSTAMP_F(3, (float)va.x, (float)wa, (float)vb.x, (float)wb);
        if (!stat_a) state_set(ba, va.x, va.y, wa);
        if (!stat_b) state_set(bb, vb.x, vb.y, wb);
        if (!last) {
            *imp_row = qi;
        } else {
            real* imp = c_imp + 4 * (size_t)(vinfo >> VINFO_CI_SHIFT);
            if (n_vel > 1) stq(imp, qi);
            else if (n_vel > 0) { imp[0] = qi.x; imp[1] = qi.y; }
        }
    }

    // One manifold point of ContactSolver.solve_position_constraints (contact_util.py:323-360 PositionSolverManifold,
    // contact.py:826-860): evaluate the separation at the bodies' CURRENT sweeps and push them apart.
    struct pos_state { V2 ca, cb; real aa, ab, min_sep; };
    B2B_D void position_point(pos_state& p, V2 lpj, int type, V2 local_normal, V2 local_point, V2 lca, V2 lcb,
                              real rad_a, real rad_b, real ma, real ia, real mb, real ib, real baumgarte = K<real>::baumgarte)
    {
        const XF xa = xf_from_sweep<real>(p.ca, p.aa, lca), xb = xf_from_sweep<real>(p.cb, p.ab, lcb);
        V2 normal, point;
        real separation;
        if (type == MANIFOLD_CIRCLES) {
            const V2 pa = xf_mul(xa, local_point), pb = xf_mul(xb, lpj);
            normal = normalize(pb - pa);
            point = real(0.5) * (pa + pb);
            separation = dot(pb - pa, normal) - rad_a - rad_b;
        } else {
            const bool face_a = (type == MANIFOLD_FACE_A);
            const XF xr = face_a ? xa : xb, xo = face_a ? xb : xa;
            normal = rot_mul(xr, local_normal);
            const V2 plane = xf_mul(xr, local_point);
            const V2 clip = xf_mul(xo, lpj);
            separation = dot(clip - plane, normal) - rad_a - rad_b;
            point = clip;
            if (!face_a) normal = -normal;
        }
        const V2 ra = point - p.ca, rb = point - p.cb;
        p.min_sep = pymin(p.min_sep, separation);
        const real C = clamp_r(baumgarte * (separation + K<real>::linear_slop), -K<real>::max_linear_correction, real(0));
        const real rn_a = cross(ra, normal), rn_b = cross(rb, normal);
        const real Keff = ma + mb + (ia * sq_r(rn_a)) + (ib * sq_r(rn_b));      // contact.py:841-842
        const real impulse = (Keff > real(0)) ? -C / Keff : real(0);
        const V2 P = impulse * normal;
        p.ca = p.ca - ma * P; p.aa = p.aa - ia * cross(ra, P);
        p.cb = p.cb + mb * P; p.ab = p.ab + ib * cross(rb, P);
    }

    // ContactSolver.solve_position_constraints for one slot (contact.py:809-864); returns the smallest separation
    // seen.  A touching contact has one or two points.
    B2B_D real constraint_solve_position(const prec& r)
    {
        const quad q7 = r.q7, q8 = r.q8, q9 = r.q9, q10 = r.q10, q11 = r.q11;
        const uint32_t bodies = bits_of_real(q10.z), info = bits_of_real(q10.w);
        const int ba = (int)(bodies & 0xffffu), bb = (int)(bodies >> 16);
        const int type = (int)((info >> 8) & 0xff), n_pos = (int)((info >> 16) & 0xff);
        const bool stat_a = (info & INFO_STATIC_A) != 0, stat_b = (info & INFO_STATIC_B) != 0;
        const real ma = q7.x, ia = q7.y, mb = q7.z, ib = q7.w;
        const V2 lca = mk<real>(q11.x, q11.y), lcb = mk<real>(q11.z, q11.w);
        const V2 local_normal = mk<real>(q8.x, q8.y), local_point = mk<real>(q8.z, q8.w);
        const real rad_a = q10.x, rad_b = q10.y;
        const quad sa = state_get(ba), sb = state_get(bb);
        pos_state p;
        p.ca = mk<real>(sa.x, sa.y); p.cb = mk<real>(sb.x, sb.y); p.aa = sa.z; p.ab = sb.z; p.min_sep = real(0);
        position_point(p, mk<real>(q9.x, q9.y), type, local_normal, local_point, lca, lcb, rad_a, rad_b, ma, ia, mb, ib);
        if (n_pos > 1)
            position_point(p, mk<real>(q9.z, q9.w), type, local_normal, local_point, lca, lcb, rad_a, rad_b, ma, ia, mb, ib);
        if (!stat_a) state_set(ba, p.ca.x, p.ca.y, p.aa);
        if (!stat_b) state_set(bb, p.cb.x, p.cb.y, p.ab);
        return p.min_sep;
    }

    // island.py:71-95 for one body, in the flat prepare kernel: an island body that is not static saves its sweep
    // start and integrates gravity / forces / damping; every body's starting velocity goes to the world's staging row
    // (the velocity kernel copies it to shared memory).
    B2B_D void prepare_body(const step_params& sp, int b)
    {
        uint32_t fl = b_flags[b];
        V2 v;
        real w;
        // world.py:520: the island walk wakes every body it visits.  A walk that was skipped because its lists are
        // still valid (sleep_invalidates) has one effect left: the sleeper it would have reached is awake again.
        if ((fl & BODY_ISLAND) && !(fl & BODY_AWAKE)) {
            fl |= BODY_AWAKE;
            b_flags[b] = fl;
            b_pos[4 * (size_t)b + 3] = real(0);
        }
        if ((fl & BODY_ISLAND) && (fl & BODY_TYPE_MASK) != BODY_STATIC) {
            const quad p = ldq(b_pos + 4 * (size_t)b);
            real* p0 = b_pos0 + 4 * (size_t)b;
            p0[0] = p.x; p0[1] = p.y; p0[2] = p.z;
            integrated_velocity(b, (real)sp.dt, mk<real>(w_gravity[0], w_gravity[1]), v, w);
        } else {
            const quad bv = ldq(b_vel + 4 * (size_t)b);
            v = mk<real>(bv.x, bv.y); w = bv.z;
        }
        real* o = cc_state + 3 * (size_t)b;
        o[0] = v.x; o[1] = v.y; o[2] = w;
    }
    // island.py:119-138 for one non-static body: clamp the step, advance the sweep.  The final velocity and the
    // advanced sweep go to global memory (the position kernel stages the sweep again).
    B2B_D void integrate_position(int b, real dt, quad pos)
    {
        const quad sv = state_get(b);
        const real p[3] = {pos.x, pos.y, pos.z};
        V2 v = mk<real>(sv.x, sv.y);
        real w = sv.z;
        const V2 translation = dt * v;
        if (dot(translation, translation) > K<real>::max_translation_sqr) {     // island.py:123-124: dot, then .length (** 2)
            const real ratio = K<real>::max_translation / sqrt_r(length_sqr(translation));
            v = ratio * v;
        }
        const real rotation = dt * w;
        if (sq_r(rotation) > K<real>::max_rotation_sqr) {                       // island.py:128
            const real ratio = K<real>::max_rotation / abs_r(rotation);
            w = w * ratio;
        }
        real* pout = b_pos + 4 * (size_t)b;
        const real nx = p[0] + dt * v.x, ny = p[1] + dt * v.y, na = p[2] + dt * w;
        pout[0] = nx; pout[1] = ny; pout[2] = na;
        real* vout = b_vel + 4 * (size_t)b;
        vout[0] = v.x; vout[1] = v.y; vout[2] = w;
    }
    // island.py:155-159 + Body._synchronize_transform (body.py:906-908) for one non-static body whose solved sweep
    // is in shared memory
    B2B_D void write_back(int b, quad mass)
    {
        const quad sp = state_get(b);
        real* p = b_pos + 4 * (size_t)b;
        p[0] = sp.x; p[1] = sp.y; p[2] = sp.z;
        const XF tf = xf_from_sweep<real>(mk<real>(sp.x, sp.y), sp.z, mk<real>(mass.z, mass.w));
        stq(b_xf + 4 * (size_t)b, mkq(tf.p.x, tf.p.y, tf.c1.x, tf.c1.y));
    }
    // A body the sleep rule put to sleep is no seed for the next island walk (world.py:499).  If it was its island's
    // seed (island index 0) the next walk differs and the lists must be rebuilt.  Otherwise the walk starts from the
    // same seed, follows the same touching contacts in the same order and reaches the sleeper exactly where it was
    // (membership and order only depend on awake flags through the seeds, world.py:496-546), wakes it (:520) and
    // produces the lists the plan rows already hold: they stay valid, and prepare_body does the waking.  This is the
    // reference's periodic event in a settled pile (the buggy rule of island.py:166-179 puts ONE body to sleep every
    // TIME_TO_SLEEP / dt + 1 steps and the next walk wakes it again); without this every world replanned on that step.
    B2B_D void sleep_invalidates(int victim)
    {
        if (b_island[victim] == 0) B2B_ATOMIC_AND(w_flags, ~(uint32_t)WORLD_GRAPH_VALID);
    }

    // Sleep bookkeeping exactly as written in the reference (island.py:163-179): the allow-sleep
    // test reads the LAST island body; only the last non-static body is put to sleep.
    B2B_D void island_sleep(int b0, int b1, real dt, bool position_solved)
    {
        real min_sleep = K<real>::max_float;
        const bool last_allow = (b_flags[l_body[b1 - 1]] & BODY_ALLOW_SLEEP) != 0;
        int victim = -1;
        for (int i = b0; i < b1; ++i) {
            const int b = l_body[i];
            if (is_static_body(b)) continue;
            victim = b;
            const real* sv = b_vel + 4 * (size_t)b;
            real* st = b_pos + 4 * (size_t)b + 3;
            if (!last_allow || sq_r(sv[2]) > K<real>::angular_sleep_tol_sqr || sq_r(sv[0]) + sq_r(sv[1]) > K<real>::linear_sleep_tol_sqr) {   // island.py:169-170
                *st = real(0);
                min_sleep = real(0);
            } else {
                const real ts = *st + dt;
                *st = ts;
                min_sleep = pymin(min_sleep, ts);
            }
        }
        if (min_sleep >= K<real>::time_to_sleep && position_solved && victim >= 0) {
            // Body.awake = False (body.py:805-812)
            b_pos[4 * (size_t)victim + 3] = real(0);
            real* v = b_vel + 4 * (size_t)victim;
            v[0] = real(0); v[1] = real(0); v[2] = real(0); v[3] = real(0);
            b_force[2 * (size_t)victim] = real(0); b_force[2 * (size_t)victim + 1] = real(0);
            b_flags[victim] &= ~BODY_AWAKE;
            sleep_invalidates(victim);
        }
    }

    // The same rule evaluated by the whole team: the per-body timer updates are independent, the island only needs
    // their minimum and the LAST non-static body of the list.
    B2B_D void island_sleep_team(int b0, int b1, real dt, bool position_solved)
    {
        real min_sleep = K<real>::max_float;
        const bool last_allow = (b_flags[l_body[b1 - 1]] & BODY_ALLOW_SLEEP) != 0;
        int last_i = -1;
        for (int i0 = b0 + t_id; i0 < b1; i0 += BODY_BATCH * t_n) {
            int bb[BODY_BATCH];
            uint32_t fl[BODY_BATCH];
            quad sv[BODY_BATCH];
            real ts0[BODY_BATCH];
#pragma unroll
            for (int k = 0; k < BODY_BATCH; ++k) {
                const int i = i0 + k * t_n;
                bb[k] = (i < b1) ? (int)l_body[i] : -1;
            }
#pragma unroll
            for (int k = 0; k < BODY_BATCH; ++k) {
                if (bb[k] < 0) continue;
                fl[k] = b_flags[bb[k]];
                sv[k] = ldq(b_vel + 4 * (size_t)bb[k]);
                ts0[k] = b_pos[4 * (size_t)bb[k] + 3];
            }
#pragma unroll
            for (int k = 0; k < BODY_BATCH; ++k) {
                if (bb[k] < 0 || (fl[k] & BODY_TYPE_MASK) == BODY_STATIC) continue;
                last_i = i0 + k * t_n;
                real* st = b_pos + 4 * (size_t)bb[k] + 3;
                if (!last_allow || sq_r(sv[k].z) > K<real>::angular_sleep_tol_sqr ||
                    sq_r(sv[k].x) + sq_r(sv[k].y) > K<real>::linear_sleep_tol_sqr) {
                    *st = real(0);
                    min_sleep = real(0);
                } else {
                    const real ts = ts0[k] + dt;
                    *st = ts;
                    min_sleep = pymin(min_sleep, ts);
                }
            }
        }
#if defined(__CUDA_ARCH__)
        for (int o = t_n >> 1; o > 0; o >>= 1) {
            const real other = __shfl_xor_sync(t_mask, min_sleep, o);
            const int other_i = __shfl_xor_sync(t_mask, last_i, o);
            min_sleep = (other < min_sleep) ? other : min_sleep;
            last_i = (other_i > last_i) ? other_i : last_i;
        }
#endif
        if (t_id == 0 && min_sleep >= K<real>::time_to_sleep && position_solved && last_i >= 0) {
            const int victim = l_body[last_i];
            b_pos[4 * (size_t)victim + 3] = real(0);
            real* v = b_vel + 4 * (size_t)victim;
            v[0] = real(0); v[1] = real(0); v[2] = real(0); v[3] = real(0);
            b_force[2 * (size_t)victim] = real(0); b_force[2 * (size_t)victim + 1] = real(0);
            b_flags[victim] &= ~BODY_AWAKE;
            sleep_invalidates(victim);
        }
    }

    // prepare kernel: the constraint records of every island slot of this world (one call per slot, any order)
    B2B_D int prepare_count() const { return w_island[2]; }
    B2B_D int prepare_bodies() const { return w_count[0]; }
    B2B_D void prepare_slot(const step_params& sp, bool first_step, int slot)
    {
        constraint_init(slot, (real)sp.dt, (real)(first_step ? sp.dt_ratio_first : sp.dt_ratio_rest));
    }

    // Island.solve for every island of the world (island.py:65-179), in two kernels.
    //   velocity kernel: integrate velocities, warm start, velocity iterations, store impulses, integrate
    //                    positions (island.py:71-138); islands without contacts are finished here, one lane each
    //   position kernel: position iterations with early exit, write-back, sleep (island.py:141-179)
    // Islands with contacts are taken one after the other by the whole team, their constraints spread over the
    // team's lanes level-step by level-step (see phase_schedule, level_sweep).  Between the kernels the bodies'
    // advanced sweeps live in b_pos and their final velocities in b_vel.
    B2B_D void island_range(int isl, int n_islands, int n_lb, int n_lc, int& b0, int& b1, int& c0, int& c1) const
    {
        b0 = l_start[2 * isl]; c0 = l_start[2 * isl + 1];
        b1 = (isl + 1 < n_islands) ? l_start[2 * isl + 2] : n_lb;
        c1 = (isl + 1 < n_islands) ? l_start[2 * isl + 3] : n_lc;
    }
    // PJ: the batch has a joint capacity (b2b_caps::nj > 0): an island's joints are initialised after the contacts' warm start and
    // visited before the contacts in every velocity iteration (island.py:107-113), by the team's first lane (b2b_joints.inl)
    // (PJ = 2: every joint of the batch is a revolute joint -- b2b_caps::joint_types -- and the type dispatch compiles away)
    template <int PJ> B2B_D void solve_velocity_phase(const step_params& sp, real dt)
    {
        const bool warm = (w_flags[0] & WORLD_WARM_STARTING) != 0;
        const bool profiling = (sp.flags & STEP_PROFILE) != 0;
        long long t = B2B_CLOCK();
        const int nb = w_count[0];
        const int n_islands = w_island[0], n_lb = w_island[1], n_lc = w_island[2];
        (void)profiling;
        // every body's starting velocity (prepare_body), body order: coalesced and independent of the island lists
        // (per-body loops of the solve kernels take BODY_BATCH bodies per lane and round and issue every load of a
        // round before its first store: one memory round trip per round instead of one per body)
        for (int b0 = t_id; b0 < nb; b0 += BODY_BATCH * t_n) {
            real v[BODY_BATCH][3];
#pragma unroll
            for (int k = 0; k < BODY_BATCH; ++k) {
                const int b = b0 + k * t_n;
                if (b < nb) { const real* src = cc_state + 3 * (size_t)b; v[k][0] = src[0]; v[k][1] = src[1]; v[k][2] = src[2]; }
            }
#pragma unroll
            for (int k = 0; k < BODY_BATCH; ++k) {
                const int b = b0 + k * t_n;
                if (b < nb) state_set(b, v[k][0], v[k][1], v[k][2]);
            }
        }
        B2B_SYNC();
        B2B_PROF(PROF_S_INIT, t);
        // ---- islands with contacts, one after the other (contact-free islands have nothing to solve)
        const bool joint = g_joint[0] != 0;                // all islands' slots as one schedule (phase_schedule)
        for (int isl = 0; isl < (joint ? 1 : n_islands); ++isl) {
            int b0, b1, c0, c1;
            island_range(isl, n_islands, n_lb, n_lc, b0, b1, c0, c1);
            if (joint) { c0 = 0; c1 = n_lc; }
            int j0 = 0, j1 = 0;
            if constexpr (PJ) island_joint_range(isl, n_islands, j0, j1);
            if (c1 == c0 && j1 == j0) continue;
            const uint32_t* lv = g_lv + lv_offset(c0, isl);
            const vel_stream vat = vel_stream_at(c0 + t_id);
            auto vready = [&](const vrec& e) { return bits_of_real(e.q0.w); };
            if (warm && c1 > c0) level_sweep<vrec>(c0, c1, lv, sp.zero, vat, [&](const vrec& e, const vel_stream&) { constraint_warm_start(e); }, vready);
            B2B_PROF(PROF_S_WARM, t);
            if constexpr (PJ) {
                if (j1 > j0) {
                    B2B_SYNC();
                    const real dt_ratio = (real)(((sp.flags & STEP_FIRST) != 0) ? sp.dt_ratio_first : sp.dt_ratio_rest);
                    joint_sweep(j0, j1, JFETCH_INIT, [&](jrow& r) {
                        if constexpr (PJ == 2) joint_init_impl<true>(r, dt, dt_ratio, warm); else joint_init(r, dt, dt_ratio, warm);
                        return true; });
                }
            }
            for (int it = 0; it < sp.vel_iters; ++it) {
                const bool last = (it + 1 == sp.vel_iters);
                if constexpr (PJ) {
                    if (j1 > j0) {
                        B2B_SYNC();
                        const real inv_dt = (real)(1.0 / sp.dt);                    // TimeStep.inv_dt (world.py:322)
                        joint_sweep(j0, j1, JFETCH_VELOCITY, [&](jrow& r) {
                            if constexpr (PJ == 2) joint_solve_velocity_impl<true>(r, dt, inv_dt); else joint_solve_velocity(r, dt, inv_dt);
                            return true; });
                    }
                }
                if (c1 > c0) level_sweep<vrec>(c0, c1, lv, sp.zero, vat, [&](const vrec& e, const vel_stream& at) { constraint_solve_velocity(e, at.imp, last); }, vready);
            }
            B2B_PROF(PROF_S_VELOCITY, t);
            // -- store_impulses (contact.py:801-806) is part of the last velocity iteration; without velocity
            // iterations the (warm-started) impulses go back as they are
            if (sp.vel_iters == 0) {
                for (int s = c0 + t_id; s < c1; s += t_n) {
                    const quad qi = *imp_rec(s);
                    const int ci = l_contact[g_sched[s]], n_vel = (int)(bits_of_real(vel_rec(s)[2].lo.w) & VINFO_NVEL_MASK);
                    real* imp = c_imp + 4 * (size_t)ci;
                    if (n_vel > 0) { imp[0] = qi.x; imp[1] = qi.y; }
                    if (n_vel > 1) { imp[2] = qi.z; imp[3] = qi.w; }
                }
            }
        }
        B2B_SYNC();
        // island.py:119-138 for every island body that is not static, body order
        for (int b0 = t_id; b0 < nb; b0 += BODY_BATCH * t_n) {
            uint32_t fl[BODY_BATCH];
            quad p[BODY_BATCH];
#pragma unroll
            for (int k = 0; k < BODY_BATCH; ++k) {
                const int b = b0 + k * t_n;
                fl[k] = 0u;
                if (b < nb) { fl[k] = b_flags[b]; p[k] = ldq(b_pos + 4 * (size_t)b); }
            }
#pragma unroll
            for (int k = 0; k < BODY_BATCH; ++k) {
                const int b = b0 + k * t_n;
                if (b < nb && (fl[k] & BODY_ISLAND) && (fl[k] & BODY_TYPE_MASK) != BODY_STATIC) integrate_position(b, dt, p[k]);
            }
        }
        B2B_SYNC();
        B2B_PROF(PROF_S_STORE, t);
    }

    // JOINT_CAPABLE = false: no world of the batch has a joint schedule (the common case: one island per world, or none
    // balanced enough), the island-ordinal and atomic paths compile away.
    // PJ (with JOINT_CAPABLE = false): the batch has a joint capacity; an island's joints are visited after its contacts in every
    // position iteration and the island is solved when both are within tolerance (island.py:141-153).
    template <bool JOINT_CAPABLE, int PJ = 0> B2B_D void solve_position_phase(const step_params& sp, real dt)
    {
        const bool allow_sleep = (w_flags[0] & WORLD_ALLOW_SLEEP) != 0;
        const bool profiling = (sp.flags & STEP_PROFILE) != 0;
        long long t = B2B_CLOCK();
        const int nb = w_count[0];
        const int n_islands = w_island[0], n_lb = w_island[1], n_lc = w_island[2];
        (void)profiling;
        int max_used = 0;
        // every body's sweep (advanced by the velocity kernel for island bodies), body order
        for (int b0 = t_id; b0 < nb; b0 += BODY_BATCH * t_n) {
            quad p[BODY_BATCH];
#pragma unroll
            for (int k = 0; k < BODY_BATCH; ++k) {
                const int b = b0 + k * t_n;
                if (b < nb) p[k] = ldq(b_pos + 4 * (size_t)b);
            }
#pragma unroll
            for (int k = 0; k < BODY_BATCH; ++k) {
                const int b = b0 + k * t_n;
                if (b < nb) state_set(b, p[k].x, p[k].y, p[k].z);
            }
        }
        B2B_SYNC();
        B2B_PROF(PROF_S_STAGE, t);
        // Islands with contacts are solved in GROUPS: one island per group, or (joint schedule, phase_schedule) one
        // group holding the slots of all islands level-step by level-step.  An island stops iterating once the smallest
        // separation of a sweep is above -3 slop (island.py:141-153); in a joint group its slots are skipped from then
        // on (the slot's island ordinal rides in the record's info word).  A one-island group keeps the sweep's worst
        // separation in a register; a joint group collects max(-separation) per island with integer atomics
        // (separations are <= 0 here: the solver starts every sweep from 0, and positive reals order like their bits)
        // in the world's record slab (the staging row of the starting velocities is free by now).  Both kinds run
        // the same loop, so the worlds of a warp do not diverge when only some of them have a joint schedule.
        if constexpr (!JOINT_CAPABLE) {
            // one island after the other; the sweep's smallest separation stays in a register (this loop's code shape is
            // the measured one: the group loop below, instantiated for one island per group, ran 20 % slower)
            for (int isl = 0; isl < n_islands; ++isl) {
                int b0, b1, c0, c1;
                island_range(isl, n_islands, n_lb, n_lc, b0, b1, c0, c1);
                int j0 = 0, j1 = 0;
                if constexpr (PJ) island_joint_range(isl, n_islands, j0, j1);
                if (c1 == c0 && j1 == j0) continue;
                const uint32_t* lv = g_lv + lv_offset(c0, isl);
                // -- position iterations with early exit (island.py:141-153)
                bool position_solved = false;
                int used = 0;
                for (int it = 0; it < sp.pos_iters; ++it) {
                    real my_min = real(0);
                    ++used;
                    if (c1 > c0)
                        level_sweep<prec>(c0, c1, lv, sp.zero, pos_stream_at(c0 + t_id),
                                          [&](const prec& e, const pos_stream&) { my_min = pymin(my_min, constraint_solve_position(e)); },
                                          [&](const prec& e) { return bits_of_real(e.q10.z); });
                    real all_min = my_min;
#if defined(__CUDA_ARCH__)
                    for (int o = t_n >> 1; o > 0; o >>= 1) {
                        const real other = __shfl_xor_sync(t_mask, all_min, o);
                        all_min = (other < all_min) ? other : all_min;
                    }
#endif
                    int joints_okay = 1;
                    if constexpr (PJ) {
                        if (j1 > j0) {
                            B2B_SYNC();
                            joints_okay = joint_sweep(j0, j1, JFETCH_POSITION, [&](jrow& r) {
                                if constexpr (PJ == 2) return joint_solve_position_impl<true>(r); else return joint_solve_position(r); });
                        }
                    }
                    if (all_min >= real(-3) * K<real>::linear_slop && joints_okay) { position_solved = true; break; }
                }
                if (used > max_used) max_used = used;
                B2B_PROF(PROF_S_POSITION, t);
                if (allow_sleep) island_sleep_team(b0, b1, dt, position_solved);
                B2B_SYNC();
                B2B_PROF(PROF_S_SLEEP, t);
            }
        } else {
            const bool joint = g_joint[0] != 0;
            const int n_groups = joint ? 1 : n_islands;
            const int n_gi = joint ? (int)g_joint[1] : 1;      // islands with contacts in a group
            const uint32_t all = (n_gi >= 32) ? 0xffffffffu : ((1u << n_gi) - 1u);
            typedef typename std::conditional<sizeof(real) == 8, unsigned long long, uint32_t>::type sep_bits;   // real's own width: exact
            sep_bits* worst = reinterpret_cast<sep_bits*>(cc_state);
            const real solved_above = real(-3) * K<real>::linear_slop;
            for (int g = 0; g < n_groups; ++g) {
                int b0, b1, c0, c1;
                island_range(g, n_islands, n_lb, n_lc, b0, b1, c0, c1);
                if (joint) { c0 = 0; c1 = n_lc; }
                if (c1 == c0) continue;
                const uint32_t* lv = g_lv + lv_offset(c0, g);
                uint32_t solved_mask = 0u;                     // islands of the group whose position iterations are over
                int used = 0;
                for (int it = 0; it < sp.pos_iters && solved_mask != all; ++it) {
                    ++used;
                    if (joint) {
                        // every access to the words is an atomic (performed at L2, so no stale L1 line and no fence: a
                        // fence here compiled to MEMBAR.SC.GPU + CCTL.IVALL, emptying the SM's L1 twice per sweep); the
                        // team barrier orders them
                        for (int i = t_id; i < n_gi; i += t_n) B2B_ATOMIC_AND(&worst[i], (sep_bits)0);
                        B2B_SYNC();
                    }
                    const uint32_t skip = solved_mask;
                    real my_neg = real(0);
                    level_sweep<prec>(c0, c1, lv, sp.zero, pos_stream_at(c0 + t_id),
                                      [&](const prec& e, const pos_stream&) {
                                          const int ci = (int)((bits_of_real(e.q10.w) >> INFO_ISLAND_SHIFT) & 31u);   // 0 unless joint
                                          if (joint && ((skip >> ci) & 1u)) return;
                                          const real neg = -constraint_solve_position(e);
                                          my_neg = (neg > my_neg) ? neg : my_neg;            // (used by one-island groups)
                                          if (joint && neg > real(0)) { sep_bits u; memcpy(&u, &neg, sizeof(u)); B2B_ATOMIC_MAX(&worst[ci], u); }
                                      },
                                      [&](const prec& e) { return bits_of_real(e.q10.z); });
                    uint32_t unsolved = 0u;
                    if (joint) {
                        B2B_SYNC();
                        for (int i = t_id; i < n_gi; i += t_n) {
                            const sep_bits u = B2B_ATOMIC_OR(&worst[i], (sep_bits)0);
                            real neg; memcpy(&neg, &u, sizeof(neg));
                            if (!(-neg >= solved_above)) unsolved |= 1u << i;
                        }
                    } else if (!(-my_neg >= solved_above)) {
                        unsolved = 1u;
                    }
#if defined(__CUDA_ARCH__)
                    for (int o = t_n >> 1; o > 0; o >>= 1) unsolved |= __shfl_xor_sync(t_mask, unsolved, o);
#endif
                    solved_mask |= ~unsolved & all;
                    if (joint) B2B_SYNC();
                }
                if (used > max_used) max_used = used;
                B2B_PROF(PROF_S_POSITION, t);
                // the group's islands: sleep bookkeeping (island.py:163-179)
                if (allow_sleep) {
                    for (int isl = joint ? 0 : g, ci = 0; isl < (joint ? n_islands : g + 1); ++isl) {
                        int ib0, ib1, ic0, ic1;
                        island_range(isl, n_islands, n_lb, n_lc, ib0, ib1, ic0, ic1);
                        if (ic1 == ic0) continue;
                        island_sleep_team(ib0, ib1, dt, ((solved_mask >> ci) & 1u) != 0);
                        ++ci;
                        B2B_SYNC();
                    }
                }
                B2B_SYNC();
                B2B_PROF(PROF_S_SLEEP, t);
            }
#if defined(__CUDA_ARCH__) && (!defined(B2B_GROUP_RECONVERGE) || B2B_GROUP_RECONVERGE)
            // The teams of the warp meet again here (every team passes this point once per world): the compiler does not
            // reconverge them after the group loop, and without this each team issued the per-body loops below on its
            // own (ncu: 13 % more warp instructions for the same thread instructions, kernel 29 % longer).
            if (w_mask != 0u) __syncwarp(w_mask);
#endif
        }
        // ---- contact-free islands: an empty constraint list is "solved" by its first position iteration
        // (contact.py:811,864); one lane per island
        B2B_PFOR(isl, n_islands) {
            int b0, b1, c0, c1;
            island_range(isl, n_islands, n_lb, n_lc, b0, b1, c0, c1);
            if (c1 != c0) continue;
            if constexpr (PJ) { int j0, j1; island_joint_range(isl, n_islands, j0, j1); if (j1 != j0) continue; }   // solved above
            if (sp.pos_iters > 0 && max_used < 1) max_used = 1;
            if (allow_sleep) island_sleep(b0, b1, dt, sp.pos_iters > 0);
        }
        B2B_SYNC();
        // island.py:155-159 + Body._synchronize_transform for every island body that is not static, body order
        for (int b0 = t_id; b0 < nb; b0 += BODY_BATCH * t_n) {
            uint32_t fl[BODY_BATCH];
            quad mass[BODY_BATCH];
#pragma unroll
            for (int k = 0; k < BODY_BATCH; ++k) {
                const int b = b0 + k * t_n;
                fl[k] = 0u;
                if (b < nb) { fl[k] = b_flags[b]; mass[k] = ldq(b_mass + 4 * (size_t)b); }
            }
#pragma unroll
            for (int k = 0; k < BODY_BATCH; ++k) {
                const int b = b0 + k * t_n;
                if (b < nb && (fl[k] & BODY_ISLAND) && (fl[k] & BODY_TYPE_MASK) != BODY_STATIC) write_back(b, mass[k]);
            }
        }
        B2B_PROF(PROF_S_WRITEBACK, t);
        // max position iterations used by any island of this world (diagnostic, w_island[3])
#if defined(__CUDA_ARCH__)
        if (max_used > 0) atomicMax(&w_island[3], max_used);
#else
        if (max_used > w_island[3]) w_island[3] = max_used;
#endif
        B2B_SYNC();
    }

    // ------------------------------------------------------------------------------------ K4
    // Body._synchronize_fixtures -> Fixture._synchronize -> BroadPhase.move_proxy -> DynamicTree.move_proxy
    // (body.py:898-904, fixture.py:287-298, broadphase.py:61-68, dynamictree.py:152-188)
    // one proxy; returns true when it went into the move buffer
    B2B_D bool sync_fixture(int f)
    {
        const int b = f_body[f];
        const uint32_t fl = b_flags[b];
        if (!(fl & BODY_ISLAND) || (fl & BODY_TYPE_MASK) == BODY_STATIC) return false;
        return sync_fixture_of(f, b);
    }
    B2B_D bool sync_fixture_of(int f, int b)
    {
        const quad p0 = ldq(b_pos0 + 4 * (size_t)b), m = ldq(b_mass + 4 * (size_t)b);
        const XF xf1 = xf_from_sweep<real>(mk<real>(p0.x, p0.y), p0.z, mk<real>(m.z, m.w));
        const XF xf2 = xf_load<real>(b_xf + 4 * (size_t)b);
        const shape_view<real> s = shape_of_fixture(f);
        real a1[4], a2[4];
        shape_aabb(s, xf1, a1);
        shape_aabb(s, xf2, a2);
        real* tight = f_tight + 4 * (size_t)f; real* fat = f_fat + 4 * (size_t)f;
        const real tlx = pymin(a1[0], a2[0]), tly = pymin(a1[1], a2[1]);
        const real tux = pymax(a1[2], a2[2]), tuy = pymax(a1[3], a2[3]);
        stq(tight, mkq(tlx, tly, tux, tuy));
        const quad fq = ldq(fat);
        if (fq.x <= tlx && fq.y <= tly && tux <= fq.z && tuy <= fq.w) return false;   // AABB.contains
        const real ext = K<real>::aabb_extension;
        real lx = tlx - ext, ly = tly - ext, ux = tux + ext, uy = tuy + ext;
        const V2 d = K<real>::aabb_multiplier * (xf2.p - xf1.p);
        if (d.x < real(0)) lx = lx + d.x; else ux = ux + d.x;
        if (d.y < real(0)) ly = ly + d.y; else uy = uy + d.y;
        stq(fat, mkq(lx, ly, ux, uy));
        f_moved[f] = 1;
        return true;
    }
    // flat sync kernel: one call per (world, index); index < nf synchronises a proxy, index < nb clears a body's
    // forces (World.clear_forces, world.py:356-369, when auto_clear_forces is set)
    B2B_D void sync_item(const step_params& sp, int i)
    {
        if (sp.dt > 0.0 && i < w_count[1] && sync_fixture(i)) {
            if (!(*g_summary & SUM_MOVED)) B2B_ATOMIC_OR(g_summary, (uint32_t)SUM_MOVED);
        }
        if (i < w_count[0] && (w_flags[0] & WORLD_AUTO_CLEAR)) {
            b_force[2 * (size_t)i] = real(0); b_force[2 * (size_t)i + 1] = real(0); b_vel[4 * (size_t)i + 3] = real(0);
        }
    }
    B2B_D int sync_items() const { return w_count[1] > w_count[0] ? w_count[1] : w_count[0]; }

    // ------------------------------------------------------------------------------------ K5 + K6
    // Pair (a, b) is emitted by mover a for every other proxy b whose fat AABB overlaps, except when b is
    // itself an earlier mover (then (b, a) is already in the pair buffer): broadphase.py:83-109 with the
    // emission order pinned to (a, b) ascending (SURVEY §0 finding 1).
    B2B_D bool pair_candidate(int a, const real* fat_a, int b) const
    {
        if (b == a) return false;
        if (!aabb_overlap<real>(s_fat + 4 * (size_t)b, fat_a)) return false;
        if (b < a && s_moved[b]) return false;
        return true;
    }

    // should a new contact be created for pair (fa, fb)?  contact_manager.py:102-158
    B2B_D bool pair_accept(int fa, int fb, bool have_adjacency)
    {
        const int ba = f_body[fa], bb = f_body[fb];
        if (ba == bb) return false;
        const int nc = w_count[2];
        const int ta = body_type(ba), tb = body_type(bb);
        if (tb != BODY_DYNAMIC && ta != BODY_DYNAMIC) return false;              // Body._should_collide
        bool exists = false;
        if (have_adjacency) {
            const int bs = (tb != BODY_STATIC) ? bb : ba;                          // a row that was built
            for (int k = g_adj_start[bs]; k < g_adj_start[bs + 1] && !exists; ++k) {
                const int c = g_adj[k];
                const int ca = c_fix[2 * c], cb = c_fix[2 * c + 1];
                exists = (ca == fa && cb == fb) || (ca == fb && cb == fa);
            }
        } else {
            // (four rows per round, loaded before they are looked at: the loop is a chain of dependent loads otherwise)
            for (int c = 0; c < nc && !exists; c += 4) {
                int ca[4], cb[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { const int ck = (c + k < nc) ? c + k : c; ca[k] = c_fix[2 * ck]; cb[k] = c_fix[2 * ck + 1]; }
#pragma unroll
                for (int k = 0; k < 4; ++k) exists = exists || (ca[k] == fa && cb[k] == fb) || (ca[k] == fb && cb[k] == fa);
            }
        }
        if (exists) return false;
        // Body._should_collide (body.py:920-926): a joint between the two bodies that does not let them collide
        if (cap.nj > 0) {
            for (int k = b_jadj[2 * bb], ke = k + b_jadj[2 * bb + 1]; k < ke; ++k) {
                const int32_t* def = j_def + 4 * (size_t)j_adj[k];
                if (((def[1] == bb && def[2] == ba) || (def[2] == bb && def[1] == ba)) && !(def[3] & JF_COLLIDE)) return false;
            }
        }
        // World.default_contact_filter (world.py:958-970)
        const uint32_t fa0 = f_filter[2 * fa], fa1 = f_filter[2 * fa + 1], fb0 = f_filter[2 * fb], fb1 = f_filter[2 * fb + 1];
        const int ga = (int)(int16_t)(fa1 & 0xffff), gb = (int)(int16_t)(fb1 & 0xffff);
        if (ga == gb && ga != 0) { if (!(ga > 0)) return false; }
        else if (!(((fa0 >> 16) & (fb0 & 0xffff)) != 0 && ((fa0 & 0xffff) & (fb0 >> 16)) != 0)) return false;
        // Contact registry (contact.py:294-311): polygon-polygon, circle-circle, polygon/edge-circle, edge-polygon
        const int t1 = s_meta[f_shape[fa]] & 0xff, t2 = s_meta[f_shape[fb]] & 0xff;
        if (t1 == SHAPE_EDGE && t2 == SHAPE_EDGE) return false;
        if (t1 == SHAPE_LOOP || t2 == SHAPE_LOOP) return false;
        return true;
    }

    B2B_D void phase_pairs(bool have_adjacency)
    {
        const int nf = w_count[1];
        // the movers x proxies tests read every fat box nf times: stage boxes and move flags in shared memory
        B2B_PFOR(f, nf) {
            const quad q = ldq(f_fat + 4 * (size_t)f);
            real* d = s_fat + 4 * (size_t)f;
            d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
            s_moved[f] = (uint8_t)(f_moved[f] != 0);
        }
        B2B_SYNC();
        // pass 1: per-mover candidate counts (and, when the bitmap fits, the candidates themselves: one bit per proxy)
        const int words = (nf + 31) >> 5;
        B2B_PFOR(a, nf) {
            int n = 0;
            if (s_moved[a]) {
                const real fat_a[4] = {s_fat[4 * a], s_fat[4 * a + 1], s_fat[4 * a + 2], s_fat[4 * a + 3]};
                if (s_pairbits) {
                    for (int wi = 0; wi < words; ++wi) {
                        uint32_t m = 0u;
                        const int b0 = wi << 5, b1 = (b0 + 32 < nf) ? b0 + 32 : nf;
                        for (int b = b0; b < b1; ++b) m |= pair_candidate(a, fat_a, b) ? (1u << (b - b0)) : 0u;
                        s_pairbits[(size_t)a * words + wi] = m;
                        n += B2B_POPC(m);
                    }
                } else {
                    for (int b = 0; b < nf; ++b) n += pair_candidate(a, fat_a, b) ? 1 : 0;
                }
            }
            scan[a] = n;
        }
        exclusive_scan(scan, nf, 8);
        const int npairs = s_ctl[8];
        // pass 2: emit in (a, b) order
        B2B_PFOR(a, nf) {
            if (!s_moved[a]) continue;
            int o = scan[a];
            if (s_pairbits) {
                for (int wi = 0; wi < words; ++wi) {
                    uint32_t m = s_pairbits[(size_t)a * words + wi];
                    while (m) {
                        const int b = (wi << 5) + B2B_CTZ(m);
                        m &= m - 1u;
                        if (o < cap.npair) { p_pairs[2 * o] = a; p_pairs[2 * o + 1] = b; }
                        ++o;
                    }
                }
                continue;
            }
            const real fat_a[4] = {s_fat[4 * a], s_fat[4 * a + 1], s_fat[4 * a + 2], s_fat[4 * a + 3]};
            for (int b = 0; b < nf; ++b) {
                if (!pair_candidate(a, fat_a, b)) continue;
                if (o < cap.npair) { p_pairs[2 * o] = a; p_pairs[2 * o + 1] = b; }
                ++o;
            }
        }
        B2B_SYNC();
        const int nuse = (npairs < cap.npair) ? npairs : cap.npair;
        // pass 3: filter, then append the survivors in pair order
        B2B_PFOR(k, nuse) pair_keep[k] = pair_accept(p_pairs[2 * k], p_pairs[2 * k + 1], have_adjacency) ? 1 : 0;
        exclusive_scan(pair_keep, nuse, 9);
        const int n_new = s_ctl[9];
        const int nc0 = w_count[2];
        B2B_PFOR(k, nuse) {
            const bool keep = (k + 1 < nuse) ? (pair_keep[k + 1] != pair_keep[k]) : (n_new != pair_keep[k]);
            if (!keep) continue;
            const int ci = nc0 + pair_keep[k];
            if (ci >= cap.nc) continue;
            int fa = p_pairs[2 * k], fb = p_pairs[2 * k + 1];
            const int t1 = s_meta[f_shape[fa]] & 0xff, t2 = s_meta[f_shape[fb]] & 0xff;
            // registry order: (circle, polygon), (circle, edge), (polygon, edge) are stored swapped (contact.py:321-324)
            const bool swap = (t1 == SHAPE_CIRCLE && (t2 == SHAPE_POLYGON || t2 == SHAPE_EDGE)) || (t1 == SHAPE_POLYGON && t2 == SHAPE_EDGE);
            if (swap) { const int t = fa; fa = fb; fb = t; }
            c_fix[2 * ci] = fa; c_fix[2 * ci + 1] = fb;
            c_flags[ci] = (((f_filter[2 * fa + 1] | f_filter[2 * fb + 1]) >> 16) & 1u) ? (uint32_t)CONTACT_SENSOR : 0u;
            c_mat[2 * ci] = sqrt_r(f_mat[2 * fa] * f_mat[2 * fb]);                 // mix_friction
            c_mat[2 * ci + 1] = pymax(f_mat[2 * fa + 1], f_mat[2 * fb + 1]);       // mix_restitution
            c_mtype[ci] = 0;
            for (int q = 0; q < 4; ++q) { c_mf[4 * (size_t)ci + q] = real(0); c_pts[4 * (size_t)ci + q] = real(0); c_imp[4 * (size_t)ci + q] = real(0); }
            c_id[2 * ci] = 0u; c_id[2 * ci + 1] = 0u;
        }
        B2B_PFOR(a, nf) f_moved[a] = 0;
        B2B_SYNC();
        if (B2B_IS_T0) {
            uint32_t st = 0;
            int nc1 = nc0 + n_new;
            if (npairs > cap.npair) st |= STATUS_PAIR_OVERFLOW;
            if (nc1 > cap.nc) { st |= STATUS_CONTACT_OVERFLOW; nc1 = cap.nc; }
            if (st) w_status[0] |= st;
            if (n_new > 0) w_flags[0] &= ~WORLD_GRAPH_VALID;
            w_count[2] = nc1;
            w_count[3] = npairs;
        }
        B2B_SYNC();
    }


    // ------------------------------------------------------------------------------------ continuous collision (TOI)
    // World.solve_toi (world.py:594-821): find the contact with the earliest time of impact among the contacts of bullets
    // and of non-dynamic bodies, advance its two bodies to that time, solve a small island around them for the rest of the
    // step, synchronise their proxies, find new contacts, repeat until no impact is left inside the step.
    //
    // Work decomposition: the times of impact of the candidate contacts are independent GJK / root-finding problems, one
    // lane each (`toi_evaluate`).  Everything that happens per EVENT is sequential in the reference (advance, narrowphase of
    // the island's contacts, the island's position / velocity sweeps) and a few contacts wide: thread 0 does it.  Proxy
    // synchronisation and pair finding are team-parallel again.  The reference equalises the two sweeps of a contact by
    // advancing the one that lags (a side effect on the BODY, in contact order); once a body has been advanced this step
    // the candidates are therefore snapshotted in list order by thread 0, in batches, and evaluated from the snapshots.
    // The reference's island flags / indices of bodies and contacts are bookkeeping of this loop only; here they live in
    // the scratch words (toi_state, wake_at) so that BODY_ISLAND / CONTACT_ISLAND / b_island keep describing the discrete
    // step's islands (the incremental plan relies on them).
    enum { TS_COUNT_MASK = 15, TS_FLAG = 16, TS_NEED = 32, TS_SEED_A = 64, TS_SEED_B = 128, TS_ISLAND = 256, TS_NEVER = 512 };
    struct toi_cc {
        real ima, imb, iia, iib, friction, restitution, rad_a, rad_b;
        V2 lca, lcb, local_normal, local_point, lp[2], normal;
        V2 ra[2], rb[2];
        real ni[2], ti[2], nmass[2], tmass[2], bias[2];
        real k11, k12, k22, nm11, nm12, nm22;
        int ia, ib, type, n_local, npoints;
    };
    static_assert(sizeof(toi_cc) <= sizeof(real) * TOI_CC_REALS, "toi_layout::off_cc is sized by TOI_CC_REALS");

    B2B_D sweep_t<real> sweep_load(int b) const
    {
        const quad p0 = ldq(b_pos0 + 4 * (size_t)b), p = ldq(b_pos + 4 * (size_t)b), m = ldq(b_mass + 4 * (size_t)b);
        sweep_t<real> s;
        s.lc = mk<real>(m.z, m.w); s.c0 = mk<real>(p0.x, p0.y); s.a0 = p0.z; s.alpha0 = p0.w;
        s.c = mk<real>(p.x, p.y); s.a = p.z;
        return s;
    }
    B2B_D void sweep_store(int b, const sweep_t<real>& s)
    {
        stq(b_pos0 + 4 * (size_t)b, mkq(s.c0.x, s.c0.y, s.a0, s.alpha0));
        real* p = b_pos + 4 * (size_t)b;
        p[0] = s.c.x; p[1] = s.c.y; p[2] = s.a;
    }
    B2B_D void body_sync_transform(int b)                                     // body.py:906-908
    {
        const quad p = ldq(b_pos + 4 * (size_t)b), m = ldq(b_mass + 4 * (size_t)b);
        const XF tf = xf_from_sweep<real>(mk<real>(p.x, p.y), p.z, mk<real>(m.z, m.w));
        stq(b_xf + 4 * (size_t)b, mkq(tf.p.x, tf.p.y, tf.c1.x, tf.c1.y));
    }
    B2B_D void body_advance(int b, real alpha)                                // body.py:930-937
    {
        sweep_t<real> s = sweep_load(b);
        sweep_advance(s, alpha);
        s.c = s.c0; s.a = s.a0;
        sweep_store(b, s);
        body_sync_transform(b);
    }
    B2B_D void body_wake(int b)                                               // body.py:800-804, 812
    {
        const uint32_t fl = b_flags[b];
        if (!(fl & BODY_AWAKE)) { b_pos[4 * (size_t)b + 3] = real(0); b_flags[b] = fl | BODY_AWAKE; }
    }
    // Contact.update with its waking of both bodies on a touching change (contact.py:257-259)
    B2B_D_NOINLINE void toi_contact_update(int c)
    {
        if (contact_update(c)) { body_wake(f_body[c_fix[2 * c]]); body_wake(f_body[c_fix[2 * c + 1]]); }
    }

    // world.py:621-653: is contact c a candidate whose time of impact has to be computed now?  0: no, 1: yes, 2: never in
    // this step (two ordinary dynamic bodies, world.py:651-653: body types and bullet flags do not change inside a step)
    B2B_D static int toi_classify(uint32_t cfl, int st, uint32_t fla, uint32_t flb)
    {
        const int type_a = (int)(fla & BODY_TYPE_MASK), type_b = (int)(flb & BODY_TYPE_MASK);
        const bool collide_a = (fla & BODY_BULLET) || type_a != BODY_DYNAMIC, collide_b = (flb & BODY_BULLET) || type_b != BODY_DYNAMIC;
        if ((!collide_a && !collide_b) || (cfl & CONTACT_SENSOR)) return 2;            // (sensors: world.py:631-633, 757-759)
        if (!(cfl & CONTACT_ENABLED) || (st & TS_COUNT_MASK) > (int)MAX_SUB_STEPS || (st & TS_FLAG)) return 0;
        const bool active_a = (fla & BODY_AWAKE) && type_a != BODY_STATIC, active_b = (flb & BODY_AWAKE) && type_b != BODY_STATIC;
        return (active_a || active_b) ? 1 : 0;
    }
    B2B_D int toi_candidate(int c, int st) const
    {
        return toi_classify(c_flags[c], st, b_flags[f_body[c_fix[2 * c]]], b_flags[f_body[c_fix[2 * c + 1]]]);
    }
    // world.py:670-683: the contact's time of impact on the two (equalised) sweeps
    B2B_D real toi_alpha_of(int c, const sweep_t<real>& sa, const sweep_t<real>& sb, real alpha0)
    {
        const shape_view<real> pa = shape_of_fixture(c_fix[2 * c]), pb = shape_of_fixture(c_fix[2 * c + 1]);
        real beta = real(1);
        bool stuck = false;
        const int state = time_of_impact<real>(pa, pb, sa, sb, real(1), beta, stuck);
        real alpha = real(1);
        if (state == TOI_TOUCHING) alpha = pymin(alpha0 + (real(1) - alpha0) * beta, real(1));
        if (stuck) B2B_ATOMIC_OR(w_status, (uint32_t)STATUS_TOI_STUCK);
        return alpha;
    }
    B2B_D void toi_evaluate(int c, const sweep_t<real>& sa, const sweep_t<real>& sb, real alpha0)
    {
        toi_alpha[c] = toi_alpha_of(c, sa, sb, alpha0);
        toi_state[c] = (toi_state[c] | TS_FLAG) & ~TS_NEED;
    }
    // out[0..n_out) = the indices c < n with (toi_state[c] & bit), ascending; returns n_out (the same value on every lane).
    // Lane teams of a warp compact with ballots, 32 (or t_n) consecutive indices at a time; the sequential build just loops.
    B2B_D int toi_compact(int n, int bit, int32_t* out)
    {
        int count = 0;
#if defined(__CUDA_ARCH__)
        if (!t_cta) {
            const int shift = __ffs((int)t_mask) - 1;
            for (int base = 0; base < n; base += t_n) {
                const int c = base + t_id;
                const bool f = (c < n) && (toi_state[c] & bit);
                const unsigned m = (__ballot_sync(t_mask, f) & t_mask) >> shift;
                if (f) out[count + __popc(m & ((1u << t_id) - 1u))] = c;
                count += __popc(m);
            }
            B2B_SYNC();
            return count;
        }
#endif
        B2B_SYNC();
        if (B2B_IS_T0) {
            for (int c = 0; c < n; ++c) if (toi_state[c] & bit) out[count++] = c;
            s_ctl[13] = count;
        }
        B2B_SYNC();
        count = s_ctl[13];
        B2B_SYNC();
        return count;
    }
    // The team's (smallest value, lowest index among equals) of every lane's (a, ci); ci < 0: the lane has none.  Returns the
    // index (or -1) and leaves the value in toi_red_a[0].
    B2B_D int toi_reduce_min(real a, int ci)
    {
        toi_red_a[t_id & 31] = a; toi_red_i[t_id & 31] = ci;
        B2B_SYNC();
        if (B2B_IS_T0) {
            const int lanes = t_n < 32 ? t_n : 32;
            for (int l = 1; l < lanes; ++l) {
                const real al = toi_red_a[l];
                const int cl = toi_red_i[l];
                if (cl >= 0 && (ci < 0 || al < a || (al == a && cl < ci))) { a = al; ci = cl; }
            }
            toi_red_a[0] = a;
            s_ctl[13] = ci;
        }
        B2B_SYNC();
        const int r = s_ctl[13];
        B2B_SYNC();
        return r;
    }

    // ContactSolver.__init__ for the sub-island's contact k (contact.py:376-412; no warm-start scaling: step.warm_starting is False)
    B2B_D void toi_constraint_ctor(toi_cc& cc, int ci)
    {
        const int fa = c_fix[2 * ci], fb = c_fix[2 * ci + 1];
        const int ba = f_body[fa], bb = f_body[fb];
        const quad mass_a = ldq(b_mass + 4 * (size_t)ba), mass_b = ldq(b_mass + 4 * (size_t)bb);
        cc.ia = wake_at[ba]; cc.ib = wake_at[bb];
        cc.ima = mass_a.x; cc.imb = mass_b.x; cc.iia = mass_a.y; cc.iib = mass_b.y;
        cc.lca = mk<real>(mass_a.z, mass_a.w); cc.lcb = mk<real>(mass_b.z, mass_b.w);
        cc.friction = c_mat[2 * ci]; cc.restitution = c_mat[2 * ci + 1];
        cc.rad_a = s_radius[f_shape[fa]]; cc.rad_b = s_radius[f_shape[fb]];
        const int mt = c_mtype[ci];
        const quad mf = ldq(c_mf + 4 * (size_t)ci), pts = ldq(c_pts + 4 * (size_t)ci), imp = ldq(c_imp + 4 * (size_t)ci);
        cc.type = mt & 0xff; cc.n_local = (mt >> 8) & 0xff; cc.npoints = cc.n_local;
        cc.local_normal = mk<real>(mf.x, mf.y); cc.local_point = mk<real>(mf.z, mf.w);
        cc.lp[0] = mk<real>(pts.x, pts.y); cc.lp[1] = mk<real>(pts.z, pts.w);
        cc.ni[0] = imp.x; cc.ti[0] = imp.y; cc.ni[1] = imp.z; cc.ti[1] = imp.w;
        cc.normal = mk<real>(real(0), real(0));
        for (int j = 0; j < 2; ++j) {
            cc.ra[j] = cc.rb[j] = mk<real>(real(0), real(0));
            cc.nmass[j] = cc.tmass[j] = cc.bias[j] = real(0);
        }
        cc.k11 = cc.k12 = cc.k22 = cc.nm11 = cc.nm12 = cc.nm22 = real(0);
    }
    // solve_toi_position_constraints (contact.py:866-933): only the two TOI bodies have mass; true when separated enough
    B2B_D bool toi_position_sweep(int ncontacts, int toi_a, int toi_b)
    {
        real min_separation = real(0);
        for (int k = 0; k < ncontacts; ++k) {
            const toi_cc& cc = toi_C[k];
            pos_state p;
            p.ca = mk<real>(toi_P[3 * cc.ia], toi_P[3 * cc.ia + 1]); p.aa = toi_P[3 * cc.ia + 2];
            p.cb = mk<real>(toi_P[3 * cc.ib], toi_P[3 * cc.ib + 1]); p.ab = toi_P[3 * cc.ib + 2];
            p.min_sep = min_separation;
            real ima = real(0), iia = real(0), imb = real(0), iib = real(0);
            if (cc.ia == toi_a || cc.ia == toi_b) { ima = cc.ima; iia = cc.iia; }
            if (cc.ib == toi_a || cc.ib == toi_b) { imb = cc.imb; iib = cc.iib; }
            for (int j = 0; j < cc.n_local; ++j)
                position_point(p, cc.lp[j], cc.type, cc.local_normal, cc.local_point, cc.lca, cc.lcb, cc.rad_a, cc.rad_b,
                               ima, iia, imb, iib, real(0.75));                       // settings.py:110 TOI_BAUMGARTE
            min_separation = p.min_sep;
            toi_P[3 * cc.ia] = p.ca.x; toi_P[3 * cc.ia + 1] = p.ca.y; toi_P[3 * cc.ia + 2] = p.aa;
            toi_P[3 * cc.ib] = p.cb.x; toi_P[3 * cc.ib + 1] = p.cb.y; toi_P[3 * cc.ib + 2] = p.ab;
        }
        return min_separation >= real(-1.5) * K<real>::linear_slop;
    }
    // initialize_velocity_constraints (contact.py:414-496) on the sub-island's positions and velocities
    B2B_D void toi_init_velocity(toi_cc& cc)
    {
        const real ma = cc.ima, mb = cc.imb, ia = cc.iia, ib = cc.iib;
        const V2 ca = mk<real>(toi_P[3 * cc.ia], toi_P[3 * cc.ia + 1]), cb = mk<real>(toi_P[3 * cc.ib], toi_P[3 * cc.ib + 1]);
        const real aa = toi_P[3 * cc.ia + 2], ab = toi_P[3 * cc.ib + 2];
        const V2 va = mk<real>(toi_V[3 * cc.ia], toi_V[3 * cc.ia + 1]), vb = mk<real>(toi_V[3 * cc.ib], toi_V[3 * cc.ib + 1]);
        const real wa = toi_V[3 * cc.ia + 2], wb = toi_V[3 * cc.ib + 2];
        const XF xa = xf_from_sweep<real>(ca, aa, cc.lca), xb = xf_from_sweep<real>(cb, ab, cc.lcb);
        V2 normal, wp[2];
        world_manifold(cc.type, cc.local_normal, cc.local_point, cc.lp, xa, xb, cc.rad_a, cc.rad_b, normal, wp);
        cc.normal = normal;
        const V2 tangent = right_perp(normal);
        for (int j = 0; j < cc.n_local; ++j) {
            cc.ra[j] = wp[j] - ca; cc.rb[j] = wp[j] - cb;
            const real rn_a = cross(cc.ra[j], normal), rn_b = cross(cc.rb[j], normal);
            const real k_normal = ma + mb + ia * sq_r(rn_a) + ib * sq_r(rn_b);
            const real rt_a = cross(cc.ra[j], tangent), rt_b = cross(cc.rb[j], tangent);
            const real k_tangent = ma + mb + ia * sq_r(rt_a) + ib * sq_r(rt_b);
            cc.nmass[j] = (k_normal > real(0)) ? real(1) / k_normal : real(0);
            cc.tmass[j] = (k_tangent > real(0)) ? real(1) / k_tangent : real(0);
            cc.bias[j] = real(0);
            const real v_rel = dot(normal, rel_velocity(vb, wb, cc.rb[j], va, wa, cc.ra[j]));
            if (v_rel < -K<real>::velocity_threshold) cc.bias[j] = -cc.restitution * v_rel;
        }
        if (cc.npoints == 2) {
            const real rn1a = cross(cc.ra[0], normal), rn1b = cross(cc.rb[0], normal);
            const real rn2a = cross(cc.ra[1], normal), rn2b = cross(cc.rb[1], normal);
            const real k11 = ma + mb + ia * rn1a * rn1a + ib * rn1b * rn1b;
            const real k22 = ma + mb + ia * rn2a * rn2a + ib * rn2b * rn2b;
            const real k12 = ma + mb + ia * rn1a * rn2a + ib * rn1b * rn2b;
            if (sq_r(k11) < K<real>::max_condition_number * (k11 * k22 - sq_r(k12))) {
                cc.k11 = k11; cc.k12 = k12; cc.k22 = k22;
                real det = k11 * k22 - k12 * k12;
                if (det != real(0)) det = real(1) / det;
                cc.nm11 = det * k22; cc.nm12 = -det * k12; cc.nm22 = det * k11;
            } else {
                cc.npoints = 1;
            }
        }
    }
    // solve_velocity_constraints (contact.py:527-799) for one contact of the sub-island
    B2B_D void toi_solve_velocity(toi_cc& cc)
    {
        const real ma = cc.ima, mb = cc.imb, ia = cc.iia, ib = cc.iib;
        const V2 normal = cc.normal, tangent = right_perp(normal);
        V2 va = mk<real>(toi_V[3 * cc.ia], toi_V[3 * cc.ia + 1]), vb = mk<real>(toi_V[3 * cc.ib], toi_V[3 * cc.ib + 1]);
        real wa = toi_V[3 * cc.ia + 2], wb = toi_V[3 * cc.ib + 2];
        for (int j = 0; j < cc.npoints; ++j) {
            const V2 dv = rel_velocity(vb, wb, cc.rb[j], va, wa, cc.ra[j]);
            const real vt = dot(dv, tangent);
            real lambda = cc.tmass[j] * (-vt);
            const real max_friction = cc.friction * cc.ni[j];
            const real new_imp = clamp_r(cc.ti[j] + lambda, -max_friction, max_friction);
            lambda = new_imp - cc.ti[j];
            const V2 P = lambda * tangent;
            va = va - ma * P; wa = wa - ia * cross(cc.ra[j], P);
            vb = vb + mb * P; wb = wb + ib * cross(cc.rb[j], P);
            cc.ti[j] = new_imp;
        }
        if (cc.npoints == 1) {
            const V2 dv = rel_velocity(vb, wb, cc.rb[0], va, wa, cc.ra[0]);
            const real vn = dot(dv, normal);
            real lambda = -cc.nmass[0] * (vn - cc.bias[0]);
            const real new_imp = pymax(cc.ni[0] + lambda, real(0));
            lambda = new_imp - cc.ni[0];
            const V2 P = lambda * normal;
            va = va - ma * P; wa = wa - ia * cross(cc.ra[0], P);
            vb = vb + mb * P; wb = wb + ib * cross(cc.rb[0], P);
            cc.ni[0] = new_imp;
        } else if (cc.npoints == 2) {
            const V2 a = mk<real>(cc.ni[0], cc.ni[1]);
            const V2 dv1 = rel_velocity(vb, wb, cc.rb[0], va, wa, cc.ra[0]);
            const V2 dv2 = rel_velocity(vb, wb, cc.rb[1], va, wa, cc.ra[1]);
            const real vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
            V2 b = mk<real>(vn1 - cc.bias[0], vn2 - cc.bias[1]);
            b = b - mk<real>(cc.k11 * a.x + cc.k12 * a.y, cc.k12 * a.x + cc.k22 * a.y);
            V2 x = -mk<real>(cc.nm11 * b.x + cc.nm12 * b.y, cc.nm12 * b.x + cc.nm22 * b.y);
            bool solved = (x.x >= real(0) && x.y >= real(0));
            if (!solved) {
                x = mk<real>(-cc.nmass[0] * b.x, real(0));
                solved = (x.x >= real(0)) && (cc.k12 * x.x + b.y >= real(0));
            }
            if (!solved) {
                x = mk<real>(real(0), -cc.nmass[1] * b.y);
                solved = (x.y >= real(0)) && (cc.k12 * x.y + b.x >= real(0));
            }
            if (!solved) {
                x = mk<real>(real(0), real(0));
                solved = (b.x >= real(0)) && (b.y >= real(0));
            }
            if (solved) {
                const V2 d = x - a;
                const V2 p1 = d.x * normal, p2 = d.y * normal;
                const V2 p12 = p1 + p2;
                va = va - ma * p12; wa = wa - ia * (cross(cc.ra[0], p1) + cross(cc.ra[1], p2));
                vb = vb + mb * p12; wb = wb + ib * (cross(cc.rb[0], p1) + cross(cc.rb[1], p2));
                cc.ni[0] = x.x; cc.ni[1] = x.y;
            }
        }
        toi_V[3 * cc.ia] = va.x; toi_V[3 * cc.ia + 1] = va.y; toi_V[3 * cc.ia + 2] = wa;
        toi_V[3 * cc.ib] = vb.x; toi_V[3 * cc.ib + 1] = vb.y; toi_V[3 * cc.ib + 2] = wb;
    }
    // Island.solve_toi (island.py:181-243) for the sub-island in toi_ib / toi_ic; thread 0 only
    B2B_D_NOINLINE void toi_island_solve(int nbodies, int ncontacts, real dt, int vel_iters, int toi_a, int toi_b)
    {
        for (int i = 0; i < nbodies; ++i) {
            const int b = toi_ib[i];
            const quad p = ldq(b_pos + 4 * (size_t)b), v = ldq(b_vel + 4 * (size_t)b);
            toi_P[3 * i] = p.x; toi_P[3 * i + 1] = p.y; toi_P[3 * i + 2] = p.z;
            toi_V[3 * i] = v.x; toi_V[3 * i + 1] = v.y; toi_V[3 * i + 2] = v.z;
        }
        for (int k = 0; k < ncontacts; ++k) toi_constraint_ctor(toi_C[k], toi_ic[k]);
        for (int it = 0; it < 20; ++it)                                       // world.py:785 pos_iters = 20
            if (toi_position_sweep(ncontacts, toi_a, toi_b)) break;
        {   // the leap of faith to the new safe state: island.py:196-202
            const int ba = toi_ib[toi_a], bb = toi_ib[toi_b];
            real* pa = b_pos0 + 4 * (size_t)ba; real* pb = b_pos0 + 4 * (size_t)bb;
            pa[0] = toi_P[3 * toi_a]; pa[1] = toi_P[3 * toi_a + 1]; pa[2] = toi_P[3 * toi_a + 2];
            pb[0] = toi_P[3 * toi_b]; pb[1] = toi_P[3 * toi_b + 1]; pb[2] = toi_P[3 * toi_b + 2];
        }
        for (int k = 0; k < ncontacts; ++k) toi_init_velocity(toi_C[k]);
        for (int it = 0; it < vel_iters; ++it)
            for (int k = 0; k < ncontacts; ++k) toi_solve_velocity(toi_C[k]);
        // (the impulses are not stored) integrate: island.py:213-241
        for (int i = 0; i < nbodies; ++i) {
            const int b = toi_ib[i];
            V2 c = mk<real>(toi_P[3 * i], toi_P[3 * i + 1]), v = mk<real>(toi_V[3 * i], toi_V[3 * i + 1]);
            real a = toi_P[3 * i + 2], w = toi_V[3 * i + 2];
            const V2 translation = dt * v;
            if (length_sqr(translation) > K<real>::max_translation_sqr) {       // island.py:219: .length_squared (** 2)
                const real ratio = K<real>::max_translation / sqrt_r(length_sqr(translation));
                v = ratio * v;
            }
            const real rotation = dt * w;
            if (sq_r(rotation) > K<real>::max_rotation_sqr) {
                const real ratio = K<real>::max_rotation / abs_r(rotation);
                w = w * ratio;
            }
            c = c + dt * v;
            a = a + dt * w;
            real* p = b_pos + 4 * (size_t)b; real* vo = b_vel + 4 * (size_t)b;
            p[0] = c.x; p[1] = c.y; p[2] = a;
            vo[0] = v.x; vo[1] = v.y; vo[2] = w;
            body_sync_transform(b);
        }
    }

    // One TOI event up to the solved sub-island (world.py:696-790), thread 0 only.  Returns the number of sub-island
    // bodies (0: the contact was not solid, the sweeps were restored).
    B2B_D int toi_event(const step_params& sp, int mc, real min_alpha, const int32_t* list_a, int n_a, const int32_t* list_b, int n_b)
    {
        const int body_a = f_body[c_fix[2 * mc]], body_b = f_body[c_fix[2 * mc + 1]];
        const sweep_t<real> backup_a = sweep_load(body_a), backup_b = sweep_load(body_b);
        body_advance(body_a, min_alpha);
        body_advance(body_b, min_alpha);
        toi_contact_update(mc);
        toi_state[mc] = ((toi_state[mc] & ~TS_FLAG) & ~TS_COUNT_MASK) | (((toi_state[mc] & TS_COUNT_MASK) + 1) & TS_COUNT_MASK);
        if (!(c_flags[mc] & CONTACT_ENABLED) || !(c_flags[mc] & CONTACT_TOUCHING)) {
            c_flags[mc] &= ~CONTACT_ENABLED;
            sweep_store(body_a, backup_a);
            sweep_store(body_b, backup_b);
            body_sync_transform(body_a);
            body_sync_transform(body_b);
            return 0;
        }
        body_wake(body_a);
        body_wake(body_b);
        int nib = 0, nic = 0;
        wake_at[body_a] = nib; toi_ib[nib++] = body_a;
        wake_at[body_b] = nib; toi_ib[nib++] = body_b;
        toi_ic[nic++] = mc;
        toi_state[mc] |= TS_ISLAND;
        for (int si = 0; si < 2; ++si) {
            const int body = si ? body_b : body_a;
            const int32_t* list = si ? list_b : list_a;                       // body._contacts: the world's list, filtered
            const int n_list = si ? n_b : n_a;
            if (body_type(body) != BODY_DYNAMIC) continue;
            const bool body_bullet = (b_flags[body] & BODY_BULLET) != 0;
            for (int k = 0; k < n_list; ++k) {
                const int c = list[k];
                const int st = toi_state[c];
                if (nib == 2 * (int)MAX_TOI_CONTACTS) break;
                if (nic == (int)MAX_TOI_CONTACTS) break;
                if (st & TS_ISLAND) continue;
                const int cba = f_body[c_fix[2 * c]], cbb = f_body[c_fix[2 * c + 1]];
                const int other = (cba == body) ? cbb : cba;
                if (body_type(other) == BODY_DYNAMIC && !body_bullet && !(b_flags[other] & BODY_BULLET)) continue;
                const sweep_t<real> backup = sweep_load(other);
                if (wake_at[other] < 0) body_advance(other, min_alpha);
                toi_contact_update(c);
                if (!(c_flags[c] & CONTACT_ENABLED) || !(c_flags[c] & CONTACT_TOUCHING)) {
                    sweep_store(other, backup);
                    body_sync_transform(other);
                    continue;
                }
                toi_state[c] = st | TS_ISLAND;
                toi_ic[nic++] = c;
                if (wake_at[other] >= 0) continue;
                if (body_type(other) != BODY_STATIC) body_wake(other);
                wake_at[other] = nib; toi_ib[nib++] = other;
            }
        }
        const real sub_dt = (real(1) - min_alpha) * (real)sp.dt;
        toi_island_solve(nib, nic, sub_dt, sp.vel_iters, wake_at[body_a], wake_at[body_b]);
        return nib;
    }

    B2B_D void phase_toi(const step_params& sp)
    {
        const int nb = w_count[0];
        // world.py:599-609 (the step is always complete here: no sub_stepping).  Sweeps only start later than the step
        // (alpha0 != 0) after a solved event: WORLD_TOI_DIRTY remembers that, so a world without events pays no reset.
        if (w_flags[0] & WORLD_TOI_DIRTY) {
            B2B_PFOR(b, nb) b_pos0[4 * (size_t)b + 3] = real(0);
            B2B_SYNC();
            if (B2B_IS_T0) w_flags[0] &= ~WORLD_TOI_DIRTY;
        }
        // First look, the common case (nothing is about to hit anything): every sweep starts at alpha0 = 0, the candidates are
        // independent, no cached value exists yet.  One pass, four contacts per lane at a time with their row loads batched (a
        // world's rows are cold in this kernel: the dependent loads of one contact after the other were most of its time);
        // nothing is written unless an event is found.
        {
            const int nc = w_count[2];
            real best = real(1);
            int best_c = -1;
            for (int base = 0; base < nc; base += 4 * t_n) {
                int cls[4];
                uint32_t cfl[4];
                int ba[4], bb[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = base + k * t_n + t_id;
                    cfl[k] = 0u; ba[k] = bb[k] = 0;
                    if (c < nc) { cfl[k] = c_flags[c]; ba[k] = c_fix[2 * c]; bb[k] = c_fix[2 * c + 1]; }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) { const int c = base + k * t_n + t_id; if (c < nc) { ba[k] = f_body[ba[k]]; bb[k] = f_body[bb[k]]; } }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = base + k * t_n + t_id;
                    cls[k] = (c < nc) ? toi_classify(cfl[k], 0, b_flags[ba[k]], b_flags[bb[k]]) : 0;
                }
                for (int k = 0; k < 4; ++k) {
                    if (cls[k] != 1) continue;
                    const int c = base + k * t_n + t_id;
                    sweep_t<real> sa = sweep_load(ba[k]), sb = sweep_load(bb[k]);
                    sa.alpha0 = sb.alpha0 = real(0);
                    const real a = toi_alpha_of(c, sa, sb, real(0));
                    if (a < best) { best = a; best_c = c; }
                }
            }
            const int first = toi_reduce_min(best, best_c);
            if (first < 0 || (real(1) - (real(10) * K<real>::epsilon)) < toi_red_a[0]) return;      // world.py:693-696
        }
        // There is an event: the loop proper, with the per-contact words in scratch
        B2B_PFOR(b, nb) wake_at[b] = -1;
        { const int nc = w_count[2]; B2B_PFOR(c, nc) { toi_alpha[c] = real(1); toi_state[c] = 0; } }
        B2B_SYNC();
        bool advanced = false;          // some body's sweep starts later than the step does (after the first solved event)
        for (;;) {
            const int nc = w_count[2];
            B2B_PFOR(c, nc) {
                const int st = toi_state[c];
                if (st & TS_NEVER) continue;
                const int cls = toi_candidate(c, st);
                toi_state[c] = (cls == 2) ? (st | TS_NEVER) : (cls == 1) ? (st | TS_NEED) : (st & ~TS_NEED);
            }
            B2B_SYNC();
            const int n_need = toi_compact(nc, TS_NEED, toi_list);
            if (!advanced) {
                // every sweep starts at alpha0 = 0: nothing to equalise, the candidates are independent
                B2B_PFOR(k, n_need) {
                    const int c = toi_list[k];
                    const sweep_t<real> sa = sweep_load(f_body[c_fix[2 * c]]), sb = sweep_load(f_body[c_fix[2 * c + 1]]);
                    toi_evaluate(c, sa, sb, real(0));
                }
                B2B_SYNC();
            } else {
                for (int k0 = 0; k0 < n_need; k0 += (int)TOI_BATCH) {
                    const int n = (n_need - k0 < (int)TOI_BATCH) ? n_need - k0 : (int)TOI_BATCH;
                    if (B2B_IS_T0) {
                        for (int k = 0; k < n; ++k) {
                            const int c = toi_list[k0 + k];
                            const int ba = f_body[c_fix[2 * c]], bb = f_body[c_fix[2 * c + 1]];
                            sweep_t<real> sa = sweep_load(ba), sb = sweep_load(bb);
                            real alpha0 = sa.alpha0;                          // world.py:657-666
                            if (sa.alpha0 < sb.alpha0) { alpha0 = sb.alpha0; sweep_advance(sa, alpha0); sweep_store(ba, sa); }
                            else if (sb.alpha0 < sa.alpha0) { sweep_advance(sb, alpha0); sweep_store(bb, sb); }
                            real* q = toi_snap + (size_t)TOI_SNAP_REALS * k;
                            q[0] = sa.c0.x; q[1] = sa.c0.y; q[2] = sa.a0; q[3] = sa.c.x; q[4] = sa.c.y; q[5] = sa.a;
                            q[6] = sb.c0.x; q[7] = sb.c0.y; q[8] = sb.a0; q[9] = sb.c.x; q[10] = sb.c.y; q[11] = sb.a;
                            q[12] = alpha0;
                        }
                    }
                    B2B_SYNC();
                    B2B_PFOR(k, n) {
                        const int c = toi_list[k0 + k];
                        const real* q = toi_snap + (size_t)TOI_SNAP_REALS * k;
                        const quad ma = ldq(b_mass + 4 * (size_t)f_body[c_fix[2 * c]]), mb = ldq(b_mass + 4 * (size_t)f_body[c_fix[2 * c + 1]]);
                        sweep_t<real> sa, sb;
                        sa.lc = mk<real>(ma.z, ma.w); sa.c0 = mk<real>(q[0], q[1]); sa.a0 = q[2]; sa.c = mk<real>(q[3], q[4]); sa.a = q[5];
                        sb.lc = mk<real>(mb.z, mb.w); sb.c0 = mk<real>(q[6], q[7]); sb.a0 = q[8]; sb.c = mk<real>(q[9], q[10]); sb.a = q[11];
                        sa.alpha0 = sb.alpha0 = q[12];
                        toi_evaluate(c, sa, sb, q[12]);
                    }
                    B2B_SYNC();
                }
            }
            // the first contact of the list with the smallest time of impact (world.py:688-691: strict <)
            real best = real(1);
            int best_c = -1;
            B2B_PFOR(c, nc) {
                const int st = toi_state[c];
                if (!(st & TS_FLAG) || (st & TS_COUNT_MASK) > (int)MAX_SUB_STEPS) continue;
                if (!(c_flags[c] & CONTACT_ENABLED)) continue;
                const real a = toi_alpha[c];
                if (a < best) { best = a; best_c = c; }
            }
            const int min_contact = toi_reduce_min(best, best_c);
            const real min_alpha = toi_red_a[0];
            B2B_SYNC();
            if (min_contact < 0 || (real(1) - (real(10) * K<real>::epsilon)) < min_alpha) break;

            // the event.  Which contacts touch the two bodies (body._contacts of the sub-island's seeds): in parallel.  Contacts
            // of two ordinary dynamic bodies (TS_NEVER) are never taken into a sub-island (world.py:751-753) and never hold a
            // cached time of impact, so they are skipped here and below.
            const int body_a = f_body[c_fix[2 * min_contact]], body_b = f_body[c_fix[2 * min_contact + 1]];
            B2B_PFOR(c, nc) {
                int st = toi_state[c];
                if (st & TS_NEVER) continue;
                const int cba = f_body[c_fix[2 * c]], cbb = f_body[c_fix[2 * c + 1]];
                st &= ~(TS_SEED_A | TS_SEED_B | TS_ISLAND);
                if (cba == body_a || cbb == body_a) st |= TS_SEED_A;
                if (cba == body_b || cbb == body_b) st |= TS_SEED_B;
                toi_state[c] = st;
            }
            B2B_SYNC();
            int32_t* list_a = toi_list + ((size_t)cap.nc + 2);
            int32_t* list_b = toi_list + 2 * ((size_t)cap.nc + 2);
            const int n_a = toi_compact(nc, TS_SEED_A, list_a), n_b = toi_compact(nc, TS_SEED_B, list_b);
            if (B2B_IS_T0) {
                const int nib = toi_event(sp, min_contact, min_alpha, list_a, n_a, list_b, n_b);
                s_ctl[14] = nib;
                // manifolds, touching flags, awake flags and positions changed under the plan; sweeps now start inside the step
                w_flags[0] = (w_flags[0] & ~WORLD_GRAPH_VALID) | (nib ? (uint32_t)WORLD_TOI_DIRTY : 0u);
            }
            B2B_SYNC();
            const int nib = s_ctl[14];
            B2B_SYNC();
            if (nib == 0) continue;
            advanced = true;
            // world.py:794-806: proxies of the sub-island's dynamic bodies, and the cached times of impact of their contacts
            const int nf = w_count[1];
            if (B2B_IS_T0) s_ctl[14] = 0;
            B2B_SYNC();
            B2B_PFOR(f, nf) {
                const int b = f_body[f];
                if (wake_at[b] >= 0 && body_type(b) == BODY_DYNAMIC && sync_fixture_of(f, b)) s_ctl[14] = 1;   // (benign same-value race)
            }
            B2B_PFOR(c, nc) {
                const int st = toi_state[c];
                if (st & TS_NEVER) continue;
                const int cba = f_body[c_fix[2 * c]], cbb = f_body[c_fix[2 * c + 1]];
                const bool hit = (wake_at[cba] >= 0 && body_type(cba) == BODY_DYNAMIC) || (wake_at[cbb] >= 0 && body_type(cbb) == BODY_DYNAMIC);
                if (hit) toi_state[c] = st & ~(TS_FLAG | TS_ISLAND);
            }
            B2B_SYNC();
            B2B_PFOR(i, nib) wake_at[toi_ib[i]] = -1;
            B2B_SYNC();
            // world.py:810: find_new_contacts (a contact created here is not enabled before the next step's narrowphase).  With
            // an empty move buffer (the usual case: the fat boxes still hold the displaced bodies) it emits no pair.
            const bool any_moved = s_ctl[14] != 0;
            B2B_SYNC();
            if (any_moved) {
                phase_pairs(false);
                { const int nc1 = w_count[2]; B2B_PFOR(c, nc1) if (c >= nc) { toi_alpha[c] = real(1); toi_state[c] = 0; } }
            } else if (B2B_IS_T0) {
                w_count[3] = 0;
            }
            B2B_SYNC();
        }
    }
    B2B_D void run_toi(const step_params& sp)
    {
        if (sp.dt > 0.0 && (w_flags[0] & WORLD_CONTINUOUS)) phase_toi(sp);
    }

    // ------------------------------------------------------------------------------------ World.step
    // plan kernel: (find_new_contacts if new fixtures) + collide + island construction + level schedule
    B2B_D void run_plan(const step_params& sp)
    {
        const bool profiling = (sp.flags & STEP_PROFILE) != 0;
        long long t = B2B_CLOCK();
        // worlds with new fixtures were skipped by the flat narrowphase kernel (their contacts are created only now)
        const bool after_flat = (sp.flags & STEP_FLAT_COLLIDE) && !(w_flags[0] & WORLD_NEW_FIXTURE);
        if ((w_flags[0] & WORLD_NEW_FIXTURE) && !(sp.flags & STEP_SOLVE_ONLY)) {                    // world.py:316-319
            phase_pairs(false);
            if (B2B_IS_T0) w_flags[0] &= ~WORLD_NEW_FIXTURE;
            B2B_SYNC();
            B2B_PROF(PROF_PAIRS0, t);
        }
        phase_collide(after_flat);                               // world.py:338
        B2B_PROF(PROF_COLLIDE, t);
        if (sp.dt > 0.0) {                                       // world.py:341-342, 475-580
            // The island walk is a pure function of the body flags and the contact graph.  While neither changes
            // (the settled regime) last step's island lists, level schedule and island flags are reused as they are.
            if (!(w_flags[0] & WORLD_GRAPH_VALID)) {
                stage_flags();
                build_adjacency();
                B2B_PROF(PROF_ADJACENCY, t);
                phase_islands();
                B2B_PROF(PROF_ISLANDS, t);
                phase_schedule(sp.solve_lanes);
                if (B2B_IS_T0 && !s_ctl[11]) w_flags[0] |= WORLD_GRAPH_VALID;   // a body woken by the walk may seed earlier next time
                B2B_PROF(PROF_S_SCHEDULE, t);
            } else if (B2B_IS_T0) {
                w_island[3] = 0;
            }
            B2B_SYNC();
        }
    }
    // solve kernel: Island.solve for every island (island.py:65-179)
    template <int PJ = 0> B2B_D void run_solve_velocity(const step_params& sp)
    {
        if (!(sp.dt > 0.0)) return;
        const bool profiling = (sp.flags & STEP_PROFILE) != 0;
        long long t = B2B_CLOCK();
#if defined(B2B_TRACE) && defined(__CUDA_ARCH__)
        for (int i = 0; i < 7; ++i) tr[i] = 0u;
        tlast = (unsigned)clock();
#endif
        solve_velocity_phase<PJ>(sp, (real)sp.dt);
#if defined(B2B_TRACE) && defined(__CUDA_ARCH__)
        if (t_id == 0) for (int i = 0; i < 7; ++i) prof[20 + i] += (long long)tr[i];
#endif
        B2B_PROF(PROF_SOLVE, t);
    }
    // The position kernel exists twice: JOINT_CAPABLE = false for batches in which no world has a joint schedule, true
    // otherwise; both are launched and the one whose turn it is not returns at once (b2b_kernels.inl).
    template <bool JOINT_CAPABLE, int PJ = 0> B2B_D void run_solve_position(const step_params& sp)
    {
        if (!(sp.dt > 0.0)) return;
        const bool profiling = (sp.flags & STEP_PROFILE) != 0;
        long long t = B2B_CLOCK();
        solve_position_phase<JOINT_CAPABLE, PJ>(sp, (real)sp.dt);
        B2B_PROF(PROF_SOLVE, t);
    }
    // the solve kernels of batches with a joint capacity
    B2B_D void run_solve_velocity_joints(const step_params& sp) { run_solve_velocity<1>(sp); }
    B2B_D void run_solve_position_joints(const step_params& sp) { run_solve_position<false, 1>(sp); }
    B2B_D void run_solve_velocity_revolute(const step_params& sp) { run_solve_velocity<2>(sp); }
    B2B_D void run_solve_position_revolute(const step_params& sp) { run_solve_position<false, 2>(sp); }
    // broadphase kernel: find_new_contacts (world.py:582-592) for the worlds whose proxies moved (the flat sync
    // kernel raised SUM_MOVED); a world without movers emits no pairs and creates no contacts
    B2B_D void run_pairs(const step_params& sp)
    {
        const bool profiling = (sp.flags & STEP_PROFILE) != 0;
        long long t = B2B_CLOCK();
        if (sp.dt > 0.0) {
            const uint32_t moved = *g_summary & SUM_MOVED;
            B2B_SYNC();
            if (moved) {
                if (B2B_IS_T0) B2B_ATOMIC_AND(g_summary, ~(uint32_t)SUM_MOVED);
                phase_pairs(true);
            } else if (B2B_IS_T0 && w_count[3] != 0) {
                w_count[3] = 0;
            }
            B2B_PROF(PROF_PAIRS, t);
        }
        B2B_SYNC();
        if (profiling && B2B_IS_T0) prof[PROF_WORLD_STEPS] += 1;
    }
    // the three kernels back to back (host check and single-world debugging)
    B2B_D void run(const step_params& sp)
    {
        for (int step = 0; step < sp.n_steps; ++step) {
            if ((sp.flags & STEP_FLAT_COLLIDE) && !(w_flags[0] & WORLD_NEW_FIXTURE)) {
                const int nc = w_count[2];
                for (int c = 0; c < nc; ++c) collide_one(c);
            }
            run_plan(sp);
            if (sp.dt > 0.0) {
                const int n_slots = prepare_count();
                for (int s = 0; s < n_slots; ++s) prepare_slot(sp, step == 0, s);
                const int n_bodies = prepare_bodies();
                for (int b = 0; b < n_bodies; ++b) prepare_body(sp, b);
            }
            if (cap.nj > 0) {
                step_params spj = sp;
                if (step == 0) spj.flags |= STEP_FIRST;
                if (cap.joint_types == (1 << JOINT_REVOLUTE)) { run_solve_velocity<2>(spj); run_solve_position<false, 2>(spj); }
                else { run_solve_velocity<1>(spj); run_solve_position<false, 1>(spj); }
            } else {
                run_solve_velocity(sp);
                run_solve_position<true>(sp);
            }
            const int n_items = sync_items();
            for (int i = 0; i < n_items; ++i) sync_item(sp, i);
            run_pairs(sp);
            run_toi(sp);
        }
    }
};

}  // namespace b2b

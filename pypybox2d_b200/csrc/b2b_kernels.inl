// b2b_kernels.inl -- __global__ entry points and their launchers for ONE scalar type.
// Included by b2b_f32.cu (B2B_REAL=float, FMA contraction on) and b2b_f64.cu (B2B_REAL=double,
// compiled with -fmad=false so the parity build rounds like the reference's unfused Python floats).
// B2B_SUFFIX names the exported launchers (launch_step_f32, ...).
#include <cuda_runtime.h>
#include <cstdlib>
#include <map>
#include "b2b_step.cuh"
#include "b2b_launch.h"
#include "b2b_query.cuh"

namespace b2b {

typedef B2B_REAL real_t;

// The solve kernels run ONE CTA per SM, a team of lanes per world: two constraint records in flight per lane need ~170
// registers, which lets 12 warps share an SM's register file (three per scheduler); how many of them a CTA gets is decided
// by the shared memory the bodies' state of its worlds needs (plan_: 11 warps = 88 pyramid worlds).  One CTA per SM pays the
// per-CTA shared-memory reservation once: one-warp CTAs stopped at 10 warps per SM.
#ifndef B2B_JOINT_MAX_WARPS
#define B2B_JOINT_MAX_WARPS 12      // warps per CTA of the joint-capable solve kernels (their register budget: 65,536 / 32 / warps).  Measured on
                                    // bridge x 131,072 worlds, velocity + position kernel: 12 warps 2.77 + 1.09 ms, 16 warps 3.28 + 0.93, 24 warps 3.79 + 1.06
                                    // (the level sweep's register prefetch spills below 168 registers)
#endif
#ifndef B2B_SOLVE_MAX_WARPS
#define B2B_SOLVE_MAX_WARPS 12
#endif

// The step is three kernels per World.step, each a persistent grid with worlds strided over CTAs:
//   plan_kernel       narrowphase (K1), contact graph + island walk (K2), level schedule      -- CTA per world
//   solve_kernel      Island.solve for every island (K3): the hot kernel, lean in registers and
//                     shared memory so that many worlds are resident per SM                     -- warp per island
//   broadphase_kernel proxy synchronisation (K4), pair finding (K5), contact creation (K6)      -- CTA per world
// State lives in the caller's SoA arrays; per-world plan rows carry the island schedule between kernels; the
// per-CTA scratch slab is reused for every world a CTA owns, so it stays L1/L2 resident.
// Flat narrowphase (K1): one thread per contact of the WHOLE batch -- the colliders are embarrassingly parallel and
// this shape gives them full lanes and their own register budget.  Worlds that still have to create their first
// contacts this step (new fixtures) are left to the plan kernel.
#ifndef B2B_COLLIDE_MIN_CTAS
#define B2B_COLLIDE_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(128, B2B_COLLIDE_MIN_CTAS) B2B_CAT(collide_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c, plan_layout P, uint16_t* plan,
                                                                                                   int w0, int n_range)
{
    const long long total = (long long)n_range * c.nc;                 // worlds [w0, w0 + n_range)
#if defined(B2B_COLLIDE_SORT) && B2B_COLLIDE_SORT
    // EXPERIMENT, off by default (measured slower on B200: 7.05 -> 7.62 ms per 131,072 pyramid worlds -- the flag load in front
    // of the row loads and the two barriers cost more than the divergence they remove, and the permuted rows load uncoalesced).
    // The 128 contacts of a CTA's round are handed to its threads sorted by what they cost: contacts that were touching last
    // step first (SAT of both polygons, incident edge, clipping: ~1,500 instructions), then the others (they usually leave
    // after the first polygon's separating axis: ~400), unused rows last.  A warp runs the long path if ONE of its contacts
    // takes it; unsorted, every warp of a settled pile has such a contact (two thirds of the rows touch).
    __shared__ int s_perm[128];
    __shared__ int s_warp[8];
    const int tid = (int)threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (long long base = (long long)blockIdx.x * 128; base < total; base += (long long)gridDim.x * 128) {
        const long long i = base + tid;
        int key = 2;
        if (i < total) {
            const int w = w0 + (int)(i / c.nc), ci = (int)(i % c.nc);
            if (ci < B.w_count[4 * (size_t)w + 2] && !(B.w_flags[w] & WORLD_NEW_FIXTURE))
                key = (B.c_flags[(size_t)w * c.nc + ci] & CONTACT_TOUCHING) ? 0 : 1;
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, key == 0), m1 = __ballot_sync(0xffffffffu, key == 1);
        if (lane == 0) { s_warp[wid] = __popc(m0); s_warp[4 + wid] = __popc(m1); }
        __syncthreads();
        int off0 = 0, off1 = 0, n0 = 0, n1 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < wid) { off0 += s_warp[k]; off1 += s_warp[4 + k]; }
            n0 += s_warp[k]; n1 += s_warp[4 + k];
        }
        const unsigned below = (1u << lane) - 1u;
        if (key == 0) s_perm[off0 + __popc(m0 & below)] = tid;
        else if (key == 1) s_perm[n0 + off1 + __popc(m1 & below)] = tid;
        __syncthreads();
        if (tid < n0 + n1) {
            const long long j = base + s_perm[tid];
            world_step<real_t> ws;
            ws.bind_rows(B, c, w0 + (int)(j / c.nc), plan, P);
            ws.collide_one((int)(j % c.nc));
        }
        __syncthreads();
    }
#else
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = w0 + (int)(i / c.nc), ci = (int)(i % c.nc);
        if (ci >= B.w_count[4 * (size_t)w + 2]) continue;
        if (B.w_flags[w] & WORLD_NEW_FIXTURE) continue;
        world_step<real_t> ws;
        ws.bind_rows(B, c, w, plan, P);
        ws.collide_one(ci);
    }
#endif
}

// The per-world kernels are one warp per CTA; the warp is split into teams of `lanes` lanes and every team steps its
// own world, so several worlds share each issued instruction (their work is mostly one or two lanes wide).
//
// plan and pair kernels: most worlds of a settled batch have nothing to do (no touching change, nothing to destroy,
// island graph still valid; no proxy moved).  The warp first looks at 32 worlds at once, one per lane, through the
// per-world summary word the flat kernels wrote; only the worlds that need work are then handed to the teams.
// Chunks of 32 worlds are handed out through an atomic counter (zeroed by the host before the launch), so the few
// expensive worlds of a mostly settled batch spread over the whole grid instead of queueing behind one another.
// `chunk` (a power of two <= 32, chosen by the launcher from batch and grid size) is the number of worlds a warp takes at a
// time: 32 when most worlds are skipped, fewer when the batch is small against the grid -- a batch in which every world needs
// its walk would otherwise keep only n_worlds / 32 warps busy, each with 32 worlds in a row (measured: deep stacks, 32,768
// worlds, plan kernel 7.7 ms = one warp's 32 worlds back to back).
template <int ONLY = 0, class Need, class Work>      // ONLY 2 / 1: the kernel's worlds are all of that class (one copy of its work, not two)
__device__ __forceinline__ void for_each_needy_world(const caps_t& c, int lanes, int chunk, unsigned int* chunk_counter, Need need_of, Work work)
{
    const int lane = (int)(threadIdx.x & 31);
    for (;;) {
        int base = 0;
        if (lane == 0) base = (int)atomicAdd(chunk_counter, (unsigned int)chunk);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= c.n_worlds) break;
        const int w = (lane < chunk) ? base + lane : c.n_worlds;
        // need_of: 0 = nothing to do, 1 = mostly sequential work (island walk, schedule: narrow teams, several worlds
        // per warp), 2 = contact-parallel work (contacts of sleeping bodies to evaluate, contacts to destroy and
        // compact, or simply many contacts: the whole warp takes one world).  Measured on B200, plan kernel per 16,384 worlds: landing pyramids
        // 2.3 ms with 8 lanes / 5.4 with 32; vertical stacks 10.0 / 6.3; mixed scene 21.4 / 10.5.
        const int cls = (w < c.n_worlds) ? (int)need_of(w) : 0;
        const unsigned heavy = __ballot_sync(0xffffffffu, cls == 2);
        if (ONLY != 1)
            for (int ord = 0; ord < __popc(heavy); ++ord) work(base + (int)__fns(heavy, 0, ord + 1), 32);
        __syncwarp();
        if (ONLY == 2) continue;
        const unsigned m = __ballot_sync(0xffffffffu, cls == 1);
        // few needy worlds in this chunk (the settled regime): widen the teams so that each world gets more lanes
        const int n = __popc(m);
        int width = lanes;
        while (width < 32 && 32 / (2 * width) >= n) width *= 2;
        const int teams_now = 32 / width, team_now = lane / width;
        for (int ord = team_now; ord < n; ord += teams_now) work(base + (int)__fns(m, 0, ord + 1), width);
        __syncwarp();
    }
}

// B2B_PLAN_MIN_CTAS: resident one-warp CTAs per SM the register allocation is capped for.  The kernel is a sequential walk per
// world -- latency bound -- so it runs on the number of resident worlds: uncapped it took 168 registers = 12 worlds per SM.
// Measured on B200 (plan kernel, ms per step; 1 = uncapped / 16 / 20 / 24): vertical stacks x 65,536 worlds 10.34 / 7.56 / 6.80 /
// 6.75 (31.6 -> 28.1 ms per step); deep stacks x 32,768 6.51 / 6.27 / 6.27 / 6.33 and mixed x 32,768 7.21 / 7.22 / 7.34 / 7.38:
// those two are capped by the shared memory of their contact capacities (14 and 12 worlds per SM) before registers matter.
#ifndef B2B_PLAN_MIN_CTAS
#define B2B_PLAN_MIN_CTAS 20
#endif
__global__ void __launch_bounds__(32, B2B_PLAN_MIN_CTAS) B2B_CAT(plan_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c, plan_layout P, uint16_t* plan,
                                                                         scratch_layout L, unsigned char* scratch, step_params sp,
                                                                         int lanes, size_t team_smem, unsigned int* chunk_counter, int chunk)
{
    extern __shared__ __align__(32) unsigned char smem[];
    world_step<real_t> ws;
    const int teams_per_cta = 32 / lanes;                 // scratch slabs and shared-memory regions of this CTA
    for_each_needy_world(
        c, lanes, chunk, chunk_counter,
        [&](int w) {
            const uint32_t fl = B.w_flags[w];
            const uint32_t sum = *reinterpret_cast<const uint32_t*>(plan + (size_t)w * P.stride + P.off_summary);
            // pending contacts alone (both bodies inactive) need no walk: without a touching change nobody was woken
            const bool settled = (sp.flags & STEP_FLAT_COLLIDE) && !(fl & WORLD_NEW_FIXTURE) && !(sum & (SUM_DESTROY | SUM_CHANGED | SUM_SENSOR)) &&
                                 (!(sp.dt > 0.0) || (fl & WORLD_GRAPH_VALID));
            if (settled && sp.dt > 0.0) B.w_island[4 * (size_t)w + 3] = 0;      // what run_plan leaves behind for such a world
            if (settled && (sum & SUM_PENDING))
                atomicAnd(reinterpret_cast<unsigned int*>(plan + (size_t)w * P.stride + P.off_summary), ~(unsigned int)SUM_PENDING);
            if (settled) return 0;
            const int n_contacts = B.w_count[4 * (size_t)w + 2];
            if (n_contacts < 64) return 1;                            // too few contacts to occupy a warp
            if ((sum & (SUM_PENDING | SUM_DESTROY)) || (fl & WORLD_NEW_FIXTURE) || !(sp.flags & STEP_FLAT_COLLIDE)) return 2;
            return n_contacts >= 700 ? 2 : 1;                         // many contacts: the contact-parallel share dominates
        },
        [&](int w, int width) {
            // every thread derives the world's row pointers from the kernel parameters, so the compiler knows their
            // address space (LDG/STG/LDS instead of generic LD/ST) and rematerialises them from the constant bank
            const int team = ws.bind_team_lanes(width);   // < teams_per_cta: wider teams use the first slabs / regions
            ws.bind_rows(B, c, w, plan, P);
            ws.bind_scratch_aux(scratch + (size_t)(blockIdx.x * teams_per_cta + team) * L.total_aux, L);
            ws.bind_smem_aux(smem + (size_t)team * team_smem);
            ws.run_plan(sp);
            __syncwarp(ws.t_mask);
        });
}

// Constraint records (K3 set-up, contact.py:376-496) for a GROUP of worlds [w0, w0 + n_group): one thread per
// (world, constraint slot), records into the group's slab `rec` (one scratch_layout::total_rec per world).
#ifndef B2B_PREPARE_MIN_CTAS
#define B2B_PREPARE_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(128, B2B_PREPARE_MIN_CTAS) B2B_CAT(prepare_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c, plan_layout P, uint16_t* plan,
                                                                            scratch_layout L, unsigned char* rec, step_params sp,
                                                                            int first_step, int w0, int n_group)
{
    const long long total = (long long)n_group * c.nc;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i / c.nc), slot = (int)(i % c.nc);
        const int w = w0 + g;
        if (slot >= B.w_island[4 * (size_t)w + 2]) continue;
        world_step<real_t> ws;
        ws.bind_rows(B, c, w, plan, P);
        ws.bind_scratch_solve(rec, (size_t)g * L.total_rec, L, nullptr);
        ws.prepare_slot(sp, first_step != 0, slot);
        if (slot == 0 && sp.joint_any != nullptr && ws.g_joint[0] != 0 && *sp.joint_any == 0u) atomicExch(sp.joint_any, 1u);
    }
    // island.py:71-95 for every body of the group (runs after no other thread of this kernel needs b_vel changed:
    // the integrated velocities go to the staging row, b_vel itself is rewritten by the velocity kernel)
    const long long total_b = (long long)n_group * c.nb;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_b; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i / c.nb), b = (int)(i % c.nb);
        const int w = w0 + g;
        if (b >= B.w_count[4 * (size_t)w]) continue;
        world_step<real_t> ws;
        ws.bind_rows(B, c, w, plan, P);
        ws.bind_scratch_solve(rec, (size_t)g * L.total_rec, L, nullptr);
        ws.prepare_body(sp, b);
    }
}

// Island.solve in two kernels (see world_step::solve_velocity_phase / solve_position_phase): one CTA per SM, every warp
// split into teams of `lanes` lanes, one world per team.  The teams of a warp start their worlds together and meet again
// after them: they share every issued instruction, so they must stay in step.
#define B2B_SOLVE_KERNEL(NAME, CALL, NOT_MY_TURN, KIND) B2B_SOLVE_KERNEL_W(NAME, CALL, NOT_MY_TURN, KIND, B2B_SOLVE_MAX_WARPS)
#define B2B_SOLVE_KERNEL_W(NAME, CALL, NOT_MY_TURN, KIND, MAXW)                                                                                \
    __global__ void __launch_bounds__(32 * MAXW, 1) B2B_CAT(NAME, B2B_SUFFIX)(                                                \
        batch_ptrs<real_t> B, caps_t c, plan_layout P, uint16_t* plan, scratch_layout L, unsigned char* rec,                   \
        unsigned char* team_prof, step_params sp, int lanes, size_t team_smem, int w0, int n_group)                            \
    {                                                                                                                          \
        extern __shared__ __align__(32) unsigned char smem[];                                                                  \
        if (NOT_MY_TURN) return;                                                                                               \
        world_step<real_t> ws;                                                                                                 \
        const int team = ws.bind_team_lanes(lanes);                /* index among the CTA's teams */                           \
        const int teams_per_warp = 32 / lanes, teams_per_cta = (int)(blockDim.x >> 5) * teams_per_warp;                        \
        const int team_in_warp = team % teams_per_warp;                                                                        \
        const int slot = blockIdx.x * teams_per_cta + team;                                                                    \
        /* Worlds are handed out a warp's worth at a time (its teams stay in step: they share every issued instruction)     */ \
        /* through a counter (sp.solve_ctr[KIND], zeroed by the launcher): a warp on a less crowded scheduler takes more of */ \
        /* them, and the last worlds of a launch go to whoever is free first.  Without a counter: strided over the warps.   */ \
        unsigned int* const ctr = sp.solve_ctr ? sp.solve_ctr + KIND : nullptr;                                                \
        for (int g0 = slot - team_in_warp;; g0 += gridDim.x * teams_per_cta) {                                                 \
            if (ctr) {                                                                                                         \
                int v = 0;                                                                                                     \
                if ((threadIdx.x & 31) == 0) v = (int)atomicAdd(ctr, (unsigned int)teams_per_warp);                            \
                g0 = __shfl_sync(0xffffffffu, v, 0);                                                                           \
            }                                                                                                                  \
            if (g0 >= n_group) break;                                                                                          \
            const int g = g0 + team_in_warp;                                                                                   \
            if (g < n_group) {                                                                                                 \
                ws.bind_rows(B, c, w0 + g, plan, P);                                                                           \
                ws.bind_scratch_solve(rec, (size_t)g * L.total_rec, L, team_prof + (size_t)slot * L.total_sprof);             \
                ws.bind_smem_solve(smem + (size_t)team * team_smem);                                                           \
                {                                                  /* lanes of this warp whose team has a world this round */  \
                    const int rem_ = n_group - g0;                                                                             \
                    const int act_ = (rem_ < teams_per_warp ? rem_ : teams_per_warp) * lanes;                                  \
                    ws.w_mask = act_ >= 32 ? 0xffffffffu : ((1u << act_) - 1u);                                                \
                }                                                                                                              \
                ws.CALL(sp);                                                                                                   \
            }                                                                                                                  \
            __syncwarp();                                                                                                      \
        }                                                                                                                      \
    }
B2B_SOLVE_KERNEL(solve_velocity_kernel_, run_solve_velocity, false, 0)
// sp.joint_any: raised by this group's prepare kernel (or an earlier group's) when some world has a joint island schedule
B2B_SOLVE_KERNEL(solve_position_kernel_, template run_solve_position<false>, *sp.joint_any != 0u, 1)
B2B_SOLVE_KERNEL(solve_position_joint_kernel_, template run_solve_position<true>, *sp.joint_any == 0u, 2)
// batches with a joint capacity (caps_t::nj > 0): the same two kernels with the joint hooks of b2b_joints.inl compiled in
B2B_SOLVE_KERNEL_W(solve_velocity_joints_kernel_, run_solve_velocity_joints, false, 0, B2B_JOINT_MAX_WARPS)
B2B_SOLVE_KERNEL_W(solve_position_joints_kernel_, run_solve_position_joints, false, 1, B2B_JOINT_MAX_WARPS)
// ... and for batches whose joints are all revolute joints (caps_t::joint_types): no type dispatch
B2B_SOLVE_KERNEL_W(solve_velocity_revolute_kernel_, run_solve_velocity_revolute, false, 0, B2B_JOINT_MAX_WARPS)
B2B_SOLVE_KERNEL_W(solve_position_revolute_kernel_, run_solve_position_revolute, false, 1, B2B_JOINT_MAX_WARPS)
#undef B2B_SOLVE_KERNEL
#undef B2B_SOLVE_KERNEL_W

// Proxy synchronisation (K4: body.py:898-904, fixture.py:287-298, dynamictree.py:152-188) and World.clear_forces as a
// flat grid: one thread per (world, proxy-or-body index).  Worlds with a mover are flagged for the pair kernel.
__global__ void __launch_bounds__(128) B2B_CAT(sync_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c, plan_layout P, uint16_t* plan,
                                                                         step_params sp)
{
    const int per_world = c.nf > c.nb ? c.nf : c.nb;
    const long long total = (long long)c.n_worlds * per_world;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i / per_world), k = (int)(i % per_world);
        world_step<real_t> ws;
        ws.bind_rows(B, c, w, plan, P);
        ws.sync_item(sp, k);
    }
}

__global__ void __launch_bounds__(32) B2B_CAT(broadphase_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c, plan_layout P, uint16_t* plan,
                                                                               scratch_layout L, unsigned char* scratch, step_params sp,
                                                                               int lanes, size_t team_smem, unsigned int* chunk_counter, int chunk)
{
    extern __shared__ __align__(32) unsigned char smem[];
    world_step<real_t> ws;
    const int teams_per_cta = 32 / lanes;
    for_each_needy_world(
        c, lanes, chunk, chunk_counter,
        [&](int w) {
            // profiling counts world-steps in run_pairs; otherwise only worlds with a mover (or a stale pair count) need it
            if (sp.flags & STEP_PROFILE) return true;
            const uint32_t sum = *reinterpret_cast<const uint32_t*>(plan + (size_t)w * P.stride + P.off_summary);
            return (sp.dt > 0.0) && ((sum & SUM_MOVED) || B.w_count[4 * (size_t)w + 3] != 0);
        },
        [&](int w, int width) {
            const int team = ws.bind_team_lanes(width);
            ws.bind_rows(B, c, w, plan, P);
            ws.bind_scratch_aux(scratch + (size_t)(blockIdx.x * teams_per_cta + team) * L.total_aux, L);
            ws.bind_smem_aux(smem + (size_t)team * team_smem, true);
            ws.run_pairs(sp);
            __syncwarp(ws.t_mask);
        });
}

// Continuous collision: World.solve_toi (world.py:594-821) for the worlds whose flags carry WORLD_CONTINUOUS, after the
// discrete step's find_new_contacts.  Same launch shape and scratch slabs as the pair kernel (the loop calls phase_pairs
// after every solved event), its shared-memory layout without the candidate bitmap.  TOI_LANES lanes per world: what happens
// per event is sequential (one lane) and the kernel is bound by instruction fetch and latency, so it runs on the number of worlds
// in flight and on how many of them share an instruction; the contact-parallel first look is a few rounds either way.  A batch
// without such worlds costs one flag word per world.  Measured on B200 (ms per step with continuous collision on: 131,072 settled
// pyramids / 65,536 landing pyramids; the same steps without it: 36.6 / 20.9): 4 lanes 40.7 / 34.1, 8 lanes 39.2 / 31.8, 16 lanes
// 39.6 / 31.9, 32 lanes (with the bitmap) 40.7 / 33.5.
#ifndef B2B_TOI_LANES
#define B2B_TOI_LANES 8
#endif
__global__ void __launch_bounds__(32) B2B_CAT(toi_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c, plan_layout P, uint16_t* plan,
                                                                        scratch_layout L, unsigned char* scratch, step_params sp,
                                                                        int lanes, size_t team_smem, unsigned int* chunk_counter, int chunk)
{
    extern __shared__ __align__(32) unsigned char smem[];
    world_step<real_t> ws;
    const int teams_per_cta = 32 / lanes;
    for_each_needy_world<1>(
        c, lanes, chunk, chunk_counter,
        [&](int w) { return ((sp.dt > 0.0) && (B.w_flags[w] & WORLD_CONTINUOUS)) ? 1 : 0; },
        [&](int w, int width) {
            const int team = ws.bind_team_lanes(width);
            ws.bind_rows(B, c, w, plan, P);
            ws.bind_scratch_aux(scratch + (size_t)(blockIdx.x * teams_per_cta + team) * L.total_aux, L);
            ws.bind_smem_aux(smem + (size_t)team * team_smem, true, false);
            ws.run_toi(sp);
            __syncwarp(ws.t_mask);
        });
}

// Fixture._create_proxies + DynamicTree.create_proxy (fixture.py:250-261, dynamictree.py:115-134)
__global__ void __launch_bounds__(256) B2B_CAT(init_proxies_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c)
{
    const long long total = (long long)c.n_worlds * c.nf;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i / c.nf), f = (int)(i % c.nf);
        const int nf = B.w_count[4 * (size_t)w + 1];
        if (f == 0) B.w_flags[w] = (B.w_flags[w] | WORLD_NEW_FIXTURE) & ~WORLD_GRAPH_VALID;
        if (f >= nf) continue;
        const size_t of = (size_t)w * c.nf + f;
        const int b = B.f_body[of];
        if (!(B.b_flags[(size_t)w * c.nb + b] & BODY_ACTIVE)) {
            // no proxies for the fixtures of an inactive body (body.py:288-297): an inverted box overlaps nothing
            real_t* tight = B.f_tight + 4 * of; real_t* fat = B.f_fat + 4 * of;
            const real_t big = real_t(1.0e30);
            tight[0] = tight[1] = fat[0] = fat[1] = big; tight[2] = tight[3] = fat[2] = fat[3] = -big;
            B.f_moved[of] = 0;
            continue;
        }
        const xform<real_t> xf = xf_load<real_t>(B.b_xf + 4 * ((size_t)w * c.nb + b));
        const shape_view<real_t> s = shape_load<real_t>(B.s_meta, B.s_radius, B.s_data, B.f_shape[of]);
        real_t a[4];
        shape_aabb(s, xf, a);
        real_t* tight = B.f_tight + 4 * of; real_t* fat = B.f_fat + 4 * of;
        const real_t ext = K<real_t>::aabb_extension;
        tight[0] = a[0]; tight[1] = a[1]; tight[2] = a[2]; tight[3] = a[3];
        fat[0] = a[0] - ext; fat[1] = a[1] - ext; fat[2] = a[2] + ext; fat[3] = a[3] + ext;
        B.f_moved[of] = 1;
    }
}

// narrowphase test hook: one thread per explicit pair, fresh manifold
__global__ void __launch_bounds__(256) B2B_CAT(collide_pairs_kernel_, B2B_SUFFIX)(
    batch_ptrs<real_t> B, int n, const int32_t* sa, const int32_t* sb, const real_t* xfa, const real_t* xfb,
    int32_t* mtype, real_t* mf, real_t* pts, uint32_t* id)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        manifold<real_t> m;
        m.type = 0; m.count = 0;
        m.local_normal = m.local_point = m.pt[0] = m.pt[1] = mk<real_t>(real_t(0), real_t(0));
        m.id[0] = m.id[1] = 0u;
        const shape_view<real_t> a = shape_load<real_t>(B.s_meta, B.s_radius, B.s_data, sa[i]);
        const shape_view<real_t> b = shape_load<real_t>(B.s_meta, B.s_radius, B.s_data, sb[i]);
        collide_shapes(m, a, xf_load<real_t>(xfa + 4 * (size_t)i), b, xf_load<real_t>(xfb + 4 * (size_t)i));
        mtype[i] = m.type | (m.count << 8);
        mf[4 * (size_t)i] = m.local_normal.x; mf[4 * (size_t)i + 1] = m.local_normal.y;
        mf[4 * (size_t)i + 2] = m.local_point.x; mf[4 * (size_t)i + 3] = m.local_point.y;
        pts[4 * (size_t)i] = m.pt[0].x; pts[4 * (size_t)i + 1] = m.pt[0].y;
        pts[4 * (size_t)i + 2] = m.pt[1].x; pts[4 * (size_t)i + 3] = m.pt[1].y;
        id[2 * (size_t)i] = m.id[0]; id[2 * (size_t)i + 1] = m.id[1];
    }
}

// host-driven step: scatter staged (fx, fy, torque) into b_force / b_vel[3]
__global__ void __launch_bounds__(256) B2B_CAT(scatter_forces_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c, const real_t* stage,
                                                                                   int w0, int n_range)
{
    const long long first = (long long)w0 * c.nb, total = (long long)(w0 + n_range) * c.nb;      // worlds [w0, w0 + n_range)
    for (long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const real_t fx = stage[3 * i], fy = stage[3 * i + 1], tq = stage[3 * i + 2];
        const uint32_t fl = B.b_flags[i];
        // Body.apply_force / apply_torque (body.py:477-545): only dynamic bodies take forces, and applying one wakes the
        // body (Body.awake = True, body.py:800-804: the sleep timer restarts).  A body woken from outside may seed the
        // next island walk, so the world's island lists are rebuilt.
        if ((fl & BODY_TYPE_MASK) != BODY_DYNAMIC) continue;
        if (!(fl & BODY_AWAKE)) {
            if (fx == real_t(0) && fy == real_t(0) && tq == real_t(0)) continue;       // nothing applied: a sleeper stays asleep
            B.b_flags[i] = fl | BODY_AWAKE;
            B.b_pos[4 * i + 3] = real_t(0);
            atomicAnd(&B.w_flags[i / c.nb], ~(uint32_t)WORLD_GRAPH_VALID);
        }
        B.b_force[2 * i] = fx; B.b_force[2 * i + 1] = fy;
        B.b_vel[4 * i + 3] = tq;
    }
}

// host-driven step: gather (c.x, c.y, a, v.x, v.y, w) for the device->host copy
__global__ void __launch_bounds__(256) B2B_CAT(gather_state_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c, real_t* stage,
                                                                                 int w0, int n_range)
{
    const long long first = (long long)w0 * c.nb, total = (long long)(w0 + n_range) * c.nb;      // worlds [w0, w0 + n_range)
    for (long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        stage[6 * i] = B.b_pos[4 * i]; stage[6 * i + 1] = B.b_pos[4 * i + 1]; stage[6 * i + 2] = B.b_pos[4 * i + 2];
        stage[6 * i + 3] = B.b_vel[4 * i]; stage[6 * i + 4] = B.b_vel[4 * i + 1]; stage[6 * i + 5] = B.b_vel[4 * i + 2];
    }
}

// per-world checksum / energy / counts for the optional end-of-run gather
__global__ void __launch_bounds__(256) B2B_CAT(checksum_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c, double* out)
{
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < c.n_worlds; w += gridDim.x * blockDim.x) {
        const int nb = B.w_count[4 * (size_t)w];
        const double gx = (double)B.w_gravity[2 * (size_t)w], gy = (double)B.w_gravity[2 * (size_t)w + 1];
        unsigned long long h = 1469598103934665603ull;
        double energy = 0.0;
        int awake = 0;
        for (int b = 0; b < nb; ++b) {
            const size_t o = (size_t)w * c.nb + b;
            const real_t* p = B.b_pos + 4 * o; const real_t* v = B.b_vel + 4 * o; const real_t* m = B.b_mass + 4 * o;
            const double vals[6] = {(double)p[0], (double)p[1], (double)p[2], (double)v[0], (double)v[1], (double)v[2]};
            for (int k = 0; k < 6; ++k) {
                unsigned long long bits = (unsigned long long)__double_as_longlong(vals[k]);
                for (int q = 0; q < 8; ++q) { h ^= (bits >> (8 * q)) & 0xffull; h *= 1099511628211ull; }
            }
            const double mass = (m[0] > real_t(0)) ? 1.0 / (double)m[0] : 0.0;
            const double inertia = (m[1] > real_t(0)) ? 1.0 / (double)m[1] : 0.0;
            energy += 0.5 * mass * (vals[3] * vals[3] + vals[4] * vals[4]) + 0.5 * inertia * vals[5] * vals[5]
                      - mass * (gx * vals[0] + gy * vals[1]);
            awake += (B.b_flags[o] & BODY_AWAKE) ? 1 : 0;
        }
        out[4 * (size_t)w] = (double)(h & ((1ull << 52) - 1ull));       // an exactly representable integer: never NaN / Inf, == works
        out[4 * (size_t)w + 1] = energy;
        out[4 * (size_t)w + 2] = (double)B.w_count[4 * (size_t)w + 2];
        out[4 * (size_t)w + 3] = (double)awake;
    }
}

__global__ void B2B_CAT(status_kernel_, B2B_SUFFIX)(const uint32_t* w_status, int n, uint32_t* out)
{
    uint32_t acc = 0;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) acc |= w_status[w];
    if (acc) atomicOr(out, acc);
}


// ---- queries (b2b_query.cuh): one warp per query, the world's proxies strided over its lanes -----------------
// World.query_aabb (world.py:371-385): out_count[q] = fixtures of world q_world[q] whose fat AABB overlaps q_box[q]; the first
// max_hits of them, ascending, in out_fix[q][.]
__global__ void __launch_bounds__(128) B2B_CAT(query_aabb_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c, int n_q, const int32_t* q_world,
                                                                                const real_t* q_box, int max_hits, int32_t* out_count,
                                                                                int32_t* out_fix)
{
    const int lane = (int)(threadIdx.x & 31);
    const int warps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int q = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); q < n_q; q += warps) {
        const int w = q_world[q];
        int total = 0;
        if (w >= 0 && w < c.n_worlds) {
            const real_t box[4] = {q_box[4 * (size_t)q], q_box[4 * (size_t)q + 1], q_box[4 * (size_t)q + 2], q_box[4 * (size_t)q + 3]};
            const int nf = B.w_count[4 * (size_t)w + 1];
            const real_t* fat = B.f_fat + 4 * (size_t)w * c.nf;
            for (int base = 0; base < nf; base += 32) {
                const int f = base + lane;
                bool hit = false;
                if (f < nf) {
                    const real_t fb[4] = {fat[4 * (size_t)f], fat[4 * (size_t)f + 1], fat[4 * (size_t)f + 2], fat[4 * (size_t)f + 3]};
                    hit = aabb_overlap<real_t>(fb, box);
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                const int at = total + __popc(m & ((1u << lane) - 1u));
                if (hit && at < max_hits) out_fix[(size_t)q * max_hits + at] = f;
                total += __popc(m);
            }
        }
        if (lane == 0) out_count[q] = total;
    }
}

// World.ray_cast (world.py:387-440).  mode 0: every fixture the segment hits (a callback that returns 1), ascending fixture
// index, out_hit[q][k] = {fraction, normal.x, normal.y, point.x, point.y}; mode 1: the closest hit only (a callback that returns
// the fraction; ties: the lowest fixture index), in slot 0.  out_count[q] = number of hits found (mode 1: 0 or 1).
__global__ void __launch_bounds__(128) B2B_CAT(ray_cast_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c, int n_q, const int32_t* q_world,
                                                                              const real_t* q_seg, int mode, int max_hits, int32_t* out_count,
                                                                              int32_t* out_fix, real_t* out_hit)
{
    typedef vec2<real_t> V2;
    const int lane = (int)(threadIdx.x & 31);
    const int warps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int q = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); q < n_q; q += warps) {
        const int w = q_world[q];
        int total = 0;
        real_t best_fr = real_t(2);
        int best_f = 0x7fffffff;
        V2 best_n = mk<real_t>(real_t(0), real_t(0));
        const V2 p1 = mk<real_t>(q_seg[4 * (size_t)q], q_seg[4 * (size_t)q + 1]), p2 = mk<real_t>(q_seg[4 * (size_t)q + 2], q_seg[4 * (size_t)q + 3]);
        const V2 dseg = p2 - p1;
        const bool valid = (w >= 0 && w < c.n_worlds) && (length_sqr(dseg) > real_t(0));      // dynamictree.py:222-223 raises ValueError
        if (valid) {
            const int nf = B.w_count[4 * (size_t)w + 1];
            const size_t of = (size_t)w * c.nf, ob = (size_t)w * c.nb;
            for (int base = 0; base < nf; base += 32) {
                const int f = base + lane;
                bool hit = false;
                V2 normal = mk<real_t>(real_t(0), real_t(0));
                real_t fraction = real_t(0);
                if (f < nf) {
                    const real_t* fp = B.f_fat + 4 * (of + f);
                    const real_t fb[4] = {fp[0], fp[1], fp[2], fp[3]};
                    if (ray_leaf_test<real_t>(fb, p1, p2)) {
                        const shape_view<real_t> s = shape_load<real_t>(B.s_meta, B.s_radius, B.s_data, B.f_shape[of + f]);
                        const xform<real_t> xf = xf_load<real_t>(B.b_xf + 4 * (ob + B.f_body[of + f]));
                        hit = shape_ray_cast<real_t>(s, xf, p1, p2, normal, fraction);
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (mode == 0) {
                    const int at = total + __popc(m & ((1u << lane) - 1u));
                    if (hit && at < max_hits) {
                        const size_t o = (size_t)q * max_hits + at;
                        const V2 point = (real_t(1) - fraction) * p1 + fraction * p2;            // world.py:427
                        out_fix[o] = f;
                        out_hit[5 * o] = fraction; out_hit[5 * o + 1] = normal.x; out_hit[5 * o + 2] = normal.y;
                        out_hit[5 * o + 3] = point.x; out_hit[5 * o + 4] = point.y;
                    }
                } else if (hit && (fraction < best_fr || (fraction == best_fr && f < best_f))) {
                    best_fr = fraction; best_f = f; best_n = normal;
                }
                total += __popc(m);
            }
            if (mode != 0) {
                // warp argmin over (fraction, fixture)
                for (int d = 16; d >= 1; d >>= 1) {
                    const real_t ofr = __shfl_xor_sync(0xffffffffu, best_fr, d);
                    const int ofx = __shfl_xor_sync(0xffffffffu, best_f, d);
                    const real_t onx = __shfl_xor_sync(0xffffffffu, best_n.x, d), ony = __shfl_xor_sync(0xffffffffu, best_n.y, d);
                    if (ofr < best_fr || (ofr == best_fr && ofx < best_f)) { best_fr = ofr; best_f = ofx; best_n = mk<real_t>(onx, ony); }
                }
                if (lane == 0 && total > 0 && max_hits > 0) {
                    const size_t o = (size_t)q * max_hits;
                    const V2 point = (real_t(1) - best_fr) * p1 + best_fr * p2;
                    out_fix[o] = best_f;
                    out_hit[5 * o] = best_fr; out_hit[5 * o + 1] = best_n.x; out_hit[5 * o + 2] = best_n.y;
                    out_hit[5 * o + 3] = point.x; out_hit[5 * o + 4] = point.y;
                }
                total = total > 0 ? 1 : 0;
            }
        }
        if (lane == 0) out_count[q] = valid ? total : -1;
    }
}

// Fixture.test_point (fixture.py:228-233) for n (world, fixture, point) triples
__global__ void __launch_bounds__(256) B2B_CAT(test_points_kernel_, B2B_SUFFIX)(batch_ptrs<real_t> B, caps_t c, int n, const int32_t* q_world,
                                                                                 const int32_t* q_fix, const real_t* q_pt, uint8_t* out)
{
    for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < n; i += (int)(gridDim.x * blockDim.x)) {
        const int w = q_world[i], f = q_fix[i];
        bool in = false;
        if (w >= 0 && w < c.n_worlds && f >= 0 && f < B.w_count[4 * (size_t)w + 1]) {
            const size_t of = (size_t)w * c.nf, ob = (size_t)w * c.nb;
            const shape_view<real_t> s = shape_load<real_t>(B.s_meta, B.s_radius, B.s_data, B.f_shape[of + f]);
            const xform<real_t> xf = xf_load<real_t>(B.b_xf + 4 * (ob + B.f_body[of + f]));
            in = shape_test_point<real_t>(s, xf, mk<real_t>(q_pt[2 * (size_t)i], q_pt[2 * (size_t)i + 1]));
        }
        out[i] = in ? 1 : 0;
    }
}

// ---- launchers (plain C++ so b2b_api.cpp stays free of templates) ----
static inline batch_ptrs<real_t> as_ptrs(const void* const* fields)
{
    batch_ptrs<real_t> B;
    static_assert(sizeof(B) == sizeof(void*) * B2B_NFIELDS, "batch_ptrs must mirror b2b_buffers");
    memcpy(&B, fields, sizeof(B));
    return B;
}

template <class Kern>
static int plan_one(Kern kern, long long n_items, int threads, size_t smem, int ctas_per_sm, int n_sm, kernel_plan* out)
{
    out->threads = threads;
    out->smem = smem;
    // the attribute belongs to the kernel, not to a context: never lower it below what an earlier context launches with
    static std::map<const void*, size_t> smem_limit_set;          // keyed by kernel (several share one signature)
    size_t& limit = smem_limit_set[reinterpret_cast<const void*>(kern)];
    cudaError_t e = cudaSuccess;
    if (smem > limit) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        limit = smem;
    }
    if (const char* env = getenv("B2B_SMEM_CARVEOUT")) {          // experiment: percent of the unified L1/shared array
        if (ctas_per_sm > 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(env));
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) return (int)cudaErrorInvalidConfiguration;     // does not fit an SM (registers or shared memory)
    if (ctas_per_sm > 0 && ctas_per_sm < per_sm) per_sm = ctas_per_sm;
    out->ctas_per_sm = per_sm;
    long long g = (long long)per_sm * n_sm;
    if (g > n_items) g = n_items;
    if (g < 1) g = 1;
    out->grid = (int)g;
    return 0;
}

int B2B_CAT(plan_, B2B_SUFFIX)(const caps_t& c, int lanes, int lanes_aux, int ctas_per_sm, int n_sm, launch_plan* out)
{
    const smem_layout S = smem_plan_layout(c, (int)sizeof(real_t));
    out->layout = scratch_plan(c, (int)sizeof(real_t));
    out->plan = plan_plan(c);
    out->lanes = lanes;
    out->lanes_aux = lanes_aux;
    out->n_sm = n_sm;
    // Teams of one warp look up the same body indices at the same time (settled worlds of one scene share their
    // schedule): skew every team's state row by 128 / teams bytes so that equal indices fall into different banks.
    out->team_smem = align_up(S.total_solve, 128) + (size_t)((128 / (32 / lanes)) < 16 ? 16 : (128 / (32 / lanes)));
    out->team_smem_plan = S.total_plan;
    out->team_smem_broad = S.total_broad;
    const int teams = 32 / lanes, teams_aux = 32 / lanes_aux;
    int rc = plan_one(B2B_CAT(plan_kernel_, B2B_SUFFIX), ((long long)c.n_worlds + teams_aux - 1) / teams_aux, 32,
                      S.total_plan * teams_aux, 0, n_sm, &out->k_plan);
    if (rc) return rc;
    // the solve kernels share one launch geometry: one CTA per SM with as many warps as the bodies' state of their worlds
    // leaves room for (`ctas_per_sm` > 0, a development knob, caps the warps per SM)
    {
        int dev = 0, smem_max = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        const size_t per_warp = out->team_smem * teams;
        int warps = (int)((size_t)smem_max / per_warp);
        const int max_warps = (c.nj > 0) ? B2B_JOINT_MAX_WARPS : B2B_SOLVE_MAX_WARPS;     // batches with joints launch the joint-capable kernels only
        if (warps > max_warps) warps = max_warps;
        if (ctas_per_sm > 0 && warps > ctas_per_sm) warps = ctas_per_sm;
        if (warps < 1) return (int)cudaErrorInvalidConfiguration;       // one warp's worlds do not fit an SM's shared memory
        const long long per_cta = (long long)warps * teams;
        kernel_plan k_pos;
        if (c.nj == 0) {
            rc = plan_one(B2B_CAT(solve_velocity_kernel_, B2B_SUFFIX), ((long long)c.n_worlds + per_cta - 1) / per_cta, 32 * warps, per_warp * warps,
                          1, n_sm, &out->k_solve);
            if (rc) return rc;
            rc = plan_one(B2B_CAT(solve_position_kernel_, B2B_SUFFIX), ((long long)c.n_worlds + per_cta - 1) / per_cta, 32 * warps, per_warp * warps,
                          1, n_sm, &k_pos);
            if (rc) return rc;
            if (k_pos.grid < out->k_solve.grid) out->k_solve = k_pos;
            rc = plan_one(B2B_CAT(solve_position_joint_kernel_, B2B_SUFFIX), ((long long)c.n_worlds + per_cta - 1) / per_cta, 32 * warps, per_warp * warps,
                          1, n_sm, &k_pos);
            if (rc) return rc;
            if (k_pos.grid < out->k_solve.grid) out->k_solve = k_pos;
        } else {
            rc = plan_one(B2B_CAT(solve_velocity_joints_kernel_, B2B_SUFFIX), ((long long)c.n_worlds + per_cta - 1) / per_cta, 32 * warps, per_warp * warps,
                          1, n_sm, &out->k_solve);
            if (rc) return rc;
            rc = plan_one(B2B_CAT(solve_position_joints_kernel_, B2B_SUFFIX), ((long long)c.n_worlds + per_cta - 1) / per_cta, 32 * warps, per_warp * warps,
                          1, n_sm, &k_pos);
            if (rc) return rc;
            if (k_pos.grid < out->k_solve.grid) out->k_solve = k_pos;
            rc = plan_one(B2B_CAT(solve_velocity_revolute_kernel_, B2B_SUFFIX), ((long long)c.n_worlds + per_cta - 1) / per_cta, 32 * warps, per_warp * warps,
                          1, n_sm, &k_pos);
            if (rc) return rc;
            if (k_pos.grid < out->k_solve.grid) out->k_solve = k_pos;
            rc = plan_one(B2B_CAT(solve_position_revolute_kernel_, B2B_SUFFIX), ((long long)c.n_worlds + per_cta - 1) / per_cta, 32 * warps, per_warp * warps,
                          1, n_sm, &k_pos);
            if (rc) return rc;
            if (k_pos.grid < out->k_solve.grid) out->k_solve = k_pos;
        }
    }
    rc = plan_one(B2B_CAT(broadphase_kernel_, B2B_SUFFIX), ((long long)c.n_worlds + teams_aux - 1) / teams_aux, 32,
                  S.total_broad * teams_aux, 0, n_sm, &out->k_broad);
    if (rc) return rc;
    {   // the TOI kernel runs on the pair kernel's geometry: its grid may only be smaller (more registers)
        kernel_plan k_toi;
        const int teams_toi = 32 / B2B_TOI_LANES;
        out->team_smem_toi = S.total_broad_nobits;
        rc = plan_one(B2B_CAT(toi_kernel_, B2B_SUFFIX), ((long long)c.n_worlds + teams_toi - 1) / teams_toi, 32,
                      S.total_broad_nobits * teams_toi, 0, n_sm, &k_toi);
        if (rc) return rc;
        out->k_toi = k_toi;
    }
    out->aux_slabs = (out->k_plan.grid > out->k_broad.grid ? out->k_plan.grid : out->k_broad.grid) * teams_aux;
    if (out->aux_slabs < out->k_toi.grid * (32 / B2B_TOI_LANES)) out->aux_slabs = out->k_toi.grid * (32 / B2B_TOI_LANES);
    out->solve_slabs = out->k_solve.grid * (out->k_solve.threads / 32) * teams;
    return 0;
}

// One World.step = narrowphase + plan + (per group of worlds: prepare, velocity, position) + sync + pairs.  The parts are
// separate launchers so that the host-buffer entry point (b2b_step_host) can run them chunk by chunk between its copies.
// worlds a warp of the plan / pair kernels takes at a time (see for_each_needy_world): at least two grabs per warp
static inline int world_chunk(int n_worlds, int grid)
{
    int chunk = 32;
    while (chunk > 1 && (long long)chunk * grid * 2 > n_worlds) chunk >>= 1;
    if (const char* env = getenv("B2B_WORLD_CHUNK")) { const int v = atoi(env); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) chunk = v; }
    return chunk;
}

static inline int flat_blocks(long long items, int n_sm)
{
    long long blocks = (items + 127) / 128;
    if (blocks > (long long)n_sm * 64) blocks = (long long)n_sm * 64;
    return blocks < 1 ? 1 : (int)blocks;
}

// flat narrowphase (K1) for worlds [w0, w0 + n)
void B2B_CAT(launch_collide_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, const launch_plan& p, uint16_t* plan, int w0, int n,
                                          cudaStream_t st)
{
    B2B_CAT(collide_kernel_, B2B_SUFFIX)<<<flat_blocks((long long)n * c.nc, p.n_sm), 128, 0, st>>>(as_ptrs(fields), c, p.plan, plan, w0, n);
}

// plan kernel (K1 ordering / destruction, K2 islands, level schedule) for every world
void B2B_CAT(launch_plan_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, const launch_plan& p, uint16_t* plan,
                                       unsigned char* scratch_aux, unsigned int* counters, const step_params& sp, cudaStream_t st)
{
    cudaMemsetAsync(counters, 0, 5 * sizeof(unsigned int), st);     // chunk counters of the plan and pair kernels, joint-schedule words, chunk counter of the TOI kernel
    if (sp.flags & STEP_GROUP_KERNEL) cudaMemsetAsync(counters + 2, 1, sizeof(unsigned int), st);   // any non-zero word
    B2B_CAT(plan_kernel_, B2B_SUFFIX)<<<p.k_plan.grid, 32, p.k_plan.smem, st>>>(as_ptrs(fields), c, p.plan, plan, p.layout, scratch_aux, sp,
                                                                                 p.lanes_aux, p.team_smem_plan, counters, world_chunk(c.n_worlds, p.k_plan.grid));
}

// K3 for worlds [w0, w0 + n_group): n_group <= launch_plan::group_worlds (one record slab each); timer may be null
void B2B_CAT(launch_solve_group_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, const launch_plan& p, uint16_t* plan,
                                              unsigned char* scratch_solve, unsigned char* scratch_sprof, const step_params& sp,
                                              int first_step, int w0, int n_group, cudaStream_t st, step_timer* tm)
{
    const batch_ptrs<real_t> B = as_ptrs(fields);
    B2B_CAT(prepare_kernel_, B2B_SUFFIX)<<<flat_blocks((long long)n_group * c.nc, p.n_sm), 128, 0, st>>>(B, c, p.plan, plan, p.layout, scratch_solve, sp,
                                                                                                         first_step, w0, n_group);
    if (tm) tm->mark(T_PREPARE, st);
    if (sp.solve_ctr) cudaMemsetAsync(sp.solve_ctr, 0, 3 * sizeof(unsigned int), st);
    const int per_cta = (p.k_solve.threads / 32) * (32 / p.lanes);       // worlds a CTA holds at a time
    int grid = p.k_solve.grid;
    if ((long long)grid * per_cta > n_group) grid = (n_group + per_cta - 1) / per_cta;
    // (sizing the grid so that the last round of worlds is as full as the others was measured slower: 34.5 vs
    // 33.5 ms per 131,072 worlds -- the stragglers of a ragged last round run faster on a nearly empty SM)
    const int threads = p.k_solve.threads;
    if (c.nj > 0) {
        // joints: island after island with the joint passes between the contact sweeps (no joint island schedule in such a batch)
        step_params spj = sp;
        if (first_step) spj.flags |= STEP_FIRST;
        const bool revolute_only = c.joint_types == (1 << 1);          // every joint of the batch is a revolute joint
        if (revolute_only)
            B2B_CAT(solve_velocity_revolute_kernel_, B2B_SUFFIX)<<<grid, threads, p.k_solve.smem, st>>>(B, c, p.plan, plan, p.layout, scratch_solve,
                                                                                                    scratch_sprof, spj, p.lanes, p.team_smem, w0, n_group);
        else
            B2B_CAT(solve_velocity_joints_kernel_, B2B_SUFFIX)<<<grid, threads, p.k_solve.smem, st>>>(B, c, p.plan, plan, p.layout, scratch_solve,
                                                                                                  scratch_sprof, spj, p.lanes, p.team_smem, w0, n_group);
        if (tm) tm->mark(T_VELOCITY, st);
        if (revolute_only)
            B2B_CAT(solve_position_revolute_kernel_, B2B_SUFFIX)<<<grid, threads, p.k_solve.smem, st>>>(B, c, p.plan, plan, p.layout, scratch_solve,
                                                                                                    scratch_sprof, spj, p.lanes, p.team_smem, w0, n_group);
        else
            B2B_CAT(solve_position_joints_kernel_, B2B_SUFFIX)<<<grid, threads, p.k_solve.smem, st>>>(B, c, p.plan, plan, p.layout, scratch_solve,
                                                                                                  scratch_sprof, spj, p.lanes, p.team_smem, w0, n_group);
        if (tm) tm->mark(T_POSITION, st);
        return;
    }
    B2B_CAT(solve_velocity_kernel_, B2B_SUFFIX)<<<grid, threads, p.k_solve.smem, st>>>(B, c, p.plan, plan, p.layout, scratch_solve,
                                                                                   scratch_sprof, sp, p.lanes, p.team_smem, w0, n_group);
    if (tm) tm->mark(T_VELOCITY, st);
    B2B_CAT(solve_position_kernel_, B2B_SUFFIX)<<<grid, threads, p.k_solve.smem, st>>>(B, c, p.plan, plan, p.layout, scratch_solve,
                                                                                   scratch_sprof, sp, p.lanes, p.team_smem, w0, n_group);
    B2B_CAT(solve_position_joint_kernel_, B2B_SUFFIX)<<<grid, threads, p.k_solve.smem, st>>>(B, c, p.plan, plan, p.layout, scratch_solve,
                                                                                         scratch_sprof, sp, p.lanes, p.team_smem, w0, n_group);
    if (tm) tm->mark(T_POSITION, st);
}

// K4 (proxy synchronisation, clear forces) + K5/K6 (pairs, new contacts) for every world
void B2B_CAT(launch_post_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, const launch_plan& p, uint16_t* plan,
                                       unsigned char* scratch_aux, unsigned int* counters, const step_params& sp, cudaStream_t st, step_timer* tm)
{
    const batch_ptrs<real_t> B = as_ptrs(fields);
    B2B_CAT(sync_kernel_, B2B_SUFFIX)<<<flat_blocks((long long)c.n_worlds * (c.nf > c.nb ? c.nf : c.nb), p.n_sm), 128, 0, st>>>(B, c, p.plan, plan, sp);
    if (tm) tm->mark(T_SYNC, st);
    B2B_CAT(broadphase_kernel_, B2B_SUFFIX)<<<p.k_broad.grid, 32, p.k_broad.smem, st>>>(B, c, p.plan, plan, p.layout, scratch_aux, sp,
                                                                                         p.lanes_aux, p.team_smem_broad, counters + 1, world_chunk(c.n_worlds, p.k_broad.grid));
    if (sp.dt > 0.0)
        B2B_CAT(toi_kernel_, B2B_SUFFIX)<<<p.k_toi.grid, 32, p.k_toi.smem, st>>>(B, c, p.plan, plan, p.layout, scratch_aux, sp,
                                                                              B2B_TOI_LANES, p.team_smem_toi, counters + 4, world_chunk(c.n_worlds, p.k_toi.grid));
    if (tm) tm->mark(T_PAIRS, st);
}

void B2B_CAT(launch_step_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, const launch_plan& p, uint16_t* plan,
                                       unsigned char* scratch_aux, unsigned char* scratch_solve, unsigned char* scratch_sprof,
                                       unsigned int* counters, const step_params& sp_in, int first_step, cudaStream_t st, step_timer* tm)
{
    step_params sp = sp_in;
    sp.joint_any = counters + 2;
    if (tm) tm->mark(T_START, st);
    if (sp.flags & STEP_FLAT_COLLIDE) B2B_CAT(launch_collide_, B2B_SUFFIX)(fields, c, p, plan, 0, c.n_worlds, st);
    if (tm) tm->mark(T_COLLIDE, st);
    B2B_CAT(launch_plan_, B2B_SUFFIX)(fields, c, p, plan, scratch_aux, counters, sp, st);
    if (tm) tm->mark(T_PLAN, st);
    if (sp.dt > 0.0) {
        // groups of worlds: the record slab is sized for one group (launch_plan::group_worlds)
        for (int w0 = 0; w0 < c.n_worlds; w0 += p.group_worlds) {
            const int n_group = (c.n_worlds - w0 < p.group_worlds) ? c.n_worlds - w0 : p.group_worlds;
            B2B_CAT(launch_solve_group_, B2B_SUFFIX)(fields, c, p, plan, scratch_solve, scratch_sprof, sp, first_step, w0, n_group, st, tm);
        }
    }
    B2B_CAT(launch_post_, B2B_SUFFIX)(fields, c, p, plan, scratch_aux, counters, sp, st, tm);
}

void B2B_CAT(launch_init_proxies_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, cudaStream_t st)
{
    long long total = (long long)c.n_worlds * c.nf;
    int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    if (grid < 1) grid = 1;
    B2B_CAT(init_proxies_kernel_, B2B_SUFFIX)<<<grid, 256, 0, st>>>(as_ptrs(fields), c);
}

void B2B_CAT(launch_collide_pairs_, B2B_SUFFIX)(const void* const* fields, int n, const void* sa, const void* sb, const void* xfa,
                                                const void* xfb, void* mtype, void* mf, void* pts, void* id, cudaStream_t st)
{
    int grid = (n + 127) / 128;
    if (grid < 1) grid = 1;
    B2B_CAT(collide_pairs_kernel_, B2B_SUFFIX)<<<grid, 128, 0, st>>>(as_ptrs(fields), n, (const int32_t*)sa, (const int32_t*)sb,
                                                                      (const real_t*)xfa, (const real_t*)xfb, (int32_t*)mtype,
                                                                      (real_t*)mf, (real_t*)pts, (uint32_t*)id);
}

void B2B_CAT(launch_scatter_forces_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, const void* stage, int w0, int n, cudaStream_t st)
{
    long long blocks = ((long long)n * c.nb + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    B2B_CAT(scatter_forces_kernel_, B2B_SUFFIX)<<<(int)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(as_ptrs(fields), c, (const real_t*)stage, w0, n);
}

void B2B_CAT(launch_gather_state_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, void* stage, int w0, int n, cudaStream_t st)
{
    long long blocks = ((long long)n * c.nb + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    B2B_CAT(gather_state_kernel_, B2B_SUFFIX)<<<(int)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(as_ptrs(fields), c, (real_t*)stage, w0, n);
}

void B2B_CAT(launch_checksum_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, double* out, cudaStream_t st)
{
    int grid = (c.n_worlds + 127) / 128;
    B2B_CAT(checksum_kernel_, B2B_SUFFIX)<<<grid, 128, 0, st>>>(as_ptrs(fields), c, out);
}

void B2B_CAT(launch_status_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, uint32_t* out, cudaStream_t st)
{
    batch_ptrs<real_t> B = as_ptrs(fields);
    int grid = (c.n_worlds + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    B2B_CAT(status_kernel_, B2B_SUFFIX)<<<grid, 256, 0, st>>>(B.w_status, c.n_worlds, out);
}

void B2B_CAT(launch_query_aabb_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, int n_q, const void* q_world, const void* q_box,
                                             int max_hits, void* out_count, void* out_fix, cudaStream_t st)
{
    const int grid = (n_q + 3) / 4 < 148 * 16 ? (n_q + 3) / 4 : 148 * 16;
    B2B_CAT(query_aabb_kernel_, B2B_SUFFIX)<<<grid < 1 ? 1 : grid, 128, 0, st>>>(as_ptrs(fields), c, n_q, (const int32_t*)q_world, (const real_t*)q_box,
                                                                                 max_hits, (int32_t*)out_count, (int32_t*)out_fix);
}

void B2B_CAT(launch_ray_cast_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, int n_q, const void* q_world, const void* q_seg, int mode,
                                           int max_hits, void* out_count, void* out_fix, void* out_hit, cudaStream_t st)
{
    const int grid = (n_q + 3) / 4 < 148 * 16 ? (n_q + 3) / 4 : 148 * 16;
    B2B_CAT(ray_cast_kernel_, B2B_SUFFIX)<<<grid < 1 ? 1 : grid, 128, 0, st>>>(as_ptrs(fields), c, n_q, (const int32_t*)q_world, (const real_t*)q_seg,
                                                                               mode, max_hits, (int32_t*)out_count, (int32_t*)out_fix, (real_t*)out_hit);
}

void B2B_CAT(launch_test_points_, B2B_SUFFIX)(const void* const* fields, const caps_t& c, int n, const void* q_world, const void* q_fix,
                                              const void* q_pt, void* out, cudaStream_t st)
{
    const int grid = (n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8;
    B2B_CAT(test_points_kernel_, B2B_SUFFIX)<<<grid < 1 ? 1 : grid, 256, 0, st>>>(as_ptrs(fields), c, n, (const int32_t*)q_world, (const int32_t*)q_fix,
                                                                                  (const real_t*)q_pt, (uint8_t*)out);
}

}  // namespace b2b

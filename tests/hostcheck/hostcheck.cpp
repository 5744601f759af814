// hostcheck.cpp -- TEST-ONLY host instantiation of the device step source.
//
// Compiles pypybox2d_b200/csrc/b2b_step.cuh with g++ (the parallel-loop macros expand to
// sequential loops) so that the exact kernel source can be debugged against the oracle in the
// CPU-only build container.  It is built by tests/test_hostcheck.py into tests/hostcheck/ and is
// never part of libb2b.so, never imported by the package, never timed.
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "../../include/b2b.h"
#include "../../pypybox2d_b200/csrc/b2b_step.cuh"

using namespace b2b;

static long long g_joint_world_steps = 0;     // world-steps whose plan row carried a joint island schedule (test bookkeeping)
static long long g_replans = 0;               // world-steps that entered with invalid island lists (a full replan follows)

template <class real>
static int run(const b2b_caps* caps, const b2b_buffers* buf, double dt, double inv_dt0, int vel_iters, int pos_iters,
               int n_steps)
{
    caps_t c;
    c.n_worlds = caps->n_worlds; c.nb = caps->nb; c.nf = caps->nf; c.nc = caps->nc; c.npair = caps->npair;
    c.nl = caps->nl; c.n_shapes = caps->n_shapes; c.nj = caps->nj; c.joint_types = caps->joint_types;
    batch_ptrs<real> B;
    static_assert(sizeof(B) == sizeof(*buf), "b2b_buffers and batch_ptrs must have the same members");
    std::memcpy(&B, buf, sizeof(B));
    scratch_layout L = scratch_plan(c, (int)sizeof(real));
    plan_layout P = plan_plan(c);
    smem_layout S = smem_plan_layout(c, (int)sizeof(real));
    std::vector<unsigned char> aux(L.total_aux + 256), solve(L.total_rec + 256), sprof(L.total_sprof + 256);
    std::vector<unsigned char> smem_aux((S.total_plan > S.total_broad ? S.total_plan : S.total_broad) + 256), smem_solve(S.total_solve + 256);
    // the plan rows persist between calls like the device context's do (the island schedule is reused while the
    // contact graph is unchanged); keyed by the batch they belong to
    static std::map<const void*, std::vector<uint16_t>> plans;
    std::vector<uint16_t>& plan = plans[buf->w_count];
    if (plan.size() != P.stride * (size_t)c.n_worlds + 16) {
        plan.assign(P.stride * (size_t)c.n_worlds + 16, 0);
        for (int w = 0; w < c.n_worlds; ++w) B.w_flags[w] &= ~WORLD_GRAPH_VALID;
    }
    step_params sp;
    sp.dt = dt; sp.dt_ratio_first = inv_dt0 * dt; sp.dt_ratio_rest = (dt > 0.0 ? (1.0 / dt) * dt : 0.0);
    sp.vel_iters = vel_iters; sp.pos_iters = pos_iters; sp.n_steps = n_steps; sp.flags = STEP_FLAT_COLLIDE;
    sp.solve_lanes = 8;
    sp.zero = 0;
    sp.joint_any = nullptr;
    for (int w = 0; w < c.n_worlds; ++w) {
        world_step<real> ws;
        ws.bind_team_cta();
        ws.bind_rows(B, c, w, plan.data(), P);
        ws.bind_scratch_aux(aux.data(), L);
        ws.bind_scratch_solve(solve.data(), 0, L, sprof.data());
        ws.bind_smem_aux(smem_aux.data(), true);      // the pair kernel's layout (candidate bitmap included): the phases run one after the other here
        ws.bind_smem_solve(smem_solve.data());
        if (dt > 0.0 && !(B.w_flags[w] & WORLD_GRAPH_VALID)) g_replans += 1;
        ws.run(sp);
        if (dt > 0.0 && ws.g_joint[0] != 0 && ws.w_island[2] > 0) g_joint_world_steps += n_steps;
    }
    return 0;
}

extern "C" long long hc_joint_world_steps(int reset)
{
    const long long v = g_joint_world_steps;
    if (reset) g_joint_world_steps = 0;
    return v;
}

extern "C" long long hc_replans(int reset)
{
    const long long v = g_replans;
    if (reset) g_replans = 0;
    return v;
}

extern "C" int hc_step(const b2b_caps* caps, const b2b_buffers* buf, double dt, double inv_dt0, int vel_iters,
                       int pos_iters, int n_steps)
{
    if (caps->real_bytes == 8) return run<double>(caps, buf, dt, inv_dt0, vel_iters, pos_iters, n_steps);
    if (caps->real_bytes == 4) return run<float>(caps, buf, dt, inv_dt0, vel_iters, pos_iters, n_steps);
    return -1;
}

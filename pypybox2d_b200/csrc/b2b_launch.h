// b2b_launch.h -- declarations shared by the per-precision kernel units and b2b_api.cpp.
#pragma once
#include <cuda_runtime.h>
#include <string.h>
#include <vector>
#include "b2b_step.cuh"

#define B2B_NFIELDS 45
#define B2B_CAT_(a, b) a##b
#define B2B_CAT(a, b) B2B_CAT_(a, b)

namespace b2b {

struct kernel_plan {
    int grid, threads, ctas_per_sm;
    size_t smem;
};
struct launch_plan {
    kernel_plan k_plan, k_solve, k_broad, k_toi;
    int aux_slabs, solve_slabs;      // scratch slabs to allocate: per plan/broadphase CTA; solve teams (cycle counters)
    int group_worlds;                // worlds whose constraint records are resident at once (one record slab each)
    int n_sm;
    int lanes, lanes_aux;            // lanes of a warp that step one world: solve kernel / plan + broadphase kernels
    size_t team_smem, team_smem_plan, team_smem_broad, team_smem_toi;
    scratch_layout layout;
    plan_layout plan;
};

// B2B_STEP_TIME: a chain of CUDA events on the launching stream; interval i ends at event i and belongs to kernel kind[i]
enum { T_START = -1, T_COLLIDE = 0, T_PLAN, T_PREPARE, T_VELOCITY, T_POSITION, T_SYNC, T_PAIRS, T_KINDS };
struct step_timer {
    std::vector<cudaEvent_t> ev;
    std::vector<int> kind;
    size_t used = 0;
    int steps = 0;
    void mark(int k, cudaStream_t st)
    {
        if (used == ev.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            ev.push_back(e); kind.push_back(0);
        }
        kind[used] = k;
        if (k == T_START) ++steps;
        cudaEventRecord(ev[used++], st);
    }
};

#define B2B_DECLARE(S)                                                                                                        \
    int plan_##S(const caps_t& c, int lanes, int lanes_aux, int ctas_per_sm, int n_sm, launch_plan* out);                     \
    void launch_step_##S(const void* const* fields, const caps_t& c, const launch_plan& p, uint16_t* plan,                    \
                         unsigned char* scratch_aux, unsigned char* scratch_solve, unsigned char* scratch_sprof,            \
                         unsigned int* counters, const step_params& sp, int first_step, cudaStream_t st, step_timer* tm);      \
    void launch_collide_##S(const void* const* fields, const caps_t& c, const launch_plan& p, uint16_t* plan, int w0, int n, cudaStream_t st); \
    void launch_plan_##S(const void* const* fields, const caps_t& c, const launch_plan& p, uint16_t* plan, unsigned char* scratch_aux,   \
                         unsigned int* counters, const step_params& sp, cudaStream_t st);                                     \
    void launch_solve_group_##S(const void* const* fields, const caps_t& c, const launch_plan& p, uint16_t* plan,               \
                                unsigned char* scratch_solve, unsigned char* scratch_sprof, const step_params& sp, int first_step, \
                                int w0, int n_group, cudaStream_t st, step_timer* tm);                                        \
    void launch_post_##S(const void* const* fields, const caps_t& c, const launch_plan& p, uint16_t* plan, unsigned char* scratch_aux,   \
                         unsigned int* counters, const step_params& sp, cudaStream_t st, step_timer* tm);                     \
    void launch_init_proxies_##S(const void* const* fields, const caps_t& c, cudaStream_t st);                                \
    void launch_collide_pairs_##S(const void* const* fields, int n, const void* sa, const void* sb, const void* xfa,          \
                                  const void* xfb, void* mtype, void* mf, void* pts, void* id, cudaStream_t st);             \
    void launch_scatter_forces_##S(const void* const* fields, const caps_t& c, const void* stage, int w0, int n, cudaStream_t st); \
    void launch_gather_state_##S(const void* const* fields, const caps_t& c, void* stage, int w0, int n, cudaStream_t st);    \
    void launch_checksum_##S(const void* const* fields, const caps_t& c, double* out, cudaStream_t st);                       \
    void launch_status_##S(const void* const* fields, const caps_t& c, uint32_t* out, cudaStream_t st);                         \
    void launch_query_aabb_##S(const void* const* fields, const caps_t& c, int n_q, const void* q_world, const void* q_box,   \
                               int max_hits, void* out_count, void* out_fix, cudaStream_t st);                                \
    void launch_ray_cast_##S(const void* const* fields, const caps_t& c, int n_q, const void* q_world, const void* q_seg, int mode, \
                             int max_hits, void* out_count, void* out_fix, void* out_hit, cudaStream_t st);                   \
    void launch_test_points_##S(const void* const* fields, const caps_t& c, int n, const void* q_world, const void* q_fix,    \
                                const void* q_pt, void* out, cudaStream_t st);

B2B_DECLARE(f32)
B2B_DECLARE(f64)

}  // namespace b2b

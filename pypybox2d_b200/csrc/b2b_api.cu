// b2b_api.cu -- the C-ABI of libb2b.so (include/b2b.h): context, buffer binding, launches.
// No physics here; the kernels live in b2b_f32.cu / b2b_f64.cu.  No CPU fallback exists: every entry point
// that computes anything launches a CUDA kernel or fails with B2B_ERR_CUDA.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/b2b.h"
#include "b2b_launch.h"

using namespace b2b;

struct b2b_ctx {
    b2b_caps caps;
    caps_t c;
    bool f64;
    int n_sm;
    launch_plan plan;
    unsigned char* scratch;        // plan / broadphase kernels: one slab per CTA
    unsigned char* scratch_solve;  // constraint records: one slab per world of a solve group
    unsigned char* scratch_sprof;  // solve kernel: cycle counters, one row per team
    uint16_t* plan_rows;
    uint32_t* d_status;            // [0] status OR, [1..2] chunk counters of the plan / pair kernels, [3..4] joint-schedule words, [8..15] world counters of the solve kernels (two sets)
    void* stage_in;       // [n_worlds][nb][3] reals
    void* stage_out;      // [n_worlds][nb][6] reals
    const void* fields[B2B_NFIELDS];
    bool bound;
    long long launches;
    step_timer* timer;             // B2B_STEP_TIME: event chain of the timed steps
    // b2b_step_host: copy streams and the events that chain copies and kernels chunk by chunk
    cudaStream_t s_in, s_out;
    // b2b_step: second stream for every other chunk of worlds of the solve kernels (their tails overlap), fork / join events
    cudaStream_t s_aux;
    cudaEvent_t ev_fork, ev_join;
    int step_streams;              // b2b_step: 1 (default): everything on the caller's stream; 2: the solve chunks alternate between two streams
    int pipe_streams;              // b2b_step_host: 2 (default): the chunks' kernels alternate between two streams (the flat kernels of one chunk fill the
                                   // tail of the previous chunk's persistent solve grids); 1: one stream
    std::vector<cudaEvent_t>* pipe_ev;
    int pipe_worlds;               // worlds per chunk of the host-buffer pipeline
    int stream_worlds;             // worlds per chunk of b2b_step's two-stream schedule
};

static thread_local std::string g_err;

static int fail(int code, const char* what, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof(buf), "%s", what);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr, what)                                              \
    do {                                                                  \
        cudaError_t e_ = (expr);                                          \
        if (e_ != cudaSuccess) return fail(B2B_ERR_CUDA, what, e_);       \
    } while (0)

static bool caps_ok(const b2b_caps* k)
{
    return k && k->n_worlds > 0 && k->nb > 0 && k->nf > 0 && k->nc > 0 && k->npair > 0 && k->nl >= k->nb && k->n_shapes > 0 &&
           (k->real_bytes == 4 || k->real_bytes == 8) && (k->threads == 0 || k->threads == 2 || k->threads == 4 || k->threads == 8 || k->threads == 16 || k->threads == 32) &&
           (k->reserved == 0 || k->reserved == 4 || k->reserved == 8 || k->reserved == 16 || k->reserved == 32) && k->nb < 65535 &&
           k->nc < 65535 && k->nj >= 0 && k->nj < 65535;
}

static caps_t to_caps(const b2b_caps* k)
{
    caps_t c;
    c.n_worlds = k->n_worlds; c.nb = k->nb; c.nf = k->nf; c.nc = k->nc; c.npair = k->npair; c.nl = k->nl; c.n_shapes = k->n_shapes;
    c.nj = k->nj;
    c.joint_types = k->joint_types;
    return c;
}

extern "C" const char* b2b_last_error(void) { return g_err.c_str(); }
extern "C" const char* b2b_version(void) { return "b2b 0.1 (sm_100a)"; }
extern "C" const char* b2b_field_names(void)
{
    return "w_count,w_flags,w_status,w_gravity,w_island,"
           "b_xf,b_pos,b_vel,b_pos0,b_force,b_mass,b_damp,b_flags,b_island,"
           "f_body,f_shape,f_mat,f_filter,f_fat,f_tight,f_moved,"
           "c_fix,c_flags,c_mat,c_mtype,c_mf,c_pts,c_imp,c_id,"
           "p_pairs,l_body,l_contact,l_start,s_meta,s_radius,s_data,"
           "w_joint,j_def,j_par,j_imp,j_tmp,j_adj,b_jadj,l_joint,l_jstart";
}

extern "C" size_t b2b_scratch_bytes(const b2b_caps* caps)
{
    if (!caps_ok(caps)) return 0;
    const scratch_layout L = scratch_plan(to_caps(caps), caps->real_bytes);
    return L.total_aux + L.total_rec;        // per resident CTA of the plan kernels / per world of a solve group
}

extern "C" void b2b_destroy(b2b_ctx* x);

extern "C" int b2b_create(const b2b_caps* caps, b2b_ctx** out)
{
    if (!out) return fail(B2B_ERR_ARG, "b2b_create: out is NULL");
    *out = nullptr;
    if (!caps_ok(caps)) return fail(B2B_ERR_ARG, "b2b_create: bad capacities");
    int n_dev = 0;
    CUDA_TRY(cudaGetDeviceCount(&n_dev), "cudaGetDeviceCount (is a CUDA device visible?)");
    if (caps->device < 0 || caps->device >= n_dev) return fail(B2B_ERR_ARG, "b2b_create: no such device");
    CUDA_TRY(cudaSetDevice(caps->device), "cudaSetDevice");
    b2b_ctx* x = new b2b_ctx();
    memset(x, 0, sizeof(*x));
    x->caps = *caps;
    x->c = to_caps(caps);
    x->f64 = caps->real_bytes == 8;
    // every failure below leaves through b2b_destroy, which frees whatever was allocated so far
#define CREATE_TRY(expr, what)                                                 \
    do {                                                                       \
        cudaError_t e_ = (expr);                                               \
        if (e_ != cudaSuccess) { b2b_destroy(x); return fail(B2B_ERR_CUDA, what, e_); } \
    } while (0)
    cudaDeviceProp prop;
    CREATE_TRY(cudaGetDeviceProperties(&prop, caps->device), "cudaGetDeviceProperties");
    x->n_sm = prop.multiProcessorCount;
    // caps->threads: lanes of a warp per world in the solve kernel (4, 8, 16 or 32); caps->reserved: the same for
    // the plan / broadphase kernels
    const int lanes = caps->threads > 0 ? caps->threads : 4;
    // default team size of the plan / pair kernels: the whole warp per world.  Narrow teams put several worlds' SEQUENTIAL island
    // walks into one warp, where they diverge and are issued one after the other, and multiply the CTA's shared memory.
    // Measured on B200 with the per-grab world count sized from the batch (b2b_kernels.inl world_chunk), plan kernel with
    // 4 / 8 / 16 / 32 lanes per world: deep stacks x 32,768 worlds 24.2 / 12.6 / 7.4 / 5.3 ms, vertical stacks x 65,536
    // 39.0 / 21.2 / 10.7 / 8.7 ms, mixed x 32,768 51.8 / 22.5 / 11.9 / 6.4 ms; pyramids from t = 0 (65,536 worlds) 24.8 ms per
    // step with 8 lanes, 24.2 with 32 (profiles/r2_other_scenes.md)
    int lanes_aux = caps->reserved > 0 ? caps->reserved : 32;
    if (const char* env = getenv("B2B_LANES_AUX")) { const int v = atoi(env); if (v == 4 || v == 8 || v == 16 || v == 32) lanes_aux = v; }   // experiments
    int rc = x->f64 ? plan_f64(x->c, lanes, lanes_aux, caps->ctas_per_sm, x->n_sm, &x->plan)
                    : plan_f32(x->c, lanes, lanes_aux, caps->ctas_per_sm, x->n_sm, &x->plan);
    if (rc != 0) { b2b_destroy(x); return fail(B2B_ERR_CUDA, "kernel launch planning (shared memory / occupancy)", (cudaError_t)rc); }
    const size_t aux_bytes = x->plan.layout.total_aux * (size_t)x->plan.aux_slabs;
    CREATE_TRY(cudaMalloc((void**)&x->scratch, aux_bytes), "cudaMalloc(scratch)");
    CREATE_TRY(cudaMemset(x->scratch, 0, aux_bytes), "cudaMemset(scratch)");
    // Constraint records live for one group of worlds at a time (prepare kernel -> solve kernel).  The group is
    // every world when that fits the record budget (B2B_RECORD_GIB, default 24 GiB, at most 40 % of the free
    // memory), else a whole number of resident-team waves.
    {
        size_t free_b = 0, total_b = 0;
        CREATE_TRY(cudaMemGetInfo(&free_b, &total_b), "cudaMemGetInfo");
        double budget = 24.0;
        if (const char* env = getenv("B2B_RECORD_GIB")) { const double v = atof(env); if (v > 0.0) budget = v; }
        size_t budget_b = (size_t)(budget * 1073741824.0);
        if (budget_b > free_b / 5 * 2) budget_b = free_b / 5 * 2;
        const size_t per_world = x->plan.layout.total_rec;
        long long group = (long long)(budget_b / per_world);
        const long long wave = x->plan.solve_slabs > 0 ? x->plan.solve_slabs : 1;
        if (group >= x->c.n_worlds) group = x->c.n_worlds;
        else if (group > wave) group = group / wave * wave;
        if (group < 1) group = 1;
        x->plan.group_worlds = (int)group;
        // host-buffer pipeline (b2b_step_host): about B2B_PIPE_CHUNKS chunks (default 12; worlds are handed to the solve warps
        // through a counter, so a chunk need not be a whole number of waves).  Measured on B200, 131,072 pyramids, two compute
        // streams: 44.3 / 42.4 / 42.6 / 41.0 ms per call with 4 / 6 / 8 / 12 chunks (device-resident step: 36.9 ms)
        int chunks = 12;
        if (const char* env = getenv("B2B_PIPE_CHUNKS")) { const int v = atoi(env); if (v > 0) chunks = v; }
        long long pipe = ((long long)x->c.n_worlds + chunks - 1) / chunks;
        if (pipe > group) pipe = group;
        if (pipe < 1) pipe = 1;
        x->pipe_worlds = (int)pipe;
        // b2b_step's two-stream schedule: about B2B_STREAM_CHUNKS chunks (default 4; measured on B200, 131,072 pyramids:
        // 43.07 ms one stream, 42.78 / 41.95 / 42.64 / 42.63 ms with 3 / 4 / 6 / 11 chunks)
        int schunks = 4;
        if (const char* env = getenv("B2B_STREAM_CHUNKS")) { const int v = atoi(env); if (v > 0) schunks = v; }
        long long sw = ((long long)x->c.n_worlds + schunks - 1) / schunks;
        if (sw > wave) sw = (sw + wave - 1) / wave * wave;
        if (sw > group) sw = group;
        if (sw < 1) sw = 1;
        x->stream_worlds = (int)sw;
    }
    CREATE_TRY(cudaMalloc((void**)&x->scratch_solve, x->plan.layout.total_rec * (size_t)x->plan.group_worlds), "cudaMalloc(constraint records)");
    const size_t sprof_bytes = x->plan.layout.total_sprof * (size_t)x->plan.solve_slabs;
    CREATE_TRY(cudaMalloc((void**)&x->scratch_sprof, sprof_bytes), "cudaMalloc(solve counters)");
    CREATE_TRY(cudaMemset(x->scratch_sprof, 0, sprof_bytes), "cudaMemset(solve counters)");
    const size_t plan_bytes = sizeof(uint16_t) * x->plan.plan.stride * (size_t)x->c.n_worlds;
    CREATE_TRY(cudaMalloc((void**)&x->plan_rows, plan_bytes), "cudaMalloc(plan rows)");
    CREATE_TRY(cudaMemset(x->plan_rows, 0, plan_bytes), "cudaMemset(plan rows)");
    CREATE_TRY(cudaMalloc((void**)&x->d_status, 16 * sizeof(uint32_t)), "cudaMalloc(status)");
    CREATE_TRY(cudaMemset(x->d_status, 0, 16 * sizeof(uint32_t)), "cudaMemset(status)");
    // (round 2, one solve CTA per SM with worlds handed out through a counter: a launch has no ragged last round any more and two
    // solve grids cannot share an SM's shared memory, so chunks on two streams only add tails: 39.9 ms vs 38.2 ms on one stream)
    x->step_streams = 1;
    if (const char* env = getenv("B2B_STEP_STREAMS")) { const int v = atoi(env); if (v == 1 || v == 2) x->step_streams = v; }
    x->pipe_streams = 2;
    if (const char* env = getenv("B2B_PIPE_STREAMS")) { const int v = atoi(env); if (v == 1 || v == 2) x->pipe_streams = v; }
#undef CREATE_TRY
    *out = x;
    return B2B_OK;
}

extern "C" void b2b_destroy(b2b_ctx* x)
{
    if (!x) return;
    cudaSetDevice(x->caps.device);
    if (x->scratch) cudaFree(x->scratch);
    if (x->scratch_solve) cudaFree(x->scratch_solve);
    if (x->scratch_sprof) cudaFree(x->scratch_sprof);
    if (x->plan_rows) cudaFree(x->plan_rows);
    if (x->d_status) cudaFree(x->d_status);
    if (x->stage_in) cudaFree(x->stage_in);
    if (x->stage_out) cudaFree(x->stage_out);
    if (x->timer) { for (cudaEvent_t e : x->timer->ev) cudaEventDestroy(e); delete x->timer; }
    if (x->pipe_ev) { for (cudaEvent_t e : *x->pipe_ev) cudaEventDestroy(e); delete x->pipe_ev; }
    if (x->ev_fork) cudaEventDestroy(x->ev_fork);
    if (x->ev_join) cudaEventDestroy(x->ev_join);
    if (x->s_aux) cudaStreamDestroy(x->s_aux);
    if (x->s_in) cudaStreamDestroy(x->s_in);
    if (x->s_out) cudaStreamDestroy(x->s_out);
    delete x;
}

extern "C" int b2b_bind(b2b_ctx* x, const b2b_buffers* buffers)
{
    if (!x || !buffers) return fail(B2B_ERR_ARG, "b2b_bind: NULL argument");
    static_assert(sizeof(b2b_buffers) == sizeof(void*) * B2B_NFIELDS, "b2b_buffers layout");
    memcpy(x->fields, buffers, sizeof(b2b_buffers));
    for (int i = 0; i < B2B_NFIELDS; ++i)
        if (!x->fields[i]) return fail(B2B_ERR_ARG, "b2b_bind: a buffer pointer is NULL");
    x->bound = true;
    return B2B_OK;
}

static int check_launch(b2b_ctx* x, const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B2B_ERR_CUDA, what, e);
    ++x->launches;
    return B2B_OK;
}

extern "C" int b2b_init_proxies(b2b_ctx* x, void* stream)
{
    if (!x || !x->bound) return fail(B2B_ERR_UNBOUND, "b2b_init_proxies: no buffers bound");
    CUDA_TRY(cudaSetDevice(x->caps.device), "cudaSetDevice");
    if (x->f64) launch_init_proxies_f64(x->fields, x->c, (cudaStream_t)stream);
    else launch_init_proxies_f32(x->fields, x->c, (cudaStream_t)stream);
    return check_launch(x, "init_proxies kernel");
}

static step_params make_params(double dt, double inv_dt0, int vel_iters, int pos_iters, int n_steps, uint32_t flags, int solve_lanes)
{
    step_params sp;
    sp.dt = dt;
    sp.dt_ratio_first = inv_dt0 * dt;                              // world.py:328
    sp.dt_ratio_rest = (dt > 0.0) ? (1.0 / dt) * dt : 0.0;         // _inv_dt0 = inv_dt after the first step (world.py:348-349)
    sp.vel_iters = vel_iters; sp.pos_iters = pos_iters; sp.n_steps = n_steps; sp.flags = flags;
    sp.solve_lanes = solve_lanes;
    sp.zero = 0;
    sp.joint_any = nullptr;                                        // bound by launch_step
    sp.solve_ctr = nullptr;                                        // bound by the callers (solve_counters)
    return sp;
}

// world counters of the solve kernels for the launches on stream `parity` (B2B_STATIC_WORLDS=1: none, worlds strided over the warps)
static unsigned int* solve_counters(const b2b_ctx* x, int parity)
{
    static const bool off = [] { const char* e = getenv("B2B_STATIC_WORLDS"); return e && atoi(e) != 0; }();
    return off ? nullptr : x->d_status + 8 + 4 * (parity & 1);
}

static int step_launches(const b2b_ctx* x, double dt, uint32_t flags, int groups)
{
    // plan, sync, pairs, (collide), and per group of worlds prepare + velocity + the two position kernels (one of them
    // returns at once, see b2b_kernels.inl); the TOI kernel (worlds without continuous collision return at once)
    return 3 + ((dt > 0.0) ? 4 * groups + 1 : 0) + ((flags & 4u) ? 1 : 0);
}

extern "C" int b2b_step(b2b_ctx* x, double dt, double inv_dt0, int vel_iters, int pos_iters, int n_steps, uint32_t flags,
                        void* stream)
{
    if (!x || !x->bound) return fail(B2B_ERR_UNBOUND, "b2b_step: no buffers bound");
    if (vel_iters < 0 || pos_iters < 0 || n_steps < 0 || !(dt >= 0.0)) return fail(B2B_ERR_ARG, "b2b_step: bad argument");
    if (n_steps == 0) return B2B_OK;
    CUDA_TRY(cudaSetDevice(x->caps.device), "cudaSetDevice");
    // the flat narrowphase kernel is the default; B2B_STEP_NO_FLAT keeps the narrowphase inside the per-world plan kernel
    step_params sp = make_params(dt, inv_dt0, vel_iters, pos_iters, 1, (flags & B2B_STEP_NO_FLAT) ? (flags & ~4u) : (flags | 4u),
                                 x->plan.lanes);
    sp.solve_ctr = solve_counters(x, 0);
    step_timer* tm = nullptr;
    if (flags & B2B_STEP_TIME) {
        if (!x->timer) x->timer = new step_timer();
        tm = x->timer;
    }
    const int groups = (x->c.n_worlds + x->plan.group_worlds - 1) / x->plan.group_worlds;
    // Two-stream schedule: the solve kernels are persistent grids whose last round of worlds leaves most SMs idle for one
    // world's latency (~2 ms of a 43 ms step at 131,072 pyramids: 11.07 rounds).  With the batch cut into chunks of whole
    // rounds that alternate between the caller's stream and a second one, the tail of one chunk's kernel runs under the
    // next chunk's.  Needs every world's record slab resident at once; timing and profiling runs stay on one stream so
    // that the events between the launches bracket one kernel each.
    const int chunk = x->stream_worlds, n_chunks = (x->c.n_worlds + chunk - 1) / chunk;
    const bool two_streams = x->step_streams == 2 && n_chunks >= 2 && groups == 1 && !(flags & (B2B_STEP_TIME | B2B_STEP_PROFILE)) && dt > 0.0;
    if (two_streams) {
        cudaStream_t st = (cudaStream_t)stream;
        if (!x->s_aux) CUDA_TRY(cudaStreamCreateWithFlags(&x->s_aux, cudaStreamNonBlocking), "cudaStreamCreate(aux)");
        if (!x->ev_fork) CUDA_TRY(cudaEventCreateWithFlags(&x->ev_fork, cudaEventDisableTiming), "cudaEventCreate");
        if (!x->ev_join) CUDA_TRY(cudaEventCreateWithFlags(&x->ev_join, cudaEventDisableTiming), "cudaEventCreate");
        for (int s = 0; s < n_steps; ++s) {
            step_params sp2 = sp;
            sp2.joint_any = x->d_status + 3;
            if (sp.flags & 4u) {
                if (x->f64) launch_collide_f64(x->fields, x->c, x->plan, x->plan_rows, 0, x->c.n_worlds, st);
                else launch_collide_f32(x->fields, x->c, x->plan, x->plan_rows, 0, x->c.n_worlds, st);
            }
            if (x->f64) launch_plan_f64(x->fields, x->c, x->plan, x->plan_rows, x->scratch, x->d_status + 1, sp2, st);
            else launch_plan_f32(x->fields, x->c, x->plan, x->plan_rows, x->scratch, x->d_status + 1, sp2, st);
            CUDA_TRY(cudaMemsetAsync(x->d_status + 4, (sp.flags & B2B_STEP_GROUP_KERNEL) ? 1 : 0, sizeof(uint32_t), st), "memset");
            CUDA_TRY(cudaEventRecord(x->ev_fork, st), "cudaEventRecord");
            CUDA_TRY(cudaStreamWaitEvent(x->s_aux, x->ev_fork, 0), "cudaStreamWaitEvent");
            for (int k = 0; k < n_chunks; ++k) {
                const int w0 = k * chunk, n = (x->c.n_worlds - w0 < chunk) ? x->c.n_worlds - w0 : chunk;
                cudaStream_t sk = (k & 1) ? x->s_aux : st;
                step_params spk = sp2;
                spk.joint_any = x->d_status + ((k & 1) ? 4 : 3);         // one joint-schedule word per stream (position kernels read it)
                spk.solve_ctr = solve_counters(x, k & 1);
                unsigned char* rec = x->scratch_solve + (size_t)w0 * x->plan.layout.total_rec;
                if (x->f64) launch_solve_group_f64(x->fields, x->c, x->plan, x->plan_rows, rec, x->scratch_sprof, spk, s == 0, w0, n, sk, nullptr);
                else launch_solve_group_f32(x->fields, x->c, x->plan, x->plan_rows, rec, x->scratch_sprof, spk, s == 0, w0, n, sk, nullptr);
            }
            CUDA_TRY(cudaEventRecord(x->ev_join, x->s_aux), "cudaEventRecord");
            CUDA_TRY(cudaStreamWaitEvent(st, x->ev_join, 0), "cudaStreamWaitEvent");
            if (x->f64) launch_post_f64(x->fields, x->c, x->plan, x->plan_rows, x->scratch, x->d_status + 1, sp2, st, nullptr);
            else launch_post_f32(x->fields, x->c, x->plan, x->plan_rows, x->scratch, x->d_status + 1, sp2, st, nullptr);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return fail(B2B_ERR_CUDA, "step kernels", e);
            x->launches += step_launches(x, dt, sp.flags, n_chunks);
        }
        return B2B_OK;
    }
    for (int s = 0; s < n_steps; ++s) {
        if (x->f64) launch_step_f64(x->fields, x->c, x->plan, x->plan_rows, x->scratch, x->scratch_solve, x->scratch_sprof, x->d_status + 1, sp, s == 0, (cudaStream_t)stream, tm);
        else launch_step_f32(x->fields, x->c, x->plan, x->plan_rows, x->scratch, x->scratch_solve, x->scratch_sprof, x->d_status + 1, sp, s == 0, (cudaStream_t)stream, tm);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(B2B_ERR_CUDA, "step kernels", e);
        x->launches += step_launches(x, dt, sp.flags, groups);
    }
    return B2B_OK;
}

// Test hook (T1): World._solve's islands and Island.solve only.
extern "C" int b2b_solve_islands(b2b_ctx* x, double dt, double inv_dt0, int vel_iters, int pos_iters, uint32_t flags, void* stream)
{
    if (!x || !x->bound) return fail(B2B_ERR_UNBOUND, "b2b_solve_islands: no buffers bound");
    if (vel_iters < 0 || pos_iters < 0 || !(dt > 0.0)) return fail(B2B_ERR_ARG, "b2b_solve_islands: bad argument");
    CUDA_TRY(cudaSetDevice(x->caps.device), "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;
    step_params sp = make_params(dt, inv_dt0, vel_iters, pos_iters, 1, (flags & B2B_STEP_PROFILE) | 4u | 64u, x->plan.lanes);
    sp.joint_any = x->d_status + 3;
    sp.solve_ctr = solve_counters(x, 0);
    if (x->f64) launch_plan_f64(x->fields, x->c, x->plan, x->plan_rows, x->scratch, x->d_status + 1, sp, st);
    else launch_plan_f32(x->fields, x->c, x->plan, x->plan_rows, x->scratch, x->d_status + 1, sp, st);
    int groups = 0;
    for (int w0 = 0; w0 < x->c.n_worlds; w0 += x->plan.group_worlds, ++groups) {
        const int n = (x->c.n_worlds - w0 < x->plan.group_worlds) ? x->c.n_worlds - w0 : x->plan.group_worlds;
        if (x->f64) launch_solve_group_f64(x->fields, x->c, x->plan, x->plan_rows, x->scratch_solve, x->scratch_sprof, sp, 1, w0, n, st, nullptr);
        else launch_solve_group_f32(x->fields, x->c, x->plan, x->plan_rows, x->scratch_solve, x->scratch_sprof, sp, 1, w0, n, st, nullptr);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B2B_ERR_CUDA, "solve kernels", e);
    x->launches += 1 + 4 * groups;
    return B2B_OK;
}

// The end-to-end call.  One World.step with its inputs and outputs in HOST memory, as a pipeline over chunks of worlds
// (b2b_ctx::pipe_worlds, a whole number of resident waves of the solve kernels) on three streams:
//   copy-in stream   forces of chunk k host -> device
//   caller's stream  scatter + narrowphase of chunk k (waits for its forces) ... plan kernel (all worlds) ...
//                    prepare / velocity / position of chunk k, gather of chunk k ... proxy synchronisation + pairs
//   copy-out stream  body state of chunk k device -> host (waits for its gather)
// so the copies run under the kernels of the neighbouring chunks; only the first chunk's forces and the last chunk's state
// are exposed.  Forces are applied (and wake their bodies) before the chunk's narrowphase, as Body.apply_force before
// World.step does in the reference.  n_steps > 1: the forces act on the first step only when the world clears forces after
// a step (auto_clear_forces), on every step otherwise; the state read back is the last step's.
extern "C" int b2b_step_host(b2b_ctx* x, const void* h_force_torque, void* h_state_out, double dt, double inv_dt0,
                             int vel_iters, int pos_iters, int n_steps, uint32_t flags, void* stream)
{
    if (!x || !x->bound) return fail(B2B_ERR_UNBOUND, "b2b_step_host: no buffers bound");
    if (vel_iters < 0 || pos_iters < 0 || n_steps < 0 || !(dt >= 0.0)) return fail(B2B_ERR_ARG, "b2b_step_host: bad argument");
    CUDA_TRY(cudaSetDevice(x->caps.device), "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t rb = (size_t)x->caps.real_bytes, per_world_in = (size_t)x->c.nb * 3 * rb, per_world_out = (size_t)x->c.nb * 6 * rb;
    const int W = x->c.n_worlds, chunk = x->pipe_worlds, n_chunks = (W + chunk - 1) / chunk;
    if (h_force_torque && !x->stage_in) CUDA_TRY(cudaMalloc(&x->stage_in, (size_t)W * per_world_in), "cudaMalloc(stage_in)");
    if (h_state_out && !x->stage_out) CUDA_TRY(cudaMalloc(&x->stage_out, (size_t)W * per_world_out), "cudaMalloc(stage_out)");
    if (!x->s_in) CUDA_TRY(cudaStreamCreateWithFlags(&x->s_in, cudaStreamNonBlocking), "cudaStreamCreate(copy-in)");
    if (!x->s_out) CUDA_TRY(cudaStreamCreateWithFlags(&x->s_out, cudaStreamNonBlocking), "cudaStreamCreate(copy-out)");
    if (!x->pipe_ev) x->pipe_ev = new std::vector<cudaEvent_t>();
    while ((int)x->pipe_ev->size() < 2 * n_chunks + 1) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
        x->pipe_ev->push_back(e);
    }
    cudaEvent_t* ev_in = x->pipe_ev->data();
    cudaEvent_t* ev_out = ev_in + n_chunks;
    cudaEvent_t ev_begin = x->pipe_ev->data()[2 * n_chunks];
    const step_params sp0 = make_params(dt, inv_dt0, vel_iters, pos_iters, 1, (flags & B2B_STEP_NO_FLAT) ? (flags & ~4u) : (flags | 4u),
                                        x->plan.lanes);
    // the copy streams start behind whatever the caller queued on its stream before this call
    CUDA_TRY(cudaEventRecord(ev_begin, st), "cudaEventRecord");
    if (h_force_torque) {
        CUDA_TRY(cudaStreamWaitEvent(x->s_in, ev_begin, 0), "cudaStreamWaitEvent");
        for (int k = 0; k < n_chunks; ++k) {
            const int w0 = k * chunk, n = (W - w0 < chunk) ? W - w0 : chunk;
            CUDA_TRY(cudaMemcpyAsync((char*)x->stage_in + (size_t)w0 * per_world_in, (const char*)h_force_torque + (size_t)w0 * per_world_in,
                                     (size_t)n * per_world_in, cudaMemcpyHostToDevice, x->s_in), "H2D forces");
            CUDA_TRY(cudaEventRecord(ev_in[k], x->s_in), "cudaEventRecord");
        }
    }
    if (h_state_out) CUDA_TRY(cudaStreamWaitEvent(x->s_out, ev_begin, 0), "cudaStreamWaitEvent");
    for (int s = 0; s < n_steps; ++s) {
        step_params sp = sp0;
        sp.joint_any = x->d_status + 3;
        const bool first = (s == 0), last = (s + 1 == n_steps);
        // forces in + narrowphase, chunk by chunk
        for (int k = 0; k < n_chunks; ++k) {
            const int w0 = k * chunk, n = (W - w0 < chunk) ? W - w0 : chunk;
            if (first && h_force_torque) {
                CUDA_TRY(cudaStreamWaitEvent(st, ev_in[k], 0), "cudaStreamWaitEvent");
                if (x->f64) launch_scatter_forces_f64(x->fields, x->c, x->stage_in, w0, n, st);
                else launch_scatter_forces_f32(x->fields, x->c, x->stage_in, w0, n, st);
                ++x->launches;
            }
            if (sp.flags & 4u) {
                if (x->f64) launch_collide_f64(x->fields, x->c, x->plan, x->plan_rows, w0, n, st);
                else launch_collide_f32(x->fields, x->c, x->plan, x->plan_rows, w0, n, st);
                ++x->launches;
            }
        }
        if (x->f64) launch_plan_f64(x->fields, x->c, x->plan, x->plan_rows, x->scratch, x->d_status + 1, sp, st);
        else launch_plan_f32(x->fields, x->c, x->plan, x->plan_rows, x->scratch, x->d_status + 1, sp, st);
        ++x->launches;
        // solve + state out, chunk by chunk (the last step's state is the one read back).  When every world's record slab is
        // resident the chunks alternate between the caller's stream and a second one, as in b2b_step: the tail of one chunk's
        // persistent solve grid runs under the next chunk's kernels.
        const bool two = x->pipe_streams == 2 && n_chunks >= 2 && x->plan.group_worlds >= W && dt > 0.0 && !(flags & B2B_STEP_PROFILE);
        if (two) {
            if (!x->s_aux) CUDA_TRY(cudaStreamCreateWithFlags(&x->s_aux, cudaStreamNonBlocking), "cudaStreamCreate(aux)");
            if (!x->ev_fork) CUDA_TRY(cudaEventCreateWithFlags(&x->ev_fork, cudaEventDisableTiming), "cudaEventCreate");
            if (!x->ev_join) CUDA_TRY(cudaEventCreateWithFlags(&x->ev_join, cudaEventDisableTiming), "cudaEventCreate");
            CUDA_TRY(cudaMemsetAsync(x->d_status + 4, (sp.flags & B2B_STEP_GROUP_KERNEL) ? 1 : 0, sizeof(uint32_t), st), "memset");
            CUDA_TRY(cudaEventRecord(x->ev_fork, st), "cudaEventRecord");
            CUDA_TRY(cudaStreamWaitEvent(x->s_aux, x->ev_fork, 0), "cudaStreamWaitEvent");
        }
        for (int k = 0; k < n_chunks; ++k) {
            const int w0 = k * chunk, n = (W - w0 < chunk) ? W - w0 : chunk;
            cudaStream_t sk = (two && (k & 1)) ? x->s_aux : st;
            if (dt > 0.0) {
                step_params spk = sp;
                if (two) spk.joint_any = x->d_status + ((k & 1) ? 4 : 3);       // one joint-schedule word per stream
                spk.solve_ctr = solve_counters(x, two ? (k & 1) : 0);
                unsigned char* rec = x->scratch_solve + (two ? (size_t)w0 * x->plan.layout.total_rec : (size_t)0);
                if (x->f64) launch_solve_group_f64(x->fields, x->c, x->plan, x->plan_rows, rec, x->scratch_sprof, spk, first, w0, n, sk, nullptr);
                else launch_solve_group_f32(x->fields, x->c, x->plan, x->plan_rows, rec, x->scratch_sprof, spk, first, w0, n, sk, nullptr);
                x->launches += 4;
            }
            if (last && h_state_out && !(flags & B2B_STEP_CONTINUOUS)) {
                if (x->f64) launch_gather_state_f64(x->fields, x->c, x->stage_out, w0, n, sk);
                else launch_gather_state_f32(x->fields, x->c, x->stage_out, w0, n, sk);
                ++x->launches;
                CUDA_TRY(cudaEventRecord(ev_out[k], sk), "cudaEventRecord");
                CUDA_TRY(cudaStreamWaitEvent(x->s_out, ev_out[k], 0), "cudaStreamWaitEvent");
                CUDA_TRY(cudaMemcpyAsync((char*)h_state_out + (size_t)w0 * per_world_out, (const char*)x->stage_out + (size_t)w0 * per_world_out,
                                         (size_t)n * per_world_out, cudaMemcpyDeviceToHost, x->s_out), "D2H state");
            }
        }
        if (two) {
            CUDA_TRY(cudaEventRecord(x->ev_join, x->s_aux), "cudaEventRecord");
            CUDA_TRY(cudaStreamWaitEvent(st, x->ev_join, 0), "cudaStreamWaitEvent");
        }
        if (x->f64) launch_post_f64(x->fields, x->c, x->plan, x->plan_rows, x->scratch, x->d_status + 1, sp, st, nullptr);
        else launch_post_f32(x->fields, x->c, x->plan, x->plan_rows, x->scratch, x->d_status + 1, sp, st, nullptr);
        x->launches += (dt > 0.0) ? 3 : 2;
        // continuous collision moves bodies after the discrete solve (World.solve_toi runs in the last kernel): read back now
        if (last && h_state_out && (flags & B2B_STEP_CONTINUOUS)) {
            for (int k = 0; k < n_chunks; ++k) {
                const int w0 = k * chunk, n = (W - w0 < chunk) ? W - w0 : chunk;
                if (x->f64) launch_gather_state_f64(x->fields, x->c, x->stage_out, w0, n, st);
                else launch_gather_state_f32(x->fields, x->c, x->stage_out, w0, n, st);
                ++x->launches;
                CUDA_TRY(cudaEventRecord(ev_out[k], st), "cudaEventRecord");
                CUDA_TRY(cudaStreamWaitEvent(x->s_out, ev_out[k], 0), "cudaStreamWaitEvent");
                CUDA_TRY(cudaMemcpyAsync((char*)h_state_out + (size_t)w0 * per_world_out, (const char*)x->stage_out + (size_t)w0 * per_world_out,
                                         (size_t)n * per_world_out, cudaMemcpyDeviceToHost, x->s_out), "D2H state");
            }
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(B2B_ERR_CUDA, "step kernels", e);
    }
    if (h_state_out) CUDA_TRY(cudaStreamSynchronize(x->s_out), "cudaStreamSynchronize(copy-out)");
    if (h_force_torque) CUDA_TRY(cudaStreamSynchronize(x->s_in), "cudaStreamSynchronize(copy-in)");
    CUDA_TRY(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return B2B_OK;
}

extern "C" int b2b_query_aabb(b2b_ctx* x, int n, const void* d_world, const void* d_box, int max_hits, void* d_count, void* d_fix,
                              void* stream)
{
    if (!x || !x->bound) return fail(B2B_ERR_UNBOUND, "b2b_query_aabb: no buffers bound");
    if (n < 0 || max_hits < 0 || (n > 0 && (!d_world || !d_box || !d_count || (max_hits > 0 && !d_fix)))) return fail(B2B_ERR_ARG, "b2b_query_aabb: bad argument");
    if (n == 0) return B2B_OK;
    CUDA_TRY(cudaSetDevice(x->caps.device), "cudaSetDevice");
    if (x->f64) launch_query_aabb_f64(x->fields, x->c, n, d_world, d_box, max_hits, d_count, d_fix, (cudaStream_t)stream);
    else launch_query_aabb_f32(x->fields, x->c, n, d_world, d_box, max_hits, d_count, d_fix, (cudaStream_t)stream);
    ++x->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B2B_ERR_CUDA, "query_aabb kernel", e);
    return B2B_OK;
}

extern "C" int b2b_ray_cast(b2b_ctx* x, int n, const void* d_world, const void* d_seg, int mode, int max_hits, void* d_count,
                            void* d_fix, void* d_hit, void* stream)
{
    if (!x || !x->bound) return fail(B2B_ERR_UNBOUND, "b2b_ray_cast: no buffers bound");
    if (n < 0 || max_hits < 1 || (mode != 0 && mode != 1) || (n > 0 && (!d_world || !d_seg || !d_count || !d_fix || !d_hit)))
        return fail(B2B_ERR_ARG, "b2b_ray_cast: bad argument");
    if (n == 0) return B2B_OK;
    CUDA_TRY(cudaSetDevice(x->caps.device), "cudaSetDevice");
    if (x->f64) launch_ray_cast_f64(x->fields, x->c, n, d_world, d_seg, mode, max_hits, d_count, d_fix, d_hit, (cudaStream_t)stream);
    else launch_ray_cast_f32(x->fields, x->c, n, d_world, d_seg, mode, max_hits, d_count, d_fix, d_hit, (cudaStream_t)stream);
    ++x->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B2B_ERR_CUDA, "ray_cast kernel", e);
    return B2B_OK;
}

extern "C" int b2b_test_points(b2b_ctx* x, int n, const void* d_world, const void* d_fixture, const void* d_point, void* d_inside,
                               void* stream)
{
    if (!x || !x->bound) return fail(B2B_ERR_UNBOUND, "b2b_test_points: no buffers bound");
    if (n < 0 || (n > 0 && (!d_world || !d_fixture || !d_point || !d_inside))) return fail(B2B_ERR_ARG, "b2b_test_points: bad argument");
    if (n == 0) return B2B_OK;
    CUDA_TRY(cudaSetDevice(x->caps.device), "cudaSetDevice");
    if (x->f64) launch_test_points_f64(x->fields, x->c, n, d_world, d_fixture, d_point, d_inside, (cudaStream_t)stream);
    else launch_test_points_f32(x->fields, x->c, n, d_world, d_fixture, d_point, d_inside, (cudaStream_t)stream);
    ++x->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B2B_ERR_CUDA, "test_points kernel", e);
    return B2B_OK;
}

extern "C" int b2b_collide_pairs(b2b_ctx* x, int n, const void* d_shape_a, const void* d_shape_b, const void* d_xf_a,
                                 const void* d_xf_b, void* d_mtype, void* d_mf, void* d_pts, void* d_id, void* stream)
{
    if (!x || !x->bound) return fail(B2B_ERR_UNBOUND, "b2b_collide_pairs: no buffers bound");
    if (n <= 0) return B2B_OK;
    CUDA_TRY(cudaSetDevice(x->caps.device), "cudaSetDevice");
    if (x->f64) launch_collide_pairs_f64(x->fields, n, d_shape_a, d_shape_b, d_xf_a, d_xf_b, d_mtype, d_mf, d_pts, d_id, (cudaStream_t)stream);
    else launch_collide_pairs_f32(x->fields, n, d_shape_a, d_shape_b, d_xf_a, d_xf_b, d_mtype, d_mf, d_pts, d_id, (cudaStream_t)stream);
    return check_launch(x, "collide_pairs kernel");
}

extern "C" int b2b_status(b2b_ctx* x, uint32_t* host_out, void* stream)
{
    if (!x || !x->bound || !host_out) return fail(B2B_ERR_UNBOUND, "b2b_status: no buffers bound");
    CUDA_TRY(cudaSetDevice(x->caps.device), "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(x->d_status, 0, sizeof(uint32_t), st), "memset status");
    if (x->f64) launch_status_f64(x->fields, x->c, x->d_status, st);
    else launch_status_f32(x->fields, x->c, x->d_status, st);
    int rc = check_launch(x, "status kernel");
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(host_out, x->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "D2H status");
    CUDA_TRY(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return (*host_out) ? fail(B2B_ERR_OVERFLOW, "a world overflowed a capacity (contacts / pairs / island list)") : B2B_OK;
}

extern "C" int b2b_checksum(b2b_ctx* x, void* d_out, void* stream)
{
    if (!x || !x->bound || !d_out) return fail(B2B_ERR_UNBOUND, "b2b_checksum: no buffers bound");
    CUDA_TRY(cudaSetDevice(x->caps.device), "cudaSetDevice");
    if (x->f64) launch_checksum_f64(x->fields, x->c, (double*)d_out, (cudaStream_t)stream);
    else launch_checksum_f32(x->fields, x->c, (double*)d_out, (cudaStream_t)stream);
    return check_launch(x, "checksum kernel");
}

extern "C" int b2b_info(b2b_ctx* x, int32_t* grid, int32_t* threads, int32_t* smem_bytes, int64_t* launches)
{
    if (!x) return fail(B2B_ERR_ARG, "b2b_info: NULL context");
    if (grid) *grid = x->plan.k_solve.grid;
    if (threads) *threads = x->plan.k_solve.threads;
    if (smem_bytes) *smem_bytes = (int32_t)x->plan.k_solve.smem;
    if (launches) *launches = x->launches;
    return B2B_OK;
}

extern "C" int b2b_profile(b2b_ctx* x, int64_t* host_out32, int reset)
{
    if (!x || !host_out32) return fail(B2B_ERR_ARG, "b2b_profile: NULL argument");
    CUDA_TRY(cudaSetDevice(x->caps.device), "cudaSetDevice");
    CUDA_TRY(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    for (int k = 0; k < PROF_SLOTS; ++k) host_out32[k] = 0;
    long long tmp[PROF_SLOTS];
    for (int pass = 0; pass < 2; ++pass) {
        const int n = pass ? x->plan.solve_slabs : x->plan.aux_slabs;
        const size_t stride = pass ? x->plan.layout.total_sprof : x->plan.layout.total_aux;
        const size_t off = pass ? 0 : x->plan.layout.off_prof;
        unsigned char* base = pass ? x->scratch_sprof : x->scratch;
        for (int g = 0; g < n; ++g) {
            unsigned char* p = base + (size_t)g * stride + off;
            CUDA_TRY(cudaMemcpy(tmp, p, sizeof(tmp), cudaMemcpyDeviceToHost), "D2H profile");
            for (int k = 0; k < PROF_SLOTS; ++k) host_out32[k] += tmp[k];
            if (reset) CUDA_TRY(cudaMemset(p, 0, sizeof(tmp)), "memset profile");
        }
    }
    return B2B_OK;
}

extern "C" int b2b_kernel_ms(b2b_ctx* x, float* out8, int32_t* n_steps)
{
    if (!x || !out8) return fail(B2B_ERR_ARG, "b2b_kernel_ms: NULL argument");
    for (int k = 0; k < 8; ++k) out8[k] = 0.f;
    if (n_steps) *n_steps = 0;
    step_timer* tm = x->timer;
    if (!tm || tm->used == 0) return B2B_OK;
    CUDA_TRY(cudaEventSynchronize(tm->ev[tm->used - 1]), "cudaEventSynchronize");
    for (size_t i = 1; i < tm->used; ++i) {
        if (tm->kind[i] == T_START) continue;                   // the gap between two steps belongs to nobody
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, tm->ev[i - 1], tm->ev[i]), "cudaEventElapsedTime");
        out8[tm->kind[i]] += ms;
        out8[7] += ms;
    }
    if (n_steps) *n_steps = tm->steps;
    tm->used = 0;
    tm->steps = 0;
    return B2B_OK;
}

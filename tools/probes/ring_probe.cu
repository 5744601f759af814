// Development probe (round 2): memory behaviour of the solve kernels' NEW access shape before the kernels are rewritten.
// Every team of `lanes` lanes walks its own world's record stream (n_slots records of Q quads, `sweeps` passes over them),
// consuming 1..4 slots per level-step; the records come through a per-team shared-memory ring of R slots filled by
// cooperative 16-byte cp.async pieces (consecutive lanes copy consecutive quads), completion by wait_group; the last quad
// of every consumed record is written back (the accumulated impulses).  A dependent FMA chain of `chain` links per
// level-step stands in for the constraint arithmetic.  Extra shared memory per team stands in for the bodies' state.
// Reports ms, GB/s of algorithmic traffic and cycles per level-step.
// usage: ring_probe <Q> <R> <lanes> <state_bytes_per_world> <chain> <policy: 0 none, 1 evict_last frac, 2 evict_first> <frac> [worlds_per_sm_waves] [ctas_limit]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__constant__ int c_pattern[16] = {1, 1, 2, 1, 1, 3, 1, 2, 1, 1, 1, 2, 1, 4, 1, 1};   // slots per level-step, mean 1.5

__device__ __forceinline__ void cp16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp16_hint(uint32_t dst, const void* src, unsigned long long pol)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int Q>
__global__ void __launch_bounds__(32) probe(float4* rec, size_t world_quads, int n_slots, int n_worlds, int sweeps, int R, int lanes,
                                            int team_smem_bytes, int chain, int policy, float frac, float* out, unsigned long long* cyc, int flags)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, t_n = lanes, t_id = lane % lanes, team = lane / lanes, teams = 32 / lanes;
    unsigned char* team_base = smem + (size_t)team * team_smem_bytes;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(team_base);
    const float4* ring = reinterpret_cast<const float4*>(team_base);
    unsigned long long pol = 0;
    if (policy == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, %1;" : "=l"(pol) : "f"(frac));
    if (policy == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, %1;" : "=l"(pol) : "f"(frac));
    float acc = 0.f;
    unsigned seg[5] = {0, 0, 0, 0, 0}, tl = 0;
#define STAMP(i, dep) do { asm volatile("" :: "r"(dep) : "memory"); const unsigned n_ = (unsigned)clock(); seg[i] += n_ - tl; tl = n_; } while (0)
    unsigned long long steps_done = 0;
    const long long t0 = clock64();
    const int total = n_slots * sweeps;
    for (int wbase = blockIdx.x * teams; wbase < n_worlds; wbase += gridDim.x * teams) {
        const int w = wbase + team;
        const bool have = w < n_worlds;
        float4* base = rec + (size_t)(have ? w : 0) * world_quads;
        // prime the ring: virtual slots [0, R)
        for (int p = t_id; p < R * Q; p += t_n) {
            const int j = p / Q, q = p - j * Q;
            if (have && j < total) {
                const int src = j % n_slots;
                const uint32_t dst = ring_s + (uint32_t)(((j & (R - 1)) * Q + (Q == 8 ? (q ^ (j & 7)) : q)) * 16);
                if (policy) cp16_hint(dst, base + (size_t)src * Q + q, pol); else cp16(dst, base + (size_t)src * Q + q);
            }
        }
        commit();
        wait_group<0>();
        __syncwarp();
        int vb = 0, in_sweep = 0, step = 0, h1 = -1000, h2 = -1000, h3 = -1000;       // vb of the previous steps
        tl = (unsigned)clock();
        for (;;) {
            const bool active = have && vb < total;
            if (!__any_sync(0xffffffffu, active)) break;
            int cnt = c_pattern[(step + 3 * w) & 15];
            if (cnt > t_n) cnt = t_n;
            if (cnt > n_slots - in_sweep) cnt = n_slots - in_sweep;
            if (!active) cnt = 0;
            // my record
            float4 r[Q];
            const bool mine = t_id < cnt;
            if (mine) {
                const int j = vb + t_id, ri = j & (R - 1);
#pragma unroll
                for (int q = 0; q < Q; ++q) r[q] = ring[ri * Q + (Q == 8 ? (q ^ (ri & 7)) : q)];
            }
            STAMP(0, __float_as_int(mine ? r[0].x : 0.f));
            // refill the consumed ring slots with virtual slots + R
            int cmax = cnt;
            if (flags & 8) cmax = 4;
            else cmax = __reduce_max_sync(0xffffffffu, cnt);
            if (flags & 4) cmax = 0;
            for (int p = t_id; p < cmax * Q; p += t_n) {
                const int j = p / Q, q = p - j * Q;
                const int v = vb + j + R;
                if (j < cnt && v < total) {
                    int src = in_sweep + j + R;
                    if (src >= n_slots) src -= n_slots;
                    const uint32_t dst = ring_s + (uint32_t)(((v & (R - 1)) * Q + (Q == 8 ? (q ^ (v & 7)) : q)) * 16);
                    if (policy) cp16_hint(dst, base + (size_t)src * Q + q, pol); else cp16(dst, base + (size_t)src * Q + q);
                }
            }
            commit();
            STAMP(1, cmax);
            // arithmetic stand-in + write-back of the last quad
            if (mine) {
                float x = r[0].x;
#pragma unroll
                for (int q = 1; q < Q; ++q) x += r[q].y;
                for (int k = 0; k < chain; ++k) x = fmaf(x, 1.0000001f, 1e-9f);
                acc += x;
                float4 wb = r[Q - 1]; wb.x = x;
                if (!(flags & 1)) base[(size_t)(in_sweep + t_id) * Q + (Q - 1)] = wb;
            }
            STAMP(2, __float_as_int(acc));
            // the records of the NEXT step must have landed: they were requested by the step that consumed
            // (their virtual index - R)
            const int ncnt_max = 4;
            const int x = vb + cnt + ncnt_max - 1 - R;               // conservative: newest slot the next step can touch
            if (flags & 2) wait_group<3>();
            else if (x >= h1) wait_group<1>();
            else if (x >= h2) wait_group<2>();
            else if (x >= h3) wait_group<3>();
            else wait_group<4>();
            STAMP(3, x);
            __syncwarp();
            STAMP(4, x);
            h3 = h2; h2 = h1; h1 = vb;
            vb += cnt;
            in_sweep += cnt;
            if (in_sweep >= n_slots) in_sweep -= n_slots;
            ++step;
            if (active) ++steps_done;
        }
        wait_group<0>();
        __syncwarp();
    }
    const long long t1 = clock64();
    if (acc == 123.456f) out[0] = acc;
    if (lane == 0 && steps_done) for (int i = 0; i < 5; ++i) atomicAdd(&cyc[3 + i], (unsigned long long)seg[i]);
    if (t_id == 0 && steps_done) { atomicAdd(&cyc[0], (unsigned long long)(t1 - t0)); atomicAdd(&cyc[1], steps_done); atomicAdd(&cyc[2], 1ull); }
}

int main(int argc, char** argv)
{
    const int Q = argc > 1 ? atoi(argv[1]) : 6, R = argc > 2 ? atoi(argv[2]) : 16, lanes = argc > 3 ? atoi(argv[3]) : 4;
    const int state_bytes = argc > 4 ? atoi(argv[4]) : 2560, chain = argc > 5 ? atoi(argv[5]) : 64;
    const int policy = argc > 6 ? atoi(argv[6]) : 0;
    const float frac = argc > 7 ? (float)atof(argv[7]) : 1.0f;
    const int waves = argc > 8 ? atoi(argv[8]) : 3;
    const int cta_limit = argc > 9 ? atoi(argv[9]) : 0;
    const int n_slots = 401, cap_slots = 640, sweeps = 9, teams = 32 / lanes;
    const int team_smem = ((state_bytes + R * Q * 16 + 127) / 128) * 128;
    const size_t smem = (size_t)team_smem * teams;
    const int flags = argc > 10 ? atoi(argv[10]) : 0;
    void (*kern)(float4*, size_t, int, int, int, int, int, int, int, int, float, float*, unsigned long long*, int) =
        Q == 6 ? probe<6> : Q == 7 ? probe<7> : Q == 8 ? probe<8> : probe<9>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem);
    if (cta_limit > 0 && cta_limit < per_sm) per_sm = cta_limit;
    const int grid = per_sm * 148;
    const int n_worlds = grid * teams * waves;
    const size_t world_quads = (size_t)cap_slots * Q;
    float4* rec; float* out; unsigned long long* cyc;
    if (cudaMalloc(&rec, (size_t)n_worlds * world_quads * 16) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 16); cudaMalloc(&cyc, 64);
    cudaMemset(rec, 0, (size_t)n_worlds * world_quads * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(cyc, 0, 64);
        cudaEventRecord(e0);
        kern<<<grid, 32, smem>>>(rec, world_quads, n_slots, n_worlds, sweeps, R, lanes, team_smem, chain, policy, frac, out, cyc, flags);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) { best = ms; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost); }
    }
    const double bytes = (double)n_worlds * n_slots * sweeps * (Q * 16.0 + 16.0);
    printf("Q=%d (%d B) R=%d lanes=%d state=%d chain=%d policy=%d frac=%.2f: CTAs/SM %d (worlds/SM %d) worlds %d: %.3f ms, %.0f GB/s algorithmic, "
           "%.3f us per world, cycles per level-step %.0f (%s)\n",
           Q, Q * 16, R, lanes, state_bytes, chain, policy, frac, per_sm, per_sm * teams, n_worlds, best, bytes / best * 1e-6,
           best * 1e3 / n_worlds, h[1] ? (double)h[0] / (double)h[1] : 0.0, cudaGetErrorString(cudaGetLastError()));
    {   // lane 0 of every warp: cycles per warp-step in the five segments
        const double wsteps = (double)h[1] / (32 / lanes);
        printf("    flags=%d segments (cycles per warp-step): load %.0f | refill %.0f | chain+store %.0f | wait %.0f | syncwarp %.0f\n", flags,
               h[3] / wsteps, h[4] / wsteps, h[5] / wsteps, h[6] / wsteps, h[7] / wsteps);
    }
    return 0;
}

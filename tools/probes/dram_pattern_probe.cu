// Development probe: HBM throughput of the solve kernels' access pattern vs a plain streaming copy.
//   mode 0: streaming read (coalesced 16 B per lane)
//   mode 1: every lane reads whole 128-byte lines (8 x 16 B) at pseudo-random line indices
//   mode 2: mode 1 plus a 16-byte read and write at a random place of a second array (the impulse rows)
//   mode 3: like mode 2, but each lane walks its own SEQUENTIAL stream of 128-byte records (+16 B rows): 'streams' lanes
// usage: dram_pattern_probe <mode> <threads_per_sm> <GiB> [iters]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long mix(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
__global__ void probe(const float4* __restrict__ a, float4* b, size_t n_lines, size_t n_rows, int mode, int iters, float4* out)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    float4 acc = make_float4(0, 0, 0, 0);
    if (mode == 0) {
        for (int it = 0; it < iters; ++it)
            for (size_t i = tid; i < n_lines * 8; i += nth) { float4 v = a[i]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    } else {
        const size_t per = n_lines / nth;                    // mode 3: this lane's stream
        for (int it = 0; it < iters; ++it) {
            for (size_t k = 0; k < per; ++k) {
                const size_t line = (mode == 3) ? tid * per + k : mix(tid * 0x9e3779b97f4a7c15ULL + k + (size_t)it * 7919) % n_lines;
                const float4* p = a + line * 8;
#pragma unroll
                for (int q = 0; q < 8; ++q) { float4 v = p[q]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
                if (mode >= 2) {
                    const size_t row = (mode == 3) ? tid * per + k : mix(line + 12345) % n_rows;
                    float4 v = b[row]; v.x += acc.x; b[row] = v;
                }
            }
        }
    }
    if (acc.x == 123.456f) out[0] = acc;
}
int main(int argc, char** argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 1, tpsm = argc > 2 ? atoi(argv[2]) : 1024;
    const double gib = argc > 3 ? atof(argv[3]) : 2.0;
    const int iters = argc > 4 ? atoi(argv[4]) : 1;
    const size_t n_lines = (size_t)(gib * 1073741824.0 / 128.0), n_rows = n_lines;
    float4 *a, *b, *out;
    cudaMalloc(&a, n_lines * 128); cudaMalloc(&b, n_rows * 16); cudaMalloc(&out, 16);
    cudaMemset(a, 0, n_lines * 128); cudaMemset(b, 0, n_rows * 16);
    const int block = 128, grid = 148 * tpsm / block;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<<<grid, block>>>(a, b, n_lines, n_rows, mode, 1, out);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<<<grid, block>>>(a, b, n_lines, n_rows, mode, iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const size_t nth = (size_t)grid * block;
    const double lines = (mode == 0) ? (double)n_lines * iters : (double)(n_lines / nth) * nth * iters;
    const double bytes = lines * 128.0 + (mode >= 2 ? lines * 32.0 : 0.0);
    printf("mode %d threads/SM %d buffer %.1f GiB: %.3f ms, %.1f GB/s (%s)\n", mode, tpsm, gib, ms, bytes / ms * 1e-6,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// Latency of the single-lane 8x8 diagonal-tile factorisation (tile_potrf) and of one panel-row substitution, out of shared memory.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_potrf tools/microbench_potrf.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../senquential-convex-programming-for-trajectory-planning_b200/csrc/ipm_core.cuh"

__global__ void k(long long *out, int reps)
{
    __shared__ double T0[64], T1[64], dinv[8];
    __shared__ int fixed;
    const int lane = threadIdx.x;
    long long tp = 0, tt = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int e = lane; e < 64; e += 32) {
            const int r = e >> 3, c = e & 7;
            T0[scp_tphys(r, c)] = (r == c) ? 50.0 + r : 1.0 / (1.0 + r + c) + 0.01 * rep;
            T1[scp_tphys(r, c)] = 0.3 + 0.01 * (r * 8 + c);
        }
        __syncwarp();
        long long t0 = clock64();
        if (lane == 0) tile_potrf(T0, dinv, &fixed);
        __syncwarp();
        long long t1 = clock64();
        if (lane < 8) {                      // one panel row per lane, as in chol_factor (b)
            const int r = lane;
            double *row = T1 + (r << 3);
            const int h0 = ((r >> 1) & 1) << 2, h1 = h0 ^ 4;
            double x[8];
#pragma unroll
            for (int c = 0; c < 4; ++c) { x[c] = row[h0 + c]; x[c + 4] = row[h1 + c]; }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                x[c] *= dinv[c];
#pragma unroll
                for (int c2 = c + 1; c2 < 8; ++c2) x[c2] -= x[c] * T0[scp_tphys(c2, c)];
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) { row[h0 + c] = x[c]; row[h1 + c] = x[c + 4]; }
        }
        __syncwarp();
        long long t2 = clock64();
        tp += t1 - t0; tt += t2 - t1;
    }
    if (lane == 0) { out[0] = tp / reps; out[1] = tt / reps; out[2] = (long long)(T0[scp_tphys(7, 7)] * 1e6) + (long long)(T1[5] * 1e6); }
}

int main()
{
    long long *out;
    cudaMallocManaged(&out, 64);
    k<<<1, 32>>>(out, 4); cudaDeviceSynchronize();
    k<<<1, 32>>>(out, 200); cudaDeviceSynchronize();
    printf("tile_potrf (one lane, shared memory): %lld cycles;  panel-row substitution (8 lanes): %lld cycles;  check %lld\n", out[0], out[1], out[2]);
    return 0;
}

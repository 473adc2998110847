// Cycle counts of the CTA-resident linear algebra of ipm_core.cuh on B200 (calibration / tuning only).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_chol tools/microbench_chol.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../senquential-convex-programming-for-trajectory-planning_b200/csrc/ipm_core.cuh"

#ifndef NT
#define NT 128
#endif
#define N1 81

__global__ void __launch_bounds__(NT, 3) k_bench(long long *out, int reps, int pad_bytes)
{
    extern __shared__ double sh[];
    Cta cta = {NT};
    IpmMem m;
    m.S_far = false; m.n1 = N1; m.n1p = ipm_padded(N1); m.T = m.n1p / 8; m.mc = 0;
    double *p = sh;
    m.S = p; p += (m.T * (m.T + 1) / 2) * 64;
    m.dinv = p; p += m.n1p;
    m.dx = p; p += m.n1p;
    m.tn = p; p += m.n1p;
    __shared__ int fixed;
    long long t_fac = 0, t_inv = 0, t_sol1 = 0, t_sol2 = 0, t_form = 0;
    for (int rep = 0; rep < reps; ++rep) {
        // SPD test matrix: S = I * n + small symmetric part; rhs row; deterministic
        long long t0 = clock64();
        for (int e = threadIdx.x; e < m.n1p * m.n1p; e += NT) {
            const int i = e / m.n1p, j = e % m.n1p;
            if (j <= i) {
                double v = (i == j) ? 100.0 + i : 1.0 / (1.0 + i + j) + ((i * 7 + j * 3 + rep) % 5) * 0.1;
                if (i >= N1 && i < m.n1p - 1) v = (i == j) ? 1.0 : 0.0;
                if (i == m.n1p - 1) v = (j == i) ? 1e300 : (j < N1 ? 1.0 + j * 0.01 : 0.0);
                m.S[scp_sidx(i, j)] = v;
            }
        }
        __syncthreads();
        long long t1 = clock64();
        chol_factor(cta, m, &fixed);
        long long t2 = clock64();
        chol_invert_diag(cta, m, m.dx);
        long long t3 = clock64();
        chol_solve(cta, m, m.dx, true);
        long long t4 = clock64();
        for (int c = threadIdx.x; c < m.n1p; c += NT) m.tn[c] = c < N1 ? 1.0 + c * 0.01 : 0.0;
        __syncthreads();
        long long t5 = clock64();
        chol_solve(cta, m, m.tn, false);
        long long t6 = clock64();
        t_form += t1 - t0; t_fac += t2 - t1; t_inv += t3 - t2; t_sol1 += t4 - t3; t_sol2 += t6 - t5;
    }
    if (threadIdx.x == 0) {
        long long *o = out + (size_t)blockIdx.x * 8;
        o[0] = t_form / reps; o[1] = t_fac / reps; o[2] = t_inv / reps; o[3] = t_sol1 / reps; o[4] = t_sol2 / reps;
        o[5] = (long long)(fabs(m.dx[3] - m.tn[3]) * 1e15);     // both solves answer the same system
    }
}

int main()
{
    long long *out;
    cudaMallocManaged(&out, 8 * 1024 * sizeof(long long));
    const int n1p = ipm_padded(N1), T = n1p / 8;
    size_t smem = ((size_t)(T * (T + 1) / 2) * 64 + 3 * n1p) * 8;
    for (int big = 0; big < 2; ++big) {
        size_t sm = big ? 76384 : smem;        // the solver's footprint: 3 CTAs per SM
        cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        for (int grid : {1, 148, 444}) {
            k_bench<<<grid, NT, sm>>>(out, 2, 0);
            cudaDeviceSynchronize();
            k_bench<<<grid, NT, sm>>>(out, 20, 0);
            cudaError_t e = cudaDeviceSynchronize();
            double a[6] = {0};
            for (int b = 0; b < grid; ++b) for (int k = 0; k < 6; ++k) a[k] += out[b * 8 + k] / (double)grid;
            printf("smem %zu grid %3d threads %d: fill %6.0f  factor %6.0f  invert-diag %5.0f  solve(back) %6.0f  solve(fwd+back) %6.0f   |dx-tn|*1e15 %.0f  (%s)\n",
                   sm, grid, NT, a[0], a[1], a[2], a[3], a[4], a[5], cudaGetErrorString(e));
        }
    }
    return 0;
}

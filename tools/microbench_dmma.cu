// DMMA (mma.sync.m8n8k4.f64) check on B200: fragment layout, dependent latency, throughput.  Calibration only.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__global__ void k_check(const double *A, const double *B, double *C)   // C[8x8] = A[8x4] B[4x8]
{
    int l = threadIdx.x;
    double d0, d1;
    dmma(d0, d1, A[(l >> 2) * 4 + (l & 3)], B[(l & 3) * 8 + (l >> 2)], 0.0, 0.0);
    C[(l >> 2) * 8 + (l & 3) * 2] = d0;
    C[(l >> 2) * 8 + (l & 3) * 2 + 1] = d1;
}
__global__ void k_lat(double *out, long long *cyc)
{
    double a = 1.0 + threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3, c0 = 0, c1 = 0;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 512; ++i) dmma(c0, c1, a, b, c0, c1);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = c0 + c1;
}
__global__ void k_tput(double *out, int iters)
{
    double a = 1.0 + threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
    double c[8][2] = {};
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma(c[j][0], c[j][1], a, b, c[j][0], c[j][1]);
    double s = 0;
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double hA[32], hB[32], hC[64], *A, *B, *C, *out; long long *cyc;
    for (int i = 0; i < 32; ++i) { hA[i] = i * 0.5 + 1; hB[i] = 3 - i * 0.25; }
    cudaMalloc(&A, sizeof hA); cudaMalloc(&B, sizeof hB); cudaMalloc(&C, sizeof hC); cudaMalloc(&out, 1 << 24);
    cudaMallocManaged(&cyc, 64);
    cudaMemcpy(A, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, sizeof hB, cudaMemcpyHostToDevice);
    k_check<<<1, 32>>>(A, B, C); cudaMemcpy(hC, C, sizeof hC, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) { double r = 0; for (int k = 0; k < 4; ++k) r += hA[i * 4 + k] * hB[k * 8 + j]; err = fmax(err, fabs(r - hC[i * 8 + j])); }
    printf("fragment layout check: max err %.3e (%s)\n", err, err < 1e-12 ? "OK" : "WRONG");
    k_lat<<<1, 32>>>(out, cyc); cudaDeviceSynchronize();
    printf("DMMA m8n8k4 dependent latency: %.1f cycles\n", cyc[0] / 512.0);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for (int nt : {128, 256}) {
        int iters = 20000, grid = p.multiProcessorCount * (1024 / nt);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_tput<<<grid, nt>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_tput<<<grid, nt>>>(out, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 256 * 8.0 * iters * (double)grid * (nt / 32);
        printf("DMMA throughput, %d threads/CTA x %d CTAs: %.2f TFLOP/s\n", nt, grid, fl / ms / 1e9);
    }
    return 0;
}

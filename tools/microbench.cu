// Latency / throughput microbenchmarks for the FP64 path on B200 (calibration for DESIGN.md; not product code).
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k_lat(double *out, long long *cyc, double x0)
{
    __shared__ double sm[256];
    sm[threadIdx.x] = x0 + threadIdx.x;
    __syncthreads();
    double a = x0, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    // dependent DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fma(a, b, c);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // dependent DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + c;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // dependent rsqrt chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 64; ++i) a = rsqrt(a + 2.0);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0) * (N / 64);
    // dependent division chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 64; ++i) a = 1.0 / (a + 2.0);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0) * (N / 64);
    // dependent shared-memory pointer chase (LDS latency)
    int idx = threadIdx.x & 31;
    __shared__ int nxt[32];
    nxt[idx] = (idx + 1) & 31;
    __syncwarp();
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = nxt[idx];
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // LDS.64 -> DFMA dependent (load, fma, address from result)
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { a = fma(sm[(idx + i) & 255], b, a); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 64-bit shuffle chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1) + c;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // syncwarp cost with a shared store/load round trip
    t0 = clock64();
    for (int i = 0; i < N; ++i) { sm[threadIdx.x] = a; __syncwarp(); a = sm[threadIdx.x ^ 1] + c; __syncwarp(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // __syncthreads round trip
    t0 = clock64();
    for (int i = 0; i < N; ++i) { sm[threadIdx.x] = a; __syncthreads(); a = sm[threadIdx.x ^ 1] + c; __syncthreads(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[8] = t1 - t0;
    out[threadIdx.x] = a + idx;
}
// FP64 throughput: independent DFMA chains, all SMs
__global__ void k_tput(double *out, int iters)
{
    double a[8];
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3 + j;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fma(a[j], b, c);
    double s = 0;
    for (int j = 0; j < 8; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 1 << 24); cudaMallocManaged(&cyc, 16 * sizeof(long long));
    const char *names[] = {"DFMA dep", "DADD dep", "rsqrt(double) dep", "1/x (double) dep", "LDS pointer chase", "LDS.64->DFMA", "SHFL.64 + DADD", "STS/syncwarp/LDS x2 roundtrip", "STS/syncthreads/LDS x2 roundtrip"};
    for (int nt : {32, 256}) {
        k_lat<<<1, nt>>>(out, cyc, 1.5);
        cudaDeviceSynchronize();
        printf("threads per CTA = %d\n", nt);
        for (int i = 0; i < 9; ++i) printf("  %-36s %.1f cycles/op\n", names[i], (double)cyc[i] / N);
    }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for (int nt : {128, 256, 512}) {
        int iters = 20000, grid = p.multiProcessorCount * (2048 / nt);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_tput<<<grid, nt>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_tput<<<grid, nt>>>(out, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * (double)grid * nt;
        printf("FP64 DFMA throughput, %d threads/CTA, %d CTAs: %.2f TFLOP/s (%.3f ms)\n", nt, grid, fl / ms / 1e9, ms);
    }
    printf("clock %d kHz, SMs %d\n", p.clockRate, p.multiProcessorCount);
    return 0;
}

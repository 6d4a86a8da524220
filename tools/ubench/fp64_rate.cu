// micro-benchmark: FP64 FMA and warp-shuffle issue rates per SM on B200 as a function of resident warps and ILP
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_dfma(double *out, long long *cyc, int iters, double c, double s) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], c, s);
    }
    long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
__global__ void k_shfl(double *out, long long *cyc, int iters) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x + i;
    int src = (threadIdx.x + 1) & 31;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = __shfl_sync(0xffffffffu, a[i], src);
    }
    long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// the rotation pattern of the checkerboard step: t = c*a; a' = fma(s, b, t)  (DMUL + dependent DFMA), ILP independent pairs
template <int ILP>
__global__ void k_rot(double *out, long long *cyc, int iters, double c, double s) {
    double a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x + i; b[i] = 1.0 / (1 + threadIdx.x + i); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            double na = fma(s, b[i], c * a[i]), nb = fma(s, a[i], c * b[i]);
            a[i] = na; b[i] = nb;
        }
    }
    long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMallocManaged(&cyc, 8);
    const int iters = 2000;
    for (int threads : {32, 64, 128, 256, 512, 1024}) {
        for (int rep = 0; rep < 2; rep++) { k_dfma<8><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9); cudaDeviceSynchronize(); }
        double d8 = (double)cyc[0];
        for (int rep = 0; rep < 2; rep++) { k_dfma<1><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9); cudaDeviceSynchronize(); }
        double d1 = (double)cyc[0];
        for (int rep = 0; rep < 2; rep++) { k_dfma<32><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9); cudaDeviceSynchronize(); }
        double d32 = (double)cyc[0];
        for (int rep = 0; rep < 2; rep++) { k_shfl<8><<<148, threads>>>(out, cyc, iters); cudaDeviceSynchronize(); }
        double s8 = (double)cyc[0];
        for (int rep = 0; rep < 2; rep++) { k_rot<16><<<148, threads>>>(out, cyc, iters, 1.0003, 0.025); cudaDeviceSynchronize(); }
        double r16 = (double)cyc[0];
        printf("threads/SM %4d: DFMA lanes/clk/SM  ILP1 %.1f (latency %.1f clk)  ILP8 %.1f  ILP32 %.1f | SHFL32 lanes/clk/SM (64-bit = 2 SHFL) %.1f | rot(ILP16 pairs) FP64 lanes/clk/SM %.1f\n",
               threads, threads * iters / d1, d1 / iters, threads * 8.0 * iters / d8, threads * 32.0 * iters / d32,
               threads * 8.0 * 2 * iters / s8, threads * 16.0 * 4 * iters / r16);
    }
    return 0;
}

// micro-benchmark: cost of one "colour step" (LDS.128 x2 -> rotation -> STS.128 x2 -> barrier) on B200
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void rot(double2 &a, double2 &b, double c, double s) {
    double2 na = make_double2(fma(s, b.x, c * a.x), fma(s, b.y, c * a.y));
    double2 nb = make_double2(fma(s, a.x, c * b.x), fma(s, a.y, c * b.y));
    a = na; b = nb;
}
template <int MODE>
__global__ void k(double2 *out, long long *cyc, int iters, double c, double s) {
    extern __shared__ double2 Y[];
    int t = threadIdx.x, n = blockDim.x;
    Y[t] = make_double2(t, 1); Y[t + n] = make_double2(1, t);
    __syncthreads();
    long long t0 = clock64();
    double2 a = Y[t], b = Y[t + n];
    for (int i = 0; i < iters; i++) {
        int oi = (t + i) % n, oj = n + (t + 3 * i) % n;      // different pairing every step
        if (MODE != 3) { a = Y[oi]; b = Y[oj]; }
        if (MODE == 0 || MODE == 3) rot(a, b, c, s);
        if (MODE == 4) { rot(a, b, c, s); a.x *= 1.0000001; b.x *= 1.0000001; rot(a, b, c, s); }
        if (MODE != 3) { Y[oi] = a; Y[oj] = b; }
        if (MODE != 2) __syncthreads();
    }
    long long t1 = clock64();
    if (t == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * n + t] = make_double2(a.x + b.x, a.y + b.y);
}
int main() {
    double2 *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double2)); cudaMallocManaged(&cyc, 8);
    const char *names[] = {"LDS+rot+STS+sync", "LDS+STS+sync (no math)", "LDS+rot... no sync (MODE2: no rot either)", "rot only (+sync)", "LDS+2rot+STS+sync (mid)"};
    for (int threads : {128, 256, 512, 1024}) {
        for (int mode = 0; mode < 5; mode++) {
            int iters = 2000;
            size_t sm = 2 * threads * sizeof(double2);
            for (int rep = 0; rep < 2; rep++) {
                switch (mode) {
                    case 0: k<0><<<112, threads, sm>>>(out, cyc, iters, 1.0003, 0.025); break;
                    case 1: k<1><<<112, threads, sm>>>(out, cyc, iters, 1.0003, 0.025); break;
                    case 2: k<2><<<112, threads, sm>>>(out, cyc, iters, 1.0003, 0.025); break;
                    case 3: k<3><<<112, threads, sm>>>(out, cyc, iters, 1.0003, 0.025); break;
                    case 4: k<4><<<112, threads, sm>>>(out, cyc, iters, 1.0003, 0.025); break;
                }
                cudaDeviceSynchronize();
            }
            printf("threads %4d  %-45s %7.1f cycles/step\n", threads, names[mode], (double)cyc[0] / iters);
        }
    }
    return 0;
}

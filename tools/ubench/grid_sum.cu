// micro-benchmark: cost of one grid-wide deterministic sum of 4 doubles over 134 resident CTAs on B200, nothing else in the
// loop.  Variants: (A) reducer CTA polls all slots with strong loads, (B) with weak .cg loads, (C) all-to-all polling,
// (D) atomic counter barrier + everybody reads the partials (the cooperative-groups style).
#include <cstdio>
#include <cuda_runtime.h>
struct Slot { double v; unsigned long long e; };
struct Slot4 { Slot q[4]; };
__device__ __forceinline__ void st_slot(Slot *p, double v, unsigned long long e) {
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(e) : "memory");
}
template <int WEAK>
__device__ __forceinline__ void ld_slot(const Slot *p, long long &v, unsigned long long &e) {
    if (WEAK) asm volatile("ld.global.cg.v2.b64 {%0, %1}, [%2];" : "=l"(v), "=l"(e) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(v), "=l"(e) : "l"(p) : "memory");
}
template <int MODE>   // 0: reducer strong, 1: reducer weak, 2: all-to-all strong, 3: all-to-all weak
__global__ void k(char *slots, unsigned stride, int iters, long long *cyc, double *out) {
    const unsigned nblk = gridDim.x, bid = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double sh[4];
    double acc = 1.0 + bid;
    long long t0 = clock64();
    for (int it = 1; it <= iters; it++) {
        const unsigned long long epoch = it;
        __syncthreads();
        if (warp == 0) {
            Slot *total = ((Slot4 *)(slots + (size_t)(nblk + (bid & 7)) * stride))->q;
            if (lane < 4 && (bid != 0 || MODE >= 2)) st_slot(((Slot4 *)(slots + (size_t)bid * stride))->q + lane, acc + lane, epoch);
            if (bid == 0 || MODE >= 2) {
                double s[4] = {0, 0, 0, 0};
                if (MODE < 2 && lane == 0) { s[0] = acc; s[1] = acc + 1; s[2] = acc + 2; s[3] = acc + 3; }
                for (unsigned base = (MODE < 2 ? 1 : 0); base < nblk; base += 128) {
                    long long val[4][4]; unsigned long long ep[4][4];
                    while (true) {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const unsigned q = base + lane + 32 * u;
                            const Slot *sl = ((const Slot4 *)(slots + (size_t)q * stride))->q;
#pragma unroll
                            for (int c = 0; c < 4; c++) { ep[u][c] = epoch; val[u][c] = 0; if (q < nblk) ld_slot<MODE & 1>(sl + c, val[u][c], ep[u][c]); }
                        }
                        bool ready = true;
#pragma unroll
                        for (int u = 0; u < 4; u++)
#pragma unroll
                            for (int c = 0; c < 4; c++) ready = ready && ep[u][c] >= epoch;
                        if (ready) break;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++)
#pragma unroll
                        for (int c = 0; c < 4; c++) s[c] += __longlong_as_double(val[u][c]);
                }
                double tot = 0;
#pragma unroll
                for (int c = 0; c < 4; c++) { double t = s[c]; for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o); if (lane == c) tot = t; }
                if (lane < 4) sh[lane] = tot;
                if (MODE < 2) {
                    const double tc = __shfl_sync(0xffffffffu, tot, lane & 3);
                    st_slot(((Slot4 *)(slots + (size_t)(nblk + (lane >> 2)) * stride))->q + (lane & 3), tc, epoch);
                }
            } else if (lane < 4) {
                long long v; unsigned long long e;
                do { ld_slot<0>(total + lane, v, e); } while (e < epoch);
                sh[lane] = __longlong_as_double(v);
            }
        }
        __syncthreads();
        acc = sh[0] * 1e-9 + bid;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[bid] = acc; if (bid == 1) cyc[0] = t1 - t0; }
}
__global__ void k_atomic(unsigned *counter, double *part, int iters, long long *cyc, double *out) {
    const unsigned nblk = gridDim.x, bid = blockIdx.x;
    __shared__ double sh[1];
    double acc = 1.0 + bid;
    long long t0 = clock64();
    for (int it = 1; it <= iters; it++) {
        __syncthreads();
        if (threadIdx.x == 0) {
            part[bid] = acc;
            __threadfence();
            atomicAdd(counter, 1u);
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < (unsigned)it * nblk);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            double t = 0;
            for (unsigned q = threadIdx.x; q < nblk; q += 32) t += __ldcg(part + q);
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (threadIdx.x == 0) sh[0] = t;
        }
        __syncthreads();
        acc = sh[0] * 1e-9 + bid;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[bid] = acc; if (bid == 1) cyc[0] = t1 - t0; }
}
int main() {
    char *slots; long long *cyc; double *out, *part; unsigned *counter;
    const unsigned stride = 1024; const int nblk = 134, iters = 2000;
    cudaMalloc(&slots, (nblk + 8) * stride); cudaMallocManaged(&cyc, 64); cudaMalloc(&out, nblk * 8); cudaMalloc(&part, nblk * 8); cudaMalloc(&counter, 4);
    const char *names[] = {"reducer CTA, strong loads", "reducer CTA, weak .cg loads", "all-to-all, strong loads", "all-to-all, weak .cg loads"};
    for (int mode = 0; mode < 4; mode += 2) {
        for (int rep = 0; rep < 2; rep++) {
            cudaMemset(slots, 0, (nblk + 8) * stride);
            void *args[] = {&slots, (void *)&stride, (void *)&iters, &cyc, &out};
            const void *f = mode == 0 ? (const void *)k<0> : mode == 1 ? (const void *)k<1> : mode == 2 ? (const void *)k<2> : (const void *)k<3>;
            cudaError_t e = cudaLaunchCooperativeKernel(f, dim3(nblk), dim3(256), args, 0, 0);
            if (e != cudaSuccess) printf("launch failed: %s\n", cudaGetErrorString(e));
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        }
        printf("%-32s %7.0f cycles per grid sum (4 doubles, %d CTAs)\n", names[mode], (double)cyc[0] / iters, nblk); fflush(stdout);
    }
    for (int rep = 0; rep < 2; rep++) {
        cudaMemset(counter, 0, 4);
        void *args[] = {&counter, &part, (void *)&iters, &cyc, &out};
        cudaLaunchCooperativeKernel((const void *)k_atomic, dim3(nblk), dim3(256), args, 0, 0);
        cudaDeviceSynchronize();
    }
    printf("%-32s %7.0f cycles per grid sum (1 double, %d CTAs)\n", "atomic counter + partial reads", (double)cyc[0] / iters, nblk);
    return 0;
}

// micro-benchmark: CTA-to-CTA signalling latency through L2 on B200 (relaxed gpu-scope store -> polling load), the building
// block of the grid-wide sums and halo handshakes of the resident CG kernel.  Also: cost of __threadfence() after N stores.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void pingpong(unsigned long long *flags, long long *cyc, int iters, int other_block) {
    // block 0 and block `other_block` bounce a counter; all other blocks exit
    if (blockIdx.x != 0 && blockIdx.x != other_block) return;
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x == 0 ? 0 : 1;
    unsigned long long *mine = flags + 32 * me, *theirs = flags + 32 * (1 - me);
    long long t0 = clock64();
    for (int i = 1; i <= iters; i++) {
        if (me == 0) {
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mine), "l"((unsigned long long)i) : "memory");
            unsigned long long e;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(e) : "l"(theirs) : "memory"); } while (e < (unsigned long long)i);
        } else {
            unsigned long long e;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(e) : "l"(theirs) : "memory"); } while (e < (unsigned long long)i);
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mine), "l"((unsigned long long)i) : "memory");
        }
    }
    long long t1 = clock64();
    if (me == 0) cyc[0] = t1 - t0;
}
__global__ void fence_cost(double2 *buf, long long *cyc, int iters, int nstores) {
    long long t0 = clock64(), tf = 0;
    for (int i = 0; i < iters; i++) {
        for (int k = 0; k < nstores; k++) buf[(size_t)blockIdx.x * 4096 + k * blockDim.x + threadIdx.x] = make_double2(i, k);
        long long a = clock64();
        __threadfence();
        tf += clock64() - a;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = tf; }
}
__global__ void load_latency(const unsigned long long *p, long long *cyc, int iters) {
    if (threadIdx.x != 0) return;
    unsigned long long idx = blockIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p + idx) : "memory"); idx = v; }
    long long t1 = clock64();
    if (blockIdx.x == 0) cyc[0] = (t1 - t0) + (idx == 12345);
}
int main() {
    unsigned long long *flags; long long *cyc; double2 *buf;
    cudaMalloc(&flags, 1 << 20); cudaMallocManaged(&cyc, 64); cudaMalloc(&buf, 148 * 4096 * sizeof(double2));
    const int iters = 2000;
    for (int other : {1, 2, 8, 37, 74, 100, 147}) {
        cudaMemset(flags, 0, 1 << 20);
        pingpong<<<148, 32>>>(flags, cyc, iters, other); cudaDeviceSynchronize();
        cudaMemset(flags, 0, 1 << 20);
        pingpong<<<148, 32>>>(flags, cyc, iters, other); cudaDeviceSynchronize();
        printf("ping-pong block 0 <-> block %3d: %.0f cycles per round trip (2 x (store -> visible -> polled))\n", other, (double)cyc[0] / iters);
    }
    for (int ns : {0, 1, 4, 16}) {
        fence_cost<<<134, 256>>>(buf, cyc, 200, ns); cudaDeviceSynchronize();
        printf("__threadfence after %2d 16-byte stores per thread (134 CTAs x 256 thr): %.0f cycles in the fence, %.0f per iteration\n", ns, (double)cyc[1] / 200, (double)cyc[0] / 200);
    }
    cudaMemset(flags, 0, 1 << 20);
    load_latency<<<1, 32>>>(flags, cyc, 2000); cudaDeviceSynchronize();
    printf("dependent ld.relaxed.gpu (L2 hit) latency: %.0f cycles\n", (double)cyc[0] / 2000);
    return 0;
}

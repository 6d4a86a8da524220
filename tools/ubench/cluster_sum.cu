// micro-benchmark for the next step of the resident CG: what do thread-block clusters buy?
//   mode 0  flat: every CTA publishes a tagged 32-byte slot in global memory and polls all slots (what k_cg_v3_resident1 does)
//   mode 1  hierarchical: the CTAs of a cluster combine their 4 partials in the leader's shared memory (DSMEM stores +
//           cluster barrier), the leader publishes ONE slot per cluster, everybody polls nblk / CS slots
//   mode 2  halo: every CTA writes an 8 KB boundary slice into its neighbour's shared memory (DSMEM) + cluster barrier
//   mode 3  halo through global memory: 8 KB store, gpu-scope fence, flag, neighbour polls the flag and reads the slice
// One CTA per SM (large dynamic shared memory), 256 threads, clock64 per iteration.  Build: nvcc -arch=sm_100a -O3 -o cluster_sum cluster_sum.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ double tagd(double x, long long tag) { return __longlong_as_double((__double_as_longlong(x) & ~3LL) | tag); }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(char *slots, double *halo, unsigned *flags, int iters, long long *cyc, double *out) {
    extern __shared__ __align__(16) double sm[];          // [0, 64): cluster partials (leader); [64, 64 + 1024): halo inbox
    cg::cluster_group cl = cg::this_cluster();
    const unsigned cs = cl.num_blocks(), cr = cl.block_rank();
    const unsigned nblk = gridDim.x, bid = blockIdx.x, ncl = nblk / cs, cid = bid / cs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double tot[4];
    double acc = 1.0 + bid, sink = 0.0;
    cl.sync();
    const long long t0 = clock64();
    for (int it = 1; it <= iters; it++) {
        const long long tag = it & 3;
        if (MODE == 0 || MODE == 1) {
            unsigned npoll = nblk, me = bid;
            bool publish = true;
            double p[4] = {acc, acc + 1, acc + 2, acc + 3};
            if (MODE == 1) {
                double *lead = cl.map_shared_rank(sm, 0);
                if (threadIdx.x < 4) lead[4 * cr + threadIdx.x] = p[threadIdx.x];       // DSMEM store into the leader's shared memory
                cl.sync();
                publish = (cr == 0);
                if (publish && threadIdx.x < 4) {
                    double s = 0;
                    for (unsigned q = 0; q < cs; q++) s += sm[4 * q + threadIdx.x];       // fixed order
                    tot[threadIdx.x] = s;
                }
                __syncthreads();
                if (publish) for (int c = 0; c < 4; c++) p[c] = tot[c];
                npoll = ncl; me = cid;
            }
            if (warp == 0) {
                char *base = slots + (size_t)(it & 1) * 65536;
                if (publish && lane < 2) {
                    asm volatile("fence.acq_rel.gpu;" ::: "memory");
                    asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(base + (size_t)me * 64 + 16 * lane), "d"(tagd(p[2 * lane], tag)),
                                 "d"(tagd(p[2 * lane + 1], tag)) : "memory");
                }
                double s[4] = {0, 0, 0, 0};
                for (unsigned b0 = 0; b0 < npoll; b0 += 32) {
                    const unsigned q = b0 + lane;
                    long long v[4] = {tag, tag, tag, tag};
                    while (true) {
                        if (q < npoll) {
                            asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(v[0]), "=l"(v[1]) : "l"(base + (size_t)q * 64) : "memory");
                            asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(v[2]), "=l"(v[3]) : "l"(base + (size_t)q * 64 + 16) : "memory");
                        }
                        const bool ok = ((v[0] & 3) == tag) && ((v[1] & 3) == tag) && ((v[2] & 3) == tag) && ((v[3] & 3) == tag);
                        if (__all_sync(0xffffffffu, ok)) break;
                        if (clock64() - t0 > 3000000000LL) break;       // never hang the box
                    }
                    if (q < npoll) for (int c = 0; c < 4; c++) s[c] += __longlong_as_double(v[c]);
                }
                for (int c = 0; c < 4; c++) { for (int o = 16; o; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o); }
                if (lane == 0) for (int c = 0; c < 4; c++) tot[c] = s[c];
            }
            __syncthreads();
            acc = 1.0 + bid + 1e-9 * tot[0];
            if (MODE == 1) cl.sync();                       // the leader's partial array may be overwritten only after everybody is through
        } else if (MODE == 2) {
            double *inbox = cl.map_shared_rank(sm + 64, (cr + 1) % cs);
            reinterpret_cast<double2 *>(inbox)[threadIdx.x] = make_double2(acc + threadIdx.x, acc);
            reinterpret_cast<double2 *>(inbox)[threadIdx.x + 256] = make_double2(acc, acc - threadIdx.x);
            cl.sync();
            sink += sm[64 + threadIdx.x] + sm[64 + 512 + threadIdx.x];
            cl.sync();
            acc += 1e-9 * sink;
        } else {
            const unsigned nb = (bid + 1) % nblk, pv = (bid + nblk - 1) % nblk;
            double2 *dst = reinterpret_cast<double2 *>(halo + ((size_t)(it & 1) * nblk + nb) * 1024);
            dst[threadIdx.x] = make_double2(acc + threadIdx.x, acc);
            dst[threadIdx.x + 256] = make_double2(acc, acc - threadIdx.x);
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flags + nb), "r"((unsigned)it) : "memory");
                unsigned f;
                do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(flags + bid) : "memory"); } while (f < (unsigned)it && clock64() - t0 < 3000000000LL);
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
            __syncthreads();
            const double2 *src = reinterpret_cast<const double2 *>(halo + ((size_t)(it & 1) * nblk + bid) * 1024);
            double2 a, b;
            asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(a.x), "=d"(a.y) : "l"(src + threadIdx.x) : "memory");
            asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(b.x), "=d"(b.y) : "l"(src + threadIdx.x + 256) : "memory");
            sink += a.x + b.y;
            acc += 1e-9 * sink;
            (void)pv;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[bid] = (t1 - t0) / iters; out[bid] = acc + sink; }
}

template <int MODE>
static void run(int cs, int iters) {
    const size_t smem = 150 * 1024;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cs > 8) cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(cs);
    int ncl = 0;
    cudaOccupancyMaxActiveClusters(&ncl, k<MODE>, &cfg);
    if (ncl <= 0) { printf("mode %d cluster %d: not launchable\n", MODE, cs); return; }
    const int nblk = ncl * cs;
    cfg.gridDim = dim3(nblk);
    char *slots; double *halo, *out; unsigned *flags; long long *cyc;
    cudaMalloc(&slots, 2 * 65536); cudaMemset(slots, 0xff, 2 * 65536);
    cudaMalloc(&halo, (size_t)2 * nblk * 1024 * sizeof(double));
    cudaMalloc(&flags, nblk * sizeof(unsigned)); cudaMemset(flags, 0, nblk * sizeof(unsigned));
    cudaMalloc(&cyc, nblk * sizeof(long long)); cudaMalloc(&out, nblk * sizeof(double));
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<MODE>, slots, halo, flags, iters, cyc, out);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d cluster %d: %s\n", MODE, cs, cudaGetErrorString(e)); return; }
    long long *h = (long long *)malloc(nblk * sizeof(long long));
    cudaMemcpy(h, cyc, nblk * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0, sum = 0;
    for (int i = 0; i < nblk; i++) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
    static const char *names[] = {"flat slots", "cluster-combined slots", "8 KB halo via DSMEM", "8 KB halo via global + flag"};
    printf("mode %d (%s) cluster %2d: %3d CTAs in %2d clusters, %lld cycles per iteration (max over CTAs %lld)\n", MODE, names[MODE], cs, nblk, ncl,
           sum / nblk, mx);
    free(h); cudaFree(slots); cudaFree(halo); cudaFree(flags); cudaFree(cyc); cudaFree(out);
}

int main() {
    const int iters = 2000;
    for (int cs : {1, 2, 4, 8, 16}) {
        run<0>(cs, iters);
        if (cs > 1) { run<1>(cs, iters); run<2>(cs, iters); }
        run<3>(cs, iters);
    }
    return 0;
}

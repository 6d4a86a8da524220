// common.cuh -- error handling, device buffers and small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstring>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

typedef int64_t i64;

struct SqError : public std::runtime_error {
    explicit SqError(const std::string &m) : std::runtime_error(m) {}
};

// Thrown by the NaN / non-finite checks (CG residual, Lanczos recurrence, fermionic action).  The reference signals these by
// throwing and its callers turn them into a rejected update (src/EFAPFFHMCUpdater.jl:168-187,215-231); hmc_update and the
// global moves catch THIS class only -- bad arguments, CUDA / NCCL errors and watchdog time-outs propagate.  ABI status 3.
struct SqNumericalInstability : public SqError {
    explicit SqNumericalInstability(const std::string &m) : SqError(m) {}
};

void sq_set_last_error(const std::string &m);

#define SQ_CUDA(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            throw SqError(std::string("CUDA error '") + cudaGetErrorString(_e) + "' at " + __FILE__ + \
                          ":" + std::to_string(__LINE__) + " in " #expr);                              \
    } while (0)

#define SQ_REQUIRE(cond, msg)                                                       \
    do {                                                                            \
        if (!(cond)) throw SqError(std::string("invalid argument: ") + (msg));     \
    } while (0)

#define SQ_LAUNCH_CHECK() SQ_CUDA(cudaGetLastError())

// RAII device buffer
// Programmatic dependent launch (PDL) for the launch-bound preconditioned CG iteration: a kernel launched with the attribute may become
// resident while its predecessor in the stream still runs; it must not touch global memory before sq_pdl_prologue(), which (i) lets ITS
// successor start launching and (ii) waits until the predecessor grid has completed and its writes are visible.  Without the attribute
// both instructions are no-ops.  g_sq_pdl is set by the solver loop around the launches that take part.
extern thread_local int g_sq_pdl;
__device__ __forceinline__ void sq_pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KA, typename... A>
inline cudaError_t sq_launch(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at;
    memset(&at, 0, sizeof(at));
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at; cfg.numAttrs = g_sq_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KA(args)...);
}

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count, bool zero = true) {
        release();
        n = count;
        size_t bytes = (count ? count : 1) * sizeof(T);
        SQ_CUDA(cudaMalloc((void **)&p, bytes));
        if (zero) {
            // cudaMemset runs on the legacy default stream, which the library's non-blocking streams do not
            // order against: finish it before anyone can enqueue work on the new buffer.
            SQ_CUDA(cudaMemset(p, 0, bytes));
            SQ_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
        }
    }
    void upload(const T *h, size_t count, cudaStream_t s) {
        if (count > n) throw SqError("DevBuf::upload overflow");
        if (count) SQ_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T *h, size_t count, cudaStream_t s) const {
        if (count > n) throw SqError("DevBuf::download overflow");
        if (count) SQ_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
    void from_vector(const std::vector<T> &v, cudaStream_t s) {
        alloc(v.size(), false);
        upload(v.data(), v.size(), s);
        SQ_CUDA(cudaStreamSynchronize(s));   // the vector may be a temporary
    }
};

// ---- device helpers ------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {   // conj(a) * b
    return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cscale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
    double d = b.x * b.x + b.y * b.y;
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block reduction of up to 4 doubles per thread.  `red` needs 4*32 doubles of shared
// memory.  The result is valid in thread 0.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double *red) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) red[k * 32 + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            double t = lane < nw ? red[k * 32 + lane] : 0.0;
            v[k] = warp_sum(t);
        }
    }
}

// Sum `n` per-block partials (stride-1 array) in a fixed order with one warp; all lanes get the result.
__device__ __forceinline__ double warp_sum_partials(const double *part, int n) {
    double t = 0.0;
    for (int k = threadIdx.x & 31; k < n; k += 32) t += part[k];
    return warp_sum(t);
}

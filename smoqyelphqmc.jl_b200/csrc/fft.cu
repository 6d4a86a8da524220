// fft.cu -- K3: batched mixed-radix Stockham FFT along the imaginary-time axis, hand written.
//
// Replaces FourierTransformer (src/FourierTransformer.jl:2-77: FFTW plan_fft!/plan_ifft! along dim 1 plus
// the antiperiodic twist theta_l = exp(-i pi l / Ltau) and the 1/sqrt(Ltau) normalisation) and, through the
// [l][i] device layout, the two transpose! passes of the preconditioner (KPMPreconditioner.jl:378,403):
// the transform of an [l][i] array is directly the frequency-major [n][i] array the KPM stage wants.
//
// One CTA transforms SB neighbouring columns (sites) for all Ltau: the Ltau x SB tile is staged in shared
// memory once (coalesced SB*16-byte segments per time slice), all radix passes run in shared memory, and
// the tile is written back once: 32 bytes of HBM traffic per element per transform.
// Radices 2, 3, 4, 5, 7 have unrolled butterflies; any other prime factor uses a generic O(r^2) butterfly
// (Ltau = beta/dtau is 20*beta for the shipped dtau = 0.05, i.e. 2^a 5^b).
#include "sq_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>

void fft_radices(i64 n, std::vector<int> &rad) {
    rad.clear();
    while (n % 4 == 0) { rad.push_back(4); n /= 4; }
    while (n % 2 == 0) { rad.push_back(2); n /= 2; }
    for (int p : {3, 5, 7}) while (n % p == 0) { rad.push_back(p); n /= p; }
    for (i64 p = 11; n > 1; p += 2) {
        while (n % p == 0) { rad.push_back((int)p); n /= p; }
        if (p * p > n && n > 1) { rad.push_back((int)n); n = 1; }
    }
}

void fft_make_twiddles(i64 n, std::vector<double2> &tw) {
    tw.resize(n);
    const long double PI = 3.141592653589793238462643383279502884L;
    for (i64 k = 0; k < n; k++) {
        long double a = -2.0L * PI * (long double)k / (long double)n;
        tw[k] = make_double2((double)cosl(a), (double)sinl(a));
    }
}

struct FftPlan {
    int L, nrad;
    int rad[24];
    int tws[24];                // twiddle stride of pass s: L / (Ns R)
    int sbshift;                // log2(SB): columns per CTA are a power of two, so w / SB and w % SB are a shift and a mask
};
// j % Ns without an integer division (j < 2^20): the float quotient is exact after the + 0.5 guard
__device__ __forceinline__ int fast_mod(int j, int Ns, float invNs) { return j - Ns * (int)(((float)j + 0.5f) * invNs); }

template <int R>
__device__ __forceinline__ void butterfly(double2 (&x)[R], const double2 *tw, int L, bool inverse) {
    // in-register DFT of size R: x[q'] = sum_q x[q] w_R^(q q'),  w_R = tw[L/R] (conjugated for the inverse)
    if (R == 2) {
        double2 a = x[0], b = x[1];
        x[0] = cadd(a, b);
        x[1] = csub(a, b);
    } else if (R == 4) {
        double2 a = cadd(x[0], x[2]), b = csub(x[0], x[2]), c = cadd(x[1], x[3]), d = csub(x[1], x[3]);
        // forward: -i*d ; inverse: +i*d
        double2 jd = inverse ? make_double2(-d.y, d.x) : make_double2(d.y, -d.x);
        x[0] = cadd(a, c);
        x[1] = cadd(b, jd);
        x[2] = csub(a, c);
        x[3] = csub(b, jd);
    } else {
        double2 y[R];
#pragma unroll
        for (int qp = 0; qp < R; qp++) {
            double2 acc = x[0];
#pragma unroll
            for (int q = 1; q < R; q++) {
                double2 w = tw[((q * qp) % R) * (L / R)];
                if (inverse) w.y = -w.y;
                acc = cadd(acc, cmul(x[q], w));
            }
            y[qp] = acc;
        }
#pragma unroll
        for (int q = 0; q < R; q++) x[q] = y[q];
    }
}

template <int R>
__device__ __forceinline__ void stockham_pass(const double2 *__restrict__ src, double2 *__restrict__ dst, int L, int Ns, int SB,
                                              const double2 *tw, bool inverse, int sbshift, int twstride) {
    const int nb = L / R;                       // butterflies per column
    const float invNs = 1.0f / (float)Ns;
    for (int w = threadIdx.x; w < nb * SB; w += blockDim.x) {
        int j = w >> sbshift, col = w & (SB - 1);
        int k = fast_mod(j, Ns, invNs);
        double2 x[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            double2 v = src[(j + q * nb) * SB + col];
            if (q > 0 && Ns > 1) {
                double2 t = tw[q * k * twstride];
                if (inverse) t.y = -t.y;
                v = cmul(v, t);
            }
            x[q] = v;
        }
        butterfly<R>(x, tw, L, inverse);
        int base = (j - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; q++) dst[(base + q * Ns) * SB + col] = x[q];
    }
}

// generic radix (prime factors other than 2, 3, 5, 7): O(r^2), inputs re-read from shared memory
__device__ __forceinline__ void stockham_pass_generic(const double2 *__restrict__ src, double2 *__restrict__ dst, int L, int Ns, int R,
                                                      int SB, const double2 *tw, bool inverse) {
    const int nb = L / R;
    for (int w = threadIdx.x; w < nb * SB * R; w += blockDim.x) {
        int qp = w / (nb * SB), rest = w - qp * nb * SB;
        int j = rest / SB, col = rest - j * SB;
        int k = j % Ns;
        double2 acc = make_double2(0, 0);
        for (int q = 0; q < R; q++) {
            double2 v = src[(j + q * nb) * SB + col];
            size_t e = ((size_t)(q * k) * (L / (Ns * R)) + (size_t)((q * qp) % R) * (L / R)) % L;
            double2 t = tw[e];
            if (inverse) t.y = -t.y;
            acc = cadd(acc, cmul(v, t));
        }
        dst[((j - k) * R + k + qp * Ns) * SB + col] = acc;
    }
}

// forward: out[n][i] = scale1[n] * sum_l e^{-2 pi i n l/L} theta_l/sqrt(L) in[l][i]
// inverse: out[l][i] = conj(theta_l)/sqrt(L) * sum_n e^{+2 pi i n l/L} in[n][i]
// scale1 (may be NULL) folds the order-1 ("scalar") frequencies of the preconditioner into the store.
// dot_with (may be NULL): per-CTA partial sums of conj(dot_with).out (re, im) -> the CG r.z dot product.
__global__ void __launch_bounds__(256, 2) k_tau_fft(const FftPlan plan, double2 *__restrict__ out, const double2 *__restrict__ in, int N, int SB, int inverse,
                          int twist, const double2 *tw, const double2 *__restrict__ theta,
                          const double *__restrict__ scale1, const double2 *__restrict__ dot_with, double *__restrict__ dot_part,
                          const CgState *__restrict__ skip, size_t bstride, const FftCgUpdate U) {
    extern __shared__ double2 sm[];
    __shared__ double red[2 * 32];
    __shared__ double ush[1];
    __shared__ int ulast;
    sq_pdl_prologue();
    // batch of vectors (multi-RHS solves): blockIdx.y selects the vector, its partial sums and its solver state
    in += (size_t)blockIdx.y * bstride;
    out += (size_t)blockIdx.y * bstride;
    if (dot_with) { dot_with += (size_t)blockIdx.y * bstride; dot_part += (size_t)blockIdx.y * 2 * SQ_MAXPART; }
    if (skip) skip += blockIdx.y;
    if (U.x != nullptr) {                              // fused CG update: the state BEFORE the update decides; a finished system keeps its state
        if (U.cur[blockIdx.y].done) {
            if (blockIdx.x == 0 && threadIdx.x == 0) U.nxt[blockIdx.y] = U.cur[blockIdx.y];
            return;
        }
    } else if (skip && skip->done) return;
    const int L = plan.L;
    double2 *bufA = sm, *bufB = sm + (size_t)L * SB;
    double2 *stw = sm + (size_t)2 * L * SB;          // twiddles staged in shared memory: no global loads inside the passes
    const int i0 = blockIdx.x * SB;
    const int ncol = min(SB, N - i0);
    const double rs = rsqrt((double)L);
    for (int k = threadIdx.x; k < L; k += blockDim.x) stw[k] = tw[k];
    // fused CG update (forward transform of the preconditioner only): alpha from the p.Ap partials and the current state
    const bool fuse = (U.x != nullptr);
    double2 alpha = make_double2(0.0, 0.0);
    double2 *ux = nullptr, *ur = nullptr;
    const double2 *up = nullptr, *uq = nullptr;
    double acc_rr = 0.0;
    if (fuse) {
        const size_t bo = (size_t)blockIdx.y * bstride;
        ux = U.x + bo; ur = U.r + bo; up = U.p + bo; uq = U.q + bo;
        if (threadIdx.x < 32) {
            const double pAp = warp_sum_partials(U.pAp_part + (size_t)blockIdx.y * U.pap_stride, U.npart);
            if (threadIdx.x == 0) ush[0] = pAp;
        }
        __syncthreads();
        const CgState st = U.cur[blockIdx.y];
        alpha = make_double2(st.rz_re / ush[0], st.rz_im / ush[0]);
    }
    // tile load, 4 independent global loads in flight per thread
    const int tot = L * SB, T = blockDim.x;
    for (int w0 = threadIdx.x; w0 < tot; w0 += 4 * T) {
        double2 v[4];
        int lq[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int w = w0 + q * T;
            v[q] = make_double2(0, 0);
            lq[q] = -1;
            if (w < tot) {
                int l = w >> plan.sbshift, col = w & (SB - 1);
                lq[q] = l;
                if (col < ncol) {
                    const size_t g = (size_t)l * N + i0 + col;
                    v[q] = fuse ? U.r[(size_t)blockIdx.y * bstride + g] : in[g];      // (fused: r is read through the pointer it is written through)
                    if (fuse) {                               // x += alpha p ; r -= alpha q ; the transform continues on the new r
                        const double2 rk = csub(v[q], cmul(alpha, uq[g]));
                        ux[g] = cadd(ux[g], cmul(alpha, up[g]));
                        ur[g] = rk;
                        acc_rr += rk.x * rk.x + rk.y * rk.y;
                        v[q] = rk;
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int w = w0 + q * T;
            if (w < tot) {
                double2 x = v[q];
                if (!inverse && twist) x = cmul(x, cscale(rs, theta[lq[q]]));
                else if (!inverse) x = cscale(rs, x);
                bufA[w] = x;
            }
        }
    }
    __syncthreads();
    if (fuse) {
        // |r|^2 partial of this CTA; the CTA that finishes last sums them in a fixed order and performs the convergence test
        double a1[1] = {acc_rr};
        block_sum<1>(a1, red);
        double *rrp = U.rr_part + (size_t)blockIdx.y * SQ_MAXPART;
        if (threadIdx.x == 0) {
            rrp[blockIdx.x] = a1[0];
            __threadfence();
            ulast = (atomicAdd(U.ticket + blockIdx.y, 1u) == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (ulast && threadIdx.x < 32) {
            __threadfence();
            const volatile double *vp = rrp;
            double t = 0;
            for (int k = threadIdx.x; k < (int)gridDim.x; k += 32) t += vp[k];
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (threadIdx.x == 0) {
                CgState c = U.cur[blockIdx.y];
                c.eps = sqrt(t) / c.normb;
                c.iters = U.iter;
                c.done = (c.eps < c.tol) ? 1 : 0;
                if (!(c.eps == c.eps)) c.done = 2;
                U.nxt[blockIdx.y] = c;
                U.ticket[blockIdx.y] = 0;
            }
        }
        __syncthreads();
    }
    tw = stw;
    double2 *src = bufA, *dst = bufB;
    int Ns = 1;
    for (int s = 0; s < plan.nrad; s++) {
        int R = plan.rad[s];
        switch (R) {
            case 2: stockham_pass<2>(src, dst, L, Ns, SB, tw, inverse, plan.sbshift, plan.tws[s]); break;
            case 3: stockham_pass<3>(src, dst, L, Ns, SB, tw, inverse, plan.sbshift, plan.tws[s]); break;
            case 4: stockham_pass<4>(src, dst, L, Ns, SB, tw, inverse, plan.sbshift, plan.tws[s]); break;
            case 5: stockham_pass<5>(src, dst, L, Ns, SB, tw, inverse, plan.sbshift, plan.tws[s]); break;
            case 7: stockham_pass<7>(src, dst, L, Ns, SB, tw, inverse, plan.sbshift, plan.tws[s]); break;
            default: stockham_pass_generic(src, dst, L, Ns, R, SB, tw, inverse); break;
        }
        __syncthreads();
        Ns *= R;
        double2 *t = src; src = dst; dst = t;
    }
    double acc_re = 0, acc_im = 0;
    for (int w = threadIdx.x; w < L * SB; w += blockDim.x) {
        int l = w >> plan.sbshift, col = w & (SB - 1);
        if (col >= ncol) continue;
        double2 v = src[w];
        if (inverse) {
            if (twist) { double2 t = theta[l]; v = cmul(v, make_double2(rs * t.x, -rs * t.y)); }
            else v = cscale(rs, v);
        } else if (scale1) v = cscale(scale1[l], v);
        size_t g = (size_t)l * N + i0 + col;
        out[g] = v;
        if (dot_with) {
            double2 r = dot_with[g];
            acc_re += r.x * v.x + r.y * v.y;
            acc_im += r.x * v.y - r.y * v.x;
        }
    }
    if (dot_with) {
        double v2[2] = {acc_re, acc_im};
        block_sum<2>(v2, red);
        if (threadIdx.x == 0) { dot_part[blockIdx.x] = v2[0]; dot_part[SQ_MAXPART + blockIdx.x] = v2[1]; }
    }
}

// Launch helper shared by the preconditioner (complex [l][i] vectors) and the EFA (phonon fields [l][p]).
// Returns the number of CTAs (= number of dot partials when dot_with != NULL).
int tau_fft_launch_batch(cudaStream_t stream, const std::vector<int> &radices, int L, int N, double2 *out, const double2 *in, bool inverse,
                         bool twist, const double2 *tw, const double2 *theta, const double *scale1, const double2 *dot_with,
                         double *dot_part, const CgState *skip, size_t smem_limit, int nbatch, size_t bstride, const FftCgUpdate *upd);
int tau_fft_launch(cudaStream_t stream, const std::vector<int> &radices, int L, int N, double2 *out, const double2 *in, bool inverse,
                   bool twist, const double2 *tw, const double2 *theta, const double *scale1, const double2 *dot_with,
                   double *dot_part, const CgState *skip, size_t smem_limit) {
    return tau_fft_launch_batch(stream, radices, L, N, out, in, inverse, twist, tw, theta, scale1, dot_with, dot_part, skip, smem_limit, 1, 0, nullptr);
}
// nbatch vectors, bstride elements apart (dot_with likewise; dot_part 2 SQ_MAXPART doubles apart; skip[] one state per vector)
int tau_fft_launch_batch(cudaStream_t stream, const std::vector<int> &radices, int L, int N, double2 *out, const double2 *in, bool inverse,
                         bool twist, const double2 *tw, const double2 *theta, const double *scale1, const double2 *dot_with,
                         double *dot_part, const CgState *skip, size_t smem_limit, int nbatch, size_t bstride, const FftCgUpdate *upd) {
    FftPlan plan;
    plan.L = L;
    plan.nrad = (int)radices.size();
    if (plan.nrad > 24) throw SqError("FFT length has too many prime factors");
    for (int s = 0; s < plan.nrad; s++) plan.rad[s] = radices[s];
    int SB = 8;
    while (SB > 1 && (size_t)(2 * SB + 1) * L * sizeof(double2) > smem_limit) SB >>= 1;
    // prefer more CTAs when the lattice is small
    while (SB > 2 && (size_t)nbatch * ((N + SB - 1) / SB) < 148) SB >>= 1;
    size_t smem = (size_t)(2 * SB + 1) * L * sizeof(double2);
    if (smem > smem_limit) throw SqError("imaginary-time axis too long for the shared-memory FFT");
    static bool attr = false;
    if (!attr) {
        SQ_CUDA(cudaFuncSetAttribute(k_tau_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
        SQ_CUDA(cudaFuncSetAttribute(k_tau_fft, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr = true;
    }
    int grid = (N + SB - 1) / SB;
    if (dot_with && grid > SQ_MAXPART) throw SqError("lattice too large for the FFT partial-sum buffer");
    if (const char *e = getenv("SQ_FFT_SB")) { int v = atoi(e); if (v >= 1 && (size_t)(2 * v + 1) * L * sizeof(double2) <= smem_limit) { SB = v; smem = (size_t)(2 * SB + 1) * L * sizeof(double2); grid = (N + SB - 1) / SB; } }
    // 128 ... 256 threads: at ~90 registers per thread two CTAs of 256 threads share an SM, so the 256 CTAs of the named size run as ONE
    // wave (416 threads = one butterfly per thread in the radix-4 passes meant one CTA per SM and two waves: ncu, profiles/r2_*fft*)
    int threads = std::min(256, std::max(128, ((L * SB / 4 + 31) / 32) * 32));
    if (const char *e = getenv("SQ_FFT_T")) threads = std::max(32, std::min(256, atoi(e)));
    plan.sbshift = 0;
    while ((1 << plan.sbshift) < SB) plan.sbshift++;
    if ((1 << plan.sbshift) != SB) throw SqError("FFT tile width must be a power of two");
    {
        int Ns = 1;
        for (int s = 0; s < plan.nrad; s++) { plan.tws[s] = L / (Ns * plan.rad[s]); Ns *= plan.rad[s]; }
    }
    FftCgUpdate U;
    memset(&U, 0, sizeof(U));
    if (upd) {
        if (inverse || grid > SQ_MAXPART) throw SqError("fused CG update: forward transform with at most SQ_MAXPART CTAs only");
        U = *upd;
    }
    SQ_CUDA(sq_launch(k_tau_fft, dim3(grid, nbatch), dim3(threads), smem, stream, plan, out, in, N, SB, inverse ? 1 : 0, twist ? 1 : 0, tw, theta, scale1,
                      dot_with, dot_part, skip, bstride, U));
    SQ_LAUNCH_CHECK();
    return grid;
}

// kpm.cu -- K4: the KPM / tau-Fourier preconditioner  P^-1 = [Mbar^T Mbar]^-1.
//
// Replaces src/KPMPreconditioner.jl: ldiv! (Sym complex :355-414, Asym complex :488-550),
// update_preconditioner! (:554-597), update_B̄! (:604-621), calculate_bounds! (:625-658),
// update_kpm_expansion_order!/coefs! (:696-795), and the un-vendored SmoQyKPMCore kernels kpm_lmul!,
// kpm_coefs!, lanczos! plus JDQMCFramework's Sym/AsymChkbrdPropagator mul! that they call.
//
// Apply = forward tau-FFT (fft.cu; also applies the order-1 "scalar" frequencies) -> one CTA per
// frequency with expansion order > 1 runs the whole Chebyshev recurrence of B-bar on its N-vector in
// shared memory -> inverse tau-FFT (optionally fused with the CG r.z reduction).  The frequency-major
// layout falls out of the [l][i] device layout, so the reference's two transposes do not exist here.
// Frequencies are scheduled longest-recurrence-first.  Bounds, orders and coefficients are tiny scalar
// work and are computed on the host in double precision exactly as the reference does.
#include "sq_internal.h"

#include <algorithm>
#include <cmath>

int tau_fft_launch(cudaStream_t stream, const std::vector<int> &radices, int L, int N, double2 *out, const double2 *in, bool inverse,
                   bool twist, const double2 *tw, const double2 *theta, const double *scale1, const double2 *dot_with,
                   double *dot_part, const CgState *skip, size_t smem_limit);
bool kpm_reg_ok(const sq_kpm *k);
void kpm_cheb_reg_launch(sq_kpm *k, double2 *z, const int *d_sched, int nsched, int nrhs, size_t rhs_stride, const CgState *skip);

struct BbarParams {
    int N, Nh, C, sym;
    int clo[SQ_MAXC], chi[SQ_MAXC];
    const int2 *nt;
    const double2 *csbar;
    const double *Dbar;
};

__device__ __forceinline__ void rotb(double2 &a, double2 &b, double c, double s) {
    double2 na = make_double2(fma(s, b.x, c * a.x), fma(s, b.y, c * a.y));
    double2 nb = make_double2(fma(s, a.x, c * b.x), fma(s, a.y, c * b.y));
    a = na;
    b = nb;
}

__device__ __forceinline__ void bbar_color(double2 *y, int c, const BbarParams &P) {
    const int lo = P.clo[c], nb = P.chi[c] - lo;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        int2 ij = __ldg(P.nt + lo + b);
        double2 cs = __ldg(P.csbar + lo + b);
        double2 a = y[ij.x], bb = y[ij.y];
        rotb(a, bb, cs.x, cs.y);
        y[ij.x] = a;
        y[ij.y] = bb;
    }
    __syncthreads();
}
__device__ __forceinline__ void bbar_diag(double2 *y, const BbarParams &P, bool squared) {
    for (int i = threadIdx.x; i < P.N; i += blockDim.x) {
        double d = __ldg(P.Dbar + i);
        if (squared) d *= d;
        double2 a = y[i];
        y[i] = make_double2(d * a.x, d * a.y);
    }
    __syncthreads();
}
// y <- B-bar y in shared memory.  Sym: Gamma D Gamma^T; Asym: D Gamma.  (KPMPreconditioner.jl:260,274)
__device__ __forceinline__ void bbar_apply(double2 *y, const BbarParams &P) {
    if (P.sym) {
        for (int c = P.C - 1; c >= 0; c--) bbar_color(y, c, P);
        bbar_diag(y, P, false);
        for (int c = 0; c < P.C; c++) bbar_color(y, c, P);
    } else {
        for (int c = 0; c < P.C; c++) bbar_color(y, c, P);
        bbar_diag(y, P, false);
    }
}
// y <- B-bar^T B-bar y (Asym Lanczos operator, :661-679)
__device__ __forceinline__ void bbar_apply_BtB(double2 *y, const BbarParams &P) {
    for (int c = 0; c < P.C; c++) bbar_color(y, c, P);
    bbar_diag(y, P, true);
    for (int c = P.C - 1; c >= 0; c--) bbar_color(y, c, P);
}

// v <- sum_q c_q T_q(B') v with B' = (B-bar - avg)/mag   [kpm_lmul!]
__device__ __forceinline__ void cheb_apply(double2 *T0, double2 *T1, double2 *Y, double2 *ACC, const double2 *__restrict__ c, int ord,
                                           double avg, double imag_, const BbarParams &P) {
    const int N = P.N;
    // entry: T0 = v.  T1 = B' v
    for (int i = threadIdx.x; i < N; i += blockDim.x) Y[i] = T0[i];
    __syncthreads();
    bbar_apply(Y, P);
    double2 c0 = c[0], c1 = c[1];
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double2 t0 = T0[i], y = Y[i];
        double2 t1 = make_double2((y.x - avg * t0.x) * imag_, (y.y - avg * t0.y) * imag_);
        T1[i] = t1;
        Y[i] = t1;
        ACC[i] = cadd(cmul(c0, t0), cmul(c1, t1));
    }
    __syncthreads();
    for (int q = 2; q < ord; q++) {
        bbar_apply(Y, P);                              // Y = B-bar T1 (Y held a copy of T1)
        double2 cq = c[q];
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            double2 t0 = T0[i], t1 = T1[i], y = Y[i];
            double2 t2 = make_double2(2.0 * (y.x - avg * t1.x) * imag_ - t0.x, 2.0 * (y.y - avg * t1.y) * imag_ - t0.y);
            T0[i] = t1;                                // shift the window: (T0, T1) <- (T1, T2)
            T1[i] = t2;
            Y[i] = t2;
            ACC[i] = cadd(ACC[i], cmul(cq, t2));
        }
        __syncthreads();
    }
}

// one CTA per scheduled frequency.  z is [n][i].  For the Asym preconditioner two expansions are applied
// back to back (coefficients of the mirrored frequency, then of this one).
__global__ void __launch_bounds__(1024, 1)
k_kpm_cheb(const __grid_constant__ BbarParams P, double2 *__restrict__ z, const int *__restrict__ sched, const int *__restrict__ order,
           const int *__restrict__ coef_off, const double2 *__restrict__ coefs, int L, double avg, double imag_,
           const CgState *__restrict__ skip) {
    extern __shared__ double2 sm[];
    sq_pdl_prologue();
    if (skip && skip->done) return;
    const int N = P.N;
    double2 *T0 = sm, *T1 = sm + N, *Y = sm + 2 * (size_t)N, *ACC = sm + 3 * (size_t)N;
    const int n = sched[blockIdx.x];
    double2 *zn = z + (size_t)n * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) T0[i] = zn[i];
    __syncthreads();
    if (P.sym) {
        int np = (n + 1 > (L + 1) / 2) ? L - 1 - n : n;                       // :387
        cheb_apply(T0, T1, Y, ACC, coefs + coef_off[np], order[np], avg, imag_, P);
    } else {
        int nm = L - 1 - n;
        cheb_apply(T0, T1, Y, ACC, coefs + coef_off[nm], order[nm], avg, imag_, P);   // :527
        for (int i = threadIdx.x; i < N; i += blockDim.x) T0[i] = ACC[i];
        __syncthreads();
        cheb_apply(T0, T1, Y, ACC, coefs + coef_off[n], order[n], avg, imag_, P);     // :530
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) zn[i] = ACC[i];
}

// ---------------------------------------------------------------------------------------------------
// Fast Chebyshev kernel (symmetric propagator, every site touched by the last colour, <= 8 colours, one bond per
// thread per colour).  The recurrence is latency bound (156 sequential B-bar applications at cfg4), so everything
// is arranged to shorten the dependent chain of one application:
//   * the thread that owns bond (i, j) of the LAST colour keeps T_{k-1}, T_k and the accumulator of sites i and j in
//     registers for the whole kernel; shared memory only holds the vector being propagated;
//   * Gamma^T starts and Gamma ends with that same colour, so an application starts from and ends in registers:
//     2C - 2 barriers per application instead of 2C + 3, no copy and no separate recurrence pass;
//   * bond slots, (cosh, sinh) means and the diagonal means live in registers (no global loads in the loop);
//   * sites are addressed through the bank-conflict-free slot numbering of the fused matvec.
// ---------------------------------------------------------------------------------------------------
struct BbarFast {
    int N, C, nunc0;
    int clo[8], chi[8];
    const int2 *nts;        // bond -> slots
    const int2 *nt;         // bond -> sites
    const int *slot, *unc0;
    const double2 *csbar;
    const double *Dbar;
    long long *dbg;
};

// B-bar is real, so the real and imaginary parts of a frequency vector are two independent real recurrences: they
// run in two CTAs on different SMs.  A colour step is bound by the shared-memory port of its SM (measured with
// tools/ubench/stage_latency.cu: ~0.6 cycles per bond item of 64 B, i.e. 85 % of 128 B/clk), so halving the bytes
// per item (LDS.64 / STS.64 on doubles) halves the length of the critical path.
template <int CMAX>
struct ChebEngine {
    int c_i[CMAX], c_j[CMAX];      // slots of this thread's bond in every colour (-1: none)
    double cc[CMAX], ss[CMAX];
    double d_i, d_j;
    int last;                      // C - 1

    __device__ __forceinline__ static void rot(double &a, double &b, double c, double s) {
        double na = fma(s, b, c * a), nb = fma(s, a, c * b);
        a = na;
        b = nb;
    }
    __device__ __forceinline__ void init(const BbarFast &P) {
        const int tx = threadIdx.x;
        last = P.C - 1;
#pragma unroll
        for (int c = 0; c < CMAX; c++) {
            c_i[c] = -1; c_j[c] = 0; cc[c] = 1.0; ss[c] = 0.0;
            if (c < P.C && tx < P.chi[c] - P.clo[c]) {
                int2 ij = __ldg(P.nts + P.clo[c] + tx);
                double2 v = __ldg(P.csbar + P.clo[c] + tx);
                c_i[c] = ij.x; c_j[c] = ij.y; cc[c] = v.x; ss[c] = v.y;
            }
        }
        d_i = d_j = 1.0;
        if (tx < P.chi[0] - P.clo[0]) {
            int2 q = __ldg(P.nt + P.clo[0] + tx);
            d_i = __ldg(P.Dbar + q.x);
            d_j = __ldg(P.Dbar + q.y);
        }
    }
    template <int Q>
    __device__ __forceinline__ void smem_step(double *Y) {
        if (c_i[Q] >= 0) {
            double a = Y[c_i[Q]], b = Y[c_j[Q]];
            rot(a, b, cc[Q], ss[Q]);
            Y[c_i[Q]] = a;
            Y[c_j[Q]] = b;
        }
        __syncthreads();
    }
    template <int Q>
    __device__ __forceinline__ void down(double *Y) {       // colours last-1 ... 1
        if constexpr (Q >= 1) {
            if (Q < last) smem_step<Q>(Y);
            down<Q - 1>(Y);
        }
    }
    template <int Q>
    __device__ __forceinline__ void up(double *Y) {         // colours 1 ... last-1
        if constexpr (Q < CMAX) {
            if (Q < last) smem_step<Q>(Y);
            up<Q + 1>(Y);
        }
    }
    __device__ __forceinline__ void mid(double &a, double &b) {
        rot(a, b, cc[0], ss[0]);
        a *= d_i;
        b *= d_j;
        rot(a, b, cc[0], ss[0]);
    }
    // (a, b) <- (B-bar y)_{i,j} for the thread's last-colour bond (slots oi, oj; coefficients cl, sl)
    __device__ __forceinline__ void apply(double &a, double &b, double *Y, const BbarFast &P, int oi, int oj, double cl, double sl) {
        if (last == 0) { mid(a, b); return; }             // single colour: all in registers
        if (oi >= 0) {
            rot(a, b, cl, sl);
            Y[oi] = a;
            Y[oj] = b;
        }
        __syncthreads();
        down<CMAX - 1>(Y);
        if (c_i[0] >= 0) {
            double u = Y[c_i[0]], v = Y[c_j[0]];
            mid(u, v);
            Y[c_i[0]] = u;
            Y[c_j[0]] = v;
        }
        for (int q = threadIdx.x; q < P.nunc0; q += blockDim.x) {
            int i = __ldg(P.unc0 + q);
            int s = __ldg(P.slot + i);
            Y[s] *= __ldg(P.Dbar + i);
        }
        __syncthreads();
        up<1>(Y);
        if (oi >= 0) {
            a = Y[oi];
            b = Y[oj];
            rot(a, b, cl, sl);
        }
    }
};

// grid = 2 * (#frequencies with order > 1): blockIdx.x = 2 * schedule index + (0: real part, 1: imaginary part)
template <int CMAX, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
k_kpm_cheb_fast(const __grid_constant__ BbarFast P, double2 *__restrict__ z, const int *__restrict__ sched, const int *__restrict__ order,
                const int *__restrict__ coef_off, const double2 *__restrict__ coefs, int L, double avg, double imag_,
                const CgState *__restrict__ skip) {
    extern __shared__ double Yr[];
    sq_pdl_prologue();
    if (skip && skip->done) return;
    ChebEngine<CMAX> E;
    E.init(P);
    const int n = sched[blockIdx.x >> 1], part = blockIdx.x & 1;
    const int np = (n + 1 > (L + 1) / 2) ? L - 1 - n : n;
    const int ord = order[np];
    const double2 *c = coefs + coef_off[np];                 // real coefficients (Sym): .x
    double *zn = reinterpret_cast<double *>(z + (size_t)n * P.N) + part;
    // the thread owns the two sites of its bond in the last colour
    int si = -1, sj = -1, oi = -1, oj = 0;
    double cl = 1.0, sl = 0.0;
    if (threadIdx.x < P.chi[E.last] - P.clo[E.last]) {
        int2 q = __ldg(P.nt + P.clo[E.last] + threadIdx.x);
        int2 o = __ldg(P.nts + P.clo[E.last] + threadIdx.x);
        double2 v = __ldg(P.csbar + P.clo[E.last] + threadIdx.x);
        si = q.x; sj = q.y; oi = o.x; oj = o.y; cl = v.x; sl = v.y;
    }
    double t0i = 0, t0j = 0;
    if (si >= 0) { t0i = zn[2 * si]; t0j = zn[2 * sj]; }
    double yi = t0i, yj = t0j;
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[0] = clock64();
    E.apply(yi, yj, Yr, P, oi, oj, cl, sl);
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) { P.dbg[1] = clock64(); P.dbg[3] = ord; }
    double t1i = (yi - avg * t0i) * imag_, t1j = (yj - avg * t0j) * imag_;
    const double c0 = c[0].x, c1 = c[1].x;
    double acci = c0 * t0i + c1 * t1i, accj = c0 * t0j + c1 * t1j;
    for (int q = 2; q < ord; q++) {
        yi = t1i; yj = t1j;
        E.apply(yi, yj, Yr, P, oi, oj, cl, sl);
        double t2i = 2.0 * (yi - avg * t1i) * imag_ - t0i, t2j = 2.0 * (yj - avg * t1j) * imag_ - t0j;
        const double cq = c[q].x;
        acci = fma(cq, t2i, acci);
        accj = fma(cq, t2j, accj);
        t0i = t1i; t0j = t1j; t1i = t2i; t1j = t2j;
    }
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[2] = clock64();
    if (si >= 0) { zn[2 * si] = acci; zn[2 * sj] = accj; }
}

// tau-means of the operator coefficients (update_B̄!, :604-621): grid over sites/bonds, 32 x 8 threads
__global__ void k_tau_means(double *__restrict__ Dbar, double2 *__restrict__ csbar, const double *__restrict__ expV,
                            const double2 *__restrict__ cs, int L, int N, int Nh) {
    __shared__ double sx[8][33], sy[8][33];
    int idx = blockIdx.x * 32 + threadIdx.x;
    double ax = 0, ay = 0;
    if (idx < N) {
        for (int l = threadIdx.y; l < L; l += 8) ax += expV[(size_t)l * N + idx];
    } else if (idx - N < Nh && idx >= N) {
        int h = idx - N;
        for (int l = threadIdx.y; l < L; l += 8) { double2 v = cs[(size_t)l * Nh + h]; ax += v.x; ay += v.y; }
    }
    sx[threadIdx.y][threadIdx.x] = ax;
    sy[threadIdx.y][threadIdx.x] = ay;
    __syncthreads();
    if (threadIdx.y == 0) {
        double tx = 0, ty = 0;
        for (int k = 0; k < 8; k++) { tx += sx[k][threadIdx.x]; ty += sy[k][threadIdx.x]; }
        if (idx < N) Dbar[idx] = tx / L;
        else if (idx - N < Nh) csbar[idx - N] = make_double2(tx / L, ty / L);
    }
}

// Lanczos on B-bar (Sym) or B-bar^T B-bar (Asym) in one CTA  [lanczos!], plain three-term recurrence.
// out[0..n) = alpha, out[n..2n-1) = beta.
__global__ void __launch_bounds__(1024, 1)
k_lanczos(const __grid_constant__ BbarParams P, const double *__restrict__ start, int n, double *__restrict__ out) {
    extern __shared__ double2 sm[];
    __shared__ double red[32];
    __shared__ double bc;
    const int N = P.N;
    double2 *vp = sm, *v = sm + N, *w = sm + 2 * (size_t)N;
    double acc = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) { double s = start[i]; acc += s * s; }
    double t1[1] = {acc};
    block_sum<1>(t1, red);
    if (threadIdx.x == 0) bc = sqrt(t1[0]);
    __syncthreads();
    double nrm = bc;
    for (int i = threadIdx.x; i < N; i += blockDim.x) { v[i] = make_double2(start[i] / nrm, 0.0); vp[i] = make_double2(0, 0); }
    __syncthreads();
    double bprev = 0;
    for (int j = 0; j < n; j++) {
        for (int i = threadIdx.x; i < N; i += blockDim.x) w[i] = v[i];
        __syncthreads();
        if (P.sym) bbar_apply(w, P); else bbar_apply_BtB(w, P);
        acc = 0;
        for (int i = threadIdx.x; i < N; i += blockDim.x) acc += v[i].x * w[i].x;
        t1[0] = acc;
        block_sum<1>(t1, red);
        if (threadIdx.x == 0) bc = t1[0];
        __syncthreads();
        double a = bc;
        acc = 0;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            double x = w[i].x - a * v[i].x - bprev * vp[i].x;
            w[i].x = x;
            acc += x * x;
        }
        t1[0] = acc;
        block_sum<1>(t1, red);
        if (threadIdx.x == 0) { bc = sqrt(t1[0]); out[j] = a; if (j < n - 1) out[n + j] = bc; }
        __syncthreads();
        double b = bc;
        if (j < n - 1) {
            // breakdown (b == 0) leaves zeros in the remaining entries; the host truncates there
            if (b < 1e-300) { for (int q = j + 1 + threadIdx.x; q < n; q += blockDim.x) { out[q] = 0; if (q < n - 1) out[n + q] = 0; } return; }
            for (int i = threadIdx.x; i < N; i += blockDim.x) { vp[i] = v[i]; v[i] = make_double2(w[i].x / b, 0.0); }
            bprev = b;
            __syncthreads();
        }
    }
}

// global-memory fallback of B-bar v is not provided: the preconditioner requires 4 N-vectors in shared memory.

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static BbarParams bbar_params(const sq_kpm *k) {
    const sq_fdm *f = k->f;
    BbarParams P;
    P.N = (int)f->N; P.Nh = (int)f->Nh; P.C = (int)f->C; P.sym = f->sym;
    for (int c = 0; c < SQ_MAXC; c++) { P.clo[c] = c < f->C ? f->clo[c] : 0; P.chi[c] = c < f->C ? f->chi[c] : 0; }
    P.nt = f->nt.p; P.csbar = k->csbar.p; P.Dbar = k->Dbar.p;
    return P;
}

static int kpm_threads(const sq_kpm *k) {
    int nbmax = 32;
    for (int c = 0; c < k->f->C; c++) nbmax = std::max(nbmax, k->f->chi[c] - k->f->clo[c]);
    int t = 32;
    while (t < nbmax && t < 1024) t <<= 1;
    return t;
}

// the fast Chebyshev kernel needs: Sym, <= 8 colours, one bond per thread per colour, every site in the last colour
static bool kpm_fast_ok(const sq_kpm *k) {
    const sq_fdm *f = k->f;
    if (!f->sym || f->C < 1 || f->C > 8 || getenv("SQ_KPM_SLOW")) return false;
    int nbmax = 0;
    for (int c = 0; c < f->C; c++) nbmax = std::max(nbmax, f->chi[c] - f->clo[c]);
    if (nbmax > 1024) return false;
    int last = (int)f->C - 1;
    return 2 * (f->chi[last] - f->clo[last]) == f->N;
}
static int kpm_fast_threads(const sq_kpm *k) {
    int nbmax = 1;
    for (int c = 0; c < k->f->C; c++) nbmax = std::max(nbmax, k->f->chi[c] - k->f->clo[c]);
    return ((nbmax + 31) / 32) * 32;
}

static void sturm_extremes(const std::vector<double> &a, const std::vector<double> &b, int n, double *emin, double *emax) {
    auto count = [&](double x) {
        int cnt = 0;
        double d = 1.0;
        for (int i = 0; i < n; i++) {
            double bb = (i == 0) ? 0.0 : b[i - 1] * b[i - 1];
            d = a[i] - x - (d != 0.0 ? bb / d : bb / 1e-300);
            if (d < 0) cnt++;
        }
        return cnt;
    };
    double lo = a[0], hi = a[0];
    for (int i = 0; i < n; i++) {
        double r = (i > 0 ? fabs(b[i - 1]) : 0) + (i < n - 1 ? fabs(b[i]) : 0);
        lo = std::min(lo, a[i] - r);
        hi = std::max(hi, a[i] + r);
    }
    double l = lo, h = hi;
    for (int it = 0; it < 200; it++) { double m = 0.5 * (l + h); if (count(m) >= 1) h = m; else l = m; }
    *emin = 0.5 * (l + h);
    l = lo; h = hi;
    for (int it = 0; it < 200; it++) { double m = 0.5 * (l + h); if (count(m) >= n) h = m; else l = m; }
    *emax = 0.5 * (l + h);
}

static void kpm_refresh_bbar(sq_kpm *k) {
    sq_fdm *f = k->f;
    if (k->bbar_version == f->coef_version) return;      // operator unchanged since the last refresh
    int tot = (int)(f->N + f->Nh);
    k_tau_means<<<(tot + 31) / 32, dim3(32, 8), 0, f->stream>>>(k->Dbar.p, k->csbar.p, f->expV.p, f->cs.p, (int)f->L, (int)f->N, (int)f->Nh);
    SQ_LAUNCH_CHECK();
    f->launches++;
    k->bbar_version = f->coef_version;
}

// Chebyshev-Gauss coefficients [kpm_coefs!]: c_q = (2 - delta_q0)/Nq sum_j f(x_j) cos(q pi (j + 1/2)/Nq), Nq = 2 order
static void cheb_coefs(std::vector<double2> &c, int order, bool sym, double phi, const double bounds[2]) {
    const double PI = 3.14159265358979323846;
    int Nq = 2 * order;
    double avg = 0.5 * (bounds[1] + bounds[0]), mag = 0.5 * (bounds[1] - bounds[0]);
    std::vector<double2> fx(Nq);
    for (int j = 0; j < Nq; j++) {
        double x = mag * cos(PI * (j + 0.5) / Nq) + avg;
        if (sym) fx[j] = make_double2(1.0 / (x * x - 2 * x * cos(phi) + 1), 0.0);            // f_B̄_sym :800
        else {                                                                                // f_B̄_asym :804
            double re = 1.0 - cos(phi) * x, im = sin(phi) * x, d = re * re + im * im;
            fx[j] = make_double2(re / d, -im / d);
        }
    }
    c.assign(order, make_double2(0, 0));
    for (int q = 0; q < order; q++) {
        double sr = 0, si = 0;
        for (int j = 0; j < Nq; j++) {
            double cs = cos(PI * q * (j + 0.5) / Nq);
            sr += fx[j].x * cs;
            si += fx[j].y * cs;
        }
        double w = (q == 0 ? 1.0 : 2.0) / Nq;
        c[q] = make_double2(w * sr, w * si);
    }
}

// update_kpm_expansion_order! + update_kpm_expansion_coefs! (:696-795) and upload
static void kpm_update_expansions(sq_kpm *k) {
    const double PI = 3.14159265358979323846;
    sq_fdm *f = k->f;
    int L = (int)f->L;
    double emin = k->bounds[0], emax = k->bounds[1];
    for (i64 l = 0; l < k->ncoef; l++) {
        double phi = 2 * PI / L * (l + 0.5);
        double ph = phi > PI ? 2 * PI - phi : phi;
        i64 n = (i64)floor((emax - emin) * (k->a1 / ph + k->a2));
        k->order[l] = std::max<i64>(1, n);
    }
    if (f->sym) {
        for (i64 l = 0; l < k->ncoef; l++) cheb_coefs(k->coefs[l], (int)k->order[l], true, 2 * PI / L * (l + 0.5), k->bounds);
    } else {
        for (i64 l = 0; l < (L + 1) / 2; l++) {
            cheb_coefs(k->coefs[l], (int)k->order[l], false, 2 * PI / L * (l + 0.5), k->bounds);
            i64 m = L - 1 - l;
            k->coefs[m] = k->coefs[l];
            for (auto &c : k->coefs[m]) c.y = -c.y;
        }
    }
    // flatten and upload; build the launch schedule (frequencies with order > 1, longest first) and the
    // per-frequency scalar for order-1 frequencies
    std::vector<int> h_order(k->ncoef), h_off(k->ncoef);
    std::vector<double2> flat;
    int maxo = 0;
    for (i64 l = 0; l < k->ncoef; l++) {
        h_order[l] = (int)k->order[l];
        h_off[l] = (int)flat.size();
        flat.insert(flat.end(), k->coefs[l].begin(), k->coefs[l].end());
        if (k->coefs[l].size() < 2) flat.push_back(make_double2(0, 0));   // c[1] is always readable
        maxo = std::max(maxo, h_order[l]);
    }
    k->max_order = maxo;
    std::vector<std::pair<int, int>> sched;
    std::vector<double> scale1(L, 1.0);
    for (int n = 0; n < L; n++) {
        int np = f->sym ? ((n + 1 > (L + 1) / 2) ? L - 1 - n : n) : n;
        if (h_order[np] > 1) sched.push_back({-h_order[np], n});
        else {
            double2 c = k->coefs[np][0];
            scale1[n] = f->sym ? c.x : (c.x * c.x + c.y * c.y);              // :398 / :534
        }
    }
    std::sort(sched.begin(), sched.end());
    std::vector<int> h_sched;
    for (auto &p : sched) h_sched.push_back(p.second);
    k->nsched = (int)h_sched.size();
    k->h_sched = h_sched;
    k->sched_version++;
    k->d_order.alloc(h_order.size() + 1, false); k->d_order.upload(h_order.data(), h_order.size(), f->stream);
    k->d_coef_off.alloc(h_off.size() + 1, false); k->d_coef_off.upload(h_off.data(), h_off.size(), f->stream);
    k->d_coefs.alloc(flat.size() + 1, false); k->d_coefs.upload(flat.data(), flat.size(), f->stream);
    k->d_freq_sched.alloc(h_sched.size() + 1, false); k->d_freq_sched.upload(h_sched.data(), h_sched.size(), f->stream);
    k->d_scale1.alloc(L, false); k->d_scale1.upload(scale1.data(), L, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
}

void kpm_create_impl(sq_kpm **out, sq_fdm *f, double rbuf, i64 n, double a1, double a2) {
    SQ_REQUIRE(out && f, "NULL argument");
    SQ_REQUIRE(n >= 2 && n <= 512, "number of Lanczos iterations out of range");
    SQ_REQUIRE((size_t)4 * f->N * sizeof(double2) <= f->smem_optin, "lattice too large for the shared-memory KPM kernel");
    SQ_CUDA(cudaSetDevice(f->device));
    sq_kpm *k = new sq_kpm();
    try {
        k->f = f; k->rbuf = rbuf; k->nlanczos = n; k->a2 = a2;
        k->a1 = f->sym ? 2 * a1 : a1;                                   // :263
        k->ncoef = f->sym ? (f->L + 1) / 2 : f->L;
        k->order.assign(k->ncoef, 0);
        k->coefs.resize(k->ncoef);
        k->Dbar.alloc(f->N); k->csbar.alloc(f->Nh + 1);
        std::vector<double2> tw, th(f->L);
        fft_make_twiddles(f->L, tw);
        const long double PI = 3.141592653589793238462643383279502884L;
        for (i64 l = 0; l < f->L; l++) {
            long double a = -PI * (long double)l / (long double)f->L;   // FourierTransformer.jl:15
            th[l] = make_double2((double)cosl(a), (double)sinl(a));
        }
        k->tw.alloc(f->L, false); k->tw.upload(tw.data(), f->L, f->stream);
        k->theta.alloc(f->L, false); k->theta.upload(th.data(), f->L, f->stream);
        SQ_CUDA(cudaStreamSynchronize(f->stream));
        fft_radices(f->L, k->radices);
        k->ztmp.alloc((size_t)f->L * f->N);
        k->lan.alloc(2 * n);
        k->lan_start.alloc(f->N);
        SQ_CUDA(cudaFuncSetAttribute(k_kpm_cheb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_optin));
        SQ_CUDA(cudaFuncSetAttribute(k_lanczos, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_optin));
        // (the fast Chebyshev kernels need at most N double2 of dynamic shared memory: below the 48 KB default for N <= 3072)
    } catch (...) {
        delete k;
        throw;
    }
    *out = k;
}

void kpm_set_bounds(sq_kpm *k, double emin, double emax) {
    kpm_refresh_bbar(k);
    k->bounds[0] = emin; k->bounds[1] = emax; k->active = 1;
    kpm_update_expansions(k);
}

// calculate_bounds! on the device; returns the raw Lanczos extremes
void kpm_lanczos(sq_kpm *k, const double *h_start, const double *d_start, double *emin, double *emax) {
    sq_fdm *f = k->f;
    kpm_refresh_bbar(k);
    const double *start = d_start;
    if (!start) {
        if (h_start) k->lan_start.upload(h_start, f->N, f->stream);
        else rng_fill_normal(k->lan_start.p, f->N, k->seed, sq_rng_stream(SQ_RNG_KPM, k->rng_counter++), f->stream);
        start = k->lan_start.p;
    }
    int n = (int)k->nlanczos;
    BbarParams P = bbar_params(k);
    k_lanczos<<<1, kpm_threads(k), 3 * f->N * sizeof(double2), f->stream>>>(P, start, n, k->lan.p);
    SQ_LAUNCH_CHECK();
    f->launches++;
    std::vector<double> h(2 * n);
    k->lan.download(h.data(), 2 * n, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    std::vector<double> a(h.begin(), h.begin() + n), b(h.begin() + n, h.begin() + 2 * n - 1);
    int used = n;
    for (int j = 0; j < n - 1; j++) if (!(b[j] >= 1e-300)) { used = j + 1; break; }
    for (int j = 0; j < used; j++) if (!(a[j] == a[j])) throw SqNumericalInstability("KPM preconditioner: NaN in the Lanczos recurrence");
    sturm_extremes(a, b, used, emin, emax);
    if (!f->sym) { *emin = sqrt(*emin); *emax = sqrt(*emax); }          // :655
}

// update_preconditioner! (:554-597)
void kpm_update(sq_kpm *k, const double *h_start, const double *d_start) {
    double emin, emax;
    kpm_lanczos(k, h_start, d_start, &emin, &emax);
    emin *= (1 - k->rbuf);
    emax *= (1 + k->rbuf);
    if (0.0 < emin && emin < 1.0 && 1.0 < emax && emax < 2.0) {
        k->active = 1;
        double e0 = k->bounds[0], e1 = k->bounds[1];
        if (e0 == 0.0 || fabs((emin - e0) / e0) > k->rbuf / 2 || fabs((emax - e1) / e1) > k->rbuf / 2) {
            k->bounds[0] = emin; k->bounds[1] = emax;
            kpm_update_expansions(k);
        }
    } else k->active = 0;
}

// The Chebyshev stage on frequency-major arrays z ([rhs][n][i], rhs_stride elements apart): sum_q c_q T_q(B') applied in place to the
// `nsched` frequencies listed in d_sched (the whole schedule, or one rank's share in tau-slab mode); skip[rhs].done skips a right-hand side.
void kpm_cheb_apply(sq_kpm *k, double2 *z, const int *d_sched, int nsched, int nrhs, size_t rhs_stride, const CgState *skip) {
    sq_fdm *f = k->f;
    if (nsched <= 0) return;
    if (kpm_reg_ok(k)) {                                                 // register engine (kpm_reg.cu): lattices of the register path
        kpm_cheb_reg_launch(k, z, d_sched, nsched, nrhs, rhs_stride, skip);
        return;
    }
    for (int rhs = 0; rhs < nrhs; rhs++) {
        double2 *zt = z + (size_t)rhs * rhs_stride;
        const CgState *sk = skip ? skip + rhs : nullptr;
        if (kpm_fast_ok(k)) {
            f->stats[SQ_STAT_KPM_SMEM]++;
            BbarFast Q;
            Q.N = (int)f->N; Q.C = (int)f->C; Q.nunc0 = f->nunc0;
            for (int c = 0; c < 8; c++) { Q.clo[c] = c < f->C ? f->clo[c] : 0; Q.chi[c] = c < f->C ? f->chi[c] : 0; }
            Q.nts = f->nts.p; Q.nt = f->nt.p; Q.slot = f->slot.p; Q.unc0 = f->unc0.p; Q.csbar = k->csbar.p; Q.Dbar = k->Dbar.p;
            static long long *dbg = nullptr;
            if (!dbg && getenv("SQ_DEBUG_STAMPS")) { SQ_CUDA(cudaMallocManaged((void **)&dbg, 8 * sizeof(long long))); for (int q = 0; q < 8; q++) dbg[q] = 0; }
            Q.dbg = dbg;
            if (dbg && getenv("SQ_DEBUG_PRINT")) {
                cudaStreamSynchronize(f->stream);
                fprintf(stderr, "cheb stamps: first apply %lld cycles, whole recurrence %lld cycles, order %lld -> %.0f cycles/step\n", dbg[1] - dbg[0],
                        dbg[2] - dbg[0], dbg[3], (double)(dbg[2] - dbg[0]) / (double)std::max<long long>(1, dbg[3] - 1));
            }
            double avg = 0.5 * (k->bounds[1] + k->bounds[0]), mag = 0.5 * (k->bounds[1] - k->bounds[0]);
            size_t smem = f->N * sizeof(double2);
            const int T = kpm_fast_threads(k);
            smem = f->N * sizeof(double);
#define SQ_CHEB(CM, MT) SQ_CUDA(sq_launch(k_kpm_cheb_fast<CM, MT>, dim3(2 * nsched), dim3(T), smem, f->stream, Q, zt, (const int *)d_sched, \
                                          (const int *)k->d_order.p, (const int *)k->d_coef_off.p, (const double2 *)k->d_coefs.p, (int)f->L, avg, 1.0 / mag, sk))
            if (f->C <= 4) { if (T <= 512) SQ_CHEB(4, 512); else SQ_CHEB(4, 1024); }
            else { if (T <= 512) SQ_CHEB(8, 512); else SQ_CHEB(8, 1024); }
#undef SQ_CHEB
            SQ_LAUNCH_CHECK();
            f->launches++;
        } else {
            f->stats[SQ_STAT_KPM_SMEM]++;
            BbarParams P = bbar_params(k);
            double avg = 0.5 * (k->bounds[1] + k->bounds[0]), mag = 0.5 * (k->bounds[1] - k->bounds[0]);
            SQ_CUDA(sq_launch(k_kpm_cheb, dim3(nsched), dim3(kpm_threads(k)), 4 * f->N * sizeof(double2), f->stream, P, zt, (const int *)d_sched,
                              (const int *)k->d_order.p, (const int *)k->d_coef_off.p, (const double2 *)k->d_coefs.p, (int)f->L, avg, 1.0 / mag, sk));
            SQ_LAUNCH_CHECK();
            f->launches++;
        }
    }
}

// out_j = P^-1 in_j for nrhs vectors `stride` elements apart (active preconditioner): forward tau-FFT (order-1 frequencies folded in),
// Chebyshev recurrences, inverse tau-FFT with the partials of conj(dot_with_j).out_j (2 SQ_MAXPART doubles per vector) if requested.
// zt: frequency-space scratch of the same shape.  skip[j].done skips vector j.  *npart = partials per vector.
// upd != NULL: the CG x / r update of this iteration rides in the forward transform's load phase (FftCgUpdate, sq_internal.h); `in` is
// then upd->r and `skip` the state array the update WRITES (upd->nxt), which the later stages read.
void kpm_fft_cheb_batch(sq_kpm *k, double2 *out, const double2 *in, double2 *zt, int nrhs, size_t stride, const CgState *skip,
                        const double2 *dot_with, double *dot_part, int *npart, const FftCgUpdate *upd) {
    sq_fdm *f = k->f;
    tau_fft_launch_batch(f->stream, k->radices, (int)f->L, (int)f->N, zt, in, false, true, k->tw.p, k->theta.p, k->d_scale1.p, nullptr, nullptr,
                         skip, f->smem_optin, nrhs, stride, upd);
    kpm_cheb_apply(k, zt, k->d_freq_sched.p, k->nsched, nrhs, stride, skip);
    const int g = tau_fft_launch_batch(f->stream, k->radices, (int)f->L, (int)f->N, out, zt, true, true, k->tw.p, k->theta.p, nullptr, dot_with,
                                       dot_part, skip, f->smem_optin, nrhs, stride);
    f->launches += 2;
    if (npart) *npart = g;
}

// out = P^-1 in  on device [l][i] vectors (out may alias in).  dot partial fusion is requested by cg.cu.
int kpm_ldiv_dev_dot(sq_kpm *k, double2 *out, const double2 *in, const CgState *skip, const double2 *dot_with, double *dot_part) {
    sq_fdm *f = k->f;
    size_t n = (size_t)f->L * f->N;
    if (!k->active) {                                                    // :407-411
        if (out != in) SQ_CUDA(cudaMemcpyAsync(out, in, n * sizeof(double2), cudaMemcpyDeviceToDevice, f->stream));
        return 0;
    }
    int g = 0;
    kpm_fft_cheb_batch(k, out, in, k->ztmp.p, 1, 0, skip, dot_with, dot_part, &g, nullptr);
    return g;
}
// z = P^-1 r with the CG x / r update fused into the forward transform (active preconditioner only); returns the number of r.z partials
int kpm_ldiv_dev_fused(sq_kpm *k, double2 *z, const FftCgUpdate &upd, double *dot_part) {
    int g = 0;
    kpm_fft_cheb_batch(k, z, upd.r, k->ztmp.p, 1, 0, upd.nxt, upd.r, dot_part, &g, &upd);
    return g;
}
void kpm_ldiv_dev(sq_kpm *k, double2 *out, const double2 *in, const CgState *skip) { kpm_ldiv_dev_dot(k, out, in, skip, nullptr, nullptr); }

void kpm_fourier_dev(sq_kpm *k, double2 *v, bool forward) {
    sq_fdm *f = k->f;
    tau_fft_launch(f->stream, k->radices, (int)f->L, (int)f->N, k->ztmp.p, v, !forward, true, k->tw.p, k->theta.p, nullptr, nullptr, nullptr,
                   nullptr, f->smem_optin);
    SQ_CUDA(cudaMemcpyAsync(v, k->ztmp.p, f->vec_bytes(), cudaMemcpyDeviceToDevice, f->stream));
    f->launches++;
}

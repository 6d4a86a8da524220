/*
 * oracle/ref_c.c -- CPU restatement of the SmoQyElPhQMC.jl hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under smoqyelphqmc.jl_b200/ links, imports or calls
 * this file; it is the checker for tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py.
 *
 * PARITY STATUS: "parity unpinned".  The reference ships no golden vectors, known-answer
 * tests or fixtures for this path and cannot be executed in the build
 * container (no julia, five un-vendored dependencies).  The restatement is pinned instead by
 * the self-owned known-answer tests in tests/test_oracle_kat.py (dense-matrix builders,
 * inverse identities, finite differences, the tau-independent P^-1 M^T M = I property).
 *
 * Layout: identical to the reference (Julia column-major).  Space-time vectors are (Ltau x N)
 * with tau fastest: element (l,i) at l + i*Ltau.  V is (N x Ltau), t is (Nh x Ltau), x is
 * (Nph x Ltau) with the site/bond/phonon index fastest.  All indices 0-based in here; the
 * Python wrapper converts from the 1-based tables the Julia side would pass.
 *
 * Pass structure: deliberately the reference's UN-FUSED structure (one full sweep over memory
 * per array operation) so that timing this file is a fair stand-in for the reference's CPU
 * cost.  OpenMP (optional, -fopenmp) splits the tau axis between threads inside each sweep;
 * the reference itself is single threaded.
 *
 * Each function cites the reference file:line it restates (paths relative to /root/reference).
 * Arithmetic that lives in un-vendored Julia dependencies (SmoQyKPMCore, JDQMCFramework,
 * SmoQyDQMC EFA) is restated from the published algorithms and marked [unvendored].
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double _Complex cplx;
typedef int64_t i64;

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------ */
/* tau-range helper: split [0,L) between OpenMP threads                                        */
/* ------------------------------------------------------------------------------------------ */
static inline void tau_range(i64 L, i64 *lo, i64 *hi) {
#ifdef _OPENMP
    int nt = omp_get_num_threads(), id = omp_get_thread_num();
    i64 chunk = (L + nt - 1) / nt;
    *lo = id * chunk; if (*lo > L) *lo = L;
    *hi = *lo + chunk; if (*hi > L) *hi = L;
#else
    *lo = 0; *hi = L;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the timed CPU arm asks for the cores it may use explicitly */
void ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* FermionDetMatrix  (src/FermionDetMatrix.jl:44-55,137-148)                                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct ref_fdm {
    int sym;                 /* 1 = SymFermionDetMatrix, 0 = AsymFermionDetMatrix */
    i64 L, N, Nh, C;
    i64 *nt;                 /* 2*Nh, checkerboard order, 0-based: (nt[2h], nt[2h+1]) */
    i64 *perm;               /* Nh: checkerboard index -> original hopping index */
    i64 *clo, *chi;          /* colour ranges [clo,chi) in checkerboard order */
    double *expV;            /* (L x N) tau-fastest */
    double *ch, *sh;         /* (L x Nh) tau-fastest */
    double tol; i64 maxiter;
    cplx *tmp1, *tmp2;       /* (L x N) */
    cplx *r, *p, *z;         /* CG workspace, src/IterativeSolvers/ConjugateGradient.jl:16-60 */
} ref_fdm;

ref_fdm *ref_fdm_create(int sym, i64 L, i64 N, i64 Nh, const i64 *nt0, const i64 *perm0, i64 C,
                        const i64 *clo0, const i64 *chi0, double tol, i64 maxiter) {
    ref_fdm *f = (ref_fdm *)calloc(1, sizeof(ref_fdm));
    f->sym = sym; f->L = L; f->N = N; f->Nh = Nh; f->C = C; f->tol = tol; f->maxiter = maxiter;
    f->nt = (i64 *)malloc(sizeof(i64) * 2 * (Nh + 1));
    f->perm = (i64 *)malloc(sizeof(i64) * (Nh + 1));
    f->clo = (i64 *)malloc(sizeof(i64) * (C + 1));
    f->chi = (i64 *)malloc(sizeof(i64) * (C + 1));
    memcpy(f->nt, nt0, sizeof(i64) * 2 * Nh);
    memcpy(f->perm, perm0, sizeof(i64) * Nh);
    memcpy(f->clo, clo0, sizeof(i64) * C);
    memcpy(f->chi, chi0, sizeof(i64) * C);
    f->expV = (double *)calloc(L * N, sizeof(double));
    f->ch = (double *)calloc(L * (Nh + 1), sizeof(double));
    f->sh = (double *)calloc(L * (Nh + 1), sizeof(double));
    f->tmp1 = (cplx *)calloc(L * N, sizeof(cplx));
    f->tmp2 = (cplx *)calloc(L * N, sizeof(cplx));
    f->r = (cplx *)calloc(L * N, sizeof(cplx));
    f->p = (cplx *)calloc(L * N, sizeof(cplx));
    f->z = (cplx *)calloc(L * N, sizeof(cplx));
    return f;
}

void ref_fdm_destroy(ref_fdm *f) {
    if (!f) return;
    free(f->nt); free(f->perm); free(f->clo); free(f->chi); free(f->expV); free(f->ch); free(f->sh);
    free(f->tmp1); free(f->tmp2); free(f->r); free(f->p); free(f->z); free(f);
}

/* update!(fdm, fpi): src/FermionDetMatrix.jl:208-236.  V is (N x L), t is (Nh x L), real. */
void ref_fdm_update(ref_fdm *f, const double *V, const double *t, double dtau) {
    i64 L = f->L, N = f->N, Nh = f->Nh;
    for (i64 i = 0; i < N; i++)
        for (i64 l = 0; l < L; l++) f->expV[l + i * L] = exp(-dtau * V[i + l * N]);
    double dtp = f->sym ? dtau / 2 : dtau;                         /* :220 */
    for (i64 h = 0; h < Nh; h++) {
        i64 hp = f->perm[h];                                       /* :224 */
        for (i64 l = 0; l < L; l++) {
            double tp = t[hp + l * Nh];
            double a = dtp * fabs(tp);
            double sg = (tp > 0) - (tp < 0);                        /* sign(conj(t)) for real t */
            f->ch[l + h * L] = cosh(a);
            f->sh[l + h * L] = sg * sinh(a);
        }
    }
}

double *ref_fdm_expV(ref_fdm *f) { return f->expV; }
double *ref_fdm_cosh(ref_fdm *f) { return f->ch; }
double *ref_fdm_sinh(ref_fdm *f) { return f->sh; }

/* checkerboard_lmul!: src/checkerboard_matrix_multiply.jl:26-72.  Bonds [lo,hi) visited in
 * increasing order, or decreasing if transposed (:44-47).  inverse=1 gives checkerboard_ldiv!
 * (:99-145): factors with s -> -s, and the order is reversed iff !transposed (:118-120). */
static void chk_apply(cplx *u, const ref_fdm *f, const double *ch, const double *sh, i64 L,
                      int transposed, int inverse, i64 lo, i64 hi) {
    int reversed = inverse ? !transposed : transposed;
    double sgn = inverse ? -1.0 : 1.0;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        i64 l0, l1; tau_range(L, &l0, &l1);
        for (i64 k = 0; k < hi - lo; k++) {
            i64 h = reversed ? (hi - 1 - k) : (lo + k);
            cplx *ui = u + f->nt[2 * h] * L, *uj = u + f->nt[2 * h + 1] * L;
            const double *c = ch + h * L, *s = sh + h * L;
            for (i64 l = l0; l < l1; l++) {
                cplx a = ui[l], b = uj[l];
                double cc = c[l], ss = sgn * s[l];
                ui[l] = cc * a + ss * b;
                uj[l] = cc * b + ss * a;                            /* conj(s) = s for real s */
            }
        }
    }
}

void ref_chk_lmul(ref_fdm *f, cplx *u, int transposed, i64 lo, i64 hi) {
    chk_apply(u, f, f->ch, f->sh, f->L, transposed, 0, lo, hi);
}
void ref_chk_ldiv(ref_fdm *f, cplx *u, int transposed, i64 lo, i64 hi) {
    chk_apply(u, f, f->ch, f->sh, f->L, transposed, 1, lo, hi);
}

static void vec_scale_real(cplx *u, const double *d, i64 n, int inverse) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (i64 k = 0; k < n; k++) u[k] = inverse ? u[k] / d[k] : d[k] * u[k];
}

/* circshift!(u', u, (1,0)): u'[l] = u[l-1], u'[0] = u[L-1]  (FermionDetMatrix.jl:398) */
static void circshift1(cplx *o, const cplx *u, i64 L, i64 N) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (i64 i = 0; i < N; i++) {
        const cplx *a = u + i * L; cplx *b = o + i * L;
        b[0] = a[L - 1];
        for (i64 l = 1; l < L; l++) b[l] = a[l - 1];
    }
}

/* mul_M!: Sym src/FermionDetMatrix.jl:385-427, Asym :430-466.  out must not alias in. */
void ref_mul_M(ref_fdm *f, cplx *out, const cplx *in) {
    i64 L = f->L, N = f->N, Nh = f->Nh;
    circshift1(out, in, L, N);
    if (f->sym) {
        chk_apply(out, f, f->ch, f->sh, L, 1, 0, 0, Nh);          /* :401 */
        vec_scale_real(out, f->expV, L * N, 0);                     /* :407 */
        chk_apply(out, f, f->ch, f->sh, L, 0, 0, 0, Nh);          /* :410 */
    } else {
        chk_apply(out, f, f->ch, f->sh, L, 0, 0, 0, Nh);          /* :445 */
        vec_scale_real(out, f->expV, L * N, 0);                     /* :451 */
    }
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (i64 i = 0; i < N; i++) {                                   /* :416-424 */
        cplx *o = out + i * L; const cplx *u = in + i * L;
        o[0] = u[0] + o[0];
        for (i64 l = 1; l < L; l++) o[l] = u[l] - o[l];
    }
}

/* mul_Mt!: Sym src/FermionDetMatrix.jl:484-525, Asym :528-563.  out must not alias in. */
void ref_mul_Mt(ref_fdm *f, cplx *out, const cplx *in) {
    i64 L = f->L, N = f->N, Nh = f->Nh;
    if (f->sym) {
        memcpy(out, in, sizeof(cplx) * L * N);                     /* checkerboard_mul! copy */
        chk_apply(out, f, f->ch, f->sh, L, 1, 0, 0, Nh);          /* :497 */
        vec_scale_real(out, f->expV, L * N, 0);                     /* :503 */
        chk_apply(out, f, f->ch, f->sh, L, 0, 0, 0, Nh);          /* :506 */
    } else {
#ifdef _OPENMP
#pragma omp parallel for
#endif
        for (i64 k = 0; k < L * N; k++) out[k] = f->expV[k] * in[k]; /* :541 */
        chk_apply(out, f, f->ch, f->sh, L, 1, 0, 0, Nh);          /* :544 */
    }
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (i64 i = 0; i < N; i++) {                                   /* :512-522 */
        cplx *o = out + i * L; const cplx *u = in + i * L;
        cplx last = u[L - 1] + o[0];
        for (i64 l = 0; l < L - 1; l++) o[l] = u[l] - o[l + 1];
        o[L - 1] = last;
    }
}

/* mul_MtM!: src/FermionDetMatrix.jl:329-340 (tmp1 = M v; v' = Mt tmp1); in-place allowed. */
void ref_mul_MtM(ref_fdm *f, cplx *out, const cplx *in) {
    ref_mul_M(f, f->tmp1, in);
    ref_mul_Mt(f, out, f->tmp1);
}
/* mul_MMt!: src/FermionDetMatrix.jl:357-368 */
void ref_mul_MMt(ref_fdm *f, cplx *out, const cplx *in) {
    ref_mul_Mt(f, f->tmp1, in);
    ref_mul_M(f, out, f->tmp1);
}

/* BLAS-1 helpers (LinearAlgebra dot/norm/axpy!/axpby!) */
static cplx zdot(const cplx *a, const cplx *b, i64 n) {           /* sum conj(a) b */
    double re = 0, im = 0;
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : re, im)
#endif
    for (i64 k = 0; k < n; k++) { cplx v = conj(a[k]) * b[k]; re += creal(v); im += cimag(v); }
    return re + im * I;
}
static double znorm(const cplx *a, i64 n) { return sqrt(creal(zdot(a, a, n))); }
static void zaxpy(cplx al, const cplx *x, cplx *y, i64 n) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (i64 k = 0; k < n; k++) y[k] += al * x[k];
}
static void zaxpby(cplx al, const cplx *x, cplx be, cplx *y, i64 n) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (i64 k = 0; k < n; k++) y[k] = al * x[k] + be * y[k];
}

/* ------------------------------------------------------------------------------------------ */
/* Mixed radix FFT along tau (stands in for FFTW; src/FourierTransformer.jl:17-18)             */
/* ------------------------------------------------------------------------------------------ */
typedef struct { i64 n; cplx *tw; cplx *scratch; } ref_fft;

static void fft_rec(const cplx *in, cplx *out, i64 n, i64 stride, const cplx *tw, i64 twstride,
                    i64 ntot) {
    if (n == 1) { out[0] = in[0]; return; }
    i64 r = 2;
    if (n % 4 == 0) r = 4;
    else { while (n % r) r++; }
    i64 m = n / r;
    for (i64 q = 0; q < r; q++) fft_rec(in + q * stride, out + q * m, m, stride * r, tw, twstride * r, ntot);
    cplx tmp[64];
    if (r > 64) { fprintf(stderr, "ref fft: prime factor %lld too large\n", (long long)r); abort(); }
    for (i64 k = 0; k < m; k++) {
        for (i64 q = 0; q < r; q++) tmp[q] = out[k + q * m] * tw[(q * k * twstride) % ntot];
        for (i64 j = 0; j < r; j++) {
            cplx acc = tmp[0];
            for (i64 q = 1; q < r; q++) acc += tmp[q] * tw[(q * j * m * twstride) % ntot];
            out[k + j * m] = acc;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* KPMPreconditioner (src/KPMPreconditioner.jl) + FourierTransformer (src/FourierTransformer.jl)*/
/* ------------------------------------------------------------------------------------------ */
typedef struct ref_kpm {
    ref_fdm *f;
    int active;
    double rbuf, a1, a2; i64 nlanczos;
    double *Dbar, *cbar, *sbar;   /* tau-means: N, Nh, Nh  (:604-621) */
    double bounds[2];
    i64 ncoef;                    /* cld(L,2) for Sym, L for Asym */
    i64 *order;
    cplx **coefs;                 /* real for Sym (imag = 0) */
    cplx *theta, *twf, *twb;      /* theta_l (FourierTransformer.jl:15), forward/backward twiddles */
    cplx *v, *vp;                 /* (L x N) and (N x L) scratch (:214) */
} ref_kpm;

static void kpm_update_Bbar(ref_kpm *k) {          /* update_B̄!: KPMPreconditioner.jl:604-621 */
    ref_fdm *f = k->f; i64 L = f->L;
    for (i64 i = 0; i < f->N; i++) { double s = 0; for (i64 l = 0; l < L; l++) s += f->expV[l + i * L]; k->Dbar[i] = s / L; }
    for (i64 h = 0; h < f->Nh; h++) {
        double c = 0, s = 0;
        for (i64 l = 0; l < L; l++) { c += f->ch[l + h * L]; s += f->sh[l + h * L]; }
        k->cbar[h] = c / L; k->sbar[h] = s / L;
    }
}

/* B̄ v on an N-vector.  [unvendored: JDQMCFramework Sym/AsymChkbrdPropagator mul!]
 * Sym: Gamma_bar D_bar Gamma_bar^T ; Asym: D_bar Gamma_bar  (KPMPreconditioner.jl:260,274) */
static void bbar_chk(const ref_kpm *k, cplx *v, int transposed) {
    const ref_fdm *f = k->f;
    for (i64 q = 0; q < f->Nh; q++) {
        i64 h = transposed ? f->Nh - 1 - q : q;
        i64 i = f->nt[2 * h], j = f->nt[2 * h + 1];
        cplx a = v[i], b = v[j];
        v[i] = k->cbar[h] * a + k->sbar[h] * b;
        v[j] = k->cbar[h] * b + k->sbar[h] * a;
    }
}
static void bbar_mul(const ref_kpm *k, cplx *out, const cplx *in) {
    const ref_fdm *f = k->f; i64 N = f->N;
    if (out != in) memcpy(out, in, sizeof(cplx) * N);
    if (f->sym) {
        bbar_chk(k, out, 1);
        for (i64 i = 0; i < N; i++) out[i] *= k->Dbar[i];
        bbar_chk(k, out, 0);
    } else {
        bbar_chk(k, out, 0);
        for (i64 i = 0; i < N; i++) out[i] *= k->Dbar[i];
    }
}
/* mul_B̄ᵀB̄!: KPMPreconditioner.jl:661-679 (Asym only) */
static void bbar_mul_BtB(const ref_kpm *k, cplx *out, const cplx *in) {
    i64 N = k->f->N;
    if (out != in) memcpy(out, in, sizeof(cplx) * N);
    bbar_chk(k, out, 0);
    for (i64 i = 0; i < N; i++) out[i] *= k->Dbar[i] * k->Dbar[i];
    bbar_chk(k, out, 1);
}

/* eigenvalue extremes of a symmetric tridiagonal (diag a[n], offdiag b[n-1]) by bisection */
static i64 sturm_count(const double *a, const double *b, i64 n, double x) {
    i64 cnt = 0; double d = 1.0;
    for (i64 i = 0; i < n; i++) {
        double bb = (i == 0) ? 0.0 : b[i - 1] * b[i - 1];
        d = a[i] - x - (d != 0.0 ? bb / d : bb / 1e-300);
        if (d < 0) cnt++;
    }
    return cnt;
}
void ref_tridiag_extremes(const double *a, const double *b, i64 n, double *emin, double *emax) {
    double lo = a[0], hi = a[0];
    for (i64 i = 0; i < n; i++) {
        double r = (i > 0 ? fabs(b[i - 1]) : 0) + (i < n - 1 ? fabs(b[i]) : 0);
        if (a[i] - r < lo) lo = a[i] - r;
        if (a[i] + r > hi) hi = a[i] + r;
    }
    double l = lo, h = hi;                     /* smallest: count(x) >= 1 */
    for (int it = 0; it < 200; it++) { double m = 0.5 * (l + h); if (sturm_count(a, b, n, m) >= 1) h = m; else l = m; }
    *emin = 0.5 * (l + h);
    l = lo; h = hi;                             /* largest: count(x) >= n */
    for (int it = 0; it < 200; it++) { double m = 0.5 * (l + h); if (sturm_count(a, b, n, m) >= n) h = m; else l = m; }
    *emax = 0.5 * (l + h);
}

/* calculate_bounds!: KPMPreconditioner.jl:625-658.  start = the randn! vector (N reals).
 * [unvendored: SmoQyKPMCore.lanczos!] plain Lanczos, no re-orthogonalisation. */
static void kpm_lanczos_bounds(ref_kpm *k, const double *start, double *emin, double *emax) {
    ref_fdm *f = k->f; i64 N = f->N, n = k->nlanczos;
    cplx *vprev = (cplx *)calloc(N, sizeof(cplx)), *v = (cplx *)calloc(N, sizeof(cplx)), *w = (cplx *)calloc(N, sizeof(cplx));
    double *al = (double *)calloc(n, sizeof(double)), *be = (double *)calloc(n, sizeof(double));
    double nrm = 0; for (i64 i = 0; i < N; i++) nrm += start[i] * start[i];
    nrm = sqrt(nrm);
    for (i64 i = 0; i < N; i++) v[i] = start[i] / nrm;
    double bprev = 0; i64 used = n;
    for (i64 j = 0; j < n; j++) {
        if (f->sym) bbar_mul(k, w, v); else bbar_mul_BtB(k, w, v);
        double a = 0; for (i64 i = 0; i < N; i++) a += creal(v[i]) * creal(w[i]);
        al[j] = a;
        for (i64 i = 0; i < N; i++) w[i] = w[i] - a * v[i] - bprev * vprev[i];
        double b = 0; for (i64 i = 0; i < N; i++) b += creal(w[i]) * creal(w[i]);
        b = sqrt(b);
        if (j < n - 1) {
            be[j] = b;
            if (b < 1e-300) { used = j + 1; break; }
            for (i64 i = 0; i < N; i++) { vprev[i] = v[i]; v[i] = w[i] / b; }
            bprev = b;
        }
    }
    ref_tridiag_extremes(al, be, used, emin, emax);
    if (!f->sym) { *emin = sqrt(*emin); *emax = sqrt(*emax); }   /* :655 */
    free(vprev); free(v); free(w); free(al); free(be);
}

static cplx f_sym(double b, double phi) { return 1.0 / (b * b - 2 * b * cos(phi) + 1); }   /* :800 */
static cplx f_asym(double b, double phi) { return 1.0 / (1.0 - cexp(-I * phi) * b); }        /* :804 */

/* kpm_coefs!  [unvendored: SmoQyKPMCore] Chebyshev-Gauss quadrature with Nq = 2*order nodes,
 * c_k = (2 - delta_k0)/Nq * sum_j f(x_j) cos(k pi (j+1/2)/Nq), x_j mapped onto the bounds. */
static void kpm_coefs(cplx *c, i64 order, int sym, double phi, const double bounds[2]) {
    i64 Nq = 2 * order;
    double avg = 0.5 * (bounds[1] + bounds[0]), mag = 0.5 * (bounds[1] - bounds[0]);
    cplx *fx = (cplx *)malloc(sizeof(cplx) * Nq);
    for (i64 j = 0; j < Nq; j++) {
        double xj = mag * cos(M_PI * (j + 0.5) / Nq) + avg;
        fx[j] = sym ? f_sym(xj, phi) : f_asym(xj, phi);
    }
    for (i64 q = 0; q < order; q++) {
        cplx s = 0;
        for (i64 j = 0; j < Nq; j++) s += fx[j] * cos(M_PI * q * (j + 0.5) / Nq);
        c[q] = (q == 0 ? 1.0 : 2.0) * s / (double)Nq;
    }
    free(fx);
}

/* update_kpm_expansion_order!/coefs!: KPMPreconditioner.jl:696-795 */
static void kpm_update_expansions(ref_kpm *k) {
    ref_fdm *f = k->f; i64 L = f->L;
    double emin = k->bounds[0], emax = k->bounds[1];
    for (i64 l = 0; l < k->ncoef; l++) {
        double phi = 2 * M_PI / L * (l + 0.5);                         /* :220 */
        double ph = phi > M_PI ? 2 * M_PI - phi : phi;                 /* :710 */
        i64 n = (i64)floor((emax - emin) * (k->a1 / ph + k->a2));     /* :711 */
        if (n < 1) n = 1;
        k->order[l] = n;
        free(k->coefs[l]);
        k->coefs[l] = (cplx *)calloc(n, sizeof(cplx));
    }
    if (f->sym) {
        for (i64 l = 0; l < k->ncoef; l++) kpm_coefs(k->coefs[l], k->order[l], 1, 2 * M_PI / L * (l + 0.5), k->bounds);
    } else {
        i64 Lo2 = (L + 1) / 2;
        for (i64 l = 0; l < Lo2; l++) {                                 /* :771-792 */
            kpm_coefs(k->coefs[l], k->order[l], 0, 2 * M_PI / L * (l + 0.5), k->bounds);
            i64 m = L - 1 - l;
            for (i64 q = 0; q < k->order[l] && q < k->order[m]; q++) k->coefs[m][q] = conj(k->coefs[l][q]);
        }
    }
}

ref_kpm *ref_kpm_create(ref_fdm *f, double rbuf, i64 nlanczos, double a1, double a2) {
    ref_kpm *k = (ref_kpm *)calloc(1, sizeof(ref_kpm));
    i64 L = f->L, N = f->N;
    k->f = f; k->rbuf = rbuf; k->nlanczos = nlanczos; k->a2 = a2;
    k->a1 = f->sym ? 2 * a1 : a1;                                       /* :263 */
    k->Dbar = (double *)calloc(N, sizeof(double));
    k->cbar = (double *)calloc(f->Nh + 1, sizeof(double));
    k->sbar = (double *)calloc(f->Nh + 1, sizeof(double));
    k->ncoef = f->sym ? (L + 1) / 2 : L;
    k->order = (i64 *)calloc(k->ncoef, sizeof(i64));
    k->coefs = (cplx **)calloc(k->ncoef, sizeof(cplx *));
    k->theta = (cplx *)malloc(sizeof(cplx) * L);
    k->twf = (cplx *)malloc(sizeof(cplx) * L);
    k->twb = (cplx *)malloc(sizeof(cplx) * L);
    for (i64 l = 0; l < L; l++) {
        k->theta[l] = cexp(-I * M_PI * (double)l / (double)L);          /* FourierTransformer.jl:15 */
        k->twf[l] = cexp(-2 * I * M_PI * (double)l / (double)L);
        k->twb[l] = cexp(+2 * I * M_PI * (double)l / (double)L);
    }
    k->v = (cplx *)calloc(L * N, sizeof(cplx));
    k->vp = (cplx *)calloc(L * N, sizeof(cplx));
    return k;
}
void ref_kpm_destroy(ref_kpm *k) {
    if (!k) return;
    for (i64 l = 0; l < k->ncoef; l++) free(k->coefs[l]);
    free(k->Dbar); free(k->cbar); free(k->sbar); free(k->order); free(k->coefs);
    free(k->theta); free(k->twf); free(k->twb); free(k->v); free(k->vp); free(k);
}

/* update_preconditioner!: KPMPreconditioner.jl:554-597.  start = N normals for Lanczos. */
void ref_kpm_update(ref_kpm *k, const double *start) {
    kpm_update_Bbar(k);
    double emin, emax;
    kpm_lanczos_bounds(k, start, &emin, &emax);
    emin *= (1 - k->rbuf); emax *= (1 + k->rbuf);                         /* :566-567 */
    if (0.0 < emin && emin < 1.0 && 1.0 < emax && emax < 2.0) {           /* :570 */
        k->active = 1;
        double e0 = k->bounds[0], e1 = k->bounds[1];
        if (fabs((emin - e0) / e0) > k->rbuf / 2 || fabs((emax - e1) / e1) > k->rbuf / 2 || e0 == 0.0) { /* :582 */
            k->bounds[0] = emin; k->bounds[1] = emax;
            kpm_update_expansions(k);
        }
    } else k->active = 0;
}
/* test hook: refresh B̄ but inject the bounds (so GPU and oracle use identical coefficients) */
void ref_kpm_set_bounds(ref_kpm *k, double emin, double emax) {
    kpm_update_Bbar(k);
    k->bounds[0] = emin; k->bounds[1] = emax; k->active = 1;
    kpm_update_expansions(k);
}
void ref_kpm_refresh_Bbar(ref_kpm *k) { kpm_update_Bbar(k); }
int ref_kpm_active(ref_kpm *k) { return k->active; }
void ref_kpm_get_bounds(ref_kpm *k, double *b) { b[0] = k->bounds[0]; b[1] = k->bounds[1]; }
i64 ref_kpm_ncoef(ref_kpm *k) { return k->ncoef; }
void ref_kpm_get_orders(ref_kpm *k, i64 *o) { memcpy(o, k->order, sizeof(i64) * k->ncoef); }
void ref_kpm_get_coefs(ref_kpm *k, i64 l, cplx *c) { memcpy(c, k->coefs[l], sizeof(cplx) * k->order[l]); }
void ref_kpm_lanczos(ref_kpm *k, const double *start, double *b) { kpm_update_Bbar(k); kpm_lanczos_bounds(k, start, b, b + 1); }
void ref_kpm_bbar_mul(ref_kpm *k, cplx *out, const cplx *in) { bbar_mul(k, out, in); }

/* kpm_lmul!  [unvendored: SmoQyKPMCore] v <- sum_q c_q T_q(B') v, B' = (B̄ - avg)/mag */
static void kpm_lmul(const ref_kpm *k, const cplx *c, i64 order, cplx *v, cplx *tmp) {
    i64 N = k->f->N;
    double avg = 0.5 * (k->bounds[1] + k->bounds[0]), mag = 0.5 * (k->bounds[1] - k->bounds[0]);
    cplx *a1 = tmp, *a2 = tmp + N, *a3 = tmp + 2 * N;
    memcpy(a1, v, sizeof(cplx) * N);                                   /* T0 v */
    bbar_mul(k, a2, v);
    for (i64 i = 0; i < N; i++) a2[i] = (a2[i] - avg * a1[i]) / mag;    /* T1 v */
    for (i64 i = 0; i < N; i++) v[i] = c[0] * a1[i] + c[1] * a2[i];
    for (i64 q = 2; q < order; q++) {
        bbar_mul(k, a3, a2);
        for (i64 i = 0; i < N; i++) a3[i] = 2.0 * (a3[i] - avg * a2[i]) / mag - a1[i];
        for (i64 i = 0; i < N; i++) v[i] += c[q] * a3[i];
        cplx *t = a1; a1 = a2; a2 = a3; a3 = t;
    }
}

/* U v (forward = 1): FourierTransformer.jl:39-50 ; U^-1 v: :53-64.  v is (L x N) tau-fastest. */
void ref_fourier(ref_kpm *k, cplx *v, int forward) {
    i64 L = k->f->L, N = k->f->N;
    double sq = sqrt((double)L);
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        cplx *buf = (cplx *)malloc(sizeof(cplx) * L);
#ifdef _OPENMP
#pragma omp for
#endif
        for (i64 i = 0; i < N; i++) {
            cplx *u = v + i * L;
            if (forward) {
                for (i64 l = 0; l < L; l++) buf[l] = k->theta[l] / sq * u[l];
                fft_rec(buf, u, L, 1, k->twf, 1, L);
            } else {
                fft_rec(u, buf, L, 1, k->twb, 1, L);
                for (i64 l = 0; l < L; l++) u[l] = (1.0 / k->theta[l]) * sq * (buf[l] / (double)L);
            }
        }
        free(buf);
    }
}

/* ldiv!(u', P, u) complex: Sym KPMPreconditioner.jl:355-414, Asym :488-550 */
void ref_kpm_ldiv(ref_kpm *k, cplx *out, const cplx *in) {
    ref_fdm *f = k->f; i64 L = f->L, N = f->N;
    if (!k->active) { if (out != in) memcpy(out, in, sizeof(cplx) * L * N); return; }   /* :407-411 */
    memcpy(k->v, in, sizeof(cplx) * L * N);
    ref_fourier(k, k->v, 1);                                                               /* :375 */
    for (i64 i = 0; i < N; i++) for (i64 n = 0; n < L; n++) k->vp[i + n * N] = k->v[n + i * L];  /* :378 */
    i64 Lo2 = (L + 1) / 2;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        cplx *tmp = (cplx *)malloc(sizeof(cplx) * 3 * N);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (i64 n = 0; n < L; n++) {
            cplx *vn = k->vp + n * N;
            if (f->sym) {
                i64 np = (n + 1 > Lo2) ? L - 1 - n : n;                                    /* :387 */
                if (k->order[np] > 1) kpm_lmul(k, k->coefs[np], k->order[np], vn, tmp);     /* :394 */
                else for (i64 i = 0; i < N; i++) vn[i] *= k->coefs[np][0];                  /* :398 */
            } else {
                if (k->order[n] > 1) {
                    kpm_lmul(k, k->coefs[L - 1 - n], k->order[L - 1 - n], vn, tmp);         /* :527 */
                    kpm_lmul(k, k->coefs[n], k->order[n], vn, tmp);                         /* :530 */
                } else {
                    double a = creal(k->coefs[n][0]) * creal(k->coefs[n][0]) + cimag(k->coefs[n][0]) * cimag(k->coefs[n][0]);
                    for (i64 i = 0; i < N; i++) vn[i] *= a;                                 /* :534 */
                }
            }
        }
        free(tmp);
    }
    for (i64 i = 0; i < N; i++) for (i64 n = 0; n < L; n++) out[n + i * L] = k->vp[i + n * N];   /* :403 */
    ref_fourier(k, out, 0);                                                                 /* :406 */
}

/* ------------------------------------------------------------------------------------------ */
/* cg_solve!: src/IterativeSolvers/ConjugateGradient.jl:93-167 (P = I), :169-249 (P given)     */
/* same != 0 <=> x === b (zero start).  Returns iterations, *eps = final |r|/|b|.              */
/* ------------------------------------------------------------------------------------------ */
i64 ref_cg(ref_fdm *f, cplx *x, const cplx *b, int same, ref_kpm *P, double tol, i64 maxiter, double *eps) {
    i64 n = f->L * f->N;
    cplx *r = f->r, *p = f->p, *z = f->z;
    double normb = znorm(b, n);
    if (same) { memcpy(r, b, sizeof(cplx) * n); memset(x, 0, sizeof(cplx) * n); }
    else { ref_mul_MtM(f, r, x); zaxpby(1.0, b, -1.0, r, n); }
    cplx rdotz;
    if (P) { ref_kpm_ldiv(P, z, r); memcpy(p, z, sizeof(cplx) * n); rdotz = zdot(r, z, n); }
    else { memcpy(p, r, sizeof(cplx) * n); rdotz = zdot(r, r, n); }
    double e = znorm(r, n) / normb;
    if (e < tol) { *eps = e; return 0; }
    for (i64 it = 1; it <= maxiter; it++) {
        ref_mul_MtM(f, z, p);
        cplx alpha = rdotz / zdot(p, z, n);
        zaxpy(alpha, p, x, n);
        zaxpy(-alpha, z, r, n);
        e = znorm(r, n) / normb;
        if (e < tol) { *eps = e; return it; }
        cplx nw;
        if (P) { ref_kpm_ldiv(P, z, r); nw = zdot(r, z, n); }
        else nw = zdot(r, r, n);
        cplx beta = nw / rdotz;
        rdotz = nw;
        zaxpby(1.0, P ? z : r, beta, p, n);
    }
    *eps = e;
    return maxiter;
}

/* ------------------------------------------------------------------------------------------ */
/* Electron-phonon model data (SmoQyDQMC ElectronPhononParameters fields used by the path)     */
/* ------------------------------------------------------------------------------------------ */
typedef struct ref_elph {
    i64 L, N, Nh, Nph; double dtau;
    double *x;                    /* (Nph x L), phonon fastest */
    double *Om, *Om4, *M;         /* PhononParameters Ω, Ω4, M */
    i64 Nhol; i64 *hol_ph, *hol_site; double *ha, *ha2, *ha3, *ha4; int *hol_sym;
    i64 Nssh; i64 *ssh_ph; i64 *ssh_hop; double *sa, *sa2, *sa3, *sa4;
    double *V0, *t0;              /* bare on-site energy minus mu (N), bare hopping (Nh, original order) */
    double *V, *t;                /* (N x L), (Nh x L) scratch = FermionPathIntegral.V / .t */
    i64 Ndisp; i64 *disp_ph; double *disp_Om, *disp_Om4;   /* DispersionParameters: phonon pairs (2 x Ndisp), Ω, Ω4 per coupling */
} ref_elph;

static double *dupd(const double *a, i64 n) { double *r = (double *)calloc(n + 1, sizeof(double)); if (n) memcpy(r, a, sizeof(double) * n); return r; }
static i64 *dupi(const i64 *a, i64 n) { i64 *r = (i64 *)calloc(n + 1, sizeof(i64)); if (n) memcpy(r, a, sizeof(i64) * n); return r; }

ref_elph *ref_elph_create(i64 L, i64 N, i64 Nh, i64 Nph, double dtau, const double *Om, const double *Om4, const double *M,
                          i64 Nhol, const i64 *hol_ph, const i64 *hol_site, const double *ha, const double *ha2,
                          const double *ha3, const double *ha4, const int *hol_sym,
                          i64 Nssh, const i64 *ssh_ph, const i64 *ssh_hop, const double *sa, const double *sa2,
                          const double *sa3, const double *sa4, const double *V0, const double *t0) {
    ref_elph *e = (ref_elph *)calloc(1, sizeof(ref_elph));
    e->L = L; e->N = N; e->Nh = Nh; e->Nph = Nph; e->dtau = dtau;
    e->x = (double *)calloc(Nph * L + 1, sizeof(double));
    e->Om = dupd(Om, Nph); e->Om4 = dupd(Om4, Nph); e->M = dupd(M, Nph);
    e->Nhol = Nhol; e->hol_ph = dupi(hol_ph, Nhol); e->hol_site = dupi(hol_site, Nhol);
    e->ha = dupd(ha, Nhol); e->ha2 = dupd(ha2, Nhol); e->ha3 = dupd(ha3, Nhol); e->ha4 = dupd(ha4, Nhol);
    e->hol_sym = (int *)calloc(Nhol + 1, sizeof(int)); if (Nhol) memcpy(e->hol_sym, hol_sym, sizeof(int) * Nhol);
    e->Nssh = Nssh; e->ssh_ph = dupi(ssh_ph, 2 * Nssh); e->ssh_hop = dupi(ssh_hop, Nssh);
    e->sa = dupd(sa, Nssh); e->sa2 = dupd(sa2, Nssh); e->sa3 = dupd(sa3, Nssh); e->sa4 = dupd(sa4, Nssh);
    e->V0 = dupd(V0, N); e->t0 = dupd(t0, Nh);
    e->V = (double *)calloc(N * L + 1, sizeof(double)); e->t = (double *)calloc(Nh * L + 1, sizeof(double));
    return e;
}
void ref_elph_destroy(ref_elph *e) {
    if (!e) return;
    free(e->x); free(e->Om); free(e->Om4); free(e->M); free(e->hol_ph); free(e->hol_site); free(e->ha); free(e->ha2);
    free(e->ha3); free(e->ha4); free(e->hol_sym); free(e->ssh_ph); free(e->ssh_hop); free(e->sa); free(e->sa2);
    free(e->sa3); free(e->sa4); free(e->V0); free(e->t0); free(e->V); free(e->t);
    free(e->disp_ph); free(e->disp_Om); free(e->disp_Om4); free(e);
}
double *ref_elph_x(ref_elph *e) { return e->x; }
double *ref_elph_V(ref_elph *e) { return e->V; }
double *ref_elph_t(ref_elph *e) { return e->t; }
void ref_elph_shift_mu(ref_elph *e, double dmu) { for (i64 i = 0; i < e->N; i++) e->V0[i] -= dmu; }

/* [unvendored: SmoQyDQMC update!(fpi, elph, x, sgn)] rebuild V and t from the phonon field:
 * V[i,l] = V0[i] + sum_c (a x + a2 x^2 + a3 x^3 + a4 x^4);  t[h,l] = t0[h] - sum_c (a dx + ... ),
 * dx = x[p',l] - x[p,l]  (sign convention fixed by fermion_det_matrix_dervative.jl:236-247). */
void ref_elph_build_Vt(ref_elph *e) {
    i64 L = e->L, N = e->N, Nh = e->Nh, Nph = e->Nph;
    for (i64 l = 0; l < L; l++) {
        for (i64 i = 0; i < N; i++) e->V[i + l * N] = e->V0[i];
        for (i64 h = 0; h < Nh; h++) e->t[h + l * Nh] = e->t0[h];
        for (i64 c = 0; c < e->Nhol; c++) {
            double x = e->x[e->hol_ph[c] + l * Nph];
            e->V[e->hol_site[c] + l * N] += e->ha[c] * x + e->ha2[c] * x * x + e->ha3[c] * x * x * x + e->ha4[c] * x * x * x * x;
        }
        for (i64 c = 0; c < e->Nssh; c++) {
            double dx = e->x[e->ssh_ph[2 * c + 1] + l * Nph] - e->x[e->ssh_ph[2 * c] + l * Nph];
            e->t[e->ssh_hop[c] + l * Nh] -= e->sa[c] * dx + e->sa2[c] * dx * dx + e->sa3[c] * dx * dx * dx + e->sa4[c] * dx * dx * dx * dx;
        }
    }
}
void ref_elph_refresh(ref_elph *e, ref_fdm *f) { ref_elph_build_Vt(e); ref_fdm_update(f, e->V, e->t, e->dtau); }

/* ------------------------------------------------------------------------------------------ */
/* Holstein shift matrix Lambda: src/holstein_shift_matrix.jl                                  */
/* ------------------------------------------------------------------------------------------ */
void ref_update_Lambda(double *Lam, const ref_elph *e) {                                 /* :2-44 */
    i64 L = e->L, N = e->N, Nph = e->Nph;
    for (i64 i = 0; i < N; i++) { Lam[i * L] = 1.0; for (i64 l = 1; l < L; l++) Lam[l + i * L] = -1.0; }
    for (i64 c = 0; c < e->Nhol; c++) if (e->hol_sym[c]) {
        double *Li = Lam + e->hol_site[c] * L;
        for (i64 l = 0; l < L; l++) {
            double x = e->x[e->hol_ph[c] + l * Nph];
            Li[l] = exp(+e->dtau * (e->ha[c] * x + e->ha3[c] * x * x * x) / 2) * Li[l];   /* :37 */
        }
    }
}
void ref_mul_Lambda(cplx *o, const double *Lam, const cplx *u, i64 L, i64 N) {           /* :47-71 */
    for (i64 n = 0; n < N; n++) {
        const cplx *v = u + n * L; cplx *w = o + n * L; const double *a = Lam + n * L;
        cplx v0 = v[0];
        for (i64 l = 0; l < L - 1; l++) w[l] = a[l + 1] * v[l + 1];
        w[L - 1] = a[0] * v0;
    }
}
void ref_ldiv_Lambda(cplx *o, const double *Lam, const cplx *u, i64 L, i64 N) {          /* :74-99 */
    for (i64 n = 0; n < N; n++) {
        const cplx *v = u + n * L; cplx *w = o + n * L; const double *a = Lam + n * L;
        cplx vl = v[L - 1];
        for (i64 l = L - 1; l >= 1; l--) w[l] = v[l - 1] / a[l];
        w[0] = vl / a[0];
    }
}
void ref_mul_LambdaT(cplx *o, const double *Lam, const cplx *u, i64 L, i64 N) {          /* :102-126 */
    for (i64 n = 0; n < N; n++) {
        const cplx *v = u + n * L; cplx *w = o + n * L; const double *a = Lam + n * L;
        cplx vl = v[L - 1];
        for (i64 l = L - 1; l >= 1; l--) w[l] = a[l] * v[l - 1];
        w[0] = a[0] * vl;
    }
}
void ref_ldiv_LambdaT(cplx *o, const double *Lam, const cplx *u, i64 L, i64 N) {         /* :129-153 */
    for (i64 n = 0; n < N; n++) {
        const cplx *v = u + n * L; cplx *w = o + n * L; const double *a = Lam + n * L;
        cplx v0 = v[0];
        for (i64 l = 0; l < L - 1; l++) w[l] = v[l + 1] / a[l + 1];
        w[L - 1] = v0 / a[0];
    }
}
/* mul_νRe∂Λ∂x!: :156-201.  F is (Nph x L). */
void ref_mul_nuRe_dLambda_dx(double *F, double nu, const cplx *up, const cplx *u, const double *Lam, const ref_elph *e) {
    i64 L = e->L, Nph = e->Nph;
    for (i64 c = 0; c < e->Nhol; c++) if (e->hol_sym[c]) {
        i64 ph = e->hol_ph[c], site = e->hol_site[c];
        for (i64 l = 0; l < L; l++) {
            double x = e->x[ph + l * Nph];
            double d = e->dtau * (e->ha[c] + 3 * e->ha3[c] * x * x) / 2 * Lam[l + site * L];       /* :192 */
            i64 lm = (l + L - 1) % L;
            F[ph + l * Nph] += nu * creal(conj(up[lm + site * L]) * d * u[l + site * L]);           /* :193 */
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Force: src/fermion_det_matrix_dervative.jl                                                   */
/* ------------------------------------------------------------------------------------------ */
/* _mul_νReΔτ∂Kc∂x!: :196-254 */
static void dKc_dx(double *F, double nu, const cplx *up, const cplx *vp, const ref_fdm *f, const ref_elph *e, double dt, i64 color,
                   const i64 *hop_first, const i64 *hop_next) {
    i64 L = f->L, Nph = e->Nph;
    for (i64 n = f->clo[color]; n < f->chi[color]; n++) {
        i64 h = f->perm[n];
        for (i64 c = hop_first[h]; c >= 0; c = hop_next[c]) {
            i64 p = e->ssh_ph[2 * c], pp = e->ssh_ph[2 * c + 1];
            int fp = isfinite(e->M[p]), fpp = isfinite(e->M[pp]);
            i64 i = f->nt[2 * n], j = f->nt[2 * n + 1];
            for (i64 l = 0; l < L; l++) {
                double dx = e->x[pp + l * Nph] - e->x[p + l * Nph];
                double g = dt * (e->sa[c] + 2 * e->sa2[c] * dx + 3 * e->sa3[c] * dx * dx + 4 * e->sa4[c] * dx * dx * dx);
                double val = nu * creal(conj(up[l + j * L]) * g * vp[l + i * L] + conj(up[l + i * L]) * g * vp[l + j * L]);
                if (fp) F[p + l * Nph] -= val;
                if (fpp) F[pp + l * Nph] += val;
            }
        }
    }
}
/* _mul_νReΔτ∂V∂x!: :258-290 */
static void dV_dx(double *F, double nu, const cplx *up, const cplx *vp, const ref_elph *e) {
    i64 L = e->L, Nph = e->Nph;
    for (i64 c = 0; c < e->Nhol; c++) {
        i64 p = e->hol_ph[c], i = e->hol_site[c];
        if (!isfinite(e->M[p])) continue;
        for (i64 l = 0; l < L; l++) {
            double x = e->x[p + l * Nph];
            double g = e->dtau * (e->ha[c] + 2 * e->ha2[c] * x + 3 * e->ha3[c] * x * x + 4 * e->ha4[c] * x * x * x);
            F[p + l * Nph] += nu * creal(conj(up[l + i * L]) * g * vp[l + i * L]);
        }
    }
}
/* mul_νRe∂M∂x!: Sym :2-114, Asym :117-191.  exact_holstein != 0 selects Gamma^-1 instead of the
 * reference's Gamma^-T in the Holstein-only Sym branch. */
void ref_mul_nuRe_dM_dx(double *F, double nu, const cplx *u, const cplx *v, ref_fdm *f, const ref_elph *e, int exact_holstein) {
    i64 L = f->L, N = f->N, Nh = f->Nh, C = f->C;
    cplx *vp = f->tmp1, *up = f->tmp2;
    /* hopping_to_couplings as linked lists */
    i64 *first = (i64 *)malloc(sizeof(i64) * (Nh + 1)), *next = (i64 *)malloc(sizeof(i64) * (e->Nssh + 1));
    for (i64 h = 0; h < Nh; h++) first[h] = -1;
    for (i64 c = e->Nssh - 1; c >= 0; c--) { next[c] = first[e->ssh_hop[c]]; first[e->ssh_hop[c]] = c; }
    circshift1(vp, v, L, N);                                                     /* :24 */
    for (i64 i = 0; i < N; i++) for (i64 l = 1; l < L; l++) vp[l + i * L] = -vp[l + i * L];   /* :27 */
    if (f->sym) {
        chk_apply(vp, f, f->ch, f->sh, L, 1, 0, 0, Nh);                        /* :30 */
        vec_scale_real(vp, f->expV, L * N, 0);                                   /* :33 */
        chk_apply(vp, f, f->ch, f->sh, L, 0, 0, 0, Nh);                        /* :36 */
        memcpy(up, u, sizeof(cplx) * L * N);                                    /* :39 */
        if (e->Nssh > 0) {
            for (i64 c = C - 1; c >= 0; c--) {                                  /* :50-63 */
                dKc_dx(F, -nu, up, vp, f, e, e->dtau / 2, c, first, next);
                chk_apply(up, f, f->ch, f->sh, L, 0, 0, f->clo[c], f->chi[c]);
                chk_apply(vp, f, f->ch, f->sh, L, 0, 1, f->clo[c], f->chi[c]);
            }
        } else {
            chk_apply(up, f, f->ch, f->sh, L, 1, 0, 0, Nh);                    /* :66-69 */
            chk_apply(vp, f, f->ch, f->sh, L, exact_holstein ? 0 : 1, 1, 0, Nh); /* :71-74 (Q1) */
        }
        if (e->Nhol > 0) dV_dx(F, -nu, up, vp, e);                               /* :82 */
        vec_scale_real(up, f->expV, L * N, 0);                                   /* :87 */
        vec_scale_real(vp, f->expV, L * N, 1);                                   /* :90 */
        if (e->Nssh > 0) {
            for (i64 c = 0; c < C; c++) {                                        /* :95-111 */
                dKc_dx(F, -nu, up, vp, f, e, e->dtau / 2, c, first, next);
                chk_apply(up, f, f->ch, f->sh, L, 0, 0, f->clo[c], f->chi[c]);
                chk_apply(vp, f, f->ch, f->sh, L, 0, 1, f->clo[c], f->chi[c]);
            }
        }
    } else {
        chk_apply(vp, f, f->ch, f->sh, L, 0, 0, 0, Nh);                        /* :144 */
        vec_scale_real(vp, f->expV, L * N, 0);                                   /* :147 */
        memcpy(up, u, sizeof(cplx) * L * N);                                    /* :150 */
        if (e->Nhol > 0) dV_dx(F, -nu, up, vp, e);                               /* :158 */
        if (e->Nssh > 0) {
            vec_scale_real(up, f->expV, L * N, 0);                               /* :166 */
            vec_scale_real(vp, f->expV, L * N, 1);                               /* :169 */
            for (i64 c = C - 1; c >= 0; c--) {                                  /* :172-187 */
                dKc_dx(F, -nu, up, vp, f, e, e->dtau, c, first, next);
                chk_apply(up, f, f->ch, f->sh, L, 0, 0, f->clo[c], f->chi[c]);
                chk_apply(vp, f, f->ch, f->sh, L, 1, 1, f->clo[c], f->chi[c]);
            }
        }
    }
    free(first); free(next);
}

/* ------------------------------------------------------------------------------------------ */
/* PFFCalculator: src/PFFCalculator.jl                                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct ref_pff { i64 L, N; cplx *Phi, *u, *up, *upp; double *Lam; int exact_holstein; } ref_pff;

ref_pff *ref_pff_create(i64 L, i64 N) {
    ref_pff *q = (ref_pff *)calloc(1, sizeof(ref_pff));
    q->L = L; q->N = N;
    q->Phi = (cplx *)calloc(L * N, sizeof(cplx)); q->u = (cplx *)calloc(L * N, sizeof(cplx));
    q->up = (cplx *)calloc(L * N, sizeof(cplx)); q->upp = (cplx *)calloc(L * N, sizeof(cplx));
    q->Lam = (double *)calloc(L * N, sizeof(double));
    return q;
}
void ref_pff_destroy(ref_pff *q) { if (!q) return; free(q->Phi); free(q->u); free(q->up); free(q->upp); free(q->Lam); free(q); }
cplx *ref_pff_Phi(ref_pff *q) { return q->Phi; }
cplx *ref_pff_Psi(ref_pff *q) { return q->u; }
double *ref_pff_Lambda(ref_pff *q) { return q->Lam; }
void ref_pff_set_exact_holstein(ref_pff *q, int flag) { q->exact_holstein = flag; }

/* sample_pseudofermion_fields!: :56-76.  R = the randn!(rng, Phi) draw, (L x N) complex. */
double ref_pff_sample(ref_pff *q, const ref_elph *e, ref_fdm *f, const cplx *R) {
    i64 n = q->L * q->N;
    ref_update_Lambda(q->Lam, e);
    memcpy(q->Phi, R, sizeof(cplx) * n);
    double Sf = creal(zdot(q->Phi, q->Phi, n));
    memcpy(f->tmp2, q->Phi, sizeof(cplx) * n);                       /* lmul_Mt!: FermionDetMatrix.jl:470-480 */
    ref_mul_Mt(f, q->Phi, f->tmp2);
    memcpy(f->tmp2, q->Phi, sizeof(cplx) * n);
    ref_mul_LambdaT(q->Phi, q->Lam, f->tmp2, q->L, q->N);            /* in-place safe in the reference (:118-125) */
    return Sf;
}
/* calculate_fermionic_action!: :79-116.  lanczos_start (N reals) is consumed iff P != NULL. */
double ref_pff_action(ref_pff *q, const ref_elph *e, ref_fdm *f, ref_kpm *P, const double *lanczos_start,
                      double tol, i64 maxiter, i64 *iters, double *eps, double *imag_part) {
    i64 n = q->L * q->N;
    ref_update_Lambda(q->Lam, e);
    ref_ldiv_LambdaT(q->u, q->Lam, q->Phi, q->L, q->N);               /* :97 */
    if (P && lanczos_start) ref_kpm_update(P, lanczos_start);          /* FermionDetMatrix.jl:259 */
    *iters = ref_cg(f, q->u, q->u, 1, P, tol, maxiter, eps);           /* :99-105 */
    memcpy(f->tmp2, q->u, sizeof(cplx) * n);
    ref_ldiv_Lambda(q->u, q->Lam, f->tmp2, q->L, q->N);               /* :107 */
    cplx Sf = zdot(q->Phi, q->u, n);                                   /* :109 */
    if (imag_part) *imag_part = cimag(Sf);
    return creal(Sf);
}
/* calculate_derivative_fermionic_action!: :119-158.  dSdx (Nph x L) += */
double ref_pff_force(double *dSdx, ref_pff *q, const ref_elph *e, ref_fdm *f, ref_kpm *P, const double *lanczos_start,
                     double tol, i64 maxiter, i64 *iters, double *eps) {
    double Sf = ref_pff_action(q, e, f, P, lanczos_start, tol, maxiter, iters, eps, NULL);
    cplx *Psi = q->u, *LPsi = q->up, *APsi = q->upp, *MtAPsi = q->up;
    ref_mul_Lambda(LPsi, q->Lam, Psi, q->L, q->N);                     /* :146 */
    ref_mul_M(f, APsi, LPsi);                                          /* :148 */
    ref_mul_nuRe_dM_dx(dSdx, -2.0, APsi, LPsi, f, e, q->exact_holstein); /* :150 */
    ref_mul_Mt(f, MtAPsi, APsi);                                       /* :153 */
    ref_mul_nuRe_dLambda_dx(dSdx, -2.0, MtAPsi, Psi, q->Lam, e);       /* :155 */
    return Sf;
}

/* ------------------------------------------------------------------------------------------ */
/* Exact Fourier acceleration  [unvendored: SmoQyDQMC ExactFourierAccelerator et al.]          */
/* Restated from arXiv:2404.09723: the free-boson action is diagonal in Matsubara space with  */
/* spring k_w = dtau M (Om^2 + 4 sin^2(pi w/L)/dtau^2); the dynamical mass Mt_w = k_w (+ eta   */
/* regularisation) makes every mode a unit-frequency oscillator that is evolved exactly.       */
/* ------------------------------------------------------------------------------------------ */
typedef struct ref_efa { i64 L, Nph; double *Mt, *wd; cplx *twf, *twb; } ref_efa;

ref_efa *ref_efa_create(const ref_elph *e, double eta) {
    ref_efa *a = (ref_efa *)calloc(1, sizeof(ref_efa));
    i64 L = e->L, Nph = e->Nph;
    a->L = L; a->Nph = Nph;
    a->Mt = (double *)calloc(Nph * L, sizeof(double)); a->wd = (double *)calloc(Nph * L, sizeof(double));
    a->twf = (cplx *)malloc(sizeof(cplx) * L); a->twb = (cplx *)malloc(sizeof(cplx) * L);
    for (i64 w = 0; w < L; w++) {
        a->twf[w] = cexp(-2 * I * M_PI * (double)w / (double)L); a->twb[w] = cexp(+2 * I * M_PI * (double)w / (double)L);
        double s = sin(M_PI * (double)w / (double)L);
        for (i64 p = 0; p < Nph; p++) {
            double Ot2 = e->Om[p] * e->Om[p] + 4 * s * s / (e->dtau * e->dtau);
            double k = e->dtau * e->M[p] * Ot2;
            double mt = e->dtau * e->M[p] * (Ot2 + eta * eta);
            a->Mt[p + w * Nph] = mt;
            a->wd[p + w * Nph] = isfinite(e->M[p]) ? sqrt(k / mt) : 0.0;
        }
    }
    return a;
}
void ref_efa_destroy(ref_efa *a) { if (!a) return; free(a->Mt); free(a->wd); free(a->twf); free(a->twb); free(a); }

/* unitary DFT along tau of a real/complex (Nph x L) array held as complex, phonon fastest */
static void efa_dft(const ref_efa *a, cplx *y, int forward) {
    i64 L = a->L, Nph = a->Nph; double sc = 1.0 / sqrt((double)L);
    cplx *in = (cplx *)malloc(sizeof(cplx) * L), *out = (cplx *)malloc(sizeof(cplx) * L);
    for (i64 p = 0; p < Nph; p++) {
        for (i64 l = 0; l < L; l++) in[l] = y[p + l * Nph];
        fft_rec(in, out, L, 1, forward ? a->twf : a->twb, 1, L);
        for (i64 l = 0; l < L; l++) y[p + l * Nph] = out[l] * sc;
    }
    free(in); free(out);
}
/* initialize_momentum!: p = F^-1 sqrt(Mt) F R, K = sum |p~|^2/(2 Mt) = |R|^2/2 over finite-mass modes */
double ref_efa_init_momentum(const ref_efa *a, const ref_elph *e, double *p, const double *R) {
    i64 L = a->L, Nph = a->Nph; cplx *y = (cplx *)malloc(sizeof(cplx) * L * Nph);
    for (i64 k = 0; k < L * Nph; k++) y[k] = R[k];
    efa_dft(a, y, 1);
    double K = 0;
    for (i64 w = 0; w < L; w++) for (i64 q = 0; q < Nph; q++) {
        i64 k = q + w * Nph;
        if (isfinite(e->M[q])) { y[k] *= sqrt(a->Mt[k]); K += (creal(y[k]) * creal(y[k]) + cimag(y[k]) * cimag(y[k])) / (2 * a->Mt[k]); }
        else y[k] = 0;
    }
    efa_dft(a, y, 0);
    for (i64 k = 0; k < L * Nph; k++) p[k] = creal(y[k]);
    free(y);
    return K;
}
double ref_efa_kinetic(const ref_efa *a, const ref_elph *e, const double *p) {
    i64 L = a->L, Nph = a->Nph; cplx *y = (cplx *)malloc(sizeof(cplx) * L * Nph);
    for (i64 k = 0; k < L * Nph; k++) y[k] = p[k];
    efa_dft(a, y, 1);
    double K = 0;
    for (i64 w = 0; w < L; w++) for (i64 q = 0; q < Nph; q++) {
        i64 k = q + w * Nph;
        if (isfinite(e->M[q])) K += (creal(y[k]) * creal(y[k]) + cimag(y[k]) * cimag(y[k])) / (2 * a->Mt[k]);
    }
    free(y);
    return K;
}
/* evolve_eom!(x, p, dt, efa): exact flow of H0 = sum p~^2/(2Mt) + k x~^2/2 */
void ref_efa_evolve(const ref_efa *a, const ref_elph *e, double *x, double *p, double dt) {
    i64 L = a->L, Nph = a->Nph;
    cplx *xs = (cplx *)malloc(sizeof(cplx) * L * Nph), *ps = (cplx *)malloc(sizeof(cplx) * L * Nph);
    for (i64 k = 0; k < L * Nph; k++) { xs[k] = x[k]; ps[k] = p[k]; }
    efa_dft(a, xs, 1); efa_dft(a, ps, 1);
    for (i64 w = 0; w < L; w++) for (i64 q = 0; q < Nph; q++) {
        i64 k = q + w * Nph;
        if (!isfinite(e->M[q])) continue;
        double wd = a->wd[k], mt = a->Mt[k], c = cos(wd * dt), s = sin(wd * dt);
        cplx X = xs[k], P = ps[k];
        if (wd > 0) { xs[k] = c * X + s / (mt * wd) * P; ps[k] = c * P - mt * wd * s * X; }
        else { xs[k] = X + dt / mt * P; }
    }
    efa_dft(a, xs, 0); efa_dft(a, ps, 0);
    for (i64 k = 0; k < L * Nph; k++) if (isfinite(e->M[k % Nph])) { x[k] = creal(xs[k]); p[k] = creal(ps[k]); }
    free(xs); free(ps);
}
/* DispersionParameters [unvendored: SmoQyDQMC; restated from the published Hamiltonian, arXiv:2311.09395 eq. for U_disp]:
 *   U_disp = sum_d M''_d [ Ω_d^2 (X_p' - X_p)^2 / 2 + Ω4_d^2 (X_p' - X_p)^4 / 24 ],  M'' = M_p M_p' / (M_p + M_p') (reduced mass; one
 *   infinite mass => the other one).  Phonon indices 0-based here. */
void ref_elph_set_dispersion(ref_elph *e, i64 Ndisp, const i64 *disp_ph, const double *Om, const double *Om4) {
    free(e->disp_ph); free(e->disp_Om); free(e->disp_Om4);
    e->Ndisp = Ndisp; e->disp_ph = dupi(disp_ph, 2 * Ndisp); e->disp_Om = dupd(Om, Ndisp); e->disp_Om4 = dupd(Om4, Ndisp);
}
static double reduced_mass(double a, double b) {
    if (!isfinite(a)) return b;
    if (!isfinite(b)) return a;
    return a * b / (a + b);
}
/* bosonic_action(elph, holstein_correction=false) [unvendored]: on-site harmonic + quartic, kinetic, dispersive */
double ref_bosonic_action(const ref_elph *e) {
    i64 L = e->L, Nph = e->Nph; double S = 0;
    for (i64 l = 0; l < L; l++) for (i64 p = 0; p < Nph; p++) {
        if (!isfinite(e->M[p])) continue;
        double x = e->x[p + l * Nph], xn = e->x[p + ((l + 1) % L) * Nph], d = xn - x;
        S += e->dtau * e->M[p] * e->Om[p] * e->Om[p] * x * x / 2 + e->dtau * e->M[p] * e->Om4[p] * e->Om4[p] * x * x * x * x / 24
           + e->M[p] * d * d / (2 * e->dtau);
    }
    for (i64 l = 0; l < L; l++) for (i64 c = 0; c < e->Ndisp; c++) {
        i64 p = e->disp_ph[2 * c], pp = e->disp_ph[2 * c + 1];
        double m = reduced_mass(e->M[p], e->M[pp]);
        if (!isfinite(m)) continue;
        double D = e->x[pp + l * Nph] - e->x[p + l * Nph];
        S += e->dtau * m * e->disp_Om[c] * e->disp_Om[c] * D * D / 2 + e->dtau * m * e->disp_Om4[c] * e->disp_Om4[c] * D * D * D * D / 24;
    }
    return S;
}
/* eval_derivative_dispersive_action! [unvendored] (src/EFAPFFHMCUpdater.jl:193): -= on the first phonon of the pair, += on the second */
void ref_dispersive_derivative(double *F, const ref_elph *e) {
    i64 L = e->L, Nph = e->Nph;
    for (i64 l = 0; l < L; l++) for (i64 c = 0; c < e->Ndisp; c++) {
        i64 p = e->disp_ph[2 * c], pp = e->disp_ph[2 * c + 1];
        double m = reduced_mass(e->M[p], e->M[pp]);
        if (!isfinite(m)) continue;
        double D = e->x[pp + l * Nph] - e->x[p + l * Nph];
        double g = e->dtau * m * (e->disp_Om[c] * e->disp_Om[c] * D + e->disp_Om4[c] * e->disp_Om4[c] * D * D * D / 6);
        if (isfinite(e->M[pp])) F[pp + l * Nph] += g;
        if (isfinite(e->M[p])) F[p + l * Nph] -= g;
    }
}
/* eval_derivative_anharmonic_action! [unvendored] */
void ref_anharmonic_derivative(double *F, const ref_elph *e) {
    for (i64 k = 0; k < e->L * e->Nph; k++) {
        i64 p = k % e->Nph; if (!isfinite(e->M[p])) continue;
        double x = e->x[k];
        F[k] += e->dtau * e->M[p] * e->Om4[p] * e->Om4[p] * x * x * x / 6;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* hmc_update!: src/EFAPFFHMCUpdater.jl:102-279                                                 */
/* Random numbers are consumed from `rnd` in this fixed order (documented in DESIGN.md):       */
/*   [0]            u for the dt jitter                                                         */
/*   2*L*N          Phi normals (re,im interleaved, host (L x N) order), each N(0,1/2)          */
/*   Nph*L          momentum normals                                                            */
/*   (Nt+1) * N     Lanczos start vectors, one per CG solve (only when P != NULL)               */
/*   1              u for the Metropolis test                                                   */
/* Returns accepted (0/1); out[0]=iters_avg, out[1]=dH, out[2]=Sf0, out[3]=Sf1, out[4]=Sb0,    */
/* out[5]=Sb1, out[6]=K0, out[7]=K1.                                                            */
/* ------------------------------------------------------------------------------------------ */
int ref_hmc_update(ref_elph *e, ref_fdm *f, ref_pff *q, ref_kpm *P, ref_efa *a, i64 Nt, double dt, double delta,
                   double tol_action, double tol_force, i64 maxiter, const double *rnd, double *out) {
    i64 L = e->L, N = e->N, Nph = e->Nph, nx = Nph * L;
    const double *rp = rnd;
    double *x0 = (double *)malloc(sizeof(double) * nx), *p = (double *)calloc(nx, sizeof(double)), *dS = (double *)malloc(sizeof(double) * nx);
    dt = dt * (1.0 + (2 * (*rp++) - 1) * delta);                                  /* :125 */
    memcpy(x0, e->x, sizeof(double) * nx);                                        /* :128 */
    cplx *R = (cplx *)malloc(sizeof(cplx) * L * N);
    for (i64 k = 0; k < L * N; k++) { R[k] = (rp[0] + I * rp[1]) * M_SQRT1_2; rp += 2; }
    double Sf = ref_pff_sample(q, e, f, R);                                       /* :131 */
    free(R);
    double Sb = ref_bosonic_action(e);                                            /* :136 */
    double K = ref_efa_init_momentum(a, e, p, rp); rp += nx;                      /* :142 */
    double H = Sf + Sb + K;
    ref_efa_evolve(a, e, e->x, p, dt / 2);                                        /* :149-153 */
    ref_elph_refresh(e, f);
    double iters_avg = 0; i64 iters; double eps;
    for (i64 t = 1; t <= Nt; t++) {                                               /* :162 */
        memset(dS, 0, sizeof(double) * nx);
        ref_pff_force(dS, q, e, f, P, P ? rp : NULL, tol_force, maxiter, &iters, &eps);   /* :171 */
        if (P) rp += N;
        iters_avg += (double)iters / (double)(Nt + 1);
        ref_anharmonic_derivative(dS, e);                                         /* :190 */
        ref_dispersive_derivative(dS, e);                                         /* :193 */
        for (i64 k = 0; k < nx; k++) p[k] -= dt * dS[k];                           /* :196 */
        ref_efa_evolve(a, e, e->x, p, t == Nt ? dt / 2 : dt);                      /* :200-205 */
        ref_elph_refresh(e, f);
    }
    double Sf1 = ref_pff_action(q, e, f, P, P ? rp : NULL, tol_action, maxiter, &iters, &eps, NULL);   /* :217 */
    if (P) rp += N;
    iters_avg += (double)iters / (double)(Nt + 1);
    double Sb1 = ref_bosonic_action(e), K1 = ref_efa_kinetic(a, e, p);            /* :238-244 */
    double dH = (Sf1 + Sb1 + K1) - H;
    double Pacc = fmin(1.0, exp(-dH));
    if (!isfinite(dH)) Pacc = 0.0;
    int accepted = (*rp++) < Pacc;                                                /* :263 */
    if (!accepted) { memcpy(e->x, x0, sizeof(double) * nx); ref_elph_refresh(e, f); }   /* :266-276 */
    if (out) { out[0] = iters_avg; out[1] = dH; out[2] = Sf; out[3] = Sf1; out[4] = Sb; out[5] = Sb1; out[6] = K; out[7] = K1; }
    free(x0); free(p); free(dS);
    return accepted;
}

/* ------------------------------------------------------------------------------------------ */
/* GreensEstimator solves + scalar measurements                                                 */
/* ------------------------------------------------------------------------------------------ */
/* update_greens_estimator!: src/Measurements/GreensEstimator.jl:125-175.  R (V x Nrv) holds the
 * unit-modulus random vectors on entry (the randn!/abs step is done by the caller so that GPU
 * and oracle see the same R); GR (V x Nrv) holds the warm start on entry, G R on exit. */
double ref_greens_update(ref_fdm *f, ref_kpm *P, const cplx *R, cplx *GR, i64 Nrv, double tol, i64 maxiter) {
    i64 V = f->L * f->N; double avg = 0, eps;
    cplx *MtR = (cplx *)malloc(sizeof(cplx) * V);
    for (i64 n = 0; n < Nrv; n++) {
        ref_mul_Mt(f, MtR, R + n * V);                                            /* :156 */
        avg += (double)ref_cg(f, GR + n * V, MtR, 0, P, tol, maxiter, &eps);      /* :159-165 */
    }
    free(MtR);
    return avg / (double)Nrv;
}
/* measure_n: src/Measurements/scalar_measurements.jl:15-28 (R = conj(Rt)) */
void ref_measure_n(const cplx *R, const cplx *GR, i64 V, i64 Nrv, double *out) {
    cplx n = 1.0 - zdot(R, GR, V * Nrv) / (double)(V * Nrv);
    out[0] = creal(n); out[1] = cimag(n);
}
/* measure_double_occ: :112-147 */
void ref_measure_double_occ(const cplx *R, const cplx *GR, i64 V, i64 Nrv, double *out) {
    cplx d = 0; i64 np = Nrv * (Nrv - 1) / 2;
    for (i64 i = 0; i < Nrv - 1; i++) for (i64 j = i + 1; j < Nrv; j++) {
        cplx s = 0;
        for (i64 r = 0; r < V; r++) s += (1 - GR[r + i * V] * conj(R[r + i * V])) * (1 - GR[r + j * V] * conj(R[r + j * V]));
        d += s / (double)V;
    }
    d /= (double)np;
    out[0] = creal(d); out[1] = cimag(d);
}
/* measure_Nsqrd: :31-96 */
void ref_measure_Nsqrd(const cplx *R, const cplx *GR, i64 V, i64 L, i64 Nrv, double *out) {
    cplx Nb = 0, T2 = 0; double np = (double)(Nrv * (Nrv - 1) / 2);
    for (i64 i = 0; i < Nrv - 1; i++) {
        cplx Ti = zdot(R + i * V, GR + i * V, V);
        for (i64 j = i + 1; j < Nrv; j++) {
            cplx Tj = zdot(R + j * V, GR + j * V, V);
            Nb += 4.0 * ((double)V - Ti) * ((double)V - Tj) / (double)(L * L);
            T2 += zdot(R + j * V, GR + i * V, V) * zdot(R + i * V, GR + j * V, V) / (double)(L * L);
        }
    }
    Nb /= np; T2 /= np;
    cplx TrG = zdot(R, GR, V * Nrv) / (double)(Nrv * L);
    cplx r = Nb + 2.0 * TrG / (double)L - 2.0 * T2;
    out[0] = creal(r); out[1] = cimag(r);
}

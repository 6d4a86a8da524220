// cg.cu -- conjugate gradient on M^T M, device resident.
//
// Replaces /root/reference/src/IterativeSolvers/ConjugateGradient.jl:93-167 (P = I) and :169-249
// (left preconditioner) plus ldiv!(x, fdm, b) at src/FermionDetMatrix.jl:248-267.
//
// The recurrence, stopping rule (|r|/|b| < tol checked before the loop and after every r update),
// complex alpha/beta and the returned (iters, eps) are the reference's.  All scalars live in a
// ping-ponged device struct; kernels of iterations launched after convergence exit immediately on the
// `done` flag, so the host only synchronises once per batch of iterations.  BLAS-1 work is fused:
//   K_A  z = M^T M p and per-CTA partials of p.A p = |M p|^2            (fdm.cu, fused matvec)
//   K_B  alpha, x += alpha p, r -= alpha z, partials of |r|^2             (one pass over x, r, p, z)
//   K_C  eps / convergence, beta, p = r + beta p  (P = I)                 (one pass over r, p)
// With a preconditioner K_C splits into the convergence check, P^-1 r, the r.z partials and the
// p update.  Reductions are deterministic (fixed-order partial sums, no floating-point atomics).
#include "sq_internal.h"

// reduce partial sums of up to 3 quantities laid out as part[q*SQ_MAXPART + k]; result in all threads of warp 0,
// broadcast through shared memory to the block.
__device__ __forceinline__ void reduce_partials(const double *part, int n, int nq, double *sh /*[3]*/) {
    if (threadIdx.x < 32) {
        for (int q = 0; q < nq; q++) {
            double s = warp_sum_partials(part + (size_t)q * SQ_MAXPART, n);
            if (threadIdx.x == 0) sh[q] = s;
        }
    }
    __syncthreads();
}

// partials of conj(a).b (re, im) and |a|^2
thread_local int g_sq_pdl = 0;
// scope guard: the launches inside take the programmatic-dependent-launch attribute (common.cuh)
struct PdlScope {
    explicit PdlScope(bool on) { g_sq_pdl = on ? 1 : 0; }
    ~PdlScope() { g_sq_pdl = 0; }
};

__global__ void k_dot_partials(const double2 *__restrict__ a, const double2 *__restrict__ b, size_t n, double *__restrict__ part,
                               const CgState *__restrict__ skip) {
    __shared__ double red[3 * 32];
    if (skip && skip->done) return;
    double v[3] = {0, 0, 0};
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        double2 x = a[k], y = b[k];
        v[0] += x.x * y.x + x.y * y.y;
        v[1] += x.x * y.y - x.y * y.x;
        v[2] += x.x * x.x + x.y * x.y;
    }
    block_sum<3>(v, red);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = v[0];
        part[SQ_MAXPART + blockIdx.x] = v[1];
        part[2 * SQ_MAXPART + blockIdx.x] = v[2];
    }
}

// r = b - r   (axpby!(1, b, -1, r), ConjugateGradient.jl:196)
__global__ void k_residual(double2 *__restrict__ r, const double2 *__restrict__ b, size_t n) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        double2 x = b[k], y = r[k];
        r[k] = make_double2(x.x - y.x, x.y - y.y);
    }
}

// initial state: normb from part_b (|b|^2 in slot 2), r.z / |r|^2 from part_rz; eps0 test (:206-214)
__global__ void k_cg_init(CgState *st, const double *part_b, int nb, const double *part_rz, int nrz, double tol) {
    __shared__ double sh[3], shb[3];
    reduce_partials(part_b, nb, 3, shb);
    reduce_partials(part_rz, nrz, 3, sh);
    if (threadIdx.x == 0) {
        CgState s;
        s.normb = sqrt(shb[2]);
        s.rz_re = sh[0];
        s.rz_im = sh[1];
        s.eps = sqrt(sh[2]) / s.normb;
        s.tol = tol;
        s.iters = 0;
        s.done = (s.eps < tol) ? 1 : 0;
        if (!(s.eps == s.eps)) s.done = 2;            // NaN: stop, reported as an error by the host
        st[0] = s;
        st[1] = s;
    }
}

// K_B: alpha = rz / pAp ; x += alpha p ; r -= alpha z ; partial |r|^2      (:219-226)
// pAp_complex != 0: p.Ap is taken from the complex partials (re, im) instead of the real |Mp|^2.
__global__ void k_cg_update_xr(const CgState *__restrict__ st, double2 *__restrict__ x, double2 *__restrict__ r,
                               const double2 *__restrict__ p, const double2 *__restrict__ z, size_t n,
                               const double *__restrict__ pAp_part, int npart, int pAp_complex, double *rr_part,
                               CgState *nxt = nullptr, unsigned *ticket = nullptr, int iter = 0) {
    // nxt != nullptr (preconditioned loop): the block that finishes last also performs the convergence test of k_cg_check on the
    // fixed-order sum of the |r|^2 partials, which saves one launch per iteration
    __shared__ double sh[3];
    __shared__ double red[32];
    __shared__ int last;
    if (st->done) {
        if (nxt && blockIdx.x == 0 && threadIdx.x == 0) *nxt = *st;
        return;
    }
    reduce_partials(pAp_part, npart, pAp_complex ? 2 : 1, sh);
    double2 pAp = make_double2(sh[0], pAp_complex ? sh[1] : 0.0);
    double2 alpha = cdiv(make_double2(st->rz_re, st->rz_im), pAp);
    double acc = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        double2 pk = p[k], zk = z[k], xk = x[k], rk = r[k];
        double2 ap = cmul(alpha, pk), az = cmul(alpha, zk);
        xk = cadd(xk, ap);
        rk = csub(rk, az);
        x[k] = xk;
        r[k] = rk;
        acc += rk.x * rk.x + rk.y * rk.y;
    }
    double v[1] = {acc};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) rr_part[blockIdx.x] = v[0];
    if (!nxt) return;
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 32) {
        const volatile double *vp = rr_part;
        double t = 0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += 32) t += vp[k];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) {
            CgState c = *st;
            c.eps = sqrt(t) / st->normb;
            c.iters = iter;
            c.done = (c.eps < st->tol) ? 1 : 0;
            if (!(c.eps == c.eps)) c.done = 2;
            *nxt = c;
            *ticket = 0;
        }
    }
}

// K_C (P = I): eps = |r|/|b|; stop or beta = rr_new/rr_old, p = r + beta p     (:229-245)
__global__ void k_cg_update_p(const CgState *__restrict__ cur, CgState *__restrict__ nxt, double2 *__restrict__ p,
                              const double2 *__restrict__ r, size_t n, const double *__restrict__ rr_part, int npart, int iter) {
    __shared__ double sh[3];
    if (cur->done) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *nxt = *cur;
        return;
    }
    reduce_partials(rr_part, npart, 1, sh);
    double rr = sh[0];
    double eps = sqrt(rr) / cur->normb;
    bool stop = (eps < cur->tol) || !(eps == eps);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        CgState s = *cur;
        s.eps = eps;
        s.iters = iter;
        s.done = stop ? ((eps == eps) ? 1 : 2) : 0;
        s.rz_re = rr;
        s.rz_im = 0.0;
        *nxt = s;
    }
    if (stop) return;
    double2 beta = cdiv(make_double2(rr, 0.0), make_double2(cur->rz_re, cur->rz_im));
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        double2 pk = p[k], rk = r[k];
        p[k] = cadd(rk, cmul(beta, pk));
    }
}

// preconditioned: convergence check only (writes nxt without touching rz)
__global__ void k_cg_check(const CgState *__restrict__ cur, CgState *__restrict__ nxt, const double *__restrict__ rr_part,
                           int npart, int iter) {
    __shared__ double sh[3];
    if (cur->done) {
        if (threadIdx.x == 0) *nxt = *cur;
        return;
    }
    reduce_partials(rr_part, npart, 1, sh);
    if (threadIdx.x == 0) {
        CgState s = *cur;
        s.eps = sqrt(sh[0]) / cur->normb;
        s.iters = iter;
        s.done = (s.eps < cur->tol) ? 1 : 0;
        if (!(s.eps == s.eps)) s.done = 2;
        *nxt = s;
    }
}
// preconditioned: beta = (r.z)_new / (r.z)_old ; p = z + beta p ; state->rz updated by block 0 AFTER all reads.
// `st` is the state written by k_cg_check; the old r.z is read from `old` (the other ping-pong slot).
__global__ void k_cg_update_p_prec(CgState *__restrict__ st, const CgState *__restrict__ old, double2 *__restrict__ p,
                                   const double2 *__restrict__ z, size_t n, const double *__restrict__ rz_part, int npart) {
    sq_pdl_prologue();
    __shared__ double sh[3];
    if (st->done) return;
    reduce_partials(rz_part, npart, 2, sh);
    double2 rz = make_double2(sh[0], sh[1]);
    double2 beta = cdiv(rz, make_double2(old->rz_re, old->rz_im));
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        double2 pk = p[k], zk = z[k];
        p[k] = cadd(zk, cmul(beta, pk));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->rz_re = rz.x; st->rz_im = rz.y; }
}

bool fdm_v2_supported(const sq_fdm *f, int mode, int S, int T);
void fdm_select_tuning(sq_fdm *f);
bool fdm_v2_cg_persistent(sq_fdm *f, double2 *x, double2 *r, double2 *p0, double2 *p1, CgState *state, double *part_a, double *part_b,
                          i64 maxiter);
int fdm_v2_launch_cg(sq_fdm *f, double2 *z, const double2 *p_old, double2 *p_new, const double2 *d, const CgState *cur, CgState *nxt,
                     const double *rr_part, int nrr, const double *beta_part, int nbeta, int beta_complex, int iter, int check,
                     double *pAp_part);

bool fdm_v3_supported(const sq_fdm *f, int S);
int fdm_v3_launch_cg(sq_fdm *f, double2 *z, const double2 *p_old, double2 *p_new, const double2 *d, const CgState *cur, CgState *nxt,
                     const double *rr_part, int nrr, const double *beta_part, int nbeta, int beta_complex, int iter, int check,
                     double *pAp_part, bool native);
void fdm_v3_prepare_native(sq_fdm *f);
bool fdm_v3_cg_resident(sq_fdm *f, double2 *x, double2 *r, CgState *state, i64 maxiter);
void fdm_v3_to_native(sq_fdm *f, double2 *dst, const double2 *src);
void fdm_v3_from_native(sq_fdm *f, double2 *dst, const double2 *src);

// end-of-batch convergence test for the fused scheme (the in-kernel test of iteration j runs in iteration j+1)
__global__ void k_cg_final_check(CgState *st, const double *__restrict__ rr_part, int npart, int iter) {
    __shared__ double sh[3];
    if (st->done) return;
    reduce_partials(rr_part, npart, 1, sh);
    if (threadIdx.x == 0) {
        double eps = sqrt(sh[0]) / st->normb;
        st->eps = eps;
        st->iters = iter;
        st->done = (eps < st->tol) ? 1 : ((eps == eps) ? 0 : 2);
    }
}

static int vec_grid(const sq_fdm *f, size_t n, int threads) {
    size_t want = (n + threads - 1) / threads;
    size_t cap = (size_t)f->num_sms * 4;
    size_t g = std::min(want, cap);
    g = std::min<size_t>(g, SQ_MAXPART);
    return (int)std::max<size_t>(g, 1);
}

// Solve M^T M x = b.  x, b device vectors in the [l][i] layout.  Follows cg_solve! step by step.
void fdm_cg_dev(sq_fdm *f, double2 *x, const double2 *b, bool zero_start, sq_kpm *kpm, double tol, i64 maxiter,
                i64 *iters, double *eps) {
    const bool kpm_on = kpm != nullptr && kpm->active;
    if (f->force_local) {
        // (a rank of a sharded chain solving one of the distributed right-hand sides on its own: fall through to the single-GPU solvers)
    } else if (f->sharded && (!kpm_on || (f->world > 1 && !getenv("SQ_SHARD_KPM_LOCAL")))) {
        // one chain over several GPUs: partitioned solve, solution gathered on every rank.  With an active preconditioner the apply is
        // sharded by Matsubara frequency (two all-to-all exchanges, slab.cu); SQ_SHARD_KPM_LOCAL=1 keeps such solves local and replicated
        fdm_cg_sharded(f, x, b, zero_start, kpm_on ? kpm : nullptr, tol, maxiter, iters, eps);
        return;
    }
    if (!f->force_local && ((f->world > 1 && !f->sharded) || f->slab_lo != 0 || f->slab_hi != (int)f->L || getenv("SQ_FORCE_SLAB_CG"))) {
        fdm_cg_slab(f, x, b, zero_start, kpm_on ? kpm : nullptr, tol, maxiter, iters, eps);
        return;
    }
    const size_t n = (size_t)f->L * f->N;
    const int TB = 256, G = vec_grid(f, n, TB);
    cudaStream_t s = f->stream;
    double *part = f->part.p;
    double *part_b = part, *part_rz = part + 3 * SQ_MAXPART, *part_pAp = part + 6 * SQ_MAXPART, *part_rr = part + 7 * SQ_MAXPART;
    double2 *r = f->r.p, *p = f->p.p, *z = f->z.p;
    CgState *st = f->cg.p;
    unsigned *ticket = f->cg_ticket.p;
    const bool prec = (kpm != nullptr) && kpm->active;

    k_dot_partials<<<G, TB, 0, s>>>(b, b, n, part_b, nullptr);                      // |b|
    if (zero_start) {
        if (x != b) {
            SQ_CUDA(cudaMemcpyAsync(r, b, n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        } else {
            SQ_CUDA(cudaMemcpyAsync(r, b, n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        }
        SQ_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double2), s));
    } else {
        fdm_mul_dev(f, SQ_OP_MTM, r, x);
        k_residual<<<G, TB, 0, s>>>(r, b, n);
    }
    if (prec) {
        kpm_ldiv_dev(kpm, z, r);
        SQ_CUDA(cudaMemcpyAsync(p, z, n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        k_dot_partials<<<G, TB, 0, s>>>(r, z, n, part_rz, nullptr);
    } else {
        SQ_CUDA(cudaMemcpyAsync(p, r, n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        k_dot_partials<<<G, TB, 0, s>>>(r, r, n, part_rz, nullptr);
    }
    k_cg_init<<<1, 64, 0, s>>>(st, part_b, G, part_rz, G, tol);
    SQ_LAUNCH_CHECK();
    f->launches += 4;
    f->stats[SQ_STAT_CG_SOLVES]++;
    if (maxiter <= 0) {                     // no iteration may run: report the initial residual (nothing else has written the partials)
        SQ_CUDA(cudaMemcpyAsync(f->h_cg, st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
        SQ_CUDA(cudaStreamSynchronize(s));
        if (f->h_cg->done == 2) throw SqNumericalInstability("conjugate gradient: NaN encountered in the residual (numerical instability)");
        *iters = 0;
        *eps = f->h_cg->eps;
        return;
    }

    int batch = prec ? 4 : 16;
    if (const char *eb = getenv("SQ_CG_BATCH")) batch = std::max(1, atoi(eb));
    i64 it = 0;
    int cur = 0;
    bool finished = false;
    fdm_select_tuning(f);
    const bool v3 = f->path == 0 && f->v3_cg && fdm_v3_supported(f, f->v3_S);
    const bool fused = !prec && f->path == 0 && (v3 || (f->use_v2 && fdm_v2_supported(f, 2, f->slab, f->threads))) && !getenv("SQ_NO_CG_FUSION");
    if (!prec && !v3 && f->path == 0 && f->v3g_ok && !getenv("SQ_NO_PERSISTENT_CG") && !getenv("SQ_NO_RESIDENT_CG") && !getenv("SQ_NO_CG_FUSION") &&
        !getenv("SQ_V3_NO_GRAPH")) {
        // small lattices (N <= 64): the whole solve in the resident register kernel with the graph engine (fdm_v3.cu)
        SQ_CUDA(cudaMemcpyAsync(f->h_cg, st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
        SQ_CUDA(cudaStreamSynchronize(s));
        if (f->h_cg->done == 2) throw SqNumericalInstability("conjugate gradient: NaN encountered in the residual (numerical instability)");
        if (f->h_cg->done) { *iters = 0; *eps = f->h_cg->eps; return; }
        if (fdm_v3g_cg(f, x, r, zero_start, st, maxiter)) {
            SQ_CUDA(cudaMemcpyAsync(f->h_cg, st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
            SQ_CUDA(cudaStreamSynchronize(s));
            f->stats[SQ_STAT_CG_RESIDENT]++;
            if (f->h_cg->done == 3) { f->stats[SQ_STAT_WATCHDOG]++; throw SqError("conjugate gradient: a grid-wide sum of the resident kernel timed out (watchdog)"); }
            if (f->h_cg->done == 2) throw SqNumericalInstability("conjugate gradient: NaN encountered in the residual (numerical instability)");
            *iters = f->h_cg->done ? f->h_cg->iters : maxiter;
            *eps = f->h_cg->eps;
            f->stats[SQ_STAT_CG_ITERS] += *iters;
            return;
        }
    }
    if (fused && !v3 && !getenv("SQ_NO_PERSISTENT_CG") && maxiter > 0) {
        // One cooperative launch for the whole solve (fdm_v2.cu: k_cg_persistent).  The state prepared by k_cg_init holds
        // |r0|^2, |b| and tol; an already converged system (eps0 < tol) is caught first.
        SQ_CUDA(cudaMemcpyAsync(f->h_cg, st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
        SQ_CUDA(cudaStreamSynchronize(s));
        if (f->h_cg->done == 2) throw SqNumericalInstability("conjugate gradient: NaN encountered in the residual (numerical instability)");
        if (f->h_cg->done) { *iters = 0; *eps = f->h_cg->eps; return; }
        if (fdm_v2_cg_persistent(f, x, r, p, f->tmp1.p, st, part_pAp, part_rr, maxiter)) {
            SQ_CUDA(cudaMemcpyAsync(f->h_cg, st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
            SQ_CUDA(cudaStreamSynchronize(s));
            f->stats[SQ_STAT_CG_PERSIST_V2]++;
            if (f->h_cg->done == 3) { f->stats[SQ_STAT_WATCHDOG]++; throw SqError("conjugate gradient: grid barrier timed out in the persistent kernel (watchdog)"); }
            if (f->h_cg->done == 2) throw SqNumericalInstability("conjugate gradient: NaN encountered in the residual (numerical instability)");
            *iters = f->h_cg->done ? f->h_cg->iters : maxiter;
            *eps = f->h_cg->eps;
            f->stats[SQ_STAT_CG_ITERS] += *iters;
            return;
        }
    }
    if (fused) {
        // Two kernels per iteration: K_A' forms p = r + beta p on load (ping-ponged p buffers), tests the convergence of
        // the previous iteration and applies M^T M; K_B updates x and r.  tmp1 is the second p buffer.
        // Register path: x, r, p, z live in the kernel's native order for the whole solve (every access coalesced); alpha and
        // beta are real without a preconditioner, so the BLAS-1 kernels are order-agnostic.  x is converted back at the end.
        double2 *pb[2] = {p, f->tmp1.p};
        int pc = 0;
        double2 *x_user = x;
        const bool native = v3 && !getenv("SQ_V3_NO_NATIVE");
        if (native) {
            fdm_v3_prepare_native(f);
            fdm_v3_to_native(f, f->v3_r.p, r);
            if (zero_start) SQ_CUDA(cudaMemsetAsync(f->v3_x.p, 0, n * sizeof(double2), s));
            else fdm_v3_to_native(f, f->v3_x.p, x);
            SQ_CUDA(cudaMemcpyAsync(pb[0], f->v3_r.p, n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
            x = f->v3_x.p;
            r = f->v3_r.p;
            if (!getenv("SQ_NO_PERSISTENT_CG") && maxiter > 0) {
                // one cooperative launch for the whole solve (fdm_v3.cu: k_cg_v3_resident1); k_cg_init left |r0|^2, |b|, tol in st
                SQ_CUDA(cudaMemcpyAsync(f->h_cg, st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
                SQ_CUDA(cudaStreamSynchronize(s));
                if (f->h_cg->done == 2) throw SqNumericalInstability("conjugate gradient: NaN encountered in the residual (numerical instability)");
                if (f->h_cg->done) { *iters = 0; *eps = f->h_cg->eps; return; }
                if (!getenv("SQ_NO_RESIDENT_CG") && fdm_v3_cg_resident(f, x, r, st, maxiter)) {
                    fdm_v3_from_native(f, x_user, x);
                    SQ_CUDA(cudaMemcpyAsync(f->h_cg, st, sizeof(CgState), cudaMemcpyDeviceToHost, s));
                    SQ_CUDA(cudaStreamSynchronize(s));
                    f->stats[SQ_STAT_CG_RESIDENT]++;
                    if (f->h_cg->done == 3) { f->stats[SQ_STAT_WATCHDOG]++; throw SqError("conjugate gradient: a grid-wide sum of the resident kernel timed out (watchdog)"); }
                    if (f->h_cg->done == 2) throw SqNumericalInstability("conjugate gradient: NaN encountered in the residual (numerical instability)");
                    *iters = f->h_cg->done ? f->h_cg->iters : maxiter;
                    *eps = f->h_cg->eps;
                    f->stats[SQ_STAT_CG_ITERS] += *iters;
                    return;
                }
                SQ_CUDA(cudaMemcpyAsync(pb[0], f->v3_r.p, n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
            }
        }
        while (!finished) {
            i64 upto = std::min<i64>(maxiter, it + batch);
            for (; it < upto;) {
                it++;
                const CgState *sc = st + cur;
                CgState *sn = st + (cur ^ 1);
                int npart = 0;
                if (it == 1) {
                    if (native) npart = fdm_v3_launch(f, 2, f->v3_S, z, pb[pc], part_pAp, sc, true);
                    else fdm_mul_dev(f, SQ_OP_MTM, z, pb[pc], part_pAp, &npart, sc);       // p0 = r0, nothing to fuse
                    k_cg_update_xr<<<G, TB, 0, s>>>(sc, x, r, pb[pc], z, n, part_pAp, npart, 0, part_rr);
                } else {
                    npart = v3 ? fdm_v3_launch_cg(f, z, pb[pc], pb[pc ^ 1], r, sc, sn, part_rr, G, part_rr, G, 0, (int)it, 1, part_pAp, native)
                               : fdm_v2_launch_cg(f, z, pb[pc], pb[pc ^ 1], r, sc, sn, part_rr, G, part_rr, G, 0, (int)it, 1, part_pAp);
                    pc ^= 1;
                    cur ^= 1;
                    k_cg_update_xr<<<G, TB, 0, s>>>(sn, x, r, pb[pc], z, n, part_pAp, npart, 0, part_rr);
                }
                f->launches += 1;
            }
            k_cg_final_check<<<1, 64, 0, s>>>(st + cur, part_rr, G, (int)it);
            SQ_LAUNCH_CHECK();
            f->launches++;
            SQ_CUDA(cudaMemcpyAsync(f->h_cg, st + cur, sizeof(CgState), cudaMemcpyDeviceToHost, s));
            SQ_CUDA(cudaStreamSynchronize(s));
            if (f->h_cg->done || it >= maxiter) finished = true;
        }
        if (native) fdm_v3_from_native(f, x_user, x);
        f->stats[SQ_STAT_CG_LOOP]++;
        if (f->h_cg->done == 2) throw SqNumericalInstability("conjugate gradient: NaN encountered in the residual (numerical instability)");
        *iters = f->h_cg->done ? f->h_cg->iters : maxiter;
        *eps = f->h_cg->eps;
        f->stats[SQ_STAT_CG_ITERS] += *iters;
        return;
    }
    // is the system already solved?  (cheap check folded into the first batch read-back)
    if (prec) SQ_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), s));
    const bool fuse_v3 = f->path == 0 && f->use_v3 && fdm_v3_supported(f, f->v3_S);
    const bool fuse_v2 = f->path == 0 && !fuse_v3 && f->use_v2 && fdm_v2_supported(f, 2, f->slab, f->threads);
    // (pays while the iteration is launch-bound; at cfg4 -- 410k elements -- the fused kernel's extra strided loads cost more than
    // the launch it saves: measured 197 against 192 us per iteration)
    if (prec && (fuse_v3 || fuse_v2) && n <= 200000 && !getenv("SQ_NO_CG_FUSION")) {
        // Preconditioned iteration in five launches: the fused matvec forms p = z + beta p on load (beta = (r.z)_new / (r.z)_old
        // from the partials the inverse FFT left behind) and writes M^T M p to its own buffer; the x / r update carries the
        // convergence test; then FFT, Chebyshev, inverse FFT.  States: A = st[0] is read by the x / r update and written by the
        // matvec, B = st[1] the other way round, so no kernel reads a state another block of the same kernel writes.
        if (f->prec_q.n < n) f->prec_q.alloc(n, false);
        double2 *q = f->prec_q.p, *pb[2] = {p, f->tmp1.p};
        CgState *A = st, *B = st + 1;
        int pc = 0, g = 0;
        while (!finished) {
            i64 step = batch;
            if (it == 0 && !getenv("SQ_CG_BATCH")) step = std::max<i64>(batch, std::min<i64>((i64)(0.85 * f->prec_iters_hint[tol < 1e-7 ? 0 : 1]), 256));
            const i64 upto = std::min<i64>(maxiter, it + step);
            PdlScope pdl(!getenv("SQ_NO_PDL"));
            for (; it < upto;) {
                it++;
                int npart = 0;
                if (it == 1) {
                    fdm_mul_dev(f, SQ_OP_MTM, q, pb[pc], part_pAp, &npart, A);
                } else {
                    npart = fuse_v3 ? fdm_v3_launch_cg(f, q, pb[pc], pb[pc ^ 1], z, B, A, part_rr, G, part_rz, g, 1, (int)it, 0, part_pAp, false)
                                    : fdm_v2_launch_cg(f, q, pb[pc], pb[pc ^ 1], z, B, A, part_rr, G, part_rz, g, 1, (int)it, 0, part_pAp);
                    pc ^= 1;
                }
                if (!getenv("SQ_NO_FFT_FUSION")) {            // x / r update + convergence test inside the forward transform's load phase
                    FftCgUpdate U = {x, r, pb[pc], q, A, B, part_pAp, npart, 0, part_rr, ticket, (int)it};
                    g = kpm_ldiv_dev_fused(kpm, z, U, part_rz);
                } else {
                    k_cg_update_xr<<<G, TB, 0, s>>>(A, x, r, pb[pc], q, n, part_pAp, npart, 0, part_rr, B, ticket, (int)it);
                    g = kpm_ldiv_dev_dot(kpm, z, r, B, r, part_rz);
                    f->launches += 1;
                }
            }
            SQ_LAUNCH_CHECK();
            SQ_CUDA(cudaMemcpyAsync(f->h_cg, B, sizeof(CgState), cudaMemcpyDeviceToHost, s));
            SQ_CUDA(cudaStreamSynchronize(s));
            if (f->h_cg->done || it >= maxiter) finished = true;
        }
        if (f->h_cg->done == 2) throw SqNumericalInstability("conjugate gradient: NaN encountered in the residual (numerical instability)");
        *iters = f->h_cg->done ? f->h_cg->iters : maxiter;
        *eps = f->h_cg->eps;
        f->stats[SQ_STAT_CG_PREC]++;
        f->stats[SQ_STAT_CG_ITERS] += *iters;
        if (f->h_cg->done) f->prec_iters_hint[tol < 1e-7 ? 0 : 1] = (int)std::min<i64>(*iters, 1 << 20);
        return;
    }
    const bool fuse_fft = prec && !getenv("SQ_NO_FFT_FUSION");
    while (!finished) {
        // preconditioned solves take a few tens of iterations and consecutive solves of a trajectory take about the same number:
        // the first read-back is placed shortly before the point where the previous solve of this tolerance class converged, later ones
        // every `batch` iterations
        i64 step = batch;
        if (prec && it == 0 && !getenv("SQ_CG_BATCH")) step = std::max<i64>(batch, std::min<i64>((i64)(0.85 * f->prec_iters_hint[tol < 1e-7 ? 0 : 1]), 256));
        i64 upto = std::min<i64>(maxiter, it + step);
        // (only while the iteration is launch-bound: at cfg4 the early-resident CTAs of the next kernel cost 3 %, at cfg1 / cfg5 the hidden
        // launch latency is worth 8 - 10 %)
        PdlScope pdl(prec && n <= 200000 && !getenv("SQ_NO_PDL"));
        for (; it < upto; ) {
            it++;
            const CgState *sc = st + cur;
            CgState *sn = st + (cur ^ 1);
            int npart = 0;
            fdm_mul_dev(f, SQ_OP_MTM, z, p, part_pAp, &npart, sc);
            if (!prec) {
                k_cg_update_xr<<<G, TB, 0, s>>>(sc, x, r, p, z, n, part_pAp, npart, 0, part_rr);
                k_cg_update_p<<<G, TB, 0, s>>>(sc, sn, p, r, n, part_rr, G, (int)it);
                f->launches += 2;
            } else {
                int g;
                if (fuse_fft) {
                    // x / r update and convergence test inside the load phase of the forward transform (which reads r anyway); z holds
                    // q = M^T M p on entry and P^-1 r on exit (the inverse transform writes it after the forward one has consumed it)
                    FftCgUpdate U = {x, r, p, z, sc, sn, part_pAp, npart, 0, part_rr, ticket, (int)it};
                    g = kpm_ldiv_dev_fused(kpm, z, U, part_rz);
                    f->launches += 1;
                } else {
                    // x, r update with the convergence test (k_cg_check) done by the last block to finish
                    k_cg_update_xr<<<G, TB, 0, s>>>(sc, x, r, p, z, n, part_pAp, npart, 0, part_rr, sn, ticket, (int)it);
                    g = kpm_ldiv_dev_dot(kpm, z, r, sn, r, part_rz);      // z = P^-1 r with the r.z partials fused in
                    f->launches += 2;
                }
                SQ_CUDA(sq_launch(k_cg_update_p_prec, dim3(G), dim3(TB), 0, s, sn, (const CgState *)sc, p, (const double2 *)z, n, (const double *)part_rz, g));
            }
            cur ^= 1;
        }
        SQ_LAUNCH_CHECK();
        SQ_CUDA(cudaMemcpyAsync(f->h_cg, st + cur, sizeof(CgState), cudaMemcpyDeviceToHost, s));
        SQ_CUDA(cudaStreamSynchronize(s));
        if (f->h_cg->done || it >= maxiter) finished = true;
    }
    if (f->h_cg->done == 2) throw SqNumericalInstability("conjugate gradient: NaN encountered in the residual (numerical instability)");
    *iters = f->h_cg->done ? f->h_cg->iters : maxiter;
    *eps = f->h_cg->eps;
    f->stats[prec ? SQ_STAT_CG_PREC : SQ_STAT_CG_LOOP]++;
    f->stats[SQ_STAT_CG_ITERS] += *iters;
    if (prec && f->h_cg->done) f->prec_iters_hint[tol < 1e-7 ? 0 : 1] = (int)std::min<i64>(*iters, 1 << 20);
}

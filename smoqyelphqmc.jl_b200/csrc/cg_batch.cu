// cg_batch.cu -- multi-RHS preconditioned CG: nrhs independent systems M^T M x_j = b_j advanced in lock step.
//
// The reference solves the Nrv right-hand sides of the Green's function estimator one after the other
// (src/Measurements/GreensEstimator.jl:152-169, each with the recurrence of src/IterativeSolvers/ConjugateGradient.jl:169-249 and a warm
// start from the previous G R column).  A preconditioned iteration is a chain of small kernels -- matvec, x / r update, forward
// tau-FFT, Chebyshev recurrences (one CTA per frequency and part: ~220 CTAs whose longest chain alone sets the time), inverse
// tau-FFT, p update -- none of which fills the chip at the named sizes.  Here every kernel takes the right-hand side as an extra
// grid dimension: one launch serves all systems, the Chebyshev stage runs nrhs times more chains in the time of the longest one,
// and the fused matvec works on nrhs x 16 MB instead of 16 MB per launch (several waves: it leaves the latency-bound regime).
// Each system keeps ITS OWN scalars, convergence test and iteration count -- the arithmetic per system is exactly that of the
// one-by-one solver (same kernels' formulas, fixed-order partial sums), so iteration counts and solutions are those of
// fdm_cg_dev; a system that has converged is skipped by every kernel (its CTAs exit on the `done` flag).
#include "sq_internal.h"

#include <algorithm>


// partials of conj(a).b (re, im) and |a|^2 per system: part[rhs][q][block]
__global__ void kb_dot_partials(const double2 *__restrict__ a, const double2 *__restrict__ b, size_t n, double *__restrict__ part) {
    __shared__ double red[3 * 32];
    a += (size_t)blockIdx.y * n;
    b += (size_t)blockIdx.y * n;
    part += (size_t)blockIdx.y * 3 * SQ_MAXPART;
    double v[3] = {0, 0, 0};
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const double2 x = a[k], y = b[k];
        v[0] += x.x * y.x + x.y * y.y;
        v[1] += x.x * y.y - x.y * y.x;
        v[2] += x.x * x.x + x.y * x.y;
    }
    block_sum<3>(v, red);
    if (threadIdx.x == 0) { part[blockIdx.x] = v[0]; part[SQ_MAXPART + blockIdx.x] = v[1]; part[2 * SQ_MAXPART + blockIdx.x] = v[2]; }
}
__global__ void kb_residual(double2 *__restrict__ r, const double2 *__restrict__ b, size_t n) {
    r += (size_t)blockIdx.y * n;
    b += (size_t)blockIdx.y * n;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const double2 x = b[k], y = r[k];
        r[k] = make_double2(x.x - y.x, x.y - y.y);
    }
}
// one block per system: |b|, r.z, eps0 and the initial convergence test (ConjugateGradient.jl:206-214); both ping-pong states
__global__ void kb_init(CgState *st, int nrhs, const double *part_b, const double *part_rz, int nb, double tol) {
    const int j = blockIdx.x;
    __shared__ double sh[6];
    if (threadIdx.x < 32) {
        const double *pb = part_b + (size_t)j * 3 * SQ_MAXPART, *pr = part_rz + (size_t)j * 3 * SQ_MAXPART;
        const double bb = warp_sum_partials(pb + 2 * SQ_MAXPART, nb);
        const double re = warp_sum_partials(pr, nb), im = warp_sum_partials(pr + SQ_MAXPART, nb), rr = warp_sum_partials(pr + 2 * SQ_MAXPART, nb);
        if (threadIdx.x == 0) { sh[0] = bb; sh[1] = re; sh[2] = im; sh[3] = rr; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        CgState s;
        s.normb = sqrt(sh[0]);
        s.rz_re = sh[1];
        s.rz_im = sh[2];
        s.eps = sqrt(sh[3]) / s.normb;
        s.tol = tol;
        s.iters = 0;
        s.done = (s.eps < tol) ? 1 : 0;
        if (!(s.eps == s.eps)) s.done = 2;
        st[j] = s;
        st[nrhs + j] = s;
    }
}
// alpha = (r.z) / (p.Ap); x += alpha p; r -= alpha q; |r|^2 partials; the block of a system that finishes last performs its
// convergence test (fixed-order sum of the partials) and writes the next state
__global__ void kb_update_xr(const CgState *__restrict__ cur, CgState *__restrict__ nxt, double2 *__restrict__ x, double2 *__restrict__ r,
                             const double2 *__restrict__ p, const double2 *__restrict__ q, size_t n, const double *__restrict__ pAp_part,
                             int npart, int pap_stride, double *__restrict__ rr_part, unsigned *__restrict__ ticket, int iter) {
    __shared__ double sh[1];
    __shared__ double red[32];
    __shared__ int last;
    const int j = blockIdx.y;
    const CgState st = cur[j];
    if (st.done) {
        if (blockIdx.x == 0 && threadIdx.x == 0) nxt[j] = st;
        return;
    }
    if (threadIdx.x < 32) {
        const double s = warp_sum_partials(pAp_part + (size_t)j * pap_stride, npart);
        if (threadIdx.x == 0) sh[0] = s;
    }
    __syncthreads();
    const double2 alpha = make_double2(st.rz_re / sh[0], st.rz_im / sh[0]);
    x += (size_t)j * n; r += (size_t)j * n; p += (size_t)j * n; q += (size_t)j * n;
    rr_part += (size_t)j * SQ_MAXPART;
    double acc = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const double2 pk = p[k], qk = q[k];
        const double2 xk = cadd(x[k], cmul(alpha, pk)), rk = csub(r[k], cmul(alpha, qk));
        x[k] = xk;
        r[k] = rk;
        acc += rk.x * rk.x + rk.y * rk.y;
    }
    double v[1] = {acc};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) rr_part[blockIdx.x] = v[0];
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(ticket + j, 1u) == gridDim.x - 1) ? 1 : 0;
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
            CgState c = st;
            c.eps = sqrt(t) / st.normb;
            c.iters = iter;
            c.done = (c.eps < st.tol) ? 1 : 0;
            if (!(c.eps == c.eps)) c.done = 2;
            nxt[j] = c;
            ticket[j] = 0;
        }
    }
}
// beta = (r.z)_new / (r.z)_old ; p = z + beta p ; the new r.z is stored by block 0 after all reads of the old one
__global__ void kb_update_p(CgState *__restrict__ st, const CgState *__restrict__ old, double2 *__restrict__ p, const double2 *__restrict__ z,
                            size_t n, const double *__restrict__ rz_part, int npart) {
    __shared__ double sh[2];
    const int j = blockIdx.y;
    if (st[j].done) return;
    if (threadIdx.x < 32) {
        const double *pr = rz_part + (size_t)j * 2 * SQ_MAXPART;
        const double re = warp_sum_partials(pr, npart), im = warp_sum_partials(pr + SQ_MAXPART, npart);
        if (threadIdx.x == 0) { sh[0] = re; sh[1] = im; }
    }
    __syncthreads();
    const double2 rz = make_double2(sh[0], sh[1]);
    const double2 beta = cdiv(rz, make_double2(old[j].rz_re, old[j].rz_im));
    p += (size_t)j * n;
    z += (size_t)j * n;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) p[k] = cadd(z[k], cmul(beta, p[k]));
    if (blockIdx.x == 0 && threadIdx.x == 0) { st[j].rz_re = rz.x; st[j].rz_im = rz.y; }
}

// z_j = M^T M p_j for all systems with the |M p_j|^2 partials (pap_stride doubles apart); returns the number of partials per system
static int mul_batch(sq_fdm *f, double2 *out, const double2 *in, int nrhs, size_t V, double *pAp_part, int pap_stride, const CgState *skip) {
    fdm_select_tuning(f);
    if (f->path == 0 && f->use_v3 && fdm_v3_supported(f, f->v3_S)) {
        const int np = fdm_v3_launch(f, 2, f->v3_S, out, in, pAp_part, skip, false, nrhs, V, pap_stride);
        return np;
    }
    int np = 0;
    for (int j = 0; j < nrhs; j++) fdm_mul_dev(f, SQ_OP_MTM, out + (size_t)j * V, in + (size_t)j * V, pAp_part + (size_t)j * pap_stride, &np, skip + j);
    return np;
}

bool fdm_cg_batch_applicable(const sq_fdm *f, const sq_kpm *kpm, int nrhs) {
    if (getenv("SQ_NO_BATCH_CG")) return false;
    const bool local = f->force_local || (f->world == 1 && !f->sharded);
    return nrhs > 1 && kpm != nullptr && kpm->active && local && f->slab_lo == 0 && f->slab_hi == (int)f->L;
}

void fdm_cg_batch_dev(sq_fdm *f, double2 *X, const double2 *B, int nrhs, bool zero_start, sq_kpm *kpm, double tol, i64 maxiter, i64 *iters,
                      double *eps) {
    SQ_REQUIRE(fdm_cg_batch_applicable(f, kpm, nrhs), "batched solve: needs an active preconditioner on one GPU and more than one right-hand side");
    const size_t V = (size_t)f->L * f->N;
    if (f->bt_cap < nrhs) {
        f->bt_r.alloc(V * nrhs, false); f->bt_p.alloc(V * nrhs, false); f->bt_z.alloc(V * nrhs, false); f->bt_q.alloc(V * nrhs, false);
        f->bt_zt.alloc(V * nrhs, false);
        f->bt_st.alloc(2 * (size_t)nrhs);
        f->bt_part.alloc((size_t)nrhs * 10 * SQ_MAXPART);
        f->bt_ticket.alloc(nrhs);
        f->bt_cap = nrhs;
    }
    cudaStream_t s = f->stream;
    const int TB = 256;
    const int G = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>((V + TB - 1) / TB, (size_t)f->num_sms * 4 / std::min(nrhs, 4) + 1), SQ_MAXPART));
    double2 *r = f->bt_r.p, *p = f->bt_p.p, *z = f->bt_z.p, *q = f->bt_q.p;
    double *part_b = f->bt_part.p, *part_rz = part_b + (size_t)nrhs * 3 * SQ_MAXPART, *part_pAp = part_rz + (size_t)nrhs * 3 * SQ_MAXPART,
           *part_rr = part_pAp + (size_t)nrhs * SQ_MAXPART, *part_dot = part_rr + (size_t)nrhs * SQ_MAXPART;     // part_dot: 2 SQ_MAXPART per system
    CgState *st = f->bt_st.p;
    SQ_CUDA(cudaMemsetAsync(f->bt_ticket.p, 0, nrhs * sizeof(unsigned), s));
    SQ_CUDA(cudaMemsetAsync(st, 0, 2 * (size_t)nrhs * sizeof(CgState), s));
    kb_dot_partials<<<dim3(G, nrhs), TB, 0, s>>>(B, B, V, part_b);
    if (zero_start) {
        SQ_CUDA(cudaMemcpyAsync(r, B, V * nrhs * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        SQ_CUDA(cudaMemsetAsync(X, 0, V * nrhs * sizeof(double2), s));
    } else {
        mul_batch(f, r, X, nrhs, V, nullptr, 0, st);                        // (states are all "not done" here)
        kb_residual<<<dim3(G, nrhs), TB, 0, s>>>(r, B, V);
    }
    int g = 0;
    kpm_fft_cheb_batch(kpm, z, r, f->bt_zt.p, nrhs, V, nullptr, nullptr, nullptr, &g);
    SQ_CUDA(cudaMemcpyAsync(p, z, V * nrhs * sizeof(double2), cudaMemcpyDeviceToDevice, s));
    kb_dot_partials<<<dim3(G, nrhs), TB, 0, s>>>(r, z, V, part_rz);
    kb_init<<<nrhs, 64, 0, s>>>(st, nrhs, part_b, part_rz, G, tol);
    SQ_LAUNCH_CHECK();
    f->launches += 5;
    std::vector<CgState> h(nrhs);
    i64 it = 0;
    int cur = 0;
    bool finished = maxiter <= 0;
    const int batch = 4;
    // (fusing the x / r update into the forward transform saves 3 - 5 us per iteration for ONE system -- cg.cu -- but measured 2 % slower for a
    //  batch, whose transform is bandwidth-bound: opt-in here)
    const bool fuse_update = getenv("SQ_BATCH_FFT_FUSION") != nullptr;
    while (!finished) {
        i64 step = batch;
        if (it == 0) step = std::max<i64>(batch, std::min<i64>((i64)(0.85 * f->prec_iters_hint[tol < 1e-7 ? 0 : 1]), 256));
        const i64 upto = std::min<i64>(maxiter, it + step);
        for (; it < upto;) {
            it++;
            CgState *sc = st + (size_t)cur * nrhs, *sn = st + (size_t)(cur ^ 1) * nrhs;
            const int np = mul_batch(f, q, p, nrhs, V, part_pAp, SQ_MAXPART, sc);
            if (fuse_update) {                                // x / r update + convergence test inside the forward transform's load phase
                FftCgUpdate U = {X, r, p, q, sc, sn, part_pAp, np, SQ_MAXPART, part_rr, f->bt_ticket.p, (int)it};
                kpm_fft_cheb_batch(kpm, z, r, f->bt_zt.p, nrhs, V, sn, r, part_dot, &g, &U);
            } else {
                kb_update_xr<<<dim3(G, nrhs), TB, 0, s>>>(sc, sn, X, r, p, q, V, part_pAp, np, SQ_MAXPART, part_rr, f->bt_ticket.p, (int)it);
                kpm_fft_cheb_batch(kpm, z, r, f->bt_zt.p, nrhs, V, sn, r, part_dot, &g);
                f->launches++;
            }
            kb_update_p<<<dim3(G, nrhs), TB, 0, s>>>(sn, sc, p, z, V, part_dot, g);
            f->launches++;
            cur ^= 1;
        }
        SQ_LAUNCH_CHECK();
        SQ_CUDA(cudaMemcpyAsync(h.data(), st + (size_t)cur * nrhs, nrhs * sizeof(CgState), cudaMemcpyDeviceToHost, s));
        SQ_CUDA(cudaStreamSynchronize(s));
        bool all = true;
        for (int j = 0; j < nrhs; j++) all = all && h[j].done != 0;
        if (all || it >= maxiter) finished = true;
    }
    if (maxiter <= 0) {
        SQ_CUDA(cudaMemcpyAsync(h.data(), st, nrhs * sizeof(CgState), cudaMemcpyDeviceToHost, s));
        SQ_CUDA(cudaStreamSynchronize(s));
    }
    i64 most = 0;
    for (int j = 0; j < nrhs; j++) {
        if (h[j].done == 2) throw SqNumericalInstability("conjugate gradient (batched): NaN encountered in the residual (numerical instability)");
        iters[j] = h[j].done ? h[j].iters : maxiter;
        eps[j] = h[j].eps;
        most = std::max(most, iters[j]);
        f->stats[SQ_STAT_CG_ITERS] += iters[j];
    }
    f->stats[SQ_STAT_CG_SOLVES] += nrhs;
    f->stats[SQ_STAT_CG_PREC] += nrhs;
    f->stats[SQ_STAT_CG_BATCHED] += nrhs;
    f->prec_iters_hint[tol < 1e-7 ? 0 : 1] = (int)std::min<i64>(most, 1 << 20);
}

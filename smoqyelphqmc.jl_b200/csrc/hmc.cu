// hmc.cu -- K8: the EFA-PFF-HMC trajectory, device resident.
//
// Replaces hmc_update! (src/EFAPFFHMCUpdater.jl:102-279) and the SmoQyDQMC pieces it calls
// [unvendored: ExactFourierAccelerator, initialize_momentum!, evolve_eom!, kinetic_energy, bosonic_action,
// eval_derivative_anharmonic_action!], restated from arXiv:2404.09723 exactly as oracle/ref_c.c does.
// x, p, dS/dx, the pseudofermion fields and the operator stay in HBM for the whole trajectory; the host only
// sees the CG convergence flags and five scalars.  x and p are transformed together as one complex field
// z = x + i p (one tau-FFT instead of two); the mode rotation separates them through the Hermitian partner.
#include "sq_internal.h"

#include <cmath>

double pff_sample_dev(sq_pff *q);
double pff_action_dev(sq_pff *q, sq_kpm *kpm, bool refresh, const double *h_lanczos, const double *d_lanczos, double tol, i64 maxiter,
                      i64 *iters, double *eps);
double pff_force_dev(sq_pff *q, sq_kpm *kpm, bool refresh, const double *h_lanczos, const double *d_lanczos, double tol, i64 maxiter,
                     i64 *iters, double *eps);
void pff_fill_phi_normals(sq_pff *q, const void *h_R, const double *d_stream);

__global__ void k_pack(double2 *__restrict__ z, const double *__restrict__ a, const double *__restrict__ b, size_t n) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        z[k] = make_double2(a[k], b ? b[k] : 0.0);
}
__global__ void k_unpack(double *__restrict__ a, double *__restrict__ b, const double2 *__restrict__ z, const int *__restrict__ fin, int Nph,
                         size_t n) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        if (fin && !fin[k % Nph]) continue;            // frozen (M = inf) modes keep x, p
        double2 v = z[k];
        if (a) a[k] = v.x;
        if (b) b[k] = v.y;
    }
}
// exact flow of every Matsubara mode: zt = FFT(x + i p) (unitary); xt = (zt[w] + conj(zt[-w]))/2 etc.
__global__ void k_efa_rotate(double2 *__restrict__ out, const double2 *__restrict__ zt, const double *__restrict__ Mt,
                             const double *__restrict__ wd, const int *__restrict__ fin, int L, int Nph, double dt) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)L * Nph) return;
    int w = (int)(idx / Nph), p = (int)(idx - (size_t)w * Nph);
    int wm = (w == 0) ? 0 : L - w;
    double2 a = zt[idx], b = zt[(size_t)wm * Nph + p];
    double2 X = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));      // x~[w]
    double2 P = make_double2(0.5 * (a.y + b.y), -0.5 * (a.x - b.x));     // p~[w] = (a - conj(b)) / (2i)
    if (fin[p]) {
        double om = wd[idx], mt = Mt[idx];
        double2 Xn, Pn;
        if (om > 0) {
            double s, c;
            sincos(om * dt, &s, &c);
            double k1 = s / (mt * om), k2 = mt * om * s;
            Xn = make_double2(c * X.x + k1 * P.x, c * X.y + k1 * P.y);
            Pn = make_double2(c * P.x - k2 * X.x, c * P.y - k2 * X.y);
        } else {
            Xn = make_double2(X.x + dt / mt * P.x, X.y + dt / mt * P.y);
            Pn = P;
        }
        X = Xn; P = Pn;
    }
    out[idx] = make_double2(X.x - P.y, X.y + P.x);                       // x~ + i p~
}
// momentum refresh in Fourier space: p~ = sqrt(Mt) R~ ; K partial = |p~|^2 / (2 Mt)
__global__ void k_efa_momentum(double2 *__restrict__ zt, const double *__restrict__ Mt, const int *__restrict__ fin, int L, int Nph,
                               int scale, double *__restrict__ part) {
    __shared__ double red[32];
    double acc = 0;
    size_t tot = (size_t)L * Nph;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < tot; k += (size_t)gridDim.x * blockDim.x) {
        int p = (int)(k % Nph);
        double2 v = zt[k];
        if (fin[p]) {
            double mt = Mt[k];
            if (scale) { double s = sqrt(mt); v = make_double2(s * v.x, s * v.y); }
            acc += (v.x * v.x + v.y * v.y) / (2 * mt);
        } else v = make_double2(0, 0);
        if (scale) zt[k] = v;
    }
    double t[1] = {acc};
    block_sum<1>(t, red);
    if (threadIdx.x == 0) part[blockIdx.x] = t[0];
}
static inline unsigned nblk(size_t n) { return (unsigned)((n + 255) / 256); }
static int red_grid(const sq_fdm *f) { return std::min(SQ_MAXPART, f->num_sms * 4); }

void hmc_create_impl(sq_hmc **out, sq_pff *q, i64 Nt, double dt, double eta, double delta, uint64_t seed) {
    SQ_REQUIRE(out && q, "NULL argument");
    SQ_REQUIRE(Nt >= 1, "Nt must be positive");
    sq_elph *e = q->e;
    sq_fdm *f = e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    sq_hmc *h = new sq_hmc();
    try {
        h->p = q; h->owner = f; h->Nt = Nt; h->dt = dt; h->eta = eta; h->delta = delta; h->seed = seed;
        if (!q->seed_set) q->seed = sq_mix_seed(seed, SQ_RNG_PFF);    // chains with different HMC seeds draw different pseudofermion noise in the global moves
        i64 L = f->L, Nph = e->Nph;
        size_t nx = (size_t)L * Nph;
        h->x0.alloc(nx); h->pm.alloc(nx); h->dS.alloc(nx);
        h->fz.alloc(nx); h->fw.alloc(nx);
        h->part.alloc(SQ_MAXPART);
        std::vector<double> Mt(nx), wd(nx), Om(Nph);
        e->Om.download(Om.data(), Nph, f->stream);
        SQ_CUDA(cudaStreamSynchronize(f->stream));
        const double PI = 3.14159265358979323846;
        for (i64 w = 0; w < L; w++) {
            double s = sin(PI * (double)w / (double)L);
            for (i64 p = 0; p < Nph; p++) {
                double M = e->h_M[p];
                double Ot2 = Om[p] * Om[p] + 4 * s * s / (e->dtau * e->dtau);
                double k = e->dtau * M * Ot2, mt = e->dtau * M * (Ot2 + eta * eta);
                Mt[p + w * Nph] = mt;
                wd[p + w * Nph] = std::isfinite(M) ? sqrt(k / mt) : 0.0;
            }
        }
        h->Mt.alloc(nx, false); h->Mt.upload(Mt.data(), nx, f->stream);
        h->wd.alloc(nx, false); h->wd.upload(wd.data(), nx, f->stream);
        std::vector<double2> tw;
        fft_make_twiddles(L, tw);
        h->tw.alloc(L, false); h->tw.upload(tw.data(), L, f->stream);
        SQ_CUDA(cudaStreamSynchronize(f->stream));
        fft_radices(L, h->radices);
    } catch (...) {
        delete h;
        throw;
    }
    *out = h;
}

static void efa_fft(sq_hmc *h, double2 *out, const double2 *in, bool inverse) {
    sq_fdm *f = h->p->e->f;
    tau_fft_launch(f->stream, h->radices, (int)f->L, (int)h->p->e->Nph, out, in, inverse, false, h->tw.p, nullptr, nullptr, nullptr, nullptr,
                   nullptr, f->smem_optin);
    f->launches++;
}

// evolve_eom!(x, p, dt, efa) on device arrays [l][p]
void hmc_evolve_dev(sq_hmc *h, double *x, double *pm, double dt) {
    sq_elph *e = h->p->e;
    sq_fdm *f = e->f;
    size_t nx = (size_t)f->L * e->Nph;
    int g = red_grid(f);
    k_pack<<<g, 256, 0, f->stream>>>(h->fz.p, x, pm, nx);
    efa_fft(h, h->fw.p, h->fz.p, false);
    k_efa_rotate<<<nblk(nx), 256, 0, f->stream>>>(h->fz.p, h->fw.p, h->Mt.p, h->wd.p, e->fin.p, (int)f->L, (int)e->Nph, dt);
    efa_fft(h, h->fw.p, h->fz.p, true);
    k_unpack<<<g, 256, 0, f->stream>>>(x, pm, h->fw.p, e->fin.p, (int)e->Nph, nx);
    SQ_LAUNCH_CHECK();
    f->launches += 3;
}
// initialize_momentum!: pm = F^-1 sqrt(Mt) F R ; returns K.  R (device, [l][p]) holds N(0,1) reals.
double hmc_init_momentum_dev(sq_hmc *h, const double *R, double *pm) {
    sq_elph *e = h->p->e;
    sq_fdm *f = e->f;
    size_t nx = (size_t)f->L * e->Nph;
    int g = red_grid(f);
    k_pack<<<g, 256, 0, f->stream>>>(h->fz.p, R, nullptr, nx);
    efa_fft(h, h->fw.p, h->fz.p, false);
    k_efa_momentum<<<g, 256, 0, f->stream>>>(h->fw.p, h->Mt.p, e->fin.p, (int)f->L, (int)e->Nph, 1, h->part.p);
    efa_fft(h, h->fz.p, h->fw.p, true);
    k_unpack<<<g, 256, 0, f->stream>>>(pm, nullptr, h->fz.p, nullptr, (int)e->Nph, nx);
    SQ_LAUNCH_CHECK();
    f->launches += 3;
    return reduce_partials_host(f, h->part.p, g);
}
double hmc_kinetic_dev(sq_hmc *h, const double *pm) {
    sq_elph *e = h->p->e;
    sq_fdm *f = e->f;
    size_t nx = (size_t)f->L * e->Nph;
    int g = red_grid(f);
    k_pack<<<g, 256, 0, f->stream>>>(h->fz.p, pm, nullptr, nx);
    efa_fft(h, h->fw.p, h->fz.p, false);
    k_efa_momentum<<<g, 256, 0, f->stream>>>(h->fw.p, h->Mt.p, e->fin.p, (int)f->L, (int)e->Nph, 0, h->part.p);
    SQ_LAUNCH_CHECK();
    f->launches += 2;
    return reduce_partials_host(f, h->part.p, g);
}

// hmc_update!: the random stream (host-supplied or Philox) is consumed in the order documented in
// include/smoqyelph_b200.h / DESIGN.md: u_dt, Phi normals (2 L N), momentum normals (Nph L),
// (Nt+1) Lanczos start vectors (N each, only with a preconditioner), u_accept.
int hmc_update_impl(sq_hmc *h, sq_kpm *kpm, double tol_action, double tol_force, i64 maxiter, const double *randoms, i64 nrandoms,
                    double *info) {
    sq_pff *q = h->p;
    sq_elph *e = q->e;
    sq_fdm *f = e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    const i64 L = f->L, N = f->N, Nph = e->Nph, Nt = h->Nt;
    const size_t nx = (size_t)L * Nph, V = (size_t)L * N;
    const i64 need = 1 + 2 * (i64)V + (i64)nx + (kpm ? (Nt + 1) * N : 0) + 1;
    if ((size_t)need > h->rnd.n) h->rnd.alloc(need, false);
    double u_dt, u_acc;
    if (randoms) {
        SQ_REQUIRE(nrandoms >= need, "random stream too short for one trajectory");
        h->rnd.upload(randoms, need, f->stream);
        u_dt = randoms[0];
        u_acc = randoms[need - 1];
    } else {
        // Philox: normals everywhere, the two uniforms separately
        rng_fill_normal(h->rnd.p, need, h->seed, sq_rng_stream(SQ_RNG_HMC, 2 * h->counter), f->stream);
        double uu[2];
        rng_fill_uniform(h->part.p, 2, h->seed, sq_rng_stream(SQ_RNG_HMC, 2 * h->counter + 1), f->stream);
        SQ_CUDA(cudaMemcpyAsync(uu, h->part.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
        SQ_CUDA(cudaStreamSynchronize(f->stream));
        u_dt = uu[0];
        u_acc = uu[1];
        h->counter++;
    }
    const double *r_phi = h->rnd.p + 1, *r_p = r_phi + 2 * V, *r_lan = r_p + nx;
    double dt = h->dt * (1.0 + (2 * u_dt - 1) * h->delta);                                   // :125
    SQ_CUDA(cudaMemcpyAsync(h->x0.p, e->x.p, nx * sizeof(double), cudaMemcpyDeviceToDevice, f->stream));   // :128
    bool stable = true;
    h->last_reject.clear();
    double Sf0 = 0, Sb0 = 0, K0 = 0, Sf1 = 0, Sb1 = 0, K1 = 0, iters_avg = 0, dH = 0;
    int solve = 0;
    try {
        pff_fill_phi_normals(q, nullptr, r_phi);
        Sf0 = pff_sample_dev(q);                                                            // :131
        Sb0 = elph_bosonic_action(e);                                                       // :136
        K0 = hmc_init_momentum_dev(h, r_p, h->pm.p);                                        // :142
        hmc_evolve_dev(h, e->x.p, h->pm.p, dt / 2);                                         // :149-150
        elph_refresh_fdm(e);                                                                // :152-153
        for (i64 t = 1; t <= Nt; t++) {                                                     // :162
            SQ_CUDA(cudaMemsetAsync(q->F.p, 0, nx * sizeof(double), f->stream));            // :165
            i64 it = 0;
            double eps = 0;
            pff_force_dev(q, kpm, kpm != nullptr, nullptr, kpm ? r_lan + (size_t)solve * N : nullptr, tol_force, maxiter, &it, &eps);   // :171
            solve++;
            iters_avg += (double)it / (double)(Nt + 1);
            elph_add_potential_derivative(e, h->pm.p, q->F.p, dt);                          // :190-196 (anharmonic + dispersive + kick)
            hmc_evolve_dev(h, e->x.p, h->pm.p, t == Nt ? dt / 2 : dt);                      // :200-203
            elph_refresh_fdm(e);                                                            // :204-205
        }
        i64 it = 0;
        double eps = 0;
        Sf1 = pff_action_dev(q, kpm, kpm != nullptr, nullptr, kpm ? r_lan + (size_t)solve * N : nullptr, tol_action, maxiter, &it, &eps);   // :217
        iters_avg += (double)it / (double)(Nt + 1);
        Sb1 = elph_bosonic_action(e);                                                       // :238
        K1 = hmc_kinetic_dev(h, h->pm.p);                                                   // :244
    } catch (const SqNumericalInstability &err) {
        // numerical instability inside the trajectory => warn and reject (EFAPFFHMCUpdater.jl:168-187,215-231: `@warn ... rejecting
        // update`).  Only this class is caught: bad arguments, CUDA / NCCL errors and watchdog time-outs propagate to the caller.
        stable = false;
        f->stats[SQ_STAT_INSTABILITY]++;
        h->last_reject = err.what();
        fprintf(stderr, "[smoqyelph_b200] warning: numerical instability in the EFA-PFF-HMC trajectory (solve %d), rejecting update: %s\n", solve, err.what());
    }
    double Pacc = 0.0;
    if (stable) {
        dH = (Sf1 + Sb1 + K1) - (Sf0 + Sb0 + K0);                                           // :246-250
        Pacc = std::isfinite(dH) ? std::min(1.0, exp(-dH)) : 0.0;                           // :253
    }
    int accepted = (u_acc < Pacc) ? 1 : 0;                                                  // :263
    if (!accepted) {                                                                        // :266-276
        SQ_CUDA(cudaMemcpyAsync(e->x.p, h->x0.p, nx * sizeof(double), cudaMemcpyDeviceToDevice, f->stream));
        elph_refresh_fdm(e);
    }
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    if (info) {
        info[0] = iters_avg; info[1] = dH; info[2] = Sf0; info[3] = Sf1; info[4] = Sb0; info[5] = Sb1; info[6] = K0; info[7] = K1;
    }
    return accepted;
}

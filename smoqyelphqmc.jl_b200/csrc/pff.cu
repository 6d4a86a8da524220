// pff.cu -- pseudofermion field calculator and K6, the fermionic force.
//
// Replaces src/PFFCalculator.jl:56-158 (sample_pseudofermion_fields!, calculate_fermionic_action!,
// calculate_derivative_fermionic_action!), src/fermion_det_matrix_dervative.jl:2-290 (mul_νRe∂M∂x! and
// its bond / site contractions) and mul_νRe∂Λ∂x! (src/holstein_shift_matrix.jl:156-201).
//
// Force layout: every coupling writes its contribution to a per-coupling array ([l][c]) exactly once per
// peel phase; one gather kernel then sums each phonon's contributions in a fixed order.  No floating-point
// atomics => bit-reproducible forces.  v' = sigma B w[l-1] is obtained from the fused matvec as M w - w.
#include "sq_internal.h"

#include <cmath>

struct ForceDev {
    int L, N, Nh, Nph, Nhol, Nssh;
    double dtau;
    const double *x;
    const int *hol_ph, *hol_site, *hol_sym, *fin;
    const double *ha;
    const int *ssh_p, *ssh_pp, *bond_ptr, *bond_cpl;
    const double *sa;
    const int2 *nt;
};

// out = a - b
__global__ void k_vec_sub(double2 *__restrict__ out, const double2 *__restrict__ a, const double2 *__restrict__ b, size_t n) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        out[k] = csub(a[k], b[k]);
}
// Re/Im of conj(a).b partials (slots 0, 1) and |a|^2 (slot 2)
__global__ void k_dot3(const double2 *__restrict__ a, const double2 *__restrict__ b, size_t n, double *__restrict__ part) {
    __shared__ double red[3 * 32];
    double v[3] = {0, 0, 0};
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        double2 x = a[k], y = b[k];
        v[0] += x.x * y.x + x.y * y.y;
        v[1] += x.x * y.y - x.y * y.x;
        v[2] += x.x * x.x + x.y * x.y;
    }
    block_sum<3>(v, red);
    if (threadIdx.x == 0) { part[blockIdx.x] = v[0]; part[SQ_MAXPART + blockIdx.x] = v[1]; part[2 * SQ_MAXPART + blockIdx.x] = v[2]; }
}
// randn!(rng, Phi): complex normals with variance 1/2 per component from a stream of N(0,1) reals laid out
// in the HOST order (re, im interleaved, tau fastest): element (l, i) at 2*(l + i*L)
__global__ void k_complex_normals_from_host_stream(double2 *__restrict__ out, const double *__restrict__ rnd, int L, int N) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)L * N) return;
    int l = (int)(idx / N), i = (int)(idx - (size_t)l * N);
    size_t h = 2 * ((size_t)l + (size_t)i * L);
    const double s = 0.70710678118654752440;
    out[idx] = make_double2(rnd[h] * s, rnd[h + 1] * s);
}
__global__ void k_scale_to_complex_normals(double2 *__restrict__ v, size_t n) {
    const double s = 0.70710678118654752440;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        v[k] = make_double2(v[k].x * s, v[k].y * s);
}

// _mul_νReΔτ∂V∂x!: fermion_det_matrix_dervative.jl:258-290 -> HV[l][c] = coef Re(conj(u')[l,i] g v'[l,i])
__global__ void k_hol_contract(const __grid_constant__ ForceDev E, double *__restrict__ HV, double coef, const double2 *__restrict__ up,
                               const double2 *__restrict__ vp) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)E.L * E.Nhol) return;
    int l = (int)(idx / E.Nhol), c = (int)(idx - (size_t)l * E.Nhol);
    int p = E.hol_ph[c], i = E.hol_site[c];
    double val = 0.0;
    if (E.fin[p]) {
        double x = E.x[(size_t)l * E.Nph + p];
        double g = E.dtau * (E.ha[c] + x * (2 * E.ha[E.Nhol + c] + x * (3 * E.ha[2 * E.Nhol + c] + x * 4 * E.ha[3 * E.Nhol + c])));
        double2 a = up[(size_t)l * E.N + i], b = vp[(size_t)l * E.N + i];
        val = coef * g * (a.x * b.x + a.y * b.y);
    }
    HV[idx] = val;
}
// _mul_νReΔτ∂Kc∂x!: :196-254, bonds [lo, lo+nb) of one colour -> SV[l][c] += val (each coupling belongs to one colour)
__global__ void k_ssh_contract(const __grid_constant__ ForceDev E, double *__restrict__ SV, double coef, double dt, const double2 *__restrict__ up,
                               const double2 *__restrict__ vp, int lo, int nb) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)E.L * nb) return;
    int l = (int)(idx / nb), n = lo + (int)(idx - (size_t)l * nb);
    int q0 = E.bond_ptr[n], q1 = E.bond_ptr[n + 1];
    if (q0 == q1) return;
    int2 ij = E.nt[n];
    double2 ui = up[(size_t)l * E.N + ij.x], uj = up[(size_t)l * E.N + ij.y];
    double2 vi = vp[(size_t)l * E.N + ij.x], vj = vp[(size_t)l * E.N + ij.y];
    double re = (uj.x * vi.x + uj.y * vi.y) + (ui.x * vj.x + ui.y * vj.y);       // Re(conj(u'_j) v'_i + conj(u'_i) v'_j), real g
    for (int q = q0; q < q1; q++) {
        int c = E.bond_cpl[q];
        double dx = E.x[(size_t)l * E.Nph + E.ssh_pp[c]] - E.x[(size_t)l * E.Nph + E.ssh_p[c]];
        double g = dt * (E.sa[c] + dx * (2 * E.sa[E.Nssh + c] + dx * (3 * E.sa[2 * E.Nssh + c] + dx * 4 * E.sa[3 * E.Nssh + c])));
        SV[(size_t)l * E.Nssh + c] += coef * g * re;
    }
}
// mul_νRe∂Λ∂x!: holstein_shift_matrix.jl:156-201 -> HL[l][c]
__global__ void k_dlambda_contract(const __grid_constant__ ForceDev E, double *__restrict__ HL, double nu, const double2 *__restrict__ up,
                                   const double2 *__restrict__ u, const double *__restrict__ Lam) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)E.L * E.Nhol) return;
    int l = (int)(idx / E.Nhol), c = (int)(idx - (size_t)l * E.Nhol);
    double val = 0.0;
    if (E.hol_sym[c]) {
        int p = E.hol_ph[c], i = E.hol_site[c];
        double x = E.x[(size_t)l * E.Nph + p];
        double d = E.dtau * (E.ha[c] + 3 * E.ha[2 * E.Nhol + c] * x * x) / 2 * Lam[(size_t)l * E.N + i];   // :192
        int lm = (l == 0) ? E.L - 1 : l - 1;
        double2 a = up[(size_t)lm * E.N + i], b = u[(size_t)l * E.N + i];
        val = nu * d * (a.x * b.x + a.y * b.y);                                                            // :193
    }
    HL[idx] = val;
}
// F[l][p] += sum over the phonon's couplings, fixed order
__global__ void k_gather_force(double *__restrict__ F, const double *__restrict__ HV, const double *__restrict__ HL, const double *__restrict__ SV,
                               const int *__restrict__ hptr, const int *__restrict__ hcpl, const int *__restrict__ sptr,
                               const int *__restrict__ scpl, const int *__restrict__ fin, int L, int Nph, int Nhol, int Nssh,
                               int use_hv, int use_hl, int use_sv) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)L * Nph) return;
    int l = (int)(idx / Nph), p = (int)(idx - (size_t)l * Nph);
    double acc = 0.0;
    for (int q = hptr[p]; q < hptr[p + 1]; q++) {
        int c = hcpl[q];
        if (use_hv) acc += HV[(size_t)l * Nhol + c];
        if (use_hl) acc += HL[(size_t)l * Nhol + c];
    }
    if (use_sv && fin[p]) {
        for (int q = sptr[p]; q < sptr[p + 1]; q++) {
            int sc = scpl[q];
            acc += (sc > 0) ? SV[(size_t)l * Nssh + (sc - 1)] : -SV[(size_t)l * Nssh + (-sc - 1)];
        }
    }
    F[idx] += acc;
}

// ---------------------------------------------------------------------------------------------------
static ForceDev force_dev(const sq_pff *q) {
    const sq_elph *e = q->e;
    const sq_fdm *f = e->f;
    ForceDev E;
    E.L = (int)f->L; E.N = (int)f->N; E.Nh = (int)f->Nh; E.Nph = (int)e->Nph; E.Nhol = (int)e->Nhol; E.Nssh = (int)e->Nssh;
    E.dtau = e->dtau; E.x = e->x.p; E.hol_ph = e->hol_ph.p; E.hol_site = e->hol_site.p; E.hol_sym = e->hol_sym.p; E.fin = e->fin.p;
    E.ha = e->ha.p; E.ssh_p = e->ssh_p.p; E.ssh_pp = e->ssh_pp.p; E.bond_ptr = e->bond_ptr.p; E.bond_cpl = e->bond_cpl.p; E.sa = e->sa.p;
    E.nt = f->nt.p;
    return E;
}
static inline unsigned nblk(size_t n) { return (unsigned)((n + 255) / 256); }
static int red_grid(const sq_fdm *f) { return std::min(SQ_MAXPART, f->num_sms * 4); }

void pff_create_impl(sq_pff **out, sq_elph *e) {
    SQ_REQUIRE(out && e, "NULL argument");
    sq_fdm *f = e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    sq_pff *q = new sq_pff();
    try {
        q->e = e;
        q->owner = f;
        size_t V = (size_t)f->L * f->N;
        q->Phi.alloc(V); q->u.alloc(V); q->up.alloc(V); q->upp.alloc(V); q->w1.alloc(V); q->w2.alloc(V);
        q->Lam.alloc(V);
        q->F.alloc((size_t)f->L * e->Nph);
        q->HV.alloc((size_t)f->L * e->Nhol + 1); q->HL.alloc((size_t)f->L * e->Nhol + 1); q->SV.alloc((size_t)f->L * e->Nssh + 1);
        q->part.alloc(3 * SQ_MAXPART);
    } catch (...) {
        delete q;
        throw;
    }
    *out = q;
}

// Sf = |R|^2 ; Phi = Lambda^T M^T R  (PFFCalculator.jl:56-76).  d_R: complex normals already in q->Phi when NULL.
double pff_sample_dev(sq_pff *q) {
    sq_elph *e = q->e;
    sq_fdm *f = e->f;
    size_t V = (size_t)f->L * f->N;
    elph_update_lambda(e, q->Lam.p);
    int g = red_grid(f);
    k_dot3<<<g, 256, 0, f->stream>>>(q->Phi.p, q->Phi.p, V, q->part.p);
    SQ_LAUNCH_CHECK();
    fdm_mul_dev(f, SQ_OP_MT, q->w1.p, q->Phi.p);                         // lmul_Mt!
    elph_lambda_op(e, 2, q->Phi.p, q->w1.p, q->Lam.p);                   // mul_Λᵀ!
    f->launches++;
    return reduce_partials_host(f, q->part.p + 2 * SQ_MAXPART, g);
}

// calculate_fermionic_action!: PFFCalculator.jl:79-116
double pff_action_dev(sq_pff *q, sq_kpm *kpm, bool refresh, const double *h_lanczos, const double *d_lanczos, double tol, i64 maxiter,
                      i64 *iters, double *eps) {
    sq_elph *e = q->e;
    sq_fdm *f = e->f;
    size_t V = (size_t)f->L * f->N;
    elph_update_lambda(e, q->Lam.p);                                      // :94
    elph_lambda_op(e, 3, q->u.p, q->Phi.p, q->Lam.p);                    // Psi = Λ⁻ᵀ Phi :97
    if (kpm && refresh) kpm_update(kpm, h_lanczos, d_lanczos);            // FermionDetMatrix.jl:259
    fdm_cg_dev(f, q->u.p, q->u.p, true, kpm, tol, maxiter, iters, eps);  // :99-105 (x === b)
    elph_lambda_op(e, 1, q->w1.p, q->u.p, q->Lam.p);                     // Psi = Λ⁻¹ Psi :107
    SQ_CUDA(cudaMemcpyAsync(q->u.p, q->w1.p, V * sizeof(double2), cudaMemcpyDeviceToDevice, f->stream));
    int g = red_grid(f);
    k_dot3<<<g, 256, 0, f->stream>>>(q->Phi.p, q->u.p, V, q->part.p);    // Sf = dot(Phi, Psi) :109
    SQ_LAUNCH_CHECK();
    f->launches++;
    double Sf = reduce_partials_host(f, q->part.p, g);
    if (!(Sf == Sf) || std::isinf(Sf)) throw SqNumericalInstability("fermionic action is not finite (numerical instability)");
    return Sf;
}

// nu Re<u| dM/dx |v> into the per-coupling arrays (mul_νRe∂M∂x!).  vp_in: v' = M v - v if already available (else NULL).
static void pff_dM_dx_dev(sq_pff *q, double nu, const double2 *u, const double2 *v, const double2 *Mv) {
    sq_elph *e = q->e;
    sq_fdm *f = e->f;
    const size_t V = (size_t)f->L * f->N;
    const int C = (int)f->C, L = (int)f->L;
    ForceDev E = force_dev(q);
    double2 *up = q->w1.p, *vp = q->w2.p;
    int g = red_grid(f);
    if (!Mv) { fdm_mul_dev(f, SQ_OP_M, vp, v); Mv = vp; }
    k_vec_sub<<<g, 256, 0, f->stream>>>(vp, Mv, v, V);                    // v'[l] = sigma_l B_l v[l-1] = (M v - v)[l]   (:24-36)
    SQ_CUDA(cudaMemcpyAsync(up, u, V * sizeof(double2), cudaMemcpyDeviceToDevice, f->stream));   // :39
    f->launches++;
    if (e->Nssh > 0) SQ_CUDA(cudaMemsetAsync(q->SV.p, 0, (size_t)L * e->Nssh * sizeof(double), f->stream));
    auto ssh = [&](int c, double dt) {
        int lo = f->clo[c], nb = f->chi[c] - lo;
        if (nb <= 0) return;
        k_ssh_contract<<<nblk((size_t)L * nb), 256, 0, f->stream>>>(E, q->SV.p, -nu, dt, up, vp, lo, nb);
        f->launches++;
    };
    auto hol = [&]() {
        if (e->Nhol > 0) {
            k_hol_contract<<<nblk((size_t)L * e->Nhol), 256, 0, f->stream>>>(E, q->HV.p, -nu, up, vp);
            f->launches++;
        }
    };
    if (f->sym) {
        if (e->Nssh > 0) {
            for (int c = C - 1; c >= 0; c--) {                                    // :50-63
                ssh(c, e->dtau / 2);
                fdm_sweep_global(f, up, f->clo[c], f->chi[c], false);
                fdm_sweep_global(f, vp, f->clo[c], f->chi[c], true);
            }
        } else {
            for (int c = C - 1; c >= 0; c--) fdm_sweep_global(f, up, f->clo[c], f->chi[c], false);   // u' <- Gamma^T u'  (:66-69)
            if (q->exact_holstein) {                                              // Gamma^-1: inverse factors, colours C..1
                for (int c = C - 1; c >= 0; c--) fdm_sweep_global(f, vp, f->clo[c], f->chi[c], true);
            } else {                                                              // reference: Gamma^-T (:71-74)
                for (int c = 0; c < C; c++) fdm_sweep_global(f, vp, f->clo[c], f->chi[c], true);
            }
        }
        hol();                                                                    // :82
        if (e->Nssh > 0) {
            fdm_scale_global(f, up, false);                                       // :87
            fdm_scale_global(f, vp, true);                                        // :90
            for (int c = 0; c < C; c++) {                                         // :95-111
                ssh(c, e->dtau / 2);
                fdm_sweep_global(f, up, f->clo[c], f->chi[c], false);
                fdm_sweep_global(f, vp, f->clo[c], f->chi[c], true);
            }
        }
    } else {
        hol();                                                                    // :158
        if (e->Nssh > 0) {
            fdm_scale_global(f, up, false);                                       // :166
            fdm_scale_global(f, vp, true);                                        // :169
            for (int c = C - 1; c >= 0; c--) {                                    // :172-187
                ssh(c, e->dtau);
                fdm_sweep_global(f, up, f->clo[c], f->chi[c], false);
                fdm_sweep_global(f, vp, f->clo[c], f->chi[c], true);
            }
        }
    }
    SQ_LAUNCH_CHECK();
}

static void pff_gather(sq_pff *q, bool hv, bool hl, bool sv) {
    sq_elph *e = q->e;
    sq_fdm *f = e->f;
    k_gather_force<<<nblk((size_t)f->L * e->Nph), 256, 0, f->stream>>>(q->F.p, q->HV.p, q->HL.p, q->SV.p, e->ph_hol_ptr.p, e->ph_hol_cpl.p,
                                                                       e->ph_ssh_ptr.p, e->ph_ssh_cpl.p, e->fin.p, (int)f->L, (int)e->Nph,
                                                                       (int)e->Nhol, (int)e->Nssh, hv && e->Nhol > 0, hl && e->Nhol > 0,
                                                                       sv && e->Nssh > 0);
    SQ_LAUNCH_CHECK();
    f->launches++;
}

// calculate_derivative_fermionic_action!: PFFCalculator.jl:119-158.  q->F (device, [l][p]) += dSf/dx
double pff_force_dev(sq_pff *q, sq_kpm *kpm, bool refresh, const double *h_lanczos, const double *d_lanczos, double tol, i64 maxiter,
                     i64 *iters, double *eps) {
    sq_elph *e = q->e;
    sq_fdm *f = e->f;
    double Sf = pff_action_dev(q, kpm, refresh, h_lanczos, d_lanczos, tol, maxiter, iters, eps);
    double2 *Psi = q->u.p, *LPsi = q->up.p, *APsi = q->upp.p;
    elph_lambda_op(e, 0, LPsi, Psi, q->Lam.p);                            // ΛΨ :146
    fdm_mul_dev(f, SQ_OP_M, APsi, LPsi);                                  // AΨ = M ΛΨ :148
    pff_dM_dx_dev(q, -2.0, APsi, LPsi, APsi);                             // :150  (M v is APsi itself)
    bool hl = e->any_phsym;
    if (hl) {
        ForceDev E = force_dev(q);
        fdm_mul_dev(f, SQ_OP_MT, q->w1.p, APsi);                          // MᵀAΨ :153
        k_dlambda_contract<<<nblk((size_t)f->L * e->Nhol), 256, 0, f->stream>>>(E, q->HL.p, -2.0, q->w1.p, Psi, q->Lam.p);   // :155
        SQ_LAUNCH_CHECK();
        f->launches++;
    }
    pff_gather(q, true, hl, true);
    return Sf;
}

// ---- ABI-level helpers on host arrays ----------------------------------------------------------------
void pff_dM_dx_host(sq_pff *q, double *F, double nu, const void *u, const void *v) {
    sq_elph *e = q->e;
    sq_fdm *f = e->f;
    size_t nF = (size_t)f->L * e->Nph;
    fdm_host_to_dev(f, q->up.p, u);
    fdm_host_to_dev(f, q->upp.p, v);
    SQ_CUDA(cudaMemsetAsync(q->F.p, 0, nF * sizeof(double), f->stream));
    pff_dM_dx_dev(q, nu, q->up.p, q->upp.p, nullptr);
    pff_gather(q, true, false, true);
    q->F.download(F, nF, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
}
void pff_dLambda_dx_host(sq_pff *q, double *F, double nu, const void *upv, const void *uv) {
    sq_elph *e = q->e;
    sq_fdm *f = e->f;
    size_t nF = (size_t)f->L * e->Nph;
    fdm_host_to_dev(f, q->up.p, upv);
    fdm_host_to_dev(f, q->upp.p, uv);
    SQ_CUDA(cudaMemsetAsync(q->F.p, 0, nF * sizeof(double), f->stream));
    elph_update_lambda(e, q->Lam.p);
    if (e->Nhol > 0) {
        ForceDev E = force_dev(q);
        k_dlambda_contract<<<nblk((size_t)f->L * e->Nhol), 256, 0, f->stream>>>(E, q->HL.p, nu, q->up.p, q->upp.p, q->Lam.p);
        SQ_LAUNCH_CHECK();
        pff_gather(q, false, true, false);
    }
    q->F.download(F, nF, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
}
void pff_fill_phi_normals(sq_pff *q, const void *h_R, const double *d_stream) {
    sq_fdm *f = q->e->f;
    size_t V = (size_t)f->L * f->N;
    if (h_R) fdm_host_to_dev(f, q->Phi.p, h_R);
    else if (d_stream) {
        k_complex_normals_from_host_stream<<<nblk(V), 256, 0, f->stream>>>(q->Phi.p, d_stream, (int)f->L, (int)f->N);
        SQ_LAUNCH_CHECK();
    } else {
        rng_fill_normal((double *)q->Phi.p, 2 * V, q->seed, sq_rng_stream(SQ_RNG_PFF, q->rng_counter++), f->stream);
        k_scale_to_complex_normals<<<red_grid(f), 256, 0, f->stream>>>(q->Phi.p, V);
        SQ_LAUNCH_CHECK();
    }
    f->launches++;
}

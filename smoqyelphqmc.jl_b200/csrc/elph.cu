// elph.cu -- phonon-field plumbing on the device: K5 (operator refresh straight from x) and K7 (Lambda).
//
// Replaces SmoQyDQMC.update!(fpi, elph, x, +-1) followed by update!(fdm, fpi)
// (src/EFAPFFHMCUpdater.jl:149-153,200-205 and src/FermionDetMatrix.jl:208-236) -- V and t are never
// materialised: exp(-dtau V), cosh, sinh are computed from x in one pass -- plus update_Λ!, mul_Λ!, ldiv_Λ!,
// mul_Λᵀ!, ldiv_Λᵀ! (src/holstein_shift_matrix.jl:2-153) and the bosonic action [unvendored, restated from
// arXiv:2404.09723; dispersive phonon couplings are not modelled].
#include "sq_internal.h"

#include <cmath>

__device__ __forceinline__ double poly4(double a, double a2, double a3, double a4, double x) {
    return x * (a + x * (a2 + x * (a3 + x * a4)));
}

struct ElphDev {
    int L, N, Nh, Nph, Nhol, Nssh;
    double dtau;
    const double *x;
    const int *hol_ph, *hol_site, *hol_sym, *site_ptr, *site_cpl;
    const double *ha;
    const int *ssh_p, *ssh_pp, *bond_ptr, *bond_cpl;
    const double *sa;
    const double *V0, *t0;
    const int *perm;
};

__device__ __forceinline__ double site_energy(const ElphDev &E, int l, int i) {
    double V = E.V0[i];
    for (int q = E.site_ptr[i]; q < E.site_ptr[i + 1]; q++) {
        int c = E.site_cpl[q];
        double x = E.x[(size_t)l * E.Nph + E.hol_ph[c]];
        V += poly4(E.ha[c], E.ha[E.Nhol + c], E.ha[2 * E.Nhol + c], E.ha[3 * E.Nhol + c], x);
    }
    return V;
}
// hopping amplitude of checkerboard bond n
__device__ __forceinline__ double bond_hopping(const ElphDev &E, int l, int n) {
    double t = E.t0[E.perm[n]];
    for (int q = E.bond_ptr[n]; q < E.bond_ptr[n + 1]; q++) {
        int c = E.bond_cpl[q];
        double dx = E.x[(size_t)l * E.Nph + E.ssh_pp[c]] - E.x[(size_t)l * E.Nph + E.ssh_p[c]];
        t -= poly4(E.sa[c], E.sa[E.Nssh + c], E.sa[2 * E.Nssh + c], E.sa[3 * E.Nssh + c], dx);
    }
    return t;
}

// K5: operator coefficients straight from the phonon field
__global__ void k_refresh_from_x(const __grid_constant__ ElphDev E, double *__restrict__ expV, double2 *__restrict__ cs, double dtaup) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t nV = (size_t)E.L * E.N, nT = (size_t)E.L * E.Nh;
    if (idx < nV) {
        int l = (int)(idx / E.N), i = (int)(idx - (size_t)l * E.N);
        expV[idx] = exp(-E.dtau * site_energy(E, l, i));
    }
    if (idx < nT) {
        int l = (int)(idx / E.Nh), n = (int)(idx - (size_t)l * E.Nh);
        double tp = bond_hopping(E, l, n);
        double a = dtaup * fabs(tp);
        double sg = (tp > 0.0) ? 1.0 : ((tp < 0.0) ? -1.0 : 0.0);
        cs[idx] = make_double2(cosh(a), sg * sinh(a));
    }
}
// materialise V [l][i] and t [l][h] (ORIGINAL hopping order) for read-back
__global__ void k_build_Vt(const __grid_constant__ ElphDev E, double *__restrict__ V, double *__restrict__ t) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t nV = (size_t)E.L * E.N, nT = (size_t)E.L * E.Nh;
    if (idx < nV) {
        int l = (int)(idx / E.N), i = (int)(idx - (size_t)l * E.N);
        V[idx] = site_energy(E, l, i);
    }
    if (idx < nT) {
        int l = (int)(idx / E.Nh), n = (int)(idx - (size_t)l * E.Nh);
        t[(size_t)l * E.Nh + E.perm[n]] = bond_hopping(E, l, n);
    }
}
// update_Λ!: holstein_shift_matrix.jl:2-44
__global__ void k_update_lambda(const __grid_constant__ ElphDev E, double *__restrict__ Lam) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)E.L * E.N) return;
    int l = (int)(idx / E.N), i = (int)(idx - (size_t)l * E.N);
    double lam = (l == 0) ? 1.0 : -1.0;
    for (int q = E.site_ptr[i]; q < E.site_ptr[i + 1]; q++) {
        int c = E.site_cpl[q];
        if (!E.hol_sym[c]) continue;
        double x = E.x[(size_t)l * E.Nph + E.hol_ph[c]];
        lam = exp(+E.dtau * (E.ha[c] * x + E.ha[2 * E.Nhol + c] * x * x * x) / 2) * lam;      // :37
    }
    Lam[idx] = lam;
}
// which: 0 mul_Λ (:47), 1 ldiv_Λ (:74), 2 mul_Λᵀ (:102), 3 ldiv_Λᵀ (:129).  out must not alias in.
__global__ void k_lambda_op(double2 *__restrict__ out, const double2 *__restrict__ in, const double *__restrict__ Lam, int L, int N, int which) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)L * N) return;
    int l = (int)(idx / N), i = (int)(idx - (size_t)l * N);
    int lp = (l + 1 == L) ? 0 : l + 1, lm = (l == 0) ? L - 1 : l - 1;
    double2 v;
    double a;
    switch (which) {
        case 0: a = Lam[(size_t)lp * N + i]; v = in[(size_t)lp * N + i]; out[idx] = make_double2(a * v.x, a * v.y); break;
        case 1: a = Lam[idx]; v = in[(size_t)lm * N + i]; out[idx] = make_double2(v.x / a, v.y / a); break;
        case 2: a = Lam[idx]; v = in[(size_t)lm * N + i]; out[idx] = make_double2(a * v.x, a * v.y); break;
        default: a = Lam[(size_t)lp * N + i]; v = in[(size_t)lp * N + i]; out[idx] = make_double2(v.x / a, v.y / a); break;
    }
}
// bosonic_action(elph, holstein_correction = false) [unvendored]: harmonic + quartic potential + kinetic term
__global__ void k_bosonic_action(const double *__restrict__ x, const double *__restrict__ Om, const double *__restrict__ Om4,
                                 const double *__restrict__ M, const int *__restrict__ fin, int L, int Nph, double dtau,
                                 double *__restrict__ part) {
    __shared__ double red[32];
    double acc = 0;
    size_t tot = (size_t)L * Nph;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < tot; k += (size_t)gridDim.x * blockDim.x) {
        int l = (int)(k / Nph), p = (int)(k - (size_t)l * Nph);
        if (!fin[p]) continue;
        int ln = (l + 1 == L) ? 0 : l + 1;
        double xv = x[k], d = x[(size_t)ln * Nph + p] - xv, m = M[p], o = Om[p], o4 = Om4[p];
        acc += dtau * m * o * o * xv * xv / 2 + dtau * m * o4 * o4 * xv * xv * xv * xv / 24 + m * d * d / (2 * dtau);
    }
    double v[1] = {acc};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) part[blockIdx.x] = v[0];
}
// dispersive part of the bosonic action [unvendored, SmoQyDQMC DispersionParameters]:
//   sum_{l, d} dtau M''_d [ Omega_d^2 D^2 / 2 + Omega4_d^2 D^4 / 24 ],  D = x[p'_d, l] - x[p_d, l];  k2 = M'' Omega^2, k4 = M'' Omega4^2
__global__ void k_dispersive_action(const double *__restrict__ x, const int *__restrict__ dp, const int *__restrict__ dpp,
                                    const double *__restrict__ k2, const double *__restrict__ k4, int L, int Nph, int Ndisp, double dtau,
                                    double *__restrict__ part) {
    __shared__ double red[32];
    double acc = 0;
    size_t tot = (size_t)L * Ndisp;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < tot; k += (size_t)gridDim.x * blockDim.x) {
        int l = (int)(k / Ndisp), c = (int)(k - (size_t)l * Ndisp);
        double D = x[(size_t)l * Nph + dpp[c]] - x[(size_t)l * Nph + dp[c]], D2 = D * D;
        acc += dtau * k2[c] * D2 / 2 + dtau * k4[c] * D2 * D2 / 24;
    }
    double v[1] = {acc};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) part[blockIdx.x] = v[0];
}
// p <- p - dt (dS/dx + anharmonic + dispersive derivative);  EFAPFFHMCUpdater.jl:190-196.  The dispersive term is gathered per phonon
// from its signed coupling list (fixed order, no floating-point atomics): eval_derivative_dispersive_action! [unvendored].
__global__ void k_kick(double *__restrict__ pm, const double *__restrict__ dS, const double *__restrict__ x, const double *__restrict__ Om4,
                       const double *__restrict__ M, const int *__restrict__ fin, int Nph, double dtau, double dt, size_t n, int Ndisp,
                       const int *__restrict__ dp, const int *__restrict__ dpp, const double *__restrict__ k2, const double *__restrict__ k4,
                       const int *__restrict__ dptr, const int *__restrict__ dcpl) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        int p = (int)(k % Nph);
        size_t base = k - p;
        double f = dS[k];
        if (fin[p]) {
            double xv = x[k];
            f += dtau * M[p] * Om4[p] * Om4[p] * xv * xv * xv / 6;
            if (Ndisp) {
                for (int q = dptr[p]; q < dptr[p + 1]; q++) {
                    int sc = dcpl[q], c = (sc > 0 ? sc : -sc) - 1;
                    double D = x[base + dpp[c]] - x[base + dp[c]];
                    double g = dtau * (k2[c] * D + k4[c] * D * D * D / 6);
                    f += sc > 0 ? g : -g;
                }
            }
        }
        pm[k] -= dt * f;
    }
}

// ---------------------------------------------------------------------------------------------------
static ElphDev elph_dev(const sq_elph *e) {
    const sq_fdm *f = e->f;
    ElphDev E;
    E.L = (int)f->L; E.N = (int)f->N; E.Nh = (int)f->Nh; E.Nph = (int)e->Nph; E.Nhol = (int)e->Nhol; E.Nssh = (int)e->Nssh;
    E.dtau = e->dtau; E.x = e->x.p;
    E.hol_ph = e->hol_ph.p; E.hol_site = e->hol_site.p; E.hol_sym = e->hol_sym.p; E.site_ptr = e->site_ptr.p; E.site_cpl = e->site_cpl.p;
    E.ha = e->ha.p; E.ssh_p = e->ssh_p.p; E.ssh_pp = e->ssh_pp.p; E.bond_ptr = e->bond_ptr.p; E.bond_cpl = e->bond_cpl.p; E.sa = e->sa.p;
    E.V0 = e->V0.p; E.t0 = e->t0.p; E.perm = f->perm.p;
    return E;
}

double reduce_partials_host(sq_fdm *f, const double *d_part, int n) {
    std::vector<double> h(n);
    SQ_CUDA(cudaMemcpyAsync(h.data(), d_part, n * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    double s = 0;
    for (int k = 0; k < n; k++) s += h[k];
    return s;
}

static void csr_build(int nrows, const std::vector<int> &row_of, std::vector<int> &ptr, std::vector<int> &items) {
    ptr.assign(nrows + 1, 0);
    for (int r : row_of) ptr[r + 1]++;
    for (int r = 0; r < nrows; r++) ptr[r + 1] += ptr[r];
    items.assign(row_of.size(), 0);
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (size_t c = 0; c < row_of.size(); c++) items[fill[row_of[c]]++] = (int)c;
}

template <class T>
static void up(DevBuf<T> &b, const std::vector<T> &v, cudaStream_t s) {
    b.alloc(v.size() + 1, false);
    b.upload(v.data(), v.size(), s);
    SQ_CUDA(cudaStreamSynchronize(s));
}

// bare on-site energies (N) and hoppings (Nh, original order): eps - mu and t of TightBindingParameters
void elph_set_bare(sq_elph *e, const double *V0, const double *t0) {
    sq_fdm *f = e->f;
    SQ_REQUIRE(V0 && (t0 || f->Nh == 0), "NULL model array");
    e->V0.upload(V0, f->N, f->stream);
    if (f->Nh) e->t0.upload(t0, f->Nh, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    // bare hoppings equal inside every colour (|t| and sign): the register path of the fused matvec applies
    e->t0_coluni = (t0 != nullptr && f->Nh > 0);
    for (i64 c = 0; c < f->C && e->t0_coluni; c++)
        for (int h = f->clo[c]; h < f->chi[c]; h++)
            if (t0[f->h_perm[h]] != t0[f->h_perm[f->clo[c]]]) { e->t0_coluni = false; break; }
    e->bare_set = true;
}

// DispersionParameters: disp_ph (2 x Ndisp) 1-based phonon pairs, Omega / Omega4 per coupling.  The reduced mass M'' = M M' / (M + M')
// (one infinite mass: the other one; both infinite: the coupling carries no dynamics and is dropped) is folded into k2, k4.
void elph_set_dispersion(sq_elph *e, i64 Ndisp, const i64 *disp_ph, const double *Om, const double *Om4) {
    sq_fdm *f = e->f;
    SQ_REQUIRE(Ndisp >= 0 && (Ndisp == 0 || (disp_ph && Om && Om4)), "bad dispersion tables");
    const i64 Nph = e->Nph;
    std::vector<int> dp, dpp;
    std::vector<double> k2, k4;
    for (i64 c = 0; c < Ndisp; c++) {
        const i64 p = disp_ph[2 * c] - 1, pp = disp_ph[2 * c + 1] - 1;
        SQ_REQUIRE(p >= 0 && p < Nph && pp >= 0 && pp < Nph && p != pp, "dispersive coupling map out of range");
        const double a = e->h_M[p], b = e->h_M[pp];
        const double m = !std::isfinite(a) ? b : (!std::isfinite(b) ? a : a * b / (a + b));
        if (!std::isfinite(m)) continue;
        dp.push_back((int)p); dpp.push_back((int)pp);
        k2.push_back(m * Om[c] * Om[c]); k4.push_back(m * Om4[c] * Om4[c]);
    }
    e->Ndisp = (i64)dp.size();
    std::vector<int> cnt(Nph + 1, 0);
    for (size_t c = 0; c < dp.size(); c++) { cnt[dp[c] + 1]++; cnt[dpp[c] + 1]++; }
    for (i64 p = 0; p < Nph; p++) cnt[p + 1] += cnt[p];
    std::vector<int> items(2 * dp.size() + 1), fill(cnt.begin(), cnt.end() - 1);
    for (size_t c = 0; c < dp.size(); c++) { items[fill[dp[c]]++] = -(int)(c + 1); items[fill[dpp[c]]++] = (int)(c + 1); }
    cudaStream_t s = f->stream;
    up(e->disp_p, dp, s); up(e->disp_pp, dpp, s); up(e->disp_k2, k2, s); up(e->disp_k4, k4, s);
    up(e->ph_disp_ptr, cnt, s); up(e->ph_disp_cpl, items, s);
}

void elph_add_potential_derivative(sq_elph *e, double *pm, const double *dS, double dt) {
    sq_fdm *f = e->f;
    const size_t nx = (size_t)f->L * e->Nph;
    const int g = std::min(SQ_MAXPART, f->num_sms * 4);
    k_kick<<<g, 256, 0, f->stream>>>(pm, dS, e->x.p, e->Om4.p, e->M.p, e->fin.p, (int)e->Nph, e->dtau, dt, nx, (int)e->Ndisp, e->disp_p.p,
                                     e->disp_pp.p, e->disp_k2.p, e->disp_k4.p, e->ph_disp_ptr.p, e->ph_disp_cpl.p);
    SQ_LAUNCH_CHECK();
    f->launches++;
}

void elph_create_impl(sq_elph **out, sq_fdm *f, double dtau, i64 Nph, const double *Om, const double *Om4, const double *M, i64 Nhol,
                      const i64 *hol_ph, const i64 *hol_site, const double *a, const double *a2, const double *a3, const double *a4,
                      const int32_t *hol_sym, i64 Nssh, const i64 *ssh_ph, const i64 *ssh_hop, const double *sa, const double *sa2,
                      const double *sa3, const double *sa4, const double *V0, const double *t0) {
    SQ_REQUIRE(out && f, "NULL argument");
    SQ_REQUIRE(Nph >= 1 && Nhol >= 0 && Nssh >= 0, "bad dimensions");
    SQ_REQUIRE(Om && Om4 && M && (!V0 || t0 || f->Nh == 0), "NULL model array");
    SQ_REQUIRE((size_t)f->L * (size_t)Nph < (size_t)1 << 31, "phonon field too large for 32-bit indexing");
    SQ_CUDA(cudaSetDevice(f->device));
    sq_elph *e = new sq_elph();
    try {
        cudaStream_t s = f->stream;
        e->f = f; e->dtau = dtau; e->Nph = Nph; e->Nhol = Nhol; e->Nssh = Nssh;
        e->h_M.assign(M, M + Nph);
        std::vector<int> fin(Nph);
        for (i64 p = 0; p < Nph; p++) fin[p] = std::isfinite(M[p]) ? 1 : 0;
        e->x.alloc((size_t)f->L * Nph);
        up(e->Om, std::vector<double>(Om, Om + Nph), s);
        up(e->Om4, std::vector<double>(Om4, Om4 + Nph), s);
        up(e->M, std::vector<double>(M, M + Nph), s);
        up(e->fin, fin, s);
        // Holstein
        std::vector<int> hp(Nhol), hs(Nhol), hy(Nhol);
        std::vector<double> ha(4 * Nhol);
        for (i64 c = 0; c < Nhol; c++) {
            hp[c] = (int)(hol_ph[c] - 1); hs[c] = (int)(hol_site[c] - 1); hy[c] = hol_sym[c] ? 1 : 0;
            SQ_REQUIRE(hp[c] >= 0 && hp[c] < Nph && hs[c] >= 0 && hs[c] < f->N, "Holstein coupling map out of range");
            ha[c] = a[c]; ha[Nhol + c] = a2[c]; ha[2 * Nhol + c] = a3[c]; ha[3 * Nhol + c] = a4[c];
            if (hy[c]) e->any_phsym = true;
        }
        up(e->hol_ph, hp, s); up(e->hol_site, hs, s); up(e->hol_sym, hy, s); up(e->ha, ha, s);
        std::vector<int> ptr, items;
        csr_build((int)f->N, hs, ptr, items);
        up(e->site_ptr, ptr, s); up(e->site_cpl, items, s);
        csr_build((int)Nph, hp, ptr, items);
        up(e->ph_hol_ptr, ptr, s); up(e->ph_hol_cpl, items, s);
        // SSH
        std::vector<int> chk_of_hop(f->Nh);
        for (i64 n = 0; n < f->Nh; n++) chk_of_hop[f->h_perm[n]] = (int)n;
        std::vector<int> sp(Nssh), spp(Nssh), sb(Nssh);
        std::vector<double> sav(4 * Nssh);
        for (i64 c = 0; c < Nssh; c++) {
            sp[c] = (int)(ssh_ph[2 * c] - 1); spp[c] = (int)(ssh_ph[2 * c + 1] - 1);
            i64 h = ssh_hop[c] - 1;
            SQ_REQUIRE(sp[c] >= 0 && sp[c] < Nph && spp[c] >= 0 && spp[c] < Nph && h >= 0 && h < f->Nh, "SSH coupling map out of range");
            sb[c] = chk_of_hop[h];
            sav[c] = sa[c]; sav[Nssh + c] = sa2[c]; sav[2 * Nssh + c] = sa3[c]; sav[3 * Nssh + c] = sa4[c];
        }
        up(e->ssh_p, sp, s); up(e->ssh_pp, spp, s); up(e->ssh_bond, sb, s); up(e->sa, sav, s);
        csr_build((int)f->Nh, sb, ptr, items);
        up(e->bond_ptr, ptr, s); up(e->bond_cpl, items, s);
        // phonon -> signed SSH couplings: -(c+1) for the first phonon of the pair, +(c+1) for the second
        // (fermion_det_matrix_dervative.jl:242-247: F[p] -= val, F[p'] += val)
        std::vector<int> cnt(Nph + 1, 0);
        for (i64 c = 0; c < Nssh; c++) { cnt[sp[c] + 1]++; cnt[spp[c] + 1]++; }
        for (i64 p = 0; p < Nph; p++) cnt[p + 1] += cnt[p];
        std::vector<int> sitems(2 * Nssh), fillp(cnt.begin(), cnt.end() - 1);
        for (i64 c = 0; c < Nssh; c++) { sitems[fillp[sp[c]]++] = -(int)(c + 1); sitems[fillp[spp[c]]++] = (int)(c + 1); }
        up(e->ph_ssh_ptr, cnt, s); up(e->ph_ssh_cpl, sitems, s);
        // bare tight-binding terms: given now, or later through sq_elph_set_bare (the Julia PFFCalculator(elph, fdm) constructor has
        // no access to them; they arrive with the first call that carries the FermionPathIntegral)
        e->V0.alloc(f->N + 1); e->t0.alloc(f->Nh + 1);
        if (V0) elph_set_bare(e, V0, t0);
    } catch (...) {
        delete e;
        throw;
    }
    *out = e;
}

void elph_refresh_fdm(sq_elph *e) {
    sq_fdm *f = e->f;
    SQ_REQUIRE(e->bare_set, "bare on-site energies / hoppings not set: create the model with V0, t0 or call sq_elph_set_bare first");
    ElphDev E = elph_dev(e);
    size_t tot = (size_t)f->L * std::max(f->N, f->Nh);
    k_refresh_from_x<<<(unsigned)((tot + 255) / 256), 256, 0, f->stream>>>(E, f->expV.p, f->cs.p, f->sym ? e->dtau / 2 : e->dtau);
    SQ_LAUNCH_CHECK();
    f->launches++;
    f->coef_version++;
    f->cs_uniform = (e->Nssh == 0) ? 1 : 0;       // no SSH coupling => hoppings (and cosh, sinh) are tau-independent
    f->cs_coluni = (f->cs_uniform && e->t0_coluni) ? 1 : 0;
}

void elph_build_Vt(sq_elph *e) {
    sq_fdm *f = e->f;
    SQ_REQUIRE(e->bare_set, "bare on-site energies / hoppings not set: create the model with V0, t0 or call sq_elph_set_bare first");
    if (!e->V.p) { e->V.alloc((size_t)f->L * f->N); e->t.alloc((size_t)f->L * f->Nh + 1); }
    ElphDev E = elph_dev(e);
    size_t tot = (size_t)f->L * std::max(f->N, f->Nh);
    k_build_Vt<<<(unsigned)((tot + 255) / 256), 256, 0, f->stream>>>(E, e->V.p, e->t.p);
    SQ_LAUNCH_CHECK();
    f->launches++;
}

void elph_update_lambda(sq_elph *e, double *Lam) {
    sq_fdm *f = e->f;
    ElphDev E = elph_dev(e);
    size_t tot = (size_t)f->L * f->N;
    k_update_lambda<<<(unsigned)((tot + 255) / 256), 256, 0, f->stream>>>(E, Lam);
    SQ_LAUNCH_CHECK();
    f->launches++;
}

void elph_lambda_op(sq_elph *e, int which, double2 *out, const double2 *in, const double *Lam) {
    sq_fdm *f = e->f;
    SQ_REQUIRE(out != in, "Lambda operations are out of place");
    size_t tot = (size_t)f->L * f->N;
    k_lambda_op<<<(unsigned)((tot + 255) / 256), 256, 0, f->stream>>>(out, in, Lam, (int)f->L, (int)f->N, which);
    SQ_LAUNCH_CHECK();
    f->launches++;
}

double elph_bosonic_action(sq_elph *e) {
    sq_fdm *f = e->f;
    int nb = std::min(SQ_MAXPART, f->num_sms * 2);
    k_bosonic_action<<<nb, 256, 0, f->stream>>>(e->x.p, e->Om.p, e->Om4.p, e->M.p, e->fin.p, (int)f->L, (int)e->Nph, e->dtau, f->part.p);
    SQ_LAUNCH_CHECK();
    f->launches++;
    if (e->Ndisp > 0) {
        k_dispersive_action<<<nb, 256, 0, f->stream>>>(e->x.p, e->disp_p.p, e->disp_pp.p, e->disp_k2.p, e->disp_k4.p, (int)f->L, (int)e->Nph,
                                                      (int)e->Ndisp, e->dtau, f->part.p + nb);
        SQ_LAUNCH_CHECK();
        f->launches++;
        return reduce_partials_host(f, f->part.p, nb) + reduce_partials_host(f, f->part.p + nb, nb);
    }
    return reduce_partials_host(f, f->part.p, nb);
}

// abi.cu -- extern "C" entry points of include/smoqyelph_b200.h: argument checks, exception -> status
// code translation, host <-> device staging.  No compute lives here.
#include "sq_internal.h"

#include <algorithm>

#include <cstring>
#include <mutex>

static thread_local std::string g_last_error;
void sq_set_last_error(const std::string &m) { g_last_error = m; }

#define SQ_TRY try {
#define SQ_CATCH                                                              \
    }                                                                         \
    catch (const SqNumericalInstability &e) { sq_set_last_error(e.what()); return 3; } \
    catch (const std::exception &e) { sq_set_last_error(e.what()); return 1; } \
    catch (...) { sq_set_last_error("unknown C++ exception"); return 2; }     \
    return 0;

void fdm_create_impl(sq_fdm **out, int sym, i64 L, i64 N, i64 Nh, const i64 *nt, const i64 *perm, i64 C,
                     const i64 *clo, const i64 *chi, double tol, i64 maxiter, int device);
void fdm_destroy_impl(sq_fdm *f);
void fdm_update_impl(sq_fdm *f, const double *V, const double *t, double dtau);
void fdm_get_coefficients_impl(sq_fdm *f, double *expV, double *ch, double *sh);
void fdm_mul_impl(sq_fdm *f, int op, void *out, const void *in);
void fdm_select_tuning(sq_fdm *f);
bool fdm_v3_supported(const sq_fdm *f, int S);
void fdm_v3_prepare_native(sq_fdm *f);
void slab_unique_id(char *out128);
void slab_init(sq_fdm *f, int rank, int world, const char *id128);
void slab_set_range(sq_fdm *f, int lo, int hi);
void slab_mailbox_create(sq_fdm *f, char *out64);
void slab_mailbox_open(sq_fdm *f, const char *handles64);

void kpm_create_impl(sq_kpm **out, sq_fdm *f, double rbuf, i64 n, double a1, double a2);
void elph_create_impl(sq_elph **out, sq_fdm *f, double dtau, i64 Nph, const double *Om, const double *Om4, const double *M, i64 Nhol,
                      const i64 *hol_ph, const i64 *hol_site, const double *a, const double *a2, const double *a3, const double *a4,
                      const int32_t *hol_sym, i64 Nssh, const i64 *ssh_ph, const i64 *ssh_hop, const double *sa, const double *sa2,
                      const double *sa3, const double *sa4, const double *V0, const double *t0);
void elph_build_Vt(sq_elph *e);
void pff_create_impl(sq_pff **out, sq_elph *e);
double pff_sample_dev(sq_pff *q);
double pff_action_dev(sq_pff *q, sq_kpm *kpm, bool refresh, const double *h_lanczos, const double *d_lanczos, double tol, i64 maxiter,
                      i64 *iters, double *eps);
double pff_force_dev(sq_pff *q, sq_kpm *kpm, bool refresh, const double *h_lanczos, const double *d_lanczos, double tol, i64 maxiter,
                     i64 *iters, double *eps);
void pff_dM_dx_host(sq_pff *q, double *F, double nu, const void *u, const void *v);
void pff_dLambda_dx_host(sq_pff *q, double *F, double nu, const void *upv, const void *uv);
void pff_fill_phi_normals(sq_pff *q, const void *h_R, const double *d_stream);
void hmc_create_impl(sq_hmc **out, sq_pff *q, i64 Nt, double dt, double eta, double delta, uint64_t seed);
int hmc_update_impl(sq_hmc *h, sq_kpm *kpm, double tol_action, double tol_force, i64 maxiter, const double *randoms, i64 nrandoms,
                    double *info);
void hmc_evolve_dev(sq_hmc *h, double *x, double *pm, double dt);
double hmc_init_momentum_dev(sq_hmc *h, const double *R, double *pm);
double hmc_kinetic_dev(sq_hmc *h, const double *pm);
void greens_create_impl(sq_greens **out, sq_fdm *f, i64 Nrv, uint64_t seed);
double greens_update_impl(sq_greens *g, sq_kpm *kpm, const void *h_R, double tol, i64 maxiter);
void greens_measure_impl(sq_greens *g, double *out);
void greens_measure_double_occ_orbital_impl(sq_greens *g, int norb, int a, double *out);
void greens_measure_c4_impl(sq_greens *g, int kind, int norb, int ndim, const i64 *dims, const int *orb, const i64 *r, void *h_out,
                            const double *h_tD, const double *h_t0);
void greens_measure_n_orbital_impl(sq_greens *g, int norb, int a, double *out);
void greens_weighted_density_impl(sq_greens *g, const double *h_w, double *out);
void greens_weighted_bonds_impl(sq_greens *g, i64 nbonds, const i64 *bonds, const void *h_w, double *out);
void greens_measure_GD0_impl(sq_greens *g, int norb, int ndim, const i64 *dims, int a, int b, void *h_out);

extern "C" {

const char *sq_last_error(void) { return g_last_error.c_str(); }
int sq_version(void) { return 100; }
int sq_device_count(int *count) {
    SQ_TRY
    SQ_REQUIRE(count != nullptr, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    SQ_CATCH
}

int sq_fdm_create(sq_fdm **out, int sym, int64_t Ltau, int64_t N, int64_t Nh, const int64_t *nt, const int64_t *perm,
                  int64_t ncolors, const int64_t *color_lo, const int64_t *color_hi, double tol, int64_t maxiter, int device) {
    SQ_TRY
    fdm_create_impl(out, sym, Ltau, N, Nh, nt, perm, ncolors, color_lo, color_hi, tol, maxiter, device);
    SQ_CATCH
}
int sq_fdm_destroy(sq_fdm *f) {
    SQ_TRY
    fdm_destroy_impl(f);
    SQ_CATCH
}
int sq_fdm_update(sq_fdm *f, const double *V, const double *t, double dtau) {
    SQ_TRY
    SQ_REQUIRE(f && V && (t || f->Nh == 0), "NULL argument");
    fdm_update_impl(f, V, t, dtau);
    SQ_CATCH
}
int sq_fdm_mul(sq_fdm *f, int op, sq_complex *out, const sq_complex *in) {
    SQ_TRY
    SQ_REQUIRE(f && out && in, "NULL argument");
    fdm_mul_impl(f, op, out, in);
    SQ_CATCH
}
int sq_fdm_get_coefficients(sq_fdm *f, double *expV, double *cosh_t, double *sinh_t) {
    SQ_TRY
    SQ_REQUIRE(f && expV && cosh_t && sinh_t, "NULL argument");
    fdm_get_coefficients_impl(f, expV, cosh_t, sinh_t);
    SQ_CATCH
}
int sq_nccl_unique_id(char *out128) {
    SQ_TRY
    SQ_REQUIRE(out128, "NULL argument");
    slab_unique_id(out128);
    SQ_CATCH
}
int sq_fdm_init_slab(sq_fdm *f, int rank, int world, const char *id128) {
    SQ_TRY
    SQ_REQUIRE(f && (id128 || world == 1), "NULL argument");
    slab_init(f, rank, world, id128);
    SQ_CATCH
}
int sq_fdm_mailbox_create(sq_fdm *f, char *out64) {
    SQ_TRY
    SQ_REQUIRE(f && out64, "NULL argument");
    slab_mailbox_create(f, out64);
    SQ_CATCH
}
int sq_fdm_mailbox_open(sq_fdm *f, const char *handles64) {
    SQ_TRY
    SQ_REQUIRE(f && handles64, "NULL argument");
    slab_mailbox_open(f, handles64);
    SQ_CATCH
}
int sq_fdm_set_sharded_solve(sq_fdm *f, int enable) {
    SQ_TRY
    SQ_REQUIRE(f, "NULL handle");
    slab_set_sharded(f, enable);
    SQ_CATCH
}
int sq_fdm_set_slab_range(sq_fdm *f, int64_t lo, int64_t hi) {
    SQ_TRY
    SQ_REQUIRE(f, "NULL handle");
    slab_set_range(f, (int)lo, (int)hi);
    SQ_CATCH
}
int sq_fdm_get_slab(sq_fdm *f, int64_t *lo, int64_t *hi, int *rank, int *world) {
    SQ_TRY
    SQ_REQUIRE(f, "NULL handle");
    if (lo) *lo = f->slab_lo;
    if (hi) *hi = f->slab_hi;
    if (rank) *rank = f->rank;
    if (world) *world = f->world;
    SQ_CATCH
}
int sq_fdm_mul_dev(sq_fdm *f, int op, void *d_out, const void *d_in) {
    SQ_TRY
    SQ_REQUIRE(f && d_out && d_in, "NULL argument");
    SQ_CUDA(cudaSetDevice(f->device));
    if (f->world > 1 && !f->sharded) {
        SQ_REQUIRE(op != SQ_OP_MMT, "M M^T is not available in tau-slab mode");
        fdm_halo_exchange(f, (double2 *)d_in);       // the neighbours' boundary slices are written into d_in at their global index
    }
    fdm_mul_dev(f, op, (double2 *)d_out, (const double2 *)d_in);
    SQ_CATCH
}
int sq_fdm_time_mul(sq_fdm *f, int op, void *d_out, const void *d_in, int reps, void *d_flush, int64_t flush_bytes,
                    double *us_per_launch) {
    SQ_TRY
    SQ_REQUIRE(f && d_out && d_in && us_per_launch && reps >= 1, "bad argument");
    SQ_REQUIRE(f->world == 1 || f->sharded, "single-GPU measurement only");
    SQ_CUDA(cudaSetDevice(f->device));
    cudaEvent_t e0, e1;
    SQ_CUDA(cudaEventCreate(&e0));
    SQ_CUDA(cudaEventCreate(&e1));
    // op 102: M^T M of the register path on vectors in its native order (what the CG solver runs), timing only
    const bool native = (op == 102);
    if (native) {
        fdm_select_tuning(f);
        SQ_REQUIRE(f->path == 0 && f->use_v3 && fdm_v3_supported(f, f->v3_S), "register path not active");
        fdm_v3_prepare_native(f);
    }
    // op 200 + n / 300 + n: the fused M^T M kernel of the register path on a BATCH of n vectors (V elements apart) in library / native
    // order -- what the multi-RHS solver launches once per iteration (cg_batch.cu); d_in / d_out hold n vectors
    const int nbatch = (op >= 300) ? op - 300 : ((op >= 200) ? op - 200 : 0);
    const bool batch_native = op >= 300;
    if (nbatch) {
        SQ_REQUIRE(nbatch >= 1 && nbatch <= 64, "batch size out of range");
        fdm_select_tuning(f);
        SQ_REQUIRE(f->path == 0 && fdm_v3_supported(f, f->v3_S), "register path not active");
        if (batch_native) fdm_v3_prepare_native(f);
    }
    auto fdm_mul_dev = [&](sq_fdm *ff, int o, double2 *out, const double2 *in) {
        if (nbatch) fdm_v3_launch(ff, 2, ff->v3_S, out, in, nullptr, nullptr, batch_native, nbatch, (size_t)ff->L * ff->N, 0);
        else if (native) fdm_v3_launch(ff, 2, ff->v3_S, out, in, nullptr, nullptr, true);
        else ::fdm_mul_dev(ff, o, out, in);
    };
    for (int k = 0; k < 3; k++) fdm_mul_dev(f, op, (double2 *)d_out, (const double2 *)d_in);
    if (!d_flush) {
        SQ_CUDA(cudaEventRecord(e0, f->stream));
        for (int k = 0; k < reps; k++) fdm_mul_dev(f, op, (double2 *)d_out, (const double2 *)d_in);
        SQ_CUDA(cudaEventRecord(e1, f->stream));
        SQ_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        SQ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        *us_per_launch = 1e3 * ms / reps;
    } else {
        // L2-cold: (flush + launch) x reps minus flush x reps, both timed as a whole -- single launches are too short for the
        // ~2 us granularity of back-to-back event pairs on this system
        float ms_both = 0, ms_flush = 0;
        SQ_CUDA(cudaMemsetAsync(d_flush, 0, (size_t)flush_bytes, f->stream));
        SQ_CUDA(cudaEventRecord(e0, f->stream));
        for (int k = 0; k < reps; k++) {
            SQ_CUDA(cudaMemsetAsync(d_flush, k & 1, (size_t)flush_bytes, f->stream));
            fdm_mul_dev(f, op, (double2 *)d_out, (const double2 *)d_in);
        }
        SQ_CUDA(cudaEventRecord(e1, f->stream));
        SQ_CUDA(cudaEventSynchronize(e1));
        SQ_CUDA(cudaEventElapsedTime(&ms_both, e0, e1));
        SQ_CUDA(cudaEventRecord(e0, f->stream));
        for (int k = 0; k < reps; k++) SQ_CUDA(cudaMemsetAsync(d_flush, k & 1, (size_t)flush_bytes, f->stream));
        SQ_CUDA(cudaEventRecord(e1, f->stream));
        SQ_CUDA(cudaEventSynchronize(e1));
        SQ_CUDA(cudaEventElapsedTime(&ms_flush, e0, e1));
        *us_per_launch = 1e3 * (ms_both - ms_flush) / reps;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    SQ_CATCH
}
int sq_fdm_set_tuning(sq_fdm *f, int slab, int threads) {
    SQ_TRY
    SQ_REQUIRE(f, "NULL handle");
    SQ_REQUIRE(f->path == 0, "fused path not available for this lattice size");
    SQ_REQUIRE(slab >= 1 && slab <= f->L, "slab out of range");
    SQ_REQUIRE(threads >= 32 && threads <= 1024 && threads % 32 == 0, "threads must be a multiple of 32 in [32, 1024]");
    SQ_REQUIRE((size_t)(2 * slab + 3) * f->N * sizeof(double2) <= f->smem_optin, "slab does not fit in shared memory");
    f->slab = slab;
    f->threads = threads;
    f->manual_tuning = 1;
    SQ_CATCH
}
int sq_fdm_set_fast_path(sq_fdm *f, int enable) {
    SQ_TRY
    SQ_REQUIRE(f, "NULL handle");
    // 0: generic shared-memory kernel, 1: fast shared-memory kernel (fdm_v2.cu), 2: register path (fdm_v3.cu; falls back
    // to 1 when the lattice / coefficients do not qualify), optionally with slices-per-CTA in bits 8..: enable = 2 + 256 * S
    f->use_v2 = enable ? 1 : 0;
    f->use_v3 = ((enable & 255) == 2) ? 1 : 0;
    if (f->use_v3 && (enable >> 8) >= 1 && (enable >> 8) <= 7) f->v3_S = enable >> 8;
    f->v3_cg = f->use_v3;
    if (f->slab < 1) { f->slab = 1; f->threads = 256; }     // never tuned: a valid default for the shared-memory kernels
    f->manual_tuning = 1;
    SQ_CATCH
}
int sq_fdm_get_tuning(sq_fdm *f, int *slab, int *threads, int *path) {
    SQ_TRY
    SQ_REQUIRE(f, "NULL handle");
    fdm_select_tuning(f);
    if (slab) *slab = f->slab;
    if (threads) *threads = f->threads;
    if (path) *path = f->path == 1 ? 1 : (f->use_v2 ? 2 : 0);
    if (f->path == 0 && f->use_v3 && fdm_v3_supported(f, f->v3_S)) {      // register path: one warp per slice, S + 1 warps
        if (slab) *slab = f->v3_S;
        if (threads) *threads = 32 * (f->v3_S + 1);
        if (path) *path = 3;
    }
    SQ_CATCH
}
int sq_fdm_stream(sq_fdm *f, void **cuda_stream) {
    SQ_TRY
    SQ_REQUIRE(f && cuda_stream, "NULL argument");
    *cuda_stream = (void *)f->stream;
    SQ_CATCH
}
int64_t sq_fdm_launch_count(sq_fdm *f) { return f ? f->launches : -1; }
int sq_fdm_stats(sq_fdm *f, int64_t *out, int n) {
    SQ_TRY
    SQ_REQUIRE(f && out && n >= 0, "bad argument");
    for (int q = 0; q < n; q++) out[q] = q < SQ_NSTATS ? f->stats[q] : 0;
    SQ_CATCH
}

int sq_fdm_cg(sq_fdm *f, sq_complex *x, const sq_complex *b, int zero_start, sq_kpm *kpm, int refresh_kpm,
              const double *lanczos_start, double tol, int64_t maxiter, int64_t *iters, double *eps) {
    SQ_TRY
    SQ_REQUIRE(f && x && b && iters && eps, "NULL argument");
    SQ_CUDA(cudaSetDevice(f->device));
    if (kpm && refresh_kpm) kpm_update(kpm, lanczos_start, nullptr);
    // device staging: b -> tmp1 (cg never touches tmp1? it does through mul); use dedicated io2 for b, tmp2... x lives in io buffers
    double2 *dx = f->tmp2.p, *db = f->io2.p;
    fdm_host_to_dev(f, db, b);
    if (!zero_start) {
        if ((const void *)x == (const void *)b) SQ_CUDA(cudaMemcpyAsync(dx, db, f->vec_bytes(), cudaMemcpyDeviceToDevice, f->stream));
        else fdm_host_to_dev(f, dx, x);
    }
    i64 it = 0;
    double e = 0;
    fdm_cg_dev(f, dx, db, zero_start != 0, kpm, tol, maxiter, &it, &e);
    fdm_dev_to_host(f, x, dx);
    *iters = it;
    *eps = e;
    SQ_CATCH
}
int sq_fdm_cg_batch(sq_fdm *f, sq_complex *X, const sq_complex *B, int64_t nrhs, int zero_start, sq_kpm *kpm, int refresh_kpm,
                    const double *lanczos_start, double tol, int64_t maxiter, int64_t *iters, double *eps) {
    SQ_TRY
    SQ_REQUIRE(f && X && B && iters && eps && nrhs >= 1 && nrhs <= 64, "bad argument");
    SQ_CUDA(cudaSetDevice(f->device));
    if (kpm && refresh_kpm) kpm_update(kpm, lanczos_start, nullptr);
    const size_t V = (size_t)f->L * f->N;
    DevBuf<double2> dX, dB;
    dX.alloc(V * nrhs, false);
    dB.alloc(V * nrhs, false);
    for (int64_t j = 0; j < nrhs; j++) {
        fdm_host_to_dev(f, dB.p + j * V, (const char *)B + j * V * sizeof(double2));
        if (!zero_start) fdm_host_to_dev(f, dX.p + j * V, (const char *)X + j * V * sizeof(double2));
    }
    if (fdm_cg_batch_applicable(f, kpm, (int)nrhs)) {
        fdm_cg_batch_dev(f, dX.p, dB.p, (int)nrhs, zero_start != 0, kpm, tol, maxiter, iters, eps);
    } else {                                              // one by one (no active preconditioner, a single system, tau-slab mode)
        for (int64_t j = 0; j < nrhs; j++) fdm_cg_dev(f, dX.p + j * V, dB.p + j * V, zero_start != 0, kpm, tol, maxiter, iters + j, eps + j);
    }
    for (int64_t j = 0; j < nrhs; j++) fdm_dev_to_host(f, (char *)X + j * V * sizeof(double2), dX.p + j * V);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    SQ_CATCH
}
int sq_fdm_cg_dev(sq_fdm *f, void *d_x, const void *d_b, int zero_start, sq_kpm *kpm, double tol, int64_t maxiter,
                  int64_t *iters, double *eps) {
    SQ_TRY
    SQ_REQUIRE(f && d_x && d_b && iters && eps, "NULL argument");
    SQ_CUDA(cudaSetDevice(f->device));
    i64 it = 0;
    double e = 0;
    fdm_cg_dev(f, (double2 *)d_x, (const double2 *)d_b, zero_start != 0, kpm, tol, maxiter, &it, &e);
    *iters = it;
    *eps = e;
    SQ_CATCH
}

// ---- KPMPreconditioner ----------------------------------------------------------------------------
int sq_kpm_create(sq_kpm **out, sq_fdm *f, double rbuf, int64_t n, double a1, double a2) {
    SQ_TRY
    kpm_create_impl(out, f, rbuf, n, a1, a2);
    SQ_CATCH
}
int sq_kpm_set_seed(sq_kpm *k, uint64_t seed) {
    SQ_TRY
    SQ_REQUIRE(k, "NULL handle");
    k->seed = seed;
    k->rng_counter = 0;
    SQ_CATCH
}
int sq_kpm_destroy(sq_kpm *k) {
    SQ_TRY
    if (k) { fdm_sync_if_alive(k->f); delete k; }
    SQ_CATCH
}
int sq_kpm_update(sq_kpm *k, const double *lanczos_start, int *active, double *bounds) {
    SQ_TRY
    SQ_REQUIRE(k, "NULL handle");
    SQ_CUDA(cudaSetDevice(k->f->device));
    kpm_update(k, lanczos_start, nullptr);
    if (active) *active = k->active;
    if (bounds) { bounds[0] = k->bounds[0]; bounds[1] = k->bounds[1]; }
    SQ_CATCH
}
int sq_kpm_set_bounds(sq_kpm *k, double emin, double emax) {
    SQ_TRY
    SQ_REQUIRE(k, "NULL handle");
    SQ_REQUIRE(emin > 0 && emax > emin, "bounds must satisfy 0 < emin < emax");
    SQ_CUDA(cudaSetDevice(k->f->device));
    kpm_set_bounds(k, emin, emax);
    SQ_CATCH
}
int sq_kpm_get_orders(sq_kpm *k, int64_t *ncoef, int64_t *orders) {
    SQ_TRY
    SQ_REQUIRE(k && ncoef, "NULL argument");
    *ncoef = k->ncoef;
    if (orders) for (i64 l = 0; l < k->ncoef; l++) orders[l] = k->order[l];
    SQ_CATCH
}
int sq_kpm_get_coefs(sq_kpm *k, int64_t l, sq_complex *coefs) {
    SQ_TRY
    SQ_REQUIRE(k && coefs && l >= 0 && l < k->ncoef, "bad argument");
    for (size_t q = 0; q < k->coefs[l].size(); q++) { coefs[q].re = k->coefs[l][q].x; coefs[q].im = k->coefs[l][q].y; }
    SQ_CATCH
}
int sq_kpm_ldiv(sq_kpm *k, sq_complex *out, const sq_complex *in) {
    SQ_TRY
    SQ_REQUIRE(k && out && in, "NULL argument");
    sq_fdm *f = k->f;
    SQ_CUDA(cudaSetDevice(f->device));
    fdm_host_to_dev(f, f->r.p, in);
    kpm_ldiv_dev(k, f->z.p, f->r.p);
    fdm_dev_to_host(f, out, f->z.p);
    SQ_CATCH
}
int sq_kpm_ldiv_dev(sq_kpm *k, void *d_out, const void *d_in) {
    SQ_TRY
    SQ_REQUIRE(k && d_out && d_in, "NULL argument");
    SQ_CUDA(cudaSetDevice(k->f->device));
    kpm_ldiv_dev(k, (double2 *)d_out, (const double2 *)d_in);
    SQ_CATCH
}
int sq_kpm_fourier(sq_kpm *k, sq_complex *v, int forward) {
    SQ_TRY
    SQ_REQUIRE(k && v, "NULL argument");
    sq_fdm *f = k->f;
    SQ_CUDA(cudaSetDevice(f->device));
    fdm_host_to_dev(f, f->r.p, v);
    kpm_fourier_dev(k, f->r.p, forward != 0);
    fdm_dev_to_host(f, v, f->r.p);
    SQ_CATCH
}

// ---- electron-phonon model --------------------------------------------------------------------------
int sq_elph_create(sq_elph **out, sq_fdm *f, double dtau, int64_t Nph, const double *Omega, const double *Omega4, const double *M,
                   int64_t Nhol, const int64_t *hol_phonon, const int64_t *hol_site, const double *hol_a, const double *hol_a2,
                   const double *hol_a3, const double *hol_a4, const int32_t *hol_phsym, int64_t Nssh, const int64_t *ssh_phonon,
                   const int64_t *ssh_hopping, const double *ssh_a, const double *ssh_a2, const double *ssh_a3, const double *ssh_a4,
                   const double *V0, const double *t0) {
    SQ_TRY
    elph_create_impl(out, f, dtau, Nph, Omega, Omega4, M, Nhol, hol_phonon, hol_site, hol_a, hol_a2, hol_a3, hol_a4, hol_phsym, Nssh,
                     ssh_phonon, ssh_hopping, ssh_a, ssh_a2, ssh_a3, ssh_a4, V0, t0);
    SQ_CATCH
}
int sq_elph_set_bare(sq_elph *e, const double *V0, const double *t0) {
    SQ_TRY
    SQ_REQUIRE(e, "NULL handle");
    SQ_CUDA(cudaSetDevice(e->f->device));
    elph_set_bare(e, V0, t0);
    SQ_CATCH
}
int sq_elph_set_dispersion(sq_elph *e, int64_t Ndisp, const int64_t *disp_phonon, const double *Omega, const double *Omega4) {
    SQ_TRY
    SQ_REQUIRE(e, "NULL handle");
    SQ_CUDA(cudaSetDevice(e->f->device));
    elph_set_dispersion(e, Ndisp, disp_phonon, Omega, Omega4);
    SQ_CATCH
}
int sq_elph_potential_derivative(sq_elph *e, double *F) {
    SQ_TRY
    SQ_REQUIRE(e && F, "NULL argument");
    sq_fdm *f = e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    const size_t nx = (size_t)f->L * e->Nph;
    DevBuf<double> pm, z;
    pm.alloc(nx); z.alloc(nx);                                  // p = 0, dS = 0:  p <- -1 * (anharmonic + dispersive)
    elph_add_potential_derivative(e, pm.p, z.p, -1.0);
    pm.download(F, nx, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    SQ_CATCH
}
int sq_elph_destroy(sq_elph *e) {
    SQ_TRY
    if (e) { fdm_sync_if_alive(e->f); delete e; }
    SQ_CATCH
}
int sq_elph_set_x(sq_elph *e, const double *x) {
    SQ_TRY
    SQ_REQUIRE(e && x, "NULL argument");
    SQ_CUDA(cudaSetDevice(e->f->device));
    e->x.upload(x, (size_t)e->f->L * e->Nph, e->f->stream);
    SQ_CUDA(cudaStreamSynchronize(e->f->stream));
    SQ_CATCH
}
int sq_elph_get_x(sq_elph *e, double *x) {
    SQ_TRY
    SQ_REQUIRE(e && x, "NULL argument");
    SQ_CUDA(cudaSetDevice(e->f->device));
    e->x.download(x, (size_t)e->f->L * e->Nph, e->f->stream);
    SQ_CUDA(cudaStreamSynchronize(e->f->stream));
    SQ_CATCH
}
// x-mutations of the global moves (reflection_update!, swap_update!, radial_update!): device resident, phonon indices 1-based
// like the Julia arrays.  x is [l][p] on the device.
__global__ void k_x_scale_rows(double *x, int L, int Nph, int p_lo, int p_hi, double factor) {
    const int np = p_hi - p_lo;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)L * np) return;
    const int l = (int)(idx / np), p = p_lo + (int)(idx % np);
    x[(size_t)l * Nph + p] *= factor;
}
__global__ void k_x_swap_rows(double *x, int L, int Nph, int pi, int pj) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const double a = x[(size_t)l * Nph + pi], b = x[(size_t)l * Nph + pj];
    x[(size_t)l * Nph + pi] = b;
    x[(size_t)l * Nph + pj] = a;
}
int sq_elph_scale_x(sq_elph *e, int64_t p_first, int64_t p_last, double factor) {
    SQ_TRY
    SQ_REQUIRE(e, "NULL handle");
    SQ_REQUIRE(p_first >= 1 && p_last >= p_first && p_last <= e->Nph, "phonon range out of bounds");
    sq_fdm *f = e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    const size_t tot = (size_t)f->L * (size_t)(p_last - p_first + 1);
    k_x_scale_rows<<<(unsigned)((tot + 255) / 256), 256, 0, f->stream>>>(e->x.p, (int)f->L, (int)e->Nph, (int)p_first - 1, (int)p_last, factor);
    SQ_LAUNCH_CHECK();
    f->launches++;
    SQ_CATCH
}
int sq_elph_swap_x(sq_elph *e, int64_t p_i, int64_t p_j) {
    SQ_TRY
    SQ_REQUIRE(e, "NULL handle");
    SQ_REQUIRE(p_i >= 1 && p_i <= e->Nph && p_j >= 1 && p_j <= e->Nph, "phonon index out of bounds");
    sq_fdm *f = e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    if (p_i != p_j) {
        k_x_swap_rows<<<(unsigned)((f->L + 127) / 128), 128, 0, f->stream>>>(e->x.p, (int)f->L, (int)e->Nph, (int)p_i - 1, (int)p_j - 1);
        SQ_LAUNCH_CHECK();
        f->launches++;
    }
    SQ_CATCH
}
int sq_elph_backup_x(sq_elph *e) {
    SQ_TRY
    SQ_REQUIRE(e, "NULL handle");
    sq_fdm *f = e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    const size_t n = (size_t)f->L * e->Nph;
    if (e->x_backup.n < n) e->x_backup.alloc(n, false);
    SQ_CUDA(cudaMemcpyAsync(e->x_backup.p, e->x.p, n * sizeof(double), cudaMemcpyDeviceToDevice, f->stream));
    SQ_CATCH
}
int sq_elph_restore_x(sq_elph *e) {
    SQ_TRY
    SQ_REQUIRE(e, "NULL handle");
    sq_fdm *f = e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    const size_t n = (size_t)f->L * e->Nph;
    SQ_REQUIRE(e->x_backup.n >= n, "no backup of the phonon field");
    SQ_CUDA(cudaMemcpyAsync(e->x.p, e->x_backup.p, n * sizeof(double), cudaMemcpyDeviceToDevice, f->stream));
    SQ_CATCH
}
int sq_elph_shift_mu(sq_elph *e, double dmu) {
    SQ_TRY
    SQ_REQUIRE(e, "NULL handle");
    sq_fdm *f = e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    std::vector<double> v(f->N);
    e->V0.download(v.data(), f->N, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    for (auto &a : v) a -= dmu;
    e->V0.upload(v.data(), f->N, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    SQ_CATCH
}
int sq_elph_refresh_fdm(sq_elph *e) {
    SQ_TRY
    SQ_REQUIRE(e, "NULL handle");
    SQ_CUDA(cudaSetDevice(e->f->device));
    elph_refresh_fdm(e);
    SQ_CUDA(cudaStreamSynchronize(e->f->stream));
    SQ_CATCH
}
int sq_elph_get_Vt(sq_elph *e, double *V, double *t) {
    SQ_TRY
    SQ_REQUIRE(e && V && t, "NULL argument");
    sq_fdm *f = e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    elph_build_Vt(e);
    e->V.download(V, (size_t)f->L * f->N, f->stream);
    e->t.download(t, (size_t)f->L * f->Nh, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    SQ_CATCH
}
int sq_elph_bosonic_action(sq_elph *e, double *Sb) {
    SQ_TRY
    SQ_REQUIRE(e && Sb, "NULL argument");
    SQ_CUDA(cudaSetDevice(e->f->device));
    *Sb = elph_bosonic_action(e);
    SQ_CATCH
}

// ---- PFFCalculator ------------------------------------------------------------------------------------
int sq_pff_create(sq_pff **out, sq_elph *e) {
    SQ_TRY
    pff_create_impl(out, e);
    SQ_CATCH
}
int sq_pff_set_seed(sq_pff *p, uint64_t seed) {
    SQ_TRY
    SQ_REQUIRE(p, "NULL handle");
    p->seed = seed;
    p->seed_set = true;
    p->rng_counter = 0;
    SQ_CATCH
}
int sq_pff_destroy(sq_pff *p) {
    SQ_TRY
    if (p) { fdm_sync_if_alive(p->owner); delete p; }
    SQ_CATCH
}
int sq_pff_set_exact_holstein(sq_pff *p, int flag) {
    SQ_TRY
    SQ_REQUIRE(p, "NULL handle");
    p->exact_holstein = flag ? 1 : 0;
    SQ_CATCH
}
int sq_pff_sample(sq_pff *p, const sq_complex *R, double *Sf) {
    SQ_TRY
    SQ_REQUIRE(p && Sf, "NULL argument");
    SQ_CUDA(cudaSetDevice(p->e->f->device));
    pff_fill_phi_normals(p, R, nullptr);
    *Sf = pff_sample_dev(p);
    SQ_CATCH
}
int sq_pff_action(sq_pff *p, sq_kpm *kpm, const double *lanczos_start, double tol, int64_t maxiter, double *Sf, int64_t *iters,
                  double *eps) {
    SQ_TRY
    SQ_REQUIRE(p && Sf && iters && eps, "NULL argument");
    SQ_CUDA(cudaSetDevice(p->e->f->device));
    i64 it = 0;
    double e = 0;
    *Sf = pff_action_dev(p, kpm, kpm != nullptr, lanczos_start, nullptr, tol, maxiter, &it, &e);
    *iters = it;
    *eps = e;
    SQ_CATCH
}
int sq_pff_force(sq_pff *p, double *dSdx, sq_kpm *kpm, const double *lanczos_start, double tol, int64_t maxiter, double *Sf,
                 int64_t *iters, double *eps) {
    SQ_TRY
    SQ_REQUIRE(p && dSdx && Sf && iters && eps, "NULL argument");
    sq_fdm *f = p->e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    size_t nF = (size_t)f->L * p->e->Nph;
    p->F.upload(dSdx, nF, f->stream);
    i64 it = 0;
    double e = 0;
    *Sf = pff_force_dev(p, kpm, kpm != nullptr, lanczos_start, nullptr, tol, maxiter, &it, &e);
    p->F.download(dSdx, nF, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    *iters = it;
    *eps = e;
    SQ_CATCH
}
int sq_pff_get_fields(sq_pff *p, sq_complex *Phi, sq_complex *Psi, double *Lambda) {
    SQ_TRY
    SQ_REQUIRE(p, "NULL handle");
    sq_fdm *f = p->e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    if (Phi) fdm_dev_to_host(f, Phi, p->Phi.p);
    if (Psi) fdm_dev_to_host(f, Psi, p->u.p);
    if (Lambda) {
        fdm_transpose_real(f, f->iod1.p, p->Lam.p, (int)f->N, (int)f->L, true);
        f->iod1.download(Lambda, (size_t)f->L * f->N, f->stream);
        SQ_CUDA(cudaStreamSynchronize(f->stream));
    }
    SQ_CATCH
}
int sq_pff_set_Phi(sq_pff *p, const sq_complex *Phi) {
    SQ_TRY
    SQ_REQUIRE(p && Phi, "NULL argument");
    SQ_CUDA(cudaSetDevice(p->e->f->device));
    fdm_host_to_dev(p->e->f, p->Phi.p, Phi);
    SQ_CUDA(cudaStreamSynchronize(p->e->f->stream));
    SQ_CATCH
}
int sq_pff_lambda_op(sq_pff *p, int which, sq_complex *out, const sq_complex *in) {
    SQ_TRY
    SQ_REQUIRE(p && out && in && which >= 0 && which <= 3, "bad argument");
    sq_fdm *f = p->e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    elph_update_lambda(p->e, p->Lam.p);
    fdm_host_to_dev(f, p->w1.p, in);
    elph_lambda_op(p->e, which, p->w2.p, p->w1.p, p->Lam.p);
    fdm_dev_to_host(f, out, p->w2.p);
    SQ_CATCH
}
int sq_pff_dM_dx(sq_pff *p, double *F, double nu, const sq_complex *u, const sq_complex *v) {
    SQ_TRY
    SQ_REQUIRE(p && F && u && v, "NULL argument");
    SQ_CUDA(cudaSetDevice(p->e->f->device));
    pff_dM_dx_host(p, F, nu, u, v);
    SQ_CATCH
}
int sq_pff_dLambda_dx(sq_pff *p, double *F, double nu, const sq_complex *up, const sq_complex *u) {
    SQ_TRY
    SQ_REQUIRE(p && F && up && u, "NULL argument");
    SQ_CUDA(cudaSetDevice(p->e->f->device));
    pff_dLambda_dx_host(p, F, nu, up, u);
    SQ_CATCH
}

// ---- EFAPFFHMCUpdater ---------------------------------------------------------------------------------
int sq_hmc_create(sq_hmc **out, sq_pff *p, int64_t Nt, double dt, double eta, double delta, uint64_t seed) {
    SQ_TRY
    hmc_create_impl(out, p, Nt, dt, eta, delta, seed);
    SQ_CATCH
}
int sq_hmc_destroy(sq_hmc *h) {
    SQ_TRY
    if (h) { fdm_sync_if_alive(h->owner); delete h; }
    SQ_CATCH
}
int sq_hmc_update(sq_hmc *h, sq_kpm *kpm, double tol_action, double tol_force, int64_t maxiter, const double *randoms,
                  int64_t nrandoms, int *accepted, double *info) {
    SQ_TRY
    SQ_REQUIRE(h && accepted, "NULL argument");
    *accepted = hmc_update_impl(h, kpm, tol_action, tol_force, maxiter, randoms, nrandoms, info);
    SQ_CATCH
}
int sq_hmc_set_seed(sq_hmc *h, uint64_t seed) {
    SQ_TRY
    SQ_REQUIRE(h, "NULL handle");
    h->seed = seed;
    h->counter = 0;
    SQ_CATCH
}
const char *sq_hmc_last_reject(sq_hmc *h) { return h ? h->last_reject.c_str() : ""; }
int sq_hmc_init_momentum(sq_hmc *h, const double *R, double *p, double *K) {
    SQ_TRY
    SQ_REQUIRE(h && R && p && K, "NULL argument");
    sq_fdm *f = h->p->e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    size_t nx = (size_t)f->L * h->p->e->Nph;
    h->dS.upload(R, nx, f->stream);
    *K = hmc_init_momentum_dev(h, h->dS.p, h->pm.p);
    h->pm.download(p, nx, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    SQ_CATCH
}
int sq_hmc_kinetic(sq_hmc *h, const double *p, double *K) {
    SQ_TRY
    SQ_REQUIRE(h && p && K, "NULL argument");
    sq_fdm *f = h->p->e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    h->dS.upload(p, (size_t)f->L * h->p->e->Nph, f->stream);
    *K = hmc_kinetic_dev(h, h->dS.p);
    SQ_CATCH
}
int sq_hmc_evolve(sq_hmc *h, double *x, double *p, double dt) {
    SQ_TRY
    SQ_REQUIRE(h && x && p, "NULL argument");
    sq_fdm *f = h->p->e->f;
    SQ_CUDA(cudaSetDevice(f->device));
    size_t nx = (size_t)f->L * h->p->e->Nph;
    h->x0.upload(x, nx, f->stream);
    h->dS.upload(p, nx, f->stream);
    hmc_evolve_dev(h, h->x0.p, h->dS.p, dt);
    h->x0.download(x, nx, f->stream);
    h->dS.download(p, nx, f->stream);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    SQ_CATCH
}

// ---- GreensEstimator ----------------------------------------------------------------------------------
int sq_greens_create(sq_greens **out, sq_fdm *f, int64_t Nrv, uint64_t seed) {
    SQ_TRY
    greens_create_impl(out, f, Nrv, seed);
    SQ_CATCH
}
int sq_greens_set_seed(sq_greens *g, uint64_t seed) {
    SQ_TRY
    SQ_REQUIRE(g, "NULL handle");
    g->seed = seed;
    g->counter = 0;
    SQ_CATCH
}
int sq_greens_destroy(sq_greens *g) {
    SQ_TRY
    if (g) { fdm_sync_if_alive(g->f); delete g; }
    SQ_CATCH
}
int sq_greens_update(sq_greens *g, sq_kpm *kpm, const sq_complex *R, double tol, int64_t maxiter, double *avg_iters) {
    SQ_TRY
    SQ_REQUIRE(g && avg_iters, "NULL argument");
    *avg_iters = greens_update_impl(g, kpm, R, tol, maxiter);
    SQ_CATCH
}
int sq_greens_get(sq_greens *g, sq_complex *R, sq_complex *GR) {
    SQ_TRY
    SQ_REQUIRE(g, "NULL handle");
    sq_fdm *f = g->f;
    SQ_CUDA(cudaSetDevice(f->device));
    size_t V = (size_t)f->L * f->N;
    for (i64 n = 0; n < g->Nrv; n++) {
        if (R) fdm_dev_to_host(f, R + n * V, g->R.p + n * V);
        if (GR) fdm_dev_to_host(f, GR + n * V, g->GR.p + n * V);
    }
    SQ_CATCH
}
int sq_greens_set_GR(sq_greens *g, const sq_complex *GR) {
    SQ_TRY
    SQ_REQUIRE(g && GR, "NULL argument");
    sq_fdm *f = g->f;
    SQ_CUDA(cudaSetDevice(f->device));
    size_t V = (size_t)f->L * f->N;
    for (i64 n = 0; n < g->Nrv; n++) fdm_host_to_dev(f, g->GR.p + n * V, GR + n * V);
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    SQ_CATCH
}
int sq_greens_measure_GD0(sq_greens *g, int norb, int ndim, const int64_t *dims, int a, int b, sq_complex *out) {
    SQ_TRY
    SQ_REQUIRE(g && out, "NULL argument");
    greens_measure_GD0_impl(g, norb, ndim, dims, a, b, out);
    SQ_CATCH
}
int sq_greens_measure_contraction(sq_greens *g, int kind, int norb, int ndim, const int64_t *dims, const int *orbitals, const int64_t *r,
                                  sq_complex *out) {
    SQ_TRY
    SQ_REQUIRE(g && out, "NULL argument");
    greens_measure_c4_impl(g, kind, norb, ndim, dims, orbitals, r, out, nullptr, nullptr);
    SQ_CATCH
}
int sq_greens_measure_contraction_weighted(sq_greens *g, int kind, int norb, int ndim, const int64_t *dims, const int *orbitals, const int64_t *r,
                                           const double *tD, const double *t0, sq_complex *out) {
    SQ_TRY
    SQ_REQUIRE(g && out, "NULL argument");
    greens_measure_c4_impl(g, kind, norb, ndim, dims, orbitals, r, out, tD, t0);
    SQ_CATCH
}
int sq_greens_measure_double_occ_orbital(sq_greens *g, int norb, int a, sq_complex *d) {
    SQ_TRY
    SQ_REQUIRE(g && d, "NULL argument");
    greens_measure_double_occ_orbital_impl(g, norb, a, (double *)d);
    SQ_CATCH
}
int sq_greens_measure_n_orbital(sq_greens *g, int norb, int a, sq_complex *n) {
    SQ_TRY
    SQ_REQUIRE(g && n, "NULL argument");
    greens_measure_n_orbital_impl(g, norb, a, (double *)n);
    SQ_CATCH
}
int sq_greens_weighted_density(sq_greens *g, const double *w, sq_complex *out) {
    SQ_TRY
    SQ_REQUIRE(g && w && out, "NULL argument");
    greens_weighted_density_impl(g, w, (double *)out);
    SQ_CATCH
}
int sq_greens_weighted_bonds(sq_greens *g, int64_t nbonds, const int64_t *bonds, const sq_complex *w, sq_complex *out) {
    SQ_TRY
    SQ_REQUIRE(g && out, "NULL argument");
    greens_weighted_bonds_impl(g, nbonds, bonds, w, (double *)out);
    SQ_CATCH
}
int sq_greens_measure(sq_greens *g, sq_complex *n, sq_complex *double_occ, sq_complex *Nsqrd) {
    SQ_TRY
    SQ_REQUIRE(g, "NULL handle");
    double out[6];
    greens_measure_impl(g, out);
    if (n) { n->re = out[0]; n->im = out[1]; }
    if (double_occ) { double_occ->re = out[2]; double_occ->im = out[3]; }
    if (Nsqrd) { Nsqrd->re = out[4]; Nsqrd->im = out[5]; }
    SQ_CATCH
}

}   // extern "C"

// greens.cu -- GreensEstimator solves and the scalar measurements built on them.
//
// Replaces update_greens_estimator! (src/Measurements/GreensEstimator.jl:125-175) and measure_n,
// measure_Nsqrd, measure_double_occ (src/Measurements/scalar_measurements.jl:2-147).  R and G R stay on the
// device ([n][l][i]); the measurements are single-pass reductions over them.
#include "sq_internal.h"

#include <cmath>
#include <cstring>
#include <map>
#include <memory>

// R <- R / |R| (unit-modulus random phases, GreensEstimator.jl:141-142)
__global__ void k_unit_modulus(double2 *__restrict__ v, size_t n) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        double2 a = v[k];
        double r = sqrt(a.x * a.x + a.y * a.y);
        v[k] = make_double2(a.x / r, a.y / r);
    }
}
// row i of the cross-dot matrix D[i][j] = sum_r conj(R_i[r]) GR_j[r]; partials part[((blk*Nrv + i)*Nrv + j)*2 + {0,1}]
__global__ void k_cross_dots(const double2 *__restrict__ R, const double2 *__restrict__ GR, size_t V, int Nrv, double *__restrict__ part) {
    __shared__ double red[2 * 32];
    int i = blockIdx.y;
    const double2 *Ri = R + (size_t)i * V;
    for (int j = 0; j < Nrv; j++) {
        const double2 *Gj = GR + (size_t)j * V;
        double v[2] = {0, 0};
        for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < V; k += (size_t)gridDim.x * blockDim.x) {
            double2 a = Ri[k], b = Gj[k];
            v[0] += a.x * b.x + a.y * b.y;
            v[1] += a.x * b.y - a.y * b.x;
        }
        block_sum<2>(v, red);
        if (threadIdx.x == 0) {
            size_t o = (((size_t)blockIdx.x * Nrv + i) * Nrv + j) * 2;
            part[o] = v[0];
            part[o + 1] = v[1];
        }
        __syncthreads();
    }
}
// double occupancy: sum_r [ (sum_i a_i)^2 - sum_i a_i^2 ] / 2 with a_i = 1 - GR_i conj(R_i)  (pairs i<j of :130-145)
// norb > 0: only the sites of orbital `orb` (site % norb == orb) contribute (measure_double_occ(greens_estimator, orbital), :98-109)
__global__ void k_double_occ(const double2 *__restrict__ R, const double2 *__restrict__ GR, size_t V, int Nrv, double *__restrict__ part,
                             int N = 1, int norb = 0, int orb = 0) {
    __shared__ double red[2 * 32];
    double v[2] = {0, 0};
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < V; k += (size_t)gridDim.x * blockDim.x) {
        if (norb > 0 && (int)((k % N) % norb) != orb) continue;
        double2 s1 = make_double2(0, 0), s2 = make_double2(0, 0);
        for (int i = 0; i < Nrv; i++) {
            double2 r = R[(size_t)i * V + k], g = GR[(size_t)i * V + k];
            double2 a = make_double2(1.0 - (g.x * r.x + g.y * r.y), -(g.y * r.x - g.x * r.y));   // 1 - g conj(r)
            s1 = cadd(s1, a);
            s2 = cadd(s2, cmul(a, a));
        }
        double2 t = csub(cmul(s1, s1), s2);
        v[0] += 0.5 * t.x;
        v[1] += 0.5 * t.y;
    }
    block_sum<2>(v, red);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = v[0]; part[2 * blockIdx.x + 1] = v[1]; }
}

void greens_create_impl(sq_greens **out, sq_fdm *f, i64 Nrv, uint64_t seed) {
    SQ_REQUIRE(out && f, "NULL argument");
    SQ_REQUIRE(Nrv >= 1 && Nrv <= 64, "Nrv out of range");
    SQ_CUDA(cudaSetDevice(f->device));
    sq_greens *g = new sq_greens();
    try {
        g->f = f; g->Nrv = Nrv; g->seed = seed;
        size_t V = (size_t)f->L * f->N;
        g->R.alloc(V * Nrv); g->GR.alloc(V * Nrv); g->MtR.alloc(V);
        g->part.alloc((size_t)2 * 64 * Nrv * Nrv + 2 * SQ_MAXPART);
    } catch (...) {
        delete g;
        throw;
    }
    *out = g;
}

// h_R: (V x Nrv) host array in the reference layout, or NULL for library randoms
double greens_update_impl(sq_greens *g, sq_kpm *kpm, const void *h_R, double tol, i64 maxiter) {
    sq_fdm *f = g->f;
    SQ_CUDA(cudaSetDevice(f->device));
    size_t V = (size_t)f->L * f->N;
    if (h_R) {
        for (i64 n = 0; n < g->Nrv; n++) fdm_host_to_dev(f, g->R.p + n * V, (const char *)h_R + n * V * sizeof(double2));
    } else {
        rng_fill_normal((double *)g->R.p, 2 * V * g->Nrv, g->seed, sq_rng_stream(SQ_RNG_GREENS, g->counter++), f->stream);
        k_unit_modulus<<<f->num_sms * 4, 256, 0, f->stream>>>(g->R.p, V * g->Nrv);
        SQ_LAUNCH_CHECK();
        f->launches++;
    }
    if (kpm) kpm_update(kpm, nullptr, nullptr);                            // GreensEstimator.jl:150
    double avg = 0;
    if (f->sharded && f->world > 1 && !getenv("SQ_GREENS_NO_RHS_SPLIT")) {
        // One chain over several GPUs: the Nrv systems are independent, so they are DISTRIBUTED over the ranks (column j on rank
        // j % world, solved there on the full lattice with the single-GPU solvers -- batched if preconditioned) instead of
        // tau-slab partitioning each of them; the solutions are then broadcast.  Every rank holds the same R (same seed) and the
        // same previous G R (warm start), and ends with the same G R.
        const int W = f->world, me = f->rank;
        std::vector<i64> mine;
        for (i64 n = 0; n < g->Nrv; n++) if (n % W == me) mine.push_back(n);
        const int nm = (int)mine.size();
        if (g->MtRb.n < V * g->Nrv) g->MtRb.alloc(V * g->Nrv, false);
        if (g->Xb.n < V * g->Nrv) g->Xb.alloc(V * g->Nrv, false);
        for (int q = 0; q < nm; q++) {
            fdm_mul_dev(f, SQ_OP_MT, g->MtRb.p + (size_t)q * V, g->R.p + mine[q] * V);
            SQ_CUDA(cudaMemcpyAsync(g->Xb.p + (size_t)q * V, g->GR.p + mine[q] * V, V * sizeof(double2), cudaMemcpyDeviceToDevice, f->stream));
        }
        std::vector<i64> its(std::max(nm, 1), 0);
        std::vector<double> epss(std::max(nm, 1), 0.0);
        f->force_local = 1;
        try {
            if (nm > 0 && fdm_cg_batch_applicable(f, kpm, nm)) fdm_cg_batch_dev(f, g->Xb.p, g->MtRb.p, nm, false, kpm, tol, maxiter, its.data(), epss.data());
            else for (int q = 0; q < nm; q++) fdm_cg_dev(f, g->Xb.p + (size_t)q * V, g->MtRb.p + (size_t)q * V, false, kpm, tol, maxiter, &its[q], &epss[q]);
        } catch (...) { f->force_local = 0; throw; }
        f->force_local = 0;
        double sum = 0;
        for (int q = 0; q < nm; q++) {
            SQ_CUDA(cudaMemcpyAsync(g->GR.p + mine[q] * V, g->Xb.p + (size_t)q * V, V * sizeof(double2), cudaMemcpyDeviceToDevice, f->stream));
            sum += (double)its[q];
        }
        slab_broadcast_columns(f, g->GR.p, V, (int)g->Nrv);
        SQ_CUDA(cudaMemcpyAsync(f->scal.p, &sum, sizeof(double), cudaMemcpyHostToDevice, f->stream));
        fdm_allreduce_sum(f, f->scal.p, 1);
        SQ_CUDA(cudaMemcpyAsync(&sum, f->scal.p, sizeof(double), cudaMemcpyDeviceToHost, f->stream));
        SQ_CUDA(cudaStreamSynchronize(f->stream));
        return sum / (double)g->Nrv;
    }
    if (fdm_cg_batch_applicable(f, kpm, (int)g->Nrv)) {
        // all Nrv systems in lock step (cg_batch.cu): the same recurrence, warm start and iteration count per system as the loop below
        if (g->MtRb.n < V * g->Nrv) g->MtRb.alloc(V * g->Nrv, false);
        for (i64 n = 0; n < g->Nrv; n++) fdm_mul_dev(f, SQ_OP_MT, g->MtRb.p + n * V, g->R.p + n * V);
        std::vector<i64> its(g->Nrv);
        std::vector<double> epss(g->Nrv);
        fdm_cg_batch_dev(f, g->GR.p, g->MtRb.p, (int)g->Nrv, false, kpm, tol, maxiter, its.data(), epss.data());
        for (i64 n = 0; n < g->Nrv; n++) avg += (double)its[n];
        return avg / (double)g->Nrv;
    }
    for (i64 n = 0; n < g->Nrv; n++) {
        fdm_mul_dev(f, SQ_OP_MT, g->MtR.p, g->R.p + n * V);                 // :156
        i64 it = 0;
        double eps = 0;
        fdm_cg_dev(f, g->GR.p + n * V, g->MtR.p, false, kpm, tol, maxiter, &it, &eps);   // :159-165, warm start
        avg += (double)it;
    }
    return avg / (double)g->Nrv;
}

// out[0..1] = n, out[2..3] = double occupancy, out[4..5] = <N^2>
void greens_measure_impl(sq_greens *g, double *out) {
    sq_fdm *f = g->f;
    SQ_CUDA(cudaSetDevice(f->device));
    size_t V = (size_t)f->L * f->N;
    int Nrv = (int)g->Nrv, nb = 64;
    double L = (double)f->L;
    k_cross_dots<<<dim3(nb, Nrv), 256, 0, f->stream>>>(g->R.p, g->GR.p, V, Nrv, g->part.p);
    double *pd = g->part.p + (size_t)2 * nb * Nrv * Nrv;
    k_double_occ<<<SQ_MAXPART / 4, 256, 0, f->stream>>>(g->R.p, g->GR.p, V, Nrv, pd);
    SQ_LAUNCH_CHECK();
    f->launches += 2;
    std::vector<double> h((size_t)2 * nb * Nrv * Nrv), hd(2 * (SQ_MAXPART / 4));
    SQ_CUDA(cudaMemcpyAsync(h.data(), g->part.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
    SQ_CUDA(cudaMemcpyAsync(hd.data(), pd, hd.size() * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    std::vector<double> Dre(Nrv * Nrv, 0.0), Dim(Nrv * Nrv, 0.0);
    for (int b = 0; b < nb; b++)
        for (int k = 0; k < Nrv * Nrv; k++) { Dre[k] += h[((size_t)b * Nrv * Nrv + k) * 2]; Dim[k] += h[((size_t)b * Nrv * Nrv + k) * 2 + 1]; }
    auto D = [&](int i, int j, double &re, double &im) { re = Dre[i * Nrv + j]; im = Dim[i * Nrv + j]; };
    // measure_n: n = 1 - dot(R, GR)/length(R)   (:15-28)
    double tr = 0, ti = 0;
    for (int i = 0; i < Nrv; i++) { tr += Dre[i * Nrv + i]; ti += Dim[i * Nrv + i]; }
    out[0] = 1.0 - tr / ((double)V * Nrv);
    out[1] = -ti / ((double)V * Nrv);
    // measure_double_occ (:112-147)
    double dr = 0, di = 0;
    for (int b = 0; b < SQ_MAXPART / 4; b++) { dr += hd[2 * b]; di += hd[2 * b + 1]; }
    double npairs = 0.5 * Nrv * (Nrv - 1);
    out[2] = Nrv > 1 ? dr / ((double)V * npairs) : 0.0;
    out[3] = Nrv > 1 ? di / ((double)V * npairs) : 0.0;
    // measure_Nsqrd (:31-96)
    double Nbr = 0, Nbi = 0, T2r = 0, T2i = 0;
    for (int i = 0; i < Nrv - 1; i++) {
        double ir, ii;
        D(i, i, ir, ii);
        for (int j = i + 1; j < Nrv; j++) {
            double jr, ji, ar, ai, br, bi;
            D(j, j, jr, ji);
            // 4 (V - TrGi)(V - TrGj) / L^2
            double xr = (double)V - ir, xi = -ii, yr = (double)V - jr, yi = -ji;
            Nbr += 4 * (xr * yr - xi * yi) / (L * L);
            Nbi += 4 * (xr * yi + xi * yr) / (L * L);
            D(j, i, ar, ai);      // dot(Rj, GRi)
            D(i, j, br, bi);      // dot(Ri, GRj)
            T2r += (ar * br - ai * bi) / (L * L);
            T2i += (ar * bi + ai * br) / (L * L);
        }
    }
    if (Nrv > 1) { Nbr /= npairs; Nbi /= npairs; T2r /= npairs; T2i /= npairs; }
    double TrGr = tr / (Nrv * L), TrGi = ti / (Nrv * L);
    out[4] = Nbr + 2 * TrGr / L - 2 * T2r;
    out[5] = Nbi + 2 * TrGi / L - 2 * T2i;
}


// ---------------------------------------------------------------------------------------------------
// Time-displaced Green's function G_ab(Delta) = G(r + Delta_r, tau + Delta_tau | r, tau), averaged over translations:
// measure_GΔ0! (src/Measurements/GreensEstimator.jl:177-233) with its helpers _aperiodic_copyto! (:656-671),
// _translational_average! (:674-705).  For every random vector the G R_a and conj(R)_b fields are extended
// antiperiodically to 2 Ltau slices, cross-correlated over the (D+1)-dimensional space-time torus by FFT, and averaged;
// the tau = beta slice is -G(tau = 0) (+ 1 at zero displacement for equal orbitals).
// The (D+1)-dimensional transform is the library's own tau-FFT kernel (fft.cu) applied to the outermost axis, followed by a
// transpose that rotates the axes -- D+1 times.  The products are accumulated in frequency space, so there is one inverse
// transform per orbital pair instead of one per random vector.
// ---------------------------------------------------------------------------------------------------
__global__ void k_gd0_fill(double2 *__restrict__ A, double2 *__restrict__ B, const double2 *__restrict__ GR, const double2 *__restrict__ R,
                           int Lt, int N, int Nc, int norb, int a, int b) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)Lt * Nc) return;
    const int l = (int)(idx / Nc), c = (int)(idx % Nc);
    const double2 g = GR[(size_t)l * N + a + norb * c], r = R[(size_t)l * N + b + norb * c];
    A[idx] = g;
    A[idx + (size_t)Lt * Nc] = make_double2(-g.x, -g.y);
    B[idx] = make_double2(r.x, -r.y);                              // Rt = conj(R)
    B[idx + (size_t)Lt * Nc] = make_double2(-r.x, r.y);
}
__global__ void k_cmul_acc(double2 *__restrict__ C, const double2 *__restrict__ A, const double2 *__restrict__ B, size_t n) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) C[k] = cadd(C[k], cmul(A[k], B[k]));
}
// src [rows][cols] (cols fastest) -> dst [cols][rows]
__global__ void k_transpose_c(double2 *__restrict__ dst, const double2 *__restrict__ src, int rows, int cols) {
    __shared__ double2 tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int r = r0 + dy, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[dy][threadIdx.x] = src[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int c = c0 + dy, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(size_t)c * rows + r] = tile[threadIdx.x][dy];
    }
}
// out: (Ltau + 1, cells) tau fastest.  s: [tau'][cell]
__global__ void k_gd0_finish(double2 *__restrict__ out, const double2 *__restrict__ s, int Lt, int Nc, double scale, int same_orbital) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)(Lt + 1) * Nc) return;
    const int c = (int)(idx / (Lt + 1)), t = (int)(idx % (Lt + 1));
    double2 v;
    if (t < Lt) {
        v = cscale(scale, s[(size_t)t * Nc + c]);
    } else {
        v = cscale(-scale, s[c]);                                   // G(r, beta) = delta(r) - G(r, 0)
        if (same_orbital && c == 0) v.x += 1.0;
    }
    out[idx] = v;
}

// (D+1)-dimensional DFT of buf (axes outermost -> innermost: dims[0..nd)), each axis scaled by 1/sqrt(length); tmp is scratch.
// On return the data is in *buf (the pointers may have been exchanged).
static void greens_fftnd(sq_greens *g, double2 **buf, double2 **tmp, const std::vector<int> &dims, bool inverse) {
    sq_fdm *f = g->f;
    size_t M = 1;
    for (int d : dims) M *= (size_t)d;
    for (size_t ax = 0; ax < dims.size(); ax++) {
        const int L = dims[ax];
        const size_t ncol = M / L;
        if (L == 1) continue;                                       // [1][M] -> [M][1]: the rotation is the identity
        auto it = g->fft_tw.find(L);
        if (it == g->fft_tw.end()) {
            std::vector<double2> tw;
            fft_make_twiddles(L, tw);
            auto ins = g->fft_tw.emplace(L, std::unique_ptr<DevBuf<double2>>(new DevBuf<double2>()));
            it = ins.first;
            it->second->from_vector(tw, f->stream);
        }
        std::vector<int> rad;
        fft_radices(L, rad);
        SQ_REQUIRE(ncol < ((size_t)1 << 31), "correlation array too large");
        tau_fft_launch(f->stream, rad, L, (int)ncol, *buf, *buf, inverse, false, it->second->p, nullptr, nullptr, nullptr, nullptr, nullptr,
                       f->smem_optin);
        dim3 grid((unsigned)((ncol + 31) / 32), (unsigned)((L + 31) / 32));
        k_transpose_c<<<grid, dim3(32, 8), 0, f->stream>>>(*tmp, *buf, L, (int)ncol);
        SQ_LAUNCH_CHECK();
        f->launches += 2;
        std::swap(*buf, *tmp);
    }
}

// h_out: (Ltau + 1) x cells complex, tau fastest (the CΔ0 array of the reference).  dims: unit cells per direction, first fastest.
void greens_measure_GD0_impl(sq_greens *g, int norb, int ndim, const i64 *dims, int a, int b, void *h_out) {
    sq_fdm *f = g->f;
    SQ_CUDA(cudaSetDevice(f->device));
    SQ_REQUIRE(norb >= 1 && ndim >= 1 && ndim <= 3 && dims && h_out, "bad argument");
    SQ_REQUIRE(a >= 1 && a <= norb && b >= 1 && b <= norb, "orbital index out of range");
    size_t Nc = 1;
    for (int d = 0; d < ndim; d++) { SQ_REQUIRE(dims[d] >= 1, "bad lattice dimension"); Nc *= (size_t)dims[d]; }
    SQ_REQUIRE((i64)(Nc * norb) == f->N, "unit cells x orbitals does not match the number of sites");
    const int Lt = (int)f->L;
    const size_t M = (size_t)2 * Lt * Nc, V = (size_t)f->L * f->N;
    if (g->wa.n < M) { g->wa.alloc(M, false); g->wb.alloc(M, false); g->wc.alloc(M, false); g->wt.alloc(M, false); }
    std::vector<int> ax;                                            // outermost -> innermost: tau', then the slowest lattice direction ...
    ax.push_back(2 * Lt);
    for (int d = ndim - 1; d >= 0; d--) ax.push_back((int)dims[d]);
    SQ_CUDA(cudaMemsetAsync(g->wc.p, 0, M * sizeof(double2), f->stream));
    double2 *A = g->wa.p, *B = g->wb.p, *C = g->wc.p, *T = g->wt.p;
    const size_t half = (size_t)Lt * Nc;
    for (i64 n = 0; n < g->Nrv; n++) {
        k_gd0_fill<<<(unsigned)((half + 255) / 256), 256, 0, f->stream>>>(A, B, g->GR.p + n * V, g->R.p + n * V, Lt, (int)f->N, (int)Nc, norb, a - 1, b - 1);
        SQ_LAUNCH_CHECK();
        greens_fftnd(g, &A, &T, ax, false);
        greens_fftnd(g, &B, &T, ax, true);
        k_cmul_acc<<<f->num_sms * 4, 256, 0, f->stream>>>(C, A, B, M);
        SQ_LAUNCH_CHECK();
        f->launches += 2;
    }
    greens_fftnd(g, &C, &T, ax, true);
    // the library's transforms carry 1/sqrt(M) each; the reference is ifft(fft(A) .* ifft(B)) with FFTW's 1/M in ifft
    const double scale = 1.0 / (std::sqrt((double)M) * (double)g->Nrv);
    const size_t nout = (size_t)(Lt + 1) * Nc;
    double2 *dout = (C == g->wa.p || T == g->wa.p) ? nullptr : nullptr;
    (void)dout;
    double2 *O = (A != C && A != T) ? A : B;                        // any work buffer that is not C or the scratch T
    k_gd0_finish<<<(unsigned)((nout + 255) / 256), 256, 0, f->stream>>>(O, C, Lt, (int)Nc, scale, a == b ? 1 : 0);
    SQ_LAUNCH_CHECK();
    f->launches++;
    SQ_CUDA(cudaMemcpyAsync(h_out, O, nout * sizeof(double2), cudaMemcpyDeviceToHost, f->stream));
    SQ_CUDA(cudaStreamSynchronize(f->stream));
}


// ---------------------------------------------------------------------------------------------------
// Four-point contractions built from pairs of random vectors (src/Measurements/GreensEstimator.jl:236-652):
//   kind 0  measure_GΔ0_GΔ0!   G(a,i+r+r1,τ | b,i+r2,0) G(c,i+r+r3,τ | d,i+r4,0)     :236-388
//   kind 1  measure_GΔΔ_G00!   G(a,i+r+r1,τ | b,i+r+r2,τ) G(c,i+r3,0 | d,i+r4,0)     :391-467
//   kind 2  measure_G0Δ_GΔ0!   G(a,i+r1,0 | b,i+r+r2,τ) G(c,i+r+r3,τ | d,i+r4,0)     :470-606
// each = average over pairs n < m of the translational average (_measure_CΔ0!, :610-652, periodic Ltau x L... torus) of two
// element-wise products of G R / conj(R) fields, plus delta-function terms at τ = 0 / β.  Without hopping weights (tΔ, t0 = nothing)
// that covers the density, pair, spin and bond correlations; with them the current correlation (Correlations/current.jl).  As for G(Δ,0) the products are accumulated in frequency space.
// ---------------------------------------------------------------------------------------------------
struct C4Field { const double2 *v; int orb, conj, sh[3]; };
struct C4Geom { int Lt, N, norb, nd, d[3]; size_t Nc; };

__device__ __forceinline__ double2 c4_value(const C4Field &f, const C4Geom &G, int l, size_t c) {
    // value at cell c displaced by sh (the reference's circshift by -r: result[i] = field[i + r])
    size_t cs = 0, mul = 1;
    size_t rem = c;
    for (int k = 0; k < G.nd; k++) {
        int ck = (int)(rem % G.d[k]);
        rem /= G.d[k];
        int q = (ck + f.sh[k]) % G.d[k];
        if (q < 0) q += G.d[k];
        cs += (size_t)q * mul;
        mul *= G.d[k];
    }
    double2 x = f.v[(size_t)l * G.N + f.orb + G.norb * cs];
    if (f.conj) x.y = -x.y;
    return x;
}
// tD, t0: optional real hopping weights (Ltau x cells, tau fastest) of _measure_CΔ0! (:626-646)
__global__ void k_c4_fill(double2 *__restrict__ X, double2 *__restrict__ Y, const C4Field f1, const C4Field f2, const C4Field f3, const C4Field f4,
                          const C4Geom G, const double *__restrict__ tD, const double *__restrict__ t0) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)G.Lt * G.Nc) return;
    const int l = (int)(idx / G.Nc);
    const size_t c = idx % G.Nc;
    double2 x = cmul(c4_value(f1, G, l, c), c4_value(f2, G, l, c));
    double2 y = cmul(c4_value(f3, G, l, c), c4_value(f4, G, l, c));
    if (tD) x = cscale(tD[(size_t)l + (size_t)G.Lt * c], x);
    if (t0) y = cscale(t0[(size_t)l + (size_t)G.Lt * c], y);
    X[idx] = x;
    Y[idx] = y;
}
// the weight tD displaced by sh cells (circshift by -sh), times t0, at (l, c); 1 without weights
struct C4Weight { const double *tD, *t0; int sh[3]; };
__device__ __forceinline__ double c4_weight(const C4Weight &w, const C4Geom &G, int l, size_t c) {
    if (!w.tD) return 1.0;
    size_t cs = 0, mul = 1, rem = c;
    for (int k = 0; k < G.nd; k++) {
        int ck = (int)(rem % G.d[k]);
        rem /= G.d[k];
        int q = (ck + w.sh[k]) % G.d[k];
        if (q < 0) q += G.d[k];
        cs += (size_t)q * mul;
        mul *= G.d[k];
    }
    return w.tD[(size_t)l + (size_t)G.Lt * cs] * w.t0[(size_t)l + (size_t)G.Lt * c];
}
// partial sums of sum_{l, c} f1(l, c) f2(l, c)
__global__ void k_c4_dot(double *__restrict__ part, const C4Field f1, const C4Field f2, const C4Geom G, const C4Weight w) {
    __shared__ double red[2 * 32];
    double v[2] = {0, 0};
    const size_t tot = (size_t)G.Lt * G.Nc;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += (size_t)gridDim.x * blockDim.x) {
        double2 p = cmul(c4_value(f1, G, (int)(idx / G.Nc), idx % G.Nc), c4_value(f2, G, (int)(idx / G.Nc), idx % G.Nc));
        p = cscale(c4_weight(w, G, (int)(idx / G.Nc), idx % G.Nc), p);
        v[0] += p.x;
        v[1] += p.y;
    }
    block_sum<2>(v, red);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = v[0]; part[2 * blockIdx.x + 1] = v[1]; }
}
__global__ void k_c4_finish(double2 *__restrict__ out, const double2 *__restrict__ s, int Lt, size_t Nc, double scale) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)(Lt + 1) * Nc) return;
    const size_t c = idx / (Lt + 1);
    const int t = (int)(idx % (Lt + 1));
    out[idx] = cscale(scale, s[(size_t)(t < Lt ? t : 0) * Nc + c]);          // S[beta] = S[0]
}

// sum over random vectors of mean_{tau, cells} GR(orb_g, cell + sh) conj(R)(orb_r, cell), divided by Nrv
static void c4_mean_GR_Rt(sq_greens *g, const C4Geom &G, int orb_g, const int *sh, int orb_r, double *re, double *im,
                          const C4Weight w = C4Weight{nullptr, nullptr, {0, 0, 0}}) {
    sq_fdm *f = g->f;
    const size_t V = (size_t)f->L * f->N;
    const int nb = 64;
    std::vector<double> h(2 * nb);
    double sr = 0, si = 0;
    for (i64 n = 0; n < g->Nrv; n++) {
        C4Field a = {g->GR.p + n * V, orb_g, 0, {sh[0], sh[1], sh[2]}}, b = {g->R.p + n * V, orb_r, 1, {0, 0, 0}};
        k_c4_dot<<<nb, 256, 0, f->stream>>>(g->part.p, a, b, G, w);
        SQ_LAUNCH_CHECK();
        f->launches++;
        SQ_CUDA(cudaMemcpyAsync(h.data(), g->part.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
        SQ_CUDA(cudaStreamSynchronize(f->stream));
        for (int k = 0; k < nb; k++) { sr += h[2 * k]; si += h[2 * k + 1]; }
    }
    const double den = (double)g->Nrv * (double)G.Lt * (double)G.Nc;
    *re = sr / den;
    *im = si / den;
}

// orbital-resolved density  n_a = 1 - dot(R_a, GR_a) / length   (scalar_measurements.jl:2-27)
void greens_measure_n_orbital_impl(sq_greens *g, int norb, int a, double *out) {
    sq_fdm *f = g->f;
    SQ_CUDA(cudaSetDevice(f->device));
    SQ_REQUIRE(norb >= 1 && a >= 1 && a <= norb && f->N % norb == 0, "bad orbital");
    C4Geom G = {(int)f->L, (int)f->N, norb, 1, {(int)(f->N / norb), 1, 1}, (size_t)(f->N / norb)};
    const int z[3] = {0, 0, 0};
    double re, im;
    c4_mean_GR_Rt(g, G, a - 1, z, a - 1, &re, &im);
    out[0] = 1.0 - re;
    out[1] = -im;
}

// measure_double_occ(greens_estimator, orbital)  (scalar_measurements.jl:98-109): the pair average restricted to the sites of one
// orbital, divided -- as the reference does -- by the TOTAL number of space-time points V
void greens_measure_double_occ_orbital_impl(sq_greens *g, int norb, int a, double *out) {
    sq_fdm *f = g->f;
    SQ_CUDA(cudaSetDevice(f->device));
    SQ_REQUIRE(norb >= 1 && a >= 1 && a <= norb && f->N % norb == 0, "bad orbital");
    const size_t V = (size_t)f->L * f->N;
    const int Nrv = (int)g->Nrv, nb = SQ_MAXPART / 4;
    out[0] = out[1] = 0.0;
    if (Nrv < 2) return;
    double *pd = g->part.p + (size_t)2 * 64 * Nrv * Nrv;
    k_double_occ<<<nb, 256, 0, f->stream>>>(g->R.p, g->GR.p, V, Nrv, pd, (int)f->N, norb, a - 1);
    SQ_LAUNCH_CHECK();
    f->launches++;
    std::vector<double> hd(2 * nb);
    SQ_CUDA(cudaMemcpyAsync(hd.data(), pd, hd.size() * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    double dr = 0, di = 0;
    for (int b = 0; b < nb; b++) { dr += hd[2 * b]; di += hd[2 * b + 1]; }
    const double npairs = 0.5 * Nrv * (Nrv - 1);
    out[0] = dr / ((double)V * npairs);
    out[1] = di / ((double)V * npairs);
}

// h_out: (Ltau + 1) x cells complex, tau fastest.  orb[4] 1-based orbitals (a, b, c, d); r: 4 x ndim static displacements.
// h_tD, h_t0: optional real hopping weights tΔ, t0 (Ltau x cells, tau fastest; the reference's PermutedDimsArray of fermion_path_integral.t,
// make_measurements.jl:316-320); real hoppings only, so the reference's conj_tΔ / conj_t0 switches have no effect.
void greens_measure_c4_impl(sq_greens *g, int kind, int norb, int ndim, const i64 *dims, const int *orb, const i64 *r, void *h_out,
                            const double *h_tD, const double *h_t0) {
    sq_fdm *f = g->f;
    SQ_CUDA(cudaSetDevice(f->device));
    SQ_REQUIRE(kind >= 0 && kind <= 2 && norb >= 1 && ndim >= 1 && ndim <= 3 && dims && orb && r && h_out, "bad argument");
    SQ_REQUIRE(g->Nrv >= 2, "the four-point contractions need at least two random vectors");
    C4Geom G;
    G.Lt = (int)f->L; G.N = (int)f->N; G.norb = norb; G.nd = ndim; G.Nc = 1;
    for (int k = 0; k < 3; k++) G.d[k] = 1;
    for (int k = 0; k < ndim; k++) { SQ_REQUIRE(dims[k] >= 1, "bad lattice dimension"); G.d[k] = (int)dims[k]; G.Nc *= (size_t)dims[k]; }
    SQ_REQUIRE((i64)(G.Nc * norb) == f->N, "unit cells x orbitals does not match the number of sites");
    for (int q = 0; q < 4; q++) SQ_REQUIRE(orb[q] >= 1 && orb[q] <= norb, "orbital index out of range");
    const int a = orb[0] - 1, b = orb[1] - 1, c = orb[2] - 1, d = orb[3] - 1;
    int R[4][3] = {{0}};
    for (int q = 0; q < 4; q++) for (int k = 0; k < ndim; k++) R[q][k] = (int)r[q * ndim + k];
    const int Lt = G.Lt;
    const size_t M = (size_t)Lt * G.Nc, V = (size_t)f->L * f->N;
    if (g->wa.n < 2 * M) { g->wa.alloc(2 * M, false); g->wb.alloc(2 * M, false); g->wc.alloc(2 * M, false); g->wt.alloc(2 * M, false); }
    std::vector<int> ax;
    ax.push_back(Lt);
    for (int k = ndim - 1; k >= 0; k--) ax.push_back(G.d[k]);
    SQ_CUDA(cudaMemsetAsync(g->wc.p, 0, M * sizeof(double2), f->stream));
    double2 *X = g->wa.p, *Y = g->wb.p, *C = g->wc.p, *T = g->wt.p;
    const double *tD = nullptr, *t0 = nullptr;
    if (h_tD || h_t0) {
        if (g->wreal.n < 2 * M) g->wreal.alloc(2 * M, false);
        if (h_tD) { SQ_CUDA(cudaMemcpyAsync(g->wreal.p, h_tD, M * sizeof(double), cudaMemcpyHostToDevice, f->stream)); tD = g->wreal.p; }
        if (h_t0) { SQ_CUDA(cudaMemcpyAsync(g->wreal.p + M, h_t0, M * sizeof(double), cudaMemcpyHostToDevice, f->stream)); t0 = g->wreal.p + M; }
    }
    const bool weighted = tD || t0;
    auto field = [&](bool gr, i64 n, int o, const int *sh) {
        C4Field x = {(gr ? g->GR.p : g->R.p) + n * V, o, gr ? 0 : 1, {sh[0], sh[1], sh[2]}};
        return x;
    };
    for (i64 n = 0; n + 1 < g->Nrv; n++)
        for (i64 m = n + 1; m < g->Nrv; m++) {
            const C4Field GRa = field(true, n, a, R[0]), Rtb = field(false, n, b, R[1]), GRc = field(true, m, c, R[2]), Rtd = field(false, m, d, R[3]);
            if (kind == 0) k_c4_fill<<<(unsigned)((M + 255) / 256), 256, 0, f->stream>>>(X, Y, GRa, GRc, Rtb, Rtd, G, tD, t0);
            else if (kind == 1) k_c4_fill<<<(unsigned)((M + 255) / 256), 256, 0, f->stream>>>(X, Y, GRa, Rtb, GRc, Rtd, G, tD, t0);
            else k_c4_fill<<<(unsigned)((M + 255) / 256), 256, 0, f->stream>>>(X, Y, Rtb, GRc, GRa, Rtd, G, tD, t0);
            SQ_LAUNCH_CHECK();
            greens_fftnd(g, &X, &T, ax, false);
            greens_fftnd(g, &Y, &T, ax, true);
            k_cmul_acc<<<f->num_sms * 4, 256, 0, f->stream>>>(C, X, Y, M);
            SQ_LAUNCH_CHECK();
            f->launches += 2;
        }
    greens_fftnd(g, &C, &T, ax, true);
    const double npairs = 0.5 * (double)g->Nrv * (double)(g->Nrv - 1);
    const size_t nout = (size_t)(Lt + 1) * G.Nc;
    k_c4_finish<<<(unsigned)((nout + 255) / 256), 256, 0, f->stream>>>(X, C, Lt, G.Nc, 1.0 / (std::sqrt((double)M) * npairs));
    SQ_LAUNCH_CHECK();
    f->launches++;
    std::vector<double2> out(nout);
    SQ_CUDA(cudaMemcpyAsync(out.data(), X, nout * sizeof(double2), cudaMemcpyDeviceToHost, f->stream));
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    // ---- delta-function terms
    auto cell_index = [&](const int *v) {                         // 0-based cell of the displacement v (mod L)
        size_t cidx = 0, mul = 1;
        for (int k = 0; k < ndim; k++) { int q = ((v[k] % G.d[k]) + G.d[k]) % G.d[k]; cidx += (size_t)q * mul; mul *= G.d[k]; }
        return cidx;
    };
    auto add = [&](int tau, const int *v, double re, double im) {
        double2 &o = out[(size_t)tau + (size_t)(Lt + 1) * cell_index(v)];
        o.x += re; o.y += im;
    };
    int sh[3] = {0, 0, 0}, at[3] = {0, 0, 0};
    double re, im;
    // with weights the delta terms carry tΔ (displaced) times t0 inside the mean (:318-326, :346-353, :560-568, :589-597)
    const bool need_delta = (kind == 0 && (a == b || c == d)) || (kind == 2 && (a == b || c == d));
    if (weighted && need_delta) SQ_REQUIRE(tD && t0, "the delta-function terms need both hopping weights (the reference shifts tΔ and multiplies by t0)");
    C4Weight W = {tD, t0, {0, 0, 0}};
    if (kind == 0) {
        if (a == b) {                                              // :305-337   -δ(a,b) δ(r, r2-r1) GR(i-r1+r2+r3-r4, c) R(i, d) at τ = β
            for (int k = 0; k < ndim; k++) { sh[k] = -(R[0][k] - R[1][k] - R[2][k] + R[3][k]); at[k] = -R[0][k] + R[1][k]; W.sh[k] = -(R[0][k] - R[1][k]); }
            c4_mean_GR_Rt(g, G, c, sh, d, &re, &im, W);
            add(Lt, at, -re, -im);
        }
        if (c == d) {                                              // :339-366
            for (int k = 0; k < ndim; k++) { sh[k] = -(-R[0][k] + R[1][k] + R[2][k] - R[3][k]); at[k] = -R[2][k] + R[3][k]; W.sh[k] = -(R[2][k] - R[3][k]); }
            c4_mean_GR_Rt(g, G, a, sh, b, &re, &im, W);
            add(Lt, at, -re, -im);
        }
        bool same = (a == b) && (c == d);
        for (int k = 0; k < ndim && same; k++) same = (((R[1][k] - R[0][k]) % G.d[k] + G.d[k]) % G.d[k]) == (((R[3][k] - R[2][k]) % G.d[k] + G.d[k]) % G.d[k]);
        if (same) {                                                // :368-385 (the weighted branch of the reference has a typo, `bonj`, and throws; its evident intent is computed)
            for (int k = 0; k < ndim; k++) at[k] = R[1][k] - R[0][k];
            double wmean = 1.0;
            if (weighted) {
                wmean = 0;
                for (size_t cc = 0; cc < G.Nc; cc++) {
                    size_t cs = 0, mul = 1, rem = cc;
                    for (int k = 0; k < ndim; k++) {
                        int ck = (int)(rem % G.d[k]); rem /= G.d[k];
                        int q = (((ck - (R[0][k] - R[1][k])) % G.d[k]) + G.d[k]) % G.d[k];
                        cs += (size_t)q * mul; mul *= G.d[k];
                    }
                    for (int l = 0; l < Lt; l++) wmean += h_tD[(size_t)l + (size_t)Lt * cs] * h_t0[(size_t)l + (size_t)Lt * cc];
                }
                wmean /= (double)M;
            }
            add(Lt, at, wmean, 0.0);
        }
    } else if (kind == 2) {
        if (a == b) {                                              // :546-573   at τ = 0, displacement r1 - r2
            for (int k = 0; k < ndim; k++) { sh[k] = -(-R[0][k] + R[1][k] - R[2][k] + R[3][k]); at[k] = R[0][k] - R[1][k]; W.sh[k] = R[0][k] - R[1][k]; }
            c4_mean_GR_Rt(g, G, c, sh, d, &re, &im, W);
            add(0, at, -re, -im);
        }
        if (c == d) {                                              // :575-602   at τ = β, displacement r4 - r3
            for (int k = 0; k < ndim; k++) { sh[k] = -(-R[0][k] + R[1][k] - R[2][k] + R[3][k]); at[k] = R[3][k] - R[2][k]; W.sh[k] = R[3][k] - R[2][k]; }
            c4_mean_GR_Rt(g, G, a, sh, b, &re, &im, W);
            add(Lt, at, -re, -im);
        }
    }
    memcpy(h_out, out.data(), nout * sizeof(double2));
}


// ---------------------------------------------------------------------------------------------------
// Local measurements: every estimator in tight_binding_measurements.jl / electron_phonon_measurements.jl is a weighted sum of
//   n(l, i)       = mean_rv (1 - GR[l,i,rv] Rt[l,i,rv])                         (measure_onsite_energy :43-63, measure_holstein_energy)
//   h(l; i -> f)  = mean_rv GR[l,i,rv] Rt[l,f,rv]                                (measure_bare_hopping_energy :66-98,
//                                                                                 measure_hopping_energy :101-133, measure_ssh_energy)
// over space-time.  The weights come from the host (they are model data: on-site energies, couplings times powers of x, hoppings).
// ---------------------------------------------------------------------------------------------------
// sum_{l,i} w[l][i] n(l, i); w in the device layout [l][i]
__global__ void k_weighted_density(double *__restrict__ part, const double2 *__restrict__ R, const double2 *__restrict__ GR, const double *__restrict__ w,
                                   size_t V, int Nrv) {
    __shared__ double red[2 * 32];
    double v[2] = {0, 0};
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < V; k += (size_t)gridDim.x * blockDim.x) {
        double sr = 0, si = 0;
        for (int n = 0; n < Nrv; n++) {
            const double2 r = R[(size_t)n * V + k], g = GR[(size_t)n * V + k];
            sr += 1.0 - (g.x * r.x + g.y * r.y);                 // 1 - g conj(r)
            si += -(g.y * r.x - g.x * r.y);
        }
        v[0] += w[k] * sr / Nrv;
        v[1] += w[k] * si / Nrv;
    }
    block_sum<2>(v, red);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = v[0]; part[2 * blockIdx.x + 1] = v[1]; }
}
// sum_{m,l} [ w[l][m] h(l; i_m -> f_m) + conj(w[l][m]) h(l; f_m -> i_m) ]; w complex in the layout [l][m]
__global__ void k_weighted_bonds(double *__restrict__ part, const double2 *__restrict__ R, const double2 *__restrict__ GR, const int2 *__restrict__ bonds,
                                 const double2 *__restrict__ w, int L, int N, int nb, int Nrv) {
    __shared__ double red[2 * 32];
    double v[2] = {0, 0};
    const size_t V = (size_t)L * N, tot = (size_t)L * nb;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < tot; k += (size_t)gridDim.x * blockDim.x) {
        const int l = (int)(k / nb), m = (int)(k % nb);
        const int2 b = bonds[m];
        double2 hf = make_double2(0, 0), hr = make_double2(0, 0);
        for (int n = 0; n < Nrv; n++) {
            const double2 gi = GR[(size_t)n * V + (size_t)l * N + b.x], gf = GR[(size_t)n * V + (size_t)l * N + b.y];
            double2 ri = R[(size_t)n * V + (size_t)l * N + b.x], rf = R[(size_t)n * V + (size_t)l * N + b.y];
            ri.y = -ri.y; rf.y = -rf.y;                             // Rt = conj(R)
            hf = cadd(hf, cmul(gi, rf));
            hr = cadd(hr, cmul(gf, ri));
        }
        const double2 wk = w[k];
        const double2 t = cadd(cmul(wk, hf), cmul(make_double2(wk.x, -wk.y), hr));
        v[0] += t.x / Nrv;
        v[1] += t.y / Nrv;
    }
    block_sum<2>(v, red);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = v[0]; part[2 * blockIdx.x + 1] = v[1]; }
}

static void greens_reduce_part(sq_greens *g, int nb, double *out) {
    sq_fdm *f = g->f;
    std::vector<double> h(2 * nb);
    SQ_CUDA(cudaMemcpyAsync(h.data(), g->part.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
    SQ_CUDA(cudaStreamSynchronize(f->stream));
    out[0] = out[1] = 0.0;
    for (int k = 0; k < nb; k++) { out[0] += h[2 * k]; out[1] += h[2 * k + 1]; }
}

// h_w: (N x Ltau) real weights, site fastest (the layout of fermion_path_integral.V).  out = sum_{i,l} w[i,l] n(l,i)
void greens_weighted_density_impl(sq_greens *g, const double *h_w, double *out) {
    sq_fdm *f = g->f;
    SQ_CUDA(cudaSetDevice(f->device));
    const size_t V = (size_t)f->L * f->N;
    if (g->wreal.n < V) g->wreal.alloc(V, false);
    g->wreal.upload(h_w, V, f->stream);
    const int nb = 128;
    k_weighted_density<<<nb, 256, 0, f->stream>>>(g->part.p, g->R.p, g->GR.p, g->wreal.p, V, (int)g->Nrv);
    SQ_LAUNCH_CHECK();
    f->launches++;
    greens_reduce_part(g, nb, out);
}
// bonds: nb pairs of 1-based sites (initial, final); h_w: (nb x Ltau) complex weights, bond fastest
void greens_weighted_bonds_impl(sq_greens *g, i64 nbonds, const i64 *bonds, const void *h_w, double *out) {
    sq_fdm *f = g->f;
    SQ_CUDA(cudaSetDevice(f->device));
    SQ_REQUIRE(nbonds >= 1 && bonds && h_w, "bad argument");
    std::vector<int2> hb(nbonds);
    for (i64 m = 0; m < nbonds; m++) {
        const i64 i = bonds[2 * m] - 1, j = bonds[2 * m + 1] - 1;
        SQ_REQUIRE(i >= 0 && i < f->N && j >= 0 && j < f->N, "bond site out of range");
        hb[m] = make_int2((int)i, (int)j);
    }
    const size_t nw = (size_t)f->L * nbonds;
    if (g->wcplx.n < nw) g->wcplx.alloc(nw, false);
    if (g->wbond.n < (size_t)nbonds) g->wbond.alloc(nbonds, false);
    g->wcplx.upload((const double2 *)h_w, nw, f->stream);
    g->wbond.upload(hb.data(), nbonds, f->stream);
    const int nb = 128;
    k_weighted_bonds<<<nb, 256, 0, f->stream>>>(g->part.p, g->R.p, g->GR.p, g->wbond.p, g->wcplx.p, (int)f->L, (int)f->N, (int)nbonds, (int)g->Nrv);
    SQ_LAUNCH_CHECK();
    f->launches++;
    greens_reduce_part(g, nb, out);
}

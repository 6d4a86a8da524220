// greens.cu -- GreensEstimator solves and the scalar measurements built on them.
//
// Replaces update_greens_estimator! (src/Measurements/GreensEstimator.jl:125-175) and measure_n,
// measure_Nsqrd, measure_double_occ (src/Measurements/scalar_measurements.jl:2-147).  R and G R stay on the
// device ([n][l][i]); the measurements are single-pass reductions over them.
#include "sq_internal.h"

#include <cmath>

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
__global__ void k_double_occ(const double2 *__restrict__ R, const double2 *__restrict__ GR, size_t V, int Nrv, double *__restrict__ part) {
    __shared__ double red[2 * 32];
    double v[2] = {0, 0};
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < V; k += (size_t)gridDim.x * blockDim.x) {
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
        rng_fill_normal((double *)g->R.p, 2 * V * g->Nrv, g->seed, g->counter++, f->stream);
        k_unit_modulus<<<f->num_sms * 4, 256, 0, f->stream>>>(g->R.p, V * g->Nrv);
        SQ_LAUNCH_CHECK();
        f->launches++;
    }
    if (kpm) kpm_update(kpm, nullptr, nullptr);                            // GreensEstimator.jl:150
    double avg = 0;
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

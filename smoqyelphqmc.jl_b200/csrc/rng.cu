// rng.cu -- K9: counter-based Philox4x32-10 fills (uniform, normal) on the device.
//
// Stands in for randn!(rng, Phi) (src/PFFCalculator.jl:67), randn!(rng, R) (src/Measurements/
// GreensEstimator.jl:141), the Lanczos start vector (src/KPMPreconditioner.jl:634) and SmoQyDQMC's
// momentum refresh.  Bit parity with Julia's Xoshiro stream is not a goal: parity tests
// inject host-supplied randoms through the ABI; production runs use this generator (statistical parity).
#include "sq_internal.h"

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
}

__device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t stream, uint64_t idx, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; r++) philox_round(c, k);
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {
    // 53 random bits -> (0, 1)
    uint64_t x = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);
    x &= ((1ULL << 53) - 1);
    return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

__global__ void k_fill(double *__restrict__ out, size_t n, uint64_t seed, uint64_t stream, int normal) {
    size_t pair = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i0 = 2 * pair;
    if (i0 >= n) return;
    uint32_t r[4];
    philox4x32_10(seed, stream, pair, r);
    double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
    double a, b;
    if (normal) {
        double rad = sqrt(-2.0 * log(u1));
        double s, c;
        sincospi(2.0 * u2, &s, &c);
        a = rad * c;
        b = rad * s;
    } else { a = u1; b = u2; }
    out[i0] = a;
    if (i0 + 1 < n) out[i0 + 1] = b;
}

void rng_fill_normal(double *d_out, size_t n, uint64_t seed, uint64_t stream, cudaStream_t s) {
    if (!n) return;
    size_t pairs = (n + 1) / 2;
    k_fill<<<(unsigned)((pairs + 255) / 256), 256, 0, s>>>(d_out, n, seed, stream, 1);
    SQ_LAUNCH_CHECK();
}
void rng_fill_uniform(double *d_out, size_t n, uint64_t seed, uint64_t stream, cudaStream_t s) {
    if (!n) return;
    size_t pairs = (n + 1) / 2;
    k_fill<<<(unsigned)((pairs + 255) / 256), 256, 0, s>>>(d_out, n, seed, stream, 0);
    SQ_LAUNCH_CHECK();
}

// speedy-b200: SPPT -- stochastically perturbed parametrisation tendencies (sppt.f90:40-146, hook physics.f90:233-248).
//
// A compile-time switch in the reference (params.f90:44, off) whose routine cannot run as written (shape mismatch of
// `sigma`, the AR(1) state freed at every return, result deallocated before it is returned, clock-seeded generator); what
// is built here is the algorithm those lines describe, restated operation for operation in oracle/physics.cpp: gen_sppt --
// complex Gaussian noise clipped to +-10, sigma(m,n) = f0 exp(-L^2 el2 / 4), AR(1) in time with phi = exp(-(24/nsteps)/6 h),
// inverse transform, clipping to +-1, tendencies (1 + pattern * mu(k)) * (X - X_dyn) + X_dyn.  Per member: its own
// pattern (extra state row off_sppt) and a COUNTER-BASED generator (splitmix64 keyed by seed, arena slot and call count)
// instead of random_number, so a run is reproducible and comparable with the oracle.  Off by default (spdy_set_sppt);
// when off, not one instruction or byte of the model step changes: the physics kernel itself is untouched, SPPT runs as
// three small kernels around it.
#include "kernels.h"

namespace spdy {

__device__ __forceinline__ unsigned long long sppt_mix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// randn(0, 1) of sppt.f90:118-133: Box-Muller with the reference's REAL(4) constant 2.0 * 6.28318530718
__device__ __forceinline__ double sppt_randn(unsigned long long key, unsigned long long ctr) {
    const double r1 = ((double)(sppt_mix64(key + 2 * ctr) >> 11) + 1.0) * 0x1p-53;
    const double r2 = (double)(sppt_mix64(key + 2 * ctr + 1) >> 11) * 0x1p-53;
    const double u = sqrt(-2.0 * log(r1));
    const double v = (double)(2.0f * 6.28318530718f) * r2;
    return 0.0 + 1.0 * u * sin(v);
}
// AR(1) step of the spectral pattern: warp per (coefficient, level), lane = member
__global__ void __launch_bounds__(128) k_sppt_update(const Ctx c, const unsigned long long seed) {
    const int lane = threadIdx.x & 31, w = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    const int k = w % KX, q = w / KX;
    if (q >= NSPC) return;
    const unsigned long long member = (unsigned long long)c.tiles[t] * TILE + lane;
    const unsigned long long calls = (unsigned long long)slot(c, t, lane, SL_SPPT_CALLS);
    const unsigned long long key = sppt_mix64(seed ^ sppt_mix64(member));
    const unsigned long long ctr = (calls * (unsigned long long)NSPC + (unsigned long long)q) * KX + k;
    const double rr = sppt_randn(key, 2 * ctr), ri = sppt_randn(key, 2 * ctr + 1);
    const double er = fmin(10.0, fabs(rr)) * copysign(1.0, rr), ei = fmin(10.0, fabs(ri)) * copysign(1.0, ri);
    const double sigma = c.G->sppt_sigma[q];
    double *x = stp(c, t, c.off_sppt + (long long)k * NSP, lane) + (size_t)(2 * (q % MX) + M2 * (q / MX)) * TILE;
    double xr, xi;
    if (calls == 0) {
        const double a = c.G->sppt_first_fac * sigma;
        xr = a * er, xi = a * ei;
    } else {
        const double phi = c.G->sppt_phi;
        xr = phi * x[0] + sigma * er, xi = phi * x[TILE] + sigma * ei;
    }
    if (lane_active(c, t, lane)) x[0] = xr, x[TILE] = xi;
}
// copies of the dynamical tendencies the physics adds to (physics.f90:80-83: utend_dyn ...; the physics changes the wind
// tendencies at the lowest level only)
__global__ void __launch_bounds__(128) k_sppt_save(const Ctx c, const ScratchLayout L) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    const size_t e = (size_t)q * TILE, lev = (size_t)NG * TILE;
#pragma unroll
    for (int k = 0; k < KX; k++) {
        *(scp(c, t, L.tdyn, lane) + e + k * lev) = *(scp(c, t, L.ttend, lane) + e + k * lev);
        *(scp(c, t, L.qdyn, lane) + e + k * lev) = *(scp(c, t, L.trtend, lane) + e + k * lev);
    }
    *(scp(c, t, L.udyn8, lane) + e) = *(scp(c, t, L.utend, lane) + e + 7 * lev);
    *(scp(c, t, L.vdyn8, lane) + e) = *(scp(c, t, L.vtend, lane) + e + 7 * lev);
}
// physics.f90:233-248 with the pattern clipped to +-1 (sppt.f90:109), mu(k) = 1 (sppt.f90:20); the thread of grid point 0
// counts the call
__global__ void __launch_bounds__(128) k_sppt_apply(const Ctx c, const ScratchLayout L) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    const size_t e = (size_t)q * TILE, lev = (size_t)NG * TILE;
    const double mu = 1.0;
    double *pat = scp(c, t, L.spptg, lane) + e;
#pragma unroll
    for (int k = 0; k < KX; k++) {
        const double p0 = pat[k * lev], p = fmin(1.0, fabs(p0)) * copysign(1.0, p0);
        pat[k * lev] = p;  // the clipped pattern stays readable (spdy_debug_get_sppt)
        double *tt = scp(c, t, L.ttend, lane) + e + k * lev, *qt = scp(c, t, L.trtend, lane) + e + k * lev;
        const double td = *(scp(c, t, L.tdyn, lane) + e + k * lev), qd = *(scp(c, t, L.qdyn, lane) + e + k * lev);
        *tt = (1 + p * mu) * (*tt - td) + td;
        *qt = (1 + p * mu) * (*qt - qd) + qd;
        if (k == KX - 1) {
            double *ut = scp(c, t, L.utend, lane) + e + k * lev, *vt = scp(c, t, L.vtend, lane) + e + k * lev;
            const double ud = *(scp(c, t, L.udyn8, lane) + e), vd = *(scp(c, t, L.vdyn8, lane) + e);
            *ut = (1 + p * mu) * (*ut - ud) + ud;
            *vt = (1 + p * mu) * (*vt - vd) + vd;
        }
    }
    if (q == 0 && lane_active(c, t, lane)) slot(c, t, lane, SL_SPPT_CALLS) = slot(c, t, lane, SL_SPPT_CALLS) + 1.0;
}
void launch_sppt_update(cudaStream_t s, const Ctx &c, unsigned long long seed) {
    k_sppt_update<<<dim3(NSPC * KX / 4, c.ntiles), 128, 0, s>>>(c, seed);
}
void launch_sppt_save(cudaStream_t s, const Ctx &c, const ScratchLayout &L) { k_sppt_save<<<dim3(NG / 4, c.ntiles), 128, 0, s>>>(c, L); }
void launch_sppt_apply(cudaStream_t s, const Ctx &c, const ScratchLayout &L) { k_sppt_apply<<<dim3(NG / 4, c.ntiles), 128, 0, s>>>(c, L); }

}  // namespace spdy

// speedy-b200: kernel launch interface shared by the translation units of the unity build (spdy_all.cu)
#pragma once
#include "spdy.cuh"

namespace spdy {

// inverse transform descriptor: spectral field -> grid field in scratch (spectral.f90:251-261)
struct InvDesc {
    FieldRef src;
    long long dst;  // scratch element offset of the (96,48) grid field
    int kcos;       // 1: plain, 2: multiply by 1/cos(lat) (fourier.f90:88-92)
    int pad;
};
// forward transform: loader modes (grid-point products fused into the FFT loads)
enum FwdMode { FM_PLAIN = 0, FM_COS = 1, FM_KE = 2, FM_FLUXT = 3, FM_FLUX = 4, FM_NMODES = 5, FM_ALL = 5 };
struct FwdDesc {
    FieldRef a, b;  // grid fields (96,48)
    double k0;      // FM_FLUXT: reference temperature subtracted from b
    int kcos;       // 2: cosgr, 3: cosgr2 (spectral.f90:229-242)
    int fidx;       // Fourier slot / index into the FwdOut list
    int mode;       // loader mode of this field in a mixed list (FM_ALL launches of the fused forward kernel)
    int pad;
};
struct FwdOut {
    FieldRef dst;   // spectral field (62,32)
};

void upload_const_tables(const ConstTables &C);
void launch_legendre_inv(cudaStream_t s, const Ctx &c, const InvDesc *d, int nf, long long four_off);
void launch_fft_inv(cudaStream_t s, const Ctx &c, const InvDesc *d, int nf, long long four_off);
void launch_fft_fwd(cudaStream_t s, const Ctx &c, int mode, const FwdDesc *d, int nf, long long four_off);
void launch_legendre_dir(cudaStream_t s, const Ctx &c, const FwdOut *o, int nf, long long four_off);
void launch_uvspec(cudaStream_t s, const Ctx &c, FieldRef vor, FieldRef dv, FieldRef u, FieldRef v, int nlev, int tri);
void launch_gradient(cudaStream_t s, const Ctx &c, FieldRef psi, FieldRef dx, FieldRef dy, int tri);
void launch_geopotential(cudaStream_t s, const Ctx &c, FieldRef tlev, FieldRef phis, FieldRef phi);

}  // namespace spdy

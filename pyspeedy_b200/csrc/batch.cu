// speedy-b200: stage-level entry points (parity tests, microbenchmarks) and ensemble diagnostics.
// Fields are treated as pseudo-members: field f lives in lane f % 32 of workspace tile f / 32, so the very same
// kernels of the model step are exercised.
namespace spdy {

struct Workspace {
    double *buf = nullptr;
    int *tiles = nullptr;
    unsigned *masks = nullptr;
    int ntiles = 0;
    InvDesc *d_inv = nullptr;
    FwdDesc *d_fwd = nullptr;
    FwdOut *d_out = nullptr;
    InvDesc *d_inv2 = nullptr;
    FwdDesc *d_fwd2 = nullptr;
    FwdOut *d_out2 = nullptr;
    double *d_lin = nullptr;
    size_t lin_cap = 0;
};
static Workspace W;
// per pseudo-member: four spectral fields, one Fourier field, two grid fields
constexpr long long WS_SPEC = 0, WS_SPEC2 = NSP, WS_SPEC3 = 2 * NSP, WS_SPEC4 = 3 * NSP, WS_FOUR = 4 * NSP,
                    WS_GRID = 4 * NSP + NFOUR, WS_GRID2 = WS_GRID + NG, WS_ELEMS = WS_GRID2 + NG;

static Ctx ws_ctx(int nfields) {
    engine_init();
    const int nt = (nfields + TILE - 1) / TILE;
    if (nt > W.ntiles) {
        if (W.buf) CK(cudaFree(W.buf)), CK(cudaFree(W.tiles)), CK(cudaFree(W.masks));
        CK(cudaMalloc(&W.buf, (size_t)nt * WS_ELEMS * TILE * sizeof(double)));
        CK(cudaMemset(W.buf, 0, (size_t)nt * WS_ELEMS * TILE * sizeof(double)));
        std::vector<int> t(nt);
        std::vector<unsigned> m(nt, 0xffffffffu);
        for (int i = 0; i < nt; i++) t[i] = i;
        CK(cudaMalloc(&W.tiles, nt * sizeof(int)));
        CK(cudaMalloc(&W.masks, nt * sizeof(unsigned)));
        CK(cudaMemcpy(W.tiles, t.data(), nt * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(W.masks, m.data(), nt * sizeof(unsigned), cudaMemcpyHostToDevice));
        W.ntiles = nt;
    }
    if (!W.d_inv) {
        CK(cudaMalloc(&W.d_inv, sizeof(InvDesc)));
        CK(cudaMalloc(&W.d_fwd, sizeof(FwdDesc)));
        CK(cudaMalloc(&W.d_out, sizeof(FwdOut)));
        const FwdDesc f{REF_SCR | WS_GRID, 0, 0.0, 2, 0};
        const FwdOut o{REF_SCR | WS_SPEC};
        CK(cudaMemcpy(W.d_fwd, &f, sizeof(f), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(W.d_out, &o, sizeof(o), cudaMemcpyHostToDevice));
        // two-field lists of the wind chain: (u cos, v cos) in WS_SPEC3/4 <-> (u, v) in WS_GRID/2
        const InvDesc i2[2] = {InvDesc{REF_SCR | WS_SPEC3, WS_GRID, 2, 0}, InvDesc{REF_SCR | WS_SPEC4, WS_GRID2, 2, 0}};
        const FwdOut o2[2] = {FwdOut{REF_SCR | WS_SPEC3}, FwdOut{REF_SCR | WS_SPEC4}};
        CK(cudaMalloc(&W.d_inv2, sizeof(i2)));
        CK(cudaMalloc(&W.d_fwd2, 2 * sizeof(FwdDesc)));
        CK(cudaMalloc(&W.d_out2, sizeof(o2)));
        CK(cudaMemcpy(W.d_inv2, i2, sizeof(i2), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(W.d_out2, o2, sizeof(o2), cudaMemcpyHostToDevice));
    }
    Ctx c = make_ctx(W.tiles, W.masks, nt);
    c.scr = W.buf, c.scr_elems = WS_ELEMS, c.st = nullptr;
    return c;
}
static void ws_set_kcos(int kcos) {
    const InvDesc d{REF_SCR | WS_SPEC, WS_GRID, kcos, 0};
    CK(cudaMemcpy(W.d_inv, &d, sizeof(d), cudaMemcpyHostToDevice));
}
__global__ void k_ws_pack(const double *lin, double *buf, long long off, long long len, long long n, int unpack) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n * len) return;
    const long long f = i / len, e = i - f * len;
    double *p = buf + ((f / TILE) * WS_ELEMS + off + e) * TILE + (f % TILE);
    if (unpack) const_cast<double *>(lin)[i] = *p; else *p = lin[i];
}
static void ws_move(const double *host_in, double *host_out, long long off, long long len, int n) {
    const size_t bytes = (size_t)n * len * sizeof(double);
    if (bytes > W.lin_cap) {
        if (W.d_lin) CK(cudaFree(W.d_lin));
        CK(cudaMalloc(&W.d_lin, bytes));
        W.lin_cap = bytes;
    }
    const long long tot = (long long)n * len;
    if (host_in) {
        CK(cudaMemcpy(W.d_lin, host_in, bytes, cudaMemcpyHostToDevice));
        k_ws_pack<<<(int)((tot + 255) / 256), 256, 0, E.stream>>>(W.d_lin, W.buf, off, len, n, 0);
        CK(cudaStreamSynchronize(E.stream));
    } else {
        k_ws_pack<<<(int)((tot + 255) / 256), 256, 0, E.stream>>>(W.d_lin, W.buf, off, len, n, 1);
        CK(cudaStreamSynchronize(E.stream));
        CK(cudaMemcpy(host_out, W.d_lin, bytes, cudaMemcpyDeviceToHost));
    }
}

}  // namespace spdy

extern "C" {

int spdy_batch_legendre_inv(const double *spec, double *four, int n) {
    API_LOCK;
    Ctx c = ws_ctx(n);
    ws_set_kcos(1);
    ws_move(spec, nullptr, WS_SPEC, NSP, n);
    launch_legendre_inv(E.stream, c, W.d_inv, 1, WS_FOUR);
    COUNT(1);
    ws_move(nullptr, four, WS_FOUR, NFOUR, n);
    return 0;
}
int spdy_batch_fourier_inv(const double *four, double *grid, int kcos, int n) {
    API_LOCK;
    Ctx c = ws_ctx(n);
    ws_set_kcos(kcos);
    ws_move(four, nullptr, WS_FOUR, NFOUR, n);
    launch_fft_inv(E.stream, c, W.d_inv, 1, WS_FOUR);
    COUNT(1);
    ws_move(nullptr, grid, WS_GRID, NG, n);
    return 0;
}
int spdy_batch_fourier_dir(const double *grid, double *four, int n) {
    API_LOCK;
    Ctx c = ws_ctx(n);
    ws_move(grid, nullptr, WS_GRID, NG, n);
    launch_fft_fwd(E.stream, c, FM_PLAIN, W.d_fwd, 1, WS_FOUR);
    COUNT(1);
    ws_move(nullptr, four, WS_FOUR, NFOUR, n);
    return 0;
}
int spdy_batch_legendre_dir(const double *four, double *spec, int n) {
    API_LOCK;
    Ctx c = ws_ctx(n);
    ws_move(four, nullptr, WS_FOUR, NFOUR, n);
    launch_legendre_dir(E.stream, c, W.d_out, 1, WS_FOUR);
    COUNT(1);
    ws_move(nullptr, spec, WS_SPEC, NSP, n);
    return 0;
}
int spdy_batch_spec2grid(const double *spec, double *grid, int kcos, int n) {
    API_LOCK;
    Ctx c = ws_ctx(n);
    ws_set_kcos(kcos);
    ws_move(spec, nullptr, WS_SPEC, NSP, n);
    if (fused_transforms()) {
        launch_spec2grid_mma4(E.stream, c, W.d_inv, 1);
    } else {
        launch_legendre_inv(E.stream, c, W.d_inv, 1, WS_FOUR);
        launch_fft_inv(E.stream, c, W.d_inv, 1, WS_FOUR);
    }
    COUNT(2);
    ws_move(nullptr, grid, WS_GRID, NG, n);
    return 0;
}
int spdy_batch_grid2spec(const double *grid, double *spec, int n) {
    API_LOCK;
    Ctx c = ws_ctx(n);
    ws_move(grid, nullptr, WS_GRID, NG, n);
    if (fused_transforms()) {  // the model step's default forward kernel (the workspace is a scratch arena)
        launch_grid2spec_mma2(E.stream, c, FM_PLAIN, W.d_fwd, W.d_out, 1, 0);
    } else {
        launch_fft_fwd(E.stream, c, FM_PLAIN, W.d_fwd, 1, WS_FOUR);
        launch_legendre_dir(E.stream, c, W.d_out, 1, WS_FOUR);
    }
    COUNT(2);
    ws_move(nullptr, spec, WS_SPEC, NSP, n);
    return 0;
}

// ---- spectral operators on batches of fields (operator-level parity tests): the kernels of the model step --------------
static void ws_set_fwd2(int kcos) {  // forward list of the wind chain: u, v times cosgr (kcos 2) or cosgr2 (3), spectral.f90:229-242
    const FwdDesc f2[2] = {FwdDesc{REF_SCR | WS_GRID, 0, 0.0, kcos, 0, FM_COS, 0}, FwdDesc{REF_SCR | WS_GRID2, 0, 0.0, kcos, 1, FM_COS, 0}};
    CK(cudaMemcpy(W.d_fwd2, f2, sizeof(f2), cudaMemcpyHostToDevice));
}
static void ws_forward2(const Ctx &c) {  // (WS_GRID, WS_GRID2) -> (WS_SPEC3, WS_SPEC4) with the cos-latitude loader
    if (fused_transforms()) {
        launch_grid2spec_mma2(E.stream, c, FM_COS, W.d_fwd2, W.d_out2, 2, 0);
        COUNT(1);
    } else {
        launch_fft_fwd(E.stream, c, FM_COS, W.d_fwd2, 2, WS_FOUR);  // Fourier slots 0, 1: the workspace has ONE Fourier field
        launch_legendre_dir(E.stream, c, W.d_out2, 2, WS_FOUR);
        COUNT(2);
    }
}
// vort2vel = uvspec (spectral.f90:190-214): (vor, div) -> (u cos, v cos)
int spdy_batch_vort2vel(const double *vor, const double *dv, double *u, double *v, int n) {
    API_LOCK;
    Ctx c = ws_ctx(n);
    ws_move(vor, nullptr, WS_SPEC, NSP, n), ws_move(dv, nullptr, WS_SPEC2, NSP, n);
    launch_uvspec(E.stream, c, REF_SCR | WS_SPEC, REF_SCR | WS_SPEC2, REF_SCR | WS_SPEC3, REF_SCR | WS_SPEC4, 1, 0);
    COUNT(1);
    ws_move(nullptr, u, WS_SPEC3, NSP, n), ws_move(nullptr, v, WS_SPEC4, NSP, n);
    return 0;
}
// vel2vort = vdspec (spectral.f90:160-186): spectral (u, v) -> (vor, div)
int spdy_batch_vel2vort(const double *u, const double *v, double *vor, double *dv, int n) {
    API_LOCK;
    Ctx c = ws_ctx(n);
    ws_move(u, nullptr, WS_SPEC3, NSP, n), ws_move(v, nullptr, WS_SPEC4, NSP, n);
    launch_vdspec(E.stream, c, REF_SCR | WS_SPEC3, REF_SCR | WS_SPEC4, REF_SCR | WS_SPEC, REF_SCR | WS_SPEC2);
    COUNT(1);
    ws_move(nullptr, vor, WS_SPEC, NSP, n), ws_move(nullptr, dv, WS_SPEC2, NSP, n);
    return 0;
}
// gradient (spectral.f90:275-296)
int spdy_batch_gradient(const double *psi, double *dx, double *dy, int n) {
    API_LOCK;
    Ctx c = ws_ctx(n);
    ws_move(psi, nullptr, WS_SPEC, NSP, n);
    launch_gradient(E.stream, c, REF_SCR | WS_SPEC, REF_SCR | WS_SPEC3, REF_SCR | WS_SPEC4, 0);
    COUNT(1);
    ws_move(nullptr, dx, WS_SPEC3, NSP, n), ws_move(nullptr, dy, WS_SPEC4, NSP, n);
    return 0;
}
// laplacian / laplacian_inv (spectral.f90:140-155)
int spdy_batch_laplacian(const double *in, double *out, int inverse, int n) {
    API_LOCK;
    Ctx c = ws_ctx(n);
    ws_move(in, nullptr, WS_SPEC, NSP, n);
    launch_laplacian(E.stream, c, REF_SCR | WS_SPEC, REF_SCR | WS_SPEC2, inverse);
    COUNT(1);
    ws_move(nullptr, out, WS_SPEC2, NSP, n);
    return 0;
}
// grid_vel2vort (spectral.f90:218-248): grid (u, v) x cosgr (kcos = 2) or cosgr2 -> two forward transforms -> vel2vort
int spdy_batch_grid_vel2vort(const double *ug, const double *vg, double *vor, double *dv, int kcos, int n) {
    API_LOCK;
    Ctx c = ws_ctx(n);
    ws_set_fwd2(kcos == 2 ? 2 : 3);
    ws_move(ug, nullptr, WS_GRID, NG, n), ws_move(vg, nullptr, WS_GRID2, NG, n);
    if (fused_transforms()) {
        ws_forward2(c);
    } else {  // one Fourier field per pseudo-member in the workspace: the two transforms one after the other
        FwdDesc f = FwdDesc{REF_SCR | WS_GRID, 0, 0.0, kcos == 2 ? 2 : 3, 0, FM_COS, 0};
        FwdOut o = FwdOut{REF_SCR | WS_SPEC3};
        for (int q = 0; q < 2; q++) {
            CK(cudaMemcpy(W.d_fwd2, &f, sizeof(f), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(W.d_out2, &o, sizeof(o), cudaMemcpyHostToDevice));
            launch_fft_fwd(E.stream, c, FM_COS, W.d_fwd2, 1, WS_FOUR);
            launch_legendre_dir(E.stream, c, W.d_out2, 1, WS_FOUR);
            COUNT(2);
            CK(cudaStreamSynchronize(E.stream));
            f.a = REF_SCR | WS_GRID2, o.dst = REF_SCR | WS_SPEC4;
        }
        const FwdOut o2[2] = {FwdOut{REF_SCR | WS_SPEC3}, FwdOut{REF_SCR | WS_SPEC4}};
        CK(cudaMemcpy(W.d_out2, o2, sizeof(o2), cudaMemcpyHostToDevice));
    }
    launch_vdspec(E.stream, c, REF_SCR | WS_SPEC3, REF_SCR | WS_SPEC4, REF_SCR | WS_SPEC, REF_SCR | WS_SPEC2);
    COUNT(1);
    ws_move(nullptr, vor, WS_SPEC, NSP, n), ws_move(nullptr, dv, WS_SPEC2, NSP, n);
    return 0;
}

// BASELINE config 4: `npairs` synthetic (vor, div) pairs resident in HBM; one rep =
//   vort2vel -> spec2grid (kcos = 2) of the 2*npairs wind fields -> grid2spec with the cos-latitude loader -> vel2vort ->
//   gradient(vor): the grid<->spectral round trip with the grad / uvspec / vdspec chain, through the kernels of the model
//   step (fused transforms).  The output (vor, div) of a rep is the input of the next (the chain is a projection up to the
//   reference's 4e-3 quadrature error, so the fields stay bounded).  ms[0] = mean device time per rep, ms[1..5] = per stage.
int spdy_bench_spectral_chain(const double *vor, const double *dv, int npairs, int reps, float *ms) {
    API_LOCK;
    if (!fused_transforms()) {
        fprintf(stderr, "speedy-b200: spdy_bench_spectral_chain times the fused transforms (unset SPDY_FUSED)\n");
        return -1;
    }
    Ctx c = ws_ctx(npairs);
    ws_set_fwd2(2);
    ws_move(vor, nullptr, WS_SPEC, NSP, npairs), ws_move(dv, nullptr, WS_SPEC2, NSP, npairs);
    cudaEvent_t ev[6];
    for (int i = 0; i < 6; i++) CK(cudaEventCreate(&ev[i]));
    float acc[6] = {0, 0, 0, 0, 0, 0};
    for (int r = -3; r < reps; r++) {  // three warm-up reps
        CK(cudaEventRecord(ev[0], E.stream));
        launch_uvspec(E.stream, c, REF_SCR | WS_SPEC, REF_SCR | WS_SPEC2, REF_SCR | WS_SPEC3, REF_SCR | WS_SPEC4, 1, 0);
        CK(cudaEventRecord(ev[1], E.stream));
        launch_spec2grid_mma4(E.stream, c, W.d_inv2, 2);
        CK(cudaEventRecord(ev[2], E.stream));
        ws_forward2(c);
        CK(cudaEventRecord(ev[3], E.stream));
        launch_vdspec(E.stream, c, REF_SCR | WS_SPEC3, REF_SCR | WS_SPEC4, REF_SCR | WS_SPEC, REF_SCR | WS_SPEC2);
        CK(cudaEventRecord(ev[4], E.stream));
        launch_gradient(E.stream, c, REF_SCR | WS_SPEC, REF_SCR | WS_SPEC3, REF_SCR | WS_SPEC4, 0);
        CK(cudaEventRecord(ev[5], E.stream));
        COUNT(5);
        CK(cudaStreamSynchronize(E.stream));
        if (r >= 0) {
            float t;
            for (int i = 0; i < 5; i++) CK(cudaEventElapsedTime(&t, ev[i], ev[i + 1])), acc[i + 1] += t;
            CK(cudaEventElapsedTime(&t, ev[0], ev[5]));
            acc[0] += t;
        }
    }
    for (int i = 0; i < 6; i++) CK(cudaEventDestroy(ev[i])), ms[i] = reps > 0 ? acc[i] / reps : 0.f;
    return 0;
}

int spdy_bench_roundtrip(const double *spec, double *spec_out, int n, int reps, float *ms_per_rep, float *per_kernel_ms) {
    API_LOCK;
    Ctx c = ws_ctx(n);
    ws_set_kcos(1);
    ws_move(spec, nullptr, WS_SPEC, NSP, n);
    cudaEvent_t ev[5];
    for (int i = 0; i < 5; i++) CK(cudaEventCreate(&ev[i]));
    float acc[4] = {0, 0, 0, 0}, tot = 0.f;
    for (int r = -2; r < reps; r++) {  // two warm-up reps
        CK(cudaEventRecord(ev[0], E.stream));
        launch_legendre_inv(E.stream, c, W.d_inv, 1, WS_FOUR);
        CK(cudaEventRecord(ev[1], E.stream));
        launch_fft_inv(E.stream, c, W.d_inv, 1, WS_FOUR);
        CK(cudaEventRecord(ev[2], E.stream));
        launch_fft_fwd(E.stream, c, FM_PLAIN, W.d_fwd, 1, WS_FOUR);
        CK(cudaEventRecord(ev[3], E.stream));
        launch_legendre_dir(E.stream, c, W.d_out, 1, WS_FOUR);
        CK(cudaEventRecord(ev[4], E.stream));
        COUNT(4);
        CK(cudaStreamSynchronize(E.stream));
        if (r >= 0) {
            float ms;
            for (int i = 0; i < 4; i++) CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1])), acc[i] += ms;
            CK(cudaEventElapsedTime(&ms, ev[0], ev[4]));
            tot += ms;
        }
    }
    for (int i = 0; i < 5; i++) CK(cudaEventDestroy(ev[i]));
    if (reps > 0) {
        *ms_per_rep = tot / reps;
        for (int i = 0; i < 4; i++) per_kernel_ms[i] = acc[i] / reps;
    }
    if (spec_out) ws_move(nullptr, spec_out, WS_SPEC, NSP, n);
    return 0;
}

}  // extern "C"

// ---- ensemble set-up / output extensions ----------------------------------------------------------------------
namespace spdy {
__global__ void __launch_bounds__(256) k_clone_lane(double *arena, long long tile_elems, int st, int sl, int dt, int dl, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) arena[((long long)dt * tile_elems + i) * TILE + dl] = arena[((long long)st * tile_elems + i) * TILE + sl];
}
// counter-based generator (splitmix64) + Box-Muller: N(0, sigma) per (member, grid point, level)
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(128) k_noise_grid(const Ctx c, long long dst, unsigned long long seed, double sigma, int nlev) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    const unsigned long long member = (unsigned long long)c.tiles[t] * TILE + lane;
    for (int k = 0; k < nlev; k++) {
        const unsigned long long ctr = (member * KX + k) * NG + q;
        const unsigned long long a = mix64(seed ^ (ctr * 2)), b = mix64(seed ^ (ctr * 2 + 1));
        const double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0), u2 = (b >> 11) * (1.0 / 9007199254740992.0);
        *(scp(c, t, dst + (long long)k * NG, lane) + (size_t)q * TILE) = sigma * sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
}
__global__ void __launch_bounds__(256) k_add_spec(const Ctx c, FieldRef src, long long dst_off, int n) {
    const int lane = threadIdx.x & 31, t = blockIdx.y;
    if (!lane_active(c, t, lane)) return;
    const double *s = refp(c, t, src, lane);
    double *d = stp(c, t, dst_off, lane);
    for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) d[(size_t)i * TILE] += s[(size_t)i * TILE];
}
}  // namespace spdy

extern "C" {
// copy the complete device state of `src` into every member of `dst` (ensemble set-up from one initialised member)
int spdy_clone_state(int64_t src, const int64_t *dst, int n) {
    API_LOCK;
    Member *ms = member_of(src);
    if (!ms) return -1;
    for (int i = 0; i < n; i++) {
        Member *md = member_of(dst[i]);
        if (!md || md == ms) continue;
        k_clone_lane<<<(int)((E.st_elems + 255) / 256), 256, 0, E.stream>>>(E.st, E.st_elems, ms->tile, ms->lane, md->tile, md->lane, E.st_elems);
        if (ms->n_months >= 0) {
            ensure_sst_months(ms->n_months + 2);
            k_clone_lane<<<(int)((E.sst_elems + 255) / 256), 256, 0, E.stream>>>(E.sst, E.sst_elems, ms->tile, ms->lane, md->tile, md->lane, E.sst_elems);
        }
        COUNT(2);
        md->initialized = ms->initialized, md->n_months = ms->n_months, md->current_step = ms->current_step;
        md->bound_ctl = -1;
        memcpy(md->lon, ms->lon, sizeof(ms->lon)), memcpy(md->lat, ms->lat, sizeof(ms->lat)), memcpy(md->lev, ms->lev, sizeof(ms->lev));
    }
    CK(cudaStreamSynchronize(E.stream));
    return 0;
}
// t_grid += N(0, sigma) i.i.d. per grid point, then t = grid2spec(t_grid) on time level 1 -- the perturbed-IC ensemble of
// examples/Ensemble_forecast.ipynb cell 8, done for all listed members on the device (linear: t += grid2spec(noise))
int spdy_perturb_temperature(const int64_t *hs, int n, unsigned long long seed, double sigma) {
    API_LOCK;
    engine_init();
    const ScratchLayout &L = E.L;
    const int nt = prepare_members(hs, n);
    std::vector<FieldRef> src, dst;
    for (int k = 0; k < KX; k++) src.push_back(REF_SCR | (L.tg + (long long)k * NG)), dst.push_back(REF_SCR | (L.sfwd + (long long)k * NSP));
    for (int t0 = 0; t0 < nt; t0 += E.chunk_tiles) {
        const int ntc = std::min(E.chunk_tiles, nt - t0);
        Ctx c = make_ctx(E.d_tiles + t0, E.d_masks + t0, ntc);
        k_noise_grid<<<dim3(NG / 4, ntc), 128, 0, E.stream>>>(c, L.tg, seed, sigma, KX);
        run_forward_plain(c, src, dst);
        k_add_spec<<<dim3(256, ntc), 256, 0, E.stream>>>(c, REF_SCR | L.sfwd, E.off[V_t], NSP * KX);
        COUNT(2);
    }
    CK(cudaStreamSynchronize(E.stream));
    return 0;
}
// transform_spectral2grid for a list of members in one go (chunked)
int spdy_batch_spectral2grid(const int64_t *hs, int n) {
    API_LOCK;
    engine_init();
    const int nt = prepare_members(hs, n);
    for (int t0 = 0; t0 < nt; t0 += E.chunk_tiles) {
        Ctx c = make_ctx(E.d_tiles + t0, E.d_masks + t0, std::min(E.chunk_tiles, nt - t0));
        s2g_ctx(c);
    }
    CK(cudaStreamSynchronize(E.stream));
    return 0;
}
int spdy_profile_intermediate(int on) {
    API_LOCK;
    g_profile_intermediate = on != 0;
    return 0;
}
// one model step of the listed members with device events between kernel classes; ms[10] per class (summed over chunks):
// forcing, pre-ops, legendre_inv, fft_inv, grid_dyn, physics, fft_fwd, legendre_dir, spec_step, post
int spdy_profile_step(const int64_t *hs, const int64_t *cs, int n, float *ms, int *err) {
    API_LOCK;
    engine_init();
    if (!P.made) {
        for (int i = 0; i < 64; i++) CK(cudaEventCreate(&P.ev[i]));
        P.made = true;
    }
    const int saved = E.chunk_tiles;
    for (int i = 0; i < PC_COUNT; i++) ms[i] = 0.f;
    // profile chunk by chunk: step_members with one chunk at a time would change semantics, so mark inside the loop
    P.on = true, P.n = 0;
    // limit event usage: only the first chunk of the step is instrumented (events are reused per class otherwise)
    E.chunk_tiles = saved;
    step_members(hs, cs, n, 1, err, true);
    CK(cudaStreamSynchronize(E.stream));  // the per-step call returns before the step's last kernel has finished
    P.on = false;
    int start = -1;
    for (int i = 0; i < P.n; i++) {
        if (P.cls[i] < 0) { start = i; continue; }
        if (start < 0 || i == 0) continue;
        float t;
        CK(cudaEventElapsedTime(&t, P.ev[i - 1], P.ev[i]));
        ms[P.cls[i]] += t;
    }
    return P.n;
}
}  // extern "C"

namespace spdy {
struct PhysSeg {
    long long src, dst, len;
};
// member i of the chunk (tile t, lane) receives column set (first + 32 t + lane) % nsets
__global__ void __launch_bounds__(256) k_scatter_sets(const Ctx c, const double *__restrict__ sets, long long setlen, int nsets,
                                                      const PhysSeg seg) {
    const int t = blockIdx.y;
    const long long e = blockIdx.x * 8ll + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (e >= seg.len) return;
    const int set = (t * TILE + lane) % nsets;
    *(scp(c, t, seg.dst + e, lane)) = sets[(size_t)set * setlen + seg.src + e];
}
}  // namespace spdy

extern "C" {
// BASELINE config 5: the column physics alone (physics.f90:103-231: convection, large-scale condensation,
// short-wave / long-wave radiation, surface fluxes, vertical diffusion) on synthetic columns.  `sets` holds nsets column
// sets of 45 x (96,48) doubles each: ug8, vg8, pslg, utend8, vtend8 (lowest level / 2-D), then tg, qg, phig, ttend, qtend
// (8 levels each); member i of the list works on set i % nsets, its surface and forcing fields come from its state.
// reps launches per phase; ms[0] = mean device time of a short-wave step (compute_shortwave), ms[1] = long-wave-only step.
int spdy_bench_physics(const int64_t *states, int n, const double *sets, int nsets, int reps, float *ms) {
    API_LOCK;
    engine_init();
    const ScratchLayout &L = E.L;
    const int nt = prepare_members(states, n);
    if (nt == 0 || nt > E.chunk_tiles || nsets < 1) return -1;
    Ctx c = make_ctx(E.d_tiles, E.d_masks, nt);
    const long long setlen = 45ll * NG, G3 = (long long)NG * KX;
    double *d_sets = nullptr;
    CK(cudaMalloc(&d_sets, (size_t)nsets * setlen * sizeof(double)));
    CK(cudaMemcpy(d_sets, sets, (size_t)nsets * setlen * sizeof(double), cudaMemcpyHostToDevice));
    const PhysSeg segs[10] = {{0, L.pug8, NG}, {NG, L.pvg8, NG}, {2ll * NG, L.pslg, NG}, {3ll * NG, L.utend + 7ll * NG, NG},
                              {4ll * NG, L.vtend + 7ll * NG, NG}, {5ll * NG, L.ptg, G3}, {5ll * NG + G3, L.pqg, G3},
                              {5ll * NG + 2 * G3, L.pphig, G3}, {5ll * NG + 3 * G3, L.ttend, G3}, {5ll * NG + 4 * G3, L.trtend, G3}};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)), CK(cudaEventCreate(&e1));
    for (int phase = 0; phase < 2; phase++) {
        k_set_slot<<<nt, 32, 0, E.stream>>>(c, SL_SW, phase == 0 ? 1.0 : 0.0);
        float acc = 0.f;
        for (int r = -2; r < reps; r++) {  // two warm-up launches; the in/out tendencies are re-initialised every time
            for (const PhysSeg &sg : segs)
                k_scatter_sets<<<dim3((unsigned)((sg.len + 7) / 8), nt), 256, 0, E.stream>>>(c, d_sets, setlen, nsets, sg);
            CK(cudaEventRecord(e0, E.stream));
            launch_physics(E.stream, c, L, nullptr);
            CK(cudaEventRecord(e1, E.stream));
            COUNT(11);
            CK(cudaStreamSynchronize(E.stream));
            float t;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (r >= 0) acc += t;
        }
        ms[phase] = reps > 0 ? acc / reps : 0.f;
    }
    CK(cudaGetLastError());
    CK(cudaEventDestroy(e0)), CK(cudaEventDestroy(e1));
    CK(cudaFree(d_sets));
    return 0;
}
// bracket a region for `ncu --profile-from-start off`
int spdy_profiler_start(void) { API_LOCK; return (int)cudaProfilerStart(); }
int spdy_profiler_stop(void) { API_LOCK; return (int)cudaProfilerStop(); }
}

// speedy-b200: host engine -- device arenas, member/control handles, the per-step kernel schedule, registry
// accessors and the C ABI (include/speedy_b200.h).  Mirrors the reference driver surface
// (registry/templates/speedy_driver.f90.j2) and the orchestration of speedy.f90:20-74, initialization.f90:13-91,
// time_stepping.f90:13-27 and prognostics.f90:125-219.  There is no CPU fallback: every entry point that
// computes needs a CUDA device and aborts loudly without one.
#include <cuda_profiler_api.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/speedy_b200.h"
#include "kernels.h"
#include "tables.h"

namespace spdy {

static inline void ck(cudaError_t e, const char *file, int line, const char *what) {
    if (e != cudaSuccess) {
        fprintf(stderr, "speedy-b200: CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e), file, line, what);
        abort();
    }
}
#define CK(call) ck((call), __FILE__, __LINE__, #call)

static long long g_launches = 0;
#define COUNT(n) (g_launches += (n))

// The engine keeps process-global state (handle tables, staging buffers, cached descriptors, graph executables) and
// ctypes releases the GIL during calls: every C-ABI entry point takes this lock, so Python threads driving different
// Speedy instances serialise on the library instead of racing (entry points call each other: recursive).
static std::recursive_mutex g_api_mutex;
#define API_LOCK std::lock_guard<std::recursive_mutex> api_lock_(g_api_mutex)

// optional per-kernel-class device timing (spdy_profile_step): events between the stages of one step
enum ProfClass { PC_FORCING = 0, PC_PREOPS, PC_LEG_INV, PC_FFT_INV, PC_GRID_DYN, PC_PHYSICS, PC_FFT_FWD, PC_LEG_DIR,
                 PC_SPEC_STEP, PC_POST, PC_COUNT };
struct Profiler {
    bool on = false;
    cudaEvent_t ev[64];
    int cls[64];
    int n = 0;
    bool made = false;
};
static Profiler P;
static void prof_mark(cudaStream_t s, int cls) {  // marks the END of a stage of class `cls` (cls < 0: start)
    if (!P.on || P.n >= 64) return;
    cudaEventRecord(P.ev[P.n], s);
    P.cls[P.n] = cls;
    P.n++;
}

// ---------------------------------------------------------------------------------------------- layouts
ScratchLayout make_scratch_layout() {
    ScratchLayout L;
    long long o = 0;
    auto take = [&](long long n) { long long r = o; o += n; return r; };
    const long long G3 = (long long)NG * KX;
    L.ug = take(G3), L.vg = take(G3), L.tg = take(G3), L.vorg = take(G3), L.divg = take(G3), L.trg = take(G3);
    L.ptg = take(G3), L.pqg = take(G3), L.pphig = take(G3), L.pug8 = take(NG), L.pvg8 = take(NG), L.pslg = take(NG);
    L.px = take(NG), L.py = take(NG), L.psdtg = take(NG);
    L.utend = take(G3), L.vtend = take(G3), L.ttend = take(G3), L.trtend = take(G3);
    L.ucos = take((long long)NSP * KX), L.vcos = take((long long)NSP * KX);
    L.ucosp8 = take(NSP), L.vcosp8 = take(NSP), L.dpx = take(NSP), L.dpy = take(NSP);
    L.sfwd = take((long long)NSP * 80);
    L.four = take((long long)NFOUR * 80);
    L.diagp = take(2ll * KX * (MX - 1));
    L.total_base = o;
    // SPPT fields at the end: part of the arena only while SPPT is switched on (spdy_set_sppt)
    L.spptg = take(G3), L.tdyn = take(G3), L.qdyn = take(G3), L.udyn8 = take(NG), L.vdyn8 = take(NG);
    L.total_sppt = o;
    L.total = L.total_base;
    return L;
}

struct Datetime {
    int y, mo, d, h, mi;
    bool alive;
};
struct Control {
    Datetime model, start, end;
    int month_idx;
    bool alive;
    int bound_member;  // member whose device slots mirror this control (-1: none)
};
struct Member {
    int tile, lane;
    bool alive, initialized;
    int n_months;  // -1: sst_anom not allocated
    int current_step;
    int bound_ctl;
    float lon[IX], lat[IL], lev[KX];
};

struct Engine {
    bool ready = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_err = nullptr;
    bool last_ms_pending = false;  // ev1 recorded, elapsed time not read yet
    ConstTables C;
    GlobTables *hG = nullptr, *dG = nullptr;
    ScratchLayout L;
    // arenas
    double *st = nullptr, *sst = nullptr, *scr = nullptr;
    long long st_elems = 0, sst_elems = 0;
    int cap_tiles = 0, sst_months = 0, scr_tiles = 0, chunk_tiles = 64;
    long long off[SPDY_NVARS], nelem[SPDY_NVARS], off_tcorh = 0, off_qcorh = 0, off_slots = 0, off_sppt = 0;
    // handles
    std::vector<Member> members;
    std::vector<Control> controls;
    std::vector<Datetime> dates;
    std::vector<int> free_members, free_controls, free_dates;
    // chunk descriptors
    int *d_tiles = nullptr;
    unsigned *d_masks = nullptr;
    int desc_cap = 0;
    std::vector<int64_t> cached_handles;
    std::vector<int> cached_tiles;
    std::vector<unsigned> cached_masks;
    // transform descriptor lists
    InvDesc *d_inv[2] = {nullptr, nullptr};  // j2 = 1, 2 (77 fields; 85 with the SPPT pattern levels appended)
    bool sppt_on = false;
    int diag_out = 1;  // Ctx::diag_out of the step being launched (step_members)
    int *d_outer = nullptr;  // Ctx::outer_zero flags of the current multi-step call: two per tile (ensure_err_capacity)
    bool outer_on = false;
    unsigned long long sppt_seed = 0;
    FwdDesc *d_fwd[FM_NMODES] = {};
    FwdDesc *d_fwd_all = nullptr;  // all fields of the step in one list (FwdDesc::mode set), two-operand modes first
    int n_fwd_all = 0;
    int n_fwd[FM_NMODES] = {};
    FwdOut *d_out = nullptr;
    InvDesc *d_inv_s2g = nullptr;  // the 41 fields of transform_spectral2grid (prognostics.f90:125-154)
    InvDesc *d_inv_tmp = nullptr;
    FwdDesc *d_fwd_tmp = nullptr;
    FwdOut *d_out_tmp = nullptr;
    // staging
    double *d_stage = nullptr, *h_stage = nullptr;
    size_t stage_elems = 0;
    int *d_err = nullptr, *h_err = nullptr;
    int err_cap = 0;
    float last_ms = 0.f;
};
static Engine E;

static Ctx make_ctx(const int *d_tiles, const unsigned *d_masks, int ntiles) {
    Ctx c;
    c.st = E.st, c.scr = E.scr, c.sst = E.sst, c.tiles = d_tiles, c.masks = d_masks, c.G = E.dG;
    c.st_elems = E.st_elems, c.scr_elems = E.L.total, c.sst_elems = E.sst_elems;
    for (int v = 0; v < SPDY_NVARS; v++) c.off[v] = E.off[v];
    c.off_tcorh = E.off_tcorh, c.off_qcorh = E.off_qcorh, c.off_slots = E.off_slots, c.off_sppt = E.off_sppt;
    c.ntiles = ntiles, c.sst_months = E.sst_months, c.diag_out = E.diag_out;
    // the flags are indexed like the tile list of the call
    c.outer_zero = (E.outer_on && d_tiles >= E.d_tiles) ? E.d_outer + 2 * (d_tiles - E.d_tiles) : nullptr;
    return c;
}

static void build_descriptor_lists() {
    const ScratchLayout &L = E.L;
    for (int j2 = 1; j2 <= 2; j2++) {
        std::vector<InvDesc> v;
        const long long tl = (long long)(j2 - 1) * NSP * KX;
        auto add = [&](FieldRef src, long long dst, int kcos) { v.push_back(InvDesc{src, dst, kcos, 0}); };
        for (int k = 0; k < KX; k++) add(E.off[V_vor] + tl + (long long)k * NSP, L.vorg + (long long)k * NG, 1);
        for (int k = 0; k < KX; k++) add(E.off[V_div] + tl + (long long)k * NSP, L.divg + (long long)k * NG, 1);
        for (int k = 0; k < KX; k++) add(E.off[V_t] + tl + (long long)k * NSP, L.tg + (long long)k * NG, 1);
        for (int k = 0; k < KX; k++) add(E.off[V_tr] + tl + (long long)k * NSP, L.trg + (long long)k * NG, 1);
        for (int k = 0; k < KX; k++) add(REF_SCR | (L.ucos + (long long)k * NSP), L.ug + (long long)k * NG, 2);
        for (int k = 0; k < KX; k++) add(REF_SCR | (L.vcos + (long long)k * NSP), L.vg + (long long)k * NG, 2);
        add(REF_SCR | L.dpx, L.px, 2);
        add(REF_SCR | L.dpy, L.py, 2);
        // physics inputs: time level 1 (physics.f90:89-101); u, v are only needed at the lowest level
        for (int k = 0; k < KX; k++) add(E.off[V_t] + (long long)k * NSP, L.ptg + (long long)k * NG, 1);
        for (int k = 0; k < KX; k++) add(E.off[V_tr] + (long long)k * NSP, L.pqg + (long long)k * NG, 1);
        for (int k = 0; k < KX; k++) add(E.off[V_phi] + (long long)k * NSP, L.pphig + (long long)k * NG, 1);
        add(REF_SCR | L.ucosp8, L.pug8, 2);
        add(REF_SCR | L.vcosp8, L.pvg8, 2);
        add(E.off[V_ps], L.pslg, 1);
        // SPPT pattern (used only when switched on: the list is then launched with 85 instead of 77 fields)
        for (int k = 0; k < KX; k++) add(E.off_sppt + (long long)k * NSP, L.spptg + (long long)k * NG, 1);
        CK(cudaMalloc(&E.d_inv[j2 - 1], v.size() * sizeof(InvDesc)));
        CK(cudaMemcpy(E.d_inv[j2 - 1], v.data(), v.size() * sizeof(InvDesc), cudaMemcpyHostToDevice));
    }
    std::vector<FwdDesc> f[FM_NMODES];
    std::vector<FwdOut> outs(FW_COUNT);
    for (int i = 0; i < FW_COUNT; i++) outs[i].dst = REF_SCR | (L.sfwd + (long long)i * NSP);
    for (int k = 0; k < KX; k++) {
        const long long g = (long long)k * NG;
        f[FM_COS].push_back(FwdDesc{REF_SCR | (L.utend + g), 0, 0.0, 2, FW_SU + k});
        f[FM_COS].push_back(FwdDesc{REF_SCR | (L.vtend + g), 0, 0.0, 2, FW_SV + k});
        f[FM_KE].push_back(FwdDesc{REF_SCR | (L.ug + g), REF_SCR | (L.vg + g), 0.0, 2, FW_KE + k});
        f[FM_FLUXT].push_back(FwdDesc{REF_SCR | (L.ug + g), REF_SCR | (L.tg + g), E.C.tref[k], 2, FW_UT + k});
        f[FM_FLUXT].push_back(FwdDesc{REF_SCR | (L.vg + g), REF_SCR | (L.tg + g), E.C.tref[k], 2, FW_VT + k});
        f[FM_FLUX].push_back(FwdDesc{REF_SCR | (L.ug + g), REF_SCR | (L.trg + g), 0.0, 2, FW_UQ + k});
        f[FM_FLUX].push_back(FwdDesc{REF_SCR | (L.vg + g), REF_SCR | (L.trg + g), 0.0, 2, FW_VQ + k});
        f[FM_PLAIN].push_back(FwdDesc{REF_SCR | (L.ttend + g), 0, 0.0, 2, FW_TT + k});
        f[FM_PLAIN].push_back(FwdDesc{REF_SCR | (L.trtend + g), 0, 0.0, 2, FW_QT + k});
    }
    f[FM_PLAIN].push_back(FwdDesc{REF_SCR | L.psdtg, 0, 0.0, 2, FW_PS});
    for (int m = 0; m < FM_NMODES; m++) {
        E.n_fwd[m] = (int)f[m].size();
        CK(cudaMalloc(&E.d_fwd[m], f[m].size() * sizeof(FwdDesc)));
        CK(cudaMemcpy(E.d_fwd[m], f[m].data(), f[m].size() * sizeof(FwdDesc), cudaMemcpyHostToDevice));
    }
    {
        std::vector<FwdDesc> all;
        const int order[FM_NMODES] = {FM_FLUXT, FM_FLUX, FM_KE, FM_PLAIN, FM_COS};
        for (int q = 0; q < FM_NMODES; q++)
            for (FwdDesc d : f[order[q]]) d.mode = order[q], d.pad = 0, all.push_back(d);
        E.n_fwd_all = (int)all.size();
        CK(cudaMalloc(&E.d_fwd_all, all.size() * sizeof(FwdDesc)));
        CK(cudaMemcpy(E.d_fwd_all, all.data(), all.size() * sizeof(FwdDesc), cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&E.d_out, outs.size() * sizeof(FwdOut)));
    CK(cudaMemcpy(E.d_out, outs.data(), outs.size() * sizeof(FwdOut), cudaMemcpyHostToDevice));
    {
        std::vector<InvDesc> v;
        for (int k = 0; k < KX; k++) {
            v.push_back(InvDesc{REF_SCR | (L.ucos + (long long)k * NSP), L.ug + (long long)k * NG, 2, 0});
            v.push_back(InvDesc{REF_SCR | (L.vcos + (long long)k * NSP), L.vg + (long long)k * NG, 2, 0});
            v.push_back(InvDesc{E.off[V_t] + (long long)k * NSP, L.tg + (long long)k * NG, 1, 0});
            v.push_back(InvDesc{E.off[V_tr] + (long long)k * NSP, L.trg + (long long)k * NG, 1, 0});
            v.push_back(InvDesc{E.off[V_phi] + (long long)k * NSP, L.pphig + (long long)k * NG, 1, 0});
        }
        v.push_back(InvDesc{E.off[V_ps], L.pslg, 1, 0});
        CK(cudaMalloc(&E.d_inv_s2g, v.size() * sizeof(InvDesc)));
        CK(cudaMemcpy(E.d_inv_s2g, v.data(), v.size() * sizeof(InvDesc), cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&E.d_inv_tmp, 128 * sizeof(InvDesc)));
    CK(cudaMalloc(&E.d_fwd_tmp, 128 * sizeof(FwdDesc)));
    CK(cudaMalloc(&E.d_out_tmp, 128 * sizeof(FwdOut)));
}

static void engine_init() {
    if (E.ready) return;
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0) {
        fprintf(stderr, "speedy-b200: no CUDA device available (%s); this library has no CPU path\n",
                cudaGetErrorString(err));
        abort();
    }
    CK(cudaSetDevice(E.device));
    CK(cudaStreamCreate(&E.stream));
    CK(cudaEventCreate(&E.ev0));
    CK(cudaEventCreate(&E.ev1));
    CK(cudaEventCreateWithFlags(&E.ev_err, cudaEventDisableTiming));
    E.hG = new GlobTables();
    build_tables(E.C, *E.hG);
    upload_const_tables(E.C);
    CK(cudaMalloc(&E.dG, sizeof(GlobTables)));
    CK(cudaMemcpy(E.dG, E.hG, sizeof(GlobTables), cudaMemcpyHostToDevice));
    E.L = make_scratch_layout();
    // state layout: all array variables of the registry that live on the device
    long long o = 0;
    for (int v = 0; v < SPDY_NVARS; v++) {
        const spdy_vardef &d = SPDY_VARDEFS[v];
        E.off[v] = -1, E.nelem[v] = 0;
        if (d.ndim == 0 || d.kind == SPDY_F4 || v == V_sst_anom) continue;
        long long n = 1;
        for (int q = 0; q < d.ndim; q++) n *= d.dims[q];
        if (d.kind == SPDY_C16) n *= 2;
        E.off[v] = o, E.nelem[v] = n;
        o += n;
    }
    E.off_tcorh = o, o += NSP;
    E.off_qcorh = o, o += NSP;
    E.off_slots = o, o += SL_COUNT;
    E.off_sppt = o, o += (long long)NSP * KX;  // SPPT AR(1) pattern of the member (sppt.cu)
    E.st_elems = o;
    E.sst_elems = 0;
    const char *ct = getenv("SPDY_CHUNK_TILES");
    if (ct && atoi(ct) > 0) E.chunk_tiles = atoi(ct);
    build_descriptor_lists();
    E.stage_elems = (size_t)NG * 12;
    CK(cudaMalloc(&E.d_stage, E.stage_elems * sizeof(double)));
    CK(cudaMallocHost(&E.h_stage, E.stage_elems * sizeof(double)));
    E.ready = true;
}

static void drop_step_graphs();  // cached graph executables hold arena addresses
static void ensure_state_tiles(int ntiles) {
    if (ntiles <= E.cap_tiles) return;
    drop_step_graphs();
    int ncap = std::max(ntiles, E.cap_tiles ? E.cap_tiles * 2 : 1);
    const size_t per_tile = (size_t)E.st_elems * TILE * sizeof(double);
    double *nst = nullptr;
    CK(cudaMalloc(&nst, per_tile * ncap));
    CK(cudaMemsetAsync(nst, 0, per_tile * ncap, E.stream));
    if (E.st) {
        CK(cudaMemcpyAsync(nst, E.st, per_tile * E.cap_tiles, cudaMemcpyDeviceToDevice, E.stream));
        CK(cudaStreamSynchronize(E.stream));
        CK(cudaFree(E.st));
    }
    E.st = nst;
    if (E.sst_months > 0) {
        const size_t sper = (size_t)E.sst_elems * TILE * sizeof(double);
        double *ns = nullptr;
        CK(cudaMalloc(&ns, sper * ncap));
        CK(cudaMemsetAsync(ns, 0, sper * ncap, E.stream));
        if (E.sst) {
            CK(cudaMemcpyAsync(ns, E.sst, sper * E.cap_tiles, cudaMemcpyDeviceToDevice, E.stream));
            CK(cudaStreamSynchronize(E.stream));
            CK(cudaFree(E.sst));
        }
        E.sst = ns;
    }
    E.cap_tiles = ncap;
    E.cached_handles.clear();
}

static void ensure_sst_months(int slabs) {  // slabs = n_months + 2
    if (slabs <= E.sst_months) return;
    drop_step_graphs();
    const long long nelems = (long long)slabs * NG;
    double *ns = nullptr;
    const int cap = std::max(E.cap_tiles, 1);
    CK(cudaMalloc(&ns, (size_t)nelems * TILE * sizeof(double) * cap));
    CK(cudaMemsetAsync(ns, 0, (size_t)nelems * TILE * sizeof(double) * cap, E.stream));
    if (E.sst) {  // re-stride existing tiles
        for (int t = 0; t < E.cap_tiles; t++)
            CK(cudaMemcpyAsync(ns + (size_t)t * nelems * TILE, E.sst + (size_t)t * E.sst_elems * TILE,
                               (size_t)E.sst_elems * TILE * sizeof(double), cudaMemcpyDeviceToDevice, E.stream));
        CK(cudaStreamSynchronize(E.stream));
        CK(cudaFree(E.sst));
    }
    E.sst = ns, E.sst_months = slabs, E.sst_elems = nelems;
}

static void ensure_scratch(int ntiles) {
    ntiles = std::min(ntiles, E.chunk_tiles);
    if (ntiles <= E.scr_tiles) return;
    drop_step_graphs();
    if (E.scr) CK(cudaFree(E.scr));
    CK(cudaMalloc(&E.scr, (size_t)E.L.total * TILE * sizeof(double) * ntiles));
    CK(cudaMemsetAsync(E.scr, 0, (size_t)E.L.total * TILE * sizeof(double) * ntiles, E.stream));
    E.scr_tiles = ntiles;
}

static void ensure_desc(int n) {
    if (n <= E.desc_cap) return;
    drop_step_graphs();
    if (E.d_tiles) CK(cudaFree(E.d_tiles)), CK(cudaFree(E.d_masks));
    E.desc_cap = std::max(n, 64);
    CK(cudaMalloc(&E.d_tiles, E.desc_cap * sizeof(int)));
    CK(cudaMalloc(&E.d_masks, E.desc_cap * sizeof(unsigned)));
    E.cached_handles.clear();
}

static Member *member_of(int64_t h) {
    if (h < 1 || h > (int64_t)E.members.size() || !E.members[h - 1].alive) return nullptr;
    return &E.members[h - 1];
}
static Control *control_of(int64_t h) {
    if (h < 1 || h > (int64_t)E.controls.size() || !E.controls[h - 1].alive) return nullptr;
    return &E.controls[h - 1];
}

// tiles/masks of a member list; uploads only when the list changed
static int prepare_members(const int64_t *hs, int n) {
    if ((int)E.cached_handles.size() == n && n > 0 && memcmp(E.cached_handles.data(), hs, n * sizeof(int64_t)) == 0)
        return (int)E.cached_tiles.size();
    std::map<int, unsigned> tm;
    for (int i = 0; i < n; i++) {
        Member *m = member_of(hs[i]);
        if (!m) continue;
        tm[m->tile] |= 1u << m->lane;
    }
    E.cached_tiles.clear(), E.cached_masks.clear();
    for (auto &kv : tm) E.cached_tiles.push_back(kv.first), E.cached_masks.push_back(kv.second);
    const int nt = (int)E.cached_tiles.size();
    ensure_desc(nt);
    if (nt) {
        CK(cudaMemcpyAsync(E.d_tiles, E.cached_tiles.data(), nt * sizeof(int), cudaMemcpyHostToDevice, E.stream));
        CK(cudaMemcpyAsync(E.d_masks, E.cached_masks.data(), nt * sizeof(unsigned), cudaMemcpyHostToDevice, E.stream));
        CK(cudaStreamSynchronize(E.stream));  // source vectors may be rebuilt by the next call
    }
    E.cached_handles.assign(hs, hs + n);
    ensure_scratch(nt);
    return nt;
}

// ---- small device helpers -----------------------------------------------------------------------------------
__global__ void k_gather(const double *arena, long long tile_elems, int tile, int lane, long long off, long long n,
                         double *dst) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = arena[((long long)tile * tile_elems + off + i) * TILE + lane];
}
__global__ void k_scatter(double *arena, long long tile_elems, int tile, int lane, long long off, long long n,
                          const double *src) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) arena[((long long)tile * tile_elems + off + i) * TILE + lane] = src[i];
}
__global__ void k_spec_trunc(const Ctx c, FieldRef f, int nfields) {  // zero l > trunc (spectral.f90:309-314)
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    if (q >= NSPC) return;
    const double trf = c.G->trfilt[q];
    double *p = refp(c, t, f, lane) + (size_t)(2 * (q % MX) + M2 * (q / MX)) * TILE;
    for (int i = 0; i < nfields; i++) {
        if (trf == 0.0) p[(size_t)i * NSP * TILE] = 0.0, p[(size_t)i * NSP * TILE + TILE] = 0.0;
    }
}
// prognostics.f90:141-150 unit conversions after the inverse transforms (masked)
__global__ void __launch_bounds__(128) k_s2g_finish(const Ctx c, const ScratchLayout L) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    if (!lane_active(c, t, lane)) return;
    const size_t e = (size_t)q * TILE, lev = (size_t)NG * TILE;
    double *u = stp(c, t, c.off[V_u_grid], lane) + e, *v = stp(c, t, c.off[V_v_grid], lane) + e,
           *tg = stp(c, t, c.off[V_t_grid], lane) + e, *qg = stp(c, t, c.off[V_q_grid], lane) + e,
           *pg = stp(c, t, c.off[V_phi_grid], lane) + e;
    for (int k = 0; k < KX; k++) {
        u[k * lev] = *(scp(c, t, L.ug, lane) + e + k * lev);
        v[k * lev] = *(scp(c, t, L.vg, lane) + e + k * lev);
        tg[k * lev] = *(scp(c, t, L.tg, lane) + e + k * lev);
        qg[k * lev] = *(scp(c, t, L.trg, lane) + e + k * lev) * FL(1.0e-3);
        pg[k * lev] = *(scp(c, t, L.pphig, lane) + e + k * lev) / FL(9.81);
    }
    *(stp(c, t, c.off[V_ps_grid], lane) + e) = FL(1.e+5) * exp(*(scp(c, t, L.pslg, lane) + e));
}
__global__ void __launch_bounds__(128) k_g2s_prepare(const Ctx c, const long long g0) {  // log(ps_grid/p0)
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    const size_t e = (size_t)q * TILE;
    *(scp(c, t, g0, lane) + e) = log(*(stp(c, t, c.off[V_ps_grid], lane) + e) / FL(1.e+5));
}
// prognostics.f90:165-174: vel2vort + unit conversions into time level 1 (masked)
__global__ void __launch_bounds__(128) k_g2s_finish(const Ctx c, const ScratchLayout L) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), t = blockIdx.y;
    if (q >= NSPC || !lane_active(c, t, lane)) return;
    const int m = q % MX, n = q / MX;
    const size_t e = (size_t)(2 * m + M2 * n) * TILE, lev = (size_t)NSP * TILE;
    const double *F = scp(c, t, L.sfwd, lane) + e;
    for (int k = 0; k < KX; k++) {
        C2 vo, dv;
        vdspec_elem(c.G, F + (0 + k) * lev, F + (8 + k) * lev, m, n, vo, dv);
        st2(stp(c, t, c.off[V_vor], lane) + e + k * lev, vo);
        st2(stp(c, t, c.off[V_div], lane) + e + k * lev, dv);
        st2(stp(c, t, c.off[V_t], lane) + e + k * lev, ld2(F + (16 + k) * lev));
        const C2 qq = ld2(F + (24 + k) * lev);
        st2(stp(c, t, c.off[V_tr], lane) + e + k * lev, C2{qq.r / FL(1.0e-3), qq.i / FL(1.0e-3)});
        st2(stp(c, t, c.off[V_phi], lane) + e + k * lev, ld2(F + (32 + k) * lev) * FL(9.81));
    }
    st2(stp(c, t, c.off[V_ps], lane) + e, ld2(F + 40 * lev));
}
__global__ void __launch_bounds__(256) k_fill_lane(double *arena, long long tile_elems, int tile, int lane, long long n, double v) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) arena[((long long)tile * tile_elems + i) * TILE + lane] = v;
}
__global__ void k_set_slot(const Ctx c, int s, double v) {
    const int lane = threadIdx.x, t = blockIdx.x;
    if (lane_active(c, t, lane)) slot(c, t, lane, s) = v;
}
// ensemble partial sums.  Stage 1: warp per (grid element, group of ENS_TG tiles), lane = member: the
// group's loads are issued together (a tile is 375 MB away from the next one: one load at a time is pure DRAM latency),
// warp-shuffle reduction over the lanes, one partial per (group, element).  Stage 2: thread per element adds the
// partials in group order -- deterministic, unlike atomics.
constexpr int ENS_TG = 8;
__global__ void __launch_bounds__(256) k_ens_sums(const Ctx c, long long off, long long n, const double *shift,
                                                  double *part /* [groups][2][n] */) {
    const int lane = threadIdx.x & 31, g = blockIdx.y;
    const long long i = blockIdx.x * 8ll + (threadIdx.x >> 5);
    if (i >= n) return;
    const double sh = shift ? shift[i] : 0.0;
    double x[ENS_TG];
    bool act[ENS_TG];
#pragma unroll
    for (int q = 0; q < ENS_TG; q++) {
        const int t = g * ENS_TG + q;
        act[q] = t < c.ntiles && lane_active(c, t, lane);
        x[q] = act[q] ? *(stp(c, t, off + i, lane)) : 0.0;
    }
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int q = 0; q < ENS_TG; q++)
        if (act[q]) s1 += x[q], s2 += (x[q] - sh) * (x[q] - sh);
    for (int o = 16; o; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o), s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (lane == 0) part[((size_t)g * 2) * n + i] = s1, part[((size_t)g * 2 + 1) * n + i] = s2;
}
__global__ void __launch_bounds__(256) k_ens_reduce(const double *__restrict__ part, int groups, long long n, double *sum,
                                                    double *sumsq) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= n) return;
    double s1 = 0.0, s2 = 0.0;
    for (int g = 0; g < groups; g++) s1 += part[((size_t)g * 2) * n + i], s2 += part[((size_t)g * 2 + 1) * n + i];
    sum[i] = s1, sumsq[i] = s2;
}

// ---------------------------------------------------------------------------------------- kernel schedules
// SPDY_FUSED selects the transform kernels: default = the fused kernels (spec -> grid: k_spec2grid_mma4, fused_mma3.cu + fused_mma4.cu;
// grid -> spec: ONE mixed-mode launch of k_grid2spec_mma2, fused_mma2.cu); SPDY_FUSED=0 = separate Legendre and FFT
// kernels both ways (transforms.cu), kept as the unfused cross-check of the parity tests.
static bool fused_transforms() {
    static int v = -1;
    if (v < 0) {
        const char *s = getenv("SPDY_FUSED");
        v = (s && atoi(s) == 0) ? 0 : 1;
    }
    return v != 0;
}
// SPDY_FUSE_PHYS=1 runs the grid-point dynamics and the column physics of a column in one kernel.  Measured: 34 loads
// and 34 stores per column less DRAM traffic but the same time (0.884 ms vs 0.234 + 0.650 ms at 512 members): the
// physics is bound by latency/issue at 12 warps per SM, not by bandwidth, so the default keeps the two kernels.
static bool fuse_dyn_physics() {
    static int v = -1;
    if (v < 0) {
        const char *s = getenv("SPDY_FUSE_PHYS");
        v = s ? atoi(s) : 0;
    }
    return v != 0;
}
static void run_inverse(const Ctx &c, const InvDesc *d, int n) {
    if (fused_transforms()) {  // parity-pure DMMA over latitude octets + two-stage FFT, Fourier rows stay in shared memory
        launch_spec2grid_mma4(E.stream, c, d, n);
        prof_mark(E.stream, PC_FFT_INV);
        COUNT(1);
        return;
    }
    launch_legendre_inv(E.stream, c, d, n, E.L.four);
    prof_mark(E.stream, PC_LEG_INV);
    launch_fft_inv(E.stream, c, d, n, E.L.four);
    prof_mark(E.stream, PC_FFT_INV);
    COUNT(2);
}
// sparse: the outputs feed the model's spectral step, which reads the rows inside the nsh2 mask only
static void run_forward_lists(const Ctx &c, FwdDesc *const *lists, const int *counts, const FwdOut *outs, int nout,
                              int sparse = 0) {
    if (fused_transforms() && lists == E.d_fwd) {  // the step's 73 fields in ONE launch of the fused forward kernel
        launch_grid2spec_mma2(E.stream, c, FM_ALL, E.d_fwd_all, outs, E.n_fwd_all, sparse);
        COUNT(1);
        prof_mark(E.stream, PC_FFT_FWD);
        return;
    }
    for (int m = 0; m < FM_NMODES; m++)
        if (counts[m]) launch_fft_fwd(E.stream, c, m, lists[m], counts[m], E.L.four), COUNT(1);
    prof_mark(E.stream, PC_FFT_FWD);
    launch_legendre_dir(E.stream, c, outs, nout, E.L.four);
    prof_mark(E.stream, PC_LEG_DIR);
    COUNT(1);
}
// forward transform of `n` plain grid fields given as refs -> spectral refs
static void run_forward_plain(const Ctx &c, const std::vector<FieldRef> &src, const std::vector<FieldRef> &dst, int mode = FM_PLAIN) {
    std::vector<FwdDesc> f(src.size());
    std::vector<FwdOut> o(src.size());
    for (size_t i = 0; i < src.size(); i++) f[i] = FwdDesc{src[i], 0, 0.0, 2, (int)i}, o[i].dst = dst[i];
    CK(cudaMemcpyAsync(E.d_fwd_tmp, f.data(), f.size() * sizeof(FwdDesc), cudaMemcpyHostToDevice, E.stream));
    CK(cudaMemcpyAsync(E.d_out_tmp, o.data(), o.size() * sizeof(FwdOut), cudaMemcpyHostToDevice, E.stream));
    CK(cudaStreamSynchronize(E.stream));
    launch_fft_fwd(E.stream, c, mode, E.d_fwd_tmp, (int)f.size(), E.L.four);
    launch_legendre_dir(E.stream, c, E.d_out_tmp, (int)f.size(), E.L.four);
    COUNT(2);
    CK(cudaStreamSynchronize(E.stream));
}
static void run_inverse_list(const Ctx &c, const std::vector<InvDesc> &v) {
    CK(cudaMemcpyAsync(E.d_inv_tmp, v.data(), v.size() * sizeof(InvDesc), cudaMemcpyHostToDevice, E.stream));
    CK(cudaStreamSynchronize(E.stream));
    run_inverse(c, E.d_inv_tmp, (int)v.size());
    CK(cudaStreamSynchronize(E.stream));
}

// set_forcing (forcing.f90:15-102): imode 0 = initialisation, 1 = daily (lanes flagged by k_control_pre)
static void run_forcing(const Ctx &c, int imode) {
    const ScratchLayout &L = E.L;
    launch_forcing(E.stream, c, imode, L.utend, L.vtend);
    COUNT(1);
    // two forward transforms into scratch, then masked copies into the per-member tcorh / qcorh
    static FwdDesc *d_f = nullptr;
    static FwdOut *d_o = nullptr;
    if (!d_f) {
        FwdDesc f[2] = {FwdDesc{REF_SCR | L.utend, 0, 0.0, 2, 0}, FwdDesc{REF_SCR | L.vtend, 0, 0.0, 2, 1}};
        FwdOut o[2] = {FwdOut{REF_SCR | L.sfwd}, FwdOut{REF_SCR | (L.sfwd + NSP)}};
        CK(cudaMalloc(&d_f, sizeof(f)));
        CK(cudaMalloc(&d_o, sizeof(o)));
        CK(cudaMemcpy(d_f, f, sizeof(f), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_o, o, sizeof(o), cudaMemcpyHostToDevice));
    }
    launch_fft_fwd(E.stream, c, FM_PLAIN, d_f, 2, L.four);
    launch_legendre_dir(E.stream, c, d_o, 2, L.four);
    launch_masked_copy(E.stream, c, REF_SCR | L.sfwd, E.off_tcorh, NSP, imode, 1.0);
    launch_masked_copy(E.stream, c, REF_SCR | (L.sfwd + NSP), E.off_qcorh, NSP, imode, 1.0);
    COUNT(4);
}

static int g_dump_stage = 2;  // tendency dump: 1 = before implicit_terms, 2 = as returned by get_tendencies
// step(j1, j2, dt) of time_stepping.f90:38-147: tendencies + diffusion + time integration
static void run_step_core(const Ctx &c, int j1, int j2, double dt, double eps, int impl_idx, long long dump = -1,
                          int with_control = 0) {
    const ScratchLayout &L = E.L;
    const long long tl2 = (long long)(j2 - 1) * NSP * KX;
    // spectral pre-operators in one launch: geopotential (time level 1), uvspec, grad(ps).  k_spec2grid_mma4 reads the rows
    // inside the nsh2 mask only: uvspec and the gradient skip the other 47 % of each field
    launch_preops(E.stream, c, L, j2, fused_transforms() ? 1 : 0, with_control);
    COUNT(1);
    prof_mark(E.stream, PC_PREOPS);
    if (E.sppt_on) launch_sppt_update(E.stream, c, E.sppt_seed), COUNT(1);
    run_inverse(c, E.d_inv[j2 - 1], E.sppt_on ? 85 : 77);
    if (E.sppt_on) {  // sppt.cu: three small kernels around the (untouched) physics kernel
        launch_grid_dyn(E.stream, c, L);
        prof_mark(E.stream, PC_GRID_DYN);
        launch_sppt_save(E.stream, c, L);
        launch_physics(E.stream, c, L, nullptr);
        launch_sppt_apply(E.stream, c, L);
        prof_mark(E.stream, PC_PHYSICS);
        COUNT(4);
    } else if (fuse_dyn_physics()) {  // one kernel: the column's dynamical tendencies stay in registers (physics.cu)
        prof_mark(E.stream, PC_GRID_DYN);
        launch_dyn_physics(E.stream, c, L);
        prof_mark(E.stream, PC_PHYSICS);
        COUNT(1);
    } else {
        launch_grid_dyn(E.stream, c, L);
        prof_mark(E.stream, PC_GRID_DYN);
        launch_physics(E.stream, c, L, nullptr);
        prof_mark(E.stream, PC_PHYSICS);
        COUNT(2);
    }
    run_forward_lists(c, E.d_fwd, E.n_fwd, E.d_out, FW_COUNT, dump < 0 ? 1 : 0);  // the tendency dump reads every row
    launch_spec_step(E.stream, c, L, (dump >= 0 && g_dump_stage == 1) ? -1 : j1, dt, eps, impl_idx, dump);
    prof_mark(E.stream, PC_SPEC_STEP);
    COUNT(2);  // k_spec_step_vq + k_spec_step_dt
}

// do_single_step (speedy.f90:20-74) for one chunk of tiles
// early_err (per-step driver call, last chunk): the error codes of the whole call are copied to the host right after the
// diagnostics check, BEFORE the coupler kernel, and E.ev_err marks that copy: the host can hand the codes back to the caller
// and launch the next step while the GPU finishes the land / sea models of this one.
static void run_model_step(const Ctx &c, bool any_daily, int *err_out, unsigned *masks, int early_err_tiles = 0) {
    prof_mark(E.stream, -1);
    if (any_daily) {  // the daily forcing kernels read the flags of the step: set them first
        launch_control_pre(E.stream, c);
        COUNT(1);
        run_forcing(c, 1);
    }
    prof_mark(E.stream, PC_FORCING);
    run_step_core(c, 2, 2, 2.0 * H_DELT, FL(0.05), 2, -1, any_daily ? 0 : 1);
    // step counter, check_diagnostics, calendar, sticky error codes of the call (k_diag_final)
    launch_diag(E.stream, c, 2, E.L.diagp, 1, err_out, masks);
    if (early_err_tiles) {
        CK(cudaMemcpyAsync(E.h_err, E.d_err, (size_t)early_err_tiles * TILE * sizeof(int), cudaMemcpyDeviceToHost, E.stream));
        CK(cudaEventRecord(E.ev_err, E.stream));
    }
    launch_couple(E.stream, c, 0);
    prof_mark(E.stream, PC_POST);
    COUNT(3);
}

static void set_slot_host(const Member &m, int s, double v);
static void upload_control(Member &m, Control &ctl) {
    // model_control.f90:73-110,166-186 evaluated on the host for the upload; the device then advances its copy
    static const int dim[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    int cum = 0;
    for (int q = 1; q < ctl.model.mo; q++) cum += dim[q - 1];
    double v[10] = {(double)m.current_step, (double)ctl.model.y, (double)ctl.model.mo, (double)ctl.model.d,
                    (double)ctl.model.h, (double)ctl.model.mi, (double)ctl.month_idx, (double)ctl.model.mo,
                    (double)(((float)ctl.model.d - 0.5f) / (float)dim[ctl.model.mo - 1]),
                    (double)(((float)(cum + ctl.model.d) - 0.5f) / 365.0f)};
    CK(cudaMemcpyAsync(E.d_stage, v, sizeof(v), cudaMemcpyHostToDevice, E.stream));
    k_scatter<<<1, 32, 0, E.stream>>>(E.st, E.st_elems, m.tile, m.lane, E.off_slots + SL_STEP, 10, E.d_stage);
    CK(cudaStreamSynchronize(E.stream));
    set_slot_host(m, SL_CPLDIRTY, 1.0);
}
static void advance_host_date(Control &c) {  // model_control.f90:113-163
    static const int dim[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    Datetime &d = c.model;
    d.mi += 24 * 60 / NSTEPS;
    if (d.mi >= 60) d.mi %= 60, d.h += 1;
    if (d.h >= 24) d.h %= 24, d.d += 1;
    if (d.y % 4 == 0 && d.mo == 2) {
        if (d.d > 29) d.d = 1, d.mo += 1, c.month_idx += 1;
    } else if (d.d > dim[d.mo - 1]) {
        d.d = 1, d.mo += 1, c.month_idx += 1;
    }
    if (d.mo > 12) d.mo = 1, d.y += 1;
}
static double get_slot_host(const Member &m, int s) {
    k_gather<<<1, 32, 0, E.stream>>>(E.st, E.st_elems, m.tile, m.lane, E.off_slots + s, 1, E.d_stage);
    double v;
    CK(cudaMemcpyAsync(&v, E.d_stage, 8, cudaMemcpyDeviceToHost, E.stream));
    CK(cudaStreamSynchronize(E.stream));
    return v;
}
static void set_slot_host(const Member &m, int s, double v) {
    CK(cudaMemcpyAsync(E.d_stage, &v, 8, cudaMemcpyHostToDevice, E.stream));
    k_scatter<<<1, 32, 0, E.stream>>>(E.st, E.st_elems, m.tile, m.lane, E.off_slots + s, 1, E.d_stage);
    CK(cudaStreamSynchronize(E.stream));
}

// ---- CUDA graphs for launch-bound ensembles ------------------------------------------------------------------
// One model step of one chunk is ~20 kernel launches; below ~500 members the gaps between them are a visible
// fraction of the step.  For chunks of at most SPDY_GRAPH_TILES tiles (default 16) the launch sequence of
// run_model_step is captured once per (chunk, daily / regular step, arena addresses) and replayed.  Every kernel
// argument is either a device pointer whose CONTENT may change (tile lists, masks, state) or a value that is part
// of the key; per-member flags (short-wave step, daily forcing) are read on the device.
struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;
};
typedef std::tuple<const void *, const void *, const void *, const void *, const void *, const void *, const void *, long long,
                   int, int, int, int, int>
    StepGraphKey;
static std::map<StepGraphKey, StepGraph> g_step_graphs;
static bool g_eager_done[2] = {false, false};  // statics inside the launchers are initialised by an eager run
static void drop_step_graphs() {
    for (auto &kv : g_step_graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    g_step_graphs.clear();
}
static int graph_max_tiles() {
    static int v = -1;
    if (v < 0) {
        const char *s = getenv("SPDY_GRAPH_TILES");
        v = s ? atoi(s) : 16;
    }
    return v;
}
// returns true if the error codes were copied to the host inside the step (early_err_tiles > 0 and an eager launch)
static bool run_chunk_step(int t0, int ntc, bool any_daily, int early_err_tiles = 0) {
    Ctx c = make_ctx(E.d_tiles + t0, E.d_masks + t0, ntc);
    const bool use_graph = ntc <= graph_max_tiles() && !P.on && g_eager_done[any_daily ? 1 : 0];
    if (!use_graph) {
        run_model_step(c, any_daily, E.d_err + t0 * TILE, E.d_masks + t0, early_err_tiles);
        g_eager_done[any_daily ? 1 : 0] = true;
        return early_err_tiles > 0;
    }
    const StepGraphKey key(E.st, E.scr, E.sst, E.d_tiles, E.d_masks, E.d_err, E.d_outer, E.st_elems, E.sst_months, t0, ntc,
                           any_daily ? 1 : 0, E.diag_out + 2 * (E.outer_on ? 1 : 0));
    auto it = g_step_graphs.find(key);
    if (it == g_step_graphs.end()) {
        StepGraph sg;
        const long long l0 = g_launches;
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(E.stream, cudaStreamCaptureModeThreadLocal));
        run_model_step(c, any_daily, E.d_err + t0 * TILE, E.d_masks + t0);
        CK(cudaStreamEndCapture(E.stream, &graph));
        CK(cudaGraphInstantiate(&sg.exec, graph, 0));
        CK(cudaGraphDestroy(graph));
        sg.launches = g_launches - l0;
        g_launches = l0;
        it = g_step_graphs.emplace(key, sg).first;
    }
    CK(cudaGraphLaunch(it->second.exec, E.stream));
    COUNT(it->second.launches);
    return false;
}

// bind controls, run nsteps for the listed members; returns per-member first error
// host-side wall time of the phases of the last step_members call (us): prologue (entry -> first launch), launch loop,
// wait for the GPU, epilogue; read with spdy_last_call_host_us (tools/e2e_breakdown.py)
static double g_host_us[4] = {0, 0, 0, 0};
static inline double now_us() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
static bool g_profile_intermediate = getenv("SPDY_PROFILE_INTERMEDIATE") && atoi(getenv("SPDY_PROFILE_INTERMEDIATE")) != 0;
// per-tile buffers of a driver call: error codes (device + pinned host) and the Ctx::outer_zero flags, grown together
static void ensure_err_capacity(int nt) {
    if (nt * TILE <= E.err_cap) return;
    if (E.d_err) CK(cudaFree(E.d_err)), CK(cudaFreeHost(E.h_err)), CK(cudaFree(E.d_outer));
    E.err_cap = nt * TILE;
    CK(cudaMalloc(&E.d_err, E.err_cap * sizeof(int)));
    CK(cudaMalloc(&E.d_outer, 2 * nt * sizeof(int)));
    CK(cudaMallocHost(&E.h_err, E.err_cap * sizeof(int)));
}
static const bool g_scan_outer = !(getenv("SPDY_SCAN_OUTER") && atoi(getenv("SPDY_SCAN_OUTER")) == 0);
static const bool g_lazy_diag = !(getenv("SPDY_LAZY_DIAG") && atoi(getenv("SPDY_LAZY_DIAG")) == 0);
static int step_members(const int64_t *hs, const int64_t *cs, int n, int nsteps, int *err_out, bool per_step_sync) {
    engine_init();
    const double t_entry = now_us();
    double t_wait = 0.0, t_launch = 0.0;
    int failed = 0;
    std::vector<int64_t> run;
    std::vector<int> idx;
    for (int i = 0; i < n; i++) {
        err_out[i] = 0;
        Member *m = member_of(hs[i]);
        Control *c = control_of(cs[i]);
        if (!m || !c || !m->initialized) {
            err_out[i] = -1;  // E_STATE_NOT_INITIALIZED (speedy.f90:41-44)
            failed++;
            continue;
        }
        const int ci = (int)cs[i] - 1, mi = (int)hs[i] - 1;
        if (m->bound_ctl != ci || c->bound_member != mi) {
            upload_control(*m, *c);
            m->bound_ctl = ci, c->bound_member = mi;
        }
        run.push_back(hs[i]);
        idx.push_back(i);
    }
    if (run.empty()) return failed;
    const int nt = prepare_members(run.data(), (int)run.size());
    ensure_err_capacity(nt);
    std::map<int, int> tile_pos;
    for (int t = 0; t < nt; t++) tile_pos[E.cached_tiles[t]] = t;
    std::vector<int> epos(run.size());  // position of each member's error code in the read-back buffer
    std::vector<Control *> ctl(run.size());
    for (size_t q = 0; q < run.size(); q++) {
        const Member *m = member_of(run[q]);
        epos[q] = tile_pos[m->tile] * TILE + m->lane;
        ctl[q] = control_of(cs[idx[q]]);
    }
    struct SavedDate {
        Datetime d;
        int month_idx;
    };
    std::vector<SavedDate> saved(per_step_sync ? run.size() : 0);
    CK(cudaMemsetAsync(E.d_err, 0, (size_t)nt * TILE * sizeof(int), E.stream));  // sticky within this call (k_diag_final)
    bool any_failed = false;
    CK(cudaEventRecord(E.ev0, E.stream));
    // multi-step call: look once at the coefficients outside the triangular truncation (k_scan_outer, dynamics.cu)
    E.outer_on = false;
    if (((nsteps >= 4 && !per_step_sync) || (P.on && g_profile_intermediate)) && g_scan_outer) {
        CK(cudaMemsetAsync(E.d_outer, 1, (size_t)2 * nt * sizeof(int), E.stream));  // bytes 0x01: non-zero = "nothing found that moves"
        for (int t0 = 0; t0 < nt; t0 += E.chunk_tiles)
            launch_scan_outer(E.stream, make_ctx(E.d_tiles + t0, E.d_masks + t0, std::min(E.chunk_tiles, nt - t0)), E.d_outer + 2 * t0);
        COUNT((nt + E.chunk_tiles - 1) / E.chunk_tiles);
        E.outer_on = true;
    }
    const double t_first = now_us();
    for (int s = 0; s < nsteps; s++) {
        const double t_s0 = now_us();
        bool any_daily = false;
        for (size_t q = 0; q < run.size(); q++) any_daily |= (member_of(run[q])->current_step % NSTEPS == 0);
        bool early = false;
        // outputs of the column physics that no kernel reads (fluxes, precipitation, rad_flux, rad_st4a: 39 of the ~160
        // doubles a column moves per step) are stored by the LAST step of a multi-step call only: nothing can observe the
        // values the intermediate steps would have left there (tests/test_ensemble_gpu.py::test_multistep_call_leaves_the_complete_state)
        E.diag_out = (per_step_sync || s == nsteps - 1 || !g_lazy_diag) ? 1 : 0;
        if (P.on && g_profile_intermediate) E.diag_out = 0;  // spdy_profile_step timing an intermediate step (bench.py)
        for (int t0 = 0; t0 < nt; t0 += E.chunk_tiles) {
            const bool last_chunk = t0 + E.chunk_tiles >= nt;
            early = run_chunk_step(t0, std::min(E.chunk_tiles, nt - t0), any_daily, (per_step_sync && last_chunk) ? nt : 0);
        }
        t_launch += now_us() - t_s0;
        for (size_t q = 0; q < run.size(); q++) member_of(run[q])->current_step += 1;
        const bool last = (s == nsteps - 1);
        const bool readback = per_step_sync || last || ((s + 1) % NSTEPS == 0);
        // host mirror of the calendar (the device copy is authoritative for the kernels).  In the per-step driver call it
        // is advanced while the GPU works on the step and taken back for the (rare) members whose check fails: the
        // reference does not advance the date of a failed step (speedy.f90:62-69)
        if (per_step_sync) {
            for (size_t q = 0; q < run.size(); q++) {
                saved[q] = SavedDate{ctl[q]->model, ctl[q]->month_idx};
                if (err_out[idx[q]] == 0) advance_host_date(*ctl[q]);
            }
        }
        if (readback) {
            // error codes: read back at most once a day in batched mode; the device keeps the first non-zero code of
            // every member and freezes a failed member for the rest of the call (k_diag_final)
            const double t_w0 = now_us();
            if (early) {
                CK(cudaEventSynchronize(E.ev_err));  // the coupler kernel of this step may still be running
            } else {
                CK(cudaMemcpyAsync(E.h_err, E.d_err, nt * TILE * sizeof(int), cudaMemcpyDeviceToHost, E.stream));
                CK(cudaStreamSynchronize(E.stream));
            }
            t_wait += now_us() - t_w0;
            for (size_t q = 0; q < run.size(); q++) {
                const int code = E.h_err[epos[q]];
                if (code != 0 && err_out[idx[q]] == 0) {
                    err_out[idx[q]] = code;
                    any_failed = true;
                    if (per_step_sync) ctl[q]->model = saved[q].d, ctl[q]->month_idx = saved[q].month_idx;
                }
            }
        }
        if (!per_step_sync) {
            for (size_t q = 0; q < run.size(); q++)
                if (err_out[idx[q]] == 0) advance_host_date(*ctl[q]);
        }
    }
    E.diag_out = 1, E.outer_on = false;
    CK(cudaEventRecord(E.ev1, E.stream));
    E.last_ms_pending = true;  // read on demand (spdy_last_elapsed_ms): the per-step call does not wait for the step's tail
    if (!per_step_sync) CK(cudaStreamSynchronize(E.stream));
    CK(cudaGetLastError());  // a rejected kernel launch is not reported by the synchronisation
    if (any_failed) {
        // the device masks of the failed members were cleared: upload them again with the next call; in a multi-step call
        // the host mirrors ran ahead of a member that stopped at its failing step -- take them from the device
        E.cached_handles.clear();
        if (!per_step_sync)
            for (size_t q = 0; q < run.size(); q++)
                if (err_out[idx[q]] != 0) {
                    Member *m = member_of(run[q]);
                    m->current_step = (int)get_slot_host(*m, SL_STEP);
                    Datetime &d = ctl[q]->model;
                    d.y = (int)get_slot_host(*m, SL_YEAR), d.mo = (int)get_slot_host(*m, SL_MONTH);
                    d.d = (int)get_slot_host(*m, SL_DAY), d.h = (int)get_slot_host(*m, SL_HOUR);
                    d.mi = (int)get_slot_host(*m, SL_MINUTE), ctl[q]->month_idx = (int)get_slot_host(*m, SL_MONTH_IDX);
                }
    }
    for (int i = 0; i < n; i++) failed += (err_out[i] != 0 && err_out[i] != -1) ? 1 : 0;
    const double t_end = now_us();
    g_host_us[0] = t_first - t_entry, g_host_us[1] = t_launch, g_host_us[2] = t_wait;
    g_host_us[3] = (t_end - t_first) - t_launch - t_wait;
    return failed;
}

static Ctx single_ctx(const Member &m) {
    int64_t h = (&m - E.members.data()) + 1;
    prepare_members(&h, 1);
    return make_ctx(E.d_tiles, E.d_masks, 1);
}

// initialize_state (initialization.f90:13-91) for one member
static int init_member(Member &m, Control &ctl) {
    engine_init();
    const ScratchLayout &L = E.L;
    m.current_step = 0;
    const int mi = (int)(&m - E.members.data()), ci = (int)(&ctl - E.controls.data());
    upload_control(m, ctl);
    m.bound_ctl = ci, ctl.bound_member = mi;
    Ctx c = single_ctx(m);
    // boundaries.f90:22-37: phi0 = g*orog ; phis0 = grid_filter(phi0)
    launch_init_grid(E.stream, c, 0, L.px);
    COUNT(1);
    run_forward_plain(c, {REF_SCR | L.px}, {REF_SCR | L.dpx});
    k_spec_trunc<<<dim3(NSPC / 4, 1), 128, 0, E.stream>>>(c, REF_SCR | L.dpx, 1);
    run_inverse_list(c, {InvDesc{REF_SCR | L.dpx, L.py, 1, 0}});
    launch_masked_copy(E.stream, c, REF_SCR | L.py, E.off[V_phis0], NG, 0, 1.0);
    COUNT(2);
    // prognostics.f90:55-112
    run_forward_plain(c, {E.off[V_phis0]}, {REF_SCR | L.dpx});
    launch_masked_copy(E.stream, c, REF_SCR | L.dpx, E.off[V_phis], NSP, 0, 1.0);
    launch_init_grid(E.stream, c, 1, L.px);
    COUNT(2);
    run_forward_plain(c, {REF_SCR | L.px}, {REF_SCR | L.dpx});
    launch_init_grid(E.stream, c, 2, L.px);
    COUNT(1);
    run_forward_plain(c, {REF_SCR | L.px}, {REF_SCR | L.dpy});
    launch_init_spec(E.stream, c, REF_SCR | L.dpx, REF_SCR | L.dpy);
    k_set_slot<<<1, 32, 0, E.stream>>>(c, SL_ERR, 0.0);
    launch_diag(E.stream, c, 1, E.L.diagp, 0);
    COUNT(3);
    const int err = (int)get_slot_host(m, SL_ERR);
    if (err != 0) return err;
    // coupler.f90:13-30, forcing.f90 (imode 0), time_stepping.f90:13-27
    launch_surface_init(E.stream, c);
    launch_couple(E.stream, c, 1);
    COUNT(2);
    run_forcing(c, 0);
    run_step_core(c, 1, 1, 0.5 * H_DELT, 0.0, 0);
    run_step_core(c, 1, 2, H_DELT, 0.0, 1);
    CK(cudaStreamSynchronize(E.stream));
    CK(cudaGetLastError());
    // initialization.f90:85-87: coordinates in default REAL
    for (int k = 0; k < KX; k++) m.lev[k] = (float)E.C.fsg[k];
    for (int k = 0; k < IX; k++) m.lon[k] = 3.75f * (float)k;
    for (int k = 0; k < IL; k++) m.lat[k] = (float)E.C.radang[k] * 90.0f / (float)0x1.921fb6p+0;
    m.initialized = true;
    set_slot_host(m, SL_INITIALIZED, 1.0);
    return 0;
}

static void s2g_ctx(const Ctx &c);
static void s2g_member(Member &m) {  // prognostics.f90:125-154
    Ctx c = single_ctx(m);
    s2g_ctx(c);
}
// uvspec + the 41 inverse transforms into scratch (no host synchronisation: the descriptor list is resident)
static void s2g_transforms(const Ctx &c) {
    const ScratchLayout &L = E.L;
    launch_uvspec(E.stream, c, E.off[V_vor], E.off[V_div], REF_SCR | L.ucos, REF_SCR | L.vcos, KX, 0);
    COUNT(1);
    run_inverse(c, E.d_inv_s2g, 5 * KX + 1);
}
static void s2g_ctx(const Ctx &c) {
    s2g_transforms(c);
    k_s2g_finish<<<dim3(NG / 4, c.ntiles), 128, 0, E.stream>>>(c, E.L);
    COUNT(1);
    CK(cudaStreamSynchronize(E.stream));
}

static void g2s_member(Member &m) {  // prognostics.f90:157-176
    const ScratchLayout &L = E.L;
    Ctx c = single_ctx(m);
    k_g2s_prepare<<<dim3(NG / 4, 1), 128, 0, E.stream>>>(c, L.px);
    COUNT(1);
    std::vector<FieldRef> su, du, sp, dp;
    for (int k = 0; k < KX; k++) su.push_back(E.off[V_u_grid] + (long long)k * NG), du.push_back(REF_SCR | (L.sfwd + (long long)(0 + k) * NSP));
    for (int k = 0; k < KX; k++) su.push_back(E.off[V_v_grid] + (long long)k * NG), du.push_back(REF_SCR | (L.sfwd + (long long)(8 + k) * NSP));
    run_forward_plain(c, su, du, FM_COS);
    // the Fourier scratch is reused: run the plain fields in a second pass
    for (int k = 0; k < KX; k++) sp.push_back(E.off[V_t_grid] + (long long)k * NG), dp.push_back(REF_SCR | (L.sfwd + (long long)(16 + k) * NSP));
    for (int k = 0; k < KX; k++) sp.push_back(E.off[V_q_grid] + (long long)k * NG), dp.push_back(REF_SCR | (L.sfwd + (long long)(24 + k) * NSP));
    for (int k = 0; k < KX; k++) sp.push_back(E.off[V_phi_grid] + (long long)k * NG), dp.push_back(REF_SCR | (L.sfwd + (long long)(32 + k) * NSP));
    sp.push_back(REF_SCR | L.px), dp.push_back(REF_SCR | (L.sfwd + 40ll * NSP));
    run_forward_plain(c, sp, dp, FM_PLAIN);
    k_g2s_finish<<<dim3(NSPC / 4, 1), 128, 0, E.stream>>>(c, L);
    COUNT(1);
    CK(cudaStreamSynchronize(E.stream));
}

static void filter_member(Member &m) {  // prognostics.f90:180-219
    const ScratchLayout &L = E.L;
    Ctx c = single_ctx(m);
    const int ids[5] = {V_u_grid, V_v_grid, V_t_grid, V_q_grid, V_phi_grid};
    std::vector<FieldRef> src, dst;
    std::vector<InvDesc> inv;
    int f = 0;
    for (int v = 0; v < 5; v++)
        for (int k = 0; k < KX; k++, f++) {
            src.push_back(E.off[ids[v]] + (long long)k * NG);
            dst.push_back(REF_SCR | (L.sfwd + (long long)f * NSP));
            inv.push_back(InvDesc{REF_SCR | (L.sfwd + (long long)f * NSP), L.ug + (long long)f * NG, 1, 0});
        }
    src.push_back(E.off[V_ps_grid]), dst.push_back(REF_SCR | (L.sfwd + (long long)f * NSP));
    inv.push_back(InvDesc{REF_SCR | (L.sfwd + (long long)f * NSP), L.ug + (long long)f * NG, 1, 0});
    f++;
    run_forward_plain(c, src, dst);
    k_spec_trunc<<<dim3(NSPC / 4, 1), 128, 0, E.stream>>>(c, REF_SCR | L.sfwd, f);
    COUNT(1);
    run_inverse_list(c, inv);
    int g = 0;
    for (int v = 0; v < 5; v++, g += KX)
        launch_masked_copy(E.stream, c, REF_SCR | (L.ug + (long long)g * NG), E.off[ids[v]], NG * KX, 0, 1.0), COUNT(1);
    launch_masked_copy(E.stream, c, REF_SCR | (L.ug + (long long)g * NG), E.off[V_ps_grid], NG, 0, 1.0);
    COUNT(1);
    CK(cudaStreamSynchronize(E.stream));
}

// ---- registry accessors -------------------------------------------------------------------------------------
static long long var_elems(const Member &m, int v) {
    if (v == V_sst_anom) return m.n_months < 0 ? 0 : (long long)(m.n_months + 2) * NG;
    return E.nelem[v];
}
static int scalar_slot(int v) {
    switch (v) {
        case V_current_step: return SL_STEP;
        case V_increase_co2: return SL_INCCO2;
        case V_compute_shortwave: return SL_SW;
        case V_air_absortivity_co2: return SL_CO2;
        case V_land_coupling_flag: return SL_LANDCPL;
        case V_sst_anomaly_coupling_flag: return SL_SSTACPL;
        case V_ablco2_ref: return SL_CO2REF;
    }
    return -1;
}
static void xfer_array(const Member &m, int v, double *host, bool to_device) {
    const long long n = var_elems(m, v);
    double *arena = (v == V_sst_anom) ? E.sst : E.st;
    const long long te = (v == V_sst_anom) ? E.sst_elems : E.st_elems, off = (v == V_sst_anom) ? 0 : E.off[v];
    for (long long o = 0; o < n; o += (long long)E.stage_elems) {
        const long long cn = std::min<long long>(E.stage_elems, n - o);
        const int nb = (int)((cn + 255) / 256);
        if (to_device) {
            memcpy(E.h_stage, host + o, cn * sizeof(double));
            CK(cudaMemcpyAsync(E.d_stage, E.h_stage, cn * sizeof(double), cudaMemcpyHostToDevice, E.stream));
            k_scatter<<<nb, 256, 0, E.stream>>>(arena, te, m.tile, m.lane, off + o, cn, E.d_stage);
            CK(cudaStreamSynchronize(E.stream));
        } else {
            k_gather<<<nb, 256, 0, E.stream>>>(arena, te, m.tile, m.lane, off + o, cn, E.d_stage);
            CK(cudaMemcpyAsync(E.h_stage, E.d_stage, cn * sizeof(double), cudaMemcpyDeviceToHost, E.stream));
            CK(cudaStreamSynchronize(E.stream));
            memcpy(host + o, E.h_stage, cn * sizeof(double));
        }
    }
}

}  // namespace spdy

// =============================================================================================== C ABI
using namespace spdy;

extern "C" {

int spdy_set_device(int ordinal) {
    API_LOCK;
    if (E.ready) return -1;
    E.device = ordinal;
    return 0;
}
int spdy_device_count(void) {
    API_LOCK;
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
int spdy_synchronize(void) {
    API_LOCK;
    if (E.ready) CK(cudaStreamSynchronize(E.stream));
    return 0;
}
void spdy_last_call_host_us(double *out4) {
    API_LOCK;
    for (int i = 0; i < 4; i++) out4[i] = g_host_us[i];
}
float spdy_last_elapsed_ms(void) {
    API_LOCK;
    if (E.last_ms_pending) {
        CK(cudaEventSynchronize(E.ev1));
        CK(cudaEventElapsedTime(&E.last_ms, E.ev0, E.ev1));
        E.last_ms_pending = false;
    }
    return E.last_ms;
}
long long spdy_kernel_launches(void) { API_LOCK; return g_launches; }

int spdy_get_model_datetime(int64_t h, int *out) {
    API_LOCK;
    Member *m = member_of(h);
    if (!m) return -1;
    const int s[5] = {SL_YEAR, SL_MONTH, SL_DAY, SL_HOUR, SL_MINUTE};
    for (int i = 0; i < 5; i++) out[i] = (int)get_slot_host(*m, s[i]);
    return 0;
}

// SPPT switch (a compile-time constant of the reference: params.f90:44 `sppt_on`); applies to every member stepped from
// now on.  The pattern generator is keyed by (seed, member slot, calls so far), see sppt.cu.
int spdy_set_sppt(int on, unsigned long long seed) {
    API_LOCK;
    engine_init();
    const bool want = on != 0;
    E.sppt_seed = seed;
    if (want != E.sppt_on) {  // the scratch arena carries the SPPT fields only while the switch is on
        CK(cudaStreamSynchronize(E.stream));
        E.sppt_on = want;
        E.L.total = want ? E.L.total_sppt : E.L.total_base;
        if (E.scr) CK(cudaFree(E.scr));
        E.scr = nullptr, E.scr_tiles = 0;
        drop_step_graphs();
        E.cached_handles.clear();
    }
    return 0;
}
// the member's spectral AR(1) pattern ((31,32,8) complex) and the clipped grid-point pattern of its last single-member
// step ((96,48,8)); calls = patterns generated so far
int spdy_debug_get_sppt(int64_t h, double *spec, double *grid, long long *calls) {
    API_LOCK;
    Member *m = member_of(h);
    if (!m || !E.sppt_on || !E.scr) return -1;
    auto down = [&](double *dst, double *arena, long long te, int tile, long long off, long long n) {
        for (long long o = 0; o < n; o += (long long)E.stage_elems) {
            const long long cn = std::min<long long>(E.stage_elems, n - o);
            k_gather<<<(int)((cn + 255) / 256), 256, 0, E.stream>>>(arena, te, tile, m->lane, off + o, cn, E.d_stage);
            CK(cudaMemcpyAsync(E.h_stage, E.d_stage, cn * 8, cudaMemcpyDeviceToHost, E.stream));
            CK(cudaStreamSynchronize(E.stream));
            memcpy(dst + o, E.h_stage, cn * 8);
        }
    };
    down(spec, E.st, E.st_elems, m->tile, E.off_sppt, (long long)NSP * KX);
    down(grid, E.scr, E.L.total, 0, E.L.spptg, (long long)NG * KX);
    *calls = (long long)get_slot_host(*m, SL_SPPT_CALLS);
    return 0;
}

int spdy_reserve(int n_members) {
    API_LOCK;
    engine_init();
    ensure_state_tiles((n_members + TILE - 1) / TILE);
    return 0;
}

int64_t spdy_modelstate_init(void) {
    API_LOCK;
    engine_init();
    int idx;
    if (!E.free_members.empty()) {
        idx = E.free_members.back();
        E.free_members.pop_back();
    } else {
        idx = (int)E.members.size();
        E.members.push_back(Member());
    }
    Member &m = E.members[idx];
    memset(&m, 0, sizeof(m));
    m.tile = idx / TILE, m.lane = idx % TILE, m.alive = true, m.initialized = false, m.n_months = -1, m.bound_ctl = -1;
    ensure_state_tiles(m.tile + 1);
    // zero this member's lane (model_state.f90:358-...: every array allocated and zeroed) and set defaults
    {
        k_fill_lane<<<(int)((E.st_elems + 255) / 256), 256, 0, E.stream>>>(E.st, E.st_elems, m.tile, m.lane, E.st_elems, 0.0);
        COUNT(1);
        CK(cudaStreamSynchronize(E.stream));
        set_slot_host(m, SL_CO2, 6.0);        // registry defaults (registry/model_state_def.py:305-325,377-383,412-417)
        set_slot_host(m, SL_SW, 1.0);
        set_slot_host(m, SL_LANDCPL, 1.0);
        set_slot_host(m, SL_SSTACPL, 1.0);
    }
    return (int64_t)idx + 1;
}

void spdy_modelstate_init_sst_anom(int64_t h, int n_months) {
    API_LOCK;
    Member *m = member_of(h);
    if (!m || n_months < 0) return;
    ensure_sst_months(n_months + 2);
    m->n_months = n_months;
    set_slot_host(*m, SL_NMONTHS, (double)n_months);
    std::vector<double> z((size_t)(n_months + 2) * NG, 0.0);
    xfer_array(*m, V_sst_anom, z.data(), true);
}

void spdy_modelstate_close(int64_t h) {
    API_LOCK;
    Member *m = member_of(h);
    if (!m) return;
    m->alive = false;
    if (m->bound_ctl >= 0 && m->bound_ctl < (int)E.controls.size()) E.controls[m->bound_ctl].bound_member = -1;
    E.free_members.push_back((int)h - 1);
    E.cached_handles.clear();
}

int64_t spdy_create_datetime(int y, int mo, int d, int h, int mi) {
    API_LOCK;
    if (!E.free_dates.empty()) {  // closed containers are recycled (the reference deallocates them, .j2:204-210)
        const int idx = E.free_dates.back();
        E.free_dates.pop_back();
        E.dates[idx] = Datetime{y, mo, d, h, mi, true};
        return (int64_t)idx + 1;
    }
    E.dates.push_back(Datetime{y, mo, d, h, mi, true});
    return (int64_t)E.dates.size();
}
void spdy_get_datetime(int64_t h, int *o) {
    API_LOCK;
    if (h < 1 || h > (int64_t)E.dates.size()) return;
    const Datetime &d = E.dates[h - 1];
    o[0] = d.y, o[1] = d.mo, o[2] = d.d, o[3] = d.h, o[4] = d.mi;
}
void spdy_close_datetime(int64_t h) {
    API_LOCK;
    if (h >= 1 && h <= (int64_t)E.dates.size() && E.dates[h - 1].alive) {
        E.dates[h - 1].alive = false;
        E.free_dates.push_back((int)h - 1);
    }
}
int64_t spdy_controlparams_init(int64_t s, int64_t e) {
    API_LOCK;
    if (s < 1 || s > (int64_t)E.dates.size() || e < 1 || e > (int64_t)E.dates.size()) return 0;
    Control c;
    c.start = E.dates[s - 1], c.end = E.dates[e - 1], c.model = c.start, c.month_idx = 1, c.alive = true, c.bound_member = -1;
    if (!E.free_controls.empty()) {
        const int idx = E.free_controls.back();
        E.free_controls.pop_back();
        E.controls[idx] = c;
        return (int64_t)idx + 1;
    }
    E.controls.push_back(c);
    return (int64_t)E.controls.size();
}
void spdy_controlparams_close(int64_t h) {
    API_LOCK;
    Control *c = control_of(h);
    if (!c) return;
    c->alive = false;
    if (c->bound_member >= 0 && c->bound_member < (int)E.members.size()) E.members[c->bound_member].bound_ctl = -1;
    E.free_controls.push_back((int)h - 1);
}

int spdy_init(int64_t sh, int64_t ch) {
    API_LOCK;
    Member *m = member_of(sh);
    Control *c = control_of(ch);
    if (!m || !c) return -1;
    return init_member(*m, *c);
}
int spdy_step(int64_t sh, int64_t ch) {
    API_LOCK;
    int err = 0;
    step_members(&sh, &ch, 1, 1, &err, true);
    return err;
}
void spdy_parallel_step(const int64_t *s, const int64_t *c, int *err, int n) { API_LOCK; step_members(s, c, n, 1, err, true); }
int spdy_run_steps(const int64_t *s, const int64_t *c, int n, int nsteps, int *err) {
    API_LOCK;
    return step_members(s, c, n, nsteps, err, false);
}
int spdy_check(int64_t h) {
    API_LOCK;
    Member *m = member_of(h);
    if (!m) return -1;
    Ctx c = single_ctx(*m);
    k_set_slot<<<1, 32, 0, E.stream>>>(c, SL_ERR, 0.0);
    launch_diag(E.stream, c, 1, E.L.diagp, 0);
    COUNT(2);
    return (int)get_slot_host(*m, SL_ERR);
}
void spdy_transform_spectral2grid(int64_t h) {
    API_LOCK;
    Member *m = member_of(h);
    if (m) s2g_member(*m);
}
void spdy_transform_grid2spectral(int64_t h) {
    API_LOCK;
    Member *m = member_of(h);
    if (m) g2s_member(*m);
}
void spdy_apply_grid_filter(int64_t h) {
    API_LOCK;
    Member *m = member_of(h);
    if (m) filter_member(*m);
}

int spdy_shape(int64_t h, int v, int *dims, int *ndim) {
    API_LOCK;
    Member *m = member_of(h);
    if (!m || v < 0 || v >= SPDY_NVARS) return -1;
    const spdy_vardef &d = SPDY_VARDEFS[v];
    *ndim = d.ndim;
    for (int q = 0; q < d.ndim; q++) dims[q] = d.dims[q] < 0 ? (m->n_months < 0 ? 0 : m->n_months + 2) : d.dims[q];
    if (v == V_sst_anom && m->n_months < 0) dims[0] = dims[1] = dims[2] = 0;  // "zeros if not allocated" (.j2:306-318)
    return 0;
}
int spdy_get(int64_t h, int v, void *dst, size_t bytes) {
    API_LOCK;
    Member *m = member_of(h);
    if (!m || v < 0 || v >= SPDY_NVARS) return -1;
    const spdy_vardef &d = SPDY_VARDEFS[v];
    if (d.ndim == 0) {
        const double x = (v == V_current_step) ? (double)m->current_step : get_slot_host(*m, scalar_slot(v));
        if (d.kind == SPDY_F8) {
            if (bytes != 8) return -2;
            memcpy(dst, &x, 8);
        } else {
            if (bytes != 4) return -2;
            const int i = (int)x;
            memcpy(dst, &i, 4);
        }
        return 0;
    }
    if (d.kind == SPDY_F4) {
        const float *f = (v == V_lon) ? m->lon : (v == V_lat) ? m->lat : m->lev;
        const size_t n = (v == V_lon) ? IX : (v == V_lat) ? IL : KX;
        if (bytes != n * 4) return -2;
        memcpy(dst, f, bytes);
        return 0;
    }
    if (bytes != (size_t)var_elems(*m, v) * 8) return -2;
    if (bytes) xfer_array(*m, v, (double *)dst, false);
    return 0;
}
int spdy_set(int64_t h, int v, const void *src, size_t bytes) {
    API_LOCK;
    Member *m = member_of(h);
    if (!m || v < 0 || v >= SPDY_NVARS) return -1;
    const spdy_vardef &d = SPDY_VARDEFS[v];
    if (d.ndim == 0) {
        double x;
        if (d.kind == SPDY_F8) {
            if (bytes != 8) return -2;
            memcpy(&x, src, 8);
        } else {
            if (bytes != 4) return -2;
            int i;
            memcpy(&i, src, 4);
            x = (d.kind == SPDY_B1) ? (i != 0 ? 1.0 : 0.0) : (double)i;
        }
        if (v == V_current_step) m->current_step = (int)x;
        set_slot_host(*m, scalar_slot(v), x);
        set_slot_host(*m, SL_CPLDIRTY, 1.0);
        return 0;
    }
    if (d.kind == SPDY_F4) {
        float *f = (v == V_lon) ? m->lon : (v == V_lat) ? m->lat : m->lev;
        const size_t n = (v == V_lon) ? IX : (v == V_lat) ? IL : KX;
        if (bytes != n * 4) return -2;
        memcpy(f, src, bytes);
        return 0;
    }
    if (bytes != (size_t)var_elems(*m, v) * 8) return -2;
    if (bytes) xfer_array(*m, v, (double *)src, true);
    set_slot_host(*m, SL_CPLDIRTY, 1.0);  // the coupler re-derives its interpolated climatology on the next step
    return 0;
}

int spdy_debug_get_corh(int64_t h, double *tcorh, double *qcorh) {
    API_LOCK;
    Member *m = member_of(h);
    if (!m) return -1;
    const int nb = (NSP + 255) / 256;
    k_gather<<<nb, 256, 0, E.stream>>>(E.st, E.st_elems, m->tile, m->lane, E.off_tcorh, NSP, E.d_stage);
    k_gather<<<nb, 256, 0, E.stream>>>(E.st, E.st_elems, m->tile, m->lane, E.off_qcorh, NSP, E.d_stage + NSP);
    CK(cudaMemcpyAsync(E.h_stage, E.d_stage, 2 * NSP * sizeof(double), cudaMemcpyDeviceToHost, E.stream));
    CK(cudaStreamSynchronize(E.stream));
    memcpy(tcorh, E.h_stage, NSP * 8);
    memcpy(qcorh, E.h_stage + NSP, NSP * 8);
    return 0;
}

int spdy_debug_raw_step(int64_t h, int j1, int j2, int dt_kind) {
    API_LOCK;
    Member *m = member_of(h);
    if (!m) return -1;
    Ctx c = single_ctx(*m);
    const double dts[3] = {0.5 * H_DELT, H_DELT, 2.0 * H_DELT};
    run_step_core(c, j1, j2, dts[dt_kind], j1 == 1 ? 0.0 : FL(0.05), dt_kind);
    CK(cudaStreamSynchronize(E.stream));
    return 0;
}

// tendencies of one member as returned by get_tendencies(state, ..., j2) (tendencies.f90:11-39): vordt, divdt, tdt
// (31,32,8) complex each, psdt (31,32), trdt (31,32,8); the prognostic state is not advanced (physics diagnostics
// are updated exactly as a model step would)
int spdy_debug_tendencies_stage(int64_t h, int j2, int stage, double *vordt, double *divdt, double *tdt, double *psdt,
                                double *trdt);
int spdy_debug_tendencies(int64_t h, int j2, double *vordt, double *divdt, double *tdt, double *psdt, double *trdt) {
    return spdy_debug_tendencies_stage(h, j2, 2, vordt, divdt, tdt, psdt, trdt);
}
// stage 1: divdt, tdt, psdt as they ENTER implicit_terms (time_stepping.f90:71-75), i.e. after get_grid_point_tendencies and
// get_spectral_tendencies; stage 2: as returned by get_tendencies (after the semi-implicit correction)
int spdy_debug_tendencies_stage(int64_t h, int j2, int stage, double *vordt, double *divdt, double *tdt, double *psdt,
                                double *trdt) {
    API_LOCK;
    Member *m = member_of(h);
    if (!m || stage < 1 || stage > 2) return -1;
    Ctx c = single_ctx(*m);
    const long long dump = E.L.four;
    g_dump_stage = stage;
    run_step_core(c, 2, j2, 2.0 * H_DELT, FL(0.05), 2, dump);
    g_dump_stage = 2;
    CK(cudaStreamSynchronize(E.stream));
    auto down = [&](double *dst, long long off, long long n) {
        for (long long o = 0; o < n; o += (long long)E.stage_elems) {
            const long long cn = std::min<long long>(E.stage_elems, n - o);
            k_gather<<<(int)((cn + 255) / 256), 256, 0, E.stream>>>(E.scr, E.L.total, 0, m->lane, off + o, cn, E.d_stage);
            CK(cudaMemcpyAsync(E.h_stage, E.d_stage, cn * 8, cudaMemcpyDeviceToHost, E.stream));
            CK(cudaStreamSynchronize(E.stream));
            memcpy(dst + o, E.h_stage, cn * 8);
        }
    };
    down(vordt, dump, (long long)NSP * KX);
    down(divdt, dump + 8ll * NSP, (long long)NSP * KX);
    down(tdt, dump + 16ll * NSP, (long long)NSP * KX);
    down(psdt, dump + 24ll * NSP, NSP);
    down(trdt, dump + 25ll * NSP, (long long)NSP * KX);
    return 0;
}

int spdy_debug_physics(int64_t h, const double *ug8, const double *vg8, const double *tg, const double *qg,
                       const double *phig, const double *pslg, double *utend8, double *vtend8, double *ttend,
                       double *qtend, int *dbg) {
    API_LOCK;
    Member *m = member_of(h);
    if (!m) return -1;
    const ScratchLayout &L = E.L;
    Ctx c = single_ctx(*m);
    auto up = [&](const double *src, long long off, long long n) {
        memcpy(E.h_stage, src, n * 8);
        CK(cudaMemcpyAsync(E.d_stage, E.h_stage, n * 8, cudaMemcpyHostToDevice, E.stream));
        k_scatter<<<(int)((n + 255) / 256), 256, 0, E.stream>>>(E.scr, E.L.total, 0, m->lane, off, n, E.d_stage);
        CK(cudaStreamSynchronize(E.stream));
    };
    auto down = [&](double *dst, long long off, long long n) {
        k_gather<<<(int)((n + 255) / 256), 256, 0, E.stream>>>(E.scr, E.L.total, 0, m->lane, off, n, E.d_stage);
        CK(cudaMemcpyAsync(E.h_stage, E.d_stage, n * 8, cudaMemcpyDeviceToHost, E.stream));
        CK(cudaStreamSynchronize(E.stream));
        memcpy(dst, E.h_stage, n * 8);
    };
    const long long G3 = (long long)NG * KX;
    up(ug8, L.pug8, NG), up(vg8, L.pvg8, NG), up(tg, L.ptg, G3), up(qg, L.pqg, G3), up(phig, L.pphig, G3), up(pslg, L.pslg, NG);
    up(utend8, L.utend + 7ll * NG, NG), up(vtend8, L.vtend + 7ll * NG, NG), up(ttend, L.ttend, G3), up(qtend, L.trtend, G3);
    int *d_dbg = nullptr;
    CK(cudaMalloc(&d_dbg, 3 * NG * TILE * sizeof(int)));
    launch_physics(E.stream, c, L, d_dbg);
    COUNT(1);
    CK(cudaStreamSynchronize(E.stream));
    down(utend8, L.utend + 7ll * NG, NG), down(vtend8, L.vtend + 7ll * NG, NG), down(ttend, L.ttend, G3), down(qtend, L.trtend, G3);
    std::vector<int> hd((size_t)3 * NG * TILE);
    CK(cudaMemcpy(hd.data(), d_dbg, hd.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int a = 0; a < 3; a++)
        for (int q = 0; q < NG; q++) dbg[a * NG + q] = hd[((size_t)a * NG + q) * TILE + m->lane];
    CK(cudaFree(d_dbg));
    return 0;
}

int spdy_table(const char *name, double *dst, int cap) {
    API_LOCK;
    // host-only: the tables are built with glibc on the CPU and need no device (tables.cu)
    static ConstTables *hc = nullptr;
    static GlobTables *hg = nullptr;
    if (!hc) {
        hc = new ConstTables();
        hg = new GlobTables();
        build_tables(*hc, *hg);
    }
    const std::string n(name);
    const ConstTables &C = *hc;
    const GlobTables &G = *hg;
    std::vector<double> o;
    auto put = [&](const double *a, int cnt) { o.assign(a, a + cnt); };
    if (n == "hsg") put(C.hsg, KX + 1);
    else if (n == "dhs") put(C.dhs, KX);
    else if (n == "fsg") put(C.fsg, KX);
    else if (n == "dhsr") put(C.dhsr, KX);
    else if (n == "fsgr") put(C.fsgr, KX);
    else if (n == "radang") put(C.radang, IL);
    else if (n == "coriol") put(C.coriol, IL);
    else if (n == "sia") put(C.sia, IL);
    else if (n == "coa") put(C.coa, IL);
    else if (n == "cosgr") put(C.cosgr, IL);
    else if (n == "cosgr2") put(C.cosgr2, IL);
    else if (n == "sigl") put(C.sigl, KX);
    else if (n == "sigh") put(C.sigh, KX + 1);
    else if (n == "grdsig") put(C.grdsig, KX);
    else if (n == "grdscp") put(C.grdscp, KX);
    else if (n == "wvi") { for (int cc = 0; cc < 2; cc++) for (int k = 0; k < KX; k++) o.push_back(C.wvi[k][cc]); }
    else if (n == "wt") put(C.wt, IY);
    else if (n == "wa") put(C.wa, IX);
    else if (n == "tcorv") put(C.tcorv, KX);
    else if (n == "qcorv") put(C.qcorv, KX);
    else if (n == "tref") put(C.tref, KX);
    else if (n == "tref2") put(C.tref2, KX);
    else if (n == "tref3") put(C.tref3, KX);
    else if (n == "xgeop1") put(C.xgeop1, KX);
    else if (n == "xgeop2") put(C.xgeop2, KX);
    else if (n == "cpol") {  // expand to the reference's (2*mx, nx, iy) layout
        o.resize((size_t)M2 * NX * IY);
        for (int j = 0; j < IY; j++) for (int nn = 0; nn < NX; nn++) for (int m = 0; m < MX; m++) {
            const double v = G.cpol[(m * NX + nn) * IY + j];
            o[(2 * m) + (size_t)M2 * (nn + (size_t)NX * j)] = v, o[(2 * m + 1) + (size_t)M2 * (nn + (size_t)NX * j)] = v;
        }
    }
    else if (n == "el2") put(G.el2, NSPC);
    else if (n == "elm2") put(G.elm2, NSPC);
    else if (n == "trfilt") put(G.trfilt, NSPC);
    else if (n == "gradx") put(G.gradx, MX);
    else if (n == "gradym") put(G.gradym, NSPC);
    else if (n == "gradyp") put(G.gradyp, NSPC);
    else if (n == "uvdx") put(G.uvdx, NSPC);
    else if (n == "uvdym") put(G.uvdym, NSPC);
    else if (n == "uvdyp") put(G.uvdyp, NSPC);
    else if (n == "vddym") put(G.vddym, NSPC);
    else if (n == "vddyp") put(G.vddyp, NSPC);
    else if (n == "dmp") put(G.dmp, NSPC);
    else if (n == "dmpd") put(G.dmpd, NSPC);
    else if (n == "dmps") put(G.dmps, NSPC);
    else if (n == "fband") put(G.fband, 301 * 4);
    else if (n.size() > 2 && n[n.size() - 2] == '@') {  // implicit tables: "<name>@<0|1|2>" (dt/2, dt, 2dt)
        const ImplTables &I = G.impl[n[n.size() - 1] - '0'];
        const std::string b = n.substr(0, n.size() - 2);
        if (b == "dmp1") put(I.dmp1, NSPC);
        else if (b == "dmp1d") put(I.dmp1d, NSPC);
        else if (b == "dmp1s") put(I.dmp1s, NSPC);
        else if (b == "elz") put(I.elz, NSPC);
        else if (b == "xc") put(I.xc, KX * KX);
        else if (b == "xd") put(I.xd, KX * KX);
        else if (b == "xj") put(I.xj, KX * KX * 64);
        else if (b == "dhsx") put(I.dhsx, KX);
        else return -1;
    } else return -1;
    if ((int)o.size() > cap) return -(int)o.size();
    memcpy(dst, o.data(), o.size() * 8);
    return (int)o.size();
}

}  // extern "C"

#include "batch.cu"
#include "ensemble.cu"

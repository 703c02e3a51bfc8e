// speedy-b200: ensemble-scale services on top of the per-member driver surface (no reference counterpart except where
// cited): bulk getters and checks for a member list, the once-a-day ensemble mean / spread of the six default outputs
// produced in the epilogue of transform_spectral2grid, and the multi-GPU side of it -- one process per GPU, members
// sharded across ranks with NO communication inside a time step (speedy_driver.f90.j2:58-79: members are independent),
// one ncclAllReduce(sum, double) over NVLink for the diagnostics.  NCCL is loaded with dlopen so that the
// library has no link-time dependency on it: single-GPU use never touches it.
#include <dlfcn.h>
#include <nccl.h>

namespace spdy {

// ---------------------------------------------------------------------------------------------- NCCL binding
struct Nccl {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double *d_small = nullptr;  // 64 doubles for barriers / small reductions
};
static Nccl NC;

static bool nccl_load() {
    if (NC.h) return true;
    const char *names[] = {getenv("SPDY_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n) continue;
        NC.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (NC.h) break;
    }
    if (!NC.h) {
        fprintf(stderr, "speedy-b200: cannot load libnccl.so.2 (%s); multi-GPU runs need NCCL\n", dlerror());
        return false;
    }
#define NC_SYM(field, name)                                                        \
    *(void **)(&NC.field) = dlsym(NC.h, name);                                     \
    if (!NC.field) {                                                               \
        fprintf(stderr, "speedy-b200: symbol %s missing in libnccl\n", name);      \
        return false;                                                              \
    }
    NC_SYM(GetUniqueId, "ncclGetUniqueId");
    NC_SYM(CommInitRank, "ncclCommInitRank");
    NC_SYM(CommDestroy, "ncclCommDestroy");
    NC_SYM(AllReduce, "ncclAllReduce");
    NC_SYM(GetErrorString, "ncclGetErrorString");
#undef NC_SYM
    return true;
}
static void nck(ncclResult_t r, const char *what) {
    if (r != ncclSuccess) {
        fprintf(stderr, "speedy-b200: NCCL error %s in %s\n", NC.GetErrorString ? NC.GetErrorString(r) : "?", what);
        abort();
    }
}

// ---------------------------------------------------------------------------------------------- kernels
// The six default outputs (pyspeedy/__init__.py: DEFAULT_OUTPUT_VARS) as one vector of ENS_NTOT grid values per member
constexpr int ENS_NV = 6;
constexpr long long ENS_NTOT = 5ll * NG * KX + NG;  // 188,928
__host__ __device__ constexpr long long ens_voff(int v) { return (long long)v * NG * KX; }

// transform_spectral2grid epilogue (prognostics.f90:141-150: unit conversions into the *_grid state variables, masked)
// fused with the ensemble partial sums: one warp per grid point, lanes = members, looping over the ENS_TG tiles of its
// group with the tile loads of a (variable, level) issued together; sum and sum of squares over tiles are kept per
// lane and reduced over the lanes once per (variable, level) -- the diagnostic costs no second pass over the 41 output
// fields of every member.  part[(group * 2 + {0: sum, 1: sum of squares}) * ENS_NTOT + element]; deterministic.
__global__ void __launch_bounds__(128) k_s2g_finish_stats(const Ctx c, const ScratchLayout L, double *__restrict__ part,
                                                          const int group0) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5), g = blockIdx.y;
    const size_t e = (size_t)q * TILE, lev = (size_t)NG * TILE;
    bool act[ENS_TG];
#pragma unroll
    for (int i = 0; i < ENS_TG; i++) {
        const int t = g * ENS_TG + i;
        act[i] = t < c.ntiles && lane_active(c, t, lane);
    }
    double *p1 = part + ((size_t)(group0 + g) * 2) * ENS_NTOT, *p2 = p1 + ENS_NTOT;
    const long long src[5] = {L.ug, L.vg, L.tg, L.trg, L.pphig};
    const int dst[5] = {V_u_grid, V_v_grid, V_t_grid, V_q_grid, V_phi_grid};
#pragma unroll 1
    for (int v = 0; v < ENS_NV; v++) {
        const int nlev = v < 5 ? KX : 1;
#pragma unroll 1
        for (int k = 0; k < nlev; k++) {
            double x[ENS_TG];
#pragma unroll
            for (int i = 0; i < ENS_TG; i++)
                if (act[i]) x[i] = *(scp(c, g * ENS_TG + i, v < 5 ? src[v] : L.pslg, lane) + e + k * lev);
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int i = 0; i < ENS_TG; i++)
                if (act[i]) {
                    double y = x[i];
                    if (v == 3) y = y * FL(1.0e-3);
                    else if (v == 4) y = y / FL(9.81);
                    else if (v == 5) y = FL(1.e+5) * exp(y);
                    *(stp(c, g * ENS_TG + i, c.off[v < 5 ? dst[v] : V_ps_grid], lane) + e + k * lev) = y;
                    s1 += y, s2 += y * y;
                }
            for (int o = 16; o; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o), s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            if (lane == 0) p1[ens_voff(v) + (size_t)k * NG + q] = s1, p2[ens_voff(v) + (size_t)k * NG + q] = s2;
        }
    }
}
__global__ void __launch_bounds__(256) k_mean_spread(double *__restrict__ s, long long n, double inv_n) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= n) return;
    const double mean = s[i] * inv_n, var = s[n + i] * inv_n - mean * mean;
    s[i] = mean, s[n + i] = sqrt(var > 0.0 ? var : 0.0);
}

// members of a list gathered member-major: out[i][e] = variable element e of list member i (optionally as float32, the
// cast pyspeedy/speedy.py:443 does on the host); 32 x 32 transpose through shared memory so that both the tile rows
// and the per-member output runs are coalesced
template <typename T>
__global__ void __launch_bounds__(256) k_gather_members(const double *__restrict__ arena, long long tile_elems, long long off,
                                                        long long nelem, const int2 *__restrict__ who, int n, int i0,
                                                        T *__restrict__ out) {
    __shared__ double s[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long e0 = blockIdx.x * 32ll;
    const int ib = i0 + blockIdx.y * 32;  // first list member of this block
    const int i = ib + tx;
    int2 w = make_int2(0, 0);
    if (i < n) w = who[i];
    for (int r = ty; r < 32; r += 8) {
        const long long e = e0 + r;
        s[r][tx] = (i < n && e < nelem) ? arena[((long long)w.x * tile_elems + off + e) * TILE + w.y] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {  // r = member within the block, tx = element
        const int im = ib + r;
        const long long e = e0 + tx;
        if (im < n && e < nelem) out[(size_t)(im - i0) * nelem + e] = (T)s[tx][r];
    }
}
__global__ void k_get_err(const Ctx c, int *out) {
    const int lane = threadIdx.x, t = blockIdx.x;
    out[t * TILE + lane] = lane_active(c, t, lane) ? (int)slot(c, t, lane, SL_ERR) : 0;
}

struct EnsBuffers {
    double *d_part = nullptr, *d_stats = nullptr, *h_stats = nullptr;
    size_t part_cap = 0;
    void *d_stage = nullptr;
    size_t stage_cap = 0;
    int2 *d_who = nullptr;
    int who_cap = 0;
    char *h_pin[2] = {nullptr, nullptr};  // pinned bounce buffers of the bulk getter
    cudaEvent_t ev_pin[2];
};
static EnsBuffers EB;
constexpr size_t ENS_PIN_CHUNK = (size_t)8 << 20;

// device -> pageable host memory through two pinned bounce buffers: the copy into pinned memory runs at PCIe speed and
// overlaps the host memcpy of the previous chunk (a direct cudaMemcpy into a numpy array runs at a third of that)
static void d2h_bounced(void *dst, const void *src, size_t bytes) {
    if (!EB.h_pin[0]) {
        for (int i = 0; i < 2; i++) {
            CK(cudaMallocHost(&EB.h_pin[i], ENS_PIN_CHUNK));
            CK(cudaEventCreateWithFlags(&EB.ev_pin[i], cudaEventDisableTiming));
        }
    }
    size_t prev_off = 0, prev_n = 0;
    int k = 0;
    for (size_t off = 0; off < bytes; off += ENS_PIN_CHUNK, k++) {
        const size_t nb = std::min(ENS_PIN_CHUNK, bytes - off);
        CK(cudaMemcpyAsync(EB.h_pin[k & 1], (const char *)src + off, nb, cudaMemcpyDeviceToHost, E.stream));
        CK(cudaEventRecord(EB.ev_pin[k & 1], E.stream));
        if (prev_n) {
            CK(cudaEventSynchronize(EB.ev_pin[(k - 1) & 1]));
            memcpy((char *)dst + prev_off, EB.h_pin[(k - 1) & 1], prev_n);
        }
        prev_off = off, prev_n = nb;
    }
    if (prev_n) {
        CK(cudaEventSynchronize(EB.ev_pin[(k - 1) & 1]));
        memcpy((char *)dst + prev_off, EB.h_pin[(k - 1) & 1], prev_n);
    }
}

}  // namespace spdy

extern "C" {

// ---- multi-GPU communicator (one process per GPU; rendezvous of the 128-byte id is the caller's business) ----------
int spdy_comm_unique_id(void *id128) {
    API_LOCK;
    if (!nccl_load()) return -1;
    ncclUniqueId id;
    nck(NC.GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id128, &id, sizeof(id));
    return 0;
}
int spdy_comm_init(int rank, int world, const void *id128) {
    API_LOCK;
    if (NC.comm) return -2;
    if (!nccl_load()) return -1;
    engine_init();  // selects E.device (spdy_set_device) and creates the stream the collectives run on
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    nck(NC.CommInitRank(&NC.comm, world, id, rank), "ncclCommInitRank");
    NC.rank = rank, NC.world = world;
    CK(cudaMalloc(&NC.d_small, 64 * sizeof(double)));
    return 0;
}
int spdy_comm_rank(void) { API_LOCK; return NC.rank; }
int spdy_comm_world(void) { API_LOCK; return NC.world; }
// in-place all-reduce of a small host vector (n <= 64): op 0 = sum, 1 = max.  Used for barriers and for the max-over-ranks
// of timings; a no-op with a single rank.
int spdy_comm_allreduce(double *host_inout, int n, int op) {
    API_LOCK;
    if (n < 0 || n > 64) return -1;
    if (!NC.comm) return 0;
    CK(cudaMemcpyAsync(NC.d_small, host_inout, n * sizeof(double), cudaMemcpyHostToDevice, E.stream));
    nck(NC.AllReduce(NC.d_small, NC.d_small, (size_t)n, ncclDouble, op == 1 ? ncclMax : ncclSum, NC.comm, E.stream), "ncclAllReduce");
    CK(cudaMemcpyAsync(host_inout, NC.d_small, n * sizeof(double), cudaMemcpyDeviceToHost, E.stream));
    CK(cudaStreamSynchronize(E.stream));
    return 0;
}
int spdy_comm_barrier(void) {
    API_LOCK;
    if (E.ready) CK(cudaStreamSynchronize(E.stream));
    double x = 1.0;
    return spdy_comm_allreduce(&x, 1, 0);
}
int spdy_comm_destroy(void) {
    API_LOCK;
    if (NC.comm) {
        CK(cudaStreamSynchronize(E.stream));
        nck(NC.CommDestroy(NC.comm), "ncclCommDestroy");
        NC.comm = nullptr, NC.rank = 0, NC.world = 1;
        cudaFree(NC.d_small), NC.d_small = nullptr;
    }
    return 0;
}

// ---- once-a-day ensemble diagnostics ---------------------------------------------------------------------------------
// transform_spectral2grid (prognostics.f90:125-154) for the listed members AND, in the same pass, the ensemble mean and
// spread (standard deviation, ddof = 0: examples/Ensemble_forecast.ipynb cells 12, 16) of u, v, t, q, phi, ps on the grid:
// partial sums in the transform's epilogue, one deterministic second stage, one in-stream ncclAllReduce over all ranks
// when a communicator is up, mean / spread formed on the device, ONE device-to-host copy and ONE synchronisation.
// out[0 : 188928] = mean, out[188928 : 377856] = spread, each the six variables in Fortran order one after another
// (u, v, t, q, phi: (96,48,8); ps: (96,48)).  n_total = ensemble size over all ranks.  out == NULL: results stay on the
// device (spdy_ensemble_stats_device_ptr).  Returns 0.
int spdy_ensemble_mean_spread(const int64_t *states, int n, long long n_total, double *out) {
    API_LOCK;
    engine_init();
    const int nt = prepare_members(states, n);
    const int groups_total = (nt + ENS_TG - 1) / ENS_TG + (nt + E.chunk_tiles - 1) / E.chunk_tiles;  // upper bound
    if ((size_t)groups_total * 2 * ENS_NTOT > EB.part_cap) {
        if (EB.d_part) CK(cudaFree(EB.d_part));
        EB.part_cap = (size_t)groups_total * 2 * ENS_NTOT;
        CK(cudaMalloc(&EB.d_part, EB.part_cap * sizeof(double)));
    }
    if (!EB.d_stats) {
        CK(cudaMalloc(&EB.d_stats, 2 * ENS_NTOT * sizeof(double)));
        CK(cudaMallocHost(&EB.h_stats, 2 * ENS_NTOT * sizeof(double)));
    }
    int group0 = 0;
    for (int t0 = 0; t0 < nt; t0 += E.chunk_tiles) {
        const int ntc = std::min(E.chunk_tiles, nt - t0), ng = (ntc + ENS_TG - 1) / ENS_TG;
        Ctx c = make_ctx(E.d_tiles + t0, E.d_masks + t0, ntc);
        s2g_transforms(c);
        k_s2g_finish_stats<<<dim3(NG / 4, ng), 128, 0, E.stream>>>(c, E.L, EB.d_part, group0);
        COUNT(1);
        group0 += ng;
    }
    k_ens_reduce<<<(unsigned)((ENS_NTOT + 255) / 256), 256, 0, E.stream>>>(EB.d_part, group0, ENS_NTOT, EB.d_stats,
                                                                           EB.d_stats + ENS_NTOT);
    COUNT(1);
    if (NC.comm)
        nck(NC.AllReduce(EB.d_stats, EB.d_stats, (size_t)(2 * ENS_NTOT), ncclDouble, ncclSum, NC.comm, E.stream), "ncclAllReduce");
    k_mean_spread<<<(unsigned)((ENS_NTOT + 255) / 256), 256, 0, E.stream>>>(EB.d_stats, ENS_NTOT, 1.0 / (double)n_total);
    COUNT(1);
    if (out) {
        CK(cudaMemcpyAsync(EB.h_stats, EB.d_stats, 2 * ENS_NTOT * sizeof(double), cudaMemcpyDeviceToHost, E.stream));
        CK(cudaStreamSynchronize(E.stream));
        memcpy(out, EB.h_stats, 2 * ENS_NTOT * sizeof(double));
    }
    return 0;
}
void *spdy_ensemble_stats_device_ptr(void) { API_LOCK; return EB.d_stats; }

// partial sums of ONE registry array variable over the listed members (any variable; mean_and_spread with a shift):
// sum[i] = sum_m x_m[i], sumsq[i] = sum_m (x_m[i]-shift[i])^2; results stay on the device (2 * nelem doubles)
static double *g_sums = nullptr, *g_part = nullptr;
static size_t g_sums_n = 0, g_part_n = 0;
int spdy_ensemble_sums_device(const int64_t *states, int n_members, int var, const double *shift_dev, void **out, size_t *nelem) {
    API_LOCK;
    engine_init();
    if (var < 0 || var >= SPDY_NVARS || E.off[var] < 0) return -1;
    const long long n = E.nelem[var];
    if ((size_t)n > g_sums_n) {
        if (g_sums) CK(cudaFree(g_sums));
        CK(cudaMalloc(&g_sums, 2 * n * sizeof(double)));
        g_sums_n = n;
    }
    const int nt = prepare_members(states, n_members);
    const int groups = (nt + ENS_TG - 1) / ENS_TG;
    if ((size_t)groups * 2 * n > g_part_n) {
        if (g_part) CK(cudaFree(g_part));
        g_part_n = (size_t)groups * 2 * n;
        CK(cudaMalloc(&g_part, g_part_n * sizeof(double)));
    }
    Ctx c = make_ctx(E.d_tiles, E.d_masks, nt);
    k_ens_sums<<<dim3((unsigned)((n + 7) / 8), groups), 256, 0, E.stream>>>(c, E.off[var], n, shift_dev, g_part);
    k_ens_reduce<<<(unsigned)((n + 255) / 256), 256, 0, E.stream>>>(g_part, groups, n, g_sums, g_sums + n);
    COUNT(2);
    CK(cudaStreamSynchronize(E.stream));
    *out = g_sums;
    *nelem = (size_t)n;
    return 0;
}
// same, summed over all ranks when a communicator is up, copied to the host
int spdy_ensemble_sums(const int64_t *states, int n_members, int var, const double *shift, double *sum, double *sumsq) {
    API_LOCK;
    engine_init();
    if (var < 0 || var >= SPDY_NVARS || E.off[var] < 0) return -1;
    const long long n = E.nelem[var];
    double *d_shift = nullptr;
    if (shift) {
        CK(cudaMalloc(&d_shift, n * sizeof(double)));
        CK(cudaMemcpy(d_shift, shift, n * sizeof(double), cudaMemcpyHostToDevice));
    }
    void *dev;
    size_t ne;
    spdy_ensemble_sums_device(states, n_members, var, d_shift, &dev, &ne);
    if (NC.comm) {
        nck(NC.AllReduce(dev, dev, (size_t)(2 * n), ncclDouble, ncclSum, NC.comm, E.stream), "ncclAllReduce");
        CK(cudaStreamSynchronize(E.stream));
    }
    CK(cudaMemcpy(sum, dev, n * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sumsq, (double *)dev + n, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (d_shift) CK(cudaFree(d_shift));
    return 0;
}

// ---- bulk accessors for a member list -----------------------------------------------------------------------------------
// get_<var> for every listed member in one call: dst[i][...] = the variable of member i in Fortran order, as float64
// (as_f32 = 0) or cast to float32 on the device (as_f32 = 1; the cast of pyspeedy/speedy.py:443).  One gather kernel and
// one device-to-host copy per block of members (staging buffer <= 256 MB) instead of n x spdy_get.
int spdy_ensemble_get(const int64_t *states, int n, int var, void *dst, size_t bytes, int as_f32) {
    API_LOCK;
    engine_init();
    if (var < 0 || var >= SPDY_NVARS || E.off[var] < 0 || n <= 0) return -1;
    const long long ne = E.nelem[var];
    const size_t esz = as_f32 ? 4 : 8;
    if (bytes != (size_t)n * ne * esz) return -2;
    std::vector<int2> who(n);
    for (int i = 0; i < n; i++) {
        Member *m = member_of(states[i]);
        if (!m) return -1;
        who[i] = make_int2(m->tile, m->lane);
    }
    if (n > EB.who_cap) {
        if (EB.d_who) CK(cudaFree(EB.d_who));
        EB.who_cap = n;
        CK(cudaMalloc(&EB.d_who, (size_t)n * sizeof(int2)));
    }
    CK(cudaMemcpyAsync(EB.d_who, who.data(), (size_t)n * sizeof(int2), cudaMemcpyHostToDevice, E.stream));
    const size_t per = (size_t)ne * esz;
    int block = (int)std::max<size_t>(32, ((size_t)256 << 20) / per / 32 * 32);
    block = std::min(block, (n + 31) / 32 * 32);
    if ((size_t)block * per > EB.stage_cap) {
        if (EB.d_stage) CK(cudaFree(EB.d_stage));
        EB.stage_cap = (size_t)block * per;
        CK(cudaMalloc(&EB.d_stage, EB.stage_cap));
    }
    for (int i0 = 0; i0 < n; i0 += block) {
        const int nb = std::min(block, n - i0);
        const dim3 grid((unsigned)((ne + 31) / 32), (unsigned)((nb + 31) / 32));
        if (as_f32)
            k_gather_members<float><<<grid, 256, 0, E.stream>>>(E.st, E.st_elems, E.off[var], ne, EB.d_who, i0 + nb, i0, (float *)EB.d_stage);
        else
            k_gather_members<double><<<grid, 256, 0, E.stream>>>(E.st, E.st_elems, E.off[var], ne, EB.d_who, i0 + nb, i0, (double *)EB.d_stage);
        COUNT(1);
        d2h_bounced((char *)dst + (size_t)i0 * per, EB.d_stage, (size_t)nb * per);
    }
    return 0;
}

// check (speedy_driver.f90.j2:81-91, diagnostics.f90:16-74 on time level 1) for every listed member in one call
int spdy_batch_check(const int64_t *states, int n, int *error_codes) {
    API_LOCK;
    engine_init();
    for (int i = 0; i < n; i++) error_codes[i] = member_of(states[i]) ? 0 : -1;
    const int nt = prepare_members(states, n);
    if (nt == 0) return 0;
    ensure_err_capacity(nt);
    for (int t0 = 0; t0 < nt; t0 += E.chunk_tiles) {
        const int ntc = std::min(E.chunk_tiles, nt - t0);
        Ctx c = make_ctx(E.d_tiles + t0, E.d_masks + t0, ntc);
        k_set_slot<<<ntc, 32, 0, E.stream>>>(c, SL_ERR, 0.0);
        launch_diag(E.stream, c, 1, E.L.diagp, 0);
        k_get_err<<<ntc, 32, 0, E.stream>>>(c, E.d_err + t0 * TILE);
        COUNT(4);
    }
    CK(cudaMemcpyAsync(E.h_err, E.d_err, (size_t)nt * TILE * sizeof(int), cudaMemcpyDeviceToHost, E.stream));
    CK(cudaStreamSynchronize(E.stream));
    std::map<int, int> tile_pos;
    for (int t = 0; t < nt; t++) tile_pos[E.cached_tiles[t]] = t;
    for (int i = 0; i < n; i++) {
        const Member *m = member_of(states[i]);
        if (m) error_codes[i] = E.h_err[tile_pos[m->tile] * TILE + m->lane];
    }
    return 0;
}

// datetime containers updated in place (the reference frees and re-creates them: pyspeedy/speedy.py:176-186); the batched
// form sets the same date in every listed container (SpeedyEns.run: one call per step instead of 2 x n)
int spdy_set_datetime(int64_t dt, int y, int mo, int d, int h, int mi) {
    API_LOCK;
    if (dt < 1 || dt > (int64_t)E.dates.size() || !E.dates[dt - 1].alive) return -1;
    E.dates[dt - 1] = Datetime{y, mo, d, h, mi, true};
    return 0;
}
int spdy_set_datetimes(const int64_t *dts, int n, const int *ymdhm) {
    API_LOCK;
    int bad = 0;
    for (int i = 0; i < n; i++) bad += spdy_set_datetime(dts[i], ymdhm[0], ymdhm[1], ymdhm[2], ymdhm[3], ymdhm[4]) != 0;
    return -bad;
}

}  // extern "C"

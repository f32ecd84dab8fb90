// mmh_api.cu — C ABI (include/mmhermite.h) over the sm_100a kernels: argument validation, per-device
// context (sqrt tables, workspace, grid-barrier counters), kernel selection, and the host-pointer
// variants that stage through device scratch.  No CPU compute path exists in this library.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "mmh_params.cuh"



// A failed runtime call leaves its code in the runtime's last-error slot; clear it so that the NEXT API call does not fail at its
// first cudaGetLastError() for something that happened here.  Device out-of-memory is the documented MemoryError class.
static inline int ck_fail(cudaError_t e) {
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? MMH_ERR_TOO_LARGE : (int)e;
}
#define CK(call)                                   \
    do {                                           \
        cudaError_t e_ = (call);                   \
        if (e_ != cudaSuccess) return ck_fail(e_); \
    } while (0)

static std::atomic<long long> g_launches{0};
static std::mutex g_mutex;

// ---- per-device context ---------------------------------------------------------------------------
struct Scratch {
    void *ptr = nullptr;
    size_t bytes = 0;
};
struct DeviceCtx {
    bool init = false;
    int sm_count = 0;
    int cc_major = 0;
    double *sq = nullptr, *rsq = nullptr;
    int table_len = 0;
    unsigned *barrier = nullptr;  // 64 counters; one is consumed per cooperative launch (round robin)
    int barrier_next = 0;
    Scratch diag_ws;              // auxiliary arrays of the compactFock sweeps
    Scratch xbuf;                 // halo exchange buffer of the tiled march; all-ones sentinel between launches
    Scratch partial;              // VJP partial sums
    Scratch norm;                 // binomial norm scalar
    Scratch timeline;             // 16 x 4 debug stamps of the last forward launches
    Scratch lattice_ws;           // lattice kept on the device by mmh_forward_contract
    Scratch ones;                 // vacuum amplitudes c = 1 of mmh_forward_contract
    Scratch ein_ws;               // offset tables of the Fock-space contraction
    Scratch dot_ws;               // per-CTA partials of mmh_overlap
    Scratch sbox_ws;              // stable box wavefront: ready flags + ticket of the current call
    Scratch gate_ws;              // gate strategies: log-factorial table, transposition buffer, masked cotangent
    Scratch host_slots[8];        // staging for the *_host entry points
    int *err_host = nullptr;      // mapped page-locked word the watchdogs of the polling kernels set when they give up
    int *err_dev = nullptr;       // its device alias
    cudaStream_t last_stream = nullptr;   // stream of the previous call that used the shared scratch (see enter_stream)
    bool have_last = false;
    cudaEvent_t xstream_ev = nullptr;
};
static DeviceCtx g_ctx[64];

static int ensure_scratch(Scratch &s, size_t bytes) {
    if (s.bytes >= bytes && s.ptr) return 0;
    if (s.ptr) { CK(cudaDeviceSynchronize()); CK(cudaFree(s.ptr)); s.ptr = nullptr; s.bytes = 0; }
    size_t want = bytes < 256 ? 256 : bytes;
    CK(cudaMalloc(&s.ptr, want));
    s.bytes = want;
    return 0;
}

static int ensure_tables(DeviceCtx &c, int need) {
    if (need <= c.table_len) return 0;
    int len = 4096;
    while (len < need) len *= 2;
    std::vector<double> hs(len), hr(len);
    for (int n = 0; n < len; n++) {
        hs[n] = std::sqrt((double)n);               // SQRT = np.sqrt(np.arange(...)) (core.py:22)
        hr[n] = n ? 1.0 / hs[n] : 0.0;              // correctly rounded reciprocal for div_by_table
    }
    if (c.sq) { CK(cudaDeviceSynchronize()); CK(cudaFree(c.sq)); CK(cudaFree(c.rsq)); c.sq = c.rsq = nullptr; c.table_len = 0; }
    CK(cudaMalloc(&c.sq, sizeof(double) * len));
    CK(cudaMalloc(&c.rsq, sizeof(double) * len));
    CK(cudaMemcpy(c.sq, hs.data(), sizeof(double) * len, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.rsq, hr.data(), sizeof(double) * len, cudaMemcpyHostToDevice));
    // the copies ran on the legacy default stream, which non-blocking streams (every torch side stream) do not synchronise
    // with: make the tables globally visible before any caller's stream can read them (once per table growth)
    CK(cudaDeviceSynchronize());
    c.table_len = len;
    return 0;
}

extern char **environ;
static std::vector<const char *> g_env_mmh;   // the "MMH_..." entries of the environment at the start of the current API call
void mmh_env_refresh() {
    g_env_mmh.clear();
    for (char **e = environ; e && *e; ++e)
        if ((*e)[0] == 'M' && (*e)[1] == 'M' && (*e)[2] == 'H' && (*e)[3] == '_') g_env_mmh.push_back(*e);
}
const char *mmh_getenv(const char *name) {
    const size_t n = strlen(name);
    for (const char *e : g_env_mmh)
        if (!strncmp(e, name, n) && e[n] == '=') return e + n + 1;
    return nullptr;
}

static int get_ctx(DeviceCtx **out) {
    mmh_env_refresh();   // every compute entry point comes through here first (under g_mutex)
    int dev = 0, ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) { cudaGetLastError(); return MMH_ERR_NO_DEVICE; }
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cudaGetLastError(); return MMH_ERR_NO_DEVICE; }
    if (dev < 0 || dev >= 64) return MMH_ERR_NO_DEVICE;
    DeviceCtx &c = g_ctx[dev];
    if (!c.init) {
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, dev);
        if (e != cudaSuccess) { cudaGetLastError(); return MMH_ERR_NO_DEVICE; }
        if (prop.major != 10) return MMH_ERR_NO_DEVICE;  // sm_100a cubin only
        c.sm_count = prop.multiProcessorCount;
        c.cc_major = prop.major;
        CK(cudaMalloc(&c.barrier, 64 * sizeof(unsigned)));
        CK(cudaMemset(c.barrier, 0, 64 * sizeof(unsigned)));
        CK(cudaHostAlloc((void **)&c.err_host, sizeof(int), cudaHostAllocMapped));
        *c.err_host = 0;
        CK(cudaHostGetDevicePointer((void **)&c.err_dev, c.err_host, 0));
        CK(cudaEventCreateWithFlags(&c.xstream_ev, cudaEventDisableTiming));
        CK(cudaDeviceSynchronize());   // legacy-stream initialisation done before any non-blocking stream runs
        c.init = true;
    }
    *out = &c;
    return 0;
}

// The per-device scratch (exchange buffer + its sentinel state, VJP partials, compactFock workspace, lattice workspace, grid-barrier
// counters) is ONE set shared by every call.  Calls on one stream are ordered by the stream; a call on a DIFFERENT stream than the
// previous one first waits (on the device, no host block) for everything enqueued so far on the previous stream, so two streams
// never run library kernels on the same scratch concurrently.  Single-stream callers pay nothing.  Callers hold g_mutex.
static int enter_stream(DeviceCtx *ctx, cudaStream_t st) {
    {   // a stream that is being captured into a CUDA graph must not wait on an event recorded outside the capture (that would
        // invalidate the capture); the graph's owner orders the replay against other work
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) return MMH_OK;
        cudaGetLastError();
    }
    if (ctx->have_last && ctx->last_stream != st) {
        cudaError_t e = cudaEventRecord(ctx->xstream_ev, ctx->last_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ctx->xstream_ev, 0);
        if (e != cudaSuccess) {   // the previous stream was destroyed meanwhile: its work is complete or abandoned; order by a full sync
            cudaGetLastError();
            CK(cudaDeviceSynchronize());
        }
    }
    ctx->last_stream = st;
    ctx->have_last = true;
    return MMH_OK;
}

// A polling kernel whose watchdog expired (a producer never delivered: a bug or a lost CTA) marched on with sentinel NaNs and left
// the exchange buffer dirty.  The flag it raised is reported by the next call on the device (and by the host-pointer entry points
// right after their own synchronisation): the exchange buffer is re-filled with the sentinel and MMH_ERR_TIMEOUT is returned.
static int check_watchdog(DeviceCtx *ctx, cudaStream_t st) {
    if (!ctx->err_host || *(volatile int *)ctx->err_host == 0) return MMH_OK;
    CK(cudaDeviceSynchronize());
    *(volatile int *)ctx->err_host = 0;
    if (ctx->xbuf.ptr) CK(cudaMemsetAsync(ctx->xbuf.ptr, 0xFF, ctx->xbuf.bytes, st));
    return MMH_ERR_TIMEOUT;
}

static int begin_call(DeviceCtx *ctx, cudaStream_t st) {
    int rc = check_watchdog(ctx, st);
    if (rc) return rc;
    return enter_stream(ctx, st);
}
// host-pointer entry points: after their own synchronisation, report a watchdog that expired during THIS call
static int end_host_call(DeviceCtx *ctx) { return check_watchdog(ctx, 0); }

static int make_desc(int ndim, const int64_t *shape, LatticeDesc *d, int *maxdim) {
    if (ndim < 1 || ndim > MMH_MAX_DIM) return MMH_ERR_BAD_NDIM;
    if (!shape) return MMH_ERR_NULL_POINTER;
    memset(d, 0, sizeof(*d));
    d->D = ndim;
    long long N = 1;
    int mx = 1;
    for (int i = 0; i < ndim; i++) {
        if (shape[i] < 1 || shape[i] > (1 << 24)) return MMH_ERR_BAD_SHAPE;
        d->shape[i] = (int)shape[i];
        if (shape[i] > mx) mx = (int)shape[i];
        if (N > (1LL << 40) / shape[i]) return MMH_ERR_TOO_LARGE;
        N *= shape[i];
    }
    d->N = N;
    d->strides[ndim - 1] = 1;
    for (int i = ndim - 1; i > 0; i--) d->strides[i - 1] = d->strides[i] * d->shape[i];
    *maxdim = mx;
    return 0;
}

static inline int round_up32(long long v) { return (int)((v + 31) / 32 * 32); }

// ---- forward --------------------------------------------------------------------------------------
static const long long kSmallPanel = 1024;     // panels up to this size are filled by one CTA
static const long long kSingleCtaN = 16384;    // lattices up to this size use the one-CTA kernel

// K2 launch plan for one stage: L lattices per CTA, R panel points per thread, T threads (mmh_march.cu)
static bool plan_march_stage(const LatticeDesc &d, int stage, long long batch, int *L_out, int *R_out, int *T_out,
                             size_t *smem_out) {
    const int npd = d.D - 1 - stage;
    if (npd < 1 || npd > 7) return false;
    const long long P = d.strides[stage];
    if (P > 1024) return false;
    const char *eL = mmh_getenv("MMH_K2_L"), *eR = mmh_getenv("MMH_K2_R");
    double best = -1.0;
    int bL = 0, bR = 0, bT = 0;
    size_t bsmem = 0;
    const int Lmax = (int)(batch < 256 ? batch : 256);
    for (int L = 1; L <= Lmax; L++) {
        const long long slots = L * P;
        if (slots > 1024) break;
        if (eL && stage == 0 && atoi(eL) != L) continue;
        const size_t smem = sizeof(c128) * (size_t)(2 * L + 2 * slots);
        const int Rs[3] = { 1, 2, 4 };
        for (int r = 0; r < 3; r++) {
            const int R = Rs[r];
            if (eR && stage == 0 && atoi(eR) != R) continue;
            if (R > 1 && R * npd > 8) continue;   // per-slot coefficients live in registers
            const int Tmax = R == 4 ? 256 : 512;
            int T = round_up32((slots + R - 1) / R);
            if (T > Tmax) continue;
            const double eff = (double)slots / ((double)T * R);
            // prefer full warps, two points per thread (ILP), CTAs of ~128-256 threads
            double score = eff + (R == 2 ? 0.02 : 0.0) - (T < 96 ? 0.05 : 0.0) - (T > 256 ? 0.03 : 0.0);
            if (score > best) { best = score; bL = L; bR = R; bT = T; bsmem = smem; }
        }
    }
    if (best < 0) return false;
    *L_out = bL; *R_out = bR; *T_out = bT; *smem_out = bsmem;
    return true;
}

// batched forward by stages: chain (stage D-1), then one march launch per stage D-2 .. 0
static int forward_staged(const FwdParams &p, DeviceCtx *ctx, cudaStream_t st, bool *done) {
    const LatticeDesc &d = p.d;
    *done = false;
    int L[MMH_MAX_DIM], R[MMH_MAX_DIM], T[MMH_MAX_DIM];
    size_t sm[MMH_MAX_DIM];
    bool box[MMH_MAX_DIM];
    BoxParams bp[MMH_MAX_DIM];
    for (int i = d.D - 2; i >= 0; i--) {
        box[i] = false;
        if (plan_march_stage(d, i, p.batch, &L[i], &R[i], &T[i], &sm[i])) continue;
        // panels beyond one CTA's shared memory: the same CTA marches the lattice box by box (mmh_box.cu)
        if (d.shape[i] == 1 || mmh_getenv("MMH_NO_BOX") || !mmh_plan_march_box(d, i, &bp[i], &T[i], &sm[i])) return MMH_OK;
        box[i] = true;
    }
    // stage D-2 of a real batch runs on the warp-synchronous lane march (mmh_lanes.cu), which then also computes the chain
    // (stage D-1) of its lattices itself: one launch instead of two, panel 0 never re-read from HBM
    int Rl = 0, ln = 0, Lw = 0;
    const bool lanes = d.D >= 2 && !box[d.D - 2] && p.batch >= 256 && d.shape[d.D - 2] > 1 && d.shape[d.D - 2] <= 4096 &&
                       !mmh_getenv("MMH_NO_LANES") && mmh_plan_march_lanes(d.shape[d.D - 1], &Rl, &ln, &Lw) &&
                       (long long)Lw * d.N < (1LL << 32);   // 32-bit store offsets
    const bool fuse_chain = lanes && !mmh_getenv("MMH_NO_FUSE_CHAIN");
    if (!fuse_chain) {
        g_launches++;
        CK(mmh_launch_chain(p, st));
    }
    for (int i = d.D - 2; i >= 0; i--) {
        if (d.shape[i] == 1) continue;   // nothing to march
        if (box[i]) {
            BoxParams &q = bp[i];
            q.A = p.A; q.b = p.b; q.G = p.G; q.sq = p.sq; q.rsq = p.rsq; q.batch = p.batch; q.lat_stride = d.N;
            g_launches++;
            CK(mmh_launch_march_box(q, ctx->sm_count, T[i], sm[i], st));
            continue;
        }
        StageParams sp;
        sp.d = d; sp.A = p.A; sp.b = p.b; sp.G = p.G; sp.sq = p.sq; sp.rsq = p.rsq;
        sp.batch = p.batch; sp.lat_stride = d.N; sp.stage = i; sp.L = L[i];
        sp.c = p.c; sp.fuse_chain = 0; sp.pdl = 0; sp.timeline = nullptr; sp.fill_n = 0;
        const long long grid = (p.batch + L[i] - 1) / L[i];
        g_launches++;
        if (i == d.D - 2 && lanes) {
            sp.fuse_chain = fuse_chain ? 1 : 0;
            CK(mmh_launch_march_lanes(sp, Rl, ln, Lw, ctx->sm_count, st));
            continue;
        }
        CK(mmh_launch_march_stage(sp, R[i], (int)grid, T[i], sm[i], st));
    }
    *done = true;
    return MMH_OK;
}

// K1r plan (mmh_rows.cu k_march_rows) for a given box grid over the panel dims of `stage` (all of them: npd = 2 or 3, so that the
// last panel dim is contiguous): lanes own R consecutive cells of a row.  R = 2 (16 compute warps, four per scheduler) measured
// best on cfg2 (121 us against 127 us at R = 5 and 131 us for k_march_tiled2); larger R only where R = 2 needs more than 512 lanes.
static bool plan_rows_for_grid(const LatticeDesc &d, int stage, const int *gin, int forcedR, TiledParams *tp, int *ntiles_out,
                               size_t *smem_out) {
    const int npd = d.D - 1 - stage;
    if (npd < 2 || npd > 3) return false;
    if (d.strides[stage] >= (1LL << 31)) return false;
    const int S = d.shape[stage];
    const int off = 3 - npd;
    int shp[3] = { 1, 1, 1 }, g[3] = { 1, 1, 1 }, em[3];
    for (int m = off; m < 3; m++) { shp[m] = d.shape[stage + 1 + m - off]; g[m] = gin[m - off]; }
    for (int m = 0; m < 3; m++) {
        if (g[m] < 1 || g[m] > shp[m]) return false;
        em[m] = (shp[m] + g[m] - 1) / g[m];
    }
    const long long HC = (long long)(g[0] > 1) * em[1] * em[2] + (long long)(g[1] > 1) * em[0] * em[2] + (long long)(g[2] > 1) * em[0] * em[1];
    const long long HCs = HC > 0 ? HC : 1;
    if ((double)g[0] * g[1] * g[2] * (double)S * (double)HCs >= 2147483648.0) return false;   // 32-bit export offsets
    for (int R = 2; R <= 6; R++) {
        if (forcedR && R != forcedR) continue;
        const int C = (em[2] + R - 1) / R;
        const long long lanes = (long long)em[0] * em[1] * C;
        if (lanes > mmh_rows_max_threads(R)) continue;
        const int TC = round_up32(lanes);
        // shared-memory row stride (16-byte cells): cell r of chunk ch sits at r * C + ch, so for a fixed r the C lanes of a row read
        // consecutive cells; with RS congruent to C modulo 8 consecutive rows continue the sequence of bank groups (conflict free)
        int RS = C * R;
        while ((RS & 7) != (C & 7)) RS++;
        const long long next = (long long)(em[0] + 1) * (em[1] + 1);
        const long long LS = next * (RS + 1) + RS + 1;   // rows | dim-2 halo cells | zero row
        const int cells_max = em[0] * em[1] * em[2];
        const size_t smem = mmh_rows_smem((int)LS, (int)HCs, S, C * R, cells_max, 0);
        if (smem > 200 * 1024) continue;
        memset(tp, 0, sizeof(*tp));
        tp->d = d; tp->stage = stage; tp->nt = npd;
        for (int m = off; m < 3; m++) tp->g[m - off] = g[m];
        for (int m = npd; m < 3; m++) tp->g[m] = 1;
        tp->tc = TC; tp->ls_max = (int)LS; tp->hc_max = (int)HCs;
        tp->rows_R = R; tp->rows_C = C; tp->rs = RS; tp->cells_max = cells_max; tp->xc_max = 0;
        *ntiles_out = g[0] * g[1] * g[2];
        *smem_out = smem;
        return true;
    }
    return false;
}

// K1 plan: tile grid over the first nt panel dims of `stage` (mmh_march.cu k_march_tiled)
static bool plan_march_tiled(const LatticeDesc &d, int stage, int sm_count, TiledParams *tp, int *R_out,
                             int *ntiles_out, size_t *smem_out, double *cost_out = nullptr) {
    // test / tuning hook: MMH_ROWS_G="g0,g1,g2" forces k_march_rows with that box grid (MMH_ROWS_R: cells per lane)
    const int rows_forcedR = mmh_getenv("MMH_ROWS_R") ? atoi(mmh_getenv("MMH_ROWS_R")) : 0;
    if (const char *eg = mmh_getenv("MMH_ROWS_G"))
        if (!mmh_getenv("MMH_NO_ROWS") && (!mmh_getenv("MMH_TILE_STAGE") || atoi(mmh_getenv("MMH_TILE_STAGE")) == stage)) {
            int fg[3] = { 1, 1, 1 };
            sscanf(eg, "%d,%d,%d", &fg[0], &fg[1], &fg[2]);
            if (plan_rows_for_grid(d, stage, fg, rows_forcedR, tp, ntiles_out, smem_out)) { *R_out = tp->rows_R; return true; }
        }
    const int npd = d.D - 1 - stage;
    if (npd < 1 || npd > 7) return false;
    const long long P = d.strides[stage];
    if (P >= (1LL << 31)) return false;
    const int S = d.shape[stage];
    const int nt = npd < 3 ? npd : 3;
    const long long inner = d.strides[stage + nt];
    int shp[3] = { 1, 1, 1 };
    for (int m = 0; m < nt; m++) shp[m] = d.shape[stage + 1 + m];
    int forced[3] = { 0, 0, 0 };
    if (const char *eg = mmh_getenv("MMH_TILE_G"))   // test / tuning hook; MMH_TILE_STAGE restricts it to one stage
        if (!mmh_getenv("MMH_TILE_STAGE") || atoi(mmh_getenv("MMH_TILE_STAGE")) == stage)
            sscanf(eg, "%d,%d,%d", &forced[0], &forced[1], &forced[2]);
    double best = 1e300;
    int bg[3] = { 0, 0, 0 }, bR = 0, bTC = 0, bLS = 0, bHC = 0;
    // one tile-owner CTA per SM (640 threads x 96 registers fill the register file); all tiles must be co-resident
    const int maxt = sm_count;
    for (int g0 = 1; g0 <= shp[0] && g0 <= maxt; g0++)
        for (int g1 = 1; g1 <= shp[1] && g0 * g1 <= maxt; g1++)
            for (int g2 = 1; g2 <= shp[2] && g0 * g1 * g2 <= maxt; g2++) {
                if (forced[0] && (g0 != forced[0] || g1 != (forced[1] ? forced[1] : 1) || g2 != (forced[2] ? forced[2] : 1)))
                    continue;
                const int g[3] = { g0, g1, g2 };
                long long e[3], TS = inner, LS = inner;
                for (int m = 0; m < 3; m++) {
                    e[m] = (shp[m] + g[m] - 1) / g[m];
                    TS *= e[m];
                    LS *= e[m] + (g[m] > 1 ? 1 : 0);
                }
                if (TS > 1024) continue;
                long long HC = 0;
                for (int m = 0; m < 3; m++) if (g[m] > 1) HC += TS / e[m];
                const long long HCs = HC > 0 ? HC : 1;   // X rows are laid out with stride hc_max >= 1
                int R = 0;
                const int Rs[2] = { 1, 2 };
                const char *eR = mmh_getenv("MMH_TILE_R");
                for (int r = 0; r < 2; r++) {
                    const int tcmax = 512;
                    if (eR && atoi(eR) != Rs[r]) continue;
                    if ((long long)Rs[r] * tcmax >= TS && (Rs[r] == 1 || Rs[r] * npd <= 10)) { R = Rs[r]; break; }
                }
                if (!R) continue;
                const int TC = round_up32((TS + R - 1) / R);
                LS = TS + HC + 2;                           // compact box + halo faces + zero cell + trash cell
                const size_t smem = mmh_tiled2_smem((int)LS, (int)HCs, S, R * TC);
                if ((double)g0 * g1 * g2 * (double)S * (double)HCs >= 2147483648.0) continue;   // 32-bit export offsets
                if (smem > 200 * 1024) continue;
                // cost model (us): issue time of a step ~ points of the tile, floor = dependent chain of a step, plus the hops of
                // the longest tile-pipeline path.  The constants are the ones the measured optimum of cfg2 follows from
                // (stage 0: 5x5x5, stage 1: 4x4; a re-fit to the absolute step times picked 5x5 for stage 1 and lost 14 us).
                double step_us = (double)TS * 0.55e-3;
                if (step_us < 0.12) step_us = 0.12;
                const double cost = (S - 1) * step_us + (g0 + g1 + g2 - 3) * 0.7;
                if (cost < best) {
                    best = cost; bg[0] = g0; bg[1] = g1; bg[2] = g2; bR = R; bTC = TC; bLS = (int)LS; bHC = (int)HC;
                    *smem_out = smem;
                }
            }
    if (best > 1e299) return false;
    // Boxes with enough arithmetic per step run on the row-lane kernel (mmh_rows.cu) with the same grid: its compute warps do
    // nothing but the recurrence (service warps import, export and drain).  Small boxes stay on k_march_tiled2, whose step is
    // bound by the hand-off latency either way (measured: (40,)^4 in 8x8x8 boxes 84 vs 80 us, (50,)^4 in 10x10x10 boxes 121 vs 131 us).
    if (npd == 3 && inner == 1 && !mmh_getenv("MMH_NO_ROWS") && !forced[0]) {
        const long long min_cells = mmh_getenv("MMH_ROWS_MIN_CELLS") ? atoll(mmh_getenv("MMH_ROWS_MIN_CELLS")) : 700;
        long long cells = 1;
        for (int m = 0; m < 3; m++) cells *= (shp[m] + bg[m] - 1) / bg[m];
        if (cells >= min_cells && plan_rows_for_grid(d, stage, bg, rows_forcedR, tp, ntiles_out, smem_out)) {
            *R_out = tp->rows_R;
            return true;
        }
    }
    memset(tp, 0, sizeof(*tp));
    tp->d = d; tp->stage = stage; tp->nt = nt;
    for (int m = 0; m < 3; m++) tp->g[m] = bg[m];
    tp->tc = bTC; tp->ls_max = bLS; tp->hc_max = bHC > 0 ? bHC : 1;
    *R_out = bR;
    *ntiles_out = bg[0] * bg[1] * bg[2];
    return true;
}

// The brute-force search above costs tens of microseconds of host time -- more than the GPU needs for a small stage, so an
// uncached plan leaves the device idle between the launches of one lattice (measured: 12 us gaps on cfg2).  Plans are cached per
// (shape, stage, SM count, tuning environment); callers hold g_mutex.
struct TiledPlan { bool ok; TiledParams tp; int R, ntiles; size_t smem; double cost_us; };
static bool plan_march_tiled_cached(const LatticeDesc &d, int stage, int sm_count, TiledParams *tp, int *R_out,
                                    int *ntiles_out, size_t *smem_out, double *cost_out = nullptr) {
    static std::map<std::string, TiledPlan> cache;
    std::string key((const char *)d.shape, sizeof(int) * (size_t)d.D);
    key.push_back((char)stage); key.push_back((char)d.D); key.append(std::to_string(sm_count));
    for (const char *name : { "MMH_TILE_G", "MMH_TILE_STAGE", "MMH_TILE_R", "MMH_NO_ROWS", "MMH_ROWS_G", "MMH_ROWS_R", "MMH_ROWS_MIN_CELLS" }) {
        const char *v = mmh_getenv(name);
        key.push_back('|');
        if (v) key.append(v);
    }
    auto it = cache.find(key);
    if (it == cache.end()) {
        TiledPlan pl;
        memset(&pl, 0, sizeof(pl));
        pl.ok = plan_march_tiled(d, stage, sm_count, &pl.tp, &pl.R, &pl.ntiles, &pl.smem, &pl.cost_us);
        if (cache.size() > 4096) cache.clear();
        it = cache.emplace(key, pl).first;
    }
    const TiledPlan &pl = it->second;
    if (!pl.ok) return false;
    *tp = pl.tp; *R_out = pl.R; *ntiles_out = pl.ntiles; *smem_out = pl.smem;
    if (cost_out) *cost_out = pl.cost_us;
    return true;
}

// one large lattice: chain, then per stage the smallest machinery that fits
// (single-CTA march / tiled multi-CTA march / plain per-step launches for giant panels)
// `seq` = index of the lattice inside one API call.  Consecutive lattices of a batch are pipelined: lattice seq+1's kernels are
// launched with programmatic stream serialization behind lattice seq's last kernel, so its latency-bound small stages and the
// fill of its tile pipeline run on the SMs that lattice seq's tile pipeline has already vacated.  Every kMaxInFlight-th lattice
// is launched in plain stream order (full wait), which makes the reuse of an exchange-buffer slot (seq % kMaxInFlight) safe.
static const int kMaxInFlight = 4;
struct PendingTrace { unsigned long long *dev; size_t words; int stage; int hdr[8]; };
static int forward_single_staged(const FwdParams &p, DeviceCtx *ctx, cudaStream_t st, long long seq) {
    const LatticeDesc &d = p.d;
    const int D = d.D;
    // The launches of one lattice are chained with programmatic dependent launch: each kernel releases its
    // successor at once and the successor's prologue (index decode, tables) overlaps it, blocking only where it first
    // touches the lattice.  The chain (stage D-1) is fused into the first march launch when that is a single CTA.
    const size_t absmem = sizeof(c128) * (size_t)(D * D + D);
    const bool use_pdl = !mmh_getenv("MMH_NO_PDL");
    if (!ctx->timeline.ptr) {
        int rc0;
        if ((rc0 = ensure_scratch(ctx->timeline, 256 * sizeof(unsigned long long)))) return rc0;
        CK(cudaMemsetAsync(ctx->timeline.ptr, 0, 256 * sizeof(unsigned long long), st));
    }
    bool chain_done = false, first = true;
    int rc;
    std::vector<PendingTrace> traces;   // MMH_TRACE_FILE: per-stage debug timelines, read back at the end
    const size_t kTraceArenaWords = 4u << 20;   // one arena, allocated and cleared before the first launch (no sync between stages)
    unsigned long long *trace_arena = nullptr;
    size_t trace_used = 0;
    if (mmh_getenv("MMH_TRACE_FILE")) {
        CK(cudaMalloc(&trace_arena, kTraceArenaWords * 8));
        CK(cudaMemsetAsync(trace_arena, 0, kTraceArenaWords * 8, st));
    }
    // Stage overlap: when the last two marched stages are both tiled, the last one does not wait for its predecessor's
    // kernel but for the amplitudes themselves (TiledParams::poll0).  Its panel 0 -- the first strides[i0] amplitudes of
    // G -- is pre-filled with the sentinel before anything runs, and the two kernels get disjoint exchange buffers.
    int i0 = -1, i1 = -1;            // last marched stage and its predecessor
    for (int i = 0; i < D - 1 && i1 < 0; i++) {
        if (d.shape[i] == 1) continue;
        if (i0 < 0) i0 = i; else i1 = i;
    }
    bool overlap = false, overlap1 = false;   // overlap1: stage i1 in turn overlaps the one-warp kernel of the trailing stages
    bool pipelined = false;          // this lattice's first kernels are chained behind the previous lattice's (see kMaxInFlight)
    size_t xbase = 0;                // exchange-buffer slot of this lattice
    size_t xoff1 = 0;                // exchange-buffer offset (bytes) of stage i0 when it overlaps stage i1
    if (use_pdl && i1 >= 0 && (!mmh_getenv("MMH_TRACE_FILE") || mmh_getenv("MMH_TRACE_OVERLAP")) && !mmh_getenv("MMH_NO_OVERLAP")) {
        int L_, R_, T_, n0 = 0, n1 = 0;
        size_t sm_;
        TiledParams t0, t1;
        const bool k2_0 = !mmh_getenv("MMH_FORCE_TILED") && plan_march_stage(d, i0, 1, &L_, &R_, &T_, &sm_);
        const bool k2_1 = !mmh_getenv("MMH_FORCE_TILED") && plan_march_stage(d, i1, 1, &L_, &R_, &T_, &sm_);
        const bool tail1 = i1 == D - 2 && d.shape[D - 1] <= 64 && !mmh_getenv("MMH_NO_WARP_TAIL") && !mmh_getenv("MMH_FORCE_TILED");
        if (!k2_0 && !k2_1 && !tail1 && plan_march_tiled_cached(d, i0, ctx->sm_count, &t0, &R_, &n0, &sm_) &&
            plan_march_tiled_cached(d, i1, ctx->sm_count, &t1, &R_, &n1, &sm_) && n0 + n1 <= ctx->sm_count &&
            d.strides[i0] * (long long)sizeof(c128) <= (64LL << 20)) {
            overlap = true;
            overlap1 = i1 == D - 3 && d.shape[D - 2] > 1 && d.shape[D - 1] <= 64 && !mmh_getenv("MMH_NO_WARP_TAIL") &&
                       !mmh_getenv("MMH_FORCE_TILED") && !mmh_getenv("MMH_NO_OVERLAP1");
            xoff1 = (sizeof(c128) * (size_t)n1 * d.shape[i1] * t1.hc_max + 255) / 256 * 256;
            const size_t xslot = (xoff1 + sizeof(c128) * (size_t)n0 * d.shape[i0] * t0.hc_max + 255) / 256 * 256;
            const size_t xtotal = xslot * (size_t)kMaxInFlight;
            if (ctx->xbuf.bytes < xtotal || !ctx->xbuf.ptr) {   // every kernel's exchange buffer, before any is launched
                if ((rc = ensure_scratch(ctx->xbuf, xtotal))) return rc;
                CK(cudaMemsetAsync(ctx->xbuf.ptr, 0xFF, ctx->xbuf.bytes, st));   // on the caller's stream: ordered before its kernels
            }
            xbase = xslot * (size_t)(seq % kMaxInFlight);
            pipelined = overlap1;   // the first kernel is then the one-warp tail, which fills panel 0 with the sentinel itself
            if (!pipelined) {       // else: a separate fill kernel in plain stream order
                g_launches++;
                CK(mmh_launch_fill_sentinel((c128 *)p.G, d.strides[i0], false, st));
            }
        }
    }
    for (int i = D - 2; i >= 0; i--) {
        if (d.shape[i] == 1) continue;
        int L, R, T, ntiles;
        size_t sm;
        TiledParams tp;
        if (i == D - 2 && !chain_done && d.shape[D - 1] <= 64 && !mmh_getenv("MMH_NO_WARP_TAIL") && !mmh_getenv("MMH_FORCE_TILED")) {
            // the two trailing stages by one warp (shuffles instead of shared memory + barriers)
            StageParams sp;
            sp.d = d; sp.A = p.A; sp.b = p.b; sp.G = p.G; sp.sq = p.sq; sp.rsq = p.rsq;
            sp.batch = 1; sp.lat_stride = d.N; sp.stage = i; sp.L = 1;
            sp.c = p.c; sp.fuse_chain = 1; sp.fill_n = pipelined ? d.strides[i0] : 0;
            sp.pdl = (pipelined && (seq % kMaxInFlight) != 0) ? 1 : 0; sp.timeline = (unsigned long long *)ctx->timeline.ptr + 64 * (seq % kMaxInFlight);
            chain_done = true; first = false;
            g_launches++;
            CK(mmh_launch_warp_tail(sp, st));
            continue;
        }
        // one CTA (K2) or many tiles (K1) for a panel that would fit one CTA?  One CTA marches ~1 ns per point and step
        // (1.0 us at 1000 points, measured); tiles of 100-200 points step in ~0.15 us and pay ~1 us per pipeline hop once.
        // Long marches of mid-size panels ((1000,1000): 999 steps of 1000 points) are 3x faster tiled.
        bool prefer_tiled = false;
        if (!mmh_getenv("MMH_FORCE_TILED") && !mmh_getenv("MMH_NO_PREFER_TILED") && d.strides[i] >= 192 && d.strides[i] <= 1024) {
            double tcost = 0.0;
            int R_, n_;
            size_t sm_;
            TiledParams t_;
            const double k2cost = (d.shape[i] - 1) * (0.05 + 1.0e-3 * (double)d.strides[i]);
            if (plan_march_tiled_cached(d, i, ctx->sm_count, &t_, &R_, &n_, &sm_, &tcost) && n_ > 1) {
                // measured step of a small tile WITH halos: ~0.27 us (hand-offs, not arithmetic), 1 ns per point above that
                double tstep = 0.05 + 1.0e-3 * (double)d.strides[i] / n_;
                if (tstep < 0.27) tstep = 0.27;
                if ((d.shape[i] - 1) * tstep < 0.75 * k2cost) prefer_tiled = true;
            }
        }
        if (!prefer_tiled && !mmh_getenv("MMH_FORCE_TILED") && plan_march_stage(d, i, 1, &L, &R, &T, &sm)) {
            StageParams sp;
            sp.d = d; sp.A = p.A; sp.b = p.b; sp.G = p.G; sp.sq = p.sq; sp.rsq = p.rsq;
            sp.batch = 1; sp.lat_stride = d.N; sp.stage = i; sp.L = L;
            sp.c = p.c; sp.fuse_chain = chain_done ? 0 : 1; sp.pdl = (use_pdl && !first) ? 1 : 0; sp.timeline = nullptr; sp.fill_n = 0;
            chain_done = true; first = false;
            g_launches++;
            CK(mmh_launch_march_stage(sp, R, 1, T, sm, st));
            continue;
        }
        if (!chain_done) {
            g_launches++;
            CK(mmh_launch_chain(p, st));
            chain_done = true; first = false;
        }
        if (0) {
        } else if (plan_march_tiled_cached(d, i, ctx->sm_count, &tp, &R, &ntiles, &sm)) {
            tp.err = ctx->err_dev;
            tp.A = p.A; tp.b = p.b; tp.G = p.G; tp.sq = p.sq; tp.rsq = p.rsq; tp.timeline = (unsigned long long *)ctx->timeline.ptr + 64 * (seq % kMaxInFlight);
            tp.pdl = (use_pdl && !first) ? 1 : 0;
            first = false;
            tp.poll0 = ((overlap && i == i0) || (overlap1 && i == i1)) ? 1 : 0;
            tp.strong_g = (overlap && i == i1) ? 1 : 0;   // only a stage that a later stage polls needs strong lattice stores
            tp.dbg = mmh_getenv("MMH_ROWS_DBG") ? atoi(mmh_getenv("MMH_ROWS_DBG")) : 0;
            if (tp.dbg & 2) tp.strong_g = 1;
            {   // exchange buffer: grow-only scratch, (re)filled with the sentinel whenever it is (re)allocated
                const size_t xoff = xbase + ((overlap && i == i0) ? xoff1 : 0);
                const size_t xbytes = xoff + sizeof(c128) * (size_t)ntiles * d.shape[i] * tp.hc_max;
                if (ctx->xbuf.bytes < xbytes || !ctx->xbuf.ptr) {
                    if ((rc = ensure_scratch(ctx->xbuf, xbytes))) return rc;
                    CK(cudaMemsetAsync(ctx->xbuf.ptr, 0xFF, ctx->xbuf.bytes, st));
                }
                tp.X = (c128 *)((char *)ctx->xbuf.ptr + xoff);
            }
            const char *trace_file = mmh_getenv("MMH_TRACE_FILE");   // debug timeline of the tile pipeline
            const size_t trace_words = (size_t)ntiles * d.shape[i] * 8;
            if (trace_file) {
                if (trace_used + trace_words > kTraceArenaWords) return MMH_ERR_TOO_LARGE;
                tp.trace = trace_arena + trace_used;
                trace_used += trace_words;
            }
            g_launches++;
            if (tp.rows_R) CK(mmh_launch_march_rows(tp, ntiles, sm, st));
            else {
                // small tile grids (stage 1 of a 4-index lattice: 12-16 tiles) can run as ONE thread-block cluster: halos are pushed
                // into the consumer's shared memory (mmh_tiled.cu, CL variant) instead of through L2
                const size_t smc = sm + mmh_tiled2_cluster_extra_smem(tp.hc_max, d.shape[i]);
                // (opt-in, MMH_CLUSTER=1: bit-identical and measured NEUTRAL on cfg2 -- 119.0 vs 119.4 us, back-to-back 100.8 vs 100.8 us
                //  per lattice: the lag of stage 1's far tiles comes from the rows of plane k1 = 0 that the one-warp tail kernel
                //  delivers last, not from the hop through L2)
                tp.cluster = (ntiles > 1 && ntiles <= 16 && d.D - 1 - i <= 3 && smc <= 200 * 1024 && mmh_getenv("MMH_CLUSTER")) ? 1 : 0;
                cudaError_t ce = cudaSuccess;
                if (tp.cluster) {
                    ce = mmh_launch_march_tiled2(tp, R, ntiles, smc, st);
                    if (mmh_getenv("MMH_DEBUG_CLUSTER")) fprintf(stderr, "[mmh] stage %d: cluster launch of %d tiles, %zu B smem: %s\n", i, ntiles, smc, cudaGetErrorString(ce));
                    if (ce != cudaSuccess) { (void)cudaGetLastError(); tp.cluster = 0; }   // no GPC can host the cluster: through L2
                }
                if (!tp.cluster) CK(mmh_launch_march_tiled2(tp, R, ntiles, sm, st));
            }
            if (trace_file) {   // read back after the whole lattice has been enqueued (the stages overlap)
                PendingTrace pt;
                pt.dev = tp.trace; pt.words = trace_words; pt.stage = i;
                pt.hdr[0] = ntiles; pt.hdr[1] = d.shape[i]; pt.hdr[2] = tp.g[0]; pt.hdr[3] = tp.g[1]; pt.hdr[4] = tp.g[2];
                pt.hdr[5] = R; pt.hdr[6] = tp.tc; pt.hdr[7] = 0;
                traces.push_back(pt);
            }
        } else {
            const long long P = d.strides[i];
            long long grid = (P + 255) / 256;
            if (grid > 8LL * ctx->sm_count) grid = 8LL * ctx->sm_count;
            for (int s = 1; s < d.shape[i]; s++) {
                g_launches++;
                CK(mmh_launch_panel_step(p, i, s, 0, P, (int)grid, absmem, st, true));   // consecutive steps of one lattice: chained with PDL
            }
            first = false;
        }
    }
    if (!chain_done) {   // D == 1 or every march stage has extent 1: the chain is the whole lattice
        g_launches++;
        CK(mmh_launch_chain(p, st));
    }
    if (!traces.empty()) {
        CK(cudaStreamSynchronize(st));
        for (const PendingTrace &pt : traces) {
            std::vector<unsigned long long> h(pt.words);
            CK(cudaMemcpy(h.data(), pt.dev, pt.words * 8, cudaMemcpyDeviceToHost));
            char name[512];
            snprintf(name, sizeof(name), "%s.stage%d.bin", mmh_getenv("MMH_TRACE_FILE"), pt.stage);
            if (FILE *fp = fopen(name, "wb")) {
                fwrite(pt.hdr, sizeof(int), 8, fp);
                fwrite(h.data(), 8, pt.words, fp);
                fclose(fp);
            }
        }
    }
    if (trace_arena) { CK(cudaStreamSynchronize(st)); CK(cudaFree(trace_arena)); }
    return MMH_OK;
}


// ---- stable rule on a wavefront of boxes (mmh_stable_boxes.cu): per-shape tables, cached on the device ---------------------------
struct StableBoxPlan { int *box_order; unsigned *cell_order; int *lvl_start; int nb[4]; int E; int nbox; int ncell; int nlev; };
static int stable_box_plan(const LatticeDesc &d, int device, StableBoxPlan **out) {
    static std::map<std::string, StableBoxPlan> cache;
    std::string key((const char *)d.shape, sizeof(int) * (size_t)d.D);
    key.push_back((char)d.D); key.append(std::to_string(device));
    auto it = cache.find(key);
    if (it == cache.end()) {
        StableBoxPlan pl;
        memset(&pl, 0, sizeof(pl));
        const int D = d.D, pad = 4 - D, E = mmh_stable_boxes_edge(D);
        pl.E = E;
        int nbox = 1;
        for (int j = 0; j < 4; j++) { pl.nb[j] = j < pad ? 1 : (d.shape[j - pad] + E - 1) / E; nbox *= pl.nb[j]; }
        pl.nbox = nbox;
        // boxes by level (sum of box coordinates), lexicographic inside a level: every box comes after all boxes below it
        std::vector<std::pair<int, int>> lv((size_t)nbox);
        for (int b = 0; b < nbox; b++) {
            int r = b, s = 0;
            for (int j = 3; j >= 0; j--) { s += r % pl.nb[j]; r /= pl.nb[j]; }
            lv[(size_t)b] = std::make_pair(s, b);
        }
        std::sort(lv.begin(), lv.end());
        std::vector<int> order((size_t)nbox);
        for (int n = 0; n < nbox; n++) order[(size_t)n] = lv[(size_t)n].second;
        // cells of a full box by local level
        int ncell = 1;
        for (int j = 0; j < D; j++) ncell *= E;
        const int nlev = D * (E - 1) + 1;
        std::vector<std::vector<unsigned>> by((size_t)nlev);
        for (int c = 0; c < ncell; c++) {
            int r = c, s = 0;
            unsigned packed = 0;
            for (int j = 3; j >= pad; j--) { const int x = r % E; r /= E; s += x; packed |= (unsigned)x << (8 * j); }
            by[(size_t)s].push_back(packed);
        }
        std::vector<unsigned> cells;
        std::vector<int> starts((size_t)nlev + 1);
        for (int m = 0; m < nlev; m++) { starts[(size_t)m] = (int)cells.size(); cells.insert(cells.end(), by[(size_t)m].begin(), by[(size_t)m].end()); }
        starts[(size_t)nlev] = (int)cells.size();
        pl.ncell = ncell; pl.nlev = nlev;
        CK(cudaMalloc(&pl.box_order, sizeof(int) * (size_t)nbox));
        CK(cudaMalloc(&pl.cell_order, sizeof(unsigned) * cells.size()));
        CK(cudaMalloc(&pl.lvl_start, sizeof(int) * starts.size()));
        CK(cudaMemcpy(pl.box_order, order.data(), sizeof(int) * (size_t)nbox, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(pl.cell_order, cells.data(), sizeof(unsigned) * cells.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(pl.lvl_start, starts.data(), sizeof(int) * starts.size(), cudaMemcpyHostToDevice));
        it = cache.emplace(key, pl).first;
    }
    *out = &it->second;
    return MMH_OK;
}

static int forward_stable_boxes(const FwdParams &p, DeviceCtx *ctx, int device, cudaStream_t st) {
    StableBoxPlan *pl;
    int rc;
    if ((rc = stable_box_plan(p.d, device, &pl))) return rc;
    if ((rc = ensure_scratch(ctx->sbox_ws, sizeof(int) * ((size_t)pl->nbox + 16)))) return rc;
    CK(cudaMemsetAsync(ctx->sbox_ws.ptr, 0, sizeof(int) * ((size_t)pl->nbox + 16), st));
    StableBoxParams q;
    memset(&q, 0, sizeof(q));
    q.d = p.d; q.A = p.A; q.b = p.b; q.c = p.c; q.G = p.G; q.sq = p.sq; q.rsq = p.rsq;
    for (int j = 0; j < 4; j++) q.nb[j] = pl->nb[j];
    q.E = pl->E; q.nbox = pl->nbox; q.ncell = pl->ncell; q.nlev = pl->nlev;
    int mxs = 0;
    for (int j = 0; j < p.d.D; j++) mxs = p.d.shape[j] > mxs ? p.d.shape[j] : mxs;
    q.ntab = mxs + 1;
    q.xshift = 0;
    while ((1 << q.xshift) < pl->E + 2) q.xshift++;
    q.box_order = pl->box_order; q.cell_order = pl->cell_order; q.lvl_start = pl->lvl_start;
    q.ticket = (int *)ctx->sbox_ws.ptr;
    q.flags = (int *)ctx->sbox_ws.ptr + 16;
    q.err = ctx->err_dev;
    const char *trace_file = mmh_getenv("MMH_SB_TRACE");
    if (trace_file) { CK(cudaMalloc(&q.trace, 64 * 8 * 8)); CK(cudaMemset(q.trace, 0, 64 * 8 * 8)); }
    g_launches++;
    CK(mmh_launch_stable_boxes(q, ctx->sm_count, st));
    if (trace_file) {
        unsigned long long h[64 * 8];
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(h, q.trace, sizeof(h), cudaMemcpyDeviceToHost));
        CK(cudaFree(q.trace));
        if (FILE *fp = fopen(trace_file, "w")) {
            for (int n = 0; n < 64 && h[n * 8]; n++)
                fprintf(fp, "box %5llu  ticket %.2f  deps %.2f  halo %.2f  levels %.2f  publish %.2f   (us; start %.2f)\n", h[n * 8 + 6],
                        (h[n * 8 + 1] - h[n * 8]) / 1e3, (h[n * 8 + 2] - h[n * 8 + 1]) / 1e3, (h[n * 8 + 3] - h[n * 8 + 2]) / 1e3,
                        (h[n * 8 + 4] - h[n * 8 + 3]) / 1e3, (h[n * 8 + 5] - h[n * 8 + 4]) / 1e3, (h[n * 8] - h[0]) / 1e3);
            fclose(fp);
        }
    }
    return MMH_OK;
}

static int forward_impl(long long batch, int ndim, const int64_t *shape, const void *dA, const void *db,
                        const void *dc, void *dG, int stable, cudaStream_t st) {
    if (batch < 0) return MMH_ERR_BAD_BATCH;
    LatticeDesc d;
    int mx;
    int rc = make_desc(ndim, shape, &d, &mx);
    if (rc) return rc;
    if (batch == 0) return MMH_OK;
    if (!dA || !db || !dc || !dG) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    if ((rc = get_ctx(&ctx))) return rc;
    if ((rc = begin_call(ctx, st))) return rc;
    if ((rc = ensure_tables(*ctx, mx + 1))) return rc;

    FwdParams p;
    p.d = d;
    p.A = (const c128 *)dA; p.b = (const c128 *)db; p.c = (const c128 *)dc; p.G = (c128 *)dG;
    p.sq = ctx->sq; p.rsq = ctx->rsq;
    p.batch = batch; p.barrier = nullptr; p.small_stage_lo = 0;
    const size_t smem = sizeof(c128) * (size_t)(ndim * ndim + ndim);

    // Batches of lattices too large for the one-CTA shared-memory march: either every lattice in turn on the whole device
    // (tiled, pipelined) or one CTA per lattice -- the box march (mmh_box.cu,
    // ~1.1 ns per amplitude and CTA, one CTA per SM; (20,)^4: 0.18 ms per wave of 148 lattices) or, for panels of more than four
    // dims, the L1/L2 kernel (~3 ns per amplitude and CTA, two CTAs per SM).  One CTA per lattice wins from a handful of lattices
    // on (32 x (30,)^4: 0.77 vs 1.78 ms; 8 x (30,)^4: 0.73 vs 0.45 ms).
    long long per_cta_batch = 2LL * ctx->sm_count;
    if (!stable && d.N > kSingleCtaN && batch > 1) {
        const bool boxed = ndim <= 5 && !mmh_getenv("MMH_NO_BOX");
        const double rate = boxed ? 1.1e-3 : 3.0e-3;   // us per amplitude and CTA
        const long long slots = (boxed ? 1LL : 2LL) * ctx->sm_count;
        long long steps = 0;
        for (int i = 0; i < ndim; i++) steps += d.shape[i] - 1;
        // per lattice, pipelined (measured: (20,)^4 39 us, (30,)^4 56 us, (64,)^3 52 us, (50,)^4 85 us)
        const double t_pipe = (double)batch * (25.0 + 0.1 * (double)steps + 8.0e-6 * (double)d.N);
        const double waves = (double)((batch + slots - 1) / slots);
        const double t_cta = waves * rate * (double)d.N;
        if (t_cta < t_pipe) per_cta_batch = batch;
    }
    if (const char *e = mmh_getenv("MMH_PER_CTA_BATCH")) per_cta_batch = atoll(e);   // tuning hook
    const bool per_cta = (d.N <= kSingleCtaN && !mmh_getenv("MMH_FORCE_TILED")) || batch >= per_cta_batch;
    if (!stable && (ndim == 1 || per_cta)) {
        bool done = false;
        if ((rc = forward_staged(p, ctx, st, &done))) return rc;
        if (done) return MMH_OK;
    }
    if (per_cta) {
        long long maxpanel = stable ? d.N / d.shape[ndim - 1] : d.strides[0];
        int block = round_up32(maxpanel);
        if (block < 32) block = 32;
        if (block > 256) block = 256;
        long long grid = batch;
        if (grid > 0x7fffffffLL) grid = 0x7fffffffLL;
        g_launches++;
        CK(mmh_launch_fwd_cta(p, stable != 0, (int)grid, block, smem, st));
        return MMH_OK;
    }
    if (!stable && ndim >= 2 && ndim <= 8 && !mmh_getenv("MMH_FORCE_COOP")) {
        // large lattices, vanilla rule: chain + per-stage march kernels, every SM on the same lattice
        for (long long l = 0; l < batch; l++) {
            FwdParams q = p;
            q.A = p.A + l * ndim * ndim; q.b = p.b + l * ndim; q.c = p.c + l; q.G = p.G + l * d.N;
            q.batch = 1;
            if ((rc = forward_single_staged(q, ctx, st, l))) return rc;
        }
        return MMH_OK;
    }
    // stable rule, 2..4 indices, a lattice of many boxes: wavefront of boxes (no grid barrier, neighbours from shared memory)
    if (stable && ndim >= 2 && ndim <= 4 && mx <= 4096 &&
        d.N >= (mmh_getenv("MMH_STABLE_BOXES_MIN_N") ? atoll(mmh_getenv("MMH_STABLE_BOXES_MIN_N")) : 200000LL) && !mmh_getenv("MMH_NO_STABLE_BOXES")) {
        int device = 0;
        CK(cudaGetDevice(&device));
        for (long long l = 0; l < batch; l++) {
            FwdParams q = p;
            q.A = p.A + l * ndim * ndim; q.b = p.b + l * ndim; q.c = p.c + l; q.G = p.G + l * d.N;
            q.batch = 1;
            if ((rc = forward_stable_boxes(q, ctx, device, st))) return rc;
        }
        return MMH_OK;
    }
    // stable rule (level wavefront) and D > 8: one cooperative launch per lattice
    const int block = 256;
    int per_sm = 0;
    CK(mmh_coop_max_blocks(stable != 0, block, smem, &per_sm));
    if (per_sm < 1) return MMH_ERR_UNSUPPORTED;
    if (per_sm > 2) per_sm = 2;
    const int grid = ctx->sm_count * per_sm;
    int lo = ndim;  // stages >= lo are filled by CTA 0 alone
    while (lo > 0 && d.strides[lo - 1] <= kSmallPanel) lo--;
    if (lo == 0) lo = 1;  // keep at least one grid-wide stage so that every CTA has work
    for (long long l = 0; l < batch; l++) {
        FwdParams q = p;
        q.A = p.A + l * ndim * ndim; q.b = p.b + l * ndim; q.c = p.c + l; q.G = p.G + l * d.N;
        q.batch = 1;
        q.small_stage_lo = lo;
        q.barrier = ctx->barrier + ctx->barrier_next;
        ctx->barrier_next = (ctx->barrier_next + 1) % 64;
        CK(cudaMemsetAsync(q.barrier, 0, sizeof(unsigned), st));
        g_launches++;
        CK(mmh_launch_fwd_coop(q, stable != 0, grid, block, smem, st));
    }
    return MMH_OK;
}

// ---- vjp ------------------------------------------------------------------------------------------
static int vjp_impl(long long batch, int ndim, const int64_t *shape, const void *dG, const void *dc,
                    const void *dg, void *oA, void *ob, void *oc, cudaStream_t st) {
    if (batch < 0) return MMH_ERR_BAD_BATCH;
    LatticeDesc d;
    int mx;
    int rc = make_desc(ndim, shape, &d, &mx);
    if (rc) return rc;
    if (batch == 0) return MMH_OK;
    if (!dG || !dc || !dg || !oA || !ob || !oc) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    if ((rc = get_ctx(&ctx))) return rc;
    if ((rc = begin_call(ctx, st))) return rc;
    if ((rc = ensure_tables(*ctx, mx + 1))) return rc;
    VjpParams p;
    p.d = d;
    p.G = (const c128 *)dG; p.g = (const c128 *)dg; p.c = (const c128 *)dc;
    p.dA = (c128 *)oA; p.db = (c128 *)ob; p.dc = (c128 *)oc;
    p.sq = ctx->sq;
    p.batch = batch;
    p.nacc = ndim + ndim * (ndim + 1) / 2 + 1;
    // batches of 2-index lattices: warp-synchronous row walk, every amplitude loaded once (k_vjp_lanes)
    {
        int Rl, ln, Lw;
        if (ndim == 2 && batch >= 256 && d.shape[0] > 1 && !mmh_getenv("MMH_NO_LANES") && mmh_plan_march_lanes(d.shape[1], &Rl, &ln, &Lw)) {
            g_launches++;
            CK(mmh_launch_vjp_lanes(p, Rl, ln, Lw, ctx->sm_count, st));
            return MMH_OK;
        }
    }
    // one (or a few) large 4-index lattices: plane tiles staged by the TMA engine (k_vjp_planes)
    const long long planes_min_n = mmh_getenv("MMH_VJP_PLANES_MIN_N") ? atoll(mmh_getenv("MMH_VJP_PLANES_MIN_N")) : (1LL << 20);   // test hook
    if (ndim == 4 && batch <= 8 && d.N >= planes_min_n && mmh_vjp_planes_smem(d) && !mmh_getenv("MMH_NO_VJP_PLANES")) {
        if ((rc = ensure_scratch(ctx->partial, sizeof(c128) * (size_t)batch * ctx->sm_count * p.nacc))) return rc;
        p.partial = (c128 *)ctx->partial.ptr;
        p.nblk = ctx->sm_count;
        int nblk = 0;
        g_launches += 2;
        CK(mmh_launch_vjp_planes(p, ctx->sm_count, &nblk, st));
        return MMH_OK;
    }
    // batched small lattices: few warps per lattice (long per-thread walks amortise the reduction); one large lattice:
    // 256-thread CTAs, ~8 per SM
    int block = 256;
    if (batch >= 4LL * ctx->sm_count && d.N <= 8192) block = d.N >= 4096 ? 128 : 64;
    else block = 128;
    if (const char *e = mmh_getenv("MMH_VJP_BLOCK")) block = atoi(e);
    long long want = (d.N + block * 4 - 1) / (block * 4);   // ~4 points per thread
    // one wave of resident CTAs over the whole batch (at most 4 per SM)
    int per_sm = mmh_vjp_blocks_per_sm(p, block);
    if (per_sm > 4) per_sm = 4;
    long long cap = ((long long)per_sm * ctx->sm_count + batch - 1) / batch;
    if (const char *e = mmh_getenv("MMH_VJP_CTAS_PER_SM")) cap = ((long long)atoi(e) * ctx->sm_count + batch - 1) / batch;
    if (cap < 1) cap = 1;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    p.nblk = (int)want;
    if ((rc = ensure_scratch(ctx->partial, sizeof(c128) * (size_t)batch * p.nblk * p.nacc))) return rc;
    p.partial = (c128 *)ctx->partial.ptr;
    g_launches += 2;
    CK(mmh_launch_vjp(p, p.nblk, block, st));
    return MMH_OK;
}

// ---- binomial -------------------------------------------------------------------------------------
static int binomial_impl(int ndim, const int64_t *shape, const void *dA, const void *db, const void *dc,
                         double max_l2, long long global_cutoff, void *dG, double *norm_out, cudaStream_t st) {
    LatticeDesc d;
    int mx;
    int rc = make_desc(ndim, shape, &d, &mx);
    if (rc) return rc;
    if (!dA || !db || !dc || !dG) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    if ((rc = get_ctx(&ctx))) return rc;
    if ((rc = begin_call(ctx, st))) return rc;
    if ((rc = ensure_tables(*ctx, mx + 1))) return rc;
    if ((rc = ensure_scratch(ctx->norm, sizeof(double)))) return rc;
    long long maxlevel = 0;
    for (int i = 0; i < ndim; i++) maxlevel += d.shape[i] - 1;
    BinomParams p;
    p.d = d;
    p.A = (const c128 *)dA; p.b = (const c128 *)db; p.c = (const c128 *)dc; p.G = (c128 *)dG;
    p.sq = ctx->sq; p.rsq = ctx->rsq;
    p.max_l2 = max_l2;
    p.global_cutoff = global_cutoff < maxlevel + 1 ? global_cutoff : maxlevel + 1;  // empty levels add 0
    p.norm_out = (double *)ctx->norm.ptr;
    CK(cudaMemsetAsync(dG, 0, sizeof(c128) * (size_t)d.N, st));  // np.zeros (binomial.py:51)
    long long Q = d.N / d.shape[ndim - 1];
    int block = round_up32(Q);
    if (block < 32) block = 32;
    if (block > 1024) block = 1024;
    const size_t smem = sizeof(c128) * (size_t)(ndim * ndim + ndim) + sizeof(double) * 40;
    g_launches++;
    CK(mmh_launch_binomial(p, block, smem, st));
    double norm = 0.0;
    CK(cudaMemcpyAsync(&norm, p.norm_out, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (norm_out) *norm_out = norm;
    return MMH_OK;
}

// ---- compactFock diagonal / one leftover mode ----------------------------------------------------------
static int stage_in(DeviceCtx &c, int slot, const void *host, size_t bytes, void **dev);
static int diagonal_impl(int M, const int64_t *cutoffs, int L0, const void *dA, const void *dB, long long nbatch,
                         const void *dG0, void *dout, cudaStream_t st) {
    if (!cutoffs) return MMH_ERR_NULL_POINTER;
    const int Md = M - L0;
    if (M < 1 + L0 || Md < 1 || Md > 8) return MMH_ERR_BAD_NDIM;
    if (nbatch < 0) return MMH_ERR_BAD_BATCH;
    if (!dA || !dB || !dG0 || !dout) return MMH_ERR_NULL_POINTER;
    DiagParams q;
    memset(&q, 0, sizeof(q));
    q.Md = Md; q.L0 = L0;
    q.c0 = L0 ? (int)cutoffs[0] : 1;
    q.nb = nbatch > 0 ? (int)nbatch : 1;
    int mx = q.c0, nlevels = 1;
    long long P = 1;
    for (int j = 0; j < M; j++) if (cutoffs[j] < 1 || cutoffs[j] > (1 << 20)) return MMH_ERR_BAD_SHAPE;
    for (int j = 0; j < Md; j++) {
        q.cut[j] = (int)cutoffs[j + L0];
        if (q.cut[j] > mx) mx = q.cut[j];
        nlevels += q.cut[j] - 1;
        if (P > (1LL << 40) / q.cut[j]) return MMH_ERR_TOO_LARGE;
        P *= q.cut[j];
    }
    q.pst[Md - 1] = 1;
    for (int j = Md - 1; j > 0; j--) q.pst[j - 1] = q.pst[j] * q.cut[j];
    q.P = P;
    q.E = (long long)q.c0 * q.c0 * P * q.nb;
    DeviceCtx *ctx;
    int rc;
    if ((rc = get_ctx(&ctx))) return rc;
    if ((rc = begin_call(ctx, st))) return rc;
    if ((rc = ensure_tables(*ctx, mx + 3))) return rc;
    const long long naux = 2LL * Md + Md + 2LL * Md * (Md > 1 ? Md - 1 : 1);
    const size_t bytes = sizeof(c128) * (size_t)naux * (size_t)q.E;
    // Large sweeps of the pure diagonal case: rolling weight-level buffers (mmh_diagonal_rolling.cu) -- the reference layout of the
    // auxiliary arrays is 0.94 TB for the 8-mode, cutoff-12 config; two level buffers are 18 GB.  MMH_DIAG_ROLLING=0/1 forces a path.
    {
        const char *er = mmh_getenv("MMH_DIAG_ROLLING");
        const bool rolling = !L0 && (er ? atoi(er) != 0 : bytes > ((size_t)1 << 30));
        if (rolling) {
            const size_t ws = mmh_diagonal_rolling_workspace(Md, q.cut, q.nb);
            if (ws > (size_t)150 << 30) return MMH_ERR_TOO_LARGE;
            if ((rc = ensure_scratch(ctx->diag_ws, ws))) return rc;
            long long launches = 0;
            CK(mmh_launch_diagonal_rolling(Md, q.cut, q.nb, (const c128 *)dA, (const c128 *)dB, (const c128 *)dG0, (c128 *)dout,
                                           ctx->sq, ctx->diag_ws.ptr, &launches, st));
            g_launches += launches;
            return MMH_OK;
        }
    }
    if (bytes > (size_t)160 << 30) return MMH_ERR_TOO_LARGE;
    if ((rc = ensure_scratch(ctx->diag_ws, bytes))) return rc;
    CK(cudaMemsetAsync(ctx->diag_ws.ptr, 0, bytes, st));
    c128 *w = (c128 *)ctx->diag_ws.ptr;
    q.arr1 = w;                      w += 2LL * Md * q.E;
    q.arr2 = w;                      w += (long long)Md * q.E;
    q.arr1010 = w;                   w += (long long)Md * (Md > 1 ? Md - 1 : 1) * q.E;
    q.arr1001 = w;
    q.arr0 = (c128 *)dout;
    q.A = (const c128 *)dA; q.B = (const c128 *)dB; q.sq = ctx->sq;
    long long launches = 0;
    CK(mmh_launch_diagonal(q, (const c128 *)dG0, nlevels, &launches, st));
    g_launches += launches;
    return MMH_OK;
}

// Jacobian kernels need the value arrays -> run the value sweep into scratch, then the tangent sweep
__global__ void k_diag_split_jacobian(const c128 *t0, const c128 *arr0, const c128 *G0, long long P, int n2,
                                      c128 *dG0_out, c128 *dA_out, c128 *dB_out) {
    const int nt = n2 * n2 + n2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= P * (nt + 1)) return;
    const long long p = gid / (nt + 1);
    const int th = (int)(gid - p * (nt + 1));
    if (th == nt) {   // arr0_dG0 = arr0 / G0 (inputValidation.py:99)
        const c128 a = arr0[p], b = G0[0];
        const double den = b.x * b.x + b.y * b.y;
        dG0_out[p] = make_double2((a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den);
    } else if (th < n2 * n2) dA_out[p * n2 * n2 + th] = t0[p * nt + th];
    else dB_out[p * n2 + (th - n2 * n2)] = t0[p * nt + th];
}

static int diagonal_grad_impl(int Mtot, const int64_t *cutoffs, int L0, const void *dA, const void *dB, const void *dG0,
                              void *o_dG0, void *o_dA, void *o_dB, cudaStream_t st) {
    if (!cutoffs) return MMH_ERR_NULL_POINTER;
    const int M = Mtot - L0;   // detected modes
    if (M < 1 || M > 8) return MMH_ERR_BAD_NDIM;
    if (!dA || !dB || !dG0 || !o_dG0 || !o_dA || !o_dB) return MMH_ERR_NULL_POINTER;
    DiagTanParams tp;
    memset(&tp, 0, sizeof(tp));
    DiagParams &q = tp.q;
    q.Md = M; q.L0 = L0; q.nb = 1;
    q.c0 = L0 ? (int)cutoffs[0] : 1;
    if (q.c0 < 1 || q.c0 > (1 << 20)) return MMH_ERR_BAD_SHAPE;
    int mx = q.c0, nlevels = 1;
    long long P = 1;
    for (int j = 0; j < M; j++) {
        if (cutoffs[j + L0] < 1 || cutoffs[j + L0] > (1 << 20)) return MMH_ERR_BAD_SHAPE;
        q.cut[j] = (int)cutoffs[j + L0];
        if (q.cut[j] > mx) mx = q.cut[j];
        nlevels += q.cut[j] - 1;
        if (P > (1LL << 34) / q.cut[j]) return MMH_ERR_TOO_LARGE;
        P *= q.cut[j];
    }
    q.pst[M - 1] = 1;
    for (int j = M - 1; j > 0; j--) q.pst[j - 1] = q.pst[j] * q.cut[j];
    q.P = P; q.E = (long long)q.c0 * q.c0 * P;
    const int n2 = 2 * Mtot;
    const long long Pv = q.E;   // amplitudes of arr0
    tp.ntheta = n2 * n2 + n2;
    DeviceCtx *ctx;
    int rc;
    if ((rc = get_ctx(&ctx))) return rc;
    if ((rc = begin_call(ctx, st))) return rc;
    if ((rc = ensure_tables(*ctx, mx + 3))) return rc;
    const long long naux = 2LL * M + M + 2LL * M * (M > 1 ? M - 1 : 1);   // arr1, arr2, arr1010, arr1001
    const size_t nval = (size_t)(naux + 1) * (size_t)Pv;
    const size_t bytes = sizeof(c128) * nval * (size_t)(1 + tp.ntheta);
    if (bytes > (size_t)120 << 30) return MMH_ERR_TOO_LARGE;
    if ((rc = ensure_scratch(ctx->diag_ws, bytes))) return rc;
    CK(cudaMemsetAsync(ctx->diag_ws.ptr, 0, bytes, st));
    c128 *w = (c128 *)ctx->diag_ws.ptr;
    q.arr0 = w;    w += Pv;
    q.arr1 = w;    w += 2LL * M * Pv;
    q.arr2 = w;    w += (long long)M * Pv;
    q.arr1010 = w; w += (long long)M * (M > 1 ? M - 1 : 1) * Pv;
    q.arr1001 = w; w += (long long)M * (M > 1 ? M - 1 : 1) * Pv;
    const long long nt = tp.ntheta;
    tp.t0 = w;     w += Pv * nt;
    tp.t1 = w;     w += 2LL * M * Pv * nt;
    tp.t2 = w;     w += (long long)M * Pv * nt;
    tp.t1010 = w;  w += (long long)M * (M > 1 ? M - 1 : 1) * Pv * nt;
    tp.t1001 = w;
    q.A = (const c128 *)dA; q.B = (const c128 *)dB; q.sq = ctx->sq;
    long long launches = 0;
    CK(mmh_launch_diagonal(q, (const c128 *)dG0, nlevels, &launches, st));
    g_launches += launches;
    CK(mmh_launch_diagonal_tangent(tp, nlevels, &launches, st));
    g_launches += launches + 1;
    const long long total = Pv * (nt + 1);
    k_diag_split_jacobian<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(tp.t0, q.arr0, (const c128 *)dG0, Pv, n2,
                                                                         (c128 *)o_dG0, (c128 *)o_dA, (c128 *)o_dB);
    CK(cudaGetLastError());
    return MMH_OK;
}

static int diagonal_host_impl(int M, const int64_t *cutoffs, int L0, const void *A, const void *B, long long nbatch,
                              const void *G0, void *out) {
    if (!cutoffs) return MMH_ERR_NULL_POINTER;
    if (M < 1 + L0 || M - L0 > 8) return MMH_ERR_BAD_NDIM;
    if (nbatch < 0) return MMH_ERR_BAD_BATCH;
    if (!A || !B || !G0 || !out) return MMH_ERR_NULL_POINTER;
    size_t n = nbatch > 0 ? (size_t)nbatch : 1;
    for (int j = 0; j < M; j++) {
        if (cutoffs[j] < 1) return MMH_ERR_BAD_SHAPE;
        n *= (size_t)cutoffs[j] * ((L0 && j == 0) ? (size_t)cutoffs[j] : 1);
    }
    DeviceCtx *ctx;
    int rc;
    if ((rc = get_ctx(&ctx))) return rc;
    const size_t n2 = 2 * (size_t)M;
    void *dA, *dB, *dG0, *dout;
    if ((rc = stage_in(*ctx, 0, A, sizeof(c128) * n2 * n2, &dA))) return rc;
    if ((rc = stage_in(*ctx, 1, B, sizeof(c128) * n2 * (nbatch > 0 ? (size_t)nbatch : 1), &dB))) return rc;
    if ((rc = stage_in(*ctx, 2, G0, sizeof(c128), &dG0))) return rc;
    if ((rc = stage_in(*ctx, 3, nullptr, sizeof(c128) * n, &dout))) return rc;
    if ((rc = diagonal_impl(M, cutoffs, L0, dA, dB, nbatch, dG0, dout, 0))) return rc;
    CK(cudaMemcpyAsync(out, dout, sizeof(c128) * n, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return end_host_call(ctx);
}

// ---- host staging ---------------------------------------------------------------------------------
static int stage_in(DeviceCtx &c, int slot, const void *host, size_t bytes, void **dev) {
    int rc = ensure_scratch(c.host_slots[slot], bytes);
    if (rc) return rc;
    *dev = c.host_slots[slot].ptr;
    if (host && bytes) CK(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, 0));
    return 0;
}

extern "C" {

int mmh_version(void) { return 100; }

const char *mmh_error_string(int status) {
    switch (status) {
        case MMH_OK: return "ok";
        case MMH_ERR_BAD_NDIM: return "ndim must be in [1, 32]";
        case MMH_ERR_BAD_SHAPE: return "every entry of shape must be >= 1 (and < 2^24)";
        case MMH_ERR_NULL_POINTER: return "required pointer is NULL";
        case MMH_ERR_BAD_BATCH: return "batch must be >= 0";
        case MMH_ERR_UNSUPPORTED: return "not supported by the CUDA path";
        case MMH_ERR_NO_DEVICE: return "no usable CUDA device (an sm_100 GPU is required; there is no CPU fallback)";
        case MMH_ERR_TOO_LARGE: return "lattice too large";
        case MMH_ERR_TIMEOUT: return "a device-side watchdog of an earlier call expired (that call's result is invalid); state reset, retry";
        default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown error";
    }
}

int mmh_device_count(int *count_out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (count_out) *count_out = n;
    return MMH_OK;
}

int mmh_set_device(int device) { CK(cudaSetDevice(device)); return MMH_OK; }
int mmh_device_synchronize(void) { CK(cudaDeviceSynchronize()); return MMH_OK; }
int64_t mmh_launch_count(void) { return g_launches.load(); }

int mmh_host_alloc(void **ptr_out, int64_t bytes) {
    if (!ptr_out) return MMH_ERR_NULL_POINTER;
    CK(cudaHostAlloc(ptr_out, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocDefault));
    return MMH_OK;
}
int mmh_host_free(void *ptr) { if (ptr) CK(cudaFreeHost(ptr)); return MMH_OK; }

int mmh_forward(int ndim, const int64_t *shape, const void *dA, const void *db, const void *dc, void *dG,
                int stable, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return forward_impl(1, ndim, shape, dA, db, dc, dG, stable, (cudaStream_t)stream);
}

int mmh_forward_panel_range(int ndim, const int64_t *shape, const void *dA, const void *db, void *dG, int stage,
                            int64_t step, int64_t f_lo, int64_t f_hi, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    LatticeDesc d;
    int mx;
    int rc = make_desc(ndim, shape, &d, &mx);
    if (rc) return rc;
    if (!dA || !db || !dG) return MMH_ERR_NULL_POINTER;
    if (stage < 0 || stage >= ndim || step < 1 || step >= d.shape[stage]) return MMH_ERR_BAD_SHAPE;
    if (f_lo < 0 || f_hi > d.strides[stage] || f_lo > f_hi) return MMH_ERR_BAD_SHAPE;
    if (f_lo == f_hi) return MMH_OK;
    DeviceCtx *ctx;
    if ((rc = get_ctx(&ctx))) return rc;
    if ((rc = begin_call(ctx, (cudaStream_t)stream))) return rc;
    if ((rc = ensure_tables(*ctx, mx + 1))) return rc;
    FwdParams p;
    memset(&p, 0, sizeof(p));
    p.d = d;
    p.A = (const c128 *)dA; p.b = (const c128 *)db; p.c = nullptr; p.G = (c128 *)dG;
    p.sq = ctx->sq; p.rsq = ctx->rsq; p.batch = 1;
    long long grid = (f_hi - f_lo + 255) / 256;
    if (grid > 8LL * ctx->sm_count) grid = 8LL * ctx->sm_count;
    g_launches++;
    // plain launch: the caller orders panel ranges with events across streams / ranks (sharding.SingleLatticePlan, also under graph capture)
    CK(mmh_launch_panel_step(p, stage, (int)step, f_lo, f_hi, (int)grid, sizeof(c128) * (size_t)(ndim * ndim + ndim),
                             (cudaStream_t)stream, false));
    return MMH_OK;
}

int mmh_forward_batched(int64_t batch, int ndim, const int64_t *shape, const void *dA, const void *db,
                        const void *dc, void *dG, int stable, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return forward_impl(batch, ndim, shape, dA, db, dc, dG, stable, (cudaStream_t)stream);
}

int mmh_forward_batched_host(int64_t batch, int ndim, const int64_t *shape, const void *A, const void *b,
                             const void *c, void *G, int stable) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (batch < 0) return MMH_ERR_BAD_BATCH;
    LatticeDesc d; int mx;
    int rc = make_desc(ndim, shape, &d, &mx);
    if (rc) return rc;
    if (batch == 0) return MMH_OK;
    if (!A || !b || !c || !G) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    if ((rc = get_ctx(&ctx))) return rc;
    void *dA, *db, *dc, *dG;
    const size_t D = ndim;
    if ((rc = stage_in(*ctx, 0, A, sizeof(c128) * batch * D * D, &dA))) return rc;
    if ((rc = stage_in(*ctx, 1, b, sizeof(c128) * batch * D, &db))) return rc;
    if ((rc = stage_in(*ctx, 2, c, sizeof(c128) * batch, &dc))) return rc;
    if ((rc = stage_in(*ctx, 3, nullptr, sizeof(c128) * batch * (size_t)d.N, &dG))) return rc;
    if ((rc = forward_impl(batch, ndim, shape, dA, db, dc, dG, stable, 0))) return rc;
    CK(cudaMemcpyAsync(G, dG, sizeof(c128) * batch * (size_t)d.N, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return end_host_call(ctx);
}

int mmh_forward_host(int ndim, const int64_t *shape, const void *A, const void *b, const void *c, void *G,
                     int stable) {
    return mmh_forward_batched_host(1, ndim, shape, A, b, c, G, stable);
}

int mmh_vjp(int ndim, const int64_t *shape, const void *dG, const void *dc, const void *ddLdG, void *oA,
            void *ob, void *oc, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return vjp_impl(1, ndim, shape, dG, dc, ddLdG, oA, ob, oc, (cudaStream_t)stream);
}

int mmh_vjp_batched(int64_t batch, int ndim, const int64_t *shape, const void *dG, const void *dc,
                    const void *ddLdG, void *oA, void *ob, void *oc, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return vjp_impl(batch, ndim, shape, dG, dc, ddLdG, oA, ob, oc, (cudaStream_t)stream);
}

int mmh_vjp_batched_host(int64_t batch, int ndim, const int64_t *shape, const void *G, const void *c,
                         const void *dLdG, void *oA, void *ob, void *oc) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (batch < 0) return MMH_ERR_BAD_BATCH;
    LatticeDesc d; int mx;
    int rc = make_desc(ndim, shape, &d, &mx);
    if (rc) return rc;
    if (batch == 0) return MMH_OK;
    if (!G || !c || !dLdG || !oA || !ob || !oc) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    if ((rc = get_ctx(&ctx))) return rc;
    const size_t D = ndim;
    void *dG, *dg, *dc, *dAo, *dbo, *dco;
    if ((rc = stage_in(*ctx, 0, G, sizeof(c128) * batch * (size_t)d.N, &dG))) return rc;
    if ((rc = stage_in(*ctx, 1, dLdG, sizeof(c128) * batch * (size_t)d.N, &dg))) return rc;
    if ((rc = stage_in(*ctx, 2, c, sizeof(c128) * batch, &dc))) return rc;
    if ((rc = stage_in(*ctx, 4, nullptr, sizeof(c128) * batch * D * D, &dAo))) return rc;
    if ((rc = stage_in(*ctx, 5, nullptr, sizeof(c128) * batch * D, &dbo))) return rc;
    if ((rc = stage_in(*ctx, 6, nullptr, sizeof(c128) * batch, &dco))) return rc;
    if ((rc = vjp_impl(batch, ndim, shape, dG, dc, dg, dAo, dbo, dco, 0))) return rc;
    CK(cudaMemcpyAsync(oA, dAo, sizeof(c128) * batch * D * D, cudaMemcpyDeviceToHost, 0));
    CK(cudaMemcpyAsync(ob, dbo, sizeof(c128) * batch * D, cudaMemcpyDeviceToHost, 0));
    CK(cudaMemcpyAsync(oc, dco, sizeof(c128) * batch, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return end_host_call(ctx);
}

int mmh_vjp_host(int ndim, const int64_t *shape, const void *G, const void *c, const void *dLdG, void *oA,
                 void *ob, void *oc) {
    return mmh_vjp_batched_host(1, ndim, shape, G, c, dLdG, oA, ob, oc);
}

int mmh_diagonal(int M, const int64_t *cutoffs, const void *dA, const void *dB, int64_t nbatch, const void *dG0,
                 void *dout, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return diagonal_impl(M, cutoffs, 0, dA, dB, nbatch, dG0, dout, (cudaStream_t)stream);
}
int mmh_diagonal_host(int M, const int64_t *cutoffs, const void *A, const void *B, int64_t nbatch, const void *G0,
                      void *out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return diagonal_host_impl(M, cutoffs, 0, A, B, nbatch, G0, out);
}
int mmh_diagonal_grad(int M, const int64_t *cutoffs, const void *dA, const void *dB, const void *dG0, void *o_dG0,
                      void *o_dA, void *o_dB, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return diagonal_grad_impl(M, cutoffs, 0, dA, dB, dG0, o_dG0, o_dA, o_dB, (cudaStream_t)stream);
}
int mmh_1leftover_grad(int M, const int64_t *cutoffs, const void *dA, const void *dB, const void *dG0, void *o_dG0,
                       void *o_dA, void *o_dB, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (M < 2) return MMH_ERR_BAD_NDIM;
    return diagonal_grad_impl(M, cutoffs, 1, dA, dB, dG0, o_dG0, o_dA, o_dB, (cudaStream_t)stream);
}
static int diagonal_grad_host_impl(int M, const int64_t *cutoffs, int L0, const void *A, const void *B, const void *G0,
                                   void *o_dG0, void *o_dA, void *o_dB);
int mmh_diagonal_grad_host(int M, const int64_t *cutoffs, const void *A, const void *B, const void *G0, void *o_dG0,
                           void *o_dA, void *o_dB) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return diagonal_grad_host_impl(M, cutoffs, 0, A, B, G0, o_dG0, o_dA, o_dB);
}
int mmh_1leftover_grad_host(int M, const int64_t *cutoffs, const void *A, const void *B, const void *G0, void *o_dG0,
                            void *o_dA, void *o_dB) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (M < 2) return MMH_ERR_BAD_NDIM;
    return diagonal_grad_host_impl(M, cutoffs, 1, A, B, G0, o_dG0, o_dA, o_dB);
}
static int diagonal_grad_host_impl(int M, const int64_t *cutoffs, int L0, const void *A, const void *B, const void *G0,
                                   void *o_dG0, void *o_dA, void *o_dB) {
    if (!cutoffs) return MMH_ERR_NULL_POINTER;
    if (M < 1 + L0 || M - L0 > 8) return MMH_ERR_BAD_NDIM;
    if (!A || !B || !G0 || !o_dG0 || !o_dA || !o_dB) return MMH_ERR_NULL_POINTER;
    size_t P = 1;
    for (int j = 0; j < M; j++) {
        if (cutoffs[j] < 1) return MMH_ERR_BAD_SHAPE;
        P *= (size_t)cutoffs[j] * ((L0 && j == 0) ? (size_t)cutoffs[j] : 1);
    }
    DeviceCtx *ctx;
    int rc;
    if ((rc = get_ctx(&ctx))) return rc;
    const size_t n2 = 2 * (size_t)M;
    void *dA, *dB, *dG0, *d0, *d1, *d2;
    if ((rc = stage_in(*ctx, 0, A, sizeof(c128) * n2 * n2, &dA))) return rc;
    if ((rc = stage_in(*ctx, 1, B, sizeof(c128) * n2, &dB))) return rc;
    if ((rc = stage_in(*ctx, 2, G0, sizeof(c128), &dG0))) return rc;
    if ((rc = stage_in(*ctx, 4, nullptr, sizeof(c128) * P, &d0))) return rc;
    if ((rc = stage_in(*ctx, 5, nullptr, sizeof(c128) * P * n2 * n2, &d1))) return rc;
    if ((rc = stage_in(*ctx, 6, nullptr, sizeof(c128) * P * n2, &d2))) return rc;
    if ((rc = diagonal_grad_impl(M, cutoffs, L0, dA, dB, dG0, d0, d1, d2, 0))) return rc;
    CK(cudaMemcpyAsync(o_dG0, d0, sizeof(c128) * P, cudaMemcpyDeviceToHost, 0));
    CK(cudaMemcpyAsync(o_dA, d1, sizeof(c128) * P * n2 * n2, cudaMemcpyDeviceToHost, 0));
    CK(cudaMemcpyAsync(o_dB, d2, sizeof(c128) * P * n2, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return end_host_call(ctx);
}
int mmh_1leftover(int M, const int64_t *cutoffs, const void *dA, const void *dB, const void *dG0, void *dout,
                  void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (M < 2) return MMH_ERR_BAD_NDIM;
    return diagonal_impl(M, cutoffs, 1, dA, dB, 0, dG0, dout, (cudaStream_t)stream);
}
int mmh_1leftover_host(int M, const int64_t *cutoffs, const void *A, const void *B, const void *G0, void *out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (M < 2) return MMH_ERR_BAD_NDIM;
    return diagonal_host_impl(M, cutoffs, 1, A, B, 0, G0, out);
}

int mmh_binomial(int ndim, const int64_t *shape, const void *dA, const void *db, const void *dc,
                 double max_l2, int64_t global_cutoff, void *dG, double *norm_out, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return binomial_impl(ndim, shape, dA, db, dc, max_l2, global_cutoff, dG, norm_out, (cudaStream_t)stream);
}

int mmh_binomial_host(int ndim, const int64_t *shape, const void *A, const void *b, const void *c,
                      double max_l2, int64_t global_cutoff, void *G, double *norm_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    LatticeDesc d; int mx;
    int rc = make_desc(ndim, shape, &d, &mx);
    if (rc) return rc;
    if (!A || !b || !c || !G) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    if ((rc = get_ctx(&ctx))) return rc;
    const size_t D = ndim;
    void *dA, *db, *dc, *dG;
    if ((rc = stage_in(*ctx, 0, A, sizeof(c128) * D * D, &dA))) return rc;
    if ((rc = stage_in(*ctx, 1, b, sizeof(c128) * D, &db))) return rc;
    if ((rc = stage_in(*ctx, 2, c, sizeof(c128), &dc))) return rc;
    if ((rc = stage_in(*ctx, 3, nullptr, sizeof(c128) * (size_t)d.N, &dG))) return rc;
    if ((rc = binomial_impl(ndim, shape, dA, db, dc, max_l2, global_cutoff, dG, norm_out, 0))) return rc;
    CK(cudaMemcpyAsync(G, dG, sizeof(c128) * (size_t)d.N, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return end_host_call(ctx);
}

}  // extern "C"

// debug aid (not part of the reference interface): copies the 16 x 4 %globaltimer stamps of the last launches (mmh_common.cuh)
extern "C" int mmh_debug_timeline(unsigned long long *out64) {
    if (!out64) return MMH_ERR_NULL_POINTER;
    CK(cudaDeviceSynchronize());
    DeviceCtx *ctx;
    int rc;
    if ((rc = get_ctx(&ctx))) return rc;
    if (!ctx->timeline.ptr) return MMH_ERR_UNSUPPORTED;
    CK(cudaMemcpy(out64, ctx->timeline.ptr, sizeof(unsigned long long) * 256, cudaMemcpyDeviceToHost));
    return MMH_OK;
}

// debug aid, host only: the launch plans of the batched lane / box kernels (tests/test_host_logic.py)
extern "C" int mmh_debug_plan(int what, int ndim, const int64_t *shape, int stage, int *out6) {
    std::lock_guard<std::mutex> lk_env(g_mutex);   // host-only planner export: takes the lock for the environment snapshot
    mmh_env_refresh();
    if (!shape || !out6) return MMH_ERR_NULL_POINTER;
    LatticeDesc d;
    int mx;
    int rc = make_desc(ndim, shape, &d, &mx);
    if (rc) return rc;
    for (int k = 0; k < 6; k++) out6[k] = 0;
    if (what == 0) {
        if (!mmh_plan_march_lanes(d.shape[ndim - 1], &out6[0], &out6[1], &out6[2])) return MMH_ERR_UNSUPPORTED;
        return MMH_OK;
    }
    if (what == 1) {
        if (stage < 0 || stage > ndim - 2) return MMH_ERR_BAD_SHAPE;
        BoxParams bp;
        int T;
        size_t smem;
        if (!mmh_plan_march_box(d, stage, &bp, &T, &smem)) return MMH_ERR_UNSUPPORTED;
        out6[0] = bp.g[0]; out6[1] = bp.g[1]; out6[2] = bp.g[2]; out6[3] = T; out6[4] = bp.nt; out6[5] = bp.ls;
        return MMH_OK;
    }
    if (what == 2) {   // single-lattice march of stage `stage` on 148 SMs: which kernel, which grid
        if (stage < 0 || stage > ndim - 2) return MMH_ERR_BAD_SHAPE;
        TiledParams tp;
        int R = 0, ntiles = 0;
        size_t smem = 0;
        if (!plan_march_tiled(d, stage, 148, &tp, &R, &ntiles, &smem)) return MMH_ERR_UNSUPPORTED;
        out6[0] = tp.rows_R ? 1 : 0;                         // 1: k_march_rows, 0: k_march_tiled2
        out6[1] = tp.g[0]; out6[2] = tp.g[1]; out6[3] = tp.g[2];
        out6[4] = tp.rows_R ? tp.rows_R * 1000 + tp.rows_C : R;
        out6[5] = tp.rows_R ? tp.rs : tp.tc;
        return MMH_OK;
    }
    if (what == 3) {   // stable box wavefront: box edge and boxes per dim (right-aligned in four dims)
        const int E = mmh_stable_boxes_edge(ndim);
        if (!E) return MMH_ERR_UNSUPPORTED;
        out6[0] = E;
        for (int j = 0; j < 4; j++) out6[1 + j] = j < 4 - ndim ? 1 : (d.shape[j - (4 - ndim)] + E - 1) / E;
        out6[5] = (int)mmh_stable_boxes_smem(ndim, mx + 1, 1, 1);
        return MMH_OK;
    }
    return MMH_ERR_UNSUPPORTED;
}

// ---- lattice + derived-variable contraction ---------------------------------------------------------
static int forward_contract_impl(long long batch, int ndim, const int64_t *shape, int ncore_dims, const void *dA, const void *db,
                                 const void *dcp, void *dout, int stable, cudaStream_t st) {
    if (batch < 0) return MMH_ERR_BAD_BATCH;
    LatticeDesc d;
    int mx;
    int rc = make_desc(ndim, shape, &d, &mx);
    if (rc) return rc;
    if (ncore_dims < 0 || ncore_dims > ndim) return MMH_ERR_BAD_NDIM;
    if (batch == 0) return MMH_OK;
    if (!dA || !db || !dcp || !dout) return MMH_ERR_NULL_POINTER;
    long long ncore = 1, nd = 1;
    for (int i = 0; i < ndim; i++) (i < ncore_dims ? ncore : nd) *= shape[i];
    if (nd > (1LL << 30)) return MMH_ERR_TOO_LARGE;
    DeviceCtx *ctx;
    if ((rc = get_ctx(&ctx))) return rc;
    if ((rc = begin_call(ctx, st))) return rc;
    if ((size_t)batch * (size_t)d.N > ((size_t)150 << 30) / sizeof(c128)) return MMH_ERR_TOO_LARGE;
    if ((rc = ensure_scratch(ctx->lattice_ws, sizeof(c128) * (size_t)batch * (size_t)d.N))) return rc;
    if (ctx->ones.bytes < sizeof(c128) * (size_t)batch) {
        if ((rc = ensure_scratch(ctx->ones, sizeof(c128) * (size_t)batch))) return rc;
        g_launches++;
        CK(mmh_launch_fill_ones((c128 *)ctx->ones.ptr, (long long)(ctx->ones.bytes / sizeof(c128)), st));
    }
    if ((rc = forward_impl(batch, ndim, shape, dA, db, ctx->ones.ptr, ctx->lattice_ws.ptr, stable, st))) return rc;
    g_launches++;
    CK(mmh_launch_contract_last((const c128 *)ctx->lattice_ws.ptr, (const c128 *)dcp, (c128 *)dout, batch * ncore, ncore, (int)nd, st));
    return MMH_OK;
}

extern "C" int mmh_forward_contract(int64_t batch, int ndim, const int64_t *shape, int ncore_dims, const void *dA, const void *db,
                                    const void *dcpoly, void *dout, int stable, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return forward_contract_impl(batch, ndim, shape, ncore_dims, dA, db, dcpoly, dout, stable, (cudaStream_t)stream);
}

extern "C" int mmh_forward_contract_host(int64_t batch, int ndim, const int64_t *shape, int ncore_dims, const void *A, const void *b,
                                         const void *cpoly, void *out, int stable) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (batch < 0) return MMH_ERR_BAD_BATCH;
    LatticeDesc d; int mx;
    int rc = make_desc(ndim, shape, &d, &mx);
    if (rc) return rc;
    if (ncore_dims < 0 || ncore_dims > ndim) return MMH_ERR_BAD_NDIM;
    if (batch == 0) return MMH_OK;
    if (!A || !b || !cpoly || !out) return MMH_ERR_NULL_POINTER;
    size_t ncore = 1, nd = 1;
    for (int i = 0; i < ndim; i++) (i < ncore_dims ? ncore : nd) *= (size_t)shape[i];
    DeviceCtx *ctx;
    if ((rc = get_ctx(&ctx))) return rc;
    void *dA, *db, *dcp, *dout;
    const size_t D = ndim;
    if ((rc = stage_in(*ctx, 0, A, sizeof(c128) * batch * D * D, &dA))) return rc;
    if ((rc = stage_in(*ctx, 1, b, sizeof(c128) * batch * D, &db))) return rc;
    if ((rc = stage_in(*ctx, 2, cpoly, sizeof(c128) * batch * nd, &dcp))) return rc;
    if ((rc = stage_in(*ctx, 3, nullptr, sizeof(c128) * batch * ncore, &dout))) return rc;
    if ((rc = forward_contract_impl(batch, ndim, shape, ncore_dims, dA, db, dcp, dout, stable, 0))) return rc;
    CK(cudaMemcpyAsync(out, dout, sizeof(c128) * batch * ncore, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return end_host_call(ctx);
}

// ---- gate-specific strategies (mmh_gates.cu) ----------------------------------------------
// The transcendental scalars are computed here, once, with libm -- the functions numba's lowering of np.cos / np.sin / np.tanh /
// np.cosh / np.exp(1j x) on scalars calls -- so that the device recurrences start from the reference's own bits.
static int gate_ctx(DeviceCtx **ctx, cudaStream_t st, int maxdim) {
    int rc;
    if ((rc = get_ctx(ctx))) return rc;
    if ((rc = begin_call(*ctx, st))) return rc;
    return ensure_tables(**ctx, maxdim + 2);
}
static int check_dims(const int64_t *shape, int n, int *mx, long long *total) {
    if (!shape) return MMH_ERR_NULL_POINTER;
    *mx = 1; *total = 1;
    for (int i = 0; i < n; i++) {
        if (shape[i] < 1 || shape[i] > (1 << 20)) return MMH_ERR_BAD_SHAPE;
        if (shape[i] > *mx) *mx = (int)shape[i];
        if (*total > (1LL << 36) / shape[i]) return MMH_ERR_TOO_LARGE;
        *total *= shape[i];
    }
    return MMH_OK;
}

static int squeezer_impl(int64_t M, int64_t N, double r, double theta, void *dS, cudaStream_t st) {
    const int64_t shape[2] = { M, N };
    int mx; long long total; int rc;
    if ((rc = check_dims(shape, 2, &mx, &total))) return rc;
    if (!dS) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    if ((rc = gate_ctx(&ctx, st, mx))) return rc;
    GateParams p; memset(&p, 0, sizeof(p));
    p.shape[0] = (int)M; p.shape[1] = (int)N; p.sq = ctx->sq; p.out = (c128 *)dS;
    const double t = std::tanh(r);
    p.z0 = make_double2(std::cos(theta) * t, std::sin(theta) * t);       // np.exp(1j * theta) * np.tanh(r)   (squeezer.py:49)
    p.r0 = 1.0 / std::cosh(r);                                          // sechr                              (:51)
    p.r1 = std::sqrt(p.r0);                                             // S[0, 0]                            (:53)
    g_launches++;
    CK(mmh_launch_squeezer(p, st));
    return MMH_OK;
}
static int squeezed_impl(int64_t cutoff, double r, double theta, void *dS, cudaStream_t st) {
    const int64_t shape[1] = { cutoff };
    int mx; long long total; int rc;
    if ((rc = check_dims(shape, 1, &mx, &total))) return rc;
    if (!dS) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    if ((rc = gate_ctx(&ctx, st, mx))) return rc;
    GateParams p; memset(&p, 0, sizeof(p));
    p.shape[0] = (int)cutoff; p.sq = ctx->sq; p.out = (c128 *)dS;
    const double t = -std::tanh(r);
    p.z0 = make_double2(std::cos(theta) * t, std::sin(theta) * t);       // np.exp(1j * theta) * -np.tanh(r)  (squeezer.py:140)
    p.r1 = std::sqrt(1.0 / std::cosh(r));                               // S[0]                               (:141)
    g_launches++;
    CK(mmh_launch_squeezed(p, st));
    return MMH_OK;
}
static int beamsplitter_impl(const int64_t *shape, double theta, double phi, int stable, void *dG, cudaStream_t st) {
    int mx; long long total; int rc;
    if ((rc = check_dims(shape, 4, &mx, &total))) return rc;
    if (!dG) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    if ((rc = gate_ctx(&ctx, st, mx))) return rc;
    GateParams p; memset(&p, 0, sizeof(p));
    for (int i = 0; i < 4; i++) p.shape[i] = (int)shape[i];
    p.sq = ctx->sq; p.out = (c128 *)dG;
    p.r0 = std::cos(theta);                                             // ct                                 (beamsplitter.py:59)
    const double s = std::sin(theta);
    p.z0 = make_double2(s * std::cos(phi), s * std::sin(phi));          // st = np.sin(theta) * np.exp(1j * phi) (:60)
    long long launches = 0;
    CK(mmh_launch_beamsplitter(p, stable != 0, &launches, st));
    g_launches += launches;
    return MMH_OK;
}
static int displacement_impl(int64_t c0, int64_t c1, double are, double aim, void *dD, cudaStream_t st) {
    const int64_t shape[2] = { c0, c1 };
    int mx; long long total; int rc;
    if ((rc = check_dims(shape, 2, &mx, &total))) return rc;
    if (!dD) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    if ((rc = gate_ctx(&ctx, st, mx))) return rc;
    const bool flipped = c0 < c1;                                       // displacement.py:40-43
    const int N = (int)(flipped ? c1 : c0), M = (int)(flipped ? c0 : c1);
    // log_k_fac = cumsum(log(arange(max) with rng[0] = 1))  (displacement.py:46-48), built on the host, staged per call
    std::vector<double> lf(mx);
    double acc = 0.0;
    for (int k = 0; k < mx; k++) { acc += std::log(k == 0 ? 1.0 : (double)k); lf[k] = acc; }
    const size_t lf_bytes = (sizeof(double) * (size_t)mx + 255) / 256 * 256;
    if ((rc = ensure_scratch(ctx->gate_ws, lf_bytes + (flipped ? sizeof(c128) * (size_t)total : 0)))) return rc;
    CK(cudaMemcpyAsync(ctx->gate_ws.ptr, lf.data(), sizeof(double) * (size_t)mx, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // lf is a local
    GateParams p; memset(&p, 0, sizeof(p));
    p.shape[0] = N; p.shape[1] = M; p.flag = flipped ? 1 : 0; p.sq = ctx->sq;
    c128 *work = flipped ? (c128 *)((char *)ctx->gate_ws.ptr + lf_bytes) : (c128 *)dD;
    p.out = work;
    p.r1 = std::hypot(are, aim);                                        // r = np.abs(alpha)
    p.r2 = std::atan2(aim, are);                                        // phi = np.angle(alpha)
    p.r0 = p.r1 * p.r1;                                                 // r ** 2.0
    g_launches += 2;
    CK(mmh_launch_displacement(p, (const double *)ctx->gate_ws.ptr, st));
    if (flipped) { g_launches++; CK(mmh_launch_transpose(work, (c128 *)dD, N, M, st)); }
    return MMH_OK;
}
static int disp_derivs_impl(int kind, int64_t M, int64_t N, const void *dD, double a0, double a1, void *o1, void *o2, cudaStream_t st) {
    const int64_t shape[2] = { M, N };
    int mx; long long total; int rc;
    if ((rc = check_dims(shape, 2, &mx, &total))) return rc;
    if (!dD || !o1 || !o2) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    if ((rc = gate_ctx(&ctx, st, mx))) return rc;
    GateParams p; memset(&p, 0, sizeof(p));
    p.shape[0] = (int)M; p.shape[1] = (int)N; p.sq = ctx->sq;
    if (kind == 0) p.z0 = make_double2(a0, a1);                         // alpha
    else {                                                              // (r, phi)
        p.r1 = a0; p.r0 = std::cos(a1); p.r2 = std::sin(a1);
        p.z0 = make_double2(a0 * p.r0, a0 * p.r2);                      // alpha = r * exp(1j * phi)
    }
    g_launches++;
    CK(mmh_launch_disp_derivs(p, (const c128 *)dD, (c128 *)o1, (c128 *)o2, kind, st));
    return MMH_OK;
}
// out: [ndim * ndim] un-symmetrised upper-triangular sums U | [ndim] dLdb | [1] sum(G * dLdG over the support)
static int gate_vjp_impl(int kind, int ndim, const int64_t *shape, const void *dG, const void *dg, void *dout, cudaStream_t st) {
    const int want = kind == 0 ? 4 : (kind == 1 ? 2 : 1);
    if (kind < 0 || kind > 2 || ndim != want) return MMH_ERR_BAD_NDIM;
    int mx; long long total; int rc;
    if ((rc = check_dims(shape, ndim, &mx, &total))) return rc;
    if (!dG || !dg || !dout) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    if ((rc = gate_ctx(&ctx, st, mx))) return rc;
    const int nout = ndim * ndim + ndim + 1;
    const size_t head = (sizeof(c128) * (size_t)(nout + 1) + 255) / 256 * 256;
    if ((rc = ensure_scratch(ctx->gate_ws, head + sizeof(c128) * (size_t)total))) return rc;
    c128 *sym = (c128 *)ctx->gate_ws.ptr, *one = sym + nout, *gm = (c128 *)((char *)ctx->gate_ws.ptr + head);
    g_launches += 3;
    CK(mmh_launch_fill_ones(one, 1, st));
    CK(mmh_launch_gate_mask((const c128 *)dg, gm, total, kind, (int)shape[ndim > 1 ? 1 : 0], ndim > 2 ? (int)shape[2] : 1,
                            ndim > 3 ? (int)shape[3] : 1, st));
    if ((rc = vjp_impl(1, ndim, shape, dG, one, gm, sym, sym + ndim * ndim, sym + ndim * ndim + ndim, st))) return rc;
    CK(mmh_launch_gate_unsym(sym, (c128 *)dout, ndim, st));
    CK(cudaMemcpyAsync((c128 *)dout + ndim * ndim, sym + ndim * ndim, sizeof(c128) * (size_t)(ndim + 1), cudaMemcpyDeviceToDevice, st));
    return MMH_OK;
}

extern "C" {
int mmh_squeezer(int64_t M, int64_t N, double r, double theta, void *dS, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return squeezer_impl(M, N, r, theta, dS, (cudaStream_t)stream);
}
int mmh_squeezed(int64_t cutoff, double r, double theta, void *dS, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return squeezed_impl(cutoff, r, theta, dS, (cudaStream_t)stream);
}
int mmh_beamsplitter(const int64_t *shape, double theta, double phi, int stable, void *dG, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return beamsplitter_impl(shape, theta, phi, stable, dG, (cudaStream_t)stream);
}
int mmh_displacement(int64_t c0, int64_t c1, double alpha_re, double alpha_im, void *dD, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return displacement_impl(c0, c1, alpha_re, alpha_im, dD, (cudaStream_t)stream);
}
int mmh_displacement_jacobian(int64_t M, int64_t N, const void *dD, double alpha_re, double alpha_im, void *d_jac_alpha,
                              void *d_jac_alphac, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return disp_derivs_impl(0, M, N, dD, alpha_re, alpha_im, d_jac_alpha, d_jac_alphac, (cudaStream_t)stream);
}
int mmh_displacement_grad(int64_t cutoff, const void *dT, double r, double phi, void *d_grad_r, void *d_grad_phi, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return disp_derivs_impl(1, cutoff, cutoff, dT, r, phi, d_grad_r, d_grad_phi, (cudaStream_t)stream);
}
int mmh_gate_vjp(int kind, int ndim, const int64_t *shape, const void *dG, const void *ddLdG, void *dout, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return gate_vjp_impl(kind, ndim, shape, dG, ddLdG, dout, (cudaStream_t)stream);
}

// host-pointer variants (numpy drop-in): what = 0 squeezer(M, N, r, theta), 1 squeezed(M, r, theta), 2 beamsplitter(shape4, theta, phi),
// 3 stable_beamsplitter, 4 displacement(c0, c1, alpha)
int mmh_gate_host(int what, const int64_t *shape, double a0, double a1, void *out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (!shape || !out) return MMH_ERR_NULL_POINTER;
    const int nd = what == 1 ? 1 : ((what == 2 || what == 3) ? 4 : 2);
    if (what < 0 || what > 4) return MMH_ERR_UNSUPPORTED;
    int mx; long long total; int rc;
    if ((rc = check_dims(shape, nd, &mx, &total))) return rc;
    DeviceCtx *ctx;
    if ((rc = get_ctx(&ctx))) return rc;
    void *dout;
    if ((rc = stage_in(*ctx, 3, nullptr, sizeof(c128) * (size_t)total, &dout))) return rc;
    switch (what) {
        case 0: rc = squeezer_impl(shape[0], shape[1], a0, a1, dout, 0); break;
        case 1: rc = squeezed_impl(shape[0], a0, a1, dout, 0); break;
        case 2: rc = beamsplitter_impl(shape, a0, a1, 0, dout, 0); break;
        case 3: rc = beamsplitter_impl(shape, a0, a1, 1, dout, 0); break;
        default: rc = displacement_impl(shape[0], shape[1], a0, a1, dout, 0); break;
    }
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, dout, sizeof(c128) * (size_t)total, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return end_host_call(ctx);
}
// what = 0 jacobian_displacement(D[M, N], alpha) -> (jac_alpha, jac_alphac); 1 grad_displacement(T[c, c], r, phi) -> (grad_r, grad_phi)
int mmh_displacement_derivs_host(int what, int64_t M, int64_t N, const void *D, double a0, double a1, void *o1, void *o2) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (!D || !o1 || !o2) return MMH_ERR_NULL_POINTER;
    if (what < 0 || what > 1) return MMH_ERR_UNSUPPORTED;
    const int64_t shape[2] = { M, N };
    int mx; long long total; int rc;
    if ((rc = check_dims(shape, 2, &mx, &total))) return rc;
    DeviceCtx *ctx;
    if ((rc = get_ctx(&ctx))) return rc;
    void *dD, *d1, *d2;
    const size_t bytes = sizeof(c128) * (size_t)total;
    if ((rc = stage_in(*ctx, 0, D, bytes, &dD))) return rc;
    if ((rc = stage_in(*ctx, 4, nullptr, bytes, &d1))) return rc;
    if ((rc = stage_in(*ctx, 5, nullptr, bytes, &d2))) return rc;
    if ((rc = disp_derivs_impl(what, M, N, dD, a0, a1, d1, d2, 0))) return rc;
    CK(cudaMemcpyAsync(o1, d1, bytes, cudaMemcpyDeviceToHost, 0));
    CK(cudaMemcpyAsync(o2, d2, bytes, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return end_host_call(ctx);
}
int mmh_gate_vjp_host(int kind, int ndim, const int64_t *shape, const void *G, const void *dLdG, void *out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (!G || !dLdG || !out) return MMH_ERR_NULL_POINTER;
    int mx; long long total; int rc;
    if (ndim < 1 || ndim > 4) return MMH_ERR_BAD_NDIM;
    if ((rc = check_dims(shape, ndim, &mx, &total))) return rc;
    DeviceCtx *ctx;
    if ((rc = get_ctx(&ctx))) return rc;
    void *dG, *dg, *dout;
    const size_t bytes = sizeof(c128) * (size_t)total, obytes = sizeof(c128) * (size_t)(ndim * ndim + ndim + 1);
    if ((rc = stage_in(*ctx, 0, G, bytes, &dG))) return rc;
    if ((rc = stage_in(*ctx, 1, dLdG, bytes, &dg))) return rc;
    if ((rc = stage_in(*ctx, 4, nullptr, obytes, &dout))) return rc;
    if ((rc = gate_vjp_impl(kind, ndim, shape, dG, dg, dout, 0))) return rc;
    CK(cudaMemcpyAsync(out, dout, obytes, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return end_host_call(ctx);
}
}  // extern "C"

// ---- autoshape (mmh_autoshape.cu) ------------------------------------------------------------
static int autoshape_impl(int M, const void *dA, const void *db, const void *dc, double max_prob, long long max_shape,
                          long long min_shape, void *dshape, cudaStream_t st) {
    if (M < 1 || M > 32) return MMH_ERR_BAD_NDIM;
    if (max_shape < 0 || max_shape > (1 << 20)) return MMH_ERR_BAD_SHAPE;
    if (!dA || !db || !dc || !dshape) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    int rc;
    if ((rc = gate_ctx(&ctx, st, (int)max_shape + 2))) return rc;
    g_launches++;
    CK(mmh_launch_autoshape(M, (const c128 *)dA, (const c128 *)db, (const c128 *)dc, max_prob, max_shape, min_shape,
                            (long long *)dshape, ctx->sq, st));
    return MMH_OK;
}
extern "C" int mmh_autoshape(int M, const void *dA, const void *db, const void *dc, double max_prob, int64_t max_shape,
                             int64_t min_shape, int64_t *dshape_out, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return autoshape_impl(M, dA, db, dc, max_prob, max_shape, min_shape, dshape_out, (cudaStream_t)stream);
}
extern "C" int mmh_autoshape_host(int M, const void *A, const void *b, const void *c, double max_prob, int64_t max_shape,
                                  int64_t min_shape, int64_t *shape_out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (M < 1 || M > 32) return MMH_ERR_BAD_NDIM;
    if (!A || !b || !c || !shape_out) return MMH_ERR_NULL_POINTER;
    DeviceCtx *ctx;
    int rc;
    if ((rc = get_ctx(&ctx))) return rc;
    void *dA, *db, *dc, *dsh;
    const size_t n2 = 2 * (size_t)M;
    if ((rc = stage_in(*ctx, 0, A, sizeof(c128) * n2 * n2, &dA))) return rc;
    if ((rc = stage_in(*ctx, 1, b, sizeof(c128) * n2, &db))) return rc;
    if ((rc = stage_in(*ctx, 2, c, sizeof(c128), &dc))) return rc;
    if ((rc = stage_in(*ctx, 4, nullptr, sizeof(int64_t) * (size_t)M, &dsh))) return rc;
    if ((rc = autoshape_impl(M, dA, db, dc, max_prob, max_shape, min_shape, dsh, 0))) return rc;
    CK(cudaMemcpyAsync(shape_out, dsh, sizeof(int64_t) * (size_t)M, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return end_host_call(ctx);
}

// ---- Fock-space contraction / reduce (mmh_einsum.cu) -------------------------------------------
static int fock_contract_impl(int nA, const int64_t *shapeA, const int *labA, int nB, const int64_t *shapeB, const int *labB, int nO,
                              const int *labO, const void *dA, const void *dB, void *dC, int64_t *out_shape, cudaStream_t st) {
    if (nA < 0 || nA > MMH_MAX_DIM || nB < 0 || nB > MMH_MAX_DIM || nO < 0 || nO > 2 * MMH_MAX_DIM) return MMH_ERR_BAD_NDIM;
    if ((nA && (!shapeA || !labA)) || (nB && (!shapeB || !labB)) || (nO && !labO)) return MMH_ERR_NULL_POINTER;
    if (!dA || !dB || !dC) return MMH_ERR_NULL_POINTER;
    const int NL = 128;
    long long dimA[NL], dimB[NL], strA[NL], strB[NL], strC[NL], dim[NL];
    bool inA[NL] = { false }, inB[NL] = { false }, inO[NL] = { false };
    for (int l = 0; l < NL; l++) { dimA[l] = dimB[l] = dim[l] = 0; strA[l] = strB[l] = strC[l] = 0; }
    long long s = 1;
    for (int i = nA - 1; i >= 0; i--) {
        const int l = labA[i];
        if (l < 0 || l >= NL || inA[l]) return MMH_ERR_UNSUPPORTED;   // repeated label inside one operand (a trace): not handled here
        if (shapeA[i] < 1) return MMH_ERR_BAD_SHAPE;
        inA[l] = true; dimA[l] = shapeA[i]; strA[l] = s; s *= shapeA[i];
    }
    s = 1;
    for (int i = nB - 1; i >= 0; i--) {
        const int l = labB[i];
        if (l < 0 || l >= NL || inB[l]) return MMH_ERR_UNSUPPORTED;
        if (shapeB[i] < 1) return MMH_ERR_BAD_SHAPE;
        inB[l] = true; dimB[l] = shapeB[i]; strB[l] = s; s *= shapeB[i];
    }
    for (int l = 0; l < NL; l++)   // shared labels run over the common minimum (array_ansatz.py:209-218)
        dim[l] = (inA[l] && inB[l]) ? (dimA[l] < dimB[l] ? dimA[l] : dimB[l]) : (inA[l] ? dimA[l] : dimB[l]);
    s = 1;
    for (int i = nO - 1; i >= 0; i--) {
        const int l = labO[i];
        if (l < 0 || l >= NL || inO[l] || !(inA[l] || inB[l])) return MMH_ERR_BAD_SHAPE;
        inO[l] = true; strC[l] = s; s *= dim[l];
        if (out_shape) out_shape[i] = dim[l];
    }
    // groups, in label order: batch (A, B, O), M (A, O), N (B, O), K (not O)
    std::vector<int> gb, gm, gn, gk;
    for (int l = 0; l < NL; l++) {
        if (!inA[l] && !inB[l]) continue;
        if (inO[l]) (inA[l] && inB[l] ? gb : (inA[l] ? gm : gn)).push_back(l);
        else gk.push_back(l);
    }
    auto count = [&](const std::vector<int> &g) { long long n = 1; for (int l : g) n *= dim[l]; return n; };
    const long long nb = count(gb), M = count(gm), N = count(gn), K = count(gk);
    if (nb > 65535 || M > (1LL << 26) || N > (1LL << 26) || K > (1LL << 26)) return MMH_ERR_TOO_LARGE;
    auto table = [&](const std::vector<int> &g, long long n, const long long *stride, std::vector<long long> &out) {
        const size_t base = out.size();
        out.resize(base + (size_t)n);
        for (long long f = 0; f < n; f++) {
            long long rem = f, off = 0;
            for (int t = (int)g.size() - 1; t >= 0; t--) { const int l = g[t]; off += (rem % dim[l]) * stride[l]; rem /= dim[l]; }
            out[base + (size_t)f] = off;
        }
        return base;
    };
    std::vector<long long> tab;
    const size_t oAb = table(gb, nb, strA, tab), oBb = table(gb, nb, strB, tab), oCb = table(gb, nb, strC, tab);
    const size_t oAm = table(gm, M, strA, tab), oCm = table(gm, M, strC, tab);
    const size_t oBn = table(gn, N, strB, tab), oCn = table(gn, N, strC, tab);
    const size_t oAk = table(gk, K, strA, tab), oBk = table(gk, K, strB, tab);
    DeviceCtx *ctx;
    int rc;
    if ((rc = get_ctx(&ctx))) return rc;
    if ((rc = begin_call(ctx, st))) return rc;
    if ((rc = ensure_scratch(ctx->ein_ws, sizeof(long long) * tab.size()))) return rc;
    CK(cudaMemcpyAsync(ctx->ein_ws.ptr, tab.data(), sizeof(long long) * tab.size(), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // tab is a local
    const long long *d = (const long long *)ctx->ein_ws.ptr;
    EinsumParams p;
    p.A = (const c128 *)dA; p.B = (const c128 *)dB; p.C = (c128 *)dC;
    p.M = M; p.N = N; p.K = K; p.nbatch = nb;
    p.offA_b = d + oAb; p.offB_b = d + oBb; p.offC_b = d + oCb;
    p.offA_m = d + oAm; p.offC_m = d + oCm; p.offB_n = d + oBn; p.offC_n = d + oCn; p.offA_k = d + oAk; p.offB_k = d + oBk;
    g_launches++;
    CK(mmh_launch_einsum(p, st));
    return MMH_OK;
}

static int fock_reduce_impl(int ndim, const int64_t *in_shape, const int64_t *out_shape, const void *din, void *dout, cudaStream_t st) {
    if (ndim < 0 || ndim > MMH_MAX_DIM) return MMH_ERR_BAD_NDIM;
    if (ndim && (!in_shape || !out_shape)) return MMH_ERR_NULL_POINTER;
    if (!din || !dout) return MMH_ERR_NULL_POINTER;
    ReduceParams p;
    memset(&p, 0, sizeof(p));
    p.ndim = ndim; p.in = (const c128 *)din; p.out = (c128 *)dout;
    long long s = 1, n = 1;
    for (int d = ndim - 1; d >= 0; d--) {
        if (in_shape[d] < 1 || out_shape[d] < 1) return MMH_ERR_BAD_SHAPE;
        p.in_shape[d] = in_shape[d]; p.out_shape[d] = out_shape[d]; p.in_stride[d] = s;
        s *= in_shape[d];
        if (n > (1LL << 40) / out_shape[d]) return MMH_ERR_TOO_LARGE;
        n *= out_shape[d];
    }
    p.n_out = n;
    DeviceCtx *ctx;
    int rc;
    if ((rc = get_ctx(&ctx))) return rc;
    if ((rc = begin_call(ctx, st))) return rc;
    g_launches++;
    CK(mmh_launch_fock_reduce(p, st));
    return MMH_OK;
}

extern "C" {
int mmh_fock_contract(int nA, const int64_t *shapeA, const int *labelsA, int nB, const int64_t *shapeB, const int *labelsB, int nOut,
                      const int *labelsOut, const void *dA, const void *dB, void *dC, int64_t *out_shape, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return fock_contract_impl(nA, shapeA, labelsA, nB, shapeB, labelsB, nOut, labelsOut, dA, dB, dC, out_shape, (cudaStream_t)stream);
}
int mmh_fock_contract_host(int nA, const int64_t *shapeA, const int *labelsA, int nB, const int64_t *shapeB, const int *labelsB, int nOut,
                           const int *labelsOut, const void *A, const void *B, void *C) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (!A || !B || !C) return MMH_ERR_NULL_POINTER;
    if (nA < 0 || nA > MMH_MAX_DIM || nB < 0 || nB > MMH_MAX_DIM || nOut < 0 || nOut > 2 * MMH_MAX_DIM) return MMH_ERR_BAD_NDIM;
    size_t na = 1, nbb = 1;
    for (int i = 0; i < nA; i++) { if (!shapeA || shapeA[i] < 1) return MMH_ERR_BAD_SHAPE; na *= (size_t)shapeA[i]; }
    for (int i = 0; i < nB; i++) { if (!shapeB || shapeB[i] < 1) return MMH_ERR_BAD_SHAPE; nbb *= (size_t)shapeB[i]; }
    // output size: product over the output labels of the (common-minimum) dims
    size_t nc = 1;
    for (int i = 0; i < nOut; i++) {
        long long d = 0;
        for (int j = 0; j < nA; j++) if (labelsA[j] == labelsOut[i]) d = shapeA[j];
        for (int j = 0; j < nB; j++) if (labelsB[j] == labelsOut[i]) d = (d == 0 || shapeB[j] < d) ? shapeB[j] : d;
        if (d < 1) return MMH_ERR_BAD_SHAPE;
        nc *= (size_t)d;
    }
    DeviceCtx *ctx;
    int rc;
    if ((rc = get_ctx(&ctx))) return rc;
    void *dA, *dB, *dC;
    if ((rc = stage_in(*ctx, 0, A, sizeof(c128) * na, &dA))) return rc;
    if ((rc = stage_in(*ctx, 1, B, sizeof(c128) * nbb, &dB))) return rc;
    if ((rc = stage_in(*ctx, 3, nullptr, sizeof(c128) * nc, &dC))) return rc;
    if ((rc = fock_contract_impl(nA, shapeA, labelsA, nB, shapeB, labelsB, nOut, labelsOut, dA, dB, dC, nullptr, 0))) return rc;
    CK(cudaMemcpyAsync(C, dC, sizeof(c128) * nc, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return end_host_call(ctx);
}
int mmh_fock_reduce(int ndim, const int64_t *in_shape, const int64_t *out_shape, const void *din, void *dout, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return fock_reduce_impl(ndim, in_shape, out_shape, din, dout, (cudaStream_t)stream);
}
}  // extern "C"

// ---- bilinear overlap of two lattices (mmh_vjp.cu k_dot_*) ----------------------------------------------------------------
extern "C" int mmh_overlap(int64_t n, const void *dx, const void *dy, void *dout, void *stream) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (n < 0) return MMH_ERR_BAD_SHAPE;
    if (!dx || !dy || !dout) return MMH_ERR_NULL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceCtx *ctx;
    int rc;
    if ((rc = get_ctx(&ctx))) return rc;
    if ((rc = begin_call(ctx, st))) return rc;
    long long want = (n + 256 * 8 - 1) / (256 * 8);
    const long long cap = 4LL * ctx->sm_count;
    int nblk = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    if ((rc = ensure_scratch(ctx->dot_ws, sizeof(c128) * (size_t)nblk))) return rc;
    g_launches += 2;
    CK(mmh_launch_dot((const c128 *)dx, (const c128 *)dy, n, (c128 *)ctx->dot_ws.ptr, nblk, (c128 *)dout, st));
    return MMH_OK;
}

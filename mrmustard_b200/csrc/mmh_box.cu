// mmh_box.cu — batches of lattices whose panels do not fit one CTA's shared memory (sm_100a): vanilla_batch_numba
// (vanilla/batch.py:27-61) over e.g. 2-mode unitaries at cutoff 20-30, (20,)^4 .. (30,)^4.
//
// One CTA marches one lattice's stage, box by box.  The panel of stage i (the sub-lattice of the stages > i) is cut into boxes
// of <= 1024 points over its first <= 3 dims exactly as the tile grid of k_march_tiled2 (mmh_tiled.cu), but the boxes of a
// lattice are marched ONE AFTER THE OTHER by the same CTA, in ascending order.  The update of a point (vanilla/core.py:97-104,
// pivot = first non-zero index) only reads lower neighbours, so when box t starts, every amplitude its low faces need -- all
// S panels of the boxes below it -- is final and sits in the lattice itself: the halo of panel s-1 is simply loaded from G
// (L2) one step ahead.  No exchange buffer, no polling, no inter-CTA dependency; different CTAs march different lattices.
// Inside a box the march is the register/shared-memory panel march of k_march_stage: G[k - e_i], G[k - 2 e_i] in registers,
// G[k - e_i - e_j] from a double-buffered shared-memory copy of the box (+ halo area), every amplitude written once.
// The k_fwd_cta kernel this replaces fetched every neighbour through L1/L2 (3 ns per amplitude and CTA).
// Arithmetic: k_march_stage's, operation for operation.
#include <cstdlib>
#include <cstring>

#include "mmh_params.cuh"

__device__ __forceinline__ c128 lds_c128_b(unsigned addr) {
    c128 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_c128_b(unsigned addr, c128 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

template <int R>
__device__ __forceinline__ void div_all_box(c128 (&v)[R], double sqs, double rsqs) {
    bool slow = false;
#ifndef MMH_BOX_EXPERIMENT_NO_RANGE_TEST   /* timing experiment only: results are wrong for tiny / huge numerators */
#pragma unroll
    for (int r = 0; r < R; r++) slow |= div_needs_slow(v[r].x) | div_needs_slow(v[r].y);
#endif
    if (!slow) {
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = c_make(div_fast(v[r].x, sqs, rsqs), div_fast(v[r].y, sqs, rsqs));
    } else {
#pragma unroll
        for (int r = 0; r < R; r++)
            v[r] = c_make(div_needs_slow(v[r].x) ? div_rare(v[r].x, sqs, rsqs) : div_fast(v[r].x, sqs, rsqs),
                          div_needs_slow(v[r].y) ? div_rare(v[r].y, sqs, rsqs) : div_fast(v[r].y, sqs, rsqs));
    }
}

#define MMH_BOX_HPT 2   // halo cells per thread (the plan keeps the halo of a box <= MMH_BOX_HPT * threads)

// smem (c128 cells): buf[2][ls] | sqtab[S] double2.   buf: [0, TS) own cells | [TS, TS + HC) low halo faces | zero | trash
// HIST: the two previous panels of a slot are re-read from shared memory (three rotating panel buffers) instead of living in
// registers.  A slot then holds ~17 registers (coefficients, offsets) instead of ~63, so four slots per thread fit 128 registers
// and two 256-thread CTAs are resident: twice the independent dependency chains per scheduler (the step is bound by the
// instruction-level parallelism in flight), for two more shared-memory loads per amplitude.
template <int R, int NPD, bool HIST>
__global__ void __launch_bounds__(HIST ? 256 : 512, HIST ? 2 : 1) k_march_box(BoxParams p) {
    constexpr int NBUF = HIST ? 3 : 2;
    extern __shared__ c128 smem[];
    const LatticeDesc &d = p.d;
    const int D = d.D, i = p.stage;
    const int P = (int)d.strides[i], S = d.shape[i];
    const int T = blockDim.x, tid = threadIdx.x, nt = p.nt;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned bstride = (unsigned)p.ls * 16u;
    const unsigned zero_off = (unsigned)(p.ls - 2) * 16u, trash_off = (unsigned)(p.ls - 1) * 16u;
    double2 *sqtab = (double2 *)(smem + (size_t)NBUF * p.ls);
    for (int s_ = tid; s_ < S; s_ += T) sqtab[s_] = make_double2(p.sq[s_], p.rsq[s_]);
    if (tid < NBUF) { smem[(size_t)tid * p.ls + p.ls - 2] = c_make(0.0, 0.0); smem[(size_t)tid * p.ls + p.ls - 1] = c_make(0.0, 0.0); }
    int g[3], shp[3], gst[3];
#pragma unroll
    for (int m = 0; m < 3; m++) {
        g[m] = m < nt ? p.g[m] : 1;
        shp[m] = m < nt ? d.shape[i + 1 + m] : 1;
        gst[m] = m < nt ? (int)d.strides[i + 1 + m] : 0;
    }
    const int inner = (int)d.strides[i + nt];
    const int ntiles = g[0] * g[1] * g[2];
    int lst[NPD];
#pragma unroll
    for (int jj = 0; jj < NPD; jj++) lst[jj] = (int)d.strides[i + 1 + jj];

#pragma unroll 1
    for (long long lat = blockIdx.x; lat < p.batch; lat += gridDim.x) {
        c128 *G = p.G + lat * p.lat_stride;
        const c128 b0 = p.b[lat * D + i], a00 = p.A[lat * D * D + i * D + i];
        const c128 *Arow = p.A + (lat * D * D + i * D + i + 1);
#pragma unroll 1
        for (int tile = 0; tile < ntiles; tile++) {
            // ---- box geometry (as k_march_tiled2) ----
            int t[3], lo[3], e[3], h[3];
            t[2] = tile % g[2];
            t[1] = (tile / g[2]) % g[1];
            t[0] = tile / (g[1] * g[2]);
#pragma unroll
            for (int m = 0; m < 3; m++) {
                lo[m] = (int)(((long long)t[m] * shp[m]) / g[m]);
                e[m] = (int)(((long long)(t[m] + 1) * shp[m]) / g[m]) - lo[m];
                h[m] = (m < nt && lo[m] > 0) ? 1 : 0;
            }
            int cst[3];
            cst[2] = inner; cst[1] = cst[2] * e[2]; cst[0] = cst[1] * e[1];
            const int TS = e[0] * e[1] * e[2] * inner;
            int faceoff[3];
            faceoff[0] = 0;
            faceoff[1] = faceoff[0] + h[0] * (TS / e[0]);
            faceoff[2] = faceoff[1] + h[1] * (TS / e[1]);
            const int HC = faceoff[2] + h[2] * (TS / e[2]);
            const unsigned halo_off = (unsigned)TS * 16u;

            // ---- per-slot constants ----
            unsigned loco[R], nbo[R][NPD], gofs[R];
            bool act[R];
            c128 coef[R][NPD], h0[R], h1[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int q = r * T + tid;
                act[r] = q < TS;
                const int qq = act[r] ? q : 0;
                const int rr = qq % inner;
                int q1 = qq / inner;
                int x[3];
                x[2] = q1 % e[2]; q1 /= e[2];
                x[1] = q1 % e[1];
                x[0] = q1 / e[1];
                loco[r] = act[r] ? (unsigned)qq * 16u : (HIST ? zero_off : trash_off);   // HIST reads through it: zero cell
                gofs[r] = (unsigned)((lo[0] + x[0]) * gst[0] + (lo[1] + x[1]) * gst[1] + (lo[2] + x[2]) * gst[2] + rr);
                int rem = rr;
#pragma unroll
                for (int jj = 0; jj < NPD; jj++) {
                    int k, nb;
                    if (jj < nt) {
                        k = lo[jj] + x[jj];
                        if (x[jj] > 0) nb = qq - cst[jj];
                        else {   // one cell below the box: halo face jj (read only when k > 0, i.e. when the face exists)
                            const int xa = jj == 0 ? x[1] : x[0], xb = jj == 2 ? x[1] : x[2], eb = jj == 2 ? e[1] : e[2];
                            nb = TS + faceoff[jj] + (xa * eb + xb) * inner + rr;
                        }
                    } else {
                        k = rem / lst[jj];
                        rem -= k * lst[jj];
                        nb = qq - lst[jj];
                    }
                    const bool has = act[r] && k > 0;
                    nbo[r][jj] = has ? (unsigned)nb * 16u : zero_off;
                    coef[r][jj] = has ? c_scale(Arow[jj], p.sq[k]) : c_make(0.0, 0.0);   // A_ij sqrt(k_j), core.py:103
                }
            }
            // ---- this thread's halo cells: panel offset of the amplitude one cell below the box ----
            unsigned hgo[MMH_BOX_HPT];
            bool hact[MMH_BOX_HPT];
#pragma unroll
            for (int w = 0; w < MMH_BOX_HPT; w++) {
                const int c = w * T + tid;
                hact[w] = c < HC;
                int m = 0;
                if (c >= faceoff[2] && h[2]) m = 2;
                else if (c >= faceoff[1] && h[1]) m = 1;
                int cc = hact[w] ? c - (m == 0 ? faceoff[0] : (m == 1 ? faceoff[1] : faceoff[2])) : 0;
                const int a = m == 0 ? 1 : 0, b = m == 2 ? 1 : 2;   // the two other box dims, in order
                const int eb = b == 1 ? e[1] : e[2];
                const int rr = cc % inner; cc /= inner;
                const int xb = cc % eb, xa = cc / eb;
                const int lom = m == 0 ? lo[0] : (m == 1 ? lo[1] : lo[2]), gm = m == 0 ? gst[0] : (m == 1 ? gst[1] : gst[2]);
                const int loa = a == 0 ? lo[0] : lo[1], ga = a == 0 ? gst[0] : gst[1];
                const int lob = b == 1 ? lo[1] : lo[2], gb = b == 1 ? gst[1] : gst[2];
                hgo[w] = (unsigned)((lom - 1) * gm + (loa + xa) * ga + (lob + xb) * gb + rr);
            }
            __syncthreads();   // the previous box is done with both buffers (and the tables are there)
            // ---- panel 0 -> buffer 0 (own cells and halo) ----
#pragma unroll
            for (int r = 0; r < R; r++) {
                h0[r] = c_make(0.0, 0.0);
                h1[r] = act[r] ? __ldcg(G + gofs[r]) : c_make(0.0, 0.0);
                sts_c128_b(sbase + (act[r] ? loco[r] : trash_off), h1[r]);
                if (HIST) sts_c128_b(sbase + 2u * bstride + (act[r] ? loco[r] : trash_off), c_make(0.0, 0.0));   // "panel -1"
            }
#pragma unroll
            for (int w = 0; w < MMH_BOX_HPT; w++)
                if (hact[w]) sts_c128_b(sbase + halo_off + 16u * (unsigned)(w * T + tid), __ldcg(G + hgo[w]));
            __syncthreads();

            // one panel step: new = (b_i P1 + A_ii sqrt(s-1) P2 + sum_j coef_j nb_j) / sqrt(s); the result replaces P2.  The halo
            // of panel s (the lower boxes' amplitudes, final since those boxes were marched before this one) is fetched at the top
            // of the step and stored beside the own cells at its end.
#define MMH_BOX_STEP(P1, P2, BPREV, BCUR)                                                                   \
            {                                                                                               \
                c128 hv[MMH_BOX_HPT];                                                                       \
                _Pragma("unroll") for (int w = 0; w < MMH_BOX_HPT; w++)                                     \
                    hv[w] = hact[w] ? __ldcg(gpan + hgo[w]) : c_make(0.0, 0.0);                             \
                c128 v[R];                                                                                  \
                const c128 as = c_scale(a00, sqm);                                                          \
                _Pragma("unroll") for (int r = 0; r < R; r++) {                                             \
                    c128 nbv[NPD];                                                                          \
                    _Pragma("unroll") for (int jj = 0; jj < NPD; jj++) nbv[jj] = lds_c128_b((BPREV) + nbo[r][jj]); \
                    v[r] = c_mul(b0, P1[r]);                                                                \
                    v[r] = c_add(v[r], c_mul(as, P2[r]));                                                   \
                    _Pragma("unroll") for (int jj = 0; jj < NPD; jj++) v[r] = c_add(v[r], c_mul(coef[r][jj], nbv[jj])); \
                }                                                                                           \
                div_all_box<R>(v, sqs, rsqs);                                                               \
                _Pragma("unroll") for (int r = 0; r < R; r++) {                                             \
                    P2[r] = v[r];                                                                           \
                    sts_c128_b((BCUR) + loco[r], v[r]);                                                     \
                    if (act[r]) gpan[gofs[r]] = v[r];                                                       \
                }                                                                                           \
                _Pragma("unroll") for (int w = 0; w < MMH_BOX_HPT; w++)                                     \
                    if (hact[w]) sts_c128_b((BCUR) + halo_off + 16u * (unsigned)(w * T + tid), hv[w]);      \
                gpan += P;                                                                                  \
                __syncthreads();                                                                            \
            }

            c128 *gpan = G + P;   // panel s
            const unsigned B0 = sbase, B1 = sbase + bstride;
            double sqm = 0.0;
            double2 tq = sqtab[S > 1 ? 1 : 0];
            double sqs = tq.x, rsqs = tq.y;
            int s = 1;
            if constexpr (HIST) {
                unsigned bpp = sbase + 2u * bstride, bprev = sbase, bcur = sbase + bstride;   // panels s-2, s-1, s
#pragma unroll 1
                for (; s < S; s++) {
                    const double2 tn = sqtab[s + 1 < S ? s + 1 : s];
                    c128 hv[MMH_BOX_HPT];
#pragma unroll
                    for (int w = 0; w < MMH_BOX_HPT; w++) hv[w] = hact[w] ? __ldcg(gpan + hgo[w]) : c_make(0.0, 0.0);
                    c128 v[R];
                    const c128 as = c_scale(a00, sqm);
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        c128 nbv[NPD];
                        const c128 P1 = lds_c128_b(bprev + loco[r]), P2 = lds_c128_b(bpp + loco[r]);
#pragma unroll
                        for (int jj = 0; jj < NPD; jj++) nbv[jj] = lds_c128_b(bprev + nbo[r][jj]);
                        v[r] = c_mul(b0, P1);
                        v[r] = c_add(v[r], c_mul(as, P2));
#pragma unroll
                        for (int jj = 0; jj < NPD; jj++) v[r] = c_add(v[r], c_mul(coef[r][jj], nbv[jj]));
                    }
                    div_all_box<R>(v, sqs, rsqs);
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        sts_c128_b(bcur + (act[r] ? loco[r] : trash_off), v[r]);
                        if (act[r]) gpan[gofs[r]] = v[r];
                    }
#pragma unroll
                    for (int w = 0; w < MMH_BOX_HPT; w++)
                        if (hact[w]) sts_c128_b(bcur + halo_off + 16u * (unsigned)(w * T + tid), hv[w]);
                    gpan += P;
                    sqm = sqs; sqs = tn.x; rsqs = tn.y;
                    const unsigned tmp = bpp; bpp = bprev; bprev = bcur; bcur = tmp;
                    __syncthreads();
                }
            }
#pragma unroll 1
            for (; s + 1 < S; s += 2) {
                const double2 t1 = sqtab[s + 1];
                const double2 t2 = sqtab[s + 2 < S ? s + 2 : s + 1];
                MMH_BOX_STEP(h1, h0, B0, B1)
                sqm = sqs; sqs = t1.x; rsqs = t1.y;
                MMH_BOX_STEP(h0, h1, B1, B0)
                sqm = sqs; sqs = t2.x; rsqs = t2.y;
            }
            if (s < S) MMH_BOX_STEP(h1, h0, B0, B1)
#undef MMH_BOX_STEP
            __threadfence();   // this box's amplitudes are the next boxes' halos
        }
    }
}

// plan: boxes of <= 2 * 512 points (R = 2) over the first <= 3 panel dims, halo <= MMH_BOX_HPT * threads
// slots per thread: 2 = the two previous panels in registers (default), 4 = in shared memory (HIST, MMH_BOX_R=4).  Measured,
// 592 lattices: (20,)^4 0.693 / 0.751 ms, (30,)^4 3.85 / 3.91 ms, (64,)^3 0.862 / 0.808 ms (R = 2 / R = 4): twice the chains in
// flight buy nothing, an SM marches ~1000 points per microsecond either way.
int mmh_box_slots() {
    const char *e = mmh_getenv("MMH_BOX_R");
    return (e && atoi(e) == 4) ? 4 : 2;
}

bool mmh_plan_march_box(const LatticeDesc &d, int stage, BoxParams *bp, int *T_out, size_t *smem_out) {
    const int Rb = mmh_box_slots();
    const int npd = d.D - 1 - stage;
    if (npd < 1 || npd > 4) return false;
    const long long P = d.strides[stage];
    if (P >= (1LL << 31)) return false;
    const int S = d.shape[stage];
    const int nt = npd < 3 ? npd : 3;
    const long long inner = d.strides[stage + nt];
    if (inner > 1024) return false;
    int shp[3] = { 1, 1, 1 };
    for (int m = 0; m < nt; m++) shp[m] = d.shape[stage + 1 + m];
    // cost ~ boxes x (points + a fixed step overhead of ~300 point-times), x 3 when the lattice's last index is cut: short rows
    // (10 points = 160 bytes) leave partially written 64-byte DRAM atoms, and once a batch outgrows L2 (126 MB) those cost
    // read-modify-writes (148 x (20,)^4 in 10 x 10 x 10 boxes: 0.69 ms; in 5 x 10 x 20 boxes: 0.20 ms; 32 lattices, L2
    // resident: 0.19 ms either way)
    long long max_ts = 1024;
    if (const char *e_ = mmh_getenv("MMH_BOX_MAXTS")) max_ts = atoll(e_);   // tuning hook
    if (inner > max_ts) return false;
    double best_cost = 1e300;
    long long best_tiles = -1;
    int bg[3] = { 1, 1, 1 }, bT = 0, bLS = 0;
    for (int g0 = 1; g0 <= shp[0]; g0++) {
        if (300.0 * g0 >= best_cost) break;
        for (int g1 = 1; g1 <= shp[1]; g1++) {
            if (300.0 * g0 * g1 >= best_cost) break;
            for (int g2 = 1; g2 <= shp[2]; g2++) {
                const int g[3] = { g0, g1, g2 };
                long long e[3], TS = inner;
                for (int m = 0; m < 3; m++) { e[m] = (shp[m] + g[m] - 1) / g[m]; TS *= e[m]; }
                if (TS > max_ts) continue;
                long long HC = 0;
                for (int m = 0; m < 3; m++) if (g[m] > 1) HC += TS / e[m];
                int T = (int)((TS + Rb - 1) / Rb + 31) / 32 * 32;
                if (T < 64) T = 64;
                if (HC > (long long)MMH_BOX_HPT * T) continue;
                const long long tiles = (long long)g0 * g1 * g2;
                const bool cut_last = inner == 1 && g[nt - 1] > 1;
                const double cost = (double)tiles * (double)(TS + 300) * (cut_last ? 3.0 : 1.0) + 1e-3 * (double)HC;
                if (cost < best_cost) {
                    best_cost = cost; best_tiles = tiles; bg[0] = g0; bg[1] = g1; bg[2] = g2; bT = T; bLS = (int)(TS + HC + 2);
                }
                break;   // a finer cut of the last boxed dim only adds boxes
            }
        }
    }
    if (best_tiles < 0) return false;
    const size_t smem = sizeof(c128) * ((size_t)(Rb == 4 ? 3 : 2) * bLS + (size_t)S);
    if (smem > 200 * 1024) return false;
    memset(bp, 0, sizeof(*bp));
    bp->d = d; bp->stage = stage; bp->nt = nt; bp->ls = bLS;
    for (int m = 0; m < 3; m++) bp->g[m] = bg[m];
    *T_out = bT; *smem_out = smem;
    return true;
}

// grid: one CTA per lattice up to one resident wave (the CTAs then take further lattices in turn)
template <int R, int NPD, bool HIST>
static cudaError_t launch_box(const BoxParams &p, int sm_count, int T, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_march_box<R, NPD, HIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_march_box<R, NPD, HIST>, T, smem) != cudaSuccess || occ < 1) occ = 1;
    long long grid = (long long)occ * sm_count;
    if (grid > p.batch) grid = p.batch;
    k_march_box<R, NPD, HIST><<<(unsigned)grid, T, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t mmh_launch_march_box(const BoxParams &p, int sm_count, int T, size_t smem, cudaStream_t st) {
    const int npd = p.d.D - 1 - p.stage;
    const bool hist = mmh_box_slots() == 4;
    switch (npd) {
        case 1: return hist ? launch_box<4, 1, true>(p, sm_count, T, smem, st) : launch_box<2, 1, false>(p, sm_count, T, smem, st);
        case 2: return hist ? launch_box<4, 2, true>(p, sm_count, T, smem, st) : launch_box<2, 2, false>(p, sm_count, T, smem, st);
        case 3: return hist ? launch_box<4, 3, true>(p, sm_count, T, smem, st) : launch_box<2, 3, false>(p, sm_count, T, smem, st);
        case 4: return hist ? launch_box<4, 4, true>(p, sm_count, T, smem, st) : launch_box<2, 4, false>(p, sm_count, T, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

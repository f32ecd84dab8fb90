// mmh_lanes.cu — warp-synchronous batched march of stage D-2 (sm_100a): the whole of a batch of 2-index lattices (cfg3,
// vanilla_batch_numba over 2-mode kets, vanilla/batch.py:27-61) and the first marched stage of deeper batched lattices.
//
// The only neighbour of stage D-2 outside the marched index is k - e_{D-2} - e_{D-1}: the previous panel one position to the
// left (vanilla/core.py:97-104 with i = D-2).  k_march_stage (mmh_march.cu) reads it from a shared-memory copy of the panel and
// pays one CTA barrier per step -- the largest stall of the cfg3 profile (2.4 of 8.9 warp cycles per issue).  Here a lane owns
// R CONSECUTIVE panel positions, so that neighbour is the lane's own register for R-1 of its R positions and one warp shuffle
// for the first; ln = ceil(n1 / R) lanes hold one lattice row, a warp marches Lw = 32 / ln lattices, and nothing in the step
// waits for another warp.  The amplitudes leave through a per-warp shared-memory transposition (lane-major in, warp-linear out:
// one __syncwarp) so that the lattice stores stay coalesced along the last index, 16 B per amplitude written once.
// The arithmetic is k_march_stage's, operation for operation (bit-identical results).
#include <cstring>

#include "mmh_params.cuh"

__device__ __forceinline__ void pdl_launch_dependents_l() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <int R>
__device__ __forceinline__ void div_all_inplace_l(c128 (&v)[R], double sqs, double rsqs) {
    bool slow = false;
#pragma unroll
    for (int r = 0; r < R; r++) slow |= div_needs_slow(v[r].x) | div_needs_slow(v[r].y);
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

// transposition buffer index of warp-linear cell q: R odd is conflict free as it is (8 lanes x R cells hit 8 distinct 16-byte
// bank groups), R even is skewed by one cell per 8
template <int R>
__device__ __forceinline__ int xp_index(int q) { return (R & 1) ? q : q + (q >> 3); }

#define MMH_LANES_XP(R) (36 * (R))   // cells of one transposition buffer

// FUSE: the CTA first computes stage D-1 of its nw * Lw lattices itself -- the 1-D chain G[n] = (b G[n-1] + A sqrt(n-1) G[n-2]) /
// sqrt(n) along the last index (vanilla/core.py:85-104 with i = D-1; arithmetic of k_fwd_chain, operation for operation), one
// thread per lattice, into a shared-memory row per lattice -- instead of a separate chain kernel writing panel 0 to HBM and this
// kernel reading it back: one launch per batch, no 39-step latency-bound kernel in front of the march (cfg3: 24 of 302 us), and
// panel 0 leaves through the same coalesced transposition as every other panel.
// smem: sqtab[max(S, FUSE ? n1 : 0)] double2 | xp[warps][2][MMH_LANES_XP(R)] c128 | FUSE: chain[nw * Lw][n1 | 1] c128
template <int R, bool FUSE>
__global__ void __launch_bounds__(128) k_march_lanes(StageParams p, int ln, int Lw) {
    extern __shared__ c128 smem[];
    const LatticeDesc &d = p.d;
    const int D = d.D, i = D - 2;
    const int n1 = d.shape[D - 1], S = d.shape[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ntab = (FUSE && n1 > S) ? n1 : S;
    double2 *sqt = (double2 *)smem;
    c128 *xp = smem + ntab + (size_t)warp * 2 * MMH_LANES_XP(R);
    c128 *chain = smem + ntab + (size_t)nw * 2 * MMH_LANES_XP(R);
    const int cpitch = n1 | 1;
    pdl_launch_dependents_l();
    for (int n = threadIdx.x; n < ntab; n += blockDim.x) sqt[n] = make_double2(p.sq[n], p.rsq[n]);
    pdl_wait();   // the caller's buffers (A, b, c, G) are touched from here on
    __syncthreads();
    if (FUSE) {
        const long long cl = (long long)blockIdx.x * nw * Lw + threadIdx.x;   // this thread's chain
        if ((int)threadIdx.x < nw * Lw && cl < p.batch) {
            const c128 Ac = p.A[cl * D * D + (D - 1) * D + (D - 1)], bc = p.b[cl * D + (D - 1)];
            c128 *row = chain + (size_t)threadIdx.x * cpitch;
            c128 p1 = p.c[cl], aterm = c_make(0.0, 0.0);   // pipelined as in k_fwd_chain (mmh_march.cu): same operations, same order
            row[0] = p1;
            double2 t = sqt[n1 > 1 ? 1 : 0];
            for (int s = 1; s < n1; s++) {
                const double2 tn = sqt[s + 1 < n1 ? s + 1 : s];
                c128 v = c_mul(bc, p1);
                if (s >= 2) v = c_add(v, aterm);
                aterm = c_mul(c_scale(Ac, t.x), p1);
                v = c_div_table_spec(v, t.x, t.y);
                row[s] = v;
                p1 = v; t = tn;
            }
        }
        __syncthreads();
    }
    const long long wl0 = ((long long)blockIdx.x * nw + warp) * Lw;   // first lattice of this warp
    if (wl0 >= p.batch) return;
    const int nlat = (int)(p.batch - wl0 < Lw ? p.batch - wl0 : Lw);

    // ---- this lane's R consecutive positions ----
    const int lw = lane / ln;
    const bool lane_act = lw < nlat;
    const int k0 = (lane - lw * ln) * R;
    const long long l = wl0 + (lane_act ? lw : 0);
    const c128 b0 = p.b[l * D + i], a00 = p.A[l * D * D + i * D + i], a01 = p.A[l * D * D + i * D + i + 1];
    c128 h0[R], h1[R], coef[R];
    {
        // panel 0 (k_i = 0): this CTA's chain rows (FUSE) or what the chain kernel wrote
        const c128 *g0 = FUSE ? chain + (size_t)(warp * Lw + (lane_act ? lw : 0)) * cpitch : p.G + l * p.lat_stride;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int k1 = k0 + r;
            const bool act = lane_act && k1 < n1;
            h0[r] = c_make(0.0, 0.0);
            h1[r] = act ? g0[k1] : c_make(0.0, 0.0);
            coef[r] = (act && k1 > 0) ? c_scale(a01, p.sq[k1]) : c_make(0.0, 0.0);   // A_ij sqrt(k_j), core.py:103
        }
    }
    // ---- the cells this lane stores after the transposition: q = j * 32 + lane (warp-linear) ----
    unsigned go[R];   // element offset from G + wl0 * lat_stride (panel 0), ~0u: padding cell
#pragma unroll
    for (int j = 0; j < R; j++) {
        const int q = j * 32 + lane;
        const int ql = q / R, qr = q - ql * R;
        const int qw = ql / ln;
        const int k1 = (ql - qw * ln) * R + qr;
        go[j] = (qw < nlat && k1 < n1) ? (unsigned)((long long)qw * p.lat_stride + k1) : 0xFFFFFFFFu;
    }
    c128 *gs = p.G + wl0 * p.lat_stride;
    const bool first = k0 == 0;   // position k_{D-1} = 0: no left neighbour (coefficient 0, own value as in k_march_stage)

    // one panel step: new = (b_i P1 + A_ii sqrt(s-1) P2 + coef nb) / sqrt(s); the result replaces P2
#define MMH_LANES_STEP(P1, P2, XB)                                                                      \
    {                                                                                                   \
        c128 v[R];                                                                                      \
        const c128 up = make_double2(__shfl_up_sync(0xffffffffu, P1[R - 1].x, 1),                       \
                                     __shfl_up_sync(0xffffffffu, P1[R - 1].y, 1));                      \
        const c128 as = c_scale(a00, sqm);                                                              \
        _Pragma("unroll") for (int r = 0; r < R; r++) {                                                 \
            const c128 nb = r == 0 ? (first ? P1[0] : up) : P1[r > 0 ? r - 1 : 0];                      \
            v[r] = c_mul(b0, P1[r]);                                                                    \
            v[r] = c_add(v[r], c_mul(as, P2[r]));                                                       \
            v[r] = c_add(v[r], c_mul(coef[r], nb));                                                     \
        }                                                                                               \
        div_all_inplace_l<R>(v, sqs, rsqs);                                                             \
        gs += n1;                                                                                       \
        _Pragma("unroll") for (int r = 0; r < R; r++) {                                                 \
            P2[r] = v[r];                                                                               \
            (XB)[xp_index<R>(lane * R + r)] = v[r];                                                     \
        }                                                                                               \
        __syncwarp();                                                                                   \
        _Pragma("unroll") for (int j = 0; j < R; j++) {                                                 \
            const c128 w = (XB)[xp_index<R>(j * 32 + lane)];                                            \
            if (go[j] != 0xFFFFFFFFu) gs[go[j]] = w;                                                    \
        }                                                                                               \
    }

    c128 *xb0 = xp, *xb1 = xp + MMH_LANES_XP(R);
    if (FUSE) {   // panel 0 leaves like every other panel: lane-major in, warp-linear (coalesced) out
#pragma unroll
        for (int r = 0; r < R; r++) xb1[xp_index<R>(lane * R + r)] = h1[r];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < R; j++) {
            const c128 w = xb1[xp_index<R>(j * 32 + lane)];
            if (go[j] != 0xFFFFFFFFu) gs[go[j]] = w;
        }
        __syncwarp();
    }
    double sqm = 0.0;
    double2 t = sqt[S > 1 ? 1 : 0];
    double sqs = t.x, rsqs = t.y;
    int s = 1;
#pragma unroll 1
    for (; s + 1 < S; s += 2) {
        const double2 t1 = sqt[s + 1];
        const double2 t2 = sqt[s + 2 < S ? s + 2 : s + 1];
        MMH_LANES_STEP(h1, h0, xb0)
        sqm = sqs; sqs = t1.x; rsqs = t1.y;
        MMH_LANES_STEP(h0, h1, xb1)
        sqm = sqs; sqs = t2.x; rsqs = t2.y;
    }
    if (s < S) MMH_LANES_STEP(h1, h0, xb0)
#undef MMH_LANES_STEP
}

// plan: R positions per lane, ln lanes per lattice row, Lw lattices per warp; false when a row does not fit one warp
bool mmh_plan_march_lanes(int n1, int *R_out, int *ln_out, int *Lw_out) {
    double best = -1.0;
    const char *eR = mmh_getenv("MMH_LANES_R");
    for (int R = 2; R <= 8; R++) {
        if (eR && atoi(eR) != R) continue;
        const int ln = (n1 + R - 1) / R;
        if (ln > 32) continue;
        const int Lw = 32 / ln;
        // slot efficiency; a slight preference for 4-5 positions per lane (enough independent chains, no register spills)
        const double score = (double)Lw * n1 / (32.0 * R) - 0.01 * (R > 5 ? R - 5 : (R < 4 ? 4 - R : 0));
        if (score > best) { best = score; *R_out = R; *ln_out = ln; *Lw_out = Lw; }
    }
    return best > 0.0;
}

cudaError_t mmh_launch_march_lanes(const StageParams &p, int R, int ln, int Lw, int sm_count, cudaStream_t st) {
    // 4 warps per CTA; 2 when the batch leaves the device less than ~8 CTAs per SM (sharded sweeps: 8,192 triples per GPU are
    // 512 four-warp CTAs = 3.46 per SM, i.e. the SMs holding 4 set the time: +16%; 1,024 two-warp CTAs balance to 1%)
    int block = 128;
    if ((p.batch + 4LL * Lw - 1) / (4LL * Lw) < 8LL * sm_count) block = 64;
    if (const char *e = mmh_getenv("MMH_LANES_BLOCK")) block = atoi(e) == 64 ? 64 : (atoi(e) == 32 ? 32 : 128);
    const int nw = block / 32;
    const long long grid = (p.batch + (long long)nw * Lw - 1) / ((long long)nw * Lw);
    if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    const int S = p.d.shape[p.d.D - 2], n1 = p.d.shape[p.d.D - 1];
    const bool fuse = p.fuse_chain != 0;
    const bool pdl = !mmh_getenv("MMH_NO_PDL");
    const size_t ntab = (fuse && n1 > S) ? n1 : S;
#define MMH_LAUNCH(N, F)                                                                                               \
    {                                                                                                                  \
        const size_t smem = sizeof(c128) * (ntab + (size_t)nw * 2 * MMH_LANES_XP(N) +                                  \
                                            ((F) ? (size_t)nw * Lw * (size_t)(n1 | 1) : 0));                           \
        if (smem > 200 * 1024) return cudaErrorInvalidValue;                                                           \
        if (smem > 48 * 1024) {                                                                                        \
            cudaError_t e = cudaFuncSetAttribute(k_march_lanes<N, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                            \
        }                                                                                                              \
        return mmh_launch_ex(k_march_lanes<N, F>, dim3((unsigned)grid), dim3(block), smem, st, pdl, p, ln, Lw);        \
    }
#define MMH_CASE(N) case N: if (fuse) MMH_LAUNCH(N, true) else MMH_LAUNCH(N, false)
    switch (R) {
        MMH_CASE(2) MMH_CASE(3) MMH_CASE(4) MMH_CASE(5) MMH_CASE(6) MMH_CASE(7) MMH_CASE(8)
        default: return cudaErrorInvalidValue;
    }
#undef MMH_CASE
#undef MMH_LAUNCH
}

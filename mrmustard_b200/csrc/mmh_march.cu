// mmh_march.cu — register/shared-memory resident "panel march" kernels for the forward path (sm_100a).
//
// Stage 0 of the panel schedule (see mmh_forward.cu) carries almost all of the work: for every step
// s = k_0 = 1 .. shape[0]-1 the panel of P0 = strides[0] points is computed from the panels s-1 and s-2.
// A thread OWNS fixed panel positions for the whole march, so
//   * G[k - e_0] and G[k - 2 e_0] (same position, previous two panels) live in registers,
//   * G[k - e_0 - e_j] (j >= 1) comes from a shared-memory copy of panel s-1 (double buffered),
//   * A_0j sqrt(k_j) comes from a per-lattice shared-memory table,
//   * every amplitude is written to HBM exactly once (coalesced along the last mode) and never re-read.
//
// K2  k_march_stage : one CTA marches L small lattices in lock step (batched path, cfg3).
//     k_fwd_chain   : stage D-1 (a 1-D chain), one thread per lattice.
//     k_warp_tail   : the two trailing stages of ONE lattice by one warp.
// K1  (one lattice over many tile-owner CTAs) lives in mmh_tiled.cu.
#include <cstdlib>
#include <cstring>

#include "mmh_params.cuh"
#include "mmh_points.cuh"


// Programmatic dependent launch (sm_90+): a kernel may let its successor in the stream start its prologue early
// (launch_dependents) and the successor blocks at `wait` until the predecessor grid has completed and flushed.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Divide all R numerators by sqrt(s) in place.  The IEEE slow path (inf/nan/subnormal-range numerators) is a
// single warp-level branch for the whole step, so the common path stays branch-free and v's registers are reused.
template <int R>
__device__ __forceinline__ void div_all_inplace(c128 (&v)[R], double sqs, double rsqs) {
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

// ---------------------------------------------------------------------------------------------------
// K2: batched stage march.  One launch per stage i (i = D-2 .. 0); stage D-1 is k_fwd_chain.
// The CTA marches L lattices in lock step; a thread owns R fixed panel positions ("slots").
// ---------------------------------------------------------------------------------------------------
// smem layout (c128 units): sba[L][2] = (b_i, A_ii) | buf[2][L*P]
#ifndef MMH_K2_MAXT
#define MMH_K2_MAXT(R) ((R) >= 4 ? 256 : 512)
#define MMH_K2_MINB(R) ((R) >= 4 ? 2 : 1)
#endif
template <int R, int NPD>
__global__ void __launch_bounds__(MMH_K2_MAXT(R), MMH_K2_MINB(R)) k_march_stage(StageParams p) {
    extern __shared__ c128 smem[];
    const LatticeDesc &d = p.d;
    const int D = d.D;
    const int i = p.stage;            // NPD == D - 1 - i
    const int L = p.L;
    const int P = (int)d.strides[i];
    const int S = d.shape[i];
    const int T = blockDim.x;
    const int tid = threadIdx.x;
    c128 *sba = smem;
    c128 *buf = smem + 2 * L;
    const int LP = L * P;
    const double *__restrict__ sq = p.sq;
    const double *__restrict__ rsq = p.rsq;
    const long long lat0 = (long long)blockIdx.x * L;
    const int nlat = (int)((p.batch - lat0) < L ? (p.batch - lat0) : L);
    const int nslots = nlat * P;

    pdl_launch_dependents();
    for (int t = tid; t < nlat; t += T) {
        sba[2 * t] = p.b[(lat0 + t) * D + i];
        sba[2 * t + 1] = p.A[(lat0 + t) * D * D + i * D + i];
    }
    if (p.fuse_chain) {
        // single-lattice path: stage D-1 (the 1-D chain) is done here by one thread per lattice instead of a
        // separate launch; its amplitudes are panel 0 of this stage (global memory, visible after the barrier)
        if (tid < nlat) {
            const long long l = lat0 + tid;
            const int ic = D - 1;
            const c128 Ac = p.A[l * D * D + ic * D + ic], bc = p.b[l * D + ic];
            c128 *g = p.G + l * p.lat_stride;
            const int Sc = d.shape[ic];
            c128 p1 = p.c[l], p2 = c_make(0.0, 0.0);
            g[0] = p1;
            for (int s = 1; s < Sc; s++) {
                c128 v = c_mul(bc, p1);
                if (s >= 2) v = c_add(v, c_mul(c_scale(Ac, sq[s - 1]), p2));
                v = c_div_table(v, sq[s], rsq[s]);
                g[s] = v;
                p2 = p1; p1 = v;
            }
        }
        __syncthreads();
    } else {
        pdl_wait();   // panel 0 was written by the previous launch in the stream
    }

    // ---- per-slot constants (registers) -------------------------------------------------------------------
    bool act[R];
    int loc[R];                 // index inside a panel buffer (l * P + f)
    int isb[R];                 // 2 * l  -> sba
    int nbi[R][NPD];            // buffer index of the neighbour k - e_i - e_j (own index when k_j == 0, coef = 0)
    c128 coef[R][NPD];          // A_ij sqrt(k_j)  (core.py:103), constant along the march
    c128 *gp[R];                // -> G[lattice][k_i = s][f]
    c128 h0[R], h1[R];          // the two previous panels at this position (ping-pong)
    int lst[NPD];
#pragma unroll
    for (int jj = 0; jj < NPD; jj++) lst[jj] = (int)d.strides[i + 1 + jj];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int q = r * T + tid;
        act[r] = q < nslots;
        loc[r] = act[r] ? q : 0;
        const int l = loc[r] / P, f = loc[r] - l * P;
        isb[r] = 2 * l;
        gp[r] = p.G + ((lat0 + l) * p.lat_stride + f);
        const c128 *Arow = p.A + ((lat0 + l) * D * D + i * D + i + 1);
        int rem = f;
#pragma unroll
        for (int jj = 0; jj < NPD; jj++) {
            const int k = rem / lst[jj];
            rem -= k * lst[jj];
            const bool has = act[r] && k > 0;
            nbi[r][jj] = has ? loc[r] - lst[jj] : loc[r];
            coef[r][jj] = has ? c_scale(Arow[jj], sq[k]) : c_make(0.0, 0.0);
        }
        h0[r] = c_make(0.0, 0.0);
        h1[r] = act[r] ? *gp[r] : c_make(0.0, 0.0);   // panel 0 (k_i = 0) was written by the previous stage
        if (act[r]) buf[loc[r]] = h1[r];
    }
    __syncthreads();

    // one panel step: new = (b_i P1 + A_ii sqrt(s-1) P2 + sum_j coef_j nb_j) / sqrt(s); the result replaces P2.
    // Branch-free per slot so that the R (x2 re/im) dependency chains interleave; the exact-division slow
    // path (inf/nan/subnormal-range numerators) is checked once per step for all slots.
#define MMH_MARCH_STEP(P1, P2, OFFP, OFFC)                                                            \
    {                                                                                                 \
        c128 v[R];                                                                                    \
        _Pragma("unroll") for (int r = 0; r < R; r++) {                                               \
            const c128 b0 = sba[isb[r]], a00 = sba[isb[r] + 1];                                       \
            v[r] = c_mul(b0, P1[r]);                                                                  \
            v[r] = c_add(v[r], c_mul(c_scale(a00, sqm), P2[r]));                                      \
            _Pragma("unroll") for (int jj = 0; jj < NPD; jj++)                                        \
                v[r] = c_add(v[r], c_mul(coef[r][jj], buf[(OFFP) + nbi[r][jj]]));                     \
        }                                                                                             \
        div_all_inplace<R>(v, sqs, rsqs);                                                             \
        _Pragma("unroll") for (int r = 0; r < R; r++) {                                               \
            P2[r] = v[r];                                                                             \
            if (act[r]) {                                                                             \
                gp[r] += P;                                                                           \
                *gp[r] = v[r];                                                                        \
                buf[(OFFC) + loc[r]] = v[r];                                                          \
            }                                                                                         \
        }                                                                                             \
    }

    double sqm = 0.0, sqs = sq[S > 1 ? 1 : 0], rsqs = rsq[S > 1 ? 1 : 0];
    int s = 1;
    for (; s + 1 < S; s += 2) {
        const double sq1 = sq[s + 1], rsq1 = rsq[s + 1];
        const int s2 = s + 2 < S ? s + 2 : s + 1;
        const double sq2 = sq[s2], rsq2 = rsq[s2];
        MMH_MARCH_STEP(h1, h0, 0, LP)
        __syncthreads();
        sqm = sqs; sqs = sq1; rsqs = rsq1;
        MMH_MARCH_STEP(h0, h1, LP, 0)
        __syncthreads();
        sqm = sqs; sqs = sq2; rsqs = rsq2;
    }
    if (s < S) MMH_MARCH_STEP(h1, h0, 0, LP)
#undef MMH_MARCH_STEP
}

// thread-per-lattice chain: stage D-1 (k_<D-1 = 0), G[n] = (b G[n-1] + A sqrt(n-1) G[n-2]) / sqrt(n).
// For D == 1 this is the whole lattice (cfg1).
// dynamic shared memory: (sqrt, 1/sqrt) table of the chain when it fits (tab != 0), see k_warp_tail
__global__ void __launch_bounds__(128) k_fwd_chain(FwdParams p, int tab) {
    extern __shared__ double2 sqt_chain[];
    pdl_launch_dependents();
    const int D = p.d.D, i = D - 1;
    const int S = p.d.shape[i];
    if (tab) {
        for (int n = threadIdx.x; n < S; n += blockDim.x) sqt_chain[n] = make_double2(p.sq[n], p.rsq[n]);
        __syncthreads();
    }
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= p.batch) return;
    const c128 A = p.A[l * D * D + i * D + i], b = p.b[l * D + i];
    c128 *g = p.G + l * p.d.N;
    // software pipelined like the chain of k_warp_tail: the A-term of step s+1, (A sqrt(s)) G[s-1], and the table entry of step
    // s+1 are issued beside the quotient of step s, and the quotient's range test resolves beside its fast sequence -- the
    // dependent chain of a step is b G (2 levels), one add and the 5-level quotient (cfg1: 134 -> ~60 ns per step).
    // Same operations in the same order as the plain loop: bit-identical.
#define MMH_CHAIN_TAB(n) (tab ? sqt_chain[(n)] : make_double2(p.sq[(n)], p.rsq[(n)]))
    c128 p1 = p.c[l], aterm = c_make(0.0, 0.0);
    g[0] = p1;
    double2 t = MMH_CHAIN_TAB(S > 1 ? 1 : 0);
    for (int s = 1; s < S; s++) {
        const double2 tn = MMH_CHAIN_TAB(s + 1 < S ? s + 1 : s);
        c128 v = c_mul(b, p1);
        if (s >= 2) v = c_add(v, aterm);
        aterm = c_mul(c_scale(A, t.x), p1);
        v = c_div_table_spec(v, t.x, t.y);
        g[s] = v;
        p1 = v; t = tn;
    }
#undef MMH_CHAIN_TAB
}

// The same chain for a batch, staged in shared memory: a thread that stores its own amplitude every step writes 16 bytes a
// lattice stride away from its neighbour's (2.6 M half-sector writes for cfg3, 28 us).  Here the CTA's T chains advance 8 steps
// at a time into a double-buffered [T][8] tile (pitch 9: conflict free both ways) that leaves as 128-byte row segments, eight
// lanes per lattice.  The tile is small (18 KB per buffer pair), so every chain of a 65,536-lattice batch is resident at once:
// with whole rows staged (64 KB per CTA) the batch needed two waves at two warps per scheduler, 24 us, issue-latency bound.
// dynamic shared memory: sqtab[S] double2 | tile[2][T][9] c128
#define MMH_CHAIN_CH 8
__global__ void __launch_bounds__(128) k_fwd_chain_rows(FwdParams p) {
    extern __shared__ double2 sqt_rows[];
    pdl_launch_dependents();
    const int D = p.d.D, i = D - 1;
    const int S = p.d.shape[i], T = blockDim.x;
    constexpr int CH = MMH_CHAIN_CH, PITCH = CH + 1;
    c128 *tile = (c128 *)(sqt_rows + S);
    for (int n = threadIdx.x; n < S; n += T) sqt_rows[n] = make_double2(p.sq[n], p.rsq[n]);
    __syncthreads();
    const long long lat0 = (long long)blockIdx.x * T;
    const long long l = lat0 + threadIdx.x;
    const bool act = l < p.batch;
    const int nlat = (int)(p.batch - lat0 < T ? p.batch - lat0 : T);
    const long long lc = act ? l : lat0;
    const c128 A = p.A[lc * D * D + i * D + i], b = p.b[lc * D + i];
    c128 p1 = p.c[lc], aterm = c_make(0.0, 0.0);
    for (int c0 = 0, buf = 0; c0 < S; c0 += CH, buf ^= 1) {
        const int cw = S - c0 < CH ? S - c0 : CH;
        c128 *mine = tile + ((size_t)buf * T + threadIdx.x) * PITCH;
        for (int j = 0; j < cw; j++) {
            const int s = c0 + j;
            if (s > 0) {   // pipelined as in k_fwd_chain: aterm = (A sqrt(s-1)) G[s-2] was formed one step earlier
                const double2 t = sqt_rows[s];
                c128 v = c_mul(b, p1);
                if (s >= 2) v = c_add(v, aterm);
                aterm = c_mul(c_scale(A, t.x), p1);
                v = c_div_table_spec(v, t.x, t.y);
                p1 = v;
            }
            mine[j] = p1;
        }
        __syncthreads();   // one barrier per chunk: the other buffer is rewritten only after everybody has passed this one
        const c128 *src = tile + (size_t)buf * T * PITCH;
        for (int idx = threadIdx.x; idx < nlat * cw; idx += T) {
            const int r = idx / cw, k = idx - r * cw;
            p.G[(lat0 + r) * p.d.N + c0 + k] = src[r * PITCH + k];
        }
    }
}

// launch with (pdl = true) or without programmatic stream serialization
template <typename P>
static cudaError_t launch_pdl(void (*kern)(P), int grid, int block, size_t smem, cudaStream_t st, bool pdl, const P &p) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

// ---------------------------------------------------------------------------------------------------
// One warp, no shared-memory hand-off, no barrier: the two trailing stages of ONE lattice.
//   stage D-1 (the 1-D chain, lane 0) and stage D-2 (panels of shape[D-1] <= 64 points, two per lane).
// The only neighbour of stage D-2 is k - e_{D-2} - e_{D-1}: the previous panel's value one lane below, i.e. a
// warp shuffle.  A step is then one dependent chain (shuffle, 2 + 1 + 5 FP64 levels, store): ~0.06 us instead of
// the ~0.22 us of the shared-memory + barrier step, and both stages are latency bound (98 dependent steps of cfg2).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ c128 shfl_up_c128(c128 v, int delta) {
    return make_double2(__shfl_up_sync(0xffffffffu, v.x, delta), __shfl_up_sync(0xffffffffu, v.y, delta));
}
__device__ __forceinline__ c128 shfl_c128(c128 v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

__device__ __forceinline__ void stg_strong(c128 *p, c128 v) {
    asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// (Round 2 also measured an anti-diagonal wavefront of the same two stages -- lane l owns rows 2l, 2l+1 and step t computes the cells
//  with k_{D-2} + k_{D-1} = t, so that the chain and the march overlap: bit-identical, but 32.7 us instead of 15.8 us on cfg2.  One
//  warp issues ~200 instructions per step either way (~0.3 us); the wavefront needs Sc + S2 such steps, this kernel Sc cheap
//  one-cell steps of 0.11 us followed by S2 of 0.2 us.)
__global__ void __launch_bounds__(32) k_warp_tail(StageParams p) {
    __shared__ double2 sqt[64];
    __shared__ c128 chain[64];
    const LatticeDesc &d = p.d;
    const int D = d.D, ic = D - 1, i2 = D - 2;
    const int Sc = d.shape[ic], S2 = d.shape[i2];
    const int lane = threadIdx.x;
    const double *__restrict__ sq = p.sq;
    const double *__restrict__ rsq = p.rsq;
    // Stage overlap (p.fill_n > 0): the dependents validate amplitudes by the all-ones sentinel, so G[0, fill_n) is filled with
    // it before they are released -- by this kernel itself (CTAs >= 1, and CTA 0 for the cells it is about to compute), never
    // by a griddepcontrol.wait: a kernel that waits also waits for everything its predecessor was chained to, which would
    // serialise consecutive lattices of a batch.  __threadfence makes the fill visible before the release.
    if (p.fill_n > 0) {
        ulonglong2 *gs = (ulonglong2 *)p.G;
        const ulonglong2 sent = make_ulonglong2(0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull);
        const long long own = (long long)Sc * S2 < p.fill_n ? (long long)Sc * S2 : p.fill_n;   // what CTA 0 computes itself
        if (blockIdx.x == 0) {
            for (long long n = lane; n < own; n += 32) gs[n] = sent;
        } else {
            const long long nf = gridDim.x - 1;
            for (long long n = own + (long long)(blockIdx.x - 1) * 32 + lane; n < p.fill_n; n += nf * 32) gs[n] = sent;
        }
        __threadfence();
    }
    pdl_launch_dependents();
    if (blockIdx.x != 0) return;
    if (lane == 0) { timeline_stamp(p.timeline, 8, 0); timeline_stamp(p.timeline, 8, 1); }
    // every cold global load of the kernel is issued here, before the first use of any of them (one DRAM round trip
    // instead of three dependent ones in front of the chain)
    const c128 Ac = p.A[ic * D + ic], bc = p.b[ic], c0 = p.c[0];
    const c128 Arow = p.A[i2 * D + ic], b0 = p.b[i2], a00 = p.A[i2 * D + i2];
    for (int n = lane; n < Sc; n += 32) sqt[n] = make_double2(sq[n], rsq[n]);
    extern __shared__ double2 sqt2[];
    const bool tab = p.L != 0;   // the host sets L = 1 when the table of the marched index fits in shared memory
    if (tab) for (int n = lane; n < S2; n += 32) sqt2[n] = make_double2(sq[n], rsq[n]);
    __syncwarp();
    if (lane == 0) {   // stage D-1: G[n] = (b G[n-1] + A sqrt(n-1) G[n-2]) / sqrt(n)   (core.py:97-104 with i = D-1)
        // software pipelined: the A-term of step s+1, (A sqrt(s)) G[s-1], and the table entry of step s+1 are issued
        // beside the quotient of step s, so the dependent chain of a step is b*G (2 levels), one add, the range test and
        // the 5-level quotient
        c128 p1 = c0, aterm = c_make(0.0, 0.0);
        chain[0] = p1;
        double2 t = sqt[Sc > 1 ? 1 : 0];
        for (int s = 1; s < Sc; s++) {
            const double2 tn = sqt[s + 1 < Sc ? s + 1 : s];
            c128 v = c_mul(bc, p1);
            if (s >= 2) v = c_add(v, aterm);
            aterm = c_mul(c_scale(Ac, t.x), p1);
            v = c_div_table_spec(v, t.x, t.y);
            chain[s] = v;
            p1 = v; t = tn;
        }
    }
    __syncwarp();
    // stage D-2: lane l owns k_{D-1} = l and l + 32
    bool act[2];
    c128 P1[2], aterm[2], coef[2];
    c128 *g[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int k = lane + 32 * r;
        act[r] = k < Sc;
        P1[r] = act[r] ? chain[k] : c_make(0.0, 0.0);
        aterm[r] = c_make(0.0, 0.0);
        coef[r] = (act[r] && k > 0) ? c_scale(Arow, sqt[k].x) : c_make(0.0, 0.0);
        g[r] = p.G + k;
        if (act[r]) stg_strong(g[r], P1[r]);   // stage i1 may validate these amplitudes by polling (stage overlap): strong stores
    }
    if (lane == 0) timeline_stamp(p.timeline, 8, 2);   // chain done
    if (S2 < 2) return;
    // (sqrt(s), 1/sqrt(s)) of the marched index in shared memory: a cold global load per step (an L2 round trip of
    // ~300 cycles every few steps) would sit on the dependent chain of a ~100-cycle step
#define MMH_TAIL_SQ(n) (tab ? sqt2[(n)] : make_double2(sq[(n)], rsq[(n)]))
    const long long P = d.strides[i2];   // = Sc
    const int rot_src = (lane + 31) & 31;
    double2 t = MMH_TAIL_SQ(1);
#pragma unroll 1
    for (int s = 1; s < S2; s++) {
        const double2 tn = MMH_TAIL_SQ(s + 1 < S2 ? s + 1 : s);
        // neighbours k_{D-1} - 1 of the previous panel: slot 0 from the lane below; slot 1 from the lane below, lane 0 from
        // lane 31's slot 0 (one rotate of a per-lane selected value: two complex shuffles per step)
        const c128 up0 = shfl_up_c128(P1[0], 1);
        const c128 rot = shfl_c128(lane == 31 ? P1[0] : P1[1], rot_src);
        c128 nb[2], v[2];
        nb[0] = lane == 0 ? c_make(0.0, 0.0) : up0;
        nb[1] = rot;
        const c128 a00s = c_scale(a00, t.x);   // A_ii sqrt(s): coefficient of this panel in step s+1
#pragma unroll
        for (int r = 0; r < 2; r++) {
            v[r] = c_mul(b0, P1[r]);
            if (s >= 2) v[r] = c_add(v[r], aterm[r]);
            v[r] = c_add(v[r], c_mul(coef[r], nb[r]));
            aterm[r] = c_mul(a00s, P1[r]);
        }
        div_all_inplace<2>(v, t.x, t.y);
#pragma unroll
        for (int r = 0; r < 2; r++) {
            P1[r] = v[r];
            g[r] += P;
            if (act[r]) stg_strong(g[r], v[r]);
        }
        t = tn;
    }
#undef MMH_TAIL_SQ
    if (lane == 0) timeline_stamp(p.timeline, 8, 3);
}


// all-ones sentinel over n amplitudes (stage overlap: panel 0 of the last tiled stage)
struct FillArgs { ulonglong2 *p; long long n; };
__global__ void __launch_bounds__(256) k_fill_sentinel_s(FillArgs a) {
    pdl_launch_dependents();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
        a.p[i] = make_ulonglong2(0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull);
}
cudaError_t mmh_launch_fill_sentinel(c128 *p, long long n, bool pdl, cudaStream_t st) {
    long long grid = (n + 255) / 256;
    if (grid > 592) grid = 592;
    FillArgs a = { (ulonglong2 *)p, n };
    return launch_pdl(k_fill_sentinel_s, (int)grid, 256, 0, st, pdl, a);
}

cudaError_t mmh_launch_warp_tail(const StageParams &p0, cudaStream_t st) {
    StageParams p = p0;
    const int S2 = p.d.shape[p.d.D - 2];
    size_t smem = sizeof(double2) * (size_t)S2;
    p.L = smem <= 160 * 1024 ? 1 : 0;
    if (!p.L) smem = 0;
    if (smem > 40 * 1024) cudaFuncSetAttribute(k_warp_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = 1;
    if (p.fill_n > (long long)p.d.shape[p.d.D - 1] * S2) {
        const long long cells = p.fill_n - (long long)p.d.shape[p.d.D - 1] * S2;
        grid += (int)((cells + 32 * 32 - 1) / (32 * 32) < 512 ? (cells + 32 * 32 - 1) / (32 * 32) : 512);
    }
    return launch_pdl(k_warp_tail, grid, 32, smem, st, p.pdl != 0, p);
}

template <int R>
static cudaError_t launch_stage_R(const StageParams &p, int grid, int block, size_t smem, cudaStream_t st) {
#define MMH_CASE(N)                                                                                   \
    case N:                                                                                           \
        if (smem > 48 * 1024)                                                                         \
            cudaFuncSetAttribute(k_march_stage<R, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        return launch_pdl(k_march_stage<R, N>, grid, block, smem, st, p.pdl != 0, p);
    const int npd = p.d.D - 1 - p.stage;
    switch (npd) {
        MMH_CASE(1) MMH_CASE(2)
        default: break;
    }
    if constexpr (R <= 2) {
        switch (npd) { MMH_CASE(3) MMH_CASE(4) default: break; }
    }
    if constexpr (R == 1) {
        switch (npd) { MMH_CASE(5) MMH_CASE(6) MMH_CASE(7) default: break; }
    }
#undef MMH_CASE
    return cudaErrorInvalidValue;
}

cudaError_t mmh_launch_march_stage(const StageParams &p, int R, int grid, int block, size_t smem, cudaStream_t st) {
    switch (R) {
        case 1: return launch_stage_R<1>(p, grid, block, smem, st);
        case 2: return launch_stage_R<2>(p, grid, block, smem, st);
        case 4: return launch_stage_R<4>(p, grid, block, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t mmh_launch_chain(const FwdParams &p, cudaStream_t st) {
    const int S_ = p.d.shape[p.d.D - 1];
    if (p.batch >= 256 && S_ >= 4 && S_ <= 8192 && !mmh_getenv("MMH_NO_CHAIN_ROWS")) {   // batches: staged chunks, coalesced stores
        const int T = 128;
        const size_t smem_ = sizeof(double2) * (size_t)S_ + sizeof(c128) * (size_t)2 * T * (MMH_CHAIN_CH + 1);
        if (smem_ > 48 * 1024) cudaFuncSetAttribute(k_fwd_chain_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_);
        k_fwd_chain_rows<<<(unsigned)((p.batch + T - 1) / T), T, smem_, st>>>(p);
        return cudaGetLastError();
    }
    const int block = 128;
    const long long grid = (p.batch + block - 1) / block;
    size_t smem = sizeof(double2) * (size_t)p.d.shape[p.d.D - 1];
    const int tab = smem <= 160 * 1024 ? 1 : 0;
    if (!tab) smem = 0;
    if (smem > 40 * 1024) cudaFuncSetAttribute(k_fwd_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_fwd_chain<<<(unsigned)grid, block, smem, st>>>(p, tab);
    return cudaGetLastError();
}

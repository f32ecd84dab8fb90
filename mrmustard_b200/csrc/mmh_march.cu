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
// K2  k_fwd_batched_march : one CTA marches L small lattices in lock step (batched path, cfg3).
// K1  k_fwd_tiled_march   : one lattice, every CTA owns a tile of the panel and exchanges one-cell halos
//                           with its lower neighbours through L2 + release/acquire flags (cfg2, cfg5).
#include "mmh_params.cuh"
#include "mmh_points.cuh"

// ---------------------------------------------------------------------------------------------------
// K2: batched stage march.  One launch per stage i (i = D-2 .. 0); stage D-1 is k_fwd_chain.
// The CTA marches L lattices in lock step; a thread owns R fixed panel positions ("slots").
// ---------------------------------------------------------------------------------------------------
// smem layout (c128 units): sba[L][2] = (b_i, A_ii) | buf[2][L*P]
template <int R, int NPD>
__global__ void __launch_bounds__(R >= 4 ? 256 : 512, R >= 4 ? 2 : 1) k_march_stage(StageParams p) {
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

    for (int t = tid; t < nlat; t += T) {
        sba[2 * t] = p.b[(lat0 + t) * D + i];
        sba[2 * t + 1] = p.A[(lat0 + t) * D * D + i * D + i];
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
        c128 v[R], qq[R];                                                                             \
        _Pragma("unroll") for (int r = 0; r < R; r++) {                                               \
            const c128 b0 = sba[isb[r]], a00 = sba[isb[r] + 1];                                       \
            v[r] = c_mul(b0, P1[r]);                                                                  \
            v[r] = c_add(v[r], c_mul(c_scale(a00, sqm), P2[r]));                                      \
            _Pragma("unroll") for (int jj = 0; jj < NPD; jj++)                                        \
                v[r] = c_add(v[r], c_mul(coef[r][jj], buf[(OFFP) + nbi[r][jj]]));                     \
        }                                                                                             \
        bool slow = false;                                                                            \
        _Pragma("unroll") for (int r = 0; r < R; r++) {                                               \
            qq[r] = c_make(div_fast(v[r].x, sqs, rsqs), div_fast(v[r].y, sqs, rsqs));                 \
            slow |= div_needs_slow(v[r].x) | div_needs_slow(v[r].y);                                  \
        }                                                                                             \
        if (slow) {                                                                                   \
            _Pragma("unroll") for (int r = 0; r < R; r++) {                                           \
                if (div_needs_slow(v[r].x)) qq[r].x = __ddiv_rn(v[r].x, sqs);                         \
                if (div_needs_slow(v[r].y)) qq[r].y = __ddiv_rn(v[r].y, sqs);                         \
            }                                                                                         \
        }                                                                                             \
        _Pragma("unroll") for (int r = 0; r < R; r++) {                                               \
            P2[r] = qq[r];                                                                            \
            if (act[r]) {                                                                             \
                gp[r] += P;                                                                           \
                *gp[r] = qq[r];                                                                       \
                buf[(OFFC) + loc[r]] = qq[r];                                                         \
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
__global__ void __launch_bounds__(128) k_fwd_chain(FwdParams p) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= p.batch) return;
    const int D = p.d.D, i = D - 1;
    const c128 A = p.A[l * D * D + i * D + i], b = p.b[l * D + i];
    c128 *g = p.G + l * p.d.N;
    const int S = p.d.shape[i];
    c128 p1 = p.c[l], p2 = c_make(0.0, 0.0);
    g[0] = p1;
    for (int s = 1; s < S; s++) {
        c128 v = c_mul(b, p1);
        if (s >= 2) v = c_add(v, c_mul(c_scale(A, p.sq[s - 1]), p2));
        v = c_div_table(v, p.sq[s], p.rsq[s]);
        g[s] = v;
        p2 = p1; p1 = v;
    }
}

template <int R>
static cudaError_t launch_stage_R(const StageParams &p, int grid, int block, size_t smem, cudaStream_t st) {
#define MMH_CASE(N)                                                                                   \
    case N:                                                                                           \
        if (smem > 48 * 1024)                                                                         \
            cudaFuncSetAttribute(k_march_stage<R, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        k_march_stage<R, N><<<grid, block, smem, st>>>(p);                                            \
        return cudaGetLastError();
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
    const int block = 128;
    const long long grid = (p.batch + block - 1) / block;
    k_fwd_chain<<<(unsigned)grid, block, 0, st>>>(p);
    return cudaGetLastError();
}

// mmh_diagonal_rolling.cu — compactFock "diagonal" sweep with ROLLING weight-level buffers (sm_100a): the form that runs
// BASELINE config 4 as written (8-mode Gaussian ket, cutoff 12: A 16x16, 430 M diagonal amplitudes).
//
// Reference: mrmustard/math/lattice/strategies/compactFock/diagonal_amps.py:19-248.  The reference keeps the four auxiliary
// families (arr1, arr2, arr1010, arr1001) at full size -- 136 arrays of prod(cutoffs) entries for M = 8, 0.94 TB at cutoff 12 --
// although a weight level w = sum(params) only ever reads levels w and w - 1 of them (the same
// observation fast_diagonal.py:65-76 exploits with three dict buffers).  Here:
//   * arr0 (the answer) is the only full-size array, addressed by the flat index of `params`;
//   * the auxiliary families live in two level buffers (current / previous), addressed by the RANK of `params` inside its level
//     in lexicographic order.  rank(params) = sum_j Pre[j][rem_j][params_j] with Pre[j][s][v] = #{tails of modes j.. whose digit j
//     is < v and which sum to s} -- tables built on the host from the bounded-composition counts C[j][s];
//   * the families of pivot d are only ever written / read where params[:d] == 0 (diagonal_amps.py:183), and those tuples are
//     exactly the first C[d][w] ranks of the level, so pivot d's arrays hold C[d][w] entries instead of C[0][w]:
//     M = 8, cutoff 12 peaks at 9 GB per level instead of 37 GB;
//   * one thread owns one `params` of the level (times the B-batch entry): it unranks itself, derives the ranks of its M lower
//     neighbours params - e_j from prefix / suffix sums of the same table entries (O(M) lookups for all of them), runs the
//     diagonal pivot (diagonal_amps.py:98-141) and then -- with arr1[2d, params] still in registers -- the off-diagonal pivots
//     d = 0 .. (first non-zero index) (diagonal_amps.py:19-94): ONE launch per level, every (params, level) pair visited once.
// Arithmetic (tolerance-gated, 1e-10 / 1e-14 like the full-layout kernels): value = (pivot B_i + sum_l A_il K_l G_in[l]) / K_i.
#include <vector>

#include "mmh_params.cuh"

#define MMH_ROLL_MAXM 8
#ifndef MMH_ROLL_MINB
#define MMH_ROLL_MINB 4   // resident CTAs per SM asked of ptxas (128 registers, 48 B spilled): the gathers are latency bound and want the
                          // occupancy -- 8-mode cutoff 12: 369 ms at 2 CTAs per SM, 271 at 3, 254 at 4
#endif

__device__ __forceinline__ c128 r_cmul(c128 x, c128 y) { return make_double2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x); }
__device__ __forceinline__ c128 r_cadd(c128 x, c128 y) { return make_double2(x.x + y.x, x.y + y.y); }
__device__ __forceinline__ c128 r_cscale(c128 x, double s) { return make_double2(x.x * s, x.y * s); }

struct DiagRollParams {
    int M, nb, level, W1, cm1;       // W1 = max level + 2 (row length of the tables), cm1 = max cutoff + 1
    int cut[MMH_ROLL_MAXM];
    long long pst[MMH_ROLL_MAXM];    // row-major strides of arr0 over `cut`
    const c128 *A, *B;               // [2M, 2M], [2M, nb]
    c128 *arr0;                      // [prod(cut), nb]
    c128 *cur;                       // level buffer of level w:   arr1[i] at i * n0 ; pivot d's arrays at baseD[d] + a * nD[d]
    const c128 *prev;                // level buffer of level w - 1 (same layout with its own sizes)
    long long n0_cur, n0_prev;       // C[0][w], C[0][w-1]
    long long baseD_cur[MMH_ROLL_MAXM], baseD_prev[MMH_ROLL_MAXM];   // in entries (before the nb factor)
    long long nD_cur[MMH_ROLL_MAXM], nD_prev[MMH_ROLL_MAXM];         // C[d][w], C[d][w-1]
    const long long *Pre;            // [M][W1][cm1]
    const double *sq;
};

__device__ __forceinline__ long long pre_at(const DiagRollParams &q, int j, int s, int v) {
    return __ldg(q.Pre + ((long long)j * q.W1 + s) * q.cm1 + v);
}

// value = (piv * B[i] + sum_l A[i, l] G_in[l]) / K.  The sum always runs over all 2M entries, fully unrolled: the entries below the
// pivot's first index are exact zeros, and a compile-time trip count keeps G_in in registers (a run-time lower bound put the array
// in local memory: 162 LDL per thread in the SASS)
template <int N2>
__device__ __forceinline__ c128 roll_value(const DiagRollParams &q, const c128 *sA, int i, c128 piv, const c128 (&G_in)[N2], int t, double K) {
    c128 v = r_cmul(piv, q.B[(long long)i * q.nb + t]);
#pragma unroll
    for (int l = 0; l < N2; l++) c_fma(v, sA[i * N2 + l], G_in[l]);   // fused multiply-adds: this path is tolerance-gated (1e-10), not bit-exact
    return make_double2(v.x / K, v.y / K);
}

template <int MT>
__global__ void __launch_bounds__(128, MMH_ROLL_MINB) k_diag_roll(DiagRollParams q) {
    extern __shared__ c128 sA[];
    constexpr int M = MT, n2 = 2 * MT;
    for (int e = threadIdx.x; e < n2 * n2; e += blockDim.x) sA[e] = q.A[e];
    __syncthreads();
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= q.n0_cur * q.nb) return;
    const int t = (int)(gid % q.nb);
    const long long r = gid / q.nb;
    const double *__restrict__ sq = q.sq;
    const int w = q.level;

    // ---- unrank: params of rank r inside level w (lexicographic) ----
    int params[M], remj[M];
    long long term[M];
    {
        long long rr = r;
        int rem = w;
#pragma unroll
        for (int j = 0; j < M; j++) {
            const int vmax = q.cut[j] - 1 < rem ? q.cut[j] - 1 : rem;
            int v = 0;
            long long lo = 0;
            while (v < vmax) {
                const long long nx = pre_at(q, j, rem, v + 1);
                if (nx > rr) break;
                lo = nx; v++;
            }
            params[j] = v; term[j] = lo; remj[j] = rem;
            rr -= lo; rem -= v;
        }
    }
    // ---- ranks of the lower neighbours params - e_j inside level w - 1 ----
    long long nbr[M];
    {
        long long suf[M];
        long long s = 0;
#pragma unroll
        for (int j = M - 1; j >= 0; j--) { suf[j] = s; s += term[j]; }
        long long pp = 0;
#pragma unroll
        for (int j = 0; j < M; j++) {
            const int rm = remj[j] > 0 ? remj[j] - 1 : 0;
            nbr[j] = params[j] > 0 ? pp + pre_at(q, j, rm, params[j] - 1) + suf[j] : 0;
            pp += pre_at(q, j, rm, params[j]);
        }
    }
    long long flat = 0;
#pragma unroll
    for (int j = 0; j < M; j++) flat += (long long)params[j] * q.pst[j];
    const long long nb = q.nb;
    const c128 a0 = q.arr0[flat * nb + t];

    // ---- diagonal pivot [a,a,b,b,...] (diagonal_amps.py:98-141) ----
    c128 a1_0 = make_double2(0.0, 0.0);   // arr1[0, params]: the pivot of the off-diagonal step d = 0 (the pivots d >= 1, used only where
                                          // params[:d] == 0, are re-read from the level buffer this thread has just written)
    c128 G_in[n2];
    if (q.cut[0] == 1 || params[0] < q.cut[0] - 1) {
#pragma unroll
        for (int l = 0; l < n2; l++) {
            const int j = l >> 1;
            G_in[l] = make_double2(0.0, 0.0);
            if (params[j] > 0) G_in[l] = r_cscale(q.prev[((long long)(l ^ 1) * q.n0_prev + nbr[j]) * nb + t], sq[params[j]]);
        }
#pragma unroll
        for (int i = 0; i < n2; i++) {
            const int j = i >> 1;
            if (params[j] + 1 < q.cut[j] && (i != 1 || params[0] + 2 < q.cut[0])) {
                const c128 v = roll_value<n2>(q, sA, i, a0, G_in, t, sq[params[j] + 1]);
                q.cur[((long long)i * q.n0_cur + r) * nb + t] = v;
                if (i == 0) a1_0 = v;
            }
        }
    }
    // ---- off-diagonal pivots [.., p_d + 1, p_d, ..] for d = 0 .. first non-zero index (diagonal_amps.py:19-94) ----
#pragma unroll
    for (int d = 0; d < M; d++) {
        if (params[d] < q.cut[d] - 1) {
#pragma unroll
            for (int l = 0; l < n2; l++) G_in[l] = make_double2(0.0, 0.0);
            G_in[2 * d] = r_cscale(a0, sq[params[d] + 1]);
            const c128 *pv = q.prev + q.baseD_prev[d] * nb;
            const long long nP = q.nD_prev[d];
            if (params[d] > 0) G_in[2 * d + 1] = r_cscale(pv[nbr[d] * nb + t], sq[params[d]]);                       // arr2[d]
#pragma unroll
            for (int i = d + 1; i < M; i++) {
                if (params[i] > 0) {
                    const int a = 1 + 2 * (i - d - 1);
                    G_in[2 * i] = r_cscale(pv[((long long)(a + 1) * nP + nbr[i]) * nb + t], sq[params[i]]);          // arr1001[d, i]
                    G_in[2 * i + 1] = r_cscale(pv[((long long)a * nP + nbr[i]) * nb + t], sq[params[i]]);            // arr1010[d, i]
                }
            }
            const c128 piv = d == 0 ? a1_0 : q.cur[((long long)(2 * d) * q.n0_cur + r) * nb + t];
            c128 *cv = q.cur + q.baseD_cur[d] * nb;
            const long long nC = q.nD_cur[d];
            q.arr0[(flat + q.pst[d]) * nb + t] = roll_value<n2>(q, sA, 2 * d + 1, piv, G_in, t, sq[params[d] + 1]);
            if (params[d] + 2 < q.cut[d]) cv[r * nb + t] = roll_value<n2>(q, sA, 2 * d, piv, G_in, t, sq[params[d] + 2]);   // arr2[d]
#pragma unroll
            for (int i = d + 1; i < M; i++) {
                if (params[i] + 1 < q.cut[i]) {
                    const int a = 1 + 2 * (i - d - 1);
                    cv[((long long)a * nC + r) * nb + t] = roll_value<n2>(q, sA, 2 * i, piv, G_in, t, sq[params[i] + 1]);            // arr1010
                    cv[((long long)(a + 1) * nC + r) * nb + t] = roll_value<n2>(q, sA, 2 * i + 1, piv, G_in, t, sq[params[i] + 1]);  // arr1001
                }
            }
        }
        if (params[d] != 0) break;   // pivot d needs params[:d] == 0
    }
}

__global__ void k_diag_roll_seed(c128 *arr0, const c128 *G0, int nb) {
    for (int t = threadIdx.x; t < nb; t += blockDim.x) arr0[t] = G0[0];
}

// bytes of ONE level buffer (the maximum over the levels) and the tables; host side
struct RollPlan {
    int M, W;                         // W = max level
    int cm1, W1;
    std::vector<long long> C;         // [(M + 1)][W1]
    std::vector<long long> Pre;       // [M][W1][cm1]
    long long max_entries;            // entries (before nb) of the largest level buffer
};

static void roll_plan(int M, const int *cut, RollPlan *pl) {
    int W = 0, cm = 1;
    for (int j = 0; j < M; j++) { W += cut[j] - 1; if (cut[j] > cm) cm = cut[j]; }
    pl->M = M; pl->W = W; pl->W1 = W + 2; pl->cm1 = cm + 1;
    const int W1 = pl->W1, cm1 = pl->cm1;
    pl->C.assign((size_t)(M + 1) * W1, 0);
    pl->Pre.assign((size_t)M * W1 * cm1, 0);
    pl->C[(size_t)M * W1 + 0] = 1;
    for (int j = M - 1; j >= 0; j--)
        for (int s = 0; s < W1; s++) {
            long long acc = 0;
            for (int v = 0; v <= cm; v++) {
                pl->Pre[((size_t)j * W1 + s) * cm1 + v] = acc;
                if (v < cut[j] && s - v >= 0) acc += pl->C[(size_t)(j + 1) * W1 + (s - v)];
            }
            pl->C[(size_t)j * W1 + s] = acc;
        }
    pl->max_entries = 0;
    for (int w = 0; w <= W; w++) {
        long long e = 2LL * M * pl->C[w];
        for (int d = 0; d < M; d++) e += (long long)(1 + 2 * (M - 1 - d)) * pl->C[(size_t)d * W1 + w];
        if (e > pl->max_entries) pl->max_entries = e;
    }
}

// workspace query: bytes of the two level buffers + the rank table
size_t mmh_diagonal_rolling_workspace(int M, const int *cut, int nb) {
    RollPlan pl;
    roll_plan(M, cut, &pl);
    const size_t tab = (sizeof(long long) * pl.Pre.size() + 255) / 256 * 256;
    return tab + 2 * sizeof(c128) * (size_t)pl.max_entries * (size_t)nb;
}

cudaError_t mmh_launch_diagonal_rolling(int M, const int *cut, int nb, const c128 *A, const c128 *B, const c128 *G0, c128 *arr0,
                                        const double *sq, void *workspace, long long *launches, cudaStream_t st) {
    RollPlan pl;
    roll_plan(M, cut, &pl);
    const size_t tab = (sizeof(long long) * pl.Pre.size() + 255) / 256 * 256;
    cudaError_t e = cudaMemcpyAsync(workspace, pl.Pre.data(), sizeof(long long) * pl.Pre.size(), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);   // pl is a local
    if (e != cudaSuccess) return e;
    c128 *buf0 = (c128 *)((char *)workspace + tab);
    c128 *buf1 = buf0 + (size_t)pl.max_entries * nb;
    DiagRollParams q;
    memset(&q, 0, sizeof(q));
    q.M = M; q.nb = nb; q.W1 = pl.W1; q.cm1 = pl.cm1;
    q.A = A; q.B = B; q.arr0 = arr0; q.Pre = (const long long *)workspace; q.sq = sq;
    for (int j = 0; j < M; j++) q.cut[j] = cut[j];
    q.pst[M - 1] = 1;
    for (int j = M - 1; j > 0; j--) q.pst[j - 1] = q.pst[j] * cut[j];
    k_diag_roll_seed<<<1, 128, 0, st>>>(arr0, G0, nb);
    *launches = 1;
    const size_t smem = sizeof(c128) * (size_t)(4 * M * M);
    for (int w = 0; w <= pl.W; w++) {
        q.level = w;
        q.cur = (w & 1) ? buf1 : buf0;
        q.prev = (w & 1) ? buf0 : buf1;
        for (int pass = 0; pass < 2; pass++) {   // layout of level w (cur) and of level w - 1 (prev)
            const int lv = w - pass;
            long long *base = pass ? q.baseD_prev : q.baseD_cur, *nD = pass ? q.nD_prev : q.nD_cur;
            const long long n0 = lv >= 0 ? pl.C[lv] : 0;
            (pass ? q.n0_prev : q.n0_cur) = n0;
            long long off = 2LL * M * n0;
            for (int d = 0; d < M; d++) {
                nD[d] = lv >= 0 ? pl.C[(size_t)d * pl.W1 + lv] : 0;
                base[d] = off;
                off += (long long)(1 + 2 * (M - 1 - d)) * nD[d];
            }
        }
        const long long threads = q.n0_cur * nb;
        if (threads <= 0) continue;
        const unsigned grid = (unsigned)((threads + 127) / 128);
        switch (M) {
            case 1: k_diag_roll<1><<<grid, 128, smem, st>>>(q); break;
            case 2: k_diag_roll<2><<<grid, 128, smem, st>>>(q); break;
            case 3: k_diag_roll<3><<<grid, 128, smem, st>>>(q); break;
            case 4: k_diag_roll<4><<<grid, 128, smem, st>>>(q); break;
            case 5: k_diag_roll<5><<<grid, 128, smem, st>>>(q); break;
            case 6: k_diag_roll<6><<<grid, 128, smem, st>>>(q); break;
            case 7: k_diag_roll<7><<<grid, 128, smem, st>>>(q); break;
            case 8: k_diag_roll<8><<<grid, 128, smem, st>>>(q); break;
            default: return cudaErrorInvalidValue;
        }
        (*launches)++;
    }
    return cudaGetLastError();
}

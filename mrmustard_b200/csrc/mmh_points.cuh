// mmh_points.cuh — per-amplitude update rules shared by the forward kernels.
// The arithmetic (term order, separate roundings, IEEE division) is the reference's, core.py:97-104 / :183-211.
#pragma once
#include "mmh_common.cuh"

// one amplitude of stage i, step s, panel offset f  (vanilla pivot rule)
__device__ __forceinline__ c128 vanilla_point(const LatticeDesc &d, const c128 *sA, const c128 *sb,
                                              const c128 *G, const double *__restrict__ sq,
                                              const double *__restrict__ rsq, int i, int s, long long f) {
    const int D = d.D;
    const long long si = d.strides[i];
    const long long pivot = (long long)(s - 1) * si + f;
    c128 val = c_mul(sb[i], G[pivot]);                                                  // core.py:97
    if (s >= 2) val = c_add(val, c_mul(c_scale(sA[i * D + i], sq[s - 1]), G[pivot - si]));  // :101
    long long rem = f;
    for (int j = i + 1; j < D; j++) {                                                   // :102-103
        const long long sj = d.strides[j];
        const int kj = (int)(rem / sj);
        rem -= (long long)kj * sj;
        if (kj > 0) val = c_add(val, c_mul(c_scale(sA[i * D + j], sq[kj]), G[pivot - sj]));
    }
    return c_div_table(val, sq[s], rsq[s]);                                             // :104
}

// 32-bit flavour of the same (N < 2^31): cheaper index arithmetic
__device__ __forceinline__ c128 vanilla_point32(const LatticeDesc &d, const c128 *sA, const c128 *sb,
                                                const c128 *G, const double *__restrict__ sq,
                                                const double *__restrict__ rsq, int i, int s, unsigned f) {
    const int D = d.D;
    const unsigned si = (unsigned)d.strides[i];
    const unsigned pivot = (unsigned)(s - 1) * si + f;
    c128 val = c_mul(sb[i], G[pivot]);
    if (s >= 2) val = c_add(val, c_mul(c_scale(sA[i * D + i], sq[s - 1]), G[pivot - si]));
    unsigned rem = f;
    for (int j = i + 1; j < D; j++) {
        const unsigned sj = (unsigned)d.strides[j];
        const unsigned kj = rem / sj;
        rem -= kj * sj;
        if (kj > 0) val = c_add(val, c_mul(c_scale(sA[i * D + j], sq[kj]), G[pivot - sj]));
    }
    return c_div_table(val, sq[s], rsq[s]);
}

// stable point (core.py:183-211): average of the update over every pivot i with k_i > 0.
// All neighbour amplitudes (k - e_i, and k - e_i - e_j once per unordered pair) are fetched FIRST, in one batch of
// independent loads, and the arithmetic then runs on registers in the reference's order: fetched one by one inside the sum
// they are D(D+1) dependent-looking L2 round trips per point (the level wavefront is L2-latency bound).
// x / np for the number of pivots np: powers of two are exact scalings (RN(x / 2^m) == RN(x * 2^-m), also in the subnormal
// range: both round the same exact quotient), everything else takes the IEEE division
__device__ __forceinline__ c128 c_div_count(c128 v, int np) {
    if (np == 1) return v;
    if (np == 2) return make_double2(__dmul_rn(v.x, 0.5), __dmul_rn(v.y, 0.5));
    if (np == 4) return make_double2(__dmul_rn(v.x, 0.25), __dmul_rn(v.y, 0.25));
    if (np == 8) return make_double2(__dmul_rn(v.x, 0.125), __dmul_rn(v.y, 0.125));
    return c_div_real(v, (double)np);
}

template <int DMAX>
static __device__ __forceinline__ c128 stable_point_batched(const LatticeDesc &d, const c128 *sA, const c128 *sb,
                                                            const c128 *G, const double *__restrict__ sq,
                                                            const double *__restrict__ rsq, const int *k, long long flat) {
    const int D = d.D;
    c128 p1[DMAX], p2[DMAX][DMAX];   // p1[i] = G[k - e_i], p2[i][j] (i <= j) = G[k - e_i - e_j]
#pragma unroll
    for (int i = 0; i < DMAX; i++) {
        if (i < D && k[i] > 0) p1[i] = __ldcg(G + (flat - d.strides[i]));
        else p1[i] = c_make(0.0, 0.0);
#pragma unroll
        for (int j = i; j < DMAX; j++) {
            const bool need = i < D && j < D && k[i] > 0 && (i == j ? k[i] > 1 : k[j] > 0);
            if (need) p2[i][j] = __ldcg(G + (flat - d.strides[i] - d.strides[j]));
            else p2[i][j] = c_make(0.0, 0.0);
        }
    }
    c128 vals = c_make(0.0, 0.0);
    int np = 0;
#pragma unroll
    for (int i = 0; i < DMAX; i++) {
        if (i >= D || k[i] == 0) continue;
        np++;
        c128 val = c_mul(sb[i], p1[i]);
#pragma unroll
        for (int j = 0; j < DMAX; j++) {
            if (j >= D) continue;
            if (j == i) { if (k[i] > 1) val = c_add(val, c_mul(c_scale(sA[i * D + i], sq[k[i] - 1]), p2[i][i])); }
            else if (k[j] > 0) val = c_add(val, c_mul(c_scale(sA[i * D + j], sq[k[j]]), j < i ? p2[j][i] : p2[i][j]));
        }
        vals = c_add(vals, c_div_table(val, sq[k[i]], rsq[k[i]]));   // == the IEEE quotient bit for bit (mmh_common.cuh), 5 ops instead of ~25
    }
    return c_div_count(vals, np);
}

static __device__ c128 stable_point_generic(const LatticeDesc &d, const c128 *sA, const c128 *sb,
                             const c128 *G, const double *__restrict__ sq, const double *__restrict__ rsq, const int *k,
                             long long flat) {
    const int D = d.D;
    c128 vals = c_make(0.0, 0.0);
    int np = 0;
    for (int i = 0; i < D; i++) {
        if (k[i] == 0) continue;
        np++;
        const long long pivot = flat - d.strides[i];
        c128 val = c_mul(sb[i], G[pivot]);
        for (int j = 0; j < i; j++)
            if (k[j] > 0) val = c_add(val, c_mul(c_scale(sA[i * D + j], sq[k[j]]), G[pivot - d.strides[j]]));
        if (k[i] > 1) val = c_add(val, c_mul(c_scale(sA[i * D + i], sq[k[i] - 1]), G[pivot - d.strides[i]]));
        for (int j = i + 1; j < D; j++)
            if (k[j] > 0) val = c_add(val, c_mul(c_scale(sA[i * D + j], sq[k[j]]), G[pivot - d.strides[j]]));
        vals = c_add(vals, c_div_table(val, sq[k[i]], rsq[k[i]]));
    }
    return c_div_count(vals, np);
}

static __device__ c128 stable_point(const LatticeDesc &d, const c128 *sA, const c128 *sb,
                             const c128 *G, const double *__restrict__ sq, const double *__restrict__ rsq, const int *k,
                             long long flat) {
    if (d.D <= 2) return stable_point_batched<2>(d, sA, sb, G, sq, rsq, k, flat);
    if (d.D <= 4) return stable_point_batched<4>(d, sA, sb, G, sq, rsq, k, flat);
    return stable_point_generic(d, sA, sb, G, sq, rsq, k, flat);
}


// mmh_diagonal.cu — compactFock "diagonal" (all modes PNR-detected) and "one leftover mode" recurrences, sm_100a.
//
// Reference: mrmustard/math/lattice/strategies/compactFock/diagonal_amps.py:19-197 and
// singleLeftoverMode_amps.py:21-421 (algorithms 1 and 2 of doi:10.22331/q-2023-08-29-1097).
// Only five families of near-diagonal amplitudes are kept (arr0 = the answer, arr1, arr2, arr1010, arr1001;
// SURVEY.md Appendix A.2).  For a weight level w = sum(params) every `params` of the level is independent:
//   * the diagonal pivot of `params` reads arr0[params] and arr1 at level w-1, writes arr1[., params];
//   * the off-diagonal pivots d of `params` read arr1[2d, params] (same level, written by the diagonal pivot),
//     arr0[params] and arr2/arr1001/arr1010 at level w-1, and write arr0[params + e_d] (level w+1) and
//     arr2/arr1010/arr1001[., params].
// So the lattice of `params` is swept level by level, two launches per level, one thread per
// (leftover block entry (m,n), params, B-batch entry).  The diagonal case is the leftover case with a 1x1 block.
#include "mmh_params.cuh"

#define MMH_DIAG_MAXMD 8

__device__ __forceinline__ c128 cmulf(c128 x, c128 y) {   // tolerance-gated path: contraction allowed
    return make_double2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
}
__device__ __forceinline__ c128 caddf(c128 x, c128 y) { return make_double2(x.x + y.x, x.y + y.y); }
__device__ __forceinline__ c128 cscalef(c128 x, double s) { return make_double2(x.x * s, x.y * s); }

struct DiagIdx {
    int params[MMH_DIAG_MAXMD];
    int sum, m, n, t;
    long long idx;   // element index inside one sub-array: ((m*c0 + n) * P + p) * nb + t
};

__device__ __forceinline__ bool diag_decode(const DiagParams &q, long long gid, DiagIdx &x) {
    const long long E = q.E;
    if (gid >= E) return false;
    x.idx = gid;
    x.t = (int)(gid % q.nb);
    long long r = gid / q.nb;
    long long p = r % q.P;
    const int mn = (int)(r / q.P);
    x.m = mn / q.c0;
    x.n = mn - x.m * q.c0;
    int sum = 0;
    for (int j = 0; j < q.Md; j++) {
        x.params[j] = (int)(p / q.pst[j]);
        p -= (long long)x.params[j] * q.pst[j];
        sum += x.params[j];
    }
    x.sum = sum;
    return sum == q.level;
}

// value = (piv * B[ii] + A[ii,0] lm + A[ii,1] ln + sum_l A[ii, 2 L0 + l] G_in[l]) / K
__device__ __forceinline__ c128 diag_write_value(const DiagParams &q, int ii, c128 piv, c128 lm, c128 ln, const c128 *G_in,
                                                 int l_lo, int t, double K) {
    const int n2 = 2 * (q.Md + q.L0);
    const c128 *Arow = q.A + (long long)ii * n2;
    c128 v = cmulf(piv, q.B[(long long)ii * q.nb + t]);
    if (q.L0) {
        v = caddf(v, cmulf(Arow[0], lm));
        v = caddf(v, cmulf(Arow[1], ln));
    }
    for (int l = l_lo; l < 2 * q.Md; l++) v = caddf(v, cmulf(Arow[2 * q.L0 + l], G_in[l]));
    return make_double2(v.x / K, v.y / K);
}

// diagonal pivot [a,a,b,b,...] (diagonal_amps.py:98-141 ; singleLeftoverMode_amps.py:225-287)
__global__ void __launch_bounds__(256) k_diag_pivot(DiagParams q) {
    DiagIdx x;
    if (!diag_decode(q, (long long)blockIdx.x * blockDim.x + threadIdx.x, x)) return;
    if (!(q.cut[0] == 1 || x.params[0] < q.cut[0] - 1)) return;
    const double *__restrict__ sq = q.sq;
    const int Md = q.Md;
    const long long E = q.E;
    c128 G_in[2 * MMH_DIAG_MAXMD];
    for (int l = 0; l < 2 * Md; l++) {
        const int j = l >> 1;
        G_in[l] = make_double2(0.0, 0.0);
        if (x.params[j] > 0)
            G_in[l] = cscalef(q.arr1[(long long)(l ^ 1) * E + x.idx - q.pst[j] * q.nb], sq[x.params[j]]);
    }
    const c128 piv = q.arr0[x.idx];
    c128 lm = make_double2(0.0, 0.0), ln = lm;
    if (q.L0) {
        if (x.m > 0) lm = cscalef(q.arr0[x.idx - (long long)q.c0 * q.P * q.nb], sq[x.m]);
        if (x.n > 0) ln = cscalef(q.arr0[x.idx - q.P * q.nb], sq[x.n]);
    }
    for (int i = 0; i < 2 * Md; i++) {
        const int j = i >> 1;
        if (x.params[j] + 1 < q.cut[j] && (i != 1 || x.params[0] + 2 < q.cut[0]))
            q.arr1[(long long)i * E + x.idx] =
                diag_write_value(q, i + 2 * q.L0, piv, lm, ln, G_in, 0, x.t, sq[x.params[j] + 1]);
    }
}

// off-diagonal pivots [.., (p_d + 1), p_d, ..] (diagonal_amps.py:19-94 ; singleLeftoverMode_amps.py:101-222)
__global__ void __launch_bounds__(256) k_diag_offdiag(DiagParams q) {
    DiagIdx x;
    if (!diag_decode(q, (long long)blockIdx.x * blockDim.x + threadIdx.x, x)) return;
    const double *__restrict__ sq = q.sq;
    const int Md = q.Md;
    const long long E = q.E;
    c128 G_in[2 * MMH_DIAG_MAXMD];
    for (int d = 0; d < Md; d++) {
        if (x.params[d] < q.cut[d] - 1) {
            for (int l = 0; l < 2 * Md; l++) G_in[l] = make_double2(0.0, 0.0);
            G_in[2 * d] = cscalef(q.arr0[x.idx], sq[x.params[d] + 1]);
            if (x.params[d] > 0) G_in[2 * d + 1] = cscalef(q.arr2[(long long)d * E + x.idx - q.pst[d] * q.nb], sq[x.params[d]]);
            for (int i = d + 1; i < Md; i++) {
                if (x.params[i] > 0) {
                    const long long o = (long long)(d * (Md - 1) + i - d - 1) * E + x.idx - q.pst[i] * q.nb;
                    G_in[2 * i] = cscalef(q.arr1001[o], sq[x.params[i]]);
                    G_in[2 * i + 1] = cscalef(q.arr1010[o], sq[x.params[i]]);
                }
            }
            const c128 *pa = q.arr1 + (long long)(2 * d) * E;
            const c128 piv = pa[x.idx];
            c128 lm = make_double2(0.0, 0.0), ln = lm;
            if (q.L0) {
                if (x.m > 0) lm = cscalef(pa[x.idx - (long long)q.c0 * q.P * q.nb], sq[x.m]);
                if (x.n > 0) ln = cscalef(pa[x.idx - q.P * q.nb], sq[x.n]);
            }
            const int o2 = 2 * q.L0;
            q.arr0[x.idx + q.pst[d] * q.nb] =
                diag_write_value(q, 2 * d + 1 + o2, piv, lm, ln, G_in, 2 * d, x.t, sq[x.params[d] + 1]);
            if (x.params[d] + 2 < q.cut[d])
                q.arr2[(long long)d * E + x.idx] =
                    diag_write_value(q, 2 * d + o2, piv, lm, ln, G_in, 2 * d, x.t, sq[x.params[d] + 2]);
            for (int i = d + 1; i < Md; i++) {
                if (x.params[i] + 1 < q.cut[i]) {
                    const long long o = (long long)(d * (Md - 1) + i - d - 1) * E + x.idx;
                    q.arr1010[o] = diag_write_value(q, 2 * i + o2, piv, lm, ln, G_in, 2 * d, x.t, sq[x.params[i] + 1]);
                    q.arr1001[o] = diag_write_value(q, 2 * i + 1 + o2, piv, lm, ln, G_in, 2 * d, x.t, sq[x.params[i] + 1]);
                }
            }
        }
        if (x.params[d] != 0) break;   // pivot d needs params[:d] == 0
    }
}

// seed: arr0[0...0] = G0 (per batch entry); for the leftover case the whole (c0 x c0) block at params = 0
// (singleLeftoverMode_amps.py:324-334): first column by a serial chain, then column n+1 from columns n, n-1.
__global__ void __launch_bounds__(256) k_diag_seed(DiagParams q, const c128 *G0) {
    const long long rowst = (long long)q.c0 * q.P * q.nb;   // (m, n) -> (m+1, n)
    const long long colst = q.P * q.nb;                     // (m, n) -> (m, n+1)
    if (!q.L0) {
        for (int t = threadIdx.x; t < q.nb; t += blockDim.x) q.arr0[t] = G0[0];
        return;
    }
    const c128 B0 = q.B[0], B1 = q.B[1];
    const int n2 = 2 * (q.Md + 1);
    const c128 A00 = q.A[0], A10 = q.A[n2], A11 = q.A[n2 + 1];
    const double *sq = q.sq;
    c128 *a = q.arr0;
    if (threadIdx.x == 0) {
        a[0] = G0[0];
        for (int m = 0; m + 1 < q.c0; m++) {
            c128 v = cmulf(a[m * rowst], B0);
            if (m > 0) v = caddf(v, cmulf(cscalef(A00, sq[m]), a[(m - 1) * rowst]));
            a[(m + 1) * rowst] = make_double2(v.x / sq[m + 1], v.y / sq[m + 1]);
        }
    }
    __syncthreads();
    for (int n = 0; n + 1 < q.c0; n++) {
        for (int m = threadIdx.x; m < q.c0; m += blockDim.x) {
            c128 v = cmulf(a[m * rowst + n * colst], B1);
            if (m > 0) v = caddf(v, cmulf(cscalef(A10, sq[m]), a[(m - 1) * rowst + n * colst]));
            if (n > 0) v = caddf(v, cmulf(cscalef(A11, sq[n]), a[m * rowst + (n - 1) * colst]));
            a[m * rowst + (n + 1) * colst] = make_double2(v.x / sq[n + 1], v.y / sq[n + 1]);
        }
        __syncthreads();
    }
}

cudaError_t mmh_launch_diagonal(DiagParams q, const c128 *G0, int nlevels, long long *launches, cudaStream_t st) {
    k_diag_seed<<<1, 256, 0, st>>>(q, G0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const long long grid = (q.E + 255) / 256;
    *launches = 1;
    for (int w = 0; w < nlevels; w++) {
        q.level = w;
        k_diag_pivot<<<(unsigned)grid, 256, 0, st>>>(q);
        k_diag_offdiag<<<(unsigned)grid, 256, 0, st>>>(q);
        *launches += 2;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Forward-mode Jacobians of the diagonal sweep (diagonal_grad.py:19-354): every stored amplitude X carries the
// tangents dX/dA[i,l] (4 M^2 of them, A entries treated as independent) and dX/dB[i] (2 M).  The tangent index
// theta is the innermost thread dimension, so the tangent sweep is the value sweep with a batch of 4M^2 + 2M
// "directions" and two injection terms (calc_dA_dB, diagonal_grad.py:19-37):
//     dX = ( dpiv * B[i] + [theta == B_i] piv + sum_l A[i,l] K_l dG_in[l] + [theta == A_il] K_l G_in[l] ) / K_i
// The value arrays (nb == 1) must have been filled by mmh_launch_diagonal first.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool tan_decode(const DiagTanParams &tp, long long gid, DiagIdx &x, int &theta) {
    const DiagParams &q = tp.q;
    const long long Et = q.E * tp.ntheta;
    if (gid >= Et) return false;
    theta = (int)(gid % tp.ntheta);
    const long long vi = gid / tp.ntheta;     // value-array index ((m*c0 + n) * P + p); nb == 1
    x.idx = vi;
    x.t = 0;
    long long p = vi % q.P;
    const int mn = (int)(vi / q.P);
    x.m = mn / q.c0;
    x.n = mn - x.m * q.c0;
    int sum = 0;
    for (int j = 0; j < q.Md; j++) {
        x.params[j] = (int)(p / q.pst[j]);
        p -= (long long)x.params[j] * q.pst[j];
        sum += x.params[j];
    }
    x.sum = sum;
    return sum == q.level;
}

// tangent of one written amplitude (calc_dA_dB: diagonal_grad.py:19-37, singleLeftoverMode_grad.py:19-47).
// ii = row of the full A; (lm, ln) / (dlm, dln): sqrt(m) piv[m-1,n], sqrt(n) piv[m,n-1] of the pivot array and their
// tangents (leftover case only).
__device__ __forceinline__ c128 tan_write_value(const DiagTanParams &tp, int ii, c128 piv, c128 dpiv, c128 lm, c128 ln,
                                                c128 dlm, c128 dln, const c128 *G_in, const c128 *dG_in, int l_lo, int theta,
                                                double K) {
    const DiagParams &q = tp.q;
    const int n2 = 2 * (q.Md + q.L0), o2 = 2 * q.L0;
    const c128 *Arow = q.A + (long long)ii * n2;
    c128 v = cmulf(dpiv, q.B[ii]);
    if (theta == n2 * n2 + ii) v = caddf(v, piv);                       // dB[i] += pivot_val
    if (q.L0) {
        v = caddf(v, cmulf(Arow[0], dlm));
        v = caddf(v, cmulf(Arow[1], dln));
    }
    for (int l = l_lo; l < 2 * q.Md; l++) v = caddf(v, cmulf(Arow[o2 + l], dG_in[l]));
    if (theta < n2 * n2 && theta / n2 == ii) {                          // dA[i, l] += (scaled) G_in[l]
        const int col = theta % n2;
        if (col >= o2) v = caddf(v, G_in[col - o2]);
        else if (q.L0) v = caddf(v, col == 0 ? lm : ln);
    }
    return make_double2(v.x / K, v.y / K);
}

__global__ void __launch_bounds__(256) k_diag_pivot_tan(DiagTanParams tp) {
    const DiagParams &q = tp.q;
    DiagIdx x;
    int theta;
    if (!tan_decode(tp, (long long)blockIdx.x * blockDim.x + threadIdx.x, x, theta)) return;
    if (!(q.cut[0] == 1 || x.params[0] < q.cut[0] - 1)) return;
    const double *__restrict__ sq = q.sq;
    const int Md = q.Md, nt = tp.ntheta;
    const long long E = q.E;
    c128 G_in[2 * MMH_DIAG_MAXMD], dG_in[2 * MMH_DIAG_MAXMD];
    for (int l = 0; l < 2 * Md; l++) {
        const int j = l >> 1;
        G_in[l] = dG_in[l] = make_double2(0.0, 0.0);
        if (x.params[j] > 0) {
            const long long o = (long long)(l ^ 1) * E + x.idx - q.pst[j];
            G_in[l] = cscalef(q.arr1[o], sq[x.params[j]]);
            dG_in[l] = cscalef(tp.t1[o * nt + theta], sq[x.params[j]]);
        }
    }
    const c128 piv = q.arr0[x.idx], dpiv = tp.t0[x.idx * nt + theta];
    c128 lm = make_double2(0.0, 0.0), ln = lm, dlm = lm, dln = lm;
    if (q.L0) {
        const long long om = x.idx - (long long)q.c0 * q.P, on = x.idx - q.P;
        if (x.m > 0) { lm = cscalef(q.arr0[om], sq[x.m]); dlm = cscalef(tp.t0[om * nt + theta], sq[x.m]); }
        if (x.n > 0) { ln = cscalef(q.arr0[on], sq[x.n]); dln = cscalef(tp.t0[on * nt + theta], sq[x.n]); }
    }
    for (int i = 0; i < 2 * Md; i++) {
        const int j = i >> 1;
        if (x.params[j] + 1 < q.cut[j] && (i != 1 || x.params[0] + 2 < q.cut[0]))
            tp.t1[((long long)i * E + x.idx) * nt + theta] =
                tan_write_value(tp, i + 2 * q.L0, piv, dpiv, lm, ln, dlm, dln, G_in, dG_in, 0, theta, sq[x.params[j] + 1]);
    }
}

__global__ void __launch_bounds__(256) k_diag_offdiag_tan(DiagTanParams tp) {
    const DiagParams &q = tp.q;
    DiagIdx x;
    int theta;
    if (!tan_decode(tp, (long long)blockIdx.x * blockDim.x + threadIdx.x, x, theta)) return;
    const double *__restrict__ sq = q.sq;
    const int Md = q.Md, nt = tp.ntheta, o2 = 2 * q.L0;
    const long long E = q.E;
    c128 G_in[2 * MMH_DIAG_MAXMD], dG_in[2 * MMH_DIAG_MAXMD];
    for (int d = 0; d < Md; d++) {
        if (x.params[d] < q.cut[d] - 1) {
            for (int l = 0; l < 2 * Md; l++) G_in[l] = dG_in[l] = make_double2(0.0, 0.0);
            G_in[2 * d] = cscalef(q.arr0[x.idx], sq[x.params[d] + 1]);
            dG_in[2 * d] = cscalef(tp.t0[x.idx * nt + theta], sq[x.params[d] + 1]);
            if (x.params[d] > 0) {
                const long long o = (long long)d * E + x.idx - q.pst[d];
                G_in[2 * d + 1] = cscalef(q.arr2[o], sq[x.params[d]]);
                dG_in[2 * d + 1] = cscalef(tp.t2[o * nt + theta], sq[x.params[d]]);
            }
            for (int i = d + 1; i < Md; i++) {
                if (x.params[i] > 0) {
                    const long long o = (long long)(d * (Md - 1) + i - d - 1) * E + x.idx - q.pst[i];
                    G_in[2 * i] = cscalef(q.arr1001[o], sq[x.params[i]]);
                    G_in[2 * i + 1] = cscalef(q.arr1010[o], sq[x.params[i]]);
                    dG_in[2 * i] = cscalef(tp.t1001[o * nt + theta], sq[x.params[i]]);
                    dG_in[2 * i + 1] = cscalef(tp.t1010[o * nt + theta], sq[x.params[i]]);
                }
            }
            const long long op = (long long)(2 * d) * E + x.idx;
            const c128 piv = q.arr1[op], dpiv = tp.t1[op * nt + theta];
            c128 lm = make_double2(0.0, 0.0), ln = lm, dlm = lm, dln = lm;
            if (q.L0) {
                const long long om = op - (long long)q.c0 * q.P, on = op - q.P;
                if (x.m > 0) { lm = cscalef(q.arr1[om], sq[x.m]); dlm = cscalef(tp.t1[om * nt + theta], sq[x.m]); }
                if (x.n > 0) { ln = cscalef(q.arr1[on], sq[x.n]); dln = cscalef(tp.t1[on * nt + theta], sq[x.n]); }
            }
            tp.t0[(x.idx + q.pst[d]) * nt + theta] =
                tan_write_value(tp, 2 * d + 1 + o2, piv, dpiv, lm, ln, dlm, dln, G_in, dG_in, 2 * d, theta, sq[x.params[d] + 1]);
            if (x.params[d] + 2 < q.cut[d])
                tp.t2[((long long)d * E + x.idx) * nt + theta] =
                    tan_write_value(tp, 2 * d + o2, piv, dpiv, lm, ln, dlm, dln, G_in, dG_in, 2 * d, theta, sq[x.params[d] + 2]);
            for (int i = d + 1; i < Md; i++) {
                if (x.params[i] + 1 < q.cut[i]) {
                    const long long o = ((long long)(d * (Md - 1) + i - d - 1) * E + x.idx) * nt + theta;
                    tp.t1010[o] = tan_write_value(tp, 2 * i + o2, piv, dpiv, lm, ln, dlm, dln, G_in, dG_in, 2 * d, theta, sq[x.params[i] + 1]);
                    tp.t1001[o] = tan_write_value(tp, 2 * i + 1 + o2, piv, dpiv, lm, ln, dlm, dln, G_in, dG_in, 2 * d, theta, sq[x.params[i] + 1]);
                }
            }
        }
        if (x.params[d] != 0) break;
    }
}

// tangents of the leftover seed block (singleLeftoverMode_grad.py:611-650): one thread per (theta, m); the first column
// is a serial chain over m, then column n+1 follows from columns n, n-1 (parallel over m).
__global__ void __launch_bounds__(1024) k_diag_seed_tan(DiagTanParams tp) {
    const DiagParams &q = tp.q;
    const int nt = tp.ntheta, c0 = q.c0;
    const int n2 = 2 * (q.Md + 1);
    const long long rowst = (long long)c0 * q.P, colst = q.P;
    const c128 B0 = q.B[0], B1 = q.B[1], A00 = q.A[0], A10 = q.A[n2], A11 = q.A[n2 + 1];
    const double *sq = q.sq;
    const c128 *a = q.arr0;
    c128 *t = tp.t0;
    const int thB0 = n2 * n2, thB1 = n2 * n2 + 1, thA00 = 0, thA10 = n2, thA11 = n2 + 1;
    for (int theta = threadIdx.x; theta < nt; theta += blockDim.x) {   // first column, serial in m
        for (int m = 0; m + 1 < c0; m++) {
            c128 v = cmulf(t[(m * rowst) * nt + theta], B0);
            if (theta == thB0) v = caddf(v, a[m * rowst]);
            if (m > 0) {
                v = caddf(v, cmulf(cscalef(A00, sq[m]), t[((m - 1) * rowst) * nt + theta]));
                if (theta == thA00) v = caddf(v, cscalef(a[(m - 1) * rowst], sq[m]));
            }
            t[((m + 1) * rowst) * nt + theta] = make_double2(v.x / sq[m + 1], v.y / sq[m + 1]);
        }
    }
    __syncthreads();
    for (int n = 0; n + 1 < c0; n++) {
        for (int w = threadIdx.x; w < nt * c0; w += blockDim.x) {
            const int theta = w % nt, m = w / nt;
            const long long o = m * rowst + n * colst;
            c128 v = cmulf(t[o * nt + theta], B1);
            if (theta == thB1) v = caddf(v, a[o]);
            if (m > 0) {
                v = caddf(v, cmulf(cscalef(A10, sq[m]), t[(o - rowst) * nt + theta]));
                if (theta == thA10) v = caddf(v, cscalef(a[o - rowst], sq[m]));
            }
            if (n > 0) {
                v = caddf(v, cmulf(cscalef(A11, sq[n]), t[(o - colst) * nt + theta]));
                if (theta == thA11) v = caddf(v, cscalef(a[o - colst], sq[n]));
            }
            t[(o + colst) * nt + theta] = make_double2(v.x / sq[n + 1], v.y / sq[n + 1]);
        }
        __syncthreads();
    }
}

cudaError_t mmh_launch_diagonal_tangent(DiagTanParams tp, int nlevels, long long *launches, cudaStream_t st) {
    const long long grid = (tp.q.E * tp.ntheta + 255) / 256;
    *launches = 0;
    if (tp.q.L0) {
        k_diag_seed_tan<<<1, 1024, 0, st>>>(tp);
        *launches += 1;
    }
    for (int w = 0; w < nlevels; w++) {
        tp.q.level = w;
        k_diag_pivot_tan<<<(unsigned)grid, 256, 0, st>>>(tp);
        k_diag_offdiag_tan<<<(unsigned)grid, 256, 0, st>>>(tp);
        *launches += 2;
    }
    return cudaGetLastError();
}

// mmh_gates.cu — gate-specific Fock strategies, sm_100a:
//   displacement / jacobian_displacement / grad_displacement   (mrmustard/math/lattice/strategies/displacement.py:24-139)
//   squeezer / squeezed                                         (strategies/squeezer.py:29-66, :127-147)
//   beamsplitter / stable_beamsplitter                          (strategies/beamsplitter.py:37-172)
//   the support masks of squeezer_vjp / squeezed_vjp / beamsplitter_vjp (squeezer.py:69-124,150-191; beamsplitter.py:175-243)
//
// These are what Dgate / Sgate / BSgate / SqueezedVacuum.fock_array call instead of the generic lattice.  Each is a short
// recurrence whose independent points form LEVELS: the squeezer's level m + n = u reads level u - 2 only, the beamsplitter's
// level m + n = L reads level L - 1 only (and, for the vanilla rule, never leaves its p-slice), the displacement's diagonals
// n - m = d are independent three-term chains.  The squeezer and the beamsplitter recurrences are numerically unstable at large
// cutoffs (a 1-ulp change of tanh r moves S[59,59] by 1e-7 relative), so parity with the reference needs the reference's own
// arithmetic: every product / sum / quotient below is an explicitly rounded IEEE operation in the reference's evaluation order
// (numba lowers float * complex as a full complex product with (s, 0) and complex / float as a component-wise division), and
// the transcendental scalars (cos, sin, tanh, cosh) are computed ONCE on the host with libm, which is what numba's lowering
// calls too (mmh_api.cu).  The results are then bit-identical to the numba strategies (tests/test_gpu_gates.py).
// The displacement evaluates log / exp per element like the reference; that one is tolerance-gated (1e-10 / 1e-14).
#include "mmh_params.cuh"

__device__ __forceinline__ c128 r_times_c(double s, c128 z) { return make_double2(__dmul_rn(s, z.x), __dmul_rn(s, z.y)); }
__device__ __forceinline__ c128 c_over_r(c128 z, double s) { return make_double2(__ddiv_rn(z.x, s), __ddiv_rn(z.y, s)); }
__device__ __forceinline__ c128 c_sub(c128 x, c128 y) { return make_double2(__dsub_rn(x.x, y.x), __dsub_rn(x.y, y.y)); }

// ---- squeezer: one CTA, levels u = m + n = 0, 2, 4, ... --------------------------------------------------------------
// S[m, 0] = (-sqrt(m-1) / sqrt(m)) e^{i theta} tanh r S[m-2, 0]
// S[m, n] = (sqrt(n-1) / sqrt(n)) conj(et) S[m, n-2] + ((sqrt(m) / sqrt(n)) sech r) S[m-1, n-1]        (m + n even)
__global__ void __launch_bounds__(1024) k_squeezer(GateParams p) {
    const int M = p.shape[0], N = p.shape[1];
    const double *__restrict__ sq = p.sq;
    c128 *S = p.out;
    const c128 et = p.z0, etc = make_double2(p.z0.x, -p.z0.y);
    const double sech = p.r0;
    for (long long f = threadIdx.x; f < (long long)M * N; f += blockDim.x) S[f] = c_make(0.0, 0.0);
    __syncthreads();
    if (threadIdx.x == 0) S[0] = c_make(p.r1, 0.0);   // sqrt(sech r)
    __syncthreads();
    for (int u = 2; u <= M + N - 2; u += 2) {
        const int m_lo = u - (N - 1) > 0 ? u - (N - 1) : 0, m_hi = u < M - 1 ? u : M - 1;
        for (int m = m_lo + (int)threadIdx.x; m <= m_hi; m += blockDim.x) {
            const int n = u - m;
            c128 v;
            if (n == 0) {
                v = c_mul(r_times_c(__ddiv_rn(-sq[m - 1], sq[m]), et), S[(long long)(m - 2) * N]);
            } else {
                v = c_make(0.0, 0.0);
                if (n >= 2) v = c_mul(r_times_c(__ddiv_rn(sq[n - 1], sq[n]), etc), S[(long long)m * N + n - 2]);
                if (m >= 1) v = c_add(v, r_times_c(__dmul_rn(__ddiv_rn(sq[m], sq[n]), sech), S[(long long)(m - 1) * N + n - 1]));
            }
            S[(long long)m * N + n] = v;
        }
        __syncthreads();
    }
}

// squeezed vacuum ket: S[m] = (sqrt(m-1) / sqrt(m)) (e^{i theta} (-tanh r)) S[m-2]  -- one dependent chain
__global__ void __launch_bounds__(256) k_squeezed(GateParams p) {
    const int M = p.shape[0];
    const double *__restrict__ sq = p.sq;
    c128 *S = p.out;
    for (int m = 1 + 2 * (int)threadIdx.x; m < M; m += 2 * blockDim.x) S[m] = c_make(0.0, 0.0);
    if (threadIdx.x == 0) {
        c128 prev = c_make(p.r1, 0.0);
        S[0] = prev;
        for (int m = 2; m < M; m += 2) {
            prev = c_mul(r_times_c(__ddiv_rn(sq[m - 1], sq[m]), p.z0), prev);
            S[m] = prev;
        }
    }
}

// ---- beamsplitter, vanilla rule ---------------------------------------------------------------------------------------------
// face q = 0:  G[m, n, p = m+n, 0] = ((ct sqrt m) / sqrt p) G[m-1, n, p-1, 0] + ((st sqrt n) / sqrt p) G[m, n-1, p-1, 0]
// one CTA, levels p = 1, 2, ...
__global__ void __launch_bounds__(1024) k_bs_face(GateParams p) {
    const int M = p.shape[0], N = p.shape[1], P = p.shape[2], Q = p.shape[3];
    const double *__restrict__ sq = p.sq;
    c128 *G = p.out;
    const double ct = p.r0;
    const c128 st = p.z0;
    const long long sM = (long long)N * P * Q, sN = (long long)P * Q, sP = Q;
    if (threadIdx.x == 0) G[0] = c_make(1.0, 0.0);
    __syncthreads();
    // the reference's face loop is `for n in range(N - m)` (beamsplitter.py:67), i.e. p = m + n < N as well as < P: for M > N the
    // face entries with p >= N stay zero there, and so they do here (identical results on identical inputs)
    const int pmax = (P - 1) < (N - 1) ? (P - 1) : (N - 1);
    for (int lv = 1; lv <= pmax; lv++) {
        const int m_lo = 0, m_hi = lv < M - 1 ? lv : M - 1;
        for (int m = m_lo + (int)threadIdx.x; m <= m_hi; m += blockDim.x) {
            const int n = lv - m;
            c128 v = c_make(0.0, 0.0);
            if (m > 0) v = r_times_c(__ddiv_rn(__dmul_rn(ct, sq[m]), sq[lv]), G[(m - 1) * sM + n * sN + (lv - 1) * sP]);
            if (n > 0) {
                const c128 t = c_mul(c_over_r(r_times_c(sq[n], st), sq[lv]), G[m * sM + (n - 1) * sN + (lv - 1) * sP]);
                v = m > 0 ? c_add(v, t) : t;
            }
            G[m * sM + n * sN + lv * sP] = v;
        }
        __syncthreads();
    }
}
// q >= 1:  G[m, n, p, q = m+n-p] = (((-conj st) sqrt m) / sqrt q) G[m-1, n, p, q-1] + ((ct sqrt n) / sqrt q) G[m, n-1, p, q-1]
// the recurrence never leaves its p-slice: CTA p walks the levels L = m + n = p + 1 ... of slice p
__global__ void __launch_bounds__(256) k_bs_slices(GateParams p) {
    const int M = p.shape[0], N = p.shape[1], P = p.shape[2], Q = p.shape[3];
    const double *__restrict__ sq = p.sq;
    c128 *G = p.out;
    const double ct = p.r0;
    const c128 mstc = make_double2(-p.z0.x, p.z0.y);   // -conj(st)
    const long long sM = (long long)N * P * Q, sN = (long long)P * Q, sP = Q;
    const int pp = blockIdx.x;
    (void)P;
    const int Lmax = (M + N - 2) < (pp + Q - 1) ? (M + N - 2) : (pp + Q - 1);
    for (int L = pp + 1; L <= Lmax; L++) {
        const int q = L - pp;
        const int m_lo = L - (N - 1) > 0 ? L - (N - 1) : 0, m_hi = L < M - 1 ? L : M - 1;
        for (int m = m_lo + (int)threadIdx.x; m <= m_hi; m += blockDim.x) {
            const int n = L - m;
            c128 v = c_make(0.0, 0.0);
            if (m > 0) v = c_mul(c_over_r(r_times_c(sq[m], mstc), sq[q]), G[(m - 1) * sM + n * sN + pp * sP + q - 1]);
            if (n > 0) {
                const c128 t = r_times_c(__ddiv_rn(__dmul_rn(ct, sq[n]), sq[q]), G[m * sM + (n - 1) * sN + pp * sP + q - 1]);
                v = m > 0 ? c_add(v, t) : t;
            }
            G[m * sM + n * sN + pp * sP + q] = v;
        }
        __syncthreads();
    }
}

// ---- beamsplitter, stable rule: average over every available pivot; one launch per level L = m + n (reads level L - 1) --------
__global__ void __launch_bounds__(256) k_bs_stable_level(GateParams p, int L) {
    const int M = p.shape[0], N = p.shape[1], P = p.shape[2], Q = p.shape[3];
    const double *__restrict__ sq = p.sq;
    c128 *G = p.out;
    const double ct = p.r0;
    const c128 st = p.z0, stc = make_double2(p.z0.x, -p.z0.y), mstc = make_double2(-p.z0.x, p.z0.y);
    const long long sM = (long long)N * P * Q, sN = (long long)P * Q, sP = Q;
    const int m_lo = L - (N - 1) > 0 ? L - (N - 1) : 0, m_hi = L < M - 1 ? L : M - 1;
    const int p_lo = L - (Q - 1) > 0 ? L - (Q - 1) : 0, p_hi = L < P - 1 ? L : P - 1;
    const int nm = m_hi - m_lo + 1, np = p_hi - p_lo + 1;
    if (nm <= 0 || np <= 0) return;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nm * np) return;
    const int m = m_lo + (int)(t / np), pp = p_lo + (int)(t % np);
    const int n = L - m, q = L - pp;
    // neighbours (zero when outside: the reference reads a wrapped entry times sqrt(0))
    const c128 zero = c_make(0.0, 0.0);
    const c128 g_mp = (m > 0 && pp > 0) ? G[(m - 1) * sM + n * sN + (pp - 1) * sP + q] : zero;   // G[m-1, n, p-1, q]
    const c128 g_mq = (m > 0 && q > 0) ? G[(m - 1) * sM + n * sN + pp * sP + q - 1] : zero;      // G[m-1, n, p, q-1]
    const c128 g_np = (n > 0 && pp > 0) ? G[m * sM + (n - 1) * sN + (pp - 1) * sP + q] : zero;   // G[m, n-1, p-1, q]
    const c128 g_nq = (n > 0 && q > 0) ? G[m * sM + (n - 1) * sN + pp * sP + q - 1] : zero;      // G[m, n-1, p, q-1]
    c128 val = zero;
    int piv = 0;
    if (q == 0) {   // the q = 0 face (beamsplitter.py:119-137): pivots m, n, p with the (p-1, 0) neighbours
        if (m > 0) { val = c_add(val, r_times_c(__ddiv_rn(__dmul_rn(ct, sq[pp]), sq[m]), g_mp)); piv++; }
        if (n > 0) { val = c_add(val, c_mul(c_over_r(r_times_c(sq[pp], st), sq[n]), g_np)); piv++; }
        if (pp > 0) {
            val = c_add(val, c_add(r_times_c(__ddiv_rn(__dmul_rn(ct, sq[m]), sq[pp]), g_mp),
                                   c_mul(c_over_r(r_times_c(sq[n], st), sq[pp]), g_np)));
            piv++;
        }
    } else {        // (beamsplitter.py:140-171)
        if (m > 0) {
            val = c_add(val, c_sub(r_times_c(__ddiv_rn(__dmul_rn(ct, sq[pp]), sq[m]), g_mp),
                                   c_mul(c_over_r(r_times_c(sq[q], stc), sq[m]), g_mq)));
            piv++;
        }
        if (n > 0) {
            val = c_add(val, c_add(c_mul(c_over_r(r_times_c(sq[pp], st), sq[n]), g_np),
                                   r_times_c(__ddiv_rn(__dmul_rn(ct, sq[q]), sq[n]), g_nq)));
            piv++;
        }
        if (pp > 0) {
            val = c_add(val, c_add(r_times_c(__ddiv_rn(__dmul_rn(ct, sq[m]), sq[pp]), g_mp),
                                   c_mul(c_over_r(r_times_c(sq[n], st), sq[pp]), g_np)));
            piv++;
        }
        val = c_add(val, c_add(c_mul(c_over_r(r_times_c(sq[m], mstc), sq[q]), g_mq),
                               r_times_c(__ddiv_rn(__dmul_rn(ct, sq[n]), sq[q]), g_nq)));
        piv++;
    }
    G[m * sM + n * sN + pp * sP + q] = c_over_r(val, (double)piv);
}

// ---- displacement ---------------------------------------------------------------------------------------------------
// pass 1: thread d walks the diagonal n - m = d: Laguerre three-term recurrence L_{m+1} = ((2m + 1 + d - x) L_m - (m + d) L_{m-1}) / (m+1)
// (displacement.py:68-82), stored at W[n, m]; pass 2: every entry through the log-domain closed form (displacement.py:52-63)
__global__ void __launch_bounds__(128) k_disp_laguerre(GateParams p) {
    const int N = p.shape[0], M = p.shape[1];   // N >= M (the host swaps)
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= N) return;
    const double x = p.r0;   // |alpha|^2
    const int m_max = M < N - d ? M : N - d;
    double *W = (double *)p.out;   // Laguerre values (real) parked in the real parts
    double l0 = 1.0, lm1 = 0.0;
    for (int m = 0; m < m_max; m++) {
        W[2 * ((long long)(m + d) * M + m)] = l0;
        const double a = __dsub_rn((double)(2 * m + 1 + d), x);
        const double nxt = __ddiv_rn(__dsub_rn(__dmul_rn(a, l0), __dmul_rn((double)(m + d), lm1)), (double)(m + 1));
        lm1 = l0; l0 = nxt;
    }
}
__global__ void __launch_bounds__(256) k_disp_fill(GateParams p, const double *logfac) {
    const int N = p.shape[0], M = p.shape[1];
    const bool flipped = p.flag != 0;
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= (long long)N * M) return;
    const int n = (int)(f / M), m = (int)(f % M);
    if (n < m) return;    // the strict upper triangle is written by its mirror entry (and never read: L is parked at n >= m only)
    const int d = n - m;
    const double r = p.r1, phi = p.r2, x = p.r0;
    c128 *D = p.out;
    const double L = ((const double *)D)[2 * f];
    // sign * exp( (logfac[m] - logfac[n]) / 2 + d log r - x / 2 + log|L| ) * sign(L) * e^{+-i phi d}
    const double sign = (flipped && n > m && (d & 1)) ? -1.0 : 1.0;
    const double cj = (flipped && n > m) ? -1.0 : 1.0;
    const double ex = 0.5 * (logfac[m] - logfac[n]) + (d ? (double)d * log(r) : 0.0) - x / 2.0 + log(fabs(L));
    const double mag = (L == 0.0) ? 0.0 : sign * (L < 0.0 ? -1.0 : 1.0) * exp(ex);
    double sn, cs;
    sincos(cj * phi * (double)d, &sn, &cs);
    const c128 v = make_double2(mag * cs, mag * sn);
    D[f] = v;
    // D[m, n] = (-1)^d conj(D[n, m]) inside the leading M x M block (displacement.py:63-64)
    if (d > 0 && n < M) D[(long long)m * M + n] = make_double2((d & 1) ? -v.x : v.x, (d & 1) ? v.y : -v.y);
}
// out-of-place transpose for the flipped case (cutoffs[0] < cutoffs[1])
__global__ void __launch_bounds__(256) k_transpose(const c128 *in, c128 *out, int rows, int cols) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= (long long)rows * cols) return;
    const int r = (int)(f / cols), c = (int)(f % cols);
    out[(long long)c * rows + r] = in[f];
}

// jacobian_displacement (displacement.py:117-139): dD/dalpha = -conj(alpha)/2 D + sqrt(m) D[m-1, n];  dD/dconj(alpha) = -alpha/2 D - sqrt(n) D[m, n-1]
// grad_displacement (displacement.py:85-114):      dT/dr = -r T + sqrt(m) e^{i phi} T[m-1, n] - sqrt(n) e^{-i phi} T[m, n-1]
//                                                  dT/dphi = sqrt(m) i alpha T[m-1, n] + sqrt(n) i conj(alpha) T[m, n-1]
__global__ void __launch_bounds__(256) k_disp_derivs(GateParams p, const c128 *D, c128 *o1, c128 *o2, int kind) {
    const int M = p.shape[0], N = p.shape[1];
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= (long long)M * N) return;
    const int m = (int)(f / N), n = (int)(f % N);
    const c128 d = D[f];
    const c128 up = m > 0 ? D[f - N] : c_make(0.0, 0.0), left = n > 0 ? D[f - 1] : c_make(0.0, 0.0);
    const double sm = sqrt((double)m), sn = sqrt((double)n);
    const c128 al = p.z0, alc = make_double2(p.z0.x, -p.z0.y);
    if (kind == 0) {
        const c128 a = c_mul(make_double2(-0.5 * alc.x, -0.5 * alc.y), d), b = c_mul(make_double2(-0.5 * al.x, -0.5 * al.y), d);
        o1[f] = make_double2(a.x + sm * up.x, a.y + sm * up.y);
        o2[f] = make_double2(b.x - sn * left.x, b.y - sn * left.y);
    } else {
        const double r = p.r1;
        const c128 ei = make_double2(p.r0, p.r2), eic = make_double2(p.r0, -p.r2);   // (cos phi, +- sin phi)
        const c128 t1 = c_mul(make_double2(sm * ei.x, sm * ei.y), up), t2 = c_mul(make_double2(sn * eic.x, sn * eic.y), left);
        o1[f] = make_double2(-r * d.x + t1.x - t2.x, -r * d.y + t1.y - t2.y);
        const c128 ia = make_double2(-al.y, al.x), iac = make_double2(-alc.y, alc.x);   // i alpha, i conj(alpha)
        const c128 u1 = c_mul(make_double2(sm * ia.x, sm * ia.y), up), u2 = c_mul(make_double2(sn * iac.x, sn * iac.y), left);
        o2[f] = make_double2(u1.x + u2.x, u1.y + u2.y);
    }
}

// ---- support masks of the gate VJPs: g_out = g on the index set the reference's loops visit, 0 elsewhere ------------------
// kind 0: beamsplitter (m + n == p + q); kind 1: squeezer ((m + n) even); kind 2: squeezed (m even).  The origin stays in: its step
// gradient is zero anyway (every weight carries sqrt(0)) and the reference's dLdC = sum(G * dLdG) includes it.
__global__ void __launch_bounds__(256) k_gate_mask(const c128 *g, c128 *out, long long n_total, int kind, int s1, int s2, int s3) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_total) return;
    bool keep;
    if (kind == 0) {
        const int q = (int)(f % s3);
        long long t = f / s3;
        const int pp = (int)(t % s2); t /= s2;
        const int n = (int)(t % s1), m = (int)(t / s1);
        keep = (m + n == pp + q);
    } else if (kind == 1) {
        const int n = (int)(f % s1), m = (int)(f / s1);
        keep = (((m + n) & 1) == 0);
    } else {
        keep = ((f & 1) == 0);
    }
    out[f] = keep ? g[f] : c_make(0.0, 0.0);
}
// generic VJP output (dLdA symmetrised) -> the un-symmetrised upper-triangular sums the gate chain rules use
__global__ void k_gate_unsym(const c128 *sym, c128 *out, int D) {
    const int t = threadIdx.x;
    if (t >= D * D) return;
    const int i = t / D, j = t % D;
    const c128 s = sym[t];
    out[t] = i == j ? s : (i < j ? make_double2(2.0 * s.x, 2.0 * s.y) : c_make(0.0, 0.0));
}

__global__ void k_set_one(c128 *p) { if (threadIdx.x == 0) p[0] = c_make(1.0, 0.0); }

// ---- launchers ------------------------------------------------------------------------------------------------------------
cudaError_t mmh_launch_squeezer(const GateParams &p, cudaStream_t st) {
    k_squeezer<<<1, 1024, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t mmh_launch_squeezed(const GateParams &p, cudaStream_t st) {
    k_squeezed<<<1, 256, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t mmh_launch_beamsplitter(const GateParams &p, bool stable, long long *launches, cudaStream_t st) {
    const int M = p.shape[0], N = p.shape[1], P = p.shape[2], Q = p.shape[3];
    cudaError_t e = cudaMemsetAsync(p.out, 0, sizeof(c128) * (size_t)M * N * P * Q, st);
    if (e != cudaSuccess) return e;
    if (!stable) {
        k_bs_face<<<1, 1024, 0, st>>>(p);
        k_bs_slices<<<P, 256, 0, st>>>(p);
        *launches = 2;
        return cudaGetLastError();
    }
    GateParams q = p;
    k_set_one<<<1, 32, 0, st>>>(p.out);   // level 0: the origin
    *launches = 1;
    for (int L = 1; L <= M + N - 2; L++) {
        const int m_lo = L - (N - 1) > 0 ? L - (N - 1) : 0, m_hi = L < M - 1 ? L : M - 1;
        const int p_lo = L - (Q - 1) > 0 ? L - (Q - 1) : 0, p_hi = L < P - 1 ? L : P - 1;
        const long long cnt = (long long)(m_hi - m_lo + 1) * (p_hi - p_lo + 1);
        if (m_hi < m_lo || p_hi < p_lo) continue;
        k_bs_stable_level<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(q, L);
        (*launches)++;
    }
    return cudaGetLastError();
}
cudaError_t mmh_launch_displacement(const GateParams &p, const double *logfac, cudaStream_t st) {
    const int N = p.shape[0], M = p.shape[1];
    cudaError_t e = cudaMemsetAsync(p.out, 0, sizeof(c128) * (size_t)N * M, st);
    if (e != cudaSuccess) return e;
    k_disp_laguerre<<<(N + 127) / 128, 128, 0, st>>>(p);
    k_disp_fill<<<(unsigned)(((long long)N * M + 255) / 256), 256, 0, st>>>(p, logfac);
    return cudaGetLastError();
}
cudaError_t mmh_launch_transpose(const c128 *in, c128 *out, int rows, int cols, cudaStream_t st) {
    k_transpose<<<(unsigned)(((long long)rows * cols + 255) / 256), 256, 0, st>>>(in, out, rows, cols);
    return cudaGetLastError();
}
cudaError_t mmh_launch_disp_derivs(const GateParams &p, const c128 *D, c128 *o1, c128 *o2, int kind, cudaStream_t st) {
    const long long n = (long long)p.shape[0] * p.shape[1];
    k_disp_derivs<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, D, o1, o2, kind);
    return cudaGetLastError();
}
cudaError_t mmh_launch_gate_mask(const c128 *g, c128 *out, long long n_total, int kind, int s1, int s2, int s3, cudaStream_t st) {
    k_gate_mask<<<(unsigned)((n_total + 255) / 256), 256, 0, st>>>(g, out, n_total, kind, s1, s2, s3);
    return cudaGetLastError();
}
cudaError_t mmh_launch_gate_unsym(const c128 *sym, c128 *out, int D, cudaStream_t st) {
    k_gate_unsym<<<1, 32, 0, st>>>(sym, out, D);
    return cudaGetLastError();
}

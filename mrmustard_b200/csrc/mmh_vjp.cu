// mmh_vjp.cu — vector-Jacobian product of the Gaussian-to-Fock map (vanilla/gradients.py:25-116), sm_100a.
//
//   dLdb_i  = sum_k sqrt(k_i)             G[k - e_i]        g_k
//   U_ii    = sum_k 1/2 sqrt(k_i (k_i-1)) G[k - 2 e_i]      g_k        (k_i > 1)
//   U_ij    = sum_k sqrt(k_i k_j)         G[k - e_i - e_j]  g_k        (j > i)
//   dLdA    = (U + U^T) / 2,   dLdc = sum_k G_k g_k / c                 (holomorphic cotangent)
//
// The gradient does not depend on A or b.  It is a pure reduction over the lattice: each thread keeps
// the D(D+1)/2 + D + 1 complex accumulators in registers, the CTA reduces them with warp shuffles in a
// fixed tree, and a second tiny kernel sums the per-CTA partials in a fixed order (deterministic; no
// atomics).  HBM traffic is one read of G and one of dLdG (32 B per amplitude); the neighbour reads of G
// are served by L1/L2.
#include "mmh_params.cuh"
#ifndef MMH_VJP_MINB
#define MMH_VJP_MINB 1
#endif


// accumulator layout: [0, D) = db ; then upper triangle row-major (i, j>=i) ; last = dc
template <int DT>
struct VjpAcc {
    static constexpr int NTRI = DT * (DT + 1) / 2;
    static constexpr int NACC = DT + NTRI + 1;
};

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// numba complex_div_impl (Smith) — only the dLdc = sum / c division uses it
__device__ __forceinline__ c128 c_div_smith(c128 a, c128 b) {
    if (fabs(b.x) >= fabs(b.y)) {
        const double ratio = b.y / b.x, denom = b.x + b.y * ratio;
        return c_make((a.x + a.y * ratio) / denom, (a.y - a.x * ratio) / denom);
    }
    const double ratio = b.x / b.y, denom = b.x * ratio + b.y;
    return c_make((a.x * ratio + a.y) / denom, (a.y * ratio - a.x) / denom);
}

// write the final gradients of one lattice from its fully reduced accumulators (layout: db | upper triangle | dc):
// symmetrisation (dLdA + dLdA^T)/2 (gradients.py:82) and dLdc = sum / c (:80)
__device__ __forceinline__ void vjp_write_entry(const VjpParams &p, long long lat, int e, c128 s) {
    const int D = p.d.D;
    if (e < D) { p.db[lat * D + e] = s; return; }
    if (e == p.nacc - 1) { p.dc[lat] = c_div_smith(s, p.c[lat]); return; }
    int r = e - D, i = 0;
    while (r >= D - i) { r -= D - i; i++; }
    const int j = i + r;
    c128 *dA = p.dA + lat * D * D;
    if (i == j) dA[i * D + i] = s;
    else { const c128 h = c_make(0.5 * s.x, 0.5 * s.y); dA[i * D + j] = h; dA[j * D + i] = h; }
}

// IT = unsigned (N < 2^31) or long long.  Threads walk the lattice with a fixed stride; the multi-index is advanced
// by adding the mixed-radix digits of the stride (precomputed on the host) with carries — no division in the loop.
template <int DT, typename IT>
__global__ void __launch_bounds__(256, MMH_VJP_MINB) k_vjp_partial(VjpParams p) {
    constexpr int NACC = VjpAcc<DT>::NACC;
    const LatticeDesc &d = p.d;
    const IT N = (IT)d.N;
    const long long lat = blockIdx.x;
    const c128 *G = p.G + lat * d.N;
    const c128 *g = p.g + lat * d.N;
    const double *__restrict__ sq = p.sq;

    IT st[DT];
    int sh[DT], dig[DT];
#pragma unroll
    for (int i = 0; i < DT; i++) { st[i] = (IT)d.strides[i]; sh[i] = d.shape[i]; dig[i] = p.stride_digits[i]; }

    c128 acc[NACC];
#pragma unroll
    for (int e = 0; e < NACC; e++) acc[e] = c_make(0.0, 0.0);

    const IT stride = (IT)gridDim.y * blockDim.x;
    IT f = (IT)blockIdx.y * blockDim.x + threadIdx.x;
    int k[DT];
    {
        IT rem = f < N ? f : 0;
#pragma unroll
        for (int i = 0; i < DT; i++) { k[i] = (int)(rem / st[i]); rem -= (IT)k[i] * st[i]; }
    }
    for (; f < N; f += stride) {
        // Every neighbour is fetched unconditionally in one batch of independent loads (a missing neighbour, k_i = 0, reads
        // G[f] instead and meets a zero weight: sq[0] = 0), so that the ~D(D+1)/2 + D + 2 loads of a point are all in flight
        // together instead of one load -> fma pair per data-dependent branch.
        const c128 gk = g[f];
        const c128 Gf = G[f];
        c128 Gi[DT], Gii[DT], Gij[NACC];
        double wi[DT];
#pragma unroll
        for (int i = 0; i < DT; i++) {
            const bool ok = k[i] >= 1;
            const IT pivot = ok ? f - st[i] : f;
            wi[i] = sq[k[i]];
            Gi[i] = G[pivot];
            Gii[i] = G[k[i] >= 2 ? pivot - st[i] : f];
#pragma unroll
            for (int j = i + 1; j < DT; j++) Gij[i * DT + j - (i + 1) * (i + 2) / 2] = G[(ok && k[j] >= 1) ? pivot - st[j] : f];
        }
        c_fma(acc[NACC - 1], Gf, gk);  // dLdc numerator (gradients.py:80)
        int e = DT;
#pragma unroll
        for (int i = 0; i < DT; i++) {
            c_fma(acc[i], c_scale(Gi[i], wi[i]), gk);                                                   // :68
            c_fma(acc[e], c_scale(Gii[i], 0.5 * wi[i] * sq[k[i] >= 1 ? k[i] - 1 : 0]), gk);             // :69-73
#pragma unroll
            for (int j = i + 1; j < DT; j++)
                c_fma(acc[e + (j - i)], c_scale(Gij[i * DT + j - (i + 1) * (i + 2) / 2], wi[i] * wi[j]), gk);   // :74-75
            e += DT - i;
        }
        // advance the multi-index by `stride` (mixed-radix add with carry, last mode first)
        int carry = 0;
#pragma unroll
        for (int i = DT - 1; i >= 0; i--) {
            int t = k[i] + dig[i] + carry;
            carry = 0;
            if (i > 0) { while (t >= sh[i]) { t -= sh[i]; carry++; } }
            k[i] = t;
        }
    }

    // CTA reduction: warp shuffles, then one smem pass (fixed order)
    __shared__ double red[8][2 * NACC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < NACC; e++) {
        const double re = warp_sum(acc[e].x), im = warp_sum(acc[e].y);
        if (lane == 0) { red[warp][2 * e] = re; red[warp][2 * e + 1] = im; }
    }
    __syncthreads();
    const int nwarps = (blockDim.x + 31) >> 5;
    if (p.nblk == 1) {   // the CTA holds the whole lattice: write the gradients directly
        for (int e = threadIdx.x; e < NACC; e += blockDim.x) {
            c128 s = c_make(0.0, 0.0);
            for (int w = 0; w < nwarps; w++) { s.x += red[w][2 * e]; s.y += red[w][2 * e + 1]; }
            vjp_write_entry(p, lat, e, s);
        }
        return;
    }
    for (int t = threadIdx.x; t < 2 * NACC; t += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < nwarps; w++) s += red[w][t];
        double *out = (double *)(p.partial + (lat * p.nblk + blockIdx.y) * (long long)NACC);
        out[t] = s;
    }
}

// Generic (any D) fallback: one launch per accumulator family would be wasteful, so each CTA row
// (blockIdx.z) owns one accumulator entry e and recomputes the index decode.  Used for D > 8.
__global__ void __launch_bounds__(256) k_vjp_partial_generic(VjpParams p) {
    const LatticeDesc &d = p.d;
    const int D = d.D;
    const long long N = d.N;
    const long long lat = blockIdx.x;
    const int e = blockIdx.z;  // entry index in the accumulator layout
    const c128 *G = p.G + lat * N;
    const c128 *g = p.g + lat * N;
    const double *__restrict__ sq = p.sq;
    // decode e -> kind
    int ei = -1, ej = -1;  // ei<0: dc ; ej<0: db[ei]
    if (e < D) { ei = e; }
    else if (e < p.nacc - 1) {
        int r = e - D;
        for (int i = 0; i < D; i++) { if (r < D - i) { ei = i; ej = i + r; break; } r -= D - i; }
    }
    c128 acc = c_make(0.0, 0.0);
    for (long long f = (long long)blockIdx.y * blockDim.x + threadIdx.x; f < N;
         f += (long long)gridDim.y * blockDim.x) {
        const c128 gk = g[f];
        if (ei < 0) { c_fma(acc, G[f], gk); continue; }
        const int ki = (int)((f / d.strides[ei]) % d.shape[ei]);
        if (ki < 1) continue;
        const long long pivot = f - d.strides[ei];
        const double wi = sq[ki];
        if (ej < 0) c_fma(acc, c_scale(G[pivot], wi), gk);
        else if (ej == ei) { if (ki > 1) c_fma(acc, c_scale(G[pivot - d.strides[ei]], 0.5 * wi * sq[ki - 1]), gk); }
        else {
            const int kj = (int)((f / d.strides[ej]) % d.shape[ej]);
            if (kj >= 1) c_fma(acc, c_scale(G[pivot - d.strides[ej]], wi * sq[kj]), gk);
        }
    }
    __shared__ double red[8][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double re = warp_sum(acc.x), im = warp_sum(acc.y);
    if (lane == 0) { red[warp][0] = re; red[warp][1] = im; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) s += red[w][threadIdx.x];
        double *out = (double *)(p.partial + (lat * p.nblk + blockIdx.y) * (long long)p.nacc + e);
        out[threadIdx.x] = s;
    }
}

// one warp per (lattice, accumulator entry): lanes sum the per-CTA partials with a fixed stride, then a fixed
// shuffle tree (deterministic), then the symmetrisation / division epilogue
__global__ void __launch_bounds__(128) k_vjp_finish(VjpParams p) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= p.batch * p.nacc) return;
    const long long lat = w / p.nacc;
    const int e = (int)(w - lat * p.nacc);
    const c128 *part = p.partial + lat * p.nblk * (long long)p.nacc + e;
    c128 s = c_make(0.0, 0.0);
    for (int blk = lane; blk < p.nblk; blk += 32) { const c128 v = part[(long long)blk * p.nacc]; s.x += v.x; s.y += v.y; }
    s.x = warp_sum(s.x); s.y = warp_sum(s.y);
    if (lane == 0) vjp_write_entry(p, lat, e, s);
}


// ---------------------------------------------------------------------------------------------------
// ONE large 4-index lattice (vanilla_vjp_numba on a 4-mode ket at cutoff 40: cfg5): plane tiles staged by the TMA engine.
// k_vjp_partial fetches the 15 neighbour amplitudes of every point with scattered 16-byte loads (160 registers, 12 warps per
// SM: latency x occupancy bound, 0.19 of HBM).  The neighbours of the points of plane (k0, k1) -- its S2 x S3 amplitudes are
// contiguous -- all lie in six planes:
//     A0 = (k0, k1): G_k, k-e2, k-e3, k-2e2, k-2e3, k-e2-e3      A1 = (k0, k1-1): k-e1, k-e1-e2, k-e1-e3      A2 = (k0, k1-2): k-2e1
//     B0 = (k0-1, k1): k-e0, k-e0-e2, k-e0-e3                     B1 = (k0-1, k1-1): k-e0-e1                   C0 = (k0-2, k1): k-2e0
// A CTA takes a task (k0, segment of k1) and walks k1: the planes A and B rotate through shared memory (4 + 3 slots), each new
// plane arrives with ONE bulk asynchronous copy (cp.async.bulk global -> shared, mbarrier complete_tx: SASS UBLKCP) issued one
// step ahead, and every thread then reads its 14 shared-memory neighbours at unit stride (LDS.128, conflict free); k-2e0 and the
// cotangent g_k are coalesced global loads.  Per point the arithmetic is the one of k_vjp_partial.  Deterministic: static task
// assignment, fixed-order CTA reduction, partials summed by k_vjp_finish.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void vp_mbar_init(unsigned addr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void vp_mbar_expect_tx(unsigned addr, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vp_mbar_wait(unsigned addr, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void vp_bulk_prefetch_l2(const void *gsrc, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vp_bulk_g2s(unsigned sdst, const void *gsrc, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}

#define MMH_VP_T 512     // threads per CTA at most
#define MMH_VP_NP 4      // points of a plane per thread at most
__global__ void __launch_bounds__(MMH_VP_T, 1) k_vjp_planes(VjpParams p, int nseg) {
    extern __shared__ c128 vp_smem[];
    constexpr int NACC = VjpAcc<4>::NACC;   // db0..3 | U00 U01 U02 U03 U11 U12 U13 U22 U23 U33 | dc
    const LatticeDesc &d = p.d;
    const int S0 = d.shape[0], S1 = d.shape[1], S2 = d.shape[2], S3 = d.shape[3];
    const int PL = S2 * S3;                               // amplitudes of one plane
    const unsigned PLB = (unsigned)PL * 16u;
    const int tid = threadIdx.x, T = blockDim.x;
    const long long lat = blockIdx.y;
    const c128 *G = p.G + lat * d.N;
    const c128 *g = p.g + lat * d.N;
    const double *__restrict__ sq = p.sq;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(vp_smem);
    int kq[MMH_VP_NP];                                    // this thread's points of a plane: k2 << 16 | k3, -1 = none (the same in every plane)
#pragma unroll
    for (int n = 0; n < MMH_VP_NP; n++) {
        const int q = tid + n * T;
        kq[n] = q < PL ? ((q / S3) << 16) | (q % S3) : -1;
    }
    const unsigned mbar = sbase + 7u * PLB;               // two mbarriers behind the seven plane slots
    if (tid == 0) { vp_mbar_init(mbar, 1u); vp_mbar_init(mbar + 8u, 1u); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    c128 acc[NACC];
#pragma unroll
    for (int e = 0; e < NACC; e++) acc[e] = c_make(0.0, 0.0);
    unsigned cnt = 0;                                     // steps done by this CTA: barrier cnt & 1, parity (cnt >> 1) & 1

    const int ntask = S0 * nseg;
    for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
        const int k0 = task / nseg, seg = task - k0 * nseg;
        const int k1a = (int)(((long long)seg * S1) / nseg), k1b = (int)(((long long)(seg + 1) * S1) / nseg);
        if (k1a >= k1b) continue;
        const c128 *Gk0 = G + (long long)k0 * S1 * PL, *Gk0m = G + (long long)(k0 >= 1 ? k0 - 1 : 0) * S1 * PL;
        const double w0 = sq[k0], h00 = 0.5 * w0 * sq[k0 >= 1 ? k0 - 1 : 0];
        __syncthreads();                                  // every thread has left the previous task's planes
        if (tid == 0) {                                   // the planes of the first step
            unsigned bytes = 0;
            for (int k1 = k1a - 2; k1 <= k1a; k1++)
                if (k1 >= 0) { vp_bulk_g2s(sbase + (unsigned)(k1 & 3) * PLB, Gk0 + (long long)k1 * PL, PLB, mbar + 8u * (cnt & 1u)); bytes += PLB; }
            if (k0 >= 1)
                for (int k1 = k1a - 1; k1 <= k1a; k1++)
                    if (k1 >= 0) { vp_bulk_g2s(sbase + (unsigned)(4 + k1 % 3) * PLB, Gk0m + (long long)k1 * PL, PLB, mbar + 8u * (cnt & 1u)); bytes += PLB; }
            vp_mbar_expect_tx(mbar + 8u * (cnt & 1u), bytes);
        }
        for (int k1 = k1a; k1 < k1b; k1++, cnt++) {
            if (tid == 0 && k1 + 1 < k1b) {               // one step ahead: A(k1+1) replaces A(k1-3), B(k1+1) replaces B(k1-2)
                const unsigned mb = mbar + 8u * ((cnt + 1u) & 1u);
                vp_bulk_g2s(sbase + (unsigned)((k1 + 1) & 3) * PLB, Gk0 + (long long)(k1 + 1) * PL, PLB, mb);
                if (k0 >= 1) vp_bulk_g2s(sbase + (unsigned)(4 + (k1 + 1) % 3) * PLB, Gk0m + (long long)(k1 + 1) * PL, PLB, mb);
                vp_mbar_expect_tx(mb, k0 >= 1 ? 2u * PLB : PLB);
                // the cotangent plane of the next step is cold (read once, from HBM): pull it into L2 a step ahead
                vp_bulk_prefetch_l2(g + ((long long)k0 * S1 + k1 + 1) * PL, PLB);
            }
            const c128 *gp = g + ((long long)k0 * S1 + k1) * PL;
            const c128 *Cp = G + ((long long)(k0 >= 2 ? k0 - 2 : k0) * S1 + k1) * PL;   // plane (k0-2, k1): weight 0 when absent
            const c128 *sA0 = vp_smem + (size_t)(k1 & 3) * PL;
            const c128 *sA1 = k1 >= 1 ? vp_smem + (size_t)((k1 - 1) & 3) * PL : sA0;
            const c128 *sA2 = k1 >= 2 ? vp_smem + (size_t)((k1 - 2) & 3) * PL : sA0;
            const c128 *sB0 = k0 >= 1 ? vp_smem + (size_t)(4 + k1 % 3) * PL : sA0;
            const c128 *sB1 = (k0 >= 1 && k1 >= 1) ? vp_smem + (size_t)(4 + (k1 - 1) % 3) * PL : sA0;
            const double w1 = sq[k1], h11 = 0.5 * w1 * sq[k1 >= 1 ? k1 - 1 : 0], w01 = w0 * w1;
            // the global operands of all of this thread's points are issued before the wait (their latency overlaps it)
            c128 gq[MMH_VP_NP], Cq[MMH_VP_NP];
#pragma unroll
            for (int n = 0; n < MMH_VP_NP; n++) {
                const int q = tid + n * T;
                gq[n] = kq[n] >= 0 ? gp[q] : c_make(0.0, 0.0);
                Cq[n] = kq[n] >= 0 ? Cp[q] : c_make(0.0, 0.0);
            }
            vp_mbar_wait(mbar + 8u * (cnt & 1u), (cnt >> 1) & 1u);
#pragma unroll
            for (int n = 0; n < MMH_VP_NP; n++) {
                if (kq[n] < 0) continue;
                const int q = tid + n * T;
                const c128 gk = gq[n], Ck = Cq[n];
                const int k2 = kq[n] >> 16, k3 = kq[n] & 0xffff;
                const double w2 = sq[k2], w3 = sq[k3];
                const int o2 = k2 >= 1 ? S3 : 0, o22 = k2 >= 2 ? 2 * S3 : 0, o3 = k3 >= 1 ? 1 : 0, o33 = k3 >= 2 ? 2 : 0;
                const int o23 = (k2 >= 1 && k3 >= 1) ? S3 + 1 : 0;
                // all shared-memory operands first (independent loads), then the arithmetic
                const c128 a_0 = sA0[q], a_2 = sA0[q - o2], a_3 = sA0[q - o3], a_22 = sA0[q - o22], a_33 = sA0[q - o33], a_23 = sA0[q - o23];
                const c128 a1_0 = sA1[q], a1_2 = sA1[q - o2], a1_3 = sA1[q - o3], a2_0 = sA2[q];
                const c128 b_0 = sB0[q], b_2 = sB0[q - o2], b_3 = sB0[q - o3], b1_0 = sB1[q];
                c_fma(acc[NACC - 1], a_0, gk);                                              // dLdc numerator (gradients.py:80)
                c_fma(acc[0], c_scale(b_0, w0), gk);                                        // db0   :68
                c_fma(acc[4], c_scale(Ck, h00), gk);                                        // U00   :69-73
                c_fma(acc[5], c_scale(b1_0, w01), gk);                                      // U01   :74-75
                c_fma(acc[6], c_scale(b_2, w0 * w2), gk);                                   // U02
                c_fma(acc[7], c_scale(b_3, w0 * w3), gk);                                   // U03
                c_fma(acc[1], c_scale(a1_0, w1), gk);                                       // db1
                c_fma(acc[8], c_scale(a2_0, h11), gk);                                      // U11
                c_fma(acc[9], c_scale(a1_2, w1 * w2), gk);                                  // U12
                c_fma(acc[10], c_scale(a1_3, w1 * w3), gk);                                 // U13
                c_fma(acc[2], c_scale(a_2, w2), gk);                                        // db2
                c_fma(acc[11], c_scale(a_22, 0.5 * w2 * sq[k2 >= 1 ? k2 - 1 : 0]), gk);     // U22
                c_fma(acc[12], c_scale(a_23, w2 * w3), gk);                                 // U23
                c_fma(acc[3], c_scale(a_3, w3), gk);                                        // db3
                c_fma(acc[13], c_scale(a_33, 0.5 * w3 * sq[k3 >= 1 ? k3 - 1 : 0]), gk);     // U33
            }
            __syncthreads();                              // plane (k0, k1) consumed: the next step's prefetch may overwrite a slot
        }
    }

    // CTA reduction: warp shuffles, then one smem pass (fixed order)
    __shared__ double red[MMH_VP_T / 32][2 * NACC];
    const int warp = tid >> 5, lane = tid & 31, nwarps = T >> 5;
#pragma unroll
    for (int e = 0; e < NACC; e++) {
        const double re = warp_sum(acc[e].x), im = warp_sum(acc[e].y);
        if (lane == 0) { red[warp][2 * e] = re; red[warp][2 * e + 1] = im; }
    }
    __syncthreads();
    for (int t = tid; t < 2 * NACC; t += T) {
        double sum = 0.0;
        for (int w = 0; w < nwarps; w++) sum += red[w][t];
        double *out = (double *)(p.partial + (lat * p.nblk + blockIdx.x) * (long long)NACC);
        out[t] = sum;
    }
}

// shared memory of k_vjp_planes: seven plane slots + two mbarriers; 0 when the lattice does not qualify
size_t mmh_vjp_planes_smem(const LatticeDesc &d) {
    if (d.D != 4) return 0;
    const size_t pl = (size_t)d.shape[2] * d.shape[3];
    const size_t smem = 7 * pl * sizeof(c128) + 16;
    if (smem > 200 * 1024 || pl < 256 || pl > (size_t)MMH_VP_NP * MMH_VP_T || d.shape[0] < 4 || d.shape[1] < 4) return 0;
    return smem;
}

// tasks = shape[0] x nseg segments of k1; nseg is the count that leaves the fewest idle CTA-steps with one CTA per SM
cudaError_t mmh_launch_vjp_planes(const VjpParams &p_in, int sm_count, int *nblk_out, cudaStream_t st) {
    VjpParams p = p_in;
    const size_t smem = mmh_vjp_planes_smem(p.d);
    if (!smem) return cudaErrorInvalidValue;
    const int S0 = p.d.shape[0], S1 = p.d.shape[1];
    int best_nseg = 1;
    double best_cost = 1e300;
    for (int nseg = 1; nseg <= S1 && nseg <= 16; nseg++) {
        const int ntask = S0 * nseg, grid = ntask < sm_count ? ntask : sm_count;
        const int per_cta = (ntask + grid - 1) / grid;
        const double L = (double)S1 / nseg;
        const double cost = per_cta * (L + 1.0);   // steps of the busiest CTA + ~1 step of prologue (three extra plane loads) per task
        if (cost < best_cost) { best_cost = cost; best_nseg = nseg; }
    }
    if (const char *e = mmh_getenv("MMH_VJP_NSEG")) best_nseg = atoi(e) > 0 ? atoi(e) : best_nseg;
    const int ntask = S0 * best_nseg, grid = ntask < sm_count ? ntask : sm_count;
    p.nblk = grid;
    *nblk_out = grid;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(k_vjp_planes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smem_set = smem;
    }
    // threads: the multiple of 32 (<= 512) that leaves the fewest idle point slots at ceil(PL / T) points per thread
    const int PL = p.d.shape[2] * p.d.shape[3];
    int T = MMH_VP_T;
    double best_waste = 1e300;
    for (int t = 256; t <= MMH_VP_T; t += 32) {
        const int np = (PL + t - 1) / t;
        if (np > MMH_VP_NP) continue;
        const double waste = (double)np * t / PL;
        if (waste < best_waste - 1e-9 || (waste < best_waste + 1e-9 && t > T)) { best_waste = waste; T = t; }
    }
    if (const char *e = mmh_getenv("MMH_VJP_T")) T = atoi(e);
    k_vjp_planes<<<dim3(grid, (unsigned)p.batch), T, smem, st>>>(p, best_nseg);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const long long warps = p.batch * p.nacc;
    k_vjp_finish<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, st>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Batches of 2-index lattices (vanilla_batch_vjp_numba over 2-mode kets, gradients.py:85-116; cfg3): warp-synchronous row walk.
// k_vjp_partial fetches the five neighbours of every point through L1 (seven 16-byte loads per amplitude: L1 wavefronts bound
// it at 0.6 of the HBM roofline).  Here a lane owns R consecutive positions of a lattice row (the layout of k_march_lanes,
// mmh_lanes.cu) and the warp walks the rows k_0 = 0 .. S-1: G[k - e_0], G[k - 2 e_0] are the lane's registers of the two previous
// rows, G[k - e_1], G[k - 2 e_1], G[k - e_0 - e_1] its own registers or two shuffles from the lane below, so G and dLdG are each
// loaded exactly once (32 B per amplitude) and nothing is re-read through L1.  The six complex accumulators stay in
// registers; the ln lanes of a lattice are summed once at the end in a fixed order (deterministic) and the gradients are
// written directly (no partials, no finish kernel).
// ---------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(128) k_vjp_lanes(VjpParams p, int ln, int Lw) {
    __shared__ double red[4][32][13];
    const LatticeDesc &d = p.d;
    const int n1 = d.shape[1], S = d.shape[0];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long wl0 = ((long long)blockIdx.x * nw + warp) * Lw;
    pdl_trigger();
    pdl_wait();   // G is normally the output of the kernel right before this one
    if (wl0 >= p.batch) return;
    const int nlat = (int)(p.batch - wl0 < Lw ? p.batch - wl0 : Lw);
    const double *__restrict__ sq = p.sq;
    const int lw = lane / ln;
    const bool lane_act = lw < nlat;
    const int k0 = (lane - lw * ln) * R;
    const long long l = wl0 + (lane_act ? lw : 0);
    const c128 *Gr = p.G + l * d.N + k0;
    const c128 *gr = p.g + l * d.N + k0;
    bool act[R];
    double w1[R], w11[R];   // sqrt(k_1), 1/2 sqrt(k_1 (k_1 - 1))
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int k1 = k0 + r;
        act[r] = lane_act && k1 < n1;
        w1[r] = act[r] ? sq[k1] : 0.0;
        w11[r] = (act[r] && k1 >= 2) ? 0.5 * sq[k1] * sq[k1 - 1] : 0.0;
    }
    const bool first = k0 == 0;   // no left neighbours: the lane below belongs to another lattice
    c128 acc[6];                  // db_0, db_1, U_00, U_01, U_11, sum G g
#pragma unroll
    for (int e = 0; e < 6; e++) acc[e] = c_make(0.0, 0.0);
    c128 Ra[R], Rb[R], Rc[R];     // three rows in rotation: current, previous, the one before
#pragma unroll
    for (int r = 0; r < R; r++) Ra[r] = Rb[r] = Rc[r] = c_make(0.0, 0.0);
    c128 upP = c_make(0.0, 0.0);  // G[k_0 - 1, k0 - 1]: left neighbour of slot 0 in the previous row

    // one row: C is loaded, P1 / P2 are the two previous rows (gradients.py:68-75)
#define MMH_VJP_ROW(C, P1, P2)                                                                                  \
    {                                                                                                           \
        c128 gk[R];                                                                                             \
        _Pragma("unroll") for (int r = 0; r < R; r++) {                                                         \
            C[r] = act[r] ? Gr[r] : c_make(0.0, 0.0);                                                           \
            gk[r] = act[r] ? gr[r] : c_make(0.0, 0.0);                                                          \
        }                                                                                                       \
        Gr += n1; gr += n1;                                                                                     \
        const double w0 = sq[s], w00 = s >= 2 ? 0.5 * w0 * sq[s - 1] : 0.0;                                     \
        c128 l1 = make_double2(__shfl_up_sync(0xffffffffu, C[R - 1].x, 1), __shfl_up_sync(0xffffffffu, C[R - 1].y, 1)); \
        c128 l2 = make_double2(__shfl_up_sync(0xffffffffu, C[R - 2].x, 1), __shfl_up_sync(0xffffffffu, C[R - 2].y, 1)); \
        if (first) { l1 = c_make(0.0, 0.0); l2 = c_make(0.0, 0.0); }                                            \
        _Pragma("unroll") for (int r = 0; r < R; r++) {                                                         \
            const c128 L1 = r == 0 ? l1 : C[r >= 1 ? r - 1 : 0];                                                \
            const c128 L2 = r == 0 ? l2 : (r == 1 ? l1 : C[r >= 2 ? r - 2 : 0]);                                \
            const c128 PL = r == 0 ? upP : P1[r >= 1 ? r - 1 : 0];                                              \
            const c128 x = make_double2(gk[r].x * w0, gk[r].y * w0);                                            \
            c_fma(acc[5], C[r], gk[r]);                                                                         \
            c_fma(acc[0], P1[r], x);                                                                            \
            c_fma(acc[2], P2[r], make_double2(gk[r].x * w00, gk[r].y * w00));                                   \
            c_fma(acc[1], L1, make_double2(gk[r].x * w1[r], gk[r].y * w1[r]));                                  \
            c_fma(acc[4], L2, make_double2(gk[r].x * w11[r], gk[r].y * w11[r]));                                \
            c_fma(acc[3], PL, make_double2(x.x * w1[r], x.y * w1[r]));                                          \
        }                                                                                                       \
        upP = l1;                                                                                               \
    }
    int s = 0;
#pragma unroll 1
    for (; s + 2 < S; s += 3) {
        MMH_VJP_ROW(Ra, Rb, Rc)
        s++;
        MMH_VJP_ROW(Rc, Ra, Rb)
        s++;
        MMH_VJP_ROW(Rb, Rc, Ra)
        s -= 2;
    }
    if (s < S) { MMH_VJP_ROW(Ra, Rb, Rc) s++; }
    if (s < S) { MMH_VJP_ROW(Rc, Ra, Rb) s++; }
#undef MMH_VJP_ROW

    // the ln lanes of a lattice, summed in lane order by one thread per (lattice, accumulator)
#pragma unroll
    for (int e = 0; e < 6; e++) { red[warp][lane][2 * e] = acc[e].x; red[warp][lane][2 * e + 1] = acc[e].y; }
    __syncwarp();
    for (int idx = lane; idx < nlat * 6; idx += 32) {
        const int q = idx / 6, e = idx - q * 6;
        c128 t = c_make(0.0, 0.0);
        for (int m = 0; m < ln; m++) { t.x += red[warp][q * ln + m][2 * e]; t.y += red[warp][q * ln + m][2 * e + 1]; }
        vjp_write_entry(p, wl0 + q, e, t);
    }
}

cudaError_t mmh_launch_vjp_lanes(const VjpParams &p, int R, int ln, int Lw, int sm_count, cudaStream_t st) {
    int block = 128;   // 2 warps per CTA when the batch leaves less than ~8 CTAs per SM (see mmh_launch_march_lanes)
    if ((p.batch + 4LL * Lw - 1) / (4LL * Lw) < 8LL * sm_count) block = 64;
    if (const char *e = mmh_getenv("MMH_LANES_BLOCK")) block = atoi(e) == 64 ? 64 : (atoi(e) == 32 ? 32 : 128);
    const int nw = block / 32;
    const long long grid = (p.batch + (long long)nw * Lw - 1) / ((long long)nw * Lw);
    if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    const bool pdl = !mmh_getenv("MMH_NO_PDL");
    const dim3 g((unsigned)grid), b(block);
    switch (R) {
        case 2: return mmh_launch_ex(k_vjp_lanes<2>, g, b, 0, st, pdl, p, ln, Lw);
        case 3: return mmh_launch_ex(k_vjp_lanes<3>, g, b, 0, st, pdl, p, ln, Lw);
        case 4: return mmh_launch_ex(k_vjp_lanes<4>, g, b, 0, st, pdl, p, ln, Lw);
        case 5: return mmh_launch_ex(k_vjp_lanes<5>, g, b, 0, st, pdl, p, ln, Lw);
        case 6: return mmh_launch_ex(k_vjp_lanes<6>, g, b, 0, st, pdl, p, ln, Lw);
        case 7: return mmh_launch_ex(k_vjp_lanes<7>, g, b, 0, st, pdl, p, ln, Lw);
        case 8: return mmh_launch_ex(k_vjp_lanes<8>, g, b, 0, st, pdl, p, ln, Lw);
        default: return cudaErrorInvalidValue;
    }
}

template <typename IT>
static void launch_partial(const VjpParams &p, dim3 grid, int block, cudaStream_t st) {
    switch (p.d.D) {
        case 1: k_vjp_partial<1, IT><<<grid, block, 0, st>>>(p); break;
        case 2: k_vjp_partial<2, IT><<<grid, block, 0, st>>>(p); break;
        case 3: k_vjp_partial<3, IT><<<grid, block, 0, st>>>(p); break;
        case 4: k_vjp_partial<4, IT><<<grid, block, 0, st>>>(p); break;
        case 5: k_vjp_partial<5, IT><<<grid, block, 0, st>>>(p); break;
        case 6: k_vjp_partial<6, IT><<<grid, block, 0, st>>>(p); break;
        case 7: k_vjp_partial<7, IT><<<grid, block, 0, st>>>(p); break;
        case 8: k_vjp_partial<8, IT><<<grid, block, 0, st>>>(p); break;
        default: break;
    }
}

// resident CTAs per SM of the partial kernel (register bound: the batched loads of D >= 4 take 160+ registers)
template <typename IT>
static int partial_blocks_per_sm(int D, int block) {
    int n = 0;
    cudaError_t e = cudaErrorInvalidValue;
    switch (D) {
        case 1: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_vjp_partial<1, IT>, block, 0); break;
        case 2: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_vjp_partial<2, IT>, block, 0); break;
        case 3: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_vjp_partial<3, IT>, block, 0); break;
        case 4: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_vjp_partial<4, IT>, block, 0); break;
        case 5: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_vjp_partial<5, IT>, block, 0); break;
        case 6: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_vjp_partial<6, IT>, block, 0); break;
        case 7: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_vjp_partial<7, IT>, block, 0); break;
        case 8: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_vjp_partial<8, IT>, block, 0); break;
        default: break;
    }
    return (e == cudaSuccess && n > 0) ? n : 1;
}
int mmh_vjp_blocks_per_sm(const VjpParams &p, int block) {
    if (p.d.D > 8) return 4;
    return p.d.N < 0x7fffffffLL ? partial_blocks_per_sm<unsigned>(p.d.D, block) : partial_blocks_per_sm<long long>(p.d.D, block);
}

cudaError_t mmh_launch_vjp(const VjpParams &p_in, int grid_y, int block, cudaStream_t st) {
    VjpParams p = p_in;
    dim3 grid((unsigned)p.batch, grid_y, 1);  // x = lattice (may exceed 65535), y = CTA within the lattice
    // mixed-radix digits of the per-thread stride grid_y * block
    {
        long long s = (long long)grid_y * block;
        for (int i = p.d.D - 1; i >= 0; i--) {
            if (i > 0) { p.stride_digits[i] = (int)(s % p.d.shape[i]); s /= p.d.shape[i]; }
            else p.stride_digits[0] = (int)(s < (1 << 30) ? s : (1 << 30));
        }
    }
    if (p.d.D <= 8) {
        if (p.d.N < 0x7fffffffLL) launch_partial<unsigned>(p, grid, block, st);
        else launch_partial<long long>(p, grid, block, st);
    } else {
        grid.z = p.nacc;
        k_vjp_partial_generic<<<grid, block, 0, st>>>(p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (p.nblk == 1 && p.d.D <= 8) return cudaSuccess;   // gradients were written by the partial kernel
    const long long warps = p.batch * p.nacc;
    k_vjp_finish<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, st>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Bilinear overlap  s = sum_k x_k y_k  of two lattices (no conjugation: the caller passes conj(target)) -- the scalar a fidelity
// cost needs from the lattice (cfg5: 1 - |<target|G>|^2), reduced on the device so that a training step reads back 22 numbers
// instead of the lattice.  Deterministic: fixed grid, fixed-order tree per CTA, partials summed in index order by one CTA.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dot_partial(const c128 *__restrict__ x, const c128 *__restrict__ y, long long n, c128 *partial) {
    __shared__ double red[8][2];
    c128 acc = c_make(0.0, 0.0);
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += (long long)gridDim.x * blockDim.x) c_fma(acc, x[f], y[f]);
    acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[warp][0] = acc.x; red[warp][1] = acc.y; }
    __syncthreads();
    if (threadIdx.x == 0) {
        c128 s = c_make(0.0, 0.0);
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { s.x += red[w][0]; s.y += red[w][1]; }
        partial[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(32) k_dot_finish(const c128 *partial, int nblk, c128 *out) {
    c128 s = c_make(0.0, 0.0);
    for (int b = threadIdx.x; b < nblk; b += 32) { s.x += partial[b].x; s.y += partial[b].y; }
    s.x = warp_sum(s.x); s.y = warp_sum(s.y);
    if (threadIdx.x == 0) out[0] = s;
}
cudaError_t mmh_launch_dot(const c128 *x, const c128 *y, long long n, c128 *partial, int nblk, c128 *out, cudaStream_t st) {
    k_dot_partial<<<nblk, 256, 0, st>>>(x, y, n, partial);
    k_dot_finish<<<1, 32, 0, st>>>(partial, nblk, out);
    return cudaGetLastError();
}

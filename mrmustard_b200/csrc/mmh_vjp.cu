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

template <int DT>
__global__ void __launch_bounds__(256) k_vjp_partial(VjpParams p) {
    constexpr int NACC = VjpAcc<DT>::NACC;
    const LatticeDesc &d = p.d;
    const long long N = d.N;
    const long long lat = blockIdx.x;
    const c128 *G = p.G + lat * N;
    const c128 *g = p.g + lat * N;
    const double *__restrict__ sq = p.sq;

    long long st[DT];
    int sh[DT];
#pragma unroll
    for (int i = 0; i < DT; i++) { st[i] = d.strides[i]; sh[i] = d.shape[i]; }

    c128 acc[NACC];
#pragma unroll
    for (int e = 0; e < NACC; e++) acc[e] = c_make(0.0, 0.0);

    for (long long f = (long long)blockIdx.y * blockDim.x + threadIdx.x; f < N;
         f += (long long)gridDim.y * blockDim.x) {
        int k[DT];
        long long rem = f;
#pragma unroll
        for (int i = 0; i < DT; i++) {
            k[i] = (int)(rem / st[i]);
            rem -= (long long)k[i] * st[i];
        }
        (void)sh;
        const c128 gk = g[f];
        c_fma(acc[NACC - 1], G[f], gk);  // dLdc numerator (gradients.py:80)
        int e = DT;
#pragma unroll
        for (int i = 0; i < DT; i++) {
            if (k[i] >= 1) {
                const long long pivot = f - st[i];
                const double wi = sq[k[i]];
                c_fma(acc[i], c_scale(G[pivot], wi), gk);                                   // :68
                if (k[i] > 1) c_fma(acc[e], c_scale(G[pivot - st[i]], 0.5 * wi * sq[k[i] - 1]), gk);  // :69-73
#pragma unroll
                for (int j = i + 1; j < DT; j++)
                    if (k[j] >= 1) c_fma(acc[e + (j - i)], c_scale(G[pivot - st[j]], wi * sq[k[j]]), gk);  // :74-75
            }
            e += DT - i;
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

// numba complex_div_impl (Smith) — only the dLdc = sum / c division uses it
__device__ __forceinline__ c128 c_div_smith(c128 a, c128 b) {
    if (fabs(b.x) >= fabs(b.y)) {
        const double ratio = b.y / b.x, denom = b.x + b.y * ratio;
        return c_make((a.x + a.y * ratio) / denom, (a.y - a.x * ratio) / denom);
    }
    const double ratio = b.x / b.y, denom = b.x * ratio + b.y;
    return c_make((a.x * ratio + a.y) / denom, (a.y * ratio - a.x) / denom);
}

// one thread per (lattice, accumulator entry): fixed-order sum over the per-CTA partials, then the
// symmetrisation (dLdA + dLdA^T)/2 (gradients.py:82) and dLdc = sum / c (:80)
__global__ void k_vjp_finish(VjpParams p) {
    const int D = p.d.D;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.batch * p.nacc) return;
    const long long lat = t / p.nacc;
    const int e = (int)(t - lat * p.nacc);
    const c128 *part = p.partial + lat * p.nblk * (long long)p.nacc + e;
    c128 s = c_make(0.0, 0.0);
    for (int blk = 0; blk < p.nblk; blk++) { const c128 v = part[(long long)blk * p.nacc]; s.x += v.x; s.y += v.y; }
    if (e < D) { p.db[lat * D + e] = s; return; }
    if (e == p.nacc - 1) { p.dc[lat] = c_div_smith(s, p.c[lat]); return; }
    int r = e - D, i = 0;
    while (r >= D - i) { r -= D - i; i++; }
    const int j = i + r;
    c128 *dA = p.dA + lat * D * D;
    if (i == j) dA[i * D + i] = s;
    else { const c128 h = c_make(0.5 * s.x, 0.5 * s.y); dA[i * D + j] = h; dA[j * D + i] = h; }
}

cudaError_t mmh_launch_vjp(const VjpParams &p, int grid_x, int block, cudaStream_t st) {
    dim3 grid((unsigned)p.batch, grid_x, 1);  // x = lattice (may exceed 65535), y = CTA within the lattice
    switch (p.d.D) {
        case 1: k_vjp_partial<1><<<grid, block, 0, st>>>(p); break;
        case 2: k_vjp_partial<2><<<grid, block, 0, st>>>(p); break;
        case 3: k_vjp_partial<3><<<grid, block, 0, st>>>(p); break;
        case 4: k_vjp_partial<4><<<grid, block, 0, st>>>(p); break;
        case 5: k_vjp_partial<5><<<grid, block, 0, st>>>(p); break;
        case 6: k_vjp_partial<6><<<grid, block, 0, st>>>(p); break;
        case 7: k_vjp_partial<7><<<grid, block, 0, st>>>(p); break;
        case 8: k_vjp_partial<8><<<grid, block, 0, st>>>(p); break;
        default: {
            grid.z = p.nacc;
            k_vjp_partial_generic<<<grid, block, 0, st>>>(p);
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const long long total = p.batch * p.nacc;
    k_vjp_finish<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(p);
    return cudaGetLastError();
}

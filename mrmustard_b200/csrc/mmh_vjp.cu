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
        const c128 gk = g[f];
        c_fma(acc[NACC - 1], G[f], gk);  // dLdc numerator (gradients.py:80)
        int e = DT;
#pragma unroll
        for (int i = 0; i < DT; i++) {
            if (k[i] >= 1) {
                const IT pivot = f - st[i];
                const double wi = sq[k[i]];
                c_fma(acc[i], c_scale(G[pivot], wi), gk);                                   // :68
                if (k[i] > 1) c_fma(acc[e], c_scale(G[pivot - st[i]], 0.5 * wi * sq[k[i] - 1]), gk);  // :69-73
#pragma unroll
                for (int j = i + 1; j < DT; j++)
                    if (k[j] >= 1) c_fma(acc[e + (j - i)], c_scale(G[pivot - st[j]], wi * sq[k[j]]), gk);  // :74-75
            }
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

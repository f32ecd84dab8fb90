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
// K2: batched march
// ---------------------------------------------------------------------------------------------------
// smem layout (c128 units): sA[L][D*D] | sb[L][D] | tab[L][tab_len] | buf[2][L][P0]
template <int R, int NPD>
__global__ void __launch_bounds__(R >= 4 ? 256 : 512) k_fwd_batched_march(MarchParams p) {
    extern __shared__ c128 smem[];
    const LatticeDesc &d = p.d;
    const int D = d.D;  // == NPD + 1
    const int L = p.L;
    const int P0 = (int)d.strides[0];
    const int T = blockDim.x;
    const int tid = threadIdx.x;
    c128 *sA = smem;
    c128 *sb = sA + L * D * D;
    c128 *tab = sb + L * D;
    c128 *buf = tab + L * p.tab_len;
    const double *__restrict__ sq = p.sq;
    const double *__restrict__ rsq = p.rsq;
    const long long lat0 = (long long)blockIdx.x * L;
    const int nlat = (int)((p.batch - lat0) < L ? (p.batch - lat0) : L);  // lattices handled by this CTA

    // ---- stage the triples and the coefficient tables tab[l][off_j + k] = A_l[0][j] * sqrt(k) --------------
    for (int t = tid; t < nlat * D * D; t += T) sA[t] = p.A[lat0 * D * D + t];
    for (int t = tid; t < nlat * D; t += T) sb[t] = p.b[lat0 * D + t];
    __syncthreads();
    for (int t = tid; t < nlat * p.tab_len; t += T) {
        const int l = t / p.tab_len, r = t - l * p.tab_len;
        int j = 1;
        while (j + 1 < D && r >= p.tab_off[j + 1]) j++;
        const int k = r - p.tab_off[j];
        tab[t] = c_scale(sA[l * D * D + j], sq[k]);  // row 0 of A
    }

    // ---- lower stages (sub-lattice k_0 = 0) in shared memory, buf[0][l][0..P0) ---------------------------
    c128 *buf0 = buf;
    if (tid < nlat) buf0[tid * P0] = p.c[lat0 + tid];
    __syncthreads();
    {
        // chain stage i = D-1 (one thread per lattice, packed in the first warp(s))
        const int i = D - 1;
        if (tid < nlat) {
            c128 *g = buf0 + tid * P0;
            const c128 *A = sA + tid * D * D, *b = sb + tid * D;
            const int S = d.shape[i];
            c128 p1 = g[0], p2 = c_make(0.0, 0.0);
            for (int s = 1; s < S; s++) {
                c128 v = c_mul(b[i], p1);
                if (s >= 2) v = c_add(v, c_mul(c_scale(A[i * D + i], sq[s - 1]), p2));
                v = c_div_table(v, sq[s], rsq[s]);
                g[s] = v;
                p2 = p1; p1 = v;
            }
        }
        __syncthreads();
        for (int i2 = D - 2; i2 >= 1; i2--) {
            const unsigned P = (unsigned)d.strides[i2];
            const int S = d.shape[i2];
            for (int s = 1; s < S; s++) {
                for (unsigned t = tid; t < (unsigned)nlat * P; t += T) {
                    const unsigned l = t / P, f = t - l * P;
                    c128 *g = buf0 + l * P0;
                    g[(unsigned)s * P + f] = vanilla_point32(d, sA + l * D * D, sb + l * D, g, sq, rsq, i2, s, f);
                }
                __syncthreads();
            }
        }
    }

    // ---- stage 0 march ----------------------------------------------------------------------------------
    const int nslots = nlat * P0;
    int loc[R];            // slot -> index inside a panel buffer (l * P0 + f); < 0: idle slot
    int kj[R][NPD];
    long long gidx[R];
    c128 prev1[R], prev2[R];
    int lst[NPD];
#pragma unroll
    for (int jj = 0; jj < NPD; jj++) lst[jj] = (int)d.strides[1 + jj];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int q = r * T + tid;
        loc[r] = -1;
        if (q < nslots) {
            const int l = q / P0, f = q - l * P0;
            loc[r] = q;
            int rem = f;
#pragma unroll
            for (int jj = 0; jj < NPD; jj++) { kj[r][jj] = rem / lst[jj]; rem -= kj[r][jj] * lst[jj]; }
            gidx[r] = (lat0 + l) * d.N + f;
            prev1[r] = buf0[q];
            prev2[r] = c_make(0.0, 0.0);
            p.G[gidx[r]] = prev1[r];  // panel 0 goes to HBM
        }
    }
    const int S0 = d.shape[0];
    for (int s = 1; s < S0; s++) {
        const c128 *bprev = buf + ((s - 1) & 1) * L * P0;
        c128 *bcur = buf + (s & 1) * L * P0;
        const double sqs = sq[s], rsqs = rsq[s], sqm = sq[s - 1];
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (loc[r] < 0) continue;
            const int l = loc[r] / P0;
            const c128 *A = sA + l * D * D;
            c128 v = c_mul(sb[l * D], prev1[r]);
            if (s >= 2) v = c_add(v, c_mul(c_scale(A[0], sqm), prev2[r]));
            const c128 *tl = tab + l * p.tab_len;
#pragma unroll
            for (int jj = 0; jj < NPD; jj++)
                if (kj[r][jj] > 0) v = c_add(v, c_mul(tl[p.tab_off[1 + jj] + kj[r][jj]], bprev[loc[r] - lst[jj]]));
            v = c_div_table(v, sqs, rsqs);
            gidx[r] += P0;
            p.G[gidx[r]] = v;
            bcur[loc[r]] = v;
            prev2[r] = prev1[r];
            prev1[r] = v;
        }
        __syncthreads();
    }
}

// thread-per-lattice chain for D == 1 (cfg1): G[n] = (b G[n-1] + A sqrt(n-1) G[n-2]) / sqrt(n)
__global__ void __launch_bounds__(128) k_fwd_chain(FwdParams p) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= p.batch) return;
    const c128 A = p.A[l], b = p.b[l];
    c128 *g = p.G + l * p.d.N;
    const int S = p.d.shape[0];
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
static cudaError_t launch_batched_R(const MarchParams &p, int grid, int block, size_t smem, cudaStream_t st) {
#define MMH_CASE(N)                                                                                   \
    case N:                                                                                           \
        if (smem > 48 * 1024)                                                                         \
            cudaFuncSetAttribute(k_fwd_batched_march<R, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        k_fwd_batched_march<R, N><<<grid, block, smem, st>>>(p);                                      \
        break;
    switch (p.d.D - 1) {
        MMH_CASE(1) MMH_CASE(2) MMH_CASE(3) MMH_CASE(4) MMH_CASE(5) MMH_CASE(6) MMH_CASE(7)
        default: return cudaErrorInvalidValue;
    }
#undef MMH_CASE
    return cudaGetLastError();
}

cudaError_t mmh_launch_batched_march(const MarchParams &p, int R, int grid, int block, size_t smem, cudaStream_t st) {
    switch (R) {
        case 1: return launch_batched_R<1>(p, grid, block, smem, st);
        case 2: return launch_batched_R<2>(p, grid, block, smem, st);
        case 4: return launch_batched_R<4>(p, grid, block, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t mmh_launch_chain(const FwdParams &p, cudaStream_t st) {
    const int block = 128;
    const long long grid = (p.batch + block - 1) / block;
    k_fwd_chain<<<(unsigned)grid, block, 0, st>>>(p);
    return cudaGetLastError();
}

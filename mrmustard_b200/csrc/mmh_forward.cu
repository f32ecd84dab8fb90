// mmh_forward.cu — forward Gaussian-to-Fock kernels (generic dimension), sm_100a.
//
// Schedule ("panel march").  With the reference's pivot rule (first non-zero index,
// pivots.py:21-34 / core.py:90-94) a point k whose first non-zero index is i reads only
//     G[k - e_i], G[k - 2 e_i], G[k - e_i - e_j]  (j > i),
// all of which have the same zero prefix k_0..k_{i-1} = 0 and k_i smaller by one or two.  Hence, for a
// fixed "stage" i and "step" s = k_i >= 1, the panel  { k : k_<i = 0, k_i = s, k_>i free }  of
// strides[i] points is embarrassingly parallel given the panels s-1 and s-2 of the same stage (panel 0
// of stage i is the whole sub-lattice of the stages > i).  The lattice is filled by stages D-1 .. 0 and
// steps 1 .. shape[i]-1: sum_i (shape[i]-1) dependent steps instead of prod(shape) dependent points.
//
// The per-point arithmetic is the reference's (core.py:97-104), see mmh_common.cuh.
#include "mmh_params.cuh"
#include "mmh_points.cuh"


// Fill stages [stage_lo, stage_hi] (descending) of one lattice with the threads of one CTA.
__device__ void cta_fill_vanilla(const LatticeDesc &d, const c128 *sA, const c128 *sb, c128 *G,
                                 const double *sq, const double *rsq, int stage_hi, int stage_lo) {
    const bool small = d.N < 0x7fffffffLL;
    for (int i = stage_hi; i >= stage_lo; i--) {
        const long long P = d.strides[i];
        const int S = d.shape[i];
        for (int s = 1; s < S; s++) {
            if (small) {
                const unsigned base = (unsigned)s * (unsigned)P;
                for (unsigned f = threadIdx.x; f < (unsigned)P; f += blockDim.x)
                    G[base + f] = vanilla_point32(d, sA, sb, G, sq, rsq, i, s, f);
            } else {
                const long long base = (long long)s * P;
                for (long long f = threadIdx.x; f < P; f += blockDim.x)
                    G[base + f] = vanilla_point(d, sA, sb, G, sq, rsq, i, s, f);
            }
            __syncthreads();
        }
    }
}

// Level-wavefront fill for the stable rule: level n = |k|; threads enumerate the prefix (k_0..k_{D-2})
// and derive k_{D-1} = n - sum(prefix).
__device__ void cta_fill_stable(const LatticeDesc &d, const c128 *sA, const c128 *sb, c128 *G,
                                const double *sq, const double *rsq) {
    const int D = d.D;
    const int last = d.shape[D - 1];
    const long long Q = d.N / last;  // number of prefixes
    int maxlevel = 0;
    for (int i = 0; i < D; i++) maxlevel += d.shape[i] - 1;
    int k[MMH_MAX_DIM];
    for (int n = 1; n <= maxlevel; n++) {
        for (long long q = threadIdx.x; q < Q; q += blockDim.x) {
            long long rem = q * last;
            int sum = 0;
            for (int j = 0; j < D - 1; j++) {
                k[j] = (int)(rem / d.strides[j]);
                rem -= (long long)k[j] * d.strides[j];
                sum += k[j];
            }
            const int kl = n - sum;
            if (kl < 0 || kl >= last) continue;
            k[D - 1] = kl;
            const long long flat = q * last + kl;
            G[flat] = stable_point(d, sA, sb, G, sq, rsq, k, flat);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void load_Ab(const FwdParams &p, long long lattice, c128 *sA, c128 *sb) {
    const int D = p.d.D;
    const c128 *A = p.A + lattice * D * D;
    const c128 *b = p.b + lattice * D;
    for (int t = threadIdx.x; t < D * D; t += blockDim.x) sA[t] = A[t];
    for (int t = threadIdx.x; t < D; t += blockDim.x) sb[t] = b[t];
}

// K2-generic: one CTA per triple (grid-stride over the batch); the lattice lives in global memory and
// is re-read through L1/L2 (the CTA is the only reader and writer of its lattice).
template <bool STABLE>
__global__ void __launch_bounds__(256) k_fwd_cta(FwdParams p) {
    extern __shared__ c128 smem[];
    c128 *sA = smem;
    c128 *sb = smem + p.d.D * p.d.D;
    for (long long l = blockIdx.x; l < p.batch; l += gridDim.x) {
        __syncthreads();
        load_Ab(p, l, sA, sb);
        c128 *G = p.G + l * p.d.N;
        if (threadIdx.x == 0) G[0] = p.c[l];
        __syncthreads();
        if (STABLE) cta_fill_stable(p.d, sA, sb, G, p.sq, p.rsq);
        else cta_fill_vanilla(p.d, sA, sb, G, p.sq, p.rsq, p.d.D - 1, 0);
    }
}

// monotone-counter grid barrier; every CTA of a cooperative launch calls it the same number of times
__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned &epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch++;
        red_release_add_u32(counter, 1u);
        const unsigned target = epoch * gridDim.x;
        while (ld_acquire_u32(counter) < target) { }
    }
    __syncthreads();
}

// K1-coop: one lattice, all SMs.  CTA 0 fills the small trailing stages alone (CTA barriers only), then
// every remaining (stage, step) panel is spread over the grid with one grid barrier per step.
__global__ void __launch_bounds__(256) k_fwd_coop(FwdParams p) {
    extern __shared__ c128 smem[];
    c128 *sA = smem;
    c128 *sb = smem + p.d.D * p.d.D;
    const LatticeDesc &d = p.d;
    load_Ab(p, 0, sA, sb);
    c128 *G = p.G;
    unsigned epoch = 0;
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) G[0] = p.c[0];
        __syncthreads();
        cta_fill_vanilla(d, sA, sb, G, p.sq, p.rsq, d.D - 1, p.small_stage_lo);
        __threadfence();
    } else {
        __syncthreads();
    }
    const bool small = d.N < 0x7fffffffLL;
    for (int i = p.small_stage_lo - 1; i >= 0; i--) {
        const long long P = d.strides[i];
        const int S = d.shape[i];
        for (int s = 1; s < S; s++) {
            grid_barrier(p.barrier, epoch);
            const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            const long long step = (long long)gridDim.x * blockDim.x;
            if (small) {
                const unsigned base = (unsigned)s * (unsigned)P;
                for (unsigned f = (unsigned)t0; f < (unsigned)P; f += (unsigned)step)
                    G[base + f] = vanilla_point32(d, sA, sb, G, p.sq, p.rsq, i, s, f);
            } else {
                const long long base = (long long)s * P;
                for (long long f = t0; f < P; f += step)
                    G[base + f] = vanilla_point(d, sA, sb, G, p.sq, p.rsq, i, s, f);
            }
        }
    }
}

// K1-stable-coop: level wavefront over the whole grid with one grid barrier per level.
__global__ void __launch_bounds__(256) k_stable_coop(FwdParams p) {
    extern __shared__ c128 smem[];
    c128 *sA = smem;
    c128 *sb = smem + p.d.D * p.d.D;
    const LatticeDesc &d = p.d;
    load_Ab(p, 0, sA, sb);
    c128 *G = p.G;
    if (blockIdx.x == 0 && threadIdx.x == 0) G[0] = p.c[0];
    const int D = d.D;
    const int last = d.shape[D - 1];
    const long long Q = d.N / last;
    int maxlevel = 0;
    for (int i = 0; i < D; i++) maxlevel += d.shape[i] - 1;
    int k[MMH_MAX_DIM];
    unsigned epoch = 0;
    // A thread owns the same prefixes (k_0..k_{D-2}) on every level, so when it owns at most kOwn of them they are decoded
    // once (the 64-bit divisions of the decode cost more than the update itself) and a level only derives k_{D-1} = n - sum.
    constexpr int kOwn = 4;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    if (D <= 8 && Q <= kOwn * nthreads) {
        int kk[kOwn][8], ksum[kOwn];
        long long kbase[kOwn];
        int nown = 0;
        for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < Q && nown < kOwn; q += nthreads, nown++) {
            long long rem = q * last;
            int sum = 0;
            for (int j = 0; j < D - 1; j++) {
                kk[nown][j] = (int)(rem / d.strides[j]);
                rem -= (long long)kk[nown][j] * d.strides[j];
                sum += kk[nown][j];
            }
            ksum[nown] = sum;
            kbase[nown] = q * last;
        }
        for (int n = 1; n <= maxlevel; n++) {
            grid_barrier(p.barrier, epoch);
            for (int o = 0; o < nown; o++) {
                const int kl = n - ksum[o];
                if (kl < 0 || kl >= last) continue;
                for (int j = 0; j < D - 1; j++) k[j] = kk[o][j];
                k[D - 1] = kl;
                const long long flat = kbase[o] + kl;
                G[flat] = stable_point(d, sA, sb, G, p.sq, p.rsq, k, flat);
            }
        }
        return;
    }
    for (int n = 1; n <= maxlevel; n++) {
        grid_barrier(p.barrier, epoch);
        for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < Q;
             q += (long long)gridDim.x * blockDim.x) {
            long long rem = q * last;
            int sum = 0;
            for (int j = 0; j < D - 1; j++) {
                k[j] = (int)(rem / d.strides[j]);
                rem -= (long long)k[j] * d.strides[j];
                sum += k[j];
            }
            const int kl = n - sum;
            if (kl < 0 || kl >= last) continue;
            k[D - 1] = kl;
            const long long flat = q * last + kl;
            G[flat] = stable_point(d, sA, sb, G, p.sq, p.rsq, k, flat);
        }
    }
}

// One (stage, step) panel as a plain grid-stride kernel: used when the panel is too large for the tiled
// march (e.g. the 35.8 M-point panels of an (12,)*8 lattice), where a launch per step costs nothing.
__global__ void __launch_bounds__(256) k_panel_step(FwdParams p, int stage, int s, long long f_lo, long long f_hi) {
    extern __shared__ c128 smem[];
    c128 *sA = smem;
    c128 *sb = smem + p.d.D * p.d.D;
    load_Ab(p, 0, sA, sb);
    __syncthreads();
    const LatticeDesc &d = p.d;
    const long long P = d.strides[stage];
    c128 *G = p.G;
    const long long base = (long long)s * P;
    (void)P;
    if (d.N < 0x7fffffffLL) {   // 32-bit index arithmetic
        for (long long f = f_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; f < f_hi; f += (long long)gridDim.x * blockDim.x)
            G[base + f] = vanilla_point32(d, sA, sb, G, p.sq, p.rsq, stage, s, (unsigned)f);
    } else {
        for (long long f = f_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; f < f_hi; f += (long long)gridDim.x * blockDim.x)
            G[base + f] = vanilla_point(d, sA, sb, G, p.sq, p.rsq, stage, s, f);
    }
}

// ---- binomial: level wavefront in one CTA with deterministic per-level norm and early stop ---------
// (binomial.py:58-71, steps.py:208-235).  The vanilla update is used for every point of a level; the
// reference additionally adds exact zeros for j < i (steps.py:63-64), which does not change the value.

__global__ void __launch_bounds__(1024) k_binomial_cta(BinomParams p) {
    extern __shared__ c128 smem[];
    const LatticeDesc &d = p.d;
    const int D = d.D;
    c128 *sA = smem;
    c128 *sb = smem + D * D;
    double *red = (double *)(sb + D);  // blockDim.x / 32 partial sums + 1 broadcast slot
    for (int t = threadIdx.x; t < D * D; t += blockDim.x) sA[t] = p.A[t];
    for (int t = threadIdx.x; t < D; t += blockDim.x) sb[t] = p.b[t];
    c128 *G = p.G;
    const c128 c0 = p.c[0];
    if (threadIdx.x == 0) G[0] = c0;
    const double a0 = hypot(c0.x, c0.y);
    double norm = __dmul_rn(a0, a0);  // np.abs(c) ** 2 (binomial.py:55)
    __syncthreads();
    const int last = d.shape[D - 1];
    const long long Q = d.N / last;
    int maxlevel = 0;
    for (int i = 0; i < D; i++) maxlevel += d.shape[i] - 1;
    for (long long n = 1; n < p.global_cutoff; n++) {
        double part = 0.0;
        if (n <= maxlevel) {
            for (long long q = threadIdx.x; q < Q; q += blockDim.x) {
                long long rem = q * last;
                int sum = 0, piv = -1, spiv = 0;
                long long fpanel = 0;  // offset of the point inside its (stage, step) panel
                for (int j = 0; j < D - 1; j++) {
                    const int kj = (int)(rem / d.strides[j]);
                    rem -= (long long)kj * d.strides[j];
                    sum += kj;
                    if (piv < 0) { if (kj > 0) { piv = j; spiv = kj; } }
                    else fpanel += (long long)kj * d.strides[j];
                }
                const long long kl = n - sum;
                if (kl < 0 || kl >= last) continue;
                if (piv < 0) { piv = D - 1; spiv = (int)kl; }
                else fpanel += kl;
                const c128 v = vanilla_point(d, sA, sb, G, p.sq, p.rsq, piv, spiv, fpanel);
                G[q * last + kl] = v;
                const double a = hypot(v.x, v.y);
                part += a * a;
            }
        }
        // deterministic block reduction (fixed tree)
        for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < (int)(blockDim.x + 31) / 32; w++) s += red[w];
            red[32] = s;
        }
        __syncthreads();
        norm += red[32];
        __syncthreads();
        if (norm > p.max_l2) break;
    }
    if (threadIdx.x == 0) *p.norm_out = norm;
}

// ---- host-side launchers (called from mmh_api.cu) -------------------------------------------------
cudaError_t mmh_launch_fwd_cta(const FwdParams &p, bool stable, int grid, int block, size_t smem,
                               cudaStream_t st) {
    if (stable) k_fwd_cta<true><<<grid, block, smem, st>>>(p);
    else k_fwd_cta<false><<<grid, block, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t mmh_coop_max_blocks(bool stable, int block, size_t smem, int *per_sm) {
    if (stable) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, k_stable_coop, block, smem);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, k_fwd_coop, block, smem);
}

cudaError_t mmh_launch_fwd_coop(const FwdParams &p, bool stable, int grid, int block, size_t smem,
                                cudaStream_t st) {
    void *args[] = { (void *)&p };
    if (stable) return cudaLaunchCooperativeKernel((void *)k_stable_coop, dim3(grid), dim3(block), args, smem, st);
    return cudaLaunchCooperativeKernel((void *)k_fwd_coop, dim3(grid), dim3(block), args, smem, st);
}

// Same panel step with a division-free index walk: each thread advances its multi-index by the mixed-radix digits of
// the walking stride (as k_vjp_partial does).  NPD = number of panel dims = D - 1 - stage (<= 8), N < 2^31.
struct PanelWalk { int dig[8]; };

template <int NPD>
__global__ void __launch_bounds__(256) k_panel_step_walk(FwdParams p, int stage, int s, unsigned f_lo, unsigned f_hi, PanelWalk w) {
    extern __shared__ c128 smem[];
    const LatticeDesc &d = p.d;
    const int D = d.D;
    c128 *sA = smem;
    c128 *sb = smem + D * D;
    load_Ab(p, 0, sA, sb);
    __syncthreads();
    const double *__restrict__ sq = p.sq;
    const double *__restrict__ rsq = p.rsq;
    unsigned st[NPD];
    int sh[NPD], k[NPD];
#pragma unroll
    for (int jj = 0; jj < NPD; jj++) { st[jj] = (unsigned)d.strides[stage + 1 + jj]; sh[jj] = d.shape[stage + 1 + jj]; }
    const unsigned P = (unsigned)d.strides[stage];
    c128 *G = p.G;
    const c128 *Gp = G + (size_t)(s - 1) * P;        // panel s-1
    const c128 *Gpp = G + (size_t)(s >= 2 ? s - 2 : 0) * P;
    c128 *Gc = G + (size_t)s * P;
    const c128 bi = sb[stage], aii = c_scale(sA[stage * D + stage], sq[s - 1]);
    const double sqs = sq[s], rsqs = rsq[s];
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned f = f_lo + blockIdx.x * blockDim.x + threadIdx.x;
    {
        unsigned rem = f < f_hi ? f : 0;
#pragma unroll
        for (int jj = 0; jj < NPD; jj++) { k[jj] = (int)(rem / st[jj]); rem -= (unsigned)k[jj] * st[jj]; }
    }
    // Launched with programmatic stream serialization (mid-size panels are one launch per step, launch-latency bound): everything
    // above -- A, b (caller inputs, never written), the tables, the index decode -- overlaps the previous step's kernel; the lattice
    // is only touched behind this wait, which returns once that kernel has completed and flushed.  Without the attribute: no-ops.
    pdl_trigger();
    pdl_wait();
    for (; f < f_hi; f += stride) {
        // all loads of the point first, unconditionally (an absent neighbour re-reads the pivot and is not used): with a branch
        // around each of them they issue one L2/DRAM latency after the other, and this kernel is latency bound
        // (ncu: long-scoreboard stall 9.8 cycles per issue, L2 24 %, FP64 29 %)
        // cache policy: panel s-2 and the output are touched once (streaming, evict-first); panel s-1 is re-read up to NPD times
        // at distances up to strides[stage+1] behind the sweep -- with the streams kept out of its way L2 retains that window
        // ((12,)^8: 48 MB), so every amplitude of panel s-1 is fetched from HBM once
        const c128 pv = Gp[f];
        const c128 pp = __ldcs(Gpp + (s >= 2 ? f : 0));
        c128 nb[NPD];
#pragma unroll
        for (int jj = 0; jj < NPD; jj++) nb[jj] = Gp[k[jj] > 0 ? f - st[jj] : f];
        c128 val = c_mul(bi, pv);
        if (s >= 2) val = c_add(val, c_mul(aii, pp));
#pragma unroll
        for (int jj = 0; jj < NPD; jj++)
            if (k[jj] > 0) val = c_add(val, c_mul(c_scale(sA[stage * D + stage + 1 + jj], sq[k[jj]]), nb[jj]));
        __stcs(Gc + f, c_div_table(val, sqs, rsqs));
        int carry = 0;
#pragma unroll
        for (int jj = NPD - 1; jj >= 0; jj--) {
            int t = k[jj] + w.dig[jj] + carry;
            carry = 0;
            if (jj > 0) { while (t >= sh[jj]) { t -= sh[jj]; carry++; } }
            k[jj] = t;
        }
    }
}

cudaError_t mmh_launch_panel_step(const FwdParams &p, int stage, int s, long long f_lo, long long f_hi, int grid,
                                  size_t smem, cudaStream_t st, bool pdl) {
    const int npd = p.d.D - 1 - stage;
    if (p.d.N < 0x7fffffffLL && npd >= 1 && npd <= 8) {
        // a grid-stride sweep wants exactly one resident wave: 8 CTAs per SM were asked for, 6 fit (40 registers), and the
        // 1.33 waves ran as two
        {
            static int per_sm[9] = { 0 }, sms = 0;
            if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
            if (!per_sm[npd]) {
                int n = 0;
                switch (npd) {
#define MMH_OCC(N) case N: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_panel_step_walk<N>, 256, smem); break;
                    MMH_OCC(1) MMH_OCC(2) MMH_OCC(3) MMH_OCC(4) MMH_OCC(5) MMH_OCC(6) MMH_OCC(7) MMH_OCC(8)
#undef MMH_OCC
                }
                per_sm[npd] = n > 0 ? n : 1;
            }
            const long long need = (f_hi - f_lo + 255) / 256, cap = (long long)per_sm[npd] * sms;
            grid = (int)(need < cap ? need : cap);
            if (grid < 1) grid = 1;
        }
        PanelWalk w;
        long long sd = (long long)grid * 256;
        for (int jj = npd - 1; jj >= 0; jj--) {
            const int shj = p.d.shape[stage + 1 + jj];
            if (jj > 0) { w.dig[jj] = (int)(sd % shj); sd /= shj; }
            else w.dig[0] = (int)(sd < (1 << 30) ? sd : (1 << 30));
        }
        switch (npd) {
#define MMH_CASE(N) case N: return mmh_launch_ex(k_panel_step_walk<N>, dim3(grid), dim3(256), smem, st, pdl && !mmh_getenv("MMH_NO_PDL"), p, stage, s, (unsigned)f_lo, (unsigned)f_hi, w);
            MMH_CASE(1) MMH_CASE(2) MMH_CASE(3) MMH_CASE(4) MMH_CASE(5) MMH_CASE(6) MMH_CASE(7) MMH_CASE(8)
#undef MMH_CASE
        }
        return cudaGetLastError();
    }
    k_panel_step<<<grid, 256, smem, st>>>(p, stage, s, f_lo, f_hi);
    return cudaGetLastError();
}

cudaError_t mmh_launch_binomial(const BinomParams &p, int block, size_t smem, cudaStream_t st) {
    k_binomial_cta<<<1, block, smem, st>>>(p);
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------
// Derived-variable contraction after the lattice (CircuitComponent.fock_array, lab/circuit_components.py:516-530;
// PolyExpAnsatz.decompose_ansatz, physics/ansatz/polyexp_ansatz.py:447-462):
//     out[l, i] = sum_d G[l, i, d] * cpoly[l, d]          i < ncore, d < nd (the trailing "derived" axes of the lattice)
// The lattice stays in device memory; only `out` (nd times smaller) travels to the host.
// nd <= 32: one thread per output, terms added in index order; nd > 32: one warp per output, lanes stride over d and a
// fixed-order shuffle tree adds the partial sums (deterministic).  Separate roundings (no FMA): the sum is what
// np.einsum's scalar loop computes up to its (unspecified) order.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_contract_last(const c128 *__restrict__ G, const c128 *__restrict__ cp, c128 *__restrict__ out,
                                                       long long nout, long long ncore, int nd) {
    if (nd <= 32) {
        const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (o >= nout) return;
        const c128 *g = G + o * nd, *c = cp + (o / ncore) * nd;
        c128 acc = c_mul(g[0], c[0]);
        for (int d = 1; d < nd; d++) acc = c_add(acc, c_mul(g[d], c[d]));
        out[o] = acc;
    } else {
        const long long o = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const int lane = threadIdx.x & 31;
        if (o >= nout) return;
        const c128 *g = G + o * nd, *c = cp + (o / ncore) * nd;
        c128 acc = c_make(0.0, 0.0);
        for (int d = lane; d < nd; d += 32) acc = c_add(acc, c_mul(g[d], c[d]));
        for (int w = 16; w > 0; w >>= 1) {
            acc.x = __dadd_rn(acc.x, __shfl_down_sync(0xffffffffu, acc.x, w));
            acc.y = __dadd_rn(acc.y, __shfl_down_sync(0xffffffffu, acc.y, w));
        }
        if (lane == 0) out[o] = acc;
    }
}

cudaError_t mmh_launch_contract_last(const c128 *G, const c128 *cp, c128 *out, long long nout, long long ncore, int nd,
                                     cudaStream_t st) {
    const long long threads = nd <= 32 ? nout : nout * 32;
    const long long grid = (threads + 255) / 256;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    k_contract_last<<<(unsigned)grid, 256, 0, st>>>(G, cp, out, nout, ncore, nd);
    return cudaGetLastError();
}

__global__ void k_fill_ones(c128 *p, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = c_make(1.0, 0.0);
}
cudaError_t mmh_launch_fill_ones(c128 *p, long long n, cudaStream_t st) {
    k_fill_ones<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n);
    return cudaGetLastError();
}

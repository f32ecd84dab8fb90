// mmh_stable_boxes.cu — stable_numba (vanilla/core.py:127-213) for ONE large lattice of 2..4 indices as a wavefront of BOXES, sm_100a.
//
// The stable rule averages the update over every pivot i with k_i > 0, so an amplitude reads k - e_i and k - e_i - e_j for ALL
// i, j: its dependencies lie in every lower direction, also inside its own panel.  k_stable_coop therefore sweeps the levels
// |k| = n of the whole lattice with one grid barrier per level and fetches the 14 neighbours of a point from L2: (50,)^4 has 197
// levels of ~8 us (barrier + an L2 round trip with one point per thread) = 1.6 ms, 13x the vanilla path.
//
// Here the lattice is cut into boxes of E^D amplitudes (E = 6 at D = 4).  A box only depends on boxes that are lower in some
// direction, so persistent CTAs take boxes from a ticket counter in a topological order (by box level, precomputed on the host) and
// wait -- on a ready flag per box, acquire / release -- only for the D face neighbours below (those waited for theirs: transitively
// every lower box is complete).  A CTA loads the two-deep lower halo of its box from the lattice (L2) into shared memory, sweeps
// the 3 D - 2 .. local levels of the box with CTA barriers and shared-memory neighbours, writes the amplitudes to the lattice and
// raises the box's flag.  No grid barrier, one L2 round trip per box instead of one per level.  The arithmetic per amplitude is
// stable_point_batched's (mmh_points.cuh), operand for operand: bit-identical to k_stable_coop and to the oracle.
#include <cstring>

#include "mmh_params.cuh"
#include "mmh_points.cuh"

namespace {

__device__ __forceinline__ unsigned long long sb_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int sb_ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sb_st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// one amplitude from the extended box in shared memory.  k: global multi-index (4 padded dims, `pad` leading dims are trivial),
// flat: index in the extended box, es: its strides, tab: (sqrt, 1/sqrt) table in shared memory.  The same operations in the same
// order as stable_point_batched<4>, but written without data-dependent branches: the four pivots' numerators and quotients are
// independent dependency chains (a level of a box is latency bound: ~150 cells on 256 threads), absent terms are computed on
// harmless operands and dropped by selects -- never added, so the result is the reference's bit for bit.
template <int D>
__device__ __forceinline__ c128 stable_point_box(const c128 *sA, const c128 *sb, const c128 *buf, const double2 *tab,
                                                 const int *k, int flat, const int *es) {
    constexpr int pad = 4 - D;
    // interior amplitudes (every k_i >= 2: all pivots, all terms): the same sums straight-line, no selects.  Only for D <= 3
    // ((1000,1000): 3.40 -> 1.93 ms); at D = 4 most boxes of a (50,)^4 lattice touch a boundary and the extra branch costs 8 %.
    if constexpr (D <= 3) {
        bool all = true;
#pragma unroll
        for (int i = pad; i < 4; i++) all &= k[i] >= 2;
        if (all) {
            c128 q[4];
            double2 tk[4];
            double w1[4];
#pragma unroll
            for (int i = pad; i < 4; i++) { tk[i] = tab[k[i]]; w1[i] = tab[k[i] - 1].x; }
#pragma unroll
            for (int i = pad; i < 4; i++) {
                c128 val = c_mul(sb[i - pad], buf[flat - es[i]]);
#pragma unroll
                for (int j = pad; j < 4; j++)
                    val = c_add(val, c_mul(c_scale(sA[(i - pad) * D + (j - pad)], j == i ? w1[i] : tk[j].x), buf[flat - es[i] - es[j]]));
                q[i] = val;
            }
            bool slow = false;
#pragma unroll
            for (int i = pad; i < 4; i++) slow |= div_needs_slow(q[i].x) | div_needs_slow(q[i].y);
            if (!slow) {
#pragma unroll
                for (int i = pad; i < 4; i++) q[i] = c_make(div_fast(q[i].x, tk[i].x, tk[i].y), div_fast(q[i].y, tk[i].x, tk[i].y));
            } else {
#pragma unroll
                for (int i = pad; i < 4; i++) q[i] = c_make(div_by_table(q[i].x, tk[i].x, tk[i].y), div_by_table(q[i].y, tk[i].x, tk[i].y));
            }
            c128 vals = c_add(c_make(0.0, 0.0), q[pad]);
#pragma unroll
            for (int i = pad + 1; i < 4; i++) vals = c_add(vals, q[i]);
            return c_div_count(vals, D);
        }
    }
    c128 p1[4], p2[4][4];   // p1[i] = G[k - e_i], p2[i][j] (i <= j) = G[k - e_i - e_j]; absent ones read the cell itself (finite)
    double2 tk[4], tk1[4];  // (sqrt, 1/sqrt) of k_i and of k_i - 1
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const bool has = i >= pad && k[i] > 0;
        p1[i] = buf[has ? flat - es[i] : flat];
        tk[i] = tab[has ? k[i] : 1];
        tk1[i] = tab[(has && k[i] > 1) ? k[i] - 1 : 0];
#pragma unroll
        for (int j = i; j < 4; j++) {
            const bool need = has && (i == j ? k[i] > 1 : k[j] > 0);   // (j >= i >= pad)
            p2[i][j] = buf[need ? flat - es[i] - es[j] : flat];
        }
    }
    c128 q[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int io = i >= pad ? i - pad : 0;
        c128 val = c_mul(sb[io], p1[i]);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int jo = j >= pad ? j - pad : 0;
            const bool use = j >= pad && (j == i ? k[i] > 1 : k[j] > 0);
            const double w = j == i ? tk1[i].x : tk[j].x;
            const c128 t = c_add(val, c_mul(c_scale(sA[io * D + jo], w), j < i ? p2[j][i] : p2[i][j]));
            val = use ? t : val;
        }
        q[i] = val;
    }
    // the four quotients behind one range test
    bool slow = false;
#pragma unroll
    for (int i = 0; i < 4; i++) slow |= div_needs_slow(q[i].x) | div_needs_slow(q[i].y);
    if (!slow) {
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = c_make(div_fast(q[i].x, tk[i].x, tk[i].y), div_fast(q[i].y, tk[i].x, tk[i].y));
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = c_make(div_by_table(q[i].x, tk[i].x, tk[i].y), div_by_table(q[i].y, tk[i].x, tk[i].y));
    }
    c128 vals = c_make(0.0, 0.0);
    int np = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const bool has = i >= pad && k[i] > 0;
        const c128 t = c_add(vals, q[i]);
        vals = has ? t : vals;
        np += has ? 1 : 0;
    }
    return c_div_count(vals, np);
}

// dynamic shared memory: ext box [X^D] c128 | A [D*D] | b [D] | (sqrt, 1/sqrt) [ntab] double2 | cell_order [E^D] u32 | lvl_start [nlev + 1] | ctl int[4]
template <int D>
__global__ void __launch_bounds__(256) k_stable_boxes(StableBoxParams p) {
    extern __shared__ c128 sbx[];
    const LatticeDesc &d = p.d;
    constexpr int pad = 4 - D;
    const int tid = threadIdx.x, T = blockDim.x;
    const int E = p.E, X = E + 2;
    int xn = 1;
    for (int j = 0; j < D; j++) xn *= X;                  // cells of the extended box
    c128 *buf = sbx;
    c128 *sA = sbx + xn;
    c128 *sb = sA + D * D;
    double2 *tab = (double2 *)(sb + D);
    unsigned *cells = (unsigned *)(tab + p.ntab);
    int *lvls = (int *)(cells + p.ncell);
    int *ctl = lvls + p.nlev + 1;
    for (int n = tid; n < D * D; n += T) sA[n] = p.A[n];
    for (int n = tid; n < D; n += T) sb[n] = p.b[n];
    for (int n = tid; n < p.ntab; n += T) tab[n] = make_double2(p.sq[n], p.rsq[n]);
    for (int n = tid; n < p.ncell; n += T) cells[n] = p.cell_order[n];
    for (int n = tid; n <= p.nlev; n += T) lvls[n] = p.lvl_start[n];
    const int xsh = p.xshift;                            // X = 1 << xshift
    // padded geometry: dim j of 4; the first `pad` dims are trivial (extent 1, no halo)
    int shp[4], gs[4], es[4], nb[4], hh[4];
    {
        int e_acc = 1;
        for (int j = 3; j >= 0; j--) {
            if (j >= pad) { shp[j] = d.shape[j - pad]; gs[j] = (int)d.strides[j - pad]; es[j] = e_acc; e_acc *= X; nb[j] = p.nb[j]; hh[j] = 2; }
            else { shp[j] = 1; gs[j] = 0; es[j] = 0; nb[j] = 1; hh[j] = 0; }
        }
    }
    const unsigned long long t_giveup = sb_timer() + 4000000000ull;
    __syncthreads();

    int nb_done = 0;   // (debug trace: boxes this CTA has processed)
#define SB_STAMP(w) do { if (p.trace && blockIdx.x == 0 && tid == 0 && nb_done < 64) p.trace[nb_done * 8 + (w)] = sb_timer(); } while (0)
    for (;;) {
        SB_STAMP(0);
        if (tid == 0) ctl[0] = atomicAdd(p.ticket, 1);
        __syncthreads();
        const int ticket = ctl[0];
        if (ticket >= p.nbox) return;
        const int box = p.box_order[ticket];
        int t[4], lo[4], e[4];
        {
            int r = box;
            for (int j = 3; j >= 0; j--) { t[j] = r % nb[j]; r /= nb[j]; lo[j] = t[j] * E; e[j] = shp[j] - lo[j] < E ? shp[j] - lo[j] : E; }
        }
        SB_STAMP(1);
        // the face neighbours below must be complete (they waited for theirs)
        if (tid < 4 && t[tid] > 0) {
            int stride = 1;
            for (int j = 3; j > tid; j--) stride *= nb[j];
            const int *f = p.flags + (box - stride);
            unsigned spins = 0;
            while (sb_ld_acquire(f) == 0) {
                if ((++spins & 1023u) == 0u && sb_timer() > t_giveup) { if (p.err) *(volatile int *)p.err = 1; break; }
                __nanosleep(64);
            }
        }
        __syncthreads();
        SB_STAMP(2);
        // lower halo (two deep) from the lattice; everything else of the extended box starts as zero
        for (int c0 = tid; c0 < xn; c0 += 8 * T) {        // batches of eight cells per thread: their loads are in flight together
            c128 hv[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int c = c0 + u * T;
                int r = c;
                bool own = true, inside = c < xn;
                long long go = 0;
#pragma unroll
                for (int j = 3; j >= 0; j--) {
                    const int xsj = j >= pad ? (r & (X - 1)) - 2 : 0;
                    if (j >= pad) r >>= xsh;
                    const int kj = lo[j] + xsj;
                    own &= xsj >= 0;
                    inside &= kj >= 0 && xsj < e[j];
                    go += (long long)kj * gs[j];
                }
                hv[u] = (!own && inside) ? __ldcg(p.G + go) : c_make(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) if (c0 + u * T < xn) buf[c0 + u * T] = hv[u];
        }
        __syncthreads();
        SB_STAMP(3);
        // local level wavefront: the cells of a full box sorted by level come from a host table (cells outside a ragged box skip)
        int mmax = 0;
        for (int j = pad; j < 4; j++) mmax += e[j] - 1;
        for (int m = 0; m <= mmax; m++) {
            const int n1 = lvls[m + 1];
            for (int n = lvls[m] + tid; n < n1; n += T) {
                const unsigned xc = cells[n];
                int x[4], k[4];
                x[0] = xc & 0xff; x[1] = (xc >> 8) & 0xff; x[2] = (xc >> 16) & 0xff; x[3] = (xc >> 24) & 0xff;
                bool in = true;
                int flat = 0;
                long long go = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    in &= x[j] < e[j];
                    k[j] = lo[j] + x[j];
                    flat += (x[j] + hh[j]) * es[j];
                    go += (long long)k[j] * gs[j];
                }
                if (!in) continue;
                c128 v;
                if ((k[0] | k[1] | k[2] | k[3]) == 0) v = p.c[0];                    // the vacuum amplitude
                else v = stable_point_box<D>(sA, sb, buf, tab, k, flat, es);
                buf[flat] = v;
                p.G[go] = v;
            }
            __syncthreads();
        }
        // (Measured and dropped: four adjacent lanes per cell, one pivot each, quotients combined by shuffles -- a level is then one
        //  pivot's dependency chain, but the ~450 mostly integer instructions of a lane-round, several rounds per level at 64 cells
        //  per round, made the box slower: 39-44 us of levels per box against ~25 us.)
        SB_STAMP(4);
        // publish: the box's amplitudes are in the lattice before its flag is raised
        __threadfence();
        __syncthreads();
        if (tid == 0) sb_st_release(p.flags + box, 1);
        SB_STAMP(5);
        if (p.trace && blockIdx.x == 0 && tid == 0 && nb_done < 64) p.trace[nb_done * 8 + 6] = (unsigned long long)box;
        nb_done++;
    }
}

}  // namespace

// extended-box edge per number of indices: 8^4 = 16^3 = 64^2 = 4096 cells (64 KB)
int mmh_stable_boxes_edge(int D) { return D == 4 ? 6 : (D == 3 ? 14 : (D == 2 ? 62 : 0)); }

size_t mmh_stable_boxes_smem(int D, int ntab, int ncell, int nlev) {
    const int X = mmh_stable_boxes_edge(D) + 2;
    size_t xn = 1;
    for (int j = 0; j < D; j++) xn *= (size_t)X;
    return sizeof(c128) * (xn + (size_t)D * D + D) + sizeof(double2) * (size_t)ntab + 4 * (size_t)ncell + 4 * (size_t)(nlev + 1) + 32;
}

template <int D>
static cudaError_t launch_stable_boxes_D(const StableBoxParams &p, int sm_count, cudaStream_t st) {
    const size_t smem = mmh_stable_boxes_smem(D, p.ntab, p.ncell, p.nlev);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(k_stable_boxes<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smem_set = smem;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stable_boxes<D>, 256, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidValue;
    long long grid = (long long)per_sm * sm_count;
    if (grid > p.nbox) grid = p.nbox;
    k_stable_boxes<D><<<(unsigned)grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t mmh_launch_stable_boxes(const StableBoxParams &p, int sm_count, cudaStream_t st) {
    switch (p.d.D) {
        case 2: return launch_stable_boxes_D<2>(p, sm_count, st);
        case 3: return launch_stable_boxes_D<3>(p, sm_count, st);
        case 4: return launch_stable_boxes_D<4>(p, sm_count, st);
        default: return cudaErrorInvalidValue;
    }
}

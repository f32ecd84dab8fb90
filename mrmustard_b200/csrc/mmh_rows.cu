// mmh_rows.cu — K1r: tiled march of ONE lattice with row-owning lanes and warp-specialised data movement, sm_100a.
//
// Same decomposition as mmh_tiled.cu (the panel of stage i is cut into a grid of boxes, CTA t owns box t for the whole march
// s = 1 .. shape[i]-1, halo faces travel through the sentinel-validated exchange buffer X), but a different division of labour
// inside the CTA.  k_march_tiled2 gives every thread R = 2 scattered cells and lets the 16 compute warps also store the lattice and
// export the faces: 255 warp-instructions per warp-step for 100 FP64 ones, every warp in the same phase of the step.  Here:
//   * COMPUTE warps: a lane owns R (2..6, R = 2 by default) CONSECUTIVE cells of a row along the last panel dim.  The neighbour
//     k - e_i - e_last of a cell is the lane's own previous cell (a register; the first cell reads the cell before it from shared
//     memory), the neighbours in the other panel dims are the same cells of the row one lower in that dim (unit-stride,
//     conflict-free LDS.128: cell r of chunk ch sits at r * C + ch of its row).  Per step a compute lane loads its neighbours, runs
//     its 2 R independent FP64 chains, pushes the cells it owns on a high face of the box to the upper neighbour's exchange rows
//     (strong stores straight after the quotient: the neighbours' critical path), stores its R cells to shared memory and arrives
//     on an mbarrier -- no lattice store, no address bookkeeping per cell.
//   * SERVICE warps (four, one per panel buffer, panels s = w mod 4): poll the exchange buffer for the low halo faces of panel s,
//     put them into the halo rows of the buffer and arrive on its halo mbarrier; wait for "panel s complete in shared memory";
//     drain the box to the lattice in cell order through a per-cell table (every STG.128 covers whole row segments, streaming
//     stores unless a later stage polls the lattice); hand the buffer back (drained mbarrier); write the sentinel back into the
//     consumed exchange rows (self-cleaning).  All of it overlaps the compute warps' next steps.
//   Hand-offs: three mbarrier families per buffer (own cells stored: count = compute threads; halo imported; drained) and a named
//   barrier "panel read" (compute arrive, service sync) before a buffer's halo cells are overwritten four panels later.
// Why no cp.async.bulk (TMA) for the drain: UBLKCP is a warp-uniform instruction; a box row is 80-160 contiguous bytes, so a box
// needs ~100 copies per step, each costing more issue slots (uniform address arithmetic or a per-lane R2UR loop, measured in the
// first version of this file: step 1.25 us) than the 2.5 instructions per row of the coalesced LDS.128 + STG.128 drain.
// Arithmetic per amplitude is the one of k_march_tiled2 / the oracle (vanilla/core.py:97-104), in the same order: bit-identical.
//
// Shared memory: NB = 4 rotating panel buffers.  A buffer is a grid of (e0+1) x (e1+1) rows (row (0, *) and (*, 0) are the low
// halo faces in panel dims 0 and 1) of RS cells, followed by one cell per row for the low halo face of panel dim 2 (the last cell
// of the same row in the lower box) and a zero row for inactive lanes.  Absent neighbours (k_j = 0) read zero-initialised cells
// with coefficient 0.
#include <cstring>

#include "mmh_params.cuh"

#define MMH_RW_NB 4
#define MMH_RW_NSV 4            // service warps (one per panel buffer): halo import, face export, drain to the lattice
#define MMH_RW_SENTINEL 0xFFFFFFFFFFFFFFFFull
#define MMH_RW_BAR_FREE 6       // + k: panel buffer k has been read by every compute warp
#ifndef MMH_RW_HANDOFF_BAR
#define MMH_RW_HANDOFF_BAR 0    // 1: debug build for compute-sanitizer racecheck -- the panel hand-offs (own cells stored + halo imported,
                                //    panel drained) go through named barriers, which the tool models, instead of mbarrier arrive / wait
#endif
#ifndef MMH_RW_DATAFLOW
#define MMH_RW_DATAFLOW 0       // 1: experiment -- a compute warp of step s+1 waits only for the warps whose rows it reads (per-warp progress
                                //    words) instead of for the whole CTA.  Measured SLOWER (cfg2 132 vs 121 us: the warps drift apart only until
                                //    the write-after-read limit of the four buffers, and the polls take issue slots), kept as a switch
#endif
#define MMH_RW_NWMAX 16         // compute warps per CTA at most (512 compute threads)
#define MMH_RW_BAR_FULL 1       // + k (debug hand-off only): panel buffer k is complete (compute threads + its service warp)
#define MMH_RW_BAR_DRAINED 10   // + k (debug hand-off only): panel buffer k has been drained

namespace {

__device__ __forceinline__ unsigned long long rw_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void rw_ldg_relaxed(const void *p, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void rw_stg_relaxed_u64(void *p, unsigned long long a, unsigned long long b) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void rw_stg_relaxed(c128 *p, c128 v) {
    asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ c128 rw_poll(const c128 *p, unsigned long long t_giveup, int *err) {
    unsigned long long a, b;
    unsigned spins = 0;
    rw_ldg_relaxed(p, a, b);
    while ((a == MMH_RW_SENTINEL || b == MMH_RW_SENTINEL) && ((++spins & 255u) != 0u || rw_timer() < t_giveup)) {
        __nanosleep(100);
        rw_ldg_relaxed(p, a, b);
    }
    if ((a == MMH_RW_SENTINEL || b == MMH_RW_SENTINEL) && err) *(volatile int *)err = 1;
    return make_double2(__longlong_as_double((long long)a), __longlong_as_double((long long)b));
}
__device__ __forceinline__ c128 rw_lds(unsigned addr) {
    c128 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void rw_sts(unsigned addr, c128 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void rw_mbar_init(unsigned addr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void rw_mbar_arrive(unsigned addr) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void rw_mbar_wait(unsigned addr, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void rw_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void rw_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
template <int R>
__device__ __forceinline__ void rw_div_all(c128 (&v)[R], double sqs, double rsqs) {
    bool slow = false;
#pragma unroll
    for (int r = 0; r < R; r++) slow |= div_needs_slow(v[r].x) | div_needs_slow(v[r].y);
    if (!slow) {
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = c_make(div_fast(v[r].x, sqs, rsqs), div_fast(v[r].y, sqs, rsqs));
    } else {
#pragma unroll
        for (int r = 0; r < R; r++)
            v[r] = c_make(div_needs_slow(v[r].x) ? div_rare(v[r].x, sqs, rsqs) : div_fast(v[r].x, sqs, rsqs),
                          div_needs_slow(v[r].y) ? div_rare(v[r].y, sqs, rsqs) : div_fast(v[r].y, sqs, rsqs));
    }
}

// smem layout (bytes): buf[NB][ls_max] c128 | sqtab[S] double2 | (b_i, A_ii, A_i,last) c128[4] | mbar[3 NB] u64 | prog[NWMAX] i32 |
//                      celltab[cells_max] uint2 | hdst[hc_max] u32
// NPD = panel dims of the stage = tiled dims (the kernel needs strides[stage + NPD] == 1); box dims are right-aligned in
// (0, 1, 2): dim 2 is always the row direction, leading dims are trivial (extent 1, no halo) when NPD < 3.
template <int R, int NPD, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_march_rows(TiledParams p) {
    extern __shared__ c128 smem[];
    const LatticeDesc &d = p.d;
    const int D = d.D;
    const int i = p.stage;
    const long long P = d.strides[i];
    const int S = d.shape[i];
    const int tid = threadIdx.x;
    const int tidc = tid - 32 * MMH_RW_NSV;
    const int TC = (int)blockDim.x - 32 * MMH_RW_NSV;
    constexpr int NB = MMH_RW_NB;
    constexpr int OFF = 3 - NPD;
    const int C = p.rows_C, RS = p.rs;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const bool tl = blockIdx.x == 0 && tid == 32 * MMH_RW_NSV;
    if (tl) timeline_stamp(p.timeline, i & 7, 0);

    // ---- box geometry -------------------------------------------------------------------------------------
    int g[3], t[3], lo[3], e[3], h[3], gst[3], shp[3];
#pragma unroll
    for (int m = 0; m < 3; m++) {
        if (m >= OFF) { g[m] = p.g[m - OFF]; shp[m] = d.shape[i + 1 + m - OFF]; gst[m] = (int)d.strides[i + 1 + m - OFF]; }
        else { g[m] = 1; shp[m] = 1; gst[m] = 0; }
    }
    const int tile = blockIdx.x;
    t[2] = tile % g[2];
    t[1] = (tile / g[2]) % g[1];
    t[0] = tile / (g[1] * g[2]);
#pragma unroll
    for (int m = 0; m < 3; m++) {
        lo[m] = (int)(((long long)t[m] * shp[m]) / g[m]);
        e[m] = (int)(((long long)(t[m] + 1) * shp[m]) / g[m]) - lo[m];
        h[m] = lo[m] > 0 ? 1 : 0;
    }
    const int W1 = e[1] + 1;                              // rows of the extended row grid per x0
    const int nrows_ext = (e[0] + 1) * W1;
    const int F0 = h[0] * e[1] * e[2], F1 = h[1] * e[0] * e[2], F2 = h[2] * e[0] * e[1];
    const int HC = F0 + F1 + F2;                          // halo cells of this box, in exchange order: face 0 | face 1 | face 2
    const int nrows = e[0] * e[1];
    // panel buffer (cells): rows [nrows_ext][RS] | halo cells of dim 2 [nrows_ext] | zero row [RS + 1]
    const unsigned h2_base = (unsigned)(nrows_ext * RS) * 16u;
    const unsigned zero_base = h2_base + (unsigned)nrows_ext * 16u;   // cell 0: "cell before", cells 1..: the row

    const unsigned bstride = (unsigned)p.ls_max * 16u;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    double2 *sqtab = (double2 *)(smem + (size_t)NB * p.ls_max);
    c128 *sba = (c128 *)(sqtab + S);
    const unsigned mb_own = (unsigned)__cvta_generic_to_shared(sba + 4);   // all compute threads have stored panel s
    const unsigned mb_halo = mb_own + 8u * NB;                               // the import warp has delivered the halo of panel s
    const unsigned mb_drained = mb_halo + 8u * NB;                           // the store warp has drained panel s
    int *prog = (int *)((double *)(sba + 4) + 3 * NB);                       // [NWMAX]: last panel compute warp w has stored completely
    uint2 *celltab = (uint2 *)(prog + MMH_RW_NWMAX);            // per box cell, row-major: (byte offset in a buffer, offset in a lattice panel)
    unsigned *hdst = (unsigned *)(celltab + p.cells_max);   // byte offset of imported halo cell c inside a panel buffer
    const size_t xtile = (size_t)S * p.hc_max;            // X cells per consumer box

    // ---- tables that do not touch the lattice (overlap the previous kernel under PDL) ------------------------
    if (tid < NB) {
        rw_mbar_init(mb_own + 8u * (unsigned)tid, (unsigned)TC);
        rw_mbar_init(mb_halo + 8u * (unsigned)tid, 1u);
        rw_mbar_init(mb_drained + 8u * (unsigned)tid, 1u);
    }
    if (tid < MMH_RW_NWMAX) prog[tid] = 0;
    for (int s_ = tid; s_ < S; s_ += blockDim.x) sqtab[s_] = make_double2(p.sq[s_], p.rsq[s_]);
    for (int c = tid; c < NB * p.ls_max; c += blockDim.x) smem[c] = c_make(0.0, 0.0);
    if (tid == 0) { sba[0] = p.b[i]; sba[1] = p.A[i * D + i]; sba[2] = p.A[i * D + i + NPD]; }
    for (int c = tid; c < nrows * e[2]; c += blockDim.x) {
        const int r = c % e[2], jj = c / e[2], y0 = jj / e[1], y1 = jj % e[1];
        celltab[c] = make_uint2((unsigned)(((y0 + 1) * W1 + (y1 + 1)) * RS + (r % R) * C + r / R) * 16u,
                                (unsigned)((lo[0] + y0) * gst[0] + (lo[1] + y1) * gst[1] + lo[2] + r));
    }
    // ---- per-lane constants (compute warps) -------------------------------------------------------------------
    const int q = tidc;
    const int j = q >= 0 ? q / C : 0, ch = q >= 0 ? q % C : 0;
    const int x0 = j / e[1], x1 = j % e[1];
    const int r0 = ch * R;
    const bool act = q >= 0 && j < nrows && r0 < e[2];
    const int nact = act ? (e[2] - r0 < R ? e[2] - r0 : R) : 0;
    const int erow = (x0 + 1) * W1 + (x1 + 1);            // row index in the extended grid
    // cell r of chunk ch sits at r * C + ch inside its row: for a fixed r the lanes of a warp read consecutive 16-byte cells (the row
    // stride RS is congruent to C modulo 8, so the rows of a quarter-warp tile the eight bank groups): LDS.128 / STS.128 in four
    // wavefronts.  (With the chunk's cells adjacent, R = 2 puts the lanes 32 bytes apart: eight wavefronts, measured 35 % of all
    // shared-memory wavefronts of the kernel.)
    const unsigned cstep = 16u * (unsigned)C;
    const unsigned own_off = act ? (unsigned)(erow * RS + ch) * 16u : zero_base + 16u;
    const unsigned n0_off = act ? own_off - (unsigned)(W1 * RS) * 16u : own_off;   // same cells, row one lower in panel dim 0
    const unsigned n1_off = act ? own_off - (unsigned)RS * 16u : own_off;          // ... in panel dim 1
    const unsigned nbl_off = act ? (ch > 0 ? own_off - 16u + (unsigned)(R - 1) * cstep : h2_base + (unsigned)erow * 16u) : zero_base;   // the cell before the lane's first: last cell of chunk ch - 1
    const c128 *Arow = p.A + i * D + i;
    c128 c0 = c_make(0.0, 0.0), c1 = c_make(0.0, 0.0);                             // A_ij sqrt(k_j)  (core.py:103)
    if (NPD >= 3 && act && lo[0] + x0 > 0) c0 = c_scale(Arow[1], p.sq[lo[0] + x0]);
    if (NPD >= 2 && act && lo[1] + x1 > 0) c1 = c_scale(Arow[NPD - 1], p.sq[lo[1] + x1]);
    const unsigned gofs = (unsigned)((lo[0] + x0) * gst[0] + (lo[1] + x1) * gst[1] + lo[2] + r0);
    // exports: where do the upper neighbour boxes expect this lane's cells (X cell offsets inside a panel row of the exchange buffer)?
    // bit 0 / 1: the lane's row lies on the high face of panel dim 0 / 1 (all its cells go); bit 2: it holds the row's last cell
    unsigned xo0 = 0, xo1 = 0, xo2 = 0, xsrc2 = 0, xflags = 0;
    if (act && t[0] + 1 < g[0] && x0 == e[0] - 1) {   // consumer = box + g1 g2; its face 0 comes first
        xflags |= 1u;
        xo0 = (unsigned)((size_t)(tile + g[1] * g[2]) * xtile) + (unsigned)(x1 * e[2] + r0);
    }
    if (act && t[1] + 1 < g[1] && x1 == e[1] - 1) {   // consumer = box + g2; its face 1 follows its face 0
        const int e1c = (int)(((long long)(t[1] + 2) * shp[1]) / g[1]) - (int)(((long long)(t[1] + 1) * shp[1]) / g[1]);
        xflags |= 2u;
        xo1 = (unsigned)((size_t)(tile + g[2]) * xtile) + (unsigned)(h[0] * e1c * e[2] + x0 * e[2] + r0);
    }
    if (act && t[2] + 1 < g[2] && r0 <= e[2] - 1 && e[2] - 1 < r0 + R) {   // consumer = box + 1; its face 2 comes last
        const int e2c = (int)(((long long)(t[2] + 2) * shp[2]) / g[2]) - (int)(((long long)(t[2] + 1) * shp[2]) / g[2]);
        xflags |= 4u;
        xo2 = (unsigned)((size_t)(tile + 1) * xtile) + (unsigned)(h[0] * e[1] * e2c + h[1] * e[0] * e2c + x0 * e[1] + x1);
        xsrc2 = own_off + (unsigned)(e[2] - 1 - r0) * cstep;
    }
    double sq2[R];                                        // sqrt(k_last) of the lane's cells (0 beyond the row / at k = 0)
#pragma unroll
    for (int r = 0; r < R; r++) sq2[r] = (act && r < nact) ? p.sq[lo[2] + r0 + r] : 0.0;
    __syncthreads();   // zero fill and tables complete

    const unsigned long long t_giveup0 = rw_timer() + 4000000000ull;
    if (!p.poll0) asm volatile("griddepcontrol.wait;" ::: "memory");
    else {
        if (tid == 0) {   // one thread watches the box's last amplitude of panel 0 before everybody polls
            const long long last = (long long)(lo[0] + e[0] - 1) * gst[0] + (long long)(lo[1] + e[1] - 1) * gst[1] + lo[2] + e[2] - 1;
            (void)rw_poll(p.G + last, t_giveup0, p.err);
        }
        __syncthreads();
    }
    if (tl) timeline_stamp(p.timeline, i & 7, 1);
    // panel 0: low halo faces -> buffer 0, and the table of halo destinations
    for (int c = tid; c < HC; c += blockDim.x) {
        unsigned dst;
        int go;
        if (c < F0) {
            const int r = c % e[2], y1 = c / e[2];
            dst = (unsigned)((y1 + 1) * RS + (r % R) * C + r / R) * 16u;
            go = (lo[0] - 1) * gst[0] + (lo[1] + y1) * gst[1] + lo[2] + r;
        } else if (c < F0 + F1) {
            const int cc = c - F0, r = cc % e[2], y0 = cc / e[2];
            dst = (unsigned)((y0 + 1) * W1 * RS + (r % R) * C + r / R) * 16u;
            go = (lo[0] + y0) * gst[0] + (lo[1] - 1) * gst[1] + lo[2] + r;
        } else {
            const int cc = c - F0 - F1, y1 = cc % e[1], y0 = cc / e[1];
            dst = h2_base + (unsigned)((y0 + 1) * W1 + y1 + 1) * 16u;
            go = (lo[0] + y0) * gst[0] + (lo[1] + y1) * gst[1] + lo[2] - 1;
        }
        hdst[c] = dst;
        rw_sts(sbase + dst, p.poll0 ? rw_poll(p.G + go, t_giveup0, p.err) : __ldcg(p.G + go));
    }
    c128 P1[R], acc[R];
    if (tidc >= 0) {
        if (p.poll0) {
#pragma unroll
            for (int r = 0; r < R; r++) P1[r] = r < nact ? rw_poll(p.G + gofs + r, t_giveup0, p.err) : c_make(0.0, 0.0);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) P1[r] = r < nact ? __ldcg(p.G + gofs + r) : c_make(0.0, 0.0);
        }
#pragma unroll
        for (int r = 0; r < R; r++) if (act) rw_sts(sbase + own_off + cstep * (unsigned)r, P1[r]);
    }
    __syncthreads();

    const int lane = tid & 31, wid = tid >> 5;
    if (wid < MMH_RW_NSV) {
        // ================= service warps: warp w owns panel buffer w and the panels s = w (mod NB) =================
        // per panel: import its low halo faces (needed by step s+1), then -- once the compute warps have stored panel s --
        // export its high faces, drain it to the lattice, hand the buffer back.  A lower neighbour box never waits for this box,
        // so polling for the halo before exporting cannot deadlock; in the steady state of the box pipeline the halo of panel s
        // arrives while the compute warps are still in step s.
        constexpr int HCL = MAXT >= 640 ? 5 : 10;   // imported cells per lane and round (register budget: 96 at 640 threads)
        constexpr int DB = MAXT >= 640 ? 4 : 8;     // drained cells per lane and batch
        const int k = wid;
        const unsigned buf = sbase + (unsigned)k * bstride;
        const c128 *xin = p.X + (size_t)tile * xtile;
        const int ncell = nrows * e[2];
        const unsigned long long t_giveup = rw_timer() + 4000000000ull;
#pragma unroll 1
        for (int s = wid == 0 ? NB : wid; s < S; s += NB) {
            const bool imp = HC > 0 && s <= S - 2;
            c128 *src = const_cast<c128 *>(xin) + (size_t)s * p.hc_max;
            if (imp) {
                if (s >= NB) rw_bar_sync(MMH_RW_BAR_FREE + k, TC + 32);   // buffer k last held panel s-NB, read during step s-NB+1
                if (p.trace && lane == 0) p.trace[((size_t)tile * S + s) * 8 + 5] = rw_timer();
#pragma unroll 1
                for (int cbase = 0; cbase < HC; cbase += 32 * HCL) {
                    const int c0_ = cbase + lane;
                    unsigned long long a[HCL], b[HCL];
                    unsigned pend = 0u;
#pragma unroll
                    for (int w = 0; w < HCL; w++) if (c0_ + 32 * w < HC) pend |= 1u << w;
                    unsigned rounds = 0;
#pragma unroll 1
                    while (true) {
#pragma unroll
                        for (int w = 0; w < HCL; w++) if ((pend >> w) & 1u) rw_ldg_relaxed(src + c0_ + 32 * w, a[w], b[w]);
#pragma unroll
                        for (int w = 0; w < HCL; w++)
                            if (((pend >> w) & 1u) && a[w] != MMH_RW_SENTINEL && b[w] != MMH_RW_SENTINEL) {
                                pend &= ~(1u << w);
                                rw_sts(buf + hdst[c0_ + 32 * w], make_double2(__longlong_as_double((long long)a[w]), __longlong_as_double((long long)b[w])));
                            }
                        if (!__any_sync(0xffffffffu, pend != 0u)) break;
                        if ((++rounds & 255u) == 0u && rw_timer() > t_giveup) {
                            if (p.err) *(volatile int *)p.err = 1;
                            break;
                        }
                        __nanosleep(40);
                        __syncwarp();
                    }
                }
                if (p.trace && lane == 0) p.trace[((size_t)tile * S + s) * 8 + 4] = rw_timer();
                __syncwarp();
#if !MMH_RW_HANDOFF_BAR
                if (lane == 0) rw_mbar_arrive(mb_halo + 8u * (unsigned)k);
#endif
                if (p.trace && lane == 0) p.trace[((size_t)tile * S + s) * 8 + 3] = rw_timer();
            }
#if MMH_RW_HANDOFF_BAR
            rw_bar_sync(MMH_RW_BAR_FULL + k, TC + 32);
#else
            rw_mbar_wait(mb_own + 8u * (unsigned)k, (unsigned)((s - 1) >> 2) & 1u);   // the compute warps have stored panel s
#endif
            c128 *gpan = p.G + (size_t)s * P;
#pragma unroll 1
            for (int cb = 0; cb < ncell; cb += 32 * DB) {
                uint2 ct[DB];
                c128 dv[DB];
#pragma unroll
                for (int w = 0; w < DB; w++) { const int c = cb + 32 * w + lane; ct[w] = celltab[c < ncell ? c : 0]; }
#pragma unroll
                for (int w = 0; w < DB; w++) dv[w] = rw_lds(buf + ct[w].x);
                if (p.strong_g) {   // the next stage validates these amplitudes by polling (stage overlap): strong stores
#pragma unroll
                    for (int w = 0; w < DB; w++) if (cb + 32 * w + lane < ncell) rw_stg_relaxed(gpan + ct[w].y, dv[w]);
                } else if (!(p.dbg & 1)) {
#pragma unroll
                    for (int w = 0; w < DB; w++) if (cb + 32 * w + lane < ncell) __stcs(gpan + ct[w].y, dv[w]);
                }
            }
            __syncwarp();
#if MMH_RW_HANDOFF_BAR
            if (s + NB < S) rw_bar_arrive(MMH_RW_BAR_DRAINED + k, TC + 32);
#else
            if (lane == 0) rw_mbar_arrive(mb_drained + 8u * (unsigned)k);
#endif
            if (imp) for (int c = lane; c < HC; c += 32) rw_stg_relaxed_u64(src + c, MMH_RW_SENTINEL, MMH_RW_SENTINEL);   // self-cleaning
        }
        return;
    }

    // ================= compute warps =================
    const c128 b0 = sba[0], a00 = sba[1], a2 = sba[2];
    const bool have_halo = HC > 0;
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = c_mul(b0, P1[r]);   // step 1: A_ii sqrt(0) P2 is skipped (core.py:100)
    if (tl) timeline_stamp(p.timeline, i & 7, 2);
#if MMH_RW_DATAFLOW && !MMH_RW_HANDOFF_BAR
    // Per-warp hand-off.  Lanes are row-major, so warp w reads (a) the previous chunk of a row (lane - 1), (b) row x1 - 1 (lane - C),
    // (c) row x0 - 1 (lane - e1 C): its own warp, the previous one and at most four more.  Read-after-write: wait for those warps' panel s before
    // step s+1.  Write-after-read: the warps that read MY rows (w + 1 and the one or two warps e1 * C lanes higher) must have
    // finished step s-3 (they read panel s-4 there) before panel s goes into the same buffer.
    const int cw = tidc >> 5, nwc = TC >> 5, dl = e[1] * C;
    constexpr int ND = 5;
    int raw[ND], war[ND];
    raw[0] = cw - 1;                                                         // previous chunk of the row (lane - 1)
    raw[1] = 32 * cw - C >= 0 ? (32 * cw - C) >> 5 : -1;                     // row x1 - 1 (lane - C)
    raw[2] = 32 * cw + 31 - C >= 0 ? (32 * cw + 31 - C) >> 5 : -1;
    raw[3] = 32 * cw - dl >= 0 ? (32 * cw - dl) >> 5 : -1;                   // row x0 - 1 (lane - e1 C)
    raw[4] = 32 * cw + 31 - dl >= 0 ? (32 * cw + 31 - dl) >> 5 : -1;
    war[0] = cw + 1;
    war[1] = (32 * cw + C) >> 5;
    war[2] = (32 * cw + 31 + C) >> 5;
    war[3] = (32 * cw + dl) >> 5;
    war[4] = (32 * cw + 31 + dl) >> 5;
#pragma unroll
    for (int a = 0; a < ND; a++) {
        if (raw[a] >= cw) raw[a] = -1;                         // own warp: __syncwarp orders it
        if (war[a] >= nwc || war[a] <= cw) war[a] = -1;
#pragma unroll
        for (int b = 0; b < a; b++) { if (raw[a] == raw[b]) raw[a] = -1; if (war[a] == war[b]) war[a] = -1; }
    }
#endif

#pragma unroll 1
    for (int s = 1; s < S; s++) {
        const int k = s & (NB - 1);
        const unsigned bprev = sbase + (unsigned)((s - 1) & (NB - 1)) * bstride;
        const unsigned bcur = sbase + (unsigned)k * bstride;
        const double2 st = sqtab[s];
        if (p.trace && tidc == 0) p.trace[((size_t)tile * S + s) * 8 + 0] = rw_timer();
        const c128 nbl = rw_lds(bprev + nbl_off);   // cell before the lane's first one: previous chunk / halo of dim 2
#pragma unroll
        for (int r = 0; r < R; r++) {
            c128 v = acc[r];
            if (NPD >= 3) v = c_add(v, c_mul(c0, rw_lds(bprev + n0_off + cstep * (unsigned)r)));
            if (NPD >= 2) v = c_add(v, c_mul(c1, rw_lds(bprev + n1_off + cstep * (unsigned)r)));
            v = c_add(v, c_mul(c_scale(a2, sq2[r]), r ? P1[r > 0 ? r - 1 : 0] : nbl));
            acc[r] = v;
        }
        if (have_halo && s + 3 <= S - 2) rw_bar_arrive(MMH_RW_BAR_FREE + ((s - 1) & (NB - 1)), TC + 32);   // panel s-1 has been read
        if (p.trace && tidc == 0) p.trace[((size_t)tile * S + s) * 8 + 6] = rw_timer() + 0 * (unsigned long long)__double_as_longlong(acc[0].x + acc[R - 1].y);
        rw_div_all<R>(acc, st.x, st.y);
        if (p.trace && tidc == 0) p.trace[((size_t)tile * S + s) * 8 + 7] = rw_timer() + 0 * (unsigned long long)__double_as_longlong(acc[0].x + acc[R - 1].y);
        if (xflags && s <= S - 2) {   // exports first (the upper neighbours' critical path): strong stores, polled by the consumers
            c128 *xpan = p.X + (size_t)s * p.hc_max;
            if (xflags & 1u) {
#pragma unroll
                for (int r = 0; r < R; r++) if (r < nact) rw_stg_relaxed(xpan + xo0 + r, acc[r]);
            }
            if (xflags & 2u) {
#pragma unroll
                for (int r = 0; r < R; r++) if (r < nact) rw_stg_relaxed(xpan + xo1 + r, acc[r]);
            }
        }
#if MMH_RW_HANDOFF_BAR
        if (s > NB) rw_bar_sync(MMH_RW_BAR_DRAINED + k, TC + 32);
#else
        if (s > NB) rw_mbar_wait(mb_drained + 8u * (unsigned)k, (unsigned)((s - NB - 1) >> 2) & 1u);   // panel s-NB has left this buffer
#if MMH_RW_DATAFLOW
        if (s > NB) {   // the readers of my rows are past panel s-NB: they have stored panel s-NB+1
            const int sr = s - NB + 1;
            int lowest;
            do {
                lowest = sr;
#pragma unroll
                for (int a = 0; a < ND; a++) if (war[a] >= 0) lowest = min(lowest, ld_acquire_cta_shared(prog + war[a]));
                if (lowest < sr) __nanosleep(64);
            } while (lowest < sr);
        }
#endif
#endif
#pragma unroll
        for (int r = 0; r < R; r++) rw_sts(bcur + own_off + cstep * (unsigned)r, acc[r]);
        if ((xflags & 4u) && s <= S - 2) rw_stg_relaxed(p.X + (size_t)s * p.hc_max + xo2, rw_lds(bcur + xsrc2));
#if MMH_RW_HANDOFF_BAR
        rw_bar_sync(MMH_RW_BAR_FULL + k, TC + 32);
#else
        rw_mbar_arrive(mb_own + 8u * (unsigned)k);   // my cells of panel s are in shared memory (the service warp waits for all of them)
#if MMH_RW_DATAFLOW
        __syncwarp();   // every lane's stores of step s before the warp's progress word (and before its other lanes' loads of step s+1)
        if ((tid & 31) == 0) st_release_cta_shared(prog + cw, s);
#endif
#endif
        // register-only part of step s+1: b_i P1 + A_ii sqrt(s) P2
        const c128 a00s = c_scale(a00, st.x);
#pragma unroll
        for (int r = 0; r < R; r++) {
            const c128 pn = c_add(c_mul(b0, acc[r]), c_mul(a00s, P1[r]));
            P1[r] = acc[r];
            acc[r] = pn;
        }
        if (p.trace && tidc == 0) p.trace[((size_t)tile * S + s) * 8 + 2] = rw_timer() + 0 * (unsigned long long)__double_as_longlong(acc[0].x + acc[R - 1].y);
#if !MMH_RW_HANDOFF_BAR
        if (s < S - 1) {
#if MMH_RW_DATAFLOW
            int lowest;
            do {
                lowest = s;
#pragma unroll
                for (int a = 0; a < ND; a++) if (raw[a] >= 0) lowest = min(lowest, ld_acquire_cta_shared(prog + raw[a]));
                if (lowest < s) __nanosleep(64);
            } while (lowest < s);
#else
            rw_mbar_wait(mb_own + 8u * (unsigned)k, (unsigned)((s - 1) >> 2) & 1u);
#endif
            if (have_halo) rw_mbar_wait(mb_halo + 8u * (unsigned)k, (unsigned)((s - 1) >> 2) & 1u);
        }
#endif
        if (p.trace && tidc == 0) p.trace[((size_t)tile * S + s) * 8 + 1] = rw_timer();
    }
    if (tl) timeline_stamp(p.timeline, i & 7, 3);
}

template <int R, int NPD, int MAXT>
cudaError_t launch_rows(const TiledParams &p, int ntiles, size_t smem, cudaStream_t st) {
    static size_t smem_set = 0;
    if (p.tc + 32 * MMH_RW_NSV > MAXT) return cudaErrorInvalidValue;
    if (smem > 48 * 1024 && smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(k_march_rows<R, NPD, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smem_set = smem;
    }
    return mmh_launch_ex(k_march_rows<R, NPD, MAXT>, dim3(ntiles), dim3(p.tc + 32 * MMH_RW_NSV), smem, st, p.pdl != 0, p);
}

template <int R, int MAXT>
cudaError_t launch_rows_npd(const TiledParams &p, int ntiles, size_t smem, cudaStream_t st) {
    switch (p.d.D - 1 - p.stage) {
        case 2: return launch_rows<R, 2, MAXT>(p, ntiles, smem, st);
        case 3: return launch_rows<R, 3, MAXT>(p, ntiles, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

// compute threads per CTA for R cells per lane: the register file (64 K) holds (threads + 128 service threads) x registers:
// R = 2: 640 x 96, R = 3: 512 x 128, R >= 4: 384 x 168
int mmh_rows_max_threads(int R) { return R <= 2 ? 512 : (R == 3 ? 384 : 256); }
bool mmh_rows_supported_R(int R) { return R >= 2 && R <= 6; }

size_t mmh_rows_smem(int ls_max, int hc_max, int S, int CR, int cells_max, int xc_max) {
    (void)CR; (void)xc_max;
    return sizeof(c128) * ((size_t)MMH_RW_NB * ls_max + (size_t)S + 4) + 8 * 3 * MMH_RW_NB + 4 * MMH_RW_NWMAX + 8 * (size_t)cells_max +
           sizeof(unsigned) * (size_t)(hc_max + 4);
}

cudaError_t mmh_launch_march_rows(const TiledParams &p, int ntiles, size_t smem, cudaStream_t st) {
    switch (p.rows_R) {
        case 2: return launch_rows_npd<2, 640>(p, ntiles, smem, st);
        case 3: return launch_rows_npd<3, 512>(p, ntiles, smem, st);
        case 4: return launch_rows_npd<4, 384>(p, ntiles, smem, st);
        case 5: return launch_rows_npd<5, 384>(p, ntiles, smem, st);
        case 6: return launch_rows_npd<6, 384>(p, ntiles, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

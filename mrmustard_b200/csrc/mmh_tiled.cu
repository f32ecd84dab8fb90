// mmh_tiled.cu — K1: tiled march of ONE lattice over many CTAs, sm_100a.
//
// The panel of stage i (all k with k_<i = 0, k_i = s) is cut into a grid of boxes over its first nt (<= 3) dims;
// CTA t owns box t for the whole march s = 1 .. shape[i]-1.  A point reads
//     G[k - e_i], G[k - 2 e_i]        same panel position, panels s-1, s-2   -> registers of the owning thread
//     G[k - e_i - e_j]  (j > i)       one cell lower in panel dim j, panel s-1 -> shared memory
// Shared memory holds NB = 4 rotating copies of the box: the box itself in slot order (consecutive slots = consecutive
// cells, so the own-cell store and the neighbour loads of a warp are unit stride and bank-conflict free), followed by the
// low halo faces owned by the lower neighbour boxes (in exchange order), a cell that always holds 0 (absent neighbours,
// k_j = 0, point there with coefficient 0) and a trash cell (inactive slots).  Four buffers let the halo warps deliver
// the halos of up to three panels ahead of the compute warps.
//
// Halo exchange without fences or flags: a producer stores every amplitude on a high face of its box also into
// X[consumer tile][panel][cell]; X is kept filled with a sentinel (all-ones, a NaN no FP64 instruction produces).  The
// consumer's halo warps -- one per panel buffer, so four panels are in flight -- load a whole panel's cells at once with
// L1-bypassing loads; both 64-bit words of a cell must differ from the sentinel (64-bit stores are single-copy atomic, so
// every word validates itself).  Missing cells are re-polled together after per-face canary cells have arrived.  The cells
// go into the halo area of the panel's buffer and the sentinel is written back afterwards (self-cleaning).
//
// Hand-offs never spin on flags: "panel s is complete in shared memory" (all compute threads + the halo warp) is an
// mbarrier per panel buffer with split arrive / wait, "buffer k may be overwritten" is a named barrier the compute warps
// arrive on and the buffer's halo warp syncs on.  With TiledParams::poll0 the kernel does not wait for its predecessor
// kernel at all but validates panel 0 by the same sentinel (stage overlap).
//
// The compute step is written for a short instruction stream (255 warp-instructions per warp-step at R = 2, 100 of them
// FP64): per-slot constants live in registers, shared memory is addressed with 32-bit offsets, and the register-only part
// of step s+1 (b_i P1 + A_ii sqrt(s) P2) is issued before the hand-off of step s.
#include <cstring>

#include "mmh_params.cuh"

#define MMH_T2_NB 4      // panel buffers (power of two)
#ifndef MMH_T2_PRE_EARLY
#define MMH_T2_PRE_EARLY 1   // 1: the register-only part of step s+1 is issued before the barrier of step s; 0: after its loads
#endif
#ifndef MMH_T2_HANDOFF_BAR
#define MMH_T2_HANDOFF_BAR 0   // 1: debug build for compute-sanitizer racecheck -- the step hand-off ("every compute thread has stored
                               //    panel s and its halo has arrived") goes through named barriers (bar.sync / bar.arrive), which the tool
                               //    models, instead of the mbarrier split arrive / wait, which it reports as unsynchronised
#endif
#define MMH_T2_BAR_FULL 1    // + k (debug hand-off only): panel buffer k is complete
#define MMH_T2_NHW 4     // halo warps
#define MMH_SENTINEL 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ void pdl_launch_dependents2() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait2() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void ldg_relaxed_v2(const void *p, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void stg_relaxed_v2(void *p, unsigned long long a, unsigned long long b) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
// amplitudes another CTA / kernel validates by polling (halo exports, and the lattice itself under stage overlap) are written with
// strong stores: the PTX memory model only orders a polling ld.relaxed.gpu against stores that are themselves strong
__device__ __forceinline__ void stg_relaxed_c128(c128 *p, c128 v) {
    asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
// panel-0 amplitude that may not have been written yet (stage overlap): poll until both words differ from the sentinel
__device__ __forceinline__ c128 poll_c128(const c128 *p, unsigned long long t_giveup, int *err) {
    unsigned long long a, b;
    unsigned spins = 0;
    ldg_relaxed_v2(p, a, b);
    while ((a == MMH_SENTINEL || b == MMH_SENTINEL) && ((++spins & 255u) != 0u || gtimer_ns() < t_giveup)) {
        __nanosleep(100);
        ldg_relaxed_v2(p, a, b);
    }
    if ((a == MMH_SENTINEL || b == MMH_SENTINEL) && err) *(volatile int *)err = 1;   // watchdog expired: reported by the next API call
    return make_double2(__longlong_as_double((long long)a), __longlong_as_double((long long)b));
}
// shared memory through 32-bit addresses (byte offsets in the shared window)
__device__ __forceinline__ c128 lds_c128(unsigned addr) {
    c128 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_c128(unsigned addr, c128 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
// thread-block cluster / distributed shared memory (cluster variant: the whole tile grid is ONE cluster, halo cells are pushed straight
// into the consumer tile's shared memory instead of through L2)
__device__ __forceinline__ unsigned mapa_shared(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_c128(unsigned addr, c128 v) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void lds_volatile_v2(unsigned addr, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// mbarrier (shared memory, CTA scope): split arrive / wait, so that the work after a thread's last shared-memory store of a
// step (lattice store, exports, the register part of the next step) overlaps the wait for the slowest warp and for the halo
__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
// named barriers: hardware-blocked waits (a waiting warp takes no issue slots, unlike a spin on a shared-memory flag)
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// (the step hand-off -- every compute thread has written panel s and the halo of panel s has arrived -- is an mbarrier per
//  panel buffer: count = compute threads + 1 arrival of the buffer's halo warp)
#define MMH_T2_BAR_HALO 10   // halo warps among themselves
#define MMH_T2_BAR_FREE 6    // + k: panel buffer k has been read by every compute warp (compute warps arrive, halo warp k syncs)

template <int R>
__device__ __forceinline__ void div_all2(c128 (&v)[R], double sqs, double rsqs) {
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

// smem layout (bytes):  buf[NB][ls_max] c128 | sqtab[S] double2 | (b_i, A_ii) c128[2] | sync[8] i32 |
//                       xo[3][R * TC] u32 (export offsets of the high-face slots)
// CL: cluster variant.  The tile grid (<= 16 tiles) is launched as one thread-block cluster; the exchange buffer lives in the
// CONSUMER's shared memory ([S][hc_max] cells, sentinel filled at kernel start), producers push their high-face amplitudes there with
// st.shared::cluster, the halo warps poll local shared memory.  Same sentinel protocol, no L2 round trip (a hop through L2 costs
// ~1-2 us for small tiles, and stage 1 of a 4-index lattice pays it on every pipeline hop), nothing to clean up afterwards.
template <int R, int NPD, bool CL>
__global__ void __launch_bounds__(512 + 32 * MMH_T2_NHW, 1) k_march_tiled2(TiledParams p) {
    extern __shared__ c128 smem[];
    const LatticeDesc &d = p.d;
    const int D = d.D;
    const int i = p.stage;
    const long long P = d.strides[i];
    const int S = d.shape[i];
    const int TC = p.tc;
    const int tid = threadIdx.x;
    const int tidc = tid - 32 * MMH_T2_NHW;
    const int nt = p.nt;
    constexpr int NB = MMH_T2_NB;

    pdl_launch_dependents2();
    const bool tl = blockIdx.x == 0 && tid == 32 * MMH_T2_NHW;   // the timeline thread: compute thread 0 of tile 0
    if (tl) timeline_stamp(p.timeline, i & 7, 0);
    // ---- tile geometry ----------------------------------------------------------------------------------
    int g[3], t[3], lo[3], e[3], h[3], gst[3], shp[3];
#pragma unroll
    for (int m = 0; m < 3; m++) g[m] = m < nt ? p.g[m] : 1;
    const int tile = blockIdx.x;
    t[2] = tile % g[2];
    t[1] = (tile / g[2]) % g[1];
    t[0] = tile / (g[1] * g[2]);
#pragma unroll
    for (int m = 0; m < 3; m++) {
        if (m < nt) {
            shp[m] = d.shape[i + 1 + m];
            lo[m] = (int)(((long long)t[m] * shp[m]) / g[m]);
            e[m] = (int)(((long long)(t[m] + 1) * shp[m]) / g[m]) - lo[m];
            h[m] = lo[m] > 0 ? 1 : 0;
            gst[m] = (int)d.strides[i + 1 + m];
        } else { shp[m] = 1; lo[m] = 0; e[m] = 1; h[m] = 0; gst[m] = 0; }
    }
    const int inner = (int)d.strides[i + nt];
    int cst[3];                                   // strides of the compact box (slot order)
    cst[2] = inner;
    cst[1] = cst[2] * e[2];
    cst[0] = cst[1] * e[1];
    const int TS = e[0] * e[1] * e[2] * inner;    // slots of this tile
    int faceoff[3];
    faceoff[0] = 0;
    faceoff[1] = faceoff[0] + h[0] * (TS / e[0]);
    faceoff[2] = faceoff[1] + h[1] * (TS / e[1]);
    const int HC = faceoff[2] + h[2] * (TS / e[2]);   // halo cells of this tile

    // panel buffer (c128 cells): [0, TS) own cells in slot order | [TS, TS + HC) low halo faces in X order | ... |
    //                            [ls_max - 2] always 0 | [ls_max - 1] trash
    // Consecutive slots are consecutive cells, so the own-cell store and the three neighbour loads of a warp are
    // unit-stride (bank-conflict free); only the lanes on a low face divert to the halo area.
    const unsigned bstride = (unsigned)p.ls_max * 16u;              // bytes between panel buffers
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned zero_off = (unsigned)(p.ls_max - 2) * 16u;
    const unsigned trash_off = (unsigned)(p.ls_max - 1) * 16u;      // written by inactive slots, never read
    const unsigned halo_off = (unsigned)TS * 16u;
    double2 *sqtab = (double2 *)(smem + (size_t)NB * p.ls_max);
    c128 *sba = (c128 *)(sqtab + S);
    const unsigned sync_base = (unsigned)__cvta_generic_to_shared(sba + 2);
    const size_t xtile = (size_t)S * p.hc_max;    // X elements per consumer tile
    const unsigned xo_base = sync_base + 32u;     // export offsets (c128 units inside X's panel row), [m][R * TC]
    // cluster variant: this tile's exchange rows [S][hc_max] in its own shared memory, 16-byte aligned behind the offset table
    const unsigned xs_base = (xo_base + 4u * (unsigned)(3 * R * TC) + 15u) & ~15u;
    unsigned xs_up[3] = { 0u, 0u, 0u };           // the same buffer in the upper neighbour tiles (cluster shared window)
    if (CL) {
#pragma unroll
        for (int m = 0; m < 3; m++)
            if (m < nt && t[m] + 1 < g[m]) xs_up[m] = mapa_shared(xs_base, (unsigned)(tile + (m == 0 ? g[1] * g[2] : (m == 1 ? g[2] : 1))));
        for (unsigned c = tid; c < (unsigned)(S * p.hc_max); c += blockDim.x)
            asm volatile("st.shared.v2.u64 [%0], {%1, %2};" ::"r"(xs_base + 16u * c), "l"(MMH_SENTINEL), "l"(MMH_SENTINEL) : "memory");
    }

    // ---- tables that do not touch the lattice (overlap the previous stage's kernel under PDL) ---------------
    if (tid < NB) mbar_init(sync_base + 8u * (unsigned)tid, (unsigned)TC + (HC > 0 ? 1u : 0u));
    for (int s_ = tid; s_ < S; s_ += blockDim.x) sqtab[s_] = make_double2(p.sq[s_], p.rsq[s_]);
    if (tid < NB) { smem[(size_t)tid * p.ls_max + p.ls_max - 2] = c_make(0.0, 0.0); smem[(size_t)tid * p.ls_max + p.ls_max - 1] = c_make(0.0, 0.0); }
    if (tid == 0) { sba[0] = p.b[i]; sba[1] = p.A[i * D + i]; }

    // ---- per-slot constants (compute warps) -----------------------------------------------------------------
    unsigned loco[R];           // byte offset of the slot inside a panel buffer (trash cell when inactive)
    unsigned nbo[R][NPD];       // byte offset of the neighbour k - e_i - e_j (zero cell when k_j == 0)
    c128 coef[R][NPD];          // A_ij sqrt(k_j)  (core.py:103), constant along the march
    unsigned gofs[R];           // panel offset f of the slot: G index = s * P + f
    unsigned flags[R];          // bit 0: active; bit 1+m: exported through the high face of tiled dim m
    c128 h0[R], h1[R];
    const c128 *Arow = p.A + i * D + i;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int q = r * TC + tidc;
        const bool act = tidc >= 0 && q < TS;
        const int qq = act ? q : 0;
        const int rr = qq % inner;
        int q1 = qq / inner;
        int x[3];
        x[2] = q1 % e[2]; q1 /= e[2];
        x[1] = q1 % e[1];
        x[0] = q1 / e[1];
        loco[r] = act ? (unsigned)qq * 16u : trash_off;
        gofs[r] = (unsigned)((lo[0] + x[0]) * gst[0] + (lo[1] + x[1]) * gst[1] + (lo[2] + x[2]) * gst[2] + rr);
        flags[r] = act ? 1u : 0u;
        int rem = rr;
#pragma unroll
        for (int jj = 0; jj < NPD; jj++) {
            int k, nb;
            if (jj < nt) {
                k = lo[jj] + x[jj];
                if (x[jj] > 0) nb = qq - cst[jj];
                else {   // one cell below the box: halo face jj (exists iff lo[jj] > 0, i.e. k > 0)
                    const int a = jj == 0 ? 1 : 0, b = jj == 2 ? 1 : 2;
                    nb = TS + faceoff[jj] + (x[a] * e[b] + x[b]) * inner + rr;
                }
            } else {
                const int sj = (int)d.strides[i + 1 + jj];
                k = rem / sj;
                rem -= k * sj;
                nb = qq - sj;
            }
            const bool has = act && k > 0;
            nbo[r][jj] = has ? (unsigned)nb * 16u : zero_off;
            coef[r][jj] = has ? c_scale(Arow[1 + jj], p.sq[k]) : c_make(0.0, 0.0);
        }
        // high faces: where does the upper neighbour in dim m expect this amplitude?
#pragma unroll
        for (int m = 0; m < 3; m++) {
            if (act && m < nt && t[m] + 1 < g[m] && x[m] == e[m] - 1) {
                const int lo_up = (int)(((long long)(t[m] + 1) * shp[m]) / g[m]);
                const int e_up = (int)(((long long)(t[m] + 2) * shp[m]) / g[m]) - lo_up;
                int ec[3] = { e[0], e[1], e[2] }, hc[3] = { h[0], h[1], h[2] };
                ec[m] = e_up; hc[m] = 1;
                const int TSc = ec[0] * ec[1] * ec[2] * inner;
                int fo = 0;
                for (int mm = 0; mm < m; mm++) fo += hc[mm] * (TSc / ec[mm]);
                const int a = m == 0 ? 1 : 0, b = m == 2 ? 1 : 2;
                const int up = tile + (m == 0 ? g[1] * g[2] : (m == 1 ? g[2] : 1));   // the consumer tile
                flags[r] |= 2u << m;
                const unsigned xo = (CL ? 0u : (unsigned)(up * xtile)) + (unsigned)(fo + (x[a] * ec[b] + x[b]) * inner + rr);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(xo_base + 4u * (unsigned)(m * R * TC + q)), "r"(xo) : "memory");
            }
        }
    }
    __syncthreads();
    if (CL) cluster_sync_all();   // every tile's exchange rows hold the sentinel before any producer pushes into them
    // panel 0 (written by the previous stages' kernels) and X are touched from here on.  Normally that waits for the previous
    // kernel to complete.  With poll0 the previous stage is still marching: every tile starts as soon as ITS part of panel 0
    // exists (the host pre-filled panel 0 with the sentinel), so the tile pipeline of this stage fills while the previous
    // stage finishes.  One thread first watches the last amplitude of the box (the previous stage writes panel 0 in
    // ascending order of this box's first dim), so that 600 threads do not poll for tens of microseconds.
    const unsigned long long t_giveup0 = gtimer_ns() + 4000000000ull;
    if (!p.poll0) pdl_wait2();
    else {
        if (tid == 0) {
            const long long last = (long long)(lo[0] + e[0] - 1) * gst[0] + (long long)(lo[1] + e[1] - 1) * gst[1] +
                                   (long long)(lo[2] + e[2] - 1) * gst[2] + inner - 1;
            (void)poll_c128(p.G + last, t_giveup0, p.err);
        }
        __syncthreads();
    }
    if (tl) timeline_stamp(p.timeline, i & 7, 1);
    // panel 0: halo faces and own cells -> buffer 0
    for (int c = tid; c < HC; c += blockDim.x) {
        int m = 0;
        if (c >= faceoff[2] && h[2]) m = 2;
        else if (c >= faceoff[1] && h[1]) m = 1;
        int cc = c - (m == 0 ? faceoff[0] : (m == 1 ? faceoff[1] : faceoff[2]));
        const int a = m == 0 ? 1 : 0, b = m == 2 ? 1 : 2;   // the two other tiled dims, in order
        const int ea = a == 0 ? e[0] : e[1], eb = b == 1 ? e[1] : e[2];
        (void)ea;
        const int rr = cc % inner; cc /= inner;
        const int xb = cc % eb, xa = cc / eb;
        const int lom = m == 0 ? lo[0] : (m == 1 ? lo[1] : lo[2]), gm = m == 0 ? gst[0] : (m == 1 ? gst[1] : gst[2]);
        const int loa = a == 0 ? lo[0] : lo[1], ga = a == 0 ? gst[0] : gst[1];
        const int lob = b == 1 ? lo[1] : lo[2], gb = b == 1 ? gst[1] : gst[2];
        const int go = (lom - 1) * gm + (loa + xa) * ga + (lob + xb) * gb + rr;
        sts_c128(sbase + halo_off + 16u * (unsigned)c, p.poll0 ? poll_c128(p.G + go, t_giveup0, p.err) : __ldcg(p.G + go));
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        h0[r] = c_make(0.0, 0.0);
        h1[r] = (flags[r] & 1u) ? (p.poll0 ? poll_c128(p.G + gofs[r], t_giveup0, p.err) : __ldcg(p.G + gofs[r])) : c_make(0.0, 0.0);
        if (tidc >= 0) sts_c128(sbase + loco[r], h1[r]);
    }
    __syncthreads();

    if (tid < 32 * MMH_T2_NHW) {
        // ================= halo warps: warp w owns panel buffer w and serves the panels u = w (mod NB) =================
        // NB panels are in flight at once, so a panel's service time is off the critical path whenever the producers are
        // ahead.  The first attempt loads all cells of the panel at once (HCL per lane): ONE L2 round trip when the data
        // is there.  If cells are missing, three lanes watch one canary cell per face (its last cell, exported by the
        // producer's last warps) so that a waiting tile does not hammer L2, then the missing cells are re-polled at once.
        if (HC == 0) { if (CL) cluster_sync_all(); return; }
        static_assert(MMH_T2_NHW == MMH_T2_NB, "one halo warp per panel buffer");
        constexpr int HCL = 12;   // cells per lane and round
        const int lane = tid & 31, hw = tid >> 5;
        const c128 *xin = p.X + (size_t)tile * xtile;
        const unsigned dst = sbase + (unsigned)hw * bstride + halo_off;
        const int can_h = lane == 0 ? h[0] : (lane == 1 ? h[1] : h[2]);
        const int can_c = (lane == 0 ? faceoff[1] : (lane == 1 ? faceoff[2] : HC)) - 1;   // last cell of face `lane`
        // watchdog: a producer that never delivers (a bug, or a lost CTA) must not hang the device: after 4 s without
        // progress the polls give up and the march runs on with whatever is in the exchange buffer (results invalid)
        const unsigned long long t_giveup = gtimer_ns() + 4000000000ull;
#pragma unroll 1
        for (int u = hw == 0 ? NB : hw; u <= S - 2; u += NB) {
            if (u - (NB - 1) >= 1) bar_sync(MMH_T2_BAR_FREE + hw, TC + 32);   // buffer hw last held panel u-NB, read during step u-NB+1
            if (p.trace && lane == 0) p.trace[((size_t)tile * S + u) * 8 + 5] = gtimer_ns();
            c128 *src = const_cast<c128 *>(xin) + (size_t)u * p.hc_max;
            const unsigned lsrc = xs_base + 16u * (unsigned)(u * p.hc_max);   // (cluster variant: the same row in local shared memory)
#pragma unroll 1
            for (int cbase = 0; cbase < HC; cbase += 32 * HCL) {   // (warp-uniform trip count: __any_sync below)
                const int c0 = cbase + lane;
                unsigned long long a[HCL], b[HCL];
                unsigned pend = 0u;
#pragma unroll
                for (int w = 0; w < HCL; w++) if (c0 + 32 * w < HC) pend |= 1u << w;
                unsigned rounds = 0;
#pragma unroll 1
                while (true) {
#pragma unroll
                    for (int w = 0; w < HCL; w++)
                        if ((pend >> w) & 1u) {
                            if (CL) lds_volatile_v2(lsrc + 16u * (unsigned)(c0 + 32 * w), a[w], b[w]);
                            else ldg_relaxed_v2(src + c0 + 32 * w, a[w], b[w]);
                        }
#pragma unroll
                    for (int w = 0; w < HCL; w++)
                        if (((pend >> w) & 1u) && a[w] != MMH_SENTINEL && b[w] != MMH_SENTINEL) {
                            pend &= ~(1u << w);
                            sts_c128(dst + 16u * (unsigned)(c0 + 32 * w), make_double2(__longlong_as_double((long long)a[w]), __longlong_as_double((long long)b[w])));
                        }
                    if (!__any_sync(0xffffffffu, pend != 0u)) break;
                    if ((++rounds & 255u) == 0u && gtimer_ns() > t_giveup) {
                        if (p.err) *(volatile int *)p.err = 1;
                        break;
                    }
                    if (lane < 3 && can_h) {   // wait for the faces' canaries before polling whole faces again
                        unsigned long long a_, b_;
                        unsigned spins = 0;
                        do {
                            if (CL) lds_volatile_v2(lsrc + 16u * (unsigned)can_c, a_, b_);
                            else ldg_relaxed_v2(src + can_c, a_, b_);
                        } while ((a_ == MMH_SENTINEL || b_ == MMH_SENTINEL) && ((++spins & 1023u) != 0u || gtimer_ns() < t_giveup));
                    }
                    __syncwarp();
                }
            }
            if (p.trace && lane == 0) p.trace[((size_t)tile * S + u) * 8 + 4] = gtimer_ns();
            __syncwarp();            // the lanes' shared-memory stores are ordered before lane 0's (release) arrive
#if MMH_T2_HANDOFF_BAR
            bar_arrive(MMH_T2_BAR_FULL + hw, TC + 32);
#else
            if (lane == 0) mbar_arrive(sync_base + 8u * (unsigned)hw);
#endif
            if (p.trace && lane == 0) p.trace[((size_t)tile * S + u) * 8 + 3] = gtimer_ns();
            // self-cleaning, off the critical path: put the sentinel back for the next launch
            if (!CL) for (int c = lane; c < HC; c += 32) stg_relaxed_v2(src + c, MMH_SENTINEL, MMH_SENTINEL);
        }
        if (CL) cluster_sync_all();   // (all threads of every tile meet here: no tile's shared memory goes away under a late push)
        return;
    }

    // ================= compute warps =================
    const c128 b0 = sba[0], a00 = sba[1];
    c128 *gpan = p.G + P;                              // panel s of the lattice
    c128 *xpan = p.X + p.hc_max;                       // panel s of the exchange rows
    const bool have_halo = HC > 0;
    c128 pre[R];                                       // register-only part of the next step
#pragma unroll
    for (int r = 0; r < R; r++) pre[r] = c_mul(b0, h1[r]);   // step 1: A_ii sqrt(0) P2 is skipped (core.py:100)

    // one panel step; P1 = panel s-1, P2 = panel s-2 (replaced by panel s)
#define MMH_T2_STEP(P1, P2, SCUR)                                                                     \
    {                                                                                                 \
        const int s_ = (SCUR);                                                                        \
        if (p.trace && tidc == 0) p.trace[((size_t)tile * S + s_) * 8 + 0] = gtimer_ns();             \
        const unsigned bprev = sbase + (unsigned)((s_ - 1) & (NB - 1)) * bstride;                     \
        const unsigned bcur = sbase + (unsigned)(s_ & (NB - 1)) * bstride;                            \
        const double2 st_ = sqtab[s_];                                                                \
        c128 v[R];                                                                                    \
        const c128 a00s = c_scale(a00, MMH_T2_PRE_EARLY ? st_.x : sqm);   /* A_ii sqrt(s-1) (late) / A_ii sqrt(s) (early) */ \
        _Pragma("unroll") for (int r = 0; r < R; r++) {                                               \
            c128 nbv[NPD];                                                                            \
            _Pragma("unroll") for (int jj = 0; jj < NPD; jj++) nbv[jj] = lds_c128(bprev + nbo[r][jj]); \
            if (!MMH_T2_PRE_EARLY) {   /* overlaps the shared-memory latency of the loads above */    \
                pre[r] = c_mul(b0, P1[r]);                                                            \
                if (s_ >= 2) pre[r] = c_add(pre[r], c_mul(a00s, P2[r]));                              \
            }                                                                                         \
            v[r] = pre[r];                                                                            \
            _Pragma("unroll") for (int jj = 0; jj < NPD; jj++) v[r] = c_add(v[r], c_mul(coef[r][jj], nbv[jj])); \
        }                                                                                             \
        if (p.trace && tidc == 0) p.trace[((size_t)tile * S + s_) * 8 + 6] = gtimer_ns() + 0 * (unsigned long long)__double_as_longlong(v[0].x + v[R - 1].y); \
        div_all2<R>(v, st_.x, st_.y);                                                                 \
        if (p.trace && tidc == 0) p.trace[((size_t)tile * S + s_) * 8 + 7] = gtimer_ns() + 0 * (unsigned long long)__double_as_longlong(v[0].x + v[R - 1].y); \
        const bool xch = s_ <= S - 2;   /* the last panel has no consumer */                          \
        /* order: exports (the neighbours' critical path), shared-memory copy + arrive, then everything nobody waits for */ \
        if (xch) {                                                                                    \
            _Pragma("unroll") for (int r = 0; r < R; r++)                                             \
                _Pragma("unroll") for (int m = 0; m < 3; m++)                                         \
                    if (m < NPD && (flags[r] & (2u << m))) {                                          \
                        unsigned xo_;                                                                 \
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(xo_) : "r"(xo_base + 4u * (unsigned)(m * R * TC + r * TC + tidc)) : "memory"); \
                        if (CL) st_cluster_c128(xs_up[m] + 16u * ((unsigned)(s_ * p.hc_max) + xo_), v[r]);             \
                        else stg_relaxed_c128(xpan + xo_, v[r]);   /* polled by the consumer: a strong (relaxed.gpu) store */ \
                    }                                                                                 \
        }                                                                                             \
        _Pragma("unroll") for (int r = 0; r < R; r++) sts_c128(bcur + loco[r], v[r]);                 \
        if (xch && !MMH_T2_HANDOFF_BAR) mbar_arrive(sync_base + 8u * (unsigned)(s_ & (NB - 1)));   /* my part of panel s is in shared memory */ \
        _Pragma("unroll") for (int r = 0; r < R; r++) {                                               \
            P2[r] = v[r];                                                                             \
            if (flags[r] & 1u) stg_relaxed_c128(gpan + gofs[r], v[r]);   /* the next stage may poll it (poll0) */ \
            if (MMH_T2_PRE_EARLY) pre[r] = c_add(c_mul(b0, v[r]), c_mul(a00s, P1[r]));                \
        }                                                                                             \
        gpan += P; xpan += p.hc_max; sqm = st_.x;                                                     \
        if (p.trace && tidc == 0) p.trace[((size_t)tile * S + s_) * 8 + 2] = gtimer_ns();             \
        /* hand-off to step s+1: every compute warp has written panel s, and the halo of panel s has arrived */ \
        if (have_halo && s_ + 3 <= S - 2) bar_arrive(MMH_T2_BAR_FREE + ((s_ - 1) & (NB - 1)), TC + 32); \
        if (xch) {                                                                                    \
            if (MMH_T2_HANDOFF_BAR) bar_sync(MMH_T2_BAR_FULL + (s_ & (NB - 1)), TC + (have_halo ? 32 : 0)); \
            else mbar_wait(sync_base + 8u * (unsigned)(s_ & (NB - 1)), (unsigned)((s_ - 1) >> 2) & 1u); \
        }                                                                                             \
        if (p.trace && tidc == 0) p.trace[((size_t)tile * S + s_) * 8 + 1] = gtimer_ns();             \
    }

    int s = 1;
    double sqm = 0.0;   // sqrt(s - 1)
    if (tl) timeline_stamp(p.timeline, i & 7, 2);
#pragma unroll 1
    for (; s + 1 < S; s += 2) {
        MMH_T2_STEP(h1, h0, s)
        MMH_T2_STEP(h0, h1, s + 1)
    }
    if (s < S) MMH_T2_STEP(h1, h0, s)
#undef MMH_T2_STEP
    if (tl) timeline_stamp(p.timeline, i & 7, 3);
    if (CL) cluster_sync_all();
}

static cudaError_t launch_pdl2(void (*kern)(TiledParams), int grid, int block, size_t smem, cudaStream_t st, bool pdl, int cluster,
                               const TiledParams &p) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    if (cluster > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = (unsigned)cluster; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        na++;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int R, bool CL>
static cudaError_t launch_tiled2_R(const TiledParams &p, int ntiles, size_t smem, cudaStream_t st) {
    const int block = p.tc + 32 * MMH_T2_NHW;
    static size_t smem_set[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };   // opt-in dynamic shared memory already granted (per instantiation)
#define MMH_CASE(N)                                                                                   \
    case N:                                                                                           \
        if (smem > 48 * 1024 && smem > smem_set[N]) {                                                 \
            cudaFuncSetAttribute(k_march_tiled2<R, N, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            smem_set[N] = smem;                                                                       \
        }                                                                                             \
        if (CL) cudaFuncSetAttribute(k_march_tiled2<R, N, CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); \
        return launch_pdl2(k_march_tiled2<R, N, CL>, ntiles, block, smem, st, p.pdl != 0, CL ? ntiles : 0, p);
    const int npd = p.d.D - 1 - p.stage;
    switch (npd) {
        MMH_CASE(1) MMH_CASE(2) MMH_CASE(3)
        default: break;
    }
    if constexpr (R <= 2 && !CL) {
        switch (npd) { MMH_CASE(4) MMH_CASE(5) default: break; }
    }
    if constexpr (R == 1 && !CL) {
        switch (npd) { MMH_CASE(6) MMH_CASE(7) default: break; }
    }
#undef MMH_CASE
    return cudaErrorInvalidValue;
}

// smem bytes of one CTA of k_march_tiled2 (host planner); cluster variant: + the tile's exchange rows [S][hc_max]
size_t mmh_tiled2_smem(int ls_max, int hc_max, int S, int slots) {
    (void)hc_max;
    return sizeof(c128) * ((size_t)MMH_T2_NB * ls_max + (size_t)S + 2) + sizeof(unsigned) * (8 + 3 * (size_t)slots);
}
size_t mmh_tiled2_cluster_extra_smem(int hc_max, int S) { return 16 + sizeof(c128) * (size_t)S * (size_t)hc_max; }

cudaError_t mmh_launch_march_tiled2(const TiledParams &p, int R, int ntiles, size_t smem, cudaStream_t st) {
    if (p.cluster) {
        switch (R) {
            case 1: return launch_tiled2_R<1, true>(p, ntiles, smem, st);
            case 2: return launch_tiled2_R<2, true>(p, ntiles, smem, st);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (R) {
        case 1: return launch_tiled2_R<1, false>(p, ntiles, smem, st);
        case 2: return launch_tiled2_R<2, false>(p, ntiles, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

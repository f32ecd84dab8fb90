// mmh_params.cuh — kernel parameter blocks and host-side launcher prototypes shared by the translation units.
#pragma once
#include "mmh_common.cuh"

// Tuning / test hooks are environment variables named MMH_*.  The entry points take ONE snapshot of them per API call
// (mmh_env_refresh, a single pass over `environ`) and every later lookup reads the snapshot: ~40 getenv() scans per single-lattice
// call were 6-12 us of host time in front of the first kernel launch.
void mmh_env_refresh();
const char *mmh_getenv(const char *name);

struct FwdParams {
    LatticeDesc d;
    const c128 *A;      // [batch, D, D]
    const c128 *b;      // [batch, D]
    const c128 *c;      // [batch]
    c128 *G;            // [batch, N]
    const double *sq;   // sqrt table  sq[n]  = RN(sqrt(n))      (SQRT, vanilla/core.py:22)
    const double *rsq;  // reciprocal  rsq[n] = RN(1 / sq[n]),   rsq[0] unused
    long long batch;
    unsigned *barrier;   // grid-barrier counter (cooperative kernels)
    int small_stage_lo;  // cooperative kernel: stages >= this are filled by CTA 0 alone
};

struct BinomParams {
    LatticeDesc d;
    const c128 *A, *b, *c;
    c128 *G;
    const double *sq, *rsq;
    double max_l2;
    long long global_cutoff;
    double *norm_out;  // device scalar
};

struct VjpParams {
    LatticeDesc d;
    const c128 *G;       // [batch, N]
    const c128 *g;       // [batch, N]  (dLdG)
    const c128 *c;       // [batch]
    c128 *partial;       // [batch, nblk, nacc]
    c128 *dA, *db, *dc;  // outputs [batch, D, D], [batch, D], [batch]
    const double *sq;
    long long batch;
    int nblk;
    int nacc;
    int stride_digits[8];  // mixed-radix digits (radices = shape) of the per-thread walking stride, D <= 8
};

struct StageParams {
    LatticeDesc d;
    const c128 *A, *b;
    c128 *G;
    const double *sq, *rsq;
    long long batch;
    long long lat_stride;        // elements between consecutive lattices in G (= N)
    int stage;                   // i: the index being marched (k_<i = 0)
    int L;                       // lattices marched in lock step by one CTA
    const c128 *c;               // vacuum amplitudes (used when fuse_chain)
    int fuse_chain;              // 1: this launch also computes stage D-1 (the chain) of its lattices first
    int pdl;                     // 1: launched with programmatic stream serialization (waits for the previous launch)
    unsigned long long *timeline;   // debug stamps (mmh_common.cuh timeline_stamp), may be NULL
    long long fill_n;            // k_warp_tail: pre-fill G[0, fill_n) with the all-ones sentinel first (stage overlap), 0 = no
};

struct BoxParams {               // k_march_box (mmh_box.cu): one CTA marches a lattice's stage box by box
    LatticeDesc d;
    const c128 *A, *b;           // [batch, D, D], [batch, D]
    c128 *G;                     // [batch, lat_stride]
    const double *sq, *rsq;
    long long batch, lat_stride;
    int stage;                   // i: the index being marched
    int nt;                      // number of boxed panel dims (1..3): dims stage+1 .. stage+nt
    int g[3];                    // box grid
    int ls;                      // shared-memory cells of one panel buffer: box + halo faces + zero cell + trash cell
};

struct TiledParams {
    LatticeDesc d;
    const c128 *A, *b;           // one triple (device)
    c128 *G;                     // one lattice
    const double *sq, *rsq;
    c128 *X;                     // halo exchange buffer [ntiles][shape[stage]][hc_max], all-sentinel between launches
    int stage;                   // i: the index being marched
    int nt;                      // number of tiled panel dims (1..3): dims stage+1 .. stage+nt
    int g[3];                    // tile grid
    int tc;                      // compute threads (multiple of 32); the CTA adds its halo warps on top
    int ls_max;                  // shared-memory stride of one panel buffer (>= local box size of any tile)
    int hc_max;                  // shared-memory stride of one halo ring slot (>= halo cells of any tile)
    int rows_R;                  // 0: k_march_tiled2 (mmh_tiled.cu); > 0: k_march_rows (mmh_rows.cu) with that many cells per lane
    int rows_C;                  // k_march_rows: lanes per row (chunks of rows_R cells along the last panel dim)
    int rs;                      // k_march_rows: cells of one shared-memory row (>= rows_C * rows_R, chosen bank-conflict free)
    int cells_max;               // k_march_rows: cells of the largest box
    int xc_max;                  // k_march_rows: exported cells of the largest box
    int cluster;                 // k_march_tiled2: 1 = the tile grid is one thread-block cluster, halos pushed through distributed shared memory
    int strong_g;                // k_march_rows: 1 = lattice stores are strong (a later stage polls them), 0 = streaming stores
    int dbg;                     // k_march_rows: tuning switches (MMH_ROWS_DBG), 0 in production
    int pdl;                     // 1: launched with programmatic stream serialization
    int poll0;                   // 1: the previous stage's kernel is still running: do not wait for its completion, validate
                                 //    every panel-0 amplitude by the sentinel the host pre-filled it with (stage overlap)
    int *err;                    // device alias of the context's watchdog word: set to 1 when a poll gives up (mmh_api.cu check_watchdog)
    unsigned long long *trace;   // debug timeline [tile][step][8] of %globaltimer stamps (MMH_TRACE_FILE), else NULL
    unsigned long long *timeline;   // per-launch debug stamps (mmh_common.cuh timeline_stamp), may be NULL
};

// stable rule on a wavefront of boxes (mmh_stable_boxes.cu)
struct StableBoxParams {
    LatticeDesc d;               // 2 <= D <= 4
    const c128 *A, *b, *c;       // one triple
    c128 *G;
    const double *sq, *rsq;
    int nb[4];                   // boxes per dim, right-aligned in four padded dims (leading dims: 1)
    int E;                       // box edge
    int nbox;
    const int *box_order;        // [nbox] box ids sorted by box level (a topological order)
    int *flags;                  // [nbox] 0 -> 1 when the box is complete (zeroed per call)
    int *ticket;                 // next position in box_order (zeroed per call)
    const unsigned *cell_order;  // [E^D] packed local coordinates of a full box's cells sorted by local level
    const int *lvl_start;        // [nlev + 1] offsets of the local levels in cell_order
    int ncell, nlev;             // E^D, D (E - 1) + 1
    int ntab;                    // entries of the sqrt table the kernel copies to shared memory (max shape + 1)
    int xshift;                  // log2 of the extended box edge E + 2
    int *err;                    // watchdog word
    unsigned long long *trace;   // debug: %globaltimer stamps of CTA 0's first 64 boxes (MMH_SB_TRACE), else NULL
};
int mmh_stable_boxes_edge(int D);
size_t mmh_stable_boxes_smem(int D, int ntab, int ncell, int nlev);
cudaError_t mmh_launch_stable_boxes(const StableBoxParams &p, int sm_count, cudaStream_t st);

cudaError_t mmh_launch_march_tiled2(const TiledParams &p, int R, int ntiles, size_t smem, cudaStream_t st);
size_t mmh_tiled2_smem(int ls_max, int hc_max, int S, int slots);
size_t mmh_tiled2_cluster_extra_smem(int hc_max, int S);
cudaError_t mmh_launch_march_rows(const TiledParams &p, int ntiles, size_t smem, cudaStream_t st);
size_t mmh_rows_smem(int ls_max, int hc_max, int S, int CR, int cells_max, int xc_max);
int mmh_rows_max_threads(int R);
bool mmh_rows_supported_R(int R);
cudaError_t mmh_launch_march_stage(const StageParams &p, int R, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t mmh_launch_warp_tail(const StageParams &p, cudaStream_t st);
bool mmh_plan_march_box(const LatticeDesc &d, int stage, BoxParams *bp, int *T_out, size_t *smem_out);
cudaError_t mmh_launch_march_box(const BoxParams &p, int sm_count, int T, size_t smem, cudaStream_t st);
bool mmh_plan_march_lanes(int n1, int *R_out, int *ln_out, int *Lw_out);
int mmh_vjp_blocks_per_sm(const VjpParams &p, int block);
cudaError_t mmh_launch_vjp_lanes(const VjpParams &p, int R, int ln, int Lw, int sm_count, cudaStream_t st);
cudaError_t mmh_launch_march_lanes(const StageParams &p, int R, int ln, int Lw, int sm_count, cudaStream_t st);
cudaError_t mmh_launch_contract_last(const c128 *G, const c128 *cp, c128 *out, long long nout, long long ncore, int nd, cudaStream_t st);
cudaError_t mmh_launch_fill_sentinel(c128 *p, long long n, bool pdl, cudaStream_t st);
cudaError_t mmh_launch_fill_ones(c128 *p, long long n, cudaStream_t st);
cudaError_t mmh_launch_chain(const FwdParams &p, cudaStream_t st);
cudaError_t mmh_launch_fwd_cta(const FwdParams &p, bool stable, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t mmh_coop_max_blocks(bool stable, int block, size_t smem, int *per_sm);
cudaError_t mmh_launch_fwd_coop(const FwdParams &p, bool stable, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t mmh_launch_panel_step(const FwdParams &p, int stage, int s, long long f_lo, long long f_hi, int grid,
                                  size_t smem, cudaStream_t st, bool pdl);
cudaError_t mmh_launch_binomial(const BinomParams &p, int block, size_t smem, cudaStream_t st);
cudaError_t mmh_launch_vjp(const VjpParams &p, int grid_y, int block, cudaStream_t st);
size_t mmh_vjp_planes_smem(const LatticeDesc &d);
cudaError_t mmh_launch_vjp_planes(const VjpParams &p, int sm_count, int *nblk_out, cudaStream_t st);

// compactFock diagonal / one-leftover-mode sweep (mmh_diagonal.cu)
struct DiagParams {
    int Md;              // PNR-detected modes
    int L0;              // 0 = diagonal, 1 = one leftover (undetected) mode occupying A/B indices 0, 1
    int c0;              // cutoff of the leftover mode (1 for the diagonal case)
    int nb;              // batch entries of B (last axis); 1 when B is a vector
    int level;           // weight level being swept
    int cut[8];          // cutoffs of the detected modes
    long long pst[8];    // row-major strides over `cut`
    long long P;         // prod(cut)
    long long E;         // elements of one sub-array: c0 * c0 * P * nb
    const c128 *A;       // [2(Md+L0)]^2, interleaved order [m0,m0,m1,m1,...]
    const c128 *B;       // [2(Md+L0)][nb]
    c128 *arr0;          // [c0][c0][cut...][nb]          (the result)
    c128 *arr1;          // [2 Md] x that
    c128 *arr2;          // [Md] x that
    c128 *arr1010, *arr1001;   // [Md][Md-1] x that
    const double *sq;
};
struct DiagTanParams {
    DiagParams q;        // value arrays (nb == 1, L0 == 0)
    int ntheta;          // 4 M^2 + 2 M tangent directions: A[i,l] row-major, then B[i]
    c128 *t0, *t1, *t2, *t1010, *t1001;   // tangents, same layout as the value arrays with theta innermost
};
cudaError_t mmh_launch_diagonal_tangent(DiagTanParams tp, int nlevels, long long *launches, cudaStream_t st);
cudaError_t mmh_launch_diagonal(DiagParams q, const c128 *G0, int nlevels, long long *launches, cudaStream_t st);

// gate-specific strategies (mmh_gates.cu)
struct GateParams {
    int shape[4];
    int flag;            // displacement: 1 when the host swapped the cutoffs (flipped)
    const double *sq;    // sqrt table
    c128 *out;
    c128 z0;             // squeezer: e^{i theta} tanh r; squeezed: e^{i theta} (-tanh r); beamsplitter: st = sin(theta) e^{i phi};
                         // displacement derivatives: alpha
    double r0, r1, r2;   // squeezer: sech r, sqrt(sech r); beamsplitter: cos(theta); displacement: |alpha|^2, |alpha|, arg(alpha)
};
cudaError_t mmh_launch_squeezer(const GateParams &p, cudaStream_t st);
cudaError_t mmh_launch_squeezed(const GateParams &p, cudaStream_t st);
cudaError_t mmh_launch_beamsplitter(const GateParams &p, bool stable, long long *launches, cudaStream_t st);
cudaError_t mmh_launch_displacement(const GateParams &p, const double *logfac, cudaStream_t st);
cudaError_t mmh_launch_transpose(const c128 *in, c128 *out, int rows, int cols, cudaStream_t st);
cudaError_t mmh_launch_disp_derivs(const GateParams &p, const c128 *D, c128 *o1, c128 *o2, int kind, cudaStream_t st);
cudaError_t mmh_launch_gate_mask(const c128 *g, c128 *out, long long n_total, int kind, int s1, int s2, int s3, cudaStream_t st);
cudaError_t mmh_launch_gate_unsym(const c128 *sym, c128 *out, int D, cudaStream_t st);

// compactFock diagonal sweep with rolling weight-level buffers (mmh_diagonal_rolling.cu)
size_t mmh_diagonal_rolling_workspace(int M, const int *cut, int nb);
cudaError_t mmh_launch_diagonal_rolling(int M, const int *cut, int nb, const c128 *A, const c128 *B, const c128 *G0, c128 *arr0,
                                        const double *sq, void *workspace, long long *launches, cudaStream_t st);

// autoshape (mmh_autoshape.cu)
cudaError_t mmh_launch_autoshape(int M, const c128 *A, const c128 *b, const c128 *c, double max_prob, long long max_shape,
                                 long long min_shape, long long *shape_out, const double *sq, cudaStream_t st);

// Fock-space contraction / reduce (mmh_einsum.cu)
struct EinsumParams {
    const c128 *A, *B;
    c128 *C;
    long long M, N, K, nbatch;
    const long long *offA_b, *offB_b, *offC_b;   // [nbatch]
    const long long *offA_m, *offC_m;            // [M]
    const long long *offB_n, *offC_n;            // [N]
    const long long *offA_k, *offB_k;            // [K]
};
struct ReduceParams {
    int ndim;
    long long in_shape[MMH_MAX_DIM], out_shape[MMH_MAX_DIM], in_stride[MMH_MAX_DIM];
    long long n_out;
    const c128 *in;
    c128 *out;
};
cudaError_t mmh_launch_einsum(const EinsumParams &p, cudaStream_t st);
cudaError_t mmh_launch_fock_reduce(const ReduceParams &p, cudaStream_t st);
cudaError_t mmh_launch_dot(const c128 *x, const c128 *y, long long n, c128 *partial, int nblk, c128 *out, cudaStream_t st);

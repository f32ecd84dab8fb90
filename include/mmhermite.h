/*
 * mmhermite.h — C ABI of libmmhermite.so, the B200 (sm_100a) Gaussian-to-Fock engine.
 *
 * The reference (XanaduAI/MrMustard) has no FFI: its plug-in interface for this path is the backend
 * method table looked up by name in BackendManager._apply (mrmustard/math/backend_manager.py:89-116)
 * plus the `strategies.*` functions the backends call.  Every entry point below replaces one of those
 * Python-level functions; the citation after "replaces:" is the reference file:line whose semantics
 * (argument meaning, layout, in-place `out` contract, error classes) the entry point reproduces.
 * The Python binding a maintainer adds on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - all tensors are C-contiguous complex128, passed as `const void*` / `void*` to interleaved
 *     (re, im) doubles; index arithmetic is int64; `shape` has `ndim` entries, every entry >= 1.
 *   - `*_host` entry points take HOST pointers and perform the H2D/D2H copies themselves (numpy
 *     drop-in).  The others take DEVICE pointers valid on the current CUDA device and enqueue work on
 *     `stream` (a cudaStream_t passed as void*; NULL = the legacy default stream) without synchronising.
 *   - return value: 0 = OK; negative = invalid-argument class (the Python shim re-raises the
 *     reference's exception type, see mmh_error_string); positive = cudaError_t of the failing call.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns an error.
 *   - streams: calls on one stream are ordered by the stream.  The library keeps ONE set of per-device scratch buffers, so a call on
 *     a different stream than the previous call first waits (on the device) for the work enqueued so far on the previous stream.
 *     Entry points that stage a host-built table synchronise `stream` once before returning: mmh_binomial (returns the norm),
 *     mmh_displacement, mmh_fock_contract, and mmh_diagonal when it takes the rolling-level path; they cannot be captured into a
 *     CUDA graph, all the others can -- after one un-captured call with the same arguments (scratch buffers, tables and launch
 *     plans are created on first use; nothing is allocated afterwards).  During capture the cross-stream wait is skipped.
 */
#ifndef MMHERMITE_H
#define MMHERMITE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMH_OK 0
#define MMH_ERR_BAD_NDIM (-1)      /* ndim < 1 or > MMH_MAX_DIM                                  -> ValueError */
#define MMH_ERR_BAD_SHAPE (-2)     /* an entry of shape < 1, or a cutoff beyond the sqrt table   -> ValueError */
#define MMH_ERR_NULL_POINTER (-3)  /* a required pointer is NULL                                 -> ValueError */
#define MMH_ERR_BAD_BATCH (-4)     /* batch < 0                                                  -> ValueError */
#define MMH_ERR_UNSUPPORTED (-5)   /* combination not implemented by the CUDA path               -> NotImplementedError */
#define MMH_ERR_NO_DEVICE (-6)     /* no CUDA device / wrong architecture                        -> RuntimeError */
#define MMH_ERR_TOO_LARGE (-7)     /* lattice does not fit the index type / device memory        -> MemoryError */
#define MMH_ERR_TIMEOUT (-8)       /* a device-side watchdog of an EARLIER call on this device expired (its result is
                                      invalid); the library state has been reset, the call may be retried -> RuntimeError */

#define MMH_MAX_DIM 32

/* library / device management ------------------------------------------------------------------ */
int mmh_version(void);
const char *mmh_error_string(int status);
int mmh_device_count(int *count_out);
int mmh_set_device(int device);            /* cudaSetDevice for the calling thread */
int mmh_device_synchronize(void);
/* number of kernel launches enqueued by this library since load (for bench.py's gpu_launches) */
int64_t mmh_launch_count(void);

/* pinned host memory for the numpy drop-in (results are handed to numpy without a pageable bounce) */
int mmh_host_alloc(void **ptr_out, int64_t bytes);
int mmh_host_free(void *ptr);

/* forward: Bargmann triple -> Fock lattice -------------------------------------------------------
 * replaces: strategies.vanilla_numba (stable=0) / strategies.stable_numba (stable=1)
 *           mrmustard/math/lattice/strategies/vanilla/core.py:25-124 and :127-213,
 *           reached through BackendNumpy.hermite_renormalized (math/backend_numpy.py:381-392).
 * A[ndim,ndim], b[ndim], c[1] -> G[prod(shape)].  Every element of G is overwritten (the reference's
 * `out` contract, core.py:73).  Pivot rule, term order and per-element IEEE arithmetic are those of the
 * reference, so the result is bit-identical to it (tests/test_gpu_forward.py).                      */
int mmh_forward(int ndim, const int64_t *shape, const void *dA, const void *db, const void *dc,
                void *dG, int stable, void *stream);
int mmh_forward_host(int ndim, const int64_t *shape, const void *A, const void *b, const void *c,
                     void *G, int stable);

/* replaces: strategies.vanilla_batch_numba  (vanilla/batch.py:27-61), reached through
 *           BackendNumpy.hermite_renormalized_batched (math/backend_numpy.py:394-403).
 * A[batch,ndim,ndim], b[batch,ndim], c[batch] -> G[batch,prod(shape)] (batch on the first axis).    */
int mmh_forward_batched(int64_t batch, int ndim, const int64_t *shape, const void *dA, const void *db,
                        const void *dc, void *dG, int stable, void *stream);
int mmh_forward_batched_host(int64_t batch, int ndim, const int64_t *shape, const void *A,
                             const void *b, const void *c, void *G, int stable);

/* lattice followed by the contraction over its trailing "derived" axes ----------
 * replaces: the pair  G = math.hermite_renormalized(A, b, ones, shape + shape_derived_vars);
 *                     ret = einsum("...abc..k,...k->...abc..", G.reshape(.., -1), c.reshape(.., -1))
 *           of CircuitComponent.fock_array (lab/circuit_components.py:516-530) and of
 *           PolyExpAnsatz.decompose_ansatz (physics/ansatz/polyexp_ansatz.py:447-462).
 * shape has ndim entries; its first ncore_dims entries are the core (Fock) axes, the rest the derived axes.
 * A[batch,ndim,ndim], b[batch,ndim], cpoly[batch, prod(derived)] -> out[batch, prod(core)].
 * The lattice is computed with vacuum amplitude 1 and never leaves the device.                                   */
int mmh_forward_contract(int64_t batch, int ndim, const int64_t *shape, int ncore_dims, const void *dA, const void *db,
                         const void *dcpoly, void *dout, int stable, void *stream);
int mmh_forward_contract_host(int64_t batch, int ndim, const int64_t *shape, int ncore_dims, const void *A, const void *b,
                              const void *cpoly, void *out, int stable);

/* One panel step of the vanilla fill restricted to a contiguous range of panel offsets: every amplitude
 * G[k] with k_<stage = 0, k_stage = step and panel offset f in [f_lo, f_hi) (f = the flat index of
 * (k_{stage+1}, ..., k_{ndim-1})), from panels step-1 and step-2 of the same lattice (vanilla/core.py:108-122 restricted
 * to a slab).  This is the unit of work of the multi-GPU decomposition of ONE large lattice
 * (mrmustard_b200/sharding.py:forward_single_sharded): ranks own ranges of f and exchange the last strides[stage+1]
 * amplitudes of their range after every step.  G is the whole lattice (device pointer); no synchronisation.      */
int mmh_forward_panel_range(int ndim, const int64_t *shape, const void *dA, const void *db, void *dG, int stage,
                            int64_t step, int64_t f_lo, int64_t f_hi, void *stream);

/* vector-Jacobian product ------------------------------------------------------------------------
 * replaces: strategies.vanilla_vjp_numba (vanilla/gradients.py:25-82), called from the jax backend's
 *           custom_vjp bwd (math/jax_vjps/hermite.py:86-102).
 * G[prod(shape)], c[1], dLdG[prod(shape)] -> dLdA[ndim,ndim] (symmetrised), dLdb[ndim], dLdc[1].
 * Holomorphic cotangent convention (no conjugation), as in the reference.                           */
int mmh_vjp(int ndim, const int64_t *shape, const void *dG, const void *dc, const void *ddLdG,
            void *dLdA_out, void *dLdb_out, void *dLdc_out, void *stream);
int mmh_vjp_host(int ndim, const int64_t *shape, const void *G, const void *c, const void *dLdG,
                 void *dLdA_out, void *dLdb_out, void *dLdc_out);

/* replaces: strategies.vanilla_batch_vjp_numba (vanilla/gradients.py:85-116;
 *           math/jax_vjps/hermite.py:146-175).  Per-triple gradients, no reduction across the batch:
 * G[batch,N], c[batch], dLdG[batch,N] -> dLdA[batch,ndim,ndim], dLdb[batch,ndim], dLdc[batch].      */
int mmh_vjp_batched(int64_t batch, int ndim, const int64_t *shape, const void *dG, const void *dc,
                    const void *ddLdG, void *dLdA_out, void *dLdb_out, void *dLdc_out, void *stream);
int mmh_vjp_batched_host(int64_t batch, int ndim, const int64_t *shape, const void *G, const void *c,
                         const void *dLdG, void *dLdA_out, void *dLdb_out, void *dLdc_out);

/* binomial (fill by total photon number with early stop) -------------------------------------------
 * replaces: strategies.binomial (strategies/binomial.py:30-72) with steps.binomial_step
 *           (lattice/steps.py:208-235) and paths.binomial_subspace_basis (lattice/paths.py:24-72),
 *           reached through BackendNumpy.hermite_renormalized_binomial (math/backend_numpy.py:405-421).
 * Levels |k| = 1 .. global_cutoff-1 are filled in order; after each level norm += sum |G_k|^2 and the
 * fill stops once norm > max_l2.  Untouched entries of G are zero.  norm_out (HOST pointer) receives
 * the accumulated norm; the call synchronises `stream` before returning.                            */
int mmh_binomial(int ndim, const int64_t *shape, const void *dA, const void *db, const void *dc,
                 double max_l2, int64_t global_cutoff, void *dG, double *norm_out, void *stream);
int mmh_binomial_host(int ndim, const int64_t *shape, const void *A, const void *b, const void *c,
                      double max_l2, int64_t global_cutoff, void *G, double *norm_out);

/* compactFock "diagonal": PNR detection amplitudes of all modes -----------------------------------------
 * replaces: hermite_multidimensional_diagonal(A, B, G0, cutoffs)[0]
 *           (strategies/compactFock/inputValidation.py:61-79 -> diagonal_amps.py:145-248), reached through
 *           BackendNumpy.hermite_renormalized_diagonal (math/backend_numpy.py:423-432) after reorder_AB_bargmann.
 * A[2M,2M], B[2M] (nbatch = 0) or B[2M, nbatch] (batch on the LAST axis), both in the interleaved order
 * [m0,m0,m1,m1,...]; G0[1]; cutoffs[M]  ->  arr0[cutoffs...(, nbatch)] = G[a,a,b,b,...].
 * M <= 8.  Auxiliary arrays (arr1, arr2, arr1010, arr1001) live in library scratch.                   */
int mmh_diagonal(int M, const int64_t *cutoffs, const void *dA, const void *dB, int64_t nbatch,
                 const void *dG0, void *darr0_out, void *stream);
int mmh_diagonal_host(int M, const int64_t *cutoffs, const void *A, const void *B, int64_t nbatch,
                      const void *G0, void *arr0_out);

/* Jacobians of the diagonal amplitudes (forward mode) ------------------------------------------------------
 * replaces: grad_hermite_multidimensional_diagonal(A, B, G0, arr0, arr2, arr1010, arr1001, arr1)
 *           (compactFock/inputValidation.py:82-100 -> diagonal_grad.py:19-354), called from the jax bwd rule
 *           (math/jax_vjps/hermite.py:292-329).  The forward arrays are recomputed on the device.
 * A[2M,2M], B[2M] interleaved, G0[1], cutoffs[M] ->
 *   arr0_dA[cutoffs..., 2M, 2M] (entries of A treated as independent), arr0_dB[cutoffs..., 2M],
 *   arr0_dG0[cutoffs...] = arr0 / G0.  M <= 8.                                                             */
int mmh_diagonal_grad(int M, const int64_t *cutoffs, const void *dA, const void *dB, const void *dG0,
                      void *darr0_dG0_out, void *darr0_dA_out, void *darr0_dB_out, void *stream);
int mmh_diagonal_grad_host(int M, const int64_t *cutoffs, const void *A, const void *B, const void *G0,
                           void *arr0_dG0_out, void *arr0_dA_out, void *arr0_dB_out);

/* replaces: grad_hermite_multidimensional_1leftoverMode (inputValidation.py:125-142 -> singleLeftoverMode_grad.py:19-724;
 *           jax bwd math/jax_vjps/hermite.py:397-439).  cutoffs = (c0, tail...), M >= 2 ->
 *   arr0_dG0[c0,c0,tail...], arr0_dA[c0,c0,tail...,2M,2M], arr0_dB[c0,c0,tail...,2M].                              */
int mmh_1leftover_grad(int M, const int64_t *cutoffs, const void *dA, const void *dB, const void *dG0,
                       void *darr0_dG0_out, void *darr0_dA_out, void *darr0_dB_out, void *stream);
int mmh_1leftover_grad_host(int M, const int64_t *cutoffs, const void *A, const void *B, const void *G0,
                            void *arr0_dG0_out, void *arr0_dA_out, void *arr0_dB_out);

/* compactFock "one leftover mode": density matrix of mode 0 conditioned on PNR outcomes of the others ----
 * replaces: hermite_multidimensional_1leftoverMode(A, B, G0, cutoffs)[0]
 *           (inputValidation.py:103-122 -> singleLeftoverMode_amps.py:290-475); the numpy backend computes
 *           the same quantity with strategies.fast_diagonal (fast_diagonal.py:32-77; backend_numpy.py:434-446).
 * A[2M,2M], B[2M] interleaved (indices 0,1 = the undetected mode); cutoffs[M] = (c0, tail...), M >= 2
 * -> arr0[c0, c0, tail...].                                                                              */
int mmh_1leftover(int M, const int64_t *cutoffs, const void *dA, const void *dB, const void *dG0,
                  void *darr0_out, void *stream);
int mmh_1leftover_host(int M, const int64_t *cutoffs, const void *A, const void *B, const void *G0,
                       void *arr0_out);

/* gate-specific Fock strategies -------------------------------------------------------------
 * What Dgate / Sgate / BSgate / SqueezedVacuum.fock_array call instead of the generic lattice (BackendNumpy.displacement /
 * beamsplitter / squeezed / squeezer, math/backend_numpy.py:452-475).  Device-pointer entry points enqueue on `stream`.
 *
 * replaces: strategies.squeezer(shape=(M, N), r, theta)   (strategies/squeezer.py:29-66)    -> S[M, N]
 *           strategies.squeezed(cutoff, r, theta)         (squeezer.py:127-147)             -> S[cutoff]
 *           strategies.beamsplitter(shape4, theta, phi)   (strategies/beamsplitter.py:37-91; stable = 0)
 *           strategies.stable_beamsplitter(shape4, ...)   (beamsplitter.py:94-172;          stable = 1) -> G[M, N, P, Q]
 *           strategies.displacement(cutoffs=(c0, c1), alpha) (strategies/displacement.py:24-65) -> D[c0, c1]
 * The squeezer / squeezed / beamsplitter results are bit-identical to the numba strategies (same IEEE operations in the same
 * order, libm scalars); the displacement evaluates log / exp per element and agrees within 1e-10 relative.                 */
int mmh_squeezer(int64_t M, int64_t N, double r, double theta, void *dS, void *stream);
int mmh_squeezed(int64_t cutoff, double r, double theta, void *dS, void *stream);
int mmh_beamsplitter(const int64_t *shape4, double theta, double phi, int stable, void *dG, void *stream);
int mmh_displacement(int64_t c0, int64_t c1, double alpha_re, double alpha_im, void *dD, void *stream);
/* what = 0 squeezer(shape2, r, theta) | 1 squeezed(shape1, r, theta) | 2 beamsplitter(shape4, theta, phi) |
 *        3 stable_beamsplitter(shape4, theta, phi) | 4 displacement(shape2, Re alpha, Im alpha); HOST output pointer             */
int mmh_gate_host(int what, const int64_t *shape, double a0, double a1, void *out);

/* replaces: strategies.jacobian_displacement(D[M, N], alpha) -> (dD/dalpha, dD/dconj(alpha))     (displacement.py:117-139)
 *           strategies.grad_displacement(T[c, c], r, phi)    -> (dT/dr, dT/dphi)                 (displacement.py:85-114)       */
int mmh_displacement_jacobian(int64_t M, int64_t N, const void *dD, double alpha_re, double alpha_im, void *d_jac_alpha,
                              void *d_jac_alphac, void *stream);
int mmh_displacement_grad(int64_t cutoff, const void *dT, double r, double phi, void *d_grad_r, void *d_grad_phi, void *stream);
/* what = 0 jacobian (a0, a1 = Re alpha, Im alpha) | 1 grad (a0, a1 = r, phi); HOST pointers                                     */
int mmh_displacement_derivs_host(int what, int64_t M, int64_t N, const void *D, double a0, double a1, void *o1, void *o2);

/* replaces: the lattice reductions of strategies.beamsplitter_vjp (kind 0, ndim 4; beamsplitter.py:175-243), squeezer_vjp (kind 1,
 *           ndim 2; squeezer.py:69-124) and squeezed_vjp (kind 2, ndim 1; squeezer.py:150-191): the sums over the gate's support
 *           of dLdG[k] * vanilla_step_grad(G, k) (lattice/steps.py:145-171).
 * out[ndim*ndim + ndim + 1] = un-symmetrised upper-triangular dLdA | dLdb | sum(G * dLdG); the chain rule to (theta, phi) /
 * (r, phi) is a handful of scalar operations done by the caller (mrmustard_b200/strategies.py).                               */
int mmh_gate_vjp(int kind, int ndim, const int64_t *shape, const void *dG, const void *ddLdG, void *dout, void *stream);
int mmh_gate_vjp_host(int kind, int ndim, const int64_t *shape, const void *G, const void *dLdG, void *out);

/* Fock-shape estimate of a Gaussian density matrix -------------------------------------------
 * replaces: autoshape_numba(A, b, c, max_prob, max_shape, min_shape)   (mrmustard/math/lattice/autoshape.py:24-154), called by
 *           State.auto_shape (lab/states/base.py:383-444) to choose the shapes of fock_array / to_fock.
 * A[2M,2M], b[2M] in bargmann order [m0.. | m0..], c[1]  ->  shape_out[M] (int64): for every mode the number of diagonal entries
 * of its single-mode marginal needed to capture max_prob of the trace, clipped to [min_shape, max_shape].  M <= 32.            */
int mmh_autoshape(int M, const void *dA, const void *db, const void *dc, double max_prob, int64_t max_shape, int64_t min_shape,
                  int64_t *dshape_out, void *stream);
int mmh_autoshape_host(int M, const void *A, const void *b, const void *c, double max_prob, int64_t max_shape, int64_t min_shape,
                       int64_t *shape_out);

/* Fock-space consumers of the lattice ---------------------------------------------------------
 * replaces: ArrayAnsatz.contract(other, idx1, idx2, idx_out)   (mrmustard/physics/ansatz/array_ansatz.py:159-225): the einsum of
 *           two Fock arrays by labels, with every label shared by both operands running over the common minimum of its two dims
 *           (:209-218) -- the contraction CircuitComponent.contract performs on `to_fock` outputs
 *           (lab/circuit_components.py:446-458, physics/mm_einsum.py:51-165).
 * labels are small non-negative integers (< 128), one per axis, no label twice inside one operand.  Labels missing from labelsOut
 * are summed.  C is written C-contiguous in the order of labelsOut; out_shape (HOST pointer, nOut entries, may be NULL) receives
 * its shape.  The call synchronises `stream` once (it stages its index tables).                                              */
int mmh_fock_contract(int nA, const int64_t *shapeA, const int *labelsA, int nB, const int64_t *shapeB, const int *labelsB, int nOut,
                      const int *labelsOut, const void *dA, const void *dB, void *dC, int64_t *out_shape, void *stream);
int mmh_fock_contract_host(int nA, const int64_t *shapeA, const int *labelsA, int nB, const int64_t *shapeB, const int *labelsB, int nOut,
                           const int *labelsOut, const void *A, const void *B, void *C);
/* replaces: ArrayAnsatz.reduce(shape) (array_ansatz.py:227-267; mm_einsum.to_fock, physics/mm_einsum.py:272-292): every axis is
 *           sliced to out_shape[d] or zero-padded up to it (batch axes are passed with equal in/out extents).                */
int mmh_fock_reduce(int ndim, const int64_t *in_shape, const int64_t *out_shape, const void *din, void *dout, void *stream);

/* Bilinear overlap of two device lattices: out[0] = sum_k x[k] * y[k] (no conjugation; pass conj(target) for <target|G>).
 * The scalar a fidelity cost takes from the lattice (BASELINE config 5: 1 - |<target|psi>|^2 inside Optimizer.minimize,
 * mrmustard/training/optimizer.py:82-105; physics/fock_utils fidelity), reduced on the device so that a training step reads
 * back a handful of numbers instead of the lattice.  Deterministic (fixed-order reduction).                                  */
int mmh_overlap(int64_t n, const void *dx, const void *dy, void *dout, void *stream);

/* debug aid, not part of the reference interface: 4 (lattice index mod 4) x 16 x 4 %globaltimer stamps (entry, dependency wait passed, first step,
 * exit of CTA 0) of the last single-lattice forward's kernels; synchronises the device.                          */
int mmh_debug_timeline(unsigned long long *out64);

/* debug aid, host only (no device needed): the launch plans of the batched kernels, for the CPU tests of the planners.
 * what = 0: lane layout of a lattice row of n1 = shape[ndim-1] positions (mmh_lanes.cu)  -> out = {R, ln, Lw}
 * what = 1: box grid of stage `stage` (mmh_box.cu)                                       -> out = {g0, g1, g2, threads, nt, ls}
 * what = 2: single-lattice march of stage `stage` planned for 148 SMs (mmh_rows.cu / mmh_tiled.cu)
 *           -> out = {1 = row-lane march | 0 = tiled march, g0, g1, g2, R * 1000 + C | R, row stride | compute threads}
 * what = 3: stable-rule box wavefront (mmh_stable_boxes.cu) -> out = {box edge, boxes per dim (four, right-aligned), -}
 * returns MMH_ERR_UNSUPPORTED when the kernel does not take the shape (the caller then uses another kernel).          */
int mmh_debug_plan(int what, int ndim, const int64_t *shape, int stage, int *out6);

#ifdef __cplusplus
}
#endif
#endif /* MMHERMITE_H */

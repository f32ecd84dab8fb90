// mmh_common.cuh — shared device helpers for the Gaussian-to-Fock kernels (sm_100a).
//
// Arithmetic contract (DESIGN.md "Numerics"): the forward kernels reproduce the reference's
// per-element IEEE-754 double arithmetic exactly (vanilla/core.py:97-104): every product and sum is
// rounded separately (explicit __dmul_rn/__dadd_rn so that ptxas can never contract them into DFMA),
// terms are accumulated in the reference's order, and the final division by sqrt(k_i) is a correctly
// rounded IEEE division.  The *schedule* (which amplitude is computed when, and by which thread) is
// free, because every amplitude only depends on amplitudes that are already final.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mmhermite.h"

typedef double2 c128;  // (x = re, y = im), 16-byte aligned

struct LatticeDesc {
    int D;
    int pad_;
    long long N;                        // prod(shape)
    int shape[MMH_MAX_DIM];
    long long strides[MMH_MAX_DIM];     // row-major element strides, strides[D-1] = 1
};

// Debug timeline (always on, four stamps per launch by one thread): tl[slot * 4 + what] = %globaltimer at kernel entry, after the
// programmatic-dependency wait, at the first march step and at exit of CTA 0.  slot = stage index (0..7), 8 = trailing-stage kernel.
// The buffer belongs to the device context; read back with mmh_debug_timeline() (mmh_api.cu), used by scripts/timeline_cfg2.py.
__device__ __forceinline__ void timeline_stamp(unsigned long long *tl, int slot, int what) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (tl) tl[slot * 4 + what] = t;
}

__device__ __forceinline__ c128 c_make(double re, double im) { return make_double2(re, im); }

// complex * complex, exactly as numba lowers it: (ac - bd, ad + bc), four products and two sums, no FMA.
__device__ __forceinline__ c128 c_mul(c128 x, c128 y) {
    return make_double2(__dsub_rn(__dmul_rn(x.x, y.x), __dmul_rn(x.y, y.y)),
                        __dadd_rn(__dmul_rn(x.x, y.y), __dmul_rn(x.y, y.x)));
}
// complex * real (the reference promotes the real to (s, +0) first; identical up to the sign of zero)
__device__ __forceinline__ c128 c_scale(c128 x, double s) {
    return make_double2(__dmul_rn(x.x, s), __dmul_rn(x.y, s));
}
__device__ __forceinline__ c128 c_add(c128 x, c128 y) {
    return make_double2(__dadd_rn(x.x, y.x), __dadd_rn(x.y, y.y));
}

// Correctly rounded x / s given r = RN(1/s), s > 0 normal.
//   q0 = RN(x r);  e0 = x - s q0 (exact, FMA);  q1 = RN(q0 + e0 r)   -> |q1 - x/s| < 1 ulp
//   e1 = x - s q1 (exact);                      q2 = RN(q1 + e1 r)   -> RN(x/s)  (Markstein's theorem)
// The exactness of e needs x away from the subnormal range, and inf/nan/0 need IEEE special-casing,
// so anything outside 2^-900 < |x| < 2^900 takes the plain IEEE division.
// true iff x needs the IEEE slow path: not (2^-899 <= |x| < 2^900) and not zero.  Integer pipe only.
// (+-0 goes through the fast path: the value is exact, only the sign of the zero may differ.)
__device__ __forceinline__ bool div_needs_slow(double x) {
    const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
    const bool inrange = hi - (124u << 20) < ((1923u - 124u) << 20);
    return !inrange && ((hi | (unsigned)__double2loint(x)) != 0u);
}
__device__ __forceinline__ double div_fast(double x, double s, double r) {
    double q = __dmul_rn(x, r);
    double e = __fma_rn(-s, q, x);
    q = __fma_rn(e, r, q);
    e = __fma_rn(-s, q, x);
    return __fma_rn(e, r, q);
}
// x outside [2^-899, 2^900) and non-zero.  Amplitudes far out in a lattice are routinely that small (they decay towards the
// cutoff), and the IEEE division is ~10x the fast sequence, so the two ranges where a power-of-two scaling is exact are taken
// through the fast sequence as well:  RN(x / s) = 2^-600 RN(2^600 x / s)  for 2^-1008 <= |x| < 2^-899  (the quotient stays normal:
// s = sqrt(k) < 2^12),  and  RN(x / s) = 2^200 RN(2^-200 x / s)  for 2^900 <= |x| < inf  (no overflow: s >= 1).  Only subnormal-range
// quotients, inf and nan take the IEEE division.
__device__ __forceinline__ double div_rare(double x, double s, double r) {
    const unsigned ex = ((unsigned)__double2hiint(x) & 0x7fffffffu) >> 20;   // biased exponent
    if (ex >= 1023u - 1008u && ex < 1023u - 899u) return __dmul_rn(div_fast(__dmul_rn(x, 0x1p600), s, r), 0x1p-600);
    if (ex >= 1023u + 900u && ex < 2047u) return __dmul_rn(div_fast(__dmul_rn(x, 0x1p-200), s, r), 0x1p200);
    return __ddiv_rn(x, s);
}
__device__ __forceinline__ double div_by_table(double x, double s, double r) {
    if (!div_needs_slow(x)) return div_fast(x, s, r);
    return div_rare(x, s, r);
}
// one range test for both components, so that the two 5-level quotient chains interleave instead of running one after
// the other behind two data-dependent branches (the 1-D chain is a pure latency loop: 170 -> ~105 cycles per step)
__device__ __forceinline__ c128 c_div_table(c128 v, double s, double r) {
    if (!(div_needs_slow(v.x) | div_needs_slow(v.y))) return make_double2(div_fast(v.x, s, r), div_fast(v.y, s, r));
    return make_double2(div_by_table(v.x, s, r), div_by_table(v.y, s, r));
}
// the same quotient for the latency-bound 1-D chains: the fast sequence is issued unconditionally and the range test resolves
// beside it (the test sits in front of the 5-level chain in c_div_table: ~15 cycles on a ~100-cycle dependent step); the rare
// out-of-range value redoes the division afterwards.  Bit-identical to c_div_table.
__device__ __forceinline__ c128 c_div_table_spec(c128 v, double s, double r) {
    c128 q = make_double2(div_fast(v.x, s, r), div_fast(v.y, s, r));
    if (div_needs_slow(v.x) | div_needs_slow(v.y)) q = make_double2(div_by_table(v.x, s, r), div_by_table(v.y, s, r));
    return q;
}
__device__ __forceinline__ c128 c_div_real(c128 v, double s) {
    return make_double2(__ddiv_rn(v.x, s), __ddiv_rn(v.y, s));
}

// fused complex multiply-accumulate for the VJP reductions (tolerance-gated, not bit-exact): acc += x*y
__device__ __forceinline__ void c_fma(c128 &acc, c128 x, c128 y) {
    acc.x = fma(x.x, y.x, acc.x);
    acc.x = fma(-x.y, y.y, acc.x);
    acc.y = fma(x.x, y.y, acc.y);
    acc.y = fma(x.y, y.x, acc.y);
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_u32(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// CTA-scope release/acquire on shared-memory words (halo warp <-> compute warps hand-off)
__device__ __forceinline__ int ld_acquire_cta_shared(const int *p) {
    int v;
    asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_shared(int *p, int v) {
    asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
// 16-byte asynchronous global -> shared copy, L2 only (.cg): never served from a stale L1 line
__device__ __forceinline__ void cp_async_cg16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- programmatic dependent launch -----------------------------------------------------------------------------------
// A kernel launched with the programmatic-stream-serialization attribute may become resident while its predecessor on the stream
// is still draining; everything it does before pdl_wait() (index decode, sqrt tables owned by the library) overlaps the
// predecessor's tail and the launch latency, and pdl_wait() -- executed before the first read of any caller-provided buffer --
// returns once the predecessor has completed and flushed.  Without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t mmh_launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                        Args... args) {
    cudaLaunchConfig_t cfg;
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
#endif

// mmh_einsum.cu — Fock-space consumers of the lattice, sm_100a:
//   * binary contraction of two complex128 tensors by labels (ArrayAnsatz.contract, physics/ansatz/array_ansatz.py:159-225:
//     einsum over label strings with the contracted dims truncated to their common minimum) -- what CircuitComponent.contract
//     does with the arrays `to_fock` produced when DEFAULT_REPRESENTATION == "Fock" (lab/circuit_components.py:446-458);
//   * reduce: slice / zero-pad the core dims of a Fock array (ArrayAnsatz.reduce, array_ansatz.py:227-267; mm_einsum.to_fock :272-292).
// Keeping these on the device means the lattice never has to cross PCIe before it is consumed.
//
// The contraction is a complex GEMM over index GROUPS: every label is a batch (in both operands and the output), M (only in the
// first operand and the output), N (only in the second and the output) or K label (not in the output: summed).  The host flattens
// each group and precomputes, per flattened index, the element offset into each operand / the output; the kernel is a shared-memory
// tiled ZGEMM whose loads and stores go through those offset tables, so arbitrary axis orders, truncated dims and output
// permutations need no transposed copies.  FP64 on the ordinary pipe (no tensor cores: complex128 has no tensor-core path).
#include "mmh_params.cuh"

#define EIN_TM 64
#define EIN_TN 64
#define EIN_TK 16

__global__ void __launch_bounds__(256) k_einsum_zgemm(EinsumParams p) {
    __shared__ c128 As[EIN_TK][EIN_TM + 1];
    __shared__ c128 Bs[EIN_TK][EIN_TN + 1];
    const long long bt = blockIdx.z;
    const long long m0 = (long long)blockIdx.y * EIN_TM, n0 = (long long)blockIdx.x * EIN_TN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4 x 4 outputs each
    const c128 *A = p.A + p.offA_b[bt];
    const c128 *B = p.B + p.offB_b[bt];
    c128 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = make_double2(0.0, 0.0);
    for (long long k0 = 0; k0 < p.K; k0 += EIN_TK) {
        for (int e = threadIdx.x; e < EIN_TK * EIN_TM; e += 256) {
            const int kk = e / EIN_TM, mm = e - kk * EIN_TM;
            const long long m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < p.M && k < p.K) ? A[p.offA_m[m] + p.offA_k[k]] : make_double2(0.0, 0.0);
        }
        for (int e = threadIdx.x; e < EIN_TK * EIN_TN; e += 256) {
            const int kk = e / EIN_TN, nn = e - kk * EIN_TN;
            const long long n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < p.N && k < p.K) ? B[p.offB_k[k] + p.offB_n[n]] : make_double2(0.0, 0.0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < EIN_TK; kk++) {
            c128 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) c_fma(acc[i][j], a[i], b[j]);
        }
        __syncthreads();
    }
    c128 *C = p.C + p.offC_b[bt];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const long long m = m0 + ty + 16 * i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long long n = n0 + tx + 16 * j;
            if (n < p.N) C[p.offC_m[m] + p.offC_n[n]] = acc[i][j];
        }
    }
}

// reduce: out[idx] = in[idx] where idx is inside the input's shape, else 0 (zero padding)
__global__ void __launch_bounds__(256) k_fock_reduce(ReduceParams p) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.n_out) return;
    long long rem = f, src = 0;
    bool inside = true;
    for (int d = p.ndim - 1; d >= 0; d--) {
        const long long i = rem % p.out_shape[d];
        rem /= p.out_shape[d];
        inside &= i < p.in_shape[d];
        src += i * p.in_stride[d];
    }
    p.out[f] = inside ? p.in[src] : make_double2(0.0, 0.0);
}

cudaError_t mmh_launch_einsum(const EinsumParams &p, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0 || p.nbatch <= 0) return cudaSuccess;
    const dim3 grid((unsigned)((p.N + EIN_TN - 1) / EIN_TN), (unsigned)((p.M + EIN_TM - 1) / EIN_TM), (unsigned)p.nbatch);
    if (grid.y > 65535 || grid.z > 65535) return cudaErrorInvalidValue;
    k_einsum_zgemm<<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t mmh_launch_fock_reduce(const ReduceParams &p, cudaStream_t st) {
    if (p.n_out <= 0) return cudaSuccess;
    k_fock_reduce<<<(unsigned)((p.n_out + 255) / 256), 256, 0, st>>>(p);
    return cudaGetLastError();
}

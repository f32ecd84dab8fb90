// mmh_autoshape.cu — Fock-shape estimate of a Gaussian density matrix, sm_100a.
//
// Reference: mrmustard/math/lattice/autoshape.py:24-154 (autoshape_numba), called by State.auto_shape (lab/states/base.py:383-444)
// to decide the shapes fed to fock_array / to_fock.  For every mode m: (1) the single-mode marginal (A_, b_, c_) of the M-mode
// Bargmann triple, a Schur complement over the other 2M - 2 variables,
//     A_ = A_mm - A_mn Z_A,  b_ = b_m - A_mn z_b,  c_ = c exp(-1/2 b_n . z_b) / sqrt(det(A_nn - X)),  (A_nn - X) [Z_A | z_b] = [A_nm | b_n]
// (the reference forms inv(A_nn - X) with LAPACK; here one Gaussian elimination with partial pivoting solves for the three
// right-hand sides and yields the determinant); (2) the diagonal of that marginal by the two-buffer recurrence of
// autoshape.py:125-152, accumulating |rho_kk| until the captured probability reaches max_prob or k reaches max_shape.
// One CTA per mode (the modes are independent); the elimination is cooperative over the CTA, the recurrence is a single
// dependent chain.  The result is integer-valued: identical to the reference unless the accumulated norm sits within rounding
// of max_prob at a step (tests/test_gpu_autoshape.py pins it to reference goldens).
#include <cuda/std/complex>

#include "mmh_params.cuh"

typedef cuda::std::complex<double> zc;

__device__ __forceinline__ zc to_z(c128 v) { return zc(v.x, v.y); }

// smem: aug[n][n + 3] zc
__global__ void __launch_bounds__(64) k_autoshape(int M, const c128 *A, const c128 *b, const c128 *cptr, double max_prob, long long max_shape,
                                                  long long min_shape, long long *shape_out, const double *sq) {
    extern __shared__ double2 smem_raw[];
    zc *aug = reinterpret_cast<zc *>(smem_raw);
    __shared__ int piv_row;
    __shared__ double2 det_acc;
    const int m = blockIdx.x;
    const int n = 2 * M - 2, ld = n + 3, n2 = 2 * M;
    const int tid = threadIdx.x, nth = blockDim.x;
    // variable v = s * (M - 1) + i'  <->  full index s * M + idx_n[i'],  idx_n = all modes but m
    auto full = [&](int v) { const int s = v / (M - 1), ip = v - s * (M - 1); return s * M + (ip < m ? ip : ip + 1); };
    for (int e = tid; e < n * ld; e += nth) {
        const int r = e / ld, col = e - r * ld;
        const int fr = full(r);
        zc v;
        if (col < n) {
            v = to_z(A[(long long)fr * n2 + full(col)]);
            const int half = M - 1;
            if ((r < half && col == r + half) || (r >= half && col == r - half)) v -= 1.0;   // X = [[0, 1], [1, 0]]
        } else if (col < n + 2) {
            v = to_z(A[(long long)((col - n) * M + m) * n2 + fr]);     // A_nm = A_mn^T:  A_mn[s, v] = A[s M + m, full(v)]
        } else {
            v = to_z(b[fr]);
        }
        aug[e] = v;
    }
    if (tid == 0) det_acc = make_double2(1.0, 0.0);
    __syncthreads();
    // Gaussian elimination with partial pivoting (n <= 62)
    for (int k = 0; k < n; k++) {
        if (tid == 0) {
            int best = k;
            double bm = cuda::std::abs(aug[k * ld + k]);
            for (int r = k + 1; r < n; r++) {
                const double a = cuda::std::abs(aug[r * ld + k]);
                if (a > bm) { bm = a; best = r; }
            }
            piv_row = best;
        }
        __syncthreads();
        const int pr = piv_row;
        if (pr != k) {
            for (int col = tid; col < ld; col += nth) { const zc t = aug[k * ld + col]; aug[k * ld + col] = aug[pr * ld + col]; aug[pr * ld + col] = t; }
        }
        __syncthreads();
        const zc pv = aug[k * ld + k];
        if (tid == 0) {
            zc d = zc(det_acc.x, det_acc.y) * pv;
            if (pr != k) d = -d;
            det_acc = make_double2(d.real(), d.imag());
        }
        for (int e = tid; e < (n - k - 1) * (ld - k - 1); e += nth) {
            const int r = k + 1 + e / (ld - k - 1), col = k + 1 + e % (ld - k - 1);
            aug[r * ld + col] -= aug[r * ld + k] / pv * aug[k * ld + col];
        }
        __syncthreads();
    }
    // back substitution for the three right-hand sides (thread t < 3 owns column n + t)
    if (tid < 3) {
        for (int r = n - 1; r >= 0; r--) {
            zc v = aug[r * ld + n + tid];
            for (int col = r + 1; col < n; col++) v -= aug[r * ld + col] * aug[col * ld + n + tid];
            aug[r * ld + n + tid] = v / aug[r * ld + r];
        }
    }
    __syncthreads();
    if (tid != 0) return;
    // single-mode triple
    zc A_[2][2], b_[2], bn_zb = 0.0;
    for (int s = 0; s < 2; s++) {
        for (int t = 0; t < 2; t++) {
            zc v = to_z(A[(long long)(s * M + m) * n2 + t * M + m]);
            for (int u = 0; u < n; u++) v -= to_z(A[(long long)(s * M + m) * n2 + full(u)]) * aug[u * ld + n + t];
            A_[s][t] = v;
        }
        zc v = to_z(b[s * M + m]);
        for (int u = 0; u < n; u++) v -= to_z(A[(long long)(s * M + m) * n2 + full(u)]) * aug[u * ld + n + 2];
        b_[s] = v;
    }
    for (int u = 0; u < n; u++) bn_zb += to_z(b[full(u)]) * aug[u * ld + n + 2];
    const zc c_ = to_z(cptr[0]) * cuda::std::exp(-0.5 * bn_zb) / cuda::std::sqrt(zc(det_acc.x, det_acc.y));
    // the two rolling buffers around the diagonal (autoshape.py:121-152)
    zc buf2[2][2] = { { 0.0, 0.0 }, { 0.0, 0.0 } }, buf3[2][3] = { { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 } };
    buf3[0][1] = c_;
    double norm = cuda::std::abs(c_);
    long long k = 0;
    while (norm < max_prob && k < max_shape) {
        const int p = (int)(k & 1), q = p ^ 1;
        const double sk = sq[k], sk1 = sq[k + 1], sk2 = sq[k + 2];
        const zc y0 = A_[0][0] * buf2[p][0] + A_[0][1] * buf2[p][1], y1 = A_[1][0] * buf2[p][0] + A_[1][1] * buf2[p][1];
        buf2[q][0] = (b_[0] * buf3[p][1] + y0 * sk) / sk1;
        buf2[q][1] = (b_[1] * buf3[p][1] + y1 * sk) / sk1;
        buf3[q][0] = (b_[0] * buf2[q][0] + A_[0][0] * buf3[p][1] * sk1 + A_[0][1] * buf3[p][0] * sk) / sk2;
        buf3[q][1] = (b_[1] * buf2[q][0] + A_[1][0] * buf3[p][1] * sk1 + A_[1][1] * buf3[p][0] * sk) / sk1;
        buf3[q][2] = (b_[1] * buf2[q][1] + A_[1][0] * buf3[p][2] * sk + A_[1][1] * buf3[p][1] * sk1) / sk2;
        norm += cuda::std::abs(buf3[q][1]);
        k++;
    }
    long long out = k;
    if (out < min_shape) out = min_shape;
    if (out > max_shape) out = max_shape;
    shape_out[m] = out;
}

cudaError_t mmh_launch_autoshape(int M, const c128 *A, const c128 *b, const c128 *c, double max_prob, long long max_shape,
                                 long long min_shape, long long *shape_out, const double *sq, cudaStream_t st) {
    const int n = 2 * M - 2;
    size_t smem = sizeof(double2) * (size_t)(n > 0 ? n * (n + 3) : 1);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_autoshape, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_autoshape<<<M, 64, smem, st>>>(M, A, b, c, max_prob, max_shape, min_shape, shape_out, sq);
    return cudaGetLastError();
}

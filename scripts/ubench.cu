// ubench.cu — latency microbenchmarks that size the single-lattice march (debug aid, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o /tmp/ubench scripts/ubench.cu && /tmp/ubench
// Prints: dependent-issue latency of DADD / DMUL / DFMA, DP throughput per SM, LDS latency, bar.sync cost,
// L2 round trip (ld.relaxed.gpu), and the SM -> L2 -> SM signalling hop (ping-pong between two CTAs).
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)); return t; }

template <int OP>
__global__ void k_dep(double *out, long long *cyc, double a, double b, int n) {
    double x = a;
    long long t0 = clk();
#pragma unroll 16
    for (int i = 0; i < n; i++) {
        if (OP == 0) x = __dadd_rn(x, b);
        if (OP == 1) x = __dmul_rn(x, b);
        if (OP == 2) x = __fma_rn(x, b, a);
    }
    long long t1 = clk();
    if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// throughput: every thread runs ILP independent chains
template <int ILP>
__global__ void k_tput(double *out, long long *cyc, double a, double b, int n) {
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = a + j;
    __syncthreads();
    long long t0 = clk();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = __fma_rn(x[j], b, a);
    }
    __syncthreads();
    long long t1 = clk();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += x[j];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DP/integer mix: per iteration 4 independent DP ops (OP: 0 DADD, 1 DMUL, 2 DFMA) and NI independent integer ops
template <int OP, int NI>
__global__ void k_mix(double *out, long long *cyc, double a, double b, int n, int ia) {
    double x[4];
    int y[8];
#pragma unroll
    for (int j = 0; j < 4; j++) x[j] = a + j;
#pragma unroll
    for (int j = 0; j < 8; j++) y[j] = ia + j + threadIdx.x;
    __syncthreads();
    long long t0 = clk();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (OP == 0) x[j] = __dadd_rn(x[j], b);
            if (OP == 1) x[j] = __dmul_rn(x[j], b);
            if (OP == 2) x[j] = __fma_rn(x[j], b, a);
        }
#pragma unroll
        for (int j = 0; j < NI; j++) y[j & 7] = y[j & 7] * ia + (j + 1);
    }
    __syncthreads();
    long long t1 = clk();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) s += x[j];
#pragma unroll
    for (int j = 0; j < 8; j++) s += y[j];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the 1-D chain recurrence itself (complex, table division), one lane: cycles and wall ns per dependent step
__device__ __forceinline__ double divf(double x, double s, double r) {
    double q = __dmul_rn(x, r);
    double e = __fma_rn(-s, q, x);
    q = __fma_rn(e, r, q);
    e = __fma_rn(-s, q, x);
    return __fma_rn(e, r, q);
}
__global__ void k_chainrec(double *out, long long *cyc, double bx, double by, double ax, double ay, int n) {
    __shared__ double2 tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) { double s = sqrt((double)(i + 1)); tab[i] = make_double2(s, 1.0 / s); }
    __syncthreads();
    double p1x = 1.0, p1y = 0.5, p2x = 0.0, p2y = 0.0;
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    long long t0 = clk();
    for (int s = 1; s < n; s++) {
        const double2 t = tab[s & 1023];
        double vx = __dsub_rn(__dmul_rn(bx, p1x), __dmul_rn(by, p1y)), vy = __dadd_rn(__dmul_rn(bx, p1y), __dmul_rn(by, p1x));
        const double cx = __dmul_rn(ax, t.x), cy = __dmul_rn(ay, t.x);
        vx = __dadd_rn(vx, __dsub_rn(__dmul_rn(cx, p2x), __dmul_rn(cy, p2y)));
        vy = __dadd_rn(vy, __dadd_rn(__dmul_rn(cx, p2y), __dmul_rn(cy, p2x)));
        vx = divf(vx, t.x, t.y); vy = divf(vy, t.x, t.y);
        p2x = p1x; p2y = p1y; p1x = vx; p1y = vy;
        if (fabs(p1x) > 1e100) { p1x *= 1e-100; p1y *= 1e-100; }
    }
    long long t1 = clk();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = (long long)(g1 - g0); }
    out[threadIdx.x] = p1x + p1y;
}

__global__ void k_lds(int *out, long long *cyc, int n) {
    __shared__ int tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = (i * 33 + 7) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clk();
#pragma unroll 8
    for (int i = 0; i < n; i++) p = tab[p];
    long long t1 = clk();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = p;
}

__global__ void k_bar(long long *cyc, int n) {
    __syncthreads();
    long long t0 = clk();
    for (int i = 0; i < n; i++) asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
    long long t1 = clk();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_l2rt(const unsigned long long *buf, unsigned long long *out, long long *cyc, int n) {
    unsigned long long idx = 0, acc = 0;
    long long t0 = clk();
    for (int i = 0; i < n; i++) {
        unsigned long long v;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(buf + idx) : "memory");
        idx = v;   // pointer chase (buf[i] = next index)
        acc += v;
    }
    long long t1 = clk();
    cyc[0] = t1 - t0;
    out[0] = acc;
}

// ping-pong between CTA 0 and CTA 1 through one L2 word each way (relaxed gpu-scope stores + polling loads)
__global__ void k_pingpong(unsigned long long *flag, long long *cyc, int n, int use_cg) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    unsigned long long *mine = flag + 32 * me, *other = flag + 32 * (1 - me);
    long long t0 = clk();
    for (int i = 1; i <= n; i++) {
        if (me == 0) {
            if (use_cg) __stcg(other, (unsigned long long)i);
            else asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(other), "l"((unsigned long long)i) : "memory");
            unsigned long long v;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory"); } while (v < (unsigned long long)i);
        } else {
            unsigned long long v;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory"); } while (v < (unsigned long long)i);
            if (use_cg) __stcg(other, (unsigned long long)i);
            else asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(other), "l"((unsigned long long)i) : "memory");
        }
    }
    long long t1 = clk();
    if (me == 0) cyc[0] = t1 - t0;
}

int main() {
    double *out; long long *cyc; unsigned long long *buf;
    CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&cyc, 4096 * 8)); CK(cudaMalloc(&buf, 1 << 24));
    long long h[4096];
    int clock_khz = 0; cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    printf("SM clock (attr) %.0f MHz\n", clock_khz / 1e3);
    const int n = 4096;
    const char *names[3] = { "DADD", "DMUL", "DFMA" };
    for (int op = 0; op < 3; op++) {
        for (int rep = 0; rep < 2; rep++) {
            if (op == 0) k_dep<0><<<1, 32>>>(out, cyc, 1.0, 1e-9, n);
            if (op == 1) k_dep<1><<<1, 32>>>(out, cyc, 1.0, 1.0000001, n);
            if (op == 2) k_dep<2><<<1, 32>>>(out, cyc, 1.0, 0.5, n);
            CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("%s dependent latency: %.2f cycles\n", names[op], (double)h[0] / n);
    }
    for (int warps = 4; warps <= 32; warps *= 2) {
        k_tput<4><<<148, warps * 32>>>(out, cyc, 1.0, 0.5, 2048); CK(cudaDeviceSynchronize());
        k_tput<4><<<148, warps * 32>>>(out, cyc, 1.0, 0.5, 2048); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("DFMA throughput, %2d warps/SM x ILP4: %.2f lane-ops/clk/SM\n", warps, 2048.0 * 4 * warps * 32 / h[0]);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        k_tput<2><<<148, warps * 32>>>(out, cyc, 1.0, 0.5, 2048); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("DFMA throughput, %2d warps/SM x ILP2: %.2f lane-ops/clk/SM\n", warps, 2048.0 * 2 * warps * 32 / h[0]);
    }
    for (int rep = 0; rep < 3; rep++) {
        k_chainrec<<<1, 32>>>(out, cyc, 0.3, 0.2, -0.4, 0.1, 20000); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
        printf("chain recurrence, one warp: %.1f cycles/step, %.1f ns/step (=> %.0f MHz effective)\n", (double)h[0] / 20000, (double)h[1] / 20000, 1e3 * h[0] / (double)h[1]);
    }
#define MIX(OP, NI) { k_mix<OP, NI><<<148, 512>>>(out, cyc, 1.0, 0.999, 2048, 3); CK(cudaDeviceSynchronize()); \
        CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost)); \
        printf("mix 16 warps/SM: 4 x %s + %2d int ops per iteration: %.1f cycles/iteration/SMSP-warp-set (DP-only floor 32)\n", names[OP], NI, (double)h[0] / 2048); }
    MIX(0, 0) MIX(1, 0) MIX(2, 0) MIX(2, 4) MIX(2, 8) MIX(2, 16) MIX(0, 8) MIX(1, 8)
    k_lds<<<1, 32>>>((int *)out, cyc, n); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("LDS dependent latency: %.2f cycles\n", (double)h[0] / n);
    for (int t = 64; t <= 1024; t *= 2) {
        k_bar<<<1, t>>>(cyc, 1024); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("bar.sync, %4d threads: %.1f cycles\n", t, (double)h[0] / 1024);
    }
    {   // pointer chase over 8 MB (L2 resident, stride 4 KB + 64)
        const int m = 2048;
        static unsigned long long host[1 << 21];
        for (int i = 0; i < m; i++) host[(size_t)i * 520] = (unsigned long long)((i + 1) % m) * 520;
        CK(cudaMemcpy(buf, host, sizeof(unsigned long long) * m * 520, cudaMemcpyHostToDevice));
        k_l2rt<<<1, 1>>>(buf, (unsigned long long *)out, cyc, m); CK(cudaDeviceSynchronize());
        k_l2rt<<<1, 1>>>(buf, (unsigned long long *)out, cyc, m); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("L2 round trip (ld.relaxed.gpu pointer chase): %.1f cycles\n", (double)h[0] / m);
    }
    for (int cg = 0; cg < 2; cg++) {
        CK(cudaMemset(buf, 0, 1024));
        k_pingpong<<<2, 32>>>(buf, cyc, 2000, cg); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("SM->L2->SM ping-pong (%s store): %.1f cycles per one-way hop\n", cg ? "st.cg" : "st.relaxed.gpu", (double)h[0] / 2000 / 2);
    }
    return 0;
}

// Micro-benchmark behind the design numbers of the on-chip Poisson kernel (profiles/ubench_fp64_r2.md): latency of dependent
// fp64 adds / multiplies, issue rate of independent ones per warp, fp64 throughput per SM, __syncthreads cost, and the
// round trip of a flag through L2 between two CTAs.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_chain(double *out, long long *cyc, int iters, double a, double b)
{
    double x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = a + i + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = __dadd_rn(x[i], b);
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
__global__ void k_chain_mul(double *out, long long *cyc, int iters, double a, double b)
{
    double x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = a + i + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = __dadd_rn(__dmul_rn(x[i], b), a);
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_bar(long long *cyc, int iters)
{
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_lds_chain(double *out, long long *cyc, int iters)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 1) & 1023);
    __syncthreads();
    int idx = threadIdx.x;
    double acc = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const double v = sm[idx & 1023];
        acc = __dadd_rn(acc, v);
        idx += 33;
    }
    const long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ping-pong of a flag through L2 between CTA 0 and CTA 1 (co-resident: 2 CTAs, any two SMs)
__global__ void k_pingpong(unsigned long long *flags, long long *cyc, int iters)
{
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x, other = me ^ 1;
    const long long t0 = clock64();
    for (unsigned long long it = 1; it <= (unsigned long long)iters; it++) {
        if (me == 0) {
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flags + 0), "l"(it) : "memory");
            unsigned long long v;
            do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + 32) : "memory"); } while (v < it);
        } else {
            unsigned long long v;
            do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + 0) : "memory"); } while (v < it);
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flags + 32), "l"(it) : "memory");
        }
    }
    const long long t1 = clock64();
    cyc[me] = t1 - t0;
}

// store 64 KB per CTA to global, then release a flag: how long until the release store has retired (fence cost)
__global__ void k_store_release(double *buf, unsigned long long *flags, long long *cyc, int iters, int doubles_per_thread)
{
    double *mine = buf + (size_t)blockIdx.x * blockDim.x * doubles_per_thread;
    long long tot = 0;
    for (int it = 0; it < iters; it++) {
        __syncthreads();
        const long long t0 = clock64();
        for (int i = 0; i < doubles_per_thread; i += 2)
            *reinterpret_cast<double2 *>(mine + ((size_t)i * blockDim.x + 2 * threadIdx.x)) = make_double2(it, i);
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flags + blockIdx.x * 16), "l"((unsigned long long)it) : "memory");
        __syncthreads();
        tot += clock64() - t0;
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = tot;
}

int main()
{
    double *out; long long *cyc; unsigned long long *flags; double *buf;
    cudaMalloc(&out, sizeof(double) * 1024 * 512);
    cudaMalloc(&cyc, sizeof(long long) * 1024);
    cudaMalloc(&flags, sizeof(unsigned long long) * 148 * 16);
    cudaMalloc(&buf, sizeof(double) * 148 * 384 * 64);
    cudaMemset(flags, 0, sizeof(unsigned long long) * 148 * 16);
    long long h[1024];
    const int iters = 2000;
    auto rd = [&](int n) { cudaDeviceSynchronize(); cudaMemcpy(h, cyc, sizeof(long long) * n, cudaMemcpyDeviceToHost); long long m = 0; for (int i = 0; i < n; i++) m = h[i] > m ? h[i] : m; return (double)m; };
#define CHAIN(K, ILP, THREADS)                                                                                   \
    K<ILP><<<1, THREADS>>>(out, cyc, iters, 1.0, 1e-9); K<ILP><<<1, THREADS>>>(out, cyc, iters, 1.0, 1e-9);       \
    printf(#K " ILP=%d threads=%d (1 CTA): %.2f cycles per fp64 instruction per warp\n", ILP, THREADS, rd(1) / iters / ILP);
    CHAIN(k_chain, 1, 32) CHAIN(k_chain, 2, 32) CHAIN(k_chain, 4, 32) CHAIN(k_chain, 8, 32) CHAIN(k_chain, 16, 32)
    CHAIN(k_chain, 8, 128) CHAIN(k_chain, 8, 256) CHAIN(k_chain, 8, 384) CHAIN(k_chain, 8, 512) CHAIN(k_chain, 8, 1024)
    k_chain_mul<1><<<1, 32>>>(out, cyc, iters, 1.0, 0.999);
    printf("dependent DMUL+DADD pair, 1 warp: %.2f cycles per pair\n", rd(1) / iters);
    k_chain_mul<8><<<1, 384>>>(out, cyc, iters, 1.0, 0.999);
    printf("DMUL+DADD ILP=8, 12 warps: %.2f cycles per fp64 instruction per warp (SM throughput = 12 warps x 32 lanes / that)\n", rd(1) / iters / 16);
    for (int th : {32, 128, 352, 384, 1024}) {
        k_bar<<<1, th>>>(cyc, iters);
        printf("__syncthreads, %d threads: %.1f cycles\n", th, rd(1) / iters);
    }
    k_lds_chain<<<1, 32>>>(out, cyc, iters);
    printf("dependent (LDS.64 -> DADD) chain... LDS independent of the chain: %.1f cycles per iteration\n", rd(1) / iters);
    k_pingpong<<<2, 32>>>(flags, cyc, 2000);
    printf("flag ping-pong between two CTAs through L2 (st.release.gpu / ld.acquire.gpu): %.0f cycles per round trip\n", rd(2) / 2000);
    for (int dpt : {16, 32}) {
        k_store_release<<<144, 384>>>(buf, flags, cyc, 200, dpt);
        printf("144 CTAs x 384 threads store %d KB each (st.v2.f64) + barrier + st.release: %.0f cycles\n", 384 * dpt * 8 / 1024, rd(144) / 200);
    }
    k_store_release<<<144, 384>>>(buf, flags, cyc, 200, 0);
    printf("  (same without the stores: %.0f cycles)\n", rd(144) / 200);
    return 0;
}

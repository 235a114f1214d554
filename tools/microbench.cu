// Developer microbenchmark: dependent-issue latencies that bound the parity-mode chains on B200.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double *out, long long *cyc, double a, double b, int n) {
    __shared__ double sm[1024];
    __shared__ int chase[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) { sm[i] = a * i; chase[i] = (i * 37 + 11) & 1023; }
    __syncthreads();
    double x = a;
    long long t0, t1;
    // 1. dependent DADD
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = __dadd_rn(x, b);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 2. dependent DMUL
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = __dmul_rn(x, 1.0000001);
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // 3. dependent DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = __fma_rn(x, 1.0000001, b);
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 4. 4 independent DADD chains (ILP 4)
    double y0 = x, y1 = x + 1, y2 = x + 2, y3 = x + 3;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n; ++i) { y0 = __dadd_rn(y0, b); y1 = __dadd_rn(y1, b); y2 = __dadd_rn(y2, b); y3 = __dadd_rn(y3, b); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    x = y0 + y1 + y2 + y3;
    // 5. LDS pointer chase (int)
    int p = threadIdx.x & 1023;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) p = chase[p];
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 6. LDS.64 + DADD chain: acc += sm[i] (loads independent)
    double acc = 0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n; ++i) acc = __dadd_rn(acc, sm[(i * 4 + (threadIdx.x & 3)) & 1023]);
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 7. dependent DDIV
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < n / 8; ++i) x = __ddiv_rn(x, 1.0000001);
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // 8. __syncthreads
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) __syncthreads();
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // 9. dependent FADD (fp32) for comparison
    float f = (float)a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) f = __fadd_rn(f, (float)b);
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
    // 10. dsqrt
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < n / 8; ++i) x = sqrt(x + 3.0);
    t1 = clock64(); if (threadIdx.x == 0) cyc[9] = t1 - t0;
    out[threadIdx.x] = x + acc + p + f;
}
int main() {
    double *out; long long *cyc, h[16];
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 16);
    const int n = 4096;
    for (int threads : {32, 128}) {
        for (int blocks : {1, 148 * 4}) {
            k_lat<<<blocks, threads>>>(out, cyc, 1.5, 1e-3, n);
            k_lat<<<blocks, threads>>>(out, cyc, 1.5, 1e-3, n);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, 8 * 16, cudaMemcpyDeviceToHost);
            printf("threads=%d blocks=%d (cycles per op, block 0)\n", threads, blocks);
            const char *nm[] = {"DADD dep", "DMUL dep", "DFMA dep", "DADD x4 ILP (per 4)", "LDS chase", "LDS.64+DADD chain", "DDIV dep", "syncthreads", "FADD dep", "DSQRT dep"};
            for (int i = 0; i < 10; ++i) printf("  %-22s %.2f\n", nm[i], (double)h[i] / ((i == 6 || i == 9) ? n / 8 : n));
        }
    }
    return 0;
}

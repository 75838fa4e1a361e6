// Latency microbenchmarks used to calibrate the level-chain model in DESIGN.md (FP64 dependent chains, global loads).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(double* out, long long* cyc, double a, double b, int n) {
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = x / b + 1.0;           // div + add
    long long t1 = clock64();
    double y = a;
    for (int i = 0; i < n; ++i) y = __dadd_rn(__dmul_rn(y, b), 1.0); // mul + add (no fma)
    long long t2 = clock64();
    double z = a;
    for (int i = 0; i < n; ++i) z = sqrt(z) + b;
    long long t3 = clock64();
    double w = a;
    for (int i = 0; i < n; ++i) w = fma(w, b, 1.0);
    long long t4 = clock64();
    double s = a, c = 0;
    for (int i = 0; i < n; ++i) { double sn, cs; sincos(s, &sn, &cs); s = sn + cs; }
    long long t5 = clock64();
    out[threadIdx.x] = x + y + z + w + s + c;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
}
__global__ void chase(int const* next, int start, int n, long long* cyc, int* sink) {
    int p = start;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) p = next[p];
    long long t1 = clock64();
    *sink = p; cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 64);
    int n = 2000;
    for (int threads : {1, 32, 512}) {
        chain<<<1, threads>>>(out, cyc, 1.2345, 1.0000001, n); cudaDeviceSynchronize();
        printf("threads %4d: div+add %.1f  mul+add %.1f  sqrt+add %.1f  fma %.1f  sincos+add %.1f cycles/iter\n", threads,
               (double)cyc[0] / n, (double)cyc[1] / n, (double)cyc[2] / n, (double)cyc[3] / n, (double)cyc[4] / n);
    }
    // pointer chase: stride 4 KB over 1 GB (DRAM, TLB friendly?) and over 8 MB (L2) and 64 KB (L1)
    for (size_t bytes : {(size_t)64 << 10, (size_t)8 << 20, (size_t)1 << 30}) {
        size_t cnt = bytes / 4; int* h = new int[cnt]; size_t stride = 1024 + 32; // ints
        for (size_t i = 0; i < cnt; ++i) h[i] = (int)((i + stride) % cnt);
        int* d; cudaMalloc(&d, bytes); cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice); int* sink; cudaMalloc(&sink, 4);
        int steps = 4000;
        chase<<<1, 1>>>(d, 0, steps, cyc, sink); cudaDeviceSynchronize();
        chase<<<1, 1>>>(d, 7, steps, cyc, sink); cudaDeviceSynchronize();
        printf("pointer chase over %zu KB: %.0f cycles/load\n", bytes >> 10, (double)cyc[0] / steps);
        cudaFree(d); delete[] h;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock %d kHz\n", clk);
    return 0;
}

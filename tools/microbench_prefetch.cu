// What does a load cost after prefetch.global.L1 / prefetch.global.L2 / a discarded load / cp.async staging issued `ahead`
// steps (of ~3200 cycles) earlier?  One warp walks a buffer with a large stride; each step: the early action for step
// i + ahead, a dependent FP64 chain, then a timed load of step i.
//   region: "L2" = the lines were touched by a warm-up pass and stay in L2; "DRAM" = a fresh 8 MB window of a 1 GB buffer
//   timed load: plain ld.global (what the solver's chains issue), ld.global.ca, ld.global.nc (__ldg)
//   early actions 5/6: cp.async.ca into a scratch word that is never read nor waited for - does the line stay in L1?
// Also: latency of re-reading a 16 KB window over and over (L1 hit).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void pf_l1(void const* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void pf_l2(void const* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__global__ void walk(double* data, size_t stride, int n, int ahead, int early, int timed, int delay, long long* cyc, double* sink) {
    extern __shared__ double ring[];
    size_t const lane = threadIdx.x;
    auto at = [&](int i) { return data + ((size_t)i * stride + lane); };
    double x = 1.0;
    long long total = 0, worst = 0;
    if (early == 3) {
        for (int k = 0; k < ahead; ++k) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(ring + (k % 8) * 32 + lane)), "l"(at(k)) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
    for (int i = 0; i < n; ++i) {
        if (early == 1) pf_l1(at(i + ahead));
        if (early == 2) pf_l2(at(i + ahead));
        if (early == 3) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(ring + ((i + ahead) % 8) * 32 + lane)), "l"(at(i + ahead)) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        if (early == 5 || early == 6) { // fire and forget: cp.async into a scratch word nobody reads (is the line left in L1?)
            if (early == 5) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(ring + 8 * 32 + lane)), "l"(at(i + ahead)) : "memory");
            else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(ring + 8 * 32 + lane)), "l"(at(i + ahead)) : "memory");
        }
        if (early == 4) {
            double dummy;
            asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(dummy) : "l"(at(i + ahead)) : "memory");
            x += dummy * 1e-300;
        }
        for (int k = 0; k < delay; ++k) x = fma(x, 1.0000001, 1e-9);
        long long const t0 = clock64();
        double v;
        if (early == 3) {
            if (ahead == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 2;" ::: "memory");
            v = ring[(i % 8) * 32 + lane];
        } else if (timed == 0) {
            v = *at(i);
        } else if (timed == 1) {
            asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(at(i)) : "memory");
        } else {
            v = __ldg(at(i));
        }
        x += v;
        long long const t1 = clock64() + (long long)(x == 12345.678);
        total += t1 - t0;
        worst = max(worst, t1 - t0);
    }
    if (lane == 0) {
        cyc[0] = total;
        cyc[1] = worst;
    }
    sink[lane] = x;
}
__global__ void reread(double const* data, int n, long long* cyc, double* sink) {
    double x = 0.0;
    for (int i = 0; i < 64; ++i) x += data[i * 32 + threadIdx.x]; // 16 KB window
    long long total = 0;
    for (int i = 0; i < n; ++i) {
        long long const t0 = clock64();
        x += data[(i % 64) * 32 + threadIdx.x];
        long long const t1 = clock64() + (long long)(x == 12345.678);
        total += t1 - t0;
    }
    if (threadIdx.x == 0) cyc[0] = total;
    sink[threadIdx.x] = x;
}
int main() {
    long long* cyc;
    cudaMallocManaged(&cyc, 64);
    double* sink;
    cudaMalloc(&sink, 4096);
    int const n = 2000;
    char const* early_names[] = {"nothing", "prefetch.global.L1", "prefetch.global.L2", "cp.async -> shared", "discarded ld.global.ca", "cp.async.ca 8 B to scratch", "cp.async.ca 4 B to scratch"};
    char const* timed_names[] = {"ld.global", "ld.global.ca", "ld.global.nc"};
    size_t const bytes = (size_t)1 << 30;
    double* d;
    cudaMalloc(&d, bytes);
    cudaMemset(d, 0, bytes);
    reread<<<1, 32>>>(d, n, cyc, sink);
    cudaDeviceSynchronize();
    printf("re-reading a 16 KB window: %.0f cycles per load (L1 hit, timing included)\n", (double)cyc[0] / n);
    size_t const stride = 4096 / 8 + 16; // doubles
    int window = 0;
    for (int dram : {0, 1}) {
        for (int early = 0; early < 7; ++early) {
            for (int timed = 0; timed < 3; ++timed) {
                if (early == 3 && timed != 0) continue;
                for (int ahead : {1, 2}) {
                    if (early == 0 && ahead != 1) continue;
                    double* base = d + (size_t)(window++ % 100) * (stride * (n + 8)); // a fresh window of the buffer per run
                    if (!dram) {
                        walk<<<1, 32, 9 * 32 * 8>>>(base, stride, n, ahead, 0, 0, 0, cyc, sink); // warm-up: lines into L2
                        cudaDeviceSynchronize();
                    }
                    walk<<<1, 32, 9 * 32 * 8>>>(base, stride, n, ahead, early, timed, 400, cyc, sink);
                    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
                    printf("%-4s early: %-22s ahead %d  timed: %-12s  %4.0f cycles avg, %5lld worst\n", dram ? "DRAM" : "L2", early_names[early],
                           ahead, early == 3 ? "wait + LDS" : timed_names[timed], (double)cyc[0] / n, cyc[1]);
                }
            }
        }
    }
    return 0;
}

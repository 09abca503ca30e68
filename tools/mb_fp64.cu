// Microbenchmark: FP64 pipe throughput of one B200 (the denominator of bench.py's secondary
// FP64 bound; SURVEY 8(d) asks for a measured figure, not the 64 lanes/clk/SM of the data sheet).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mb_fp64 tools/mb_fp64.cu
// Prints one JSON line; bench.py runs the binary (if built) and falls back to the constant.
// Three instruction mixes, 8 independent chains per thread, 1024 threads x 2 blocks per SM:
//   dfma   a = fma(a, b, c)                       (what a contraction could reach)
//   dmul+dadd  a = a * b; a = a + c  (-fmad=false arithmetic of the exact kernels: two pipe slots
//                                     per multiply-add)
//   ddiv   a = c / a  (IEEE division, as in the bracket coordinate of every point and direction)
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(double* out, int iters, double b, double c) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fma(a[i], b, c);
            else if (MODE == 1) a[i] = __dadd_rn(__dmul_rn(a[i], b), c);
            else a[i] = __ddiv_rn(c, a[i]);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(double* out, int sms, int iters) {
    const int blocks = sms * 2, threads = 1024;
    k<MODE><<<blocks, threads>>>(out, 10, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return 8.0 * iters * blocks * threads / (best * 1e-3);      // operations of the mix per second
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = prop.multiProcessorCount;
    double* out;
    cudaMalloc(&out, (size_t)sms * 2 * 1024 * 8);
    const double fma_s = run<0>(out, sms, 20000);
    const double ma_s = run<1>(out, sms, 20000);
    const double div_s = run<2>(out, sms, 2000);
    const double per_clk = 1.0 / sms / (clk_khz * 1e3);
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_max_khz\": %d, "
           "\"dfma_per_s\": %.4e, \"dfma_tflops\": %.2f, \"dfma_lanes_per_clk_sm\": %.1f, "
           "\"dmul_dadd_pairs_per_s\": %.4e, \"dmul_dadd_lanes_per_clk_sm\": %.1f, "
           "\"ddiv_per_s\": %.4e, \"ddiv_lanes_per_clk_sm\": %.2f, "
           "\"note\": \"per-clock figures use the maximum SM clock; 8 independent chains per thread, "
           "2 x 1024 threads per SM, best of 5\"}\n",
           prop.name, sms, clk_khz, fma_s, 2 * fma_s / 1e12, fma_s * per_clk,
           ma_s, 2 * ma_s * per_clk, div_s, div_s * per_clk);
    return 0;
}

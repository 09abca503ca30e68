// Microbenchmark: L1 data-pipe cost of warp-wide loads by width and coherence.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mb_l1 tools/mb_l1.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int WIDTH, int MODE>   // WIDTH 8|16 bytes; MODE 0 broadcast, 1 random in 648 B, 2 coalesced
__global__ void k(const double* __restrict__ tab, double* out, int iters) {
    const int lane = threadIdx.x & 31;
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    double acc = 0.0;
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        int idx;
        if (MODE == 0) idx = (i * 7) % 40;
        else if (MODE == 1) idx = (s >> 8) % 40;
        else idx = lane;
        if (WIDTH == 8) {
            acc += __ldg(tab + 2 * idx) + __ldg(tab + 2 * idx + 1);      // two 8-byte loads
        } else {
            const double2 v = __ldg(reinterpret_cast<const double2*>(tab) + idx);   // one 16-byte load
            acc += v.x + v.y;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int WIDTH, int MODE>
void run(const char* name, const double* tab, double* out) {
    const int iters = 20000, blocks = 148 * 2, threads = 512;
    k<WIDTH, MODE><<<blocks, threads>>>(tab, out, 100);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<WIDTH, MODE><<<blocks, threads>>>(tab, out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double values = 2.0 * iters * blocks * threads;            // 8-byte values delivered
    printf("%-44s %8.3f ms  %7.2f values/clk/SM (1.965 GHz)\n", name, ms,
           values / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    double *tab, *out;
    cudaMalloc(&tab, 4096); cudaMemset(tab, 0, 4096); cudaMalloc(&out, 148 * 2 * 512 * 8);
    run<8, 0>("2 x LDG.64  broadcast", tab, out);
    run<16, 0>("1 x LDG.128 broadcast", tab, out);
    run<8, 1>("2 x LDG.64  random within 640 B", tab, out);
    run<16, 1>("1 x LDG.128 random within 640 B", tab, out);
    run<8, 2>("2 x LDG.64  coalesced", tab, out);
    run<16, 2>("1 x LDG.128 coalesced", tab, out);
    return 0;
}

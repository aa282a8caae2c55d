// Microbenchmark: fp64 DFMA throughput and dependent-issue latency per SM on this GPU, as a function of resident warps and ILP.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, typename T>
__global__ void k(T* out, int iters, T x) {
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], x, (T)1.0);
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == (T)123456789.0) out[0] = s;
}
template <int ILP, typename T>
void run(const char* name, int warps, int n_sm, double ghz) {
    T* d; cudaMalloc(&d, 64);
    const int iters = 4096;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<ILP, T><<<n_sm, warps * 32>>>(d, iters, (T)1.0000001);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<ILP, T><<<n_sm, warps * 32>>>(d, iters, (T)1.0000001);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double cycles = ms * 1e-3 * ghz * 1e9;
    double lanes_per_clk = (double)warps * 32 * iters * ILP / cycles;
    printf("%s warps=%2d ilp=%d  %.3f ms  lanes/clk/SM=%.1f  cycles per dependent op per warp=%.1f\n", name, warps, ILP, ms, lanes_per_clk, cycles / iters);
    cudaFree(d);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz / 1e6;
    printf("%s SMs=%d clock=%.3f GHz (nominal max; lanes/clk assumes this clock)\n", p.name, p.multiProcessorCount, ghz);
    for (int w : {1, 2, 4, 6, 8, 12, 16, 32}) {
        run<1, double>("f64", w, p.multiProcessorCount, ghz);
        run<4, double>("f64", w, p.multiProcessorCount, ghz);
        run<9, double>("f64", w, p.multiProcessorCount, ghz);
    }
    for (int w : {4, 16, 32}) { run<1, float>("f32", w, p.multiProcessorCount, ghz); run<8, float>("f32", w, p.multiProcessorCount, ghz); }
    return 0;
}

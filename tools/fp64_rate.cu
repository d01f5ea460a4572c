// micro-benchmark: DFMA and FFMA issue rate per SM (how much room the float64 recurrences of the effects chain have)
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void k(T* out, int iters) {
    T a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T m = (T)0.999999, c = (T)1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <typename T>
void run(const char* name) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    T* d; cudaMalloc(&d, sizeof(T) * sms * 4 * 512);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<T><<<sms * 4, 512>>>(d, 100);
    cudaEventRecord(e0);
    k<T><<<sms * 4, 512>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)sms * 4 * 512 * iters * 8;
    printf("%s: %.3f ms, %.1f TFMA/s, %.1f FMA/clk/SM at nominal %d MHz\n", name, ms, fma / ms * 1e-9, fma / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    cudaFree(d);
}
int main() { run<double>("f64"); run<float>("f32"); return 0; }

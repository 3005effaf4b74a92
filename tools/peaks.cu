// peaks.cu -- measures the FP64 / FP32 FMA issue peak and a streaming-read bandwidth of the
// GPU it runs on; the FP64 figure is the denominator of the "FP64 pipe" fractions in
// DESIGN.md / profiles (MEASURED_PEAKS.json records only HBM copy and bf16 GEMM peaks).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks.bin tools/peaks.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <typename T, int CHAINS>
__global__ void fma_chain(T *out, T a, T b, int iters)
{
    T x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = fma(x[i], a, b);
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i];
    if (s == (T)123456789) out[0] = s;
}

__global__ void read_stream(const double2 *__restrict__ p, size_t n, double *out)
{
    double acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double2 v = __ldcs(p + i);
        acc += v.x + v.y;
    }
    if (acc == 1.2345) out[0] = acc;
}

template <typename T>
static double run_fma(int sms, const char *name)
{
    T *out;
    cudaMalloc(&out, 64);
    const int iters = 4096, chains = 8, threads = 512, blocks = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        for (int k = 0; k < 10; ++k) fma_chain<T, chains><<<blocks, threads>>>(out, (T)1.0000001, (T)1e-9, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fmas = 10.0 * (double)blocks * threads * chains * iters;
        double tf = 2.0 * fmas / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    printf("{\"what\": \"%s\", \"tflops\": %.3f, \"gfma_per_s\": %.1f}\n", name, best, best * 1e3 / 2);
    cudaFree(out);
    return best;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
    run_fma<double>(p.multiProcessorCount, "fp64_fma");
    run_fma<float>(p.multiProcessorCount, "fp32_fma");
    size_t n = (size_t)1 << 28;   // 4 GiB of double2
    double2 *buf; double *out;
    cudaMalloc(&buf, n * sizeof(double2)); cudaMalloc(&out, 64);
    cudaMemset(buf, 0, n * sizeof(double2));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        read_stream<<<p.multiProcessorCount * 16, 512>>>(buf, n, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double gbs = (double)n * sizeof(double2) / (ms * 1e-3) / 1e9;
        if (gbs > best) best = gbs;
    }
    printf("{\"what\": \"stream_read_4GiB\", \"gbs\": %.1f}\n", best);
    return 0;
}

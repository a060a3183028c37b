// fp64_peak.cu -- measures the FP64 denominators of the `prepare` roofline on this GPU:
//   (1) vector DFMA throughput (independent FMA chains per thread, all SMs),
//   (2) DMMA m8n8k4 (mma.sync.aligned.m8n8k4.row.col.f64) throughput,
//   (3) shared-memory-fed DFMA (one LDS.64 operand per FMA), the regime of small-matrix kernels.
// MEASURED_PEAKS.json has no FP64 figure, so bench.py reads the JSON this prints
// (profiles/fp64_peak.json is the committed copy).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o benchmarks/_build/fp64_peak benchmarks/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double fa = a + threadIdx.x, fb = b;
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[0]), "+d"(c0[1]) : "d"(fa), "d"(fb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c1[0]), "+d"(c1[1]) : "d"(fa), "d"(fb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c2[0]), "+d"(c2[1]) : "d"(fa), "d"(fb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c3[0]), "+d"(c3[1]) : "d"(fa), "d"(fb));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

__global__ void lds_dfma_kernel(double* out, int iters, double b) {
    __shared__ double s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    const int l = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
        const int o = (i * 4) & 511;
        x0 = fma(x0, s[o + l], b); x1 = fma(x1, s[o + l + 32], b);
        x2 = fma(x2, s[o + l + 64], b); x3 = fma(x3, s[o + l + 96], b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (x0 + x1) + (x2 + x3);
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, threads = 512, blocks = sms * 4, iters = 20000;
    double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    float t1 = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    float t2 = time_ms([&] { dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    float t3 = time_ms([&] { lds_dfma_kernel<<<blocks, threads>>>(out, iters, 1e-9); });
    const double n = (double)blocks * threads;
    const double dfma = n * iters * 8 * 2 / (t1 * 1e-3) / 1e12;
    const double dmma = (n / 32) * iters * 4 * (8 * 8 * 4 * 2) / (t2 * 1e-3) / 1e12;
    const double lds = n * iters * 4 * 2 / (t3 * 1e-3) / 1e12;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.3f, \"dmma_m8n8k4_tflops\": %.3f, \"lds_fed_dfma_tflops\": %.3f, "
           "\"how\": \"8 independent DFMA chains/thread, 4 DMMA chains/warp, 4 LDS.64-fed DFMA chains/thread; %d blocks x %d threads, best of 5, CUDA events\"}\n",
           p.name, sms, dfma, dmma, lds, blocks, threads);
    return cudaGetLastError() != cudaSuccess;
}

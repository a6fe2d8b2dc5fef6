// tools/microbench.cu -- FP64 pipe measurements that size the Gram kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/microbench tools/microbench.cu
// Prints one JSON object: DFMA / DMMA / mixed / exp throughput per chip.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dfma(double *out, int iters, double a, double b)
{
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma(double *out, int iters, double a, double b)
{
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x; c1[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// NMMA DMMAs + NFMA DFMAs per iteration, independent chains
template <int NMMA, int NFMA>
__global__ void k_mixed(double *out, int iters, double a, double b)
{
    double c0[NMMA > 0 ? NMMA : 1], c1[NMMA > 0 ? NMMA : 1], f[NFMA > 0 ? NFMA : 1];
#pragma unroll
    for (int i = 0; i < NMMA; i++) { c0[i] = threadIdx.x; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < NFMA; i++) f[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NMMA; i++) dmma(c0[i], c1[i], a, b);
#pragma unroll
        for (int i = 0; i < NFMA; i++) f[i] = fma(f[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NMMA; i++) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NFMA; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double *out, int iters, double x0)
{
    double x = x0 - 1e-3 * threadIdx.x, s = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            double v = exp(x);
            s += v;
            x -= 1e-4;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_it(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int threads = 256, blocks = sms * 4;   // 1024 threads = 32 warps per SM
    double *out;
    CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    const int iters = 4096;
    const double nthreads = (double)blocks * threads, nwarps = nthreads / 32;

    float t_fma = time_it([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double dfma_tflops = 2.0 * nthreads * 8 * iters / (t_fma * 1e-3) / 1e12;

    float t_mma = time_it([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double dmma_tflops = 2.0 * 256 * nwarps * 8 * iters / (t_mma * 1e-3) / 1e12;

    // same DMMA work plus extra DFMAs: if the pipes are shared the time adds up, if separate it hides
    float t_mix8 = time_it([&] { k_mixed<8, 8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    float t_mix16 = time_it([&] { k_mixed<8, 16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    float t_mix0 = time_it([&] { k_mixed<8, 0><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    float t_f8 = time_it([&] { k_mixed<0, 8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });

    const int eiters = 512;
    float t_exp0 = time_it([&] { k_exp<<<blocks, threads>>>(out, eiters, -0.1); });
    double exp0 = nthreads * 8 * eiters / (t_exp0 * 1e-3) / 1e12;

    // occupancy sensitivity of DMMA: 8 warps per SM (the Gram kernel's shape)
    float t_mma8w = time_it([&] { k_dmma<16><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); });
    double dmma8w = 2.0 * 256 * (sms * 8.0) * 16 * iters / (t_mma8w * 1e-3) / 1e12;
    float t_mma16w = time_it([&] { k_dmma<16><<<sms, 512>>>(out, iters, 1.0000001, 1e-9); });
    double dmma16w = 2.0 * 256 * (sms * 16.0) * 16 * iters / (t_mma16w * 1e-3) / 1e12;
    float t_mma4w = time_it([&] { k_dmma<16><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); });
    double dmma4w = 2.0 * 256 * (sms * 4.0) * 16 * iters / (t_mma4w * 1e-3) / 1e12;

    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, "
           "\"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, "
           "\"ms_dmma8\": %.4f, \"ms_dmma8_plus_dfma8\": %.4f, \"ms_dmma8_plus_dfma16\": %.4f, \"ms_dfma8\": %.4f, "
           "\"exp_libdevice_Tops\": %.3f, "
           "\"dmma_tflops_4warps_per_sm\": %.2f, \"dmma_tflops_8warps_per_sm\": %.2f, \"dmma_tflops_16warps_per_sm\": %.2f}\n",
           prop.name, sms, prop.clockRate, dfma_tflops, dmma_tflops, t_mix0, t_mix8, t_mix16, t_f8, exp0,
           dmma4w, dmma8w, dmma16w);
    cudaFree(out);
    return 0;
}

// cvt_rate.cu -- issue cost of int -> double conversions against DADD / DFMA on sm_100a (one CTA of 256 threads per SM,
// 8 independent chains per thread).  Prints SM-clock cycles per warp-instruction per scheduler.
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, double *out, long long *cyc, long long seed)
{
    double a[8];
    long long q[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-3 + i; q[i] = seed + threadIdx.x * 977 + i * 131071; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = a[i] + 1.000001;                                   // DADD
            if (MODE == 1) a[i] = fma(a[i], 1.0000001, 0.5);                         // DFMA
            if (MODE == 2) { a[i] += (double)q[i]; q[i] += it; }                     // I2F.F64.S64 + DADD + int add
            if (MODE == 3) { a[i] += (double)(int)q[i]; q[i] += it; }                // I2F.F64.S32 + DADD
            if (MODE == 4) {                                                         // magic: |q| < 2^51
                a[i] += __longlong_as_double(q[i] + 0x4338000000000000ll) - 6755399441055744.0; q[i] += it; }
            if (MODE == 5) { q[i] += it; a[i] += 1.0; }                              // baseline of modes 2..4: int add + DADD
        }
    }
    const long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; i++) s += a[i] + (double)q[i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    double *out; long long *cyc; CK(cudaMalloc(&out, 148 * 256 * 8)); CK(cudaMalloc(&cyc, 148 * 8));
    const int iters = 20000;
    const char *names[] = {"DADD", "DFMA", "I2F.F64.S64 + DADD", "I2F.F64.S32 + DADD", "magic (IADD64 + DADD) + DADD", "int add + DADD"};
    for (int m = 0; m < 6; m++) {
        for (int rep = 0; rep < 2; rep++) {
            switch (m) {
            case 0: k<0><<<148, 256>>>(iters, out, cyc, 12345); break;
            case 1: k<1><<<148, 256>>>(iters, out, cyc, 12345); break;
            case 2: k<2><<<148, 256>>>(iters, out, cyc, 12345); break;
            case 3: k<3><<<148, 256>>>(iters, out, cyc, 12345); break;
            case 4: k<4><<<148, 256>>>(iters, out, cyc, 12345); break;
            case 5: k<5><<<148, 256>>>(iters, out, cyc, 12345); break;
            }
            CK(cudaDeviceSynchronize());
        }
        long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        // 8 warps / 4 schedulers = 2 warps per scheduler, 8 statements per iteration
        printf("{\"op\": \"%s\", \"clk_per_statement_per_scheduler\": %.2f}\n", names[m], (double)c / iters / 8 / 2);
    }
    return 0;
}

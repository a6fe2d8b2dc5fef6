// Prototype (NOT product code): FP64-equivalent Gram G = K^T K on the INT8 tensor cores of sm_100a
// (tcgen05.mma.kind::i8, accumulators in TMEM), Ozaki-style error-free splitting.
//
// Why: both row passes of the sparse-GP evaluation are FP64 GEMMs bounded by the DMMA pipe (37 TF/s, DESIGN.md
// section 8).  K = sigma^2 exp(-d^2/2) lies in (0, sigma^2], so k = K/sigma^2 has an exact 62-bit fixed-point
// image q = rint(k 2^62); q = sum_t d_t 256^t with balanced digits d_t in [-128, 127] (8 INT8 slices, slice
// s = 7 - t has weight 2^(-6-8s)).  Then
//     (K^T K)_ij / sigma^4 = sum_{sa,sb} 2^(-12-8(sa+sb)) * sum_r d_sa[r,i] d_sb[r,j]
// where every inner sum is an exact INT32 dot product (|d d| <= 2^14, <= 8 pairs and 4096 rows per accumulator:
// < 2^30).  Pairs with sa + sb > 7 are dropped (< 2^-58 relative to sigma^4 per row).  36 INT8 MMAs replace one
// FP64 MMA; the INT8 pipe is ~120x the DMMA pipe.
//
// Kernel: one CTA per (128 x 64 output tile of the lower triangle, K split); all 8 levels L = sa + sb stay
// resident in TMEM (8 x 64 = 512 columns) over the CTA's whole row range, so each operand slice tile is loaded
// once per 64-row k-block (2-stage ring of 8 A + 8 B slice tiles = 192 KB) and 72 MMAs (128 x 64 x 32) are issued
// per stage by one thread.  Operands are K-major, SWIZZLE_NONE "interleaved" canonical layout; the splitter
// writes the slices to global memory already in that shared-memory image, so a tile is one contiguous bulk copy
// (cp.async.bulk + mbarrier complete_tx, no tensor map).  Epilogue: tcgen05.ld 32x32b, INT32 -> FP64, levels
// summed from the least significant one, one deterministic partial per (tile, split).
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ozaki_syrk ozaki_syrk.cu
// Run:    ./ozaki_syrk [rows=8192] [m=1024] [reps=20]      (prints one JSON line; exit code != 0 on mismatch)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

#include "tc_i8_r01.cuh"      // the product's INT8 engine: layout constants, PTX wrappers, issue_stage
using namespace srgp::i8;

// ---- splitter: doubles in (0, 1] (column j of K contiguous over rows) -> 8 INT8 slices in the smem image --------
// image of slice s: [jblk = j/128][kb = r/64][c = (r%64)/16][r1 = (j%128)/8][r0 = j%8][r%16]   (8 KB per (jblk, kb))
__global__ void split_kernel(const double *__restrict__ Kcm, int rows, int m, int8_t *__restrict__ slices, size_t slice_stride)
{
    const int groups = rows / 16;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)groups * m) return;
    const int j = (int)(gid / groups), g = (int)(gid % groups), r = g * 16;
    const double *src = Kcm + (size_t)j * rows + r;
    uint32_t w[NS][4];
#pragma unroll
    for (int s = 0; s < NS; ++s) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0;
#pragma unroll
    for (int e = 0; e < 16; e += 4) split_quad(src[e], src[e + 1], src[e + 2], src[e + 3], e >> 2, w);   // the product's splitter
    const int KB = rows / BK;
    const size_t off = ((size_t)(j / 128) * KB + r / BK) * (size_t)(128 * BK) + (size_t)((r % BK) / 16) * 2048 + (size_t)((j % 128) / 8) * 128 + (size_t)(j % 8) * 16;
#pragma unroll
    for (int s = 0; s < NS; ++s)
        *reinterpret_cast<uint4 *>(slices + s * slice_stride + off) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
}

// ---- the INT8 Gram kernel ----------------------------------------------------------------------------------------
struct Tile { int I, J; };       // output rows 128 I .., columns 64 J ..  (J <= 2 I + 1: lower block triangle)

__global__ void __launch_bounds__(THREADS, 1)
ozaki_syrk_kernel(const int8_t *__restrict__ slices, size_t slice_stride, int rows, int nsplit, const Tile *__restrict__ tiles,
                  double *__restrict__ Gpart, int *__restrict__ dbg_levels, int mode)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *tmem_full = empty + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Tile tile = tiles[blockIdx.x / nsplit];
    const int split = blockIdx.x % nsplit;
    const int KB = rows / BK;                       // k-blocks in the chunk
    const int kb_per = KB / nsplit, kb0 = split * kb_per;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc_all(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== producer: 8 A slice tiles (8 KB each) + 8 B slice tiles (4 x 1 KB each) per stage =====
            for (int it = 0; it < 2 * kb_per; ++it) {
                const int st = it % STAGES;
                if (it >= STAGES) mbar_wait(&empty[st], ((it / STAGES) - 1) & 1);
                const uint32_t sbase = smem_u32(smem + st * STAGE_BYTES);
                if (mode == 1 && it >= STAGES) {            // diagnostic: no operand traffic (stale smem), MMA-bound rate
                    mbar_expect_tx(&full[st], 16);
                    bulk_g2s(sbase, slices, 16, &full[st]);
                    continue;
                }
                load_stage(sbase, &full[st], slices, slice_stride, ((size_t)tile.I * KB + kb0) * IMG_BLOCK, slices, slice_stride,
                           ((size_t)(tile.J >> 1) * KB + kb0) * IMG_BLOCK, tile.J & 1, it);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer: 36 slice pairs x 2 k-steps per stage, level L = sa + sb -> TMEM columns [64 L, 64 L + 64) =====
            for (int it = 0; it < 2 * kb_per; ++it) {
                const int st = it % STAGES;
                mbar_wait(&full[st], (it / STAGES) & 1);
                tc_fence_after();
                const uint32_t sbase = smem_u32(smem + st * STAGE_BYTES);
                if (mode == 2) {                            // diagnostic: no MMAs, load-bound rate
                    if (it == 0) mma_i8(tmem_base, make_desc(sbase, 2048, 128), make_desc(sbase + NS * A_TILE, 1024, 128), 0u);
                    mma_commit(&empty[st]);
                    continue;
                }
                issue_stage(sbase, tmem_base, it == 0);
                mma_commit(&empty[st]);               // frees the stage when these MMAs have read it
            }
            mma_commit(tmem_full);                     // all levels complete
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes 32 (w % 4) .. + 31 = tile rows; INT32 -> FP64, levels LSB first =====
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row = q * 32 + lane;
        double *out = Gpart + ((size_t)blockIdx.x * BM + row) * BN;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            double acc[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = 0.0;
#pragma unroll 1
            for (int L = NS - 1; L >= 0; --L) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(L * BN + half * 32), v);
                const double wgt = exp2(-12.0 - 8.0 * L);
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[c] += wgt * (double)(int)v[c];
                if (dbg_levels) {
                    int *d = dbg_levels + (((size_t)blockIdx.x * NS + L) * BM + row) * BN + half * 32;
#pragma unroll
                    for (int c = 0; c < 32; ++c) d[c] = (int)v[c];
                }
            }
#pragma unroll
            for (int c = 0; c < 32; c += 2) *reinterpret_cast<double2 *>(out + half * 32 + c) = make_double2(acc[c], acc[c + 1]);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_free_all(tmem_base);
    }
}

// ---- exact per-level reference on one tile (plain integer loops) and an FP64 reference of the same tile -------------
__global__ void ref_levels_kernel(const double *__restrict__ Kcm, int rows, Tile tile, int r_begin, int r_end, long long *__restrict__ lev, double *__restrict__ g64)
{
    const int i = tile.I * BM + blockIdx.x, j = tile.J * BN + threadIdx.x;          // grid 128, block 64
    long long L[NS] = {0, 0, 0, 0, 0, 0, 0, 0};
    double hi = 0.0, lo = 0.0;
    for (int r = r_begin; r < r_end; ++r) {
        const double a = Kcm[(size_t)i * rows + r], b = Kcm[(size_t)j * rows + r];
        int da[NS], db[NS];
        long long qa = __double2ll_rn(a * 4611686018427387904.0), qb = __double2ll_rn(b * 4611686018427387904.0);
        for (int t = 0; t < NS; ++t) {
            long long d = ((qa + 128) & 255) - 128; qa = (qa - d) >> 8; da[NS - 1 - t] = (int)d;
            d = ((qb + 128) & 255) - 128; qb = (qb - d) >> 8; db[NS - 1 - t] = (int)d;
        }
        for (int sa = 0; sa < NS; ++sa)
            for (int sb = 0; sa + sb < NS; ++sb) L[sa + sb] += da[sa] * db[sb];
        const double p = a * b, e = fma(a, b, -p);                                    // double-double accumulation
        const double s = hi + p, bb = s - hi;
        lo += ((hi - (s - bb)) + (p - bb)) + e;
        hi = s;
    }
    for (int l = 0; l < NS; ++l) lev[((size_t)l * BM + blockIdx.x) * BN + threadIdx.x] = L[l];
    g64[(size_t)blockIdx.x * BN + threadIdx.x] = hi + lo;
}

int main(int argc, char **argv)
{
    const int rows = argc > 1 ? atoi(argv[1]) : 8192, m = argc > 2 ? atoi(argv[2]) : 1024, reps = argc > 3 ? atoi(argv[3]) : 20;
    const int nsplit = 2;
    if (rows % (BK * nsplit) || m % 128) { fprintf(stderr, "rows must be a multiple of %d, m of 128\n", BK * nsplit); return 2; }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));

    // K chunk: k = exp(-u), u ~ the squared-distance profile of the headline config (mostly small entries, some near 1)
    std::vector<double> hK((size_t)rows * m);
    uint64_t seed = 0x9E3779B97F4A7C15ull;
    for (size_t e = 0; e < hK.size(); ++e) {
        seed = seed * 6364136223846793005ull + 1442695040888963407ull;
        const double u = (double)(seed >> 11) / 9007199254740992.0;
        hK[e] = exp(-12.0 * u * u);
    }
    hK[5] = 1.0;                                       // a coincident pair: K == sigma^2 exactly
    double *dK; CK(cudaMalloc(&dK, hK.size() * 8)); CK(cudaMemcpy(dK, hK.data(), hK.size() * 8, cudaMemcpyHostToDevice));
    const size_t slice_stride = (size_t)rows * m;
    int8_t *dS; CK(cudaMalloc(&dS, slice_stride * NS));

    std::vector<Tile> tiles;
    for (int I = 0; I < m / BM; ++I) for (int J = 0; J <= 2 * I + 1; ++J) tiles.push_back({I, J});
    const int ntiles = (int)tiles.size(), grid = ntiles * nsplit;
    Tile *dT; CK(cudaMalloc(&dT, tiles.size() * sizeof(Tile))); CK(cudaMemcpy(dT, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice));
    double *dG; CK(cudaMalloc(&dG, (size_t)grid * BM * BN * 8));
    int *dDbg; CK(cudaMalloc(&dDbg, (size_t)grid * NS * BM * BN * 4));

    const int smem_bytes = STAGES * STAGE_BYTES + 256;
    CK(cudaFuncSetAttribute(ozaki_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    const size_t ngroups = (size_t)(rows / 16) * m;
    split_kernel<<<(unsigned)((ngroups + 255) / 256), 256>>>(dK, rows, m, dS, slice_stride);
    CK(cudaGetLastError());
    ozaki_syrk_kernel<<<grid, THREADS, smem_bytes>>>(dS, slice_stride, rows, nsplit, dT, dG, dDbg, 0);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());

    // ---- check three tiles: every INT32 level bit for bit, and the FP64 result against double-double ----
    std::vector<double> hG((size_t)grid * BM * BN);
    std::vector<int> hDbg((size_t)grid * NS * BM * BN);
    CK(cudaMemcpy(hG.data(), dG, hG.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hDbg.data(), dDbg, hDbg.size() * 4, cudaMemcpyDeviceToHost));
    long long *dLev; double *dG64; CK(cudaMalloc(&dLev, (size_t)NS * BM * BN * 8)); CK(cudaMalloc(&dG64, (size_t)BM * BN * 8));
    std::vector<long long> hLev((size_t)NS * BM * BN); std::vector<double> hG64((size_t)BM * BN);
    long long level_mismatch = 0; double max_rel = 0.0;
    const int check[3] = {0, ntiles / 2, ntiles - 1};
    for (int c = 0; c < 3; ++c) {
        const int t = check[c];
        std::vector<double> sum((size_t)BM * BN, 0.0), ref((size_t)BM * BN, 0.0);
        for (int sp = 0; sp < nsplit; ++sp) {
            ref_levels_kernel<<<BM, BN>>>(dK, rows, tiles[t], sp * rows / nsplit, (sp + 1) * rows / nsplit, dLev, dG64);
            CK(cudaGetLastError());
            CK(cudaMemcpy(hLev.data(), dLev, hLev.size() * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hG64.data(), dG64, hG64.size() * 8, cudaMemcpyDeviceToHost));
            const size_t cta = (size_t)t * nsplit + sp;
            for (int l = 0; l < NS; ++l)
                for (int e = 0; e < BM * BN; ++e)
                    if ((long long)hDbg[(cta * NS + l) * BM * BN + e] != hLev[(size_t)l * BM * BN + e]) ++level_mismatch;
            for (int e = 0; e < BM * BN; ++e) { sum[e] += hG[cta * BM * BN + e]; ref[e] += hG64[e]; }
        }
        for (int e = 0; e < BM * BN; ++e) max_rel = fmax(max_rel, fabs(sum[e] - ref[e]) / fabs(ref[e]));
    }

    // ---- timing ----
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) split_kernel<<<(unsigned)((ngroups + 255) / 256), 256>>>(dK, rows, m, dS, slice_stride);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms_split; CK(cudaEventElapsedTime(&ms_split, e0, e1)); ms_split /= reps;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) ozaki_syrk_kernel<<<grid, THREADS, smem_bytes>>>(dS, slice_stride, rows, nsplit, dT, dG, nullptr, 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    float ms_mode[3] = {ms, 0.f, 0.f};
    for (int mode = 1; mode <= 2; ++mode) {
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) ozaki_syrk_kernel<<<grid, THREADS, smem_bytes>>>(dS, slice_stride, rows, nsplit, dT, dG, nullptr, mode);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms_mode[mode], e0, e1)); ms_mode[mode] /= reps;
    }
    const double int8_ops = 36.0 * 2.0 * (double)rows * BM * BN * ntiles;            // executed INT8 multiply-adds x 2
    const double f64_flops = (double)rows * m * (m + 1);                                // the SYRK count bench.py uses
    printf("{\"gpu\": \"%s\", \"rows\": %d, \"m\": %d, \"ctas\": %d, \"ms_gram\": %.4f, \"ms_split\": %.4f, \"int8_tops\": %.1f, "
           "\"fp64_equiv_tflops\": %.2f, \"fp64_equiv_tflops_with_split\": %.2f, \"ms_mma_only\": %.4f, \"ms_load_only\": %.4f, \"l2_to_sm_TBps\": %.2f, \"level_mismatches\": %lld, \"max_rel_err_vs_double_double\": %.3e}\n",
           prop.name, rows, m, grid, ms, ms_split, int8_ops / ms * 1e-9, f64_flops / ms * 1e-9, f64_flops / (ms + ms_split) * 1e-9,
           ms_mode[1], ms_mode[2], (double)grid * (2 * rows / BK / nsplit) * STAGE_BYTES / ms * 1e-9, level_mismatch, max_rel);
    return (level_mismatch == 0 && max_rel < 1e-14) ? 0 : 1;
}

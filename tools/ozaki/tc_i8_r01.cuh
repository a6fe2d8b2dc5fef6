// Frozen copy of sparsergps_b200/csrc/tc_i8.cuh as of round 1 / early round 2 (128 x 64 single-sweep tiles, slice-major
// operand images): the engine the probes in this directory were written and measured against.  NOT product code.
// tc_i8.cuh -- FP64-equivalent products on the INT8 tensor cores of sm_100a (tcgen05.mma.kind::i8, accumulators in
// TMEM): the operand layout, the PTX wrappers and the digit splitter shared by the Gram (pass 1) and K*M (pass 2)
// kernels of gauss_i8.cu.
//
// Scheme (Ozaki-style error-free splitting, fixed point): a value v in [-1, 1] is q = rint(v 2^62), and
// q = sum_t d_t 256^t with balanced digits d_t in [-128, 127] -- NS = 8 INT8 slices, slice s = 7 - t carries weight
// 2^(-6-8s).  A product sum_r a_r b_r is then sum_{sa,sb} 2^(-12-8(sa+sb)) (sum_r da_sa[r] db_sb[r]) where every
// inner sum is an exact INT32 dot product; pairs with sa + sb > 7 are dropped (< 2^-58 of the operand scales per
// term).  All pairs of one level L = sa + sb accumulate into the same TMEM accumulator (8 levels x 64 columns = the
// whole 512-column TMEM), so a 128 x 64 output tile costs 36 INT8 MMAs per k-step instead of one FP64 MMA, on a pipe
// that is ~120x wider than DMMA.  Overflow bound: (L + 1) <= 8 pairs x 2^14 x rows <= 2^30 for rows <= 8192
// (the K*M pass sums over the mp knots: mp <= 16384, checked by i8_pass2_supported).
//
// Operand layout: K-major, SWIZZLE_NONE ("interleaved") canonical UMMA layout -- 8 rows x 16 bytes core matrices
// stored as 128 contiguous bytes.  The producers of the slices (generator kernels) write them to global memory
// ALREADY in the shared-memory image,
//     image(slice s)[blk = row / 128][kb = k / 64][c = (k % 64) / 16][r1 = (row % 128) / 8][r0 = row % 8][k % 16]
// (8 KB per (blk, kb)), so an operand tile is one contiguous cp.async.bulk (TMA bulk copy, mbarrier complete_tx;
// no tensor map) and its matrix descriptor has LBO = 2048 (1024 for a 64-row half tile), SBO = 128.
#pragma once
#include <stdint.h>

namespace srgp {
namespace i8 {

// Slices per operand: a compile-time choice.  7 (default): q = rint(v 2^54), absolute 2^-55 on |v| <= 1 -- at or below
// half an ulp of the double itself for |v| >= 1/4 -- and 28 slice pairs (levels 0..6; the dropped pairs weigh
// < 2^-51 of the operand scales per term, typically 2^-54: one FP64 product rounding).  8: q = rint(v 2^62), 36 pairs
// (dropped < 2^-58): the validation build, `-DSRGP_I8_NS=8`; tests/test_i8_gpu.py holds the 7-slice default to the
// long-double yardstick.
#ifndef SRGP_I8_NS
#define SRGP_I8_NS 7
#endif
constexpr int NS = SRGP_I8_NS;
static_assert(NS == 7 || NS == 8, "digit slices per operand");
constexpr int NPAIRS = NS * (NS + 1) / 2;   // MMAs per k-step
constexpr int FIX_BITS = 8 * NS - 2;  // fixed-point fraction bits: 54 / 62
constexpr int BM = 128, BN = 64;      // output tile (BN x NS <= 512 TMEM columns)
constexpr int BK = 64;                // k extent of one block of the operand image, in INT8 elements = bytes
constexpr int KS = 32;                // k extent of one pipeline stage = one MMA k-step (half an image block)
constexpr int STAGES = NS == 8 ? 4 : 5;   // 4 x 48 KB / 5 x 42 KB: all but one stage in flight while one is consumed.  With 2 x 96 KB
                                      // only one stage was ever in flight and the K*M pass, whose K slices come from
                                      // HBM rather than L2, ran at 2.66 us per 64-byte block instead of 2.17
constexpr int A_TILE = BM * KS;       // 4 KB
constexpr int B_TILE = BN * KS;       // 2 KB
constexpr int STAGE_BYTES = NS * (A_TILE + B_TILE);   // 48 KB
constexpr int IMG_BLOCK = 128 * BK;   // bytes of one (128-row block, k-block) image
constexpr int THREADS = 192;          // warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: epilogue
constexpr int MAX_ROWS_PER_SPLIT = 8192;
constexpr double FIX_SCALE = (double)(1ull << FIX_BITS);   // 2^54 / 2^62
constexpr long long FIX_ONE = 1ll << FIX_BITS;             // the image of exp(0) = 1: candidate for quirk Q4
// instruction descriptor: D = S32, A = B = signed INT8, both K-major, N = 64, M = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// A protocol bug must surface as a launch error (SRGP_ERR_CUDA), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try(bar, parity))
        if (++spins > (1u << 28)) __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// K-major SWIZZLE_NONE matrix descriptor (version 1 = Blackwell): start address, LBO = byte distance between the
// two 16-byte k-chunks of one MMA, SBO = byte distance between 8-row groups; all in units of 16 bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}" ::"r"(tmem_d),
                 "l"(da), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc_all(uint32_t *slot)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free_all(uint32_t base)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}
// 32 lanes x 32 consecutive columns: thread t of the warp receives row (lane base + t), columns col .. col + 31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
                 "%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 16 columns of the NS level accumulators (columns col + L * BN) -> two exact 64-bit integers per entry:
// hi = sum_{k<4} lev_k 256^(3-k) (levels 0..3, weight 2^-36) and lo = sum_{k<NS-4} lev_{4+k} 256^(NS-5-k) (levels 4..NS-1,
// weight 2^(-12-8(NS-1))); |lev| < 2^30, so both stay below 2^55.  The INT32 -> FP64 conversions and the weighting of
// the levels cost 3 FP64 instructions per entry instead of 2 NS (the FP64 pipe is the contended one).
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// NL consecutive levels starting at taddr: acc = sum_k lev_k 256^(NL-1-k); the NL loads are in flight together (one
// TMEM round trip per 16 x NL block)
template <int NL>
__device__ __forceinline__ void drain16(uint32_t taddr, long long (&acc)[16])
{
    static_assert(NL == 3 || NL == 4, "levels per exact 64-bit group");
    uint32_t v[NL][16];
#pragma unroll
    for (int k = 0; k < NL; ++k) tmem_ld16_nowait(taddr + (uint32_t)(k * BN), v[k]);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        long long a = (long long)(int)v[0][c];
#pragma unroll
        for (int k = 1; k < NL; ++k) a = a * 256 + (long long)(int)v[k][c];
        acc[c] = a;
    }
}
constexpr double W_LEVELS_HI = 1.4551915228366852e-11;                           // 2^-36: levels 0..3 as one integer
constexpr double W_LEVELS_LO = NS == 8 ? 3.3881317890172014e-21 : 8.673617379884035e-19;   // 2^-68 / 2^-60: levels 4..NS-1
constexpr double FIX_INV = 1.0 / FIX_SCALE;                                      // 2^-54 / 2^-62

// 4 x 4 byte transpose: out[t] = (a.byte_t, b.byte_t, c.byte_t, d.byte_t), a's byte in bits 0..7  (8 PRMT)
__device__ __forceinline__ void transpose4x4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t (&out)[4])
{
    const uint32_t t0 = __byte_perm(a, b, 0x5140), t1 = __byte_perm(a, b, 0x7362);
    const uint32_t t2 = __byte_perm(c, d, 0x5140), t3 = __byte_perm(c, d, 0x7362);
    out[0] = __byte_perm(t0, t2, 0x5410);
    out[1] = __byte_perm(t0, t2, 0x7632);
    out[2] = __byte_perm(t1, t3, 0x5410);
    out[3] = __byte_perm(t1, t3, 0x7632);
}
constexpr unsigned long long DIGIT_BIAS = NS == 8 ? 0x8080808080808080ull : 0x0080808080808080ull;   // 128 in each of the NS digit bytes

// The inverse of split_quad for the 4 entries of word `wi` (0..3) of a 16-byte k-chunk: exact q of each entry.
__device__ __forceinline__ void join_quad(const uint4 (&w)[NS], int wi, long long (&q)[4])
{
    uint32_t in[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) in[s] = wi == 0 ? w[s].x : wi == 1 ? w[s].y : wi == 2 ? w[s].z : w[s].w;
    uint32_t lo[4], hi[4];
    transpose4x4(in[NS - 1], in[NS - 2], in[NS - 3], in[NS - 4], lo);            // digits 0..3 live in slices NS-1..NS-4
    transpose4x4(in[NS - 5], in[NS - 6], in[NS - 7], NS == 8 ? in[0] : 0u, hi);  // digits 4..NS-1 in slices NS-5..0
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned long long y = ((unsigned long long)hi[k] << 32) | lo[k];
        q[k] = (long long)((y ^ DIGIT_BIAS) - DIGIT_BIAS);
    }
}

// Balanced base-256 digits of q = rint(v 2^FIX_BITS), |v| <= 1, for 4 consecutive entries (word `wi` of a 16-byte k-chunk).
// q = sum_t d_t 256^t with d_t in [-128, 127]  <=>  q + B = sum_t (d_t + 128) 256^t with B = 0x80...80 (NS bytes), i.e. the
// plain bytes of q + B (< 2^(8 NS), the bytes above stay 0); and d_t as a two's-complement INT8 is (d_t + 128) ^ 0x80.  So all NS digits of an entry are
// the bytes of (q + B) ^ B: two 64-bit integer operations instead of a carry chain, then two 4 x 4 byte transposes
// put digit t of the 4 entries into one 32-bit word of slice NS - 1 - t.
__device__ __forceinline__ void split_quad(double v0, double v1, double v2, double v3, int wi, uint32_t (&w)[NS][4])
{
    unsigned long long y[4];
    const double v[4] = {v0, v1, v2, v3};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        y[k] = ((unsigned long long)__double2ll_rn(v[k] * FIX_SCALE) + DIGIT_BIAS) ^ DIGIT_BIAS;
    uint32_t lo[4], hi[4];
    transpose4x4((uint32_t)y[0], (uint32_t)y[1], (uint32_t)y[2], (uint32_t)y[3], lo);
    transpose4x4((uint32_t)(y[0] >> 32), (uint32_t)(y[1] >> 32), (uint32_t)(y[2] >> 32), (uint32_t)(y[3] >> 32), hi);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        w[NS - 1 - t][wi] = lo[t];
        if (t + 4 < NS) w[NS - 5 - t][wi] = hi[t];
    }
}

// shared-memory carve-up of the two INT8 kernels
struct Bars {
    uint64_t full[STAGES], empty[STAGES], tmem_full;
    uint32_t tmem_slot, pad;
};
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + (int)sizeof(Bars);

// The issue sequence of one stage (one 32-byte k-step): NPAIRS slice pairs; level L = sa + sb -> TMEM columns [64 L, 64 L + 64).
// `fresh` = this is the first stage of the accumulation (the first MMA of every level overwrites).
// One thread issues all NPAIRS MMAs, so the sequence is fully unrolled and every descriptor is the stage's base
// descriptor plus a compile-time constant (the start-address field is the low 14 bits, in 16-byte units; the sums
// stay below 2^14): 2 integer adds per MMA.  With descriptors rebuilt per MMA the issuing thread, not the tensor
// pipe, was the limit (85 instead of ~45 cycles per MMA, profiles/r01_ozaki_proto.json).
__device__ __forceinline__ void issue_stage(uint32_t sbase, uint32_t tmem_base, bool fresh)
{
    const uint64_t da0 = make_desc(sbase, 2048, 128);
    const uint64_t db0 = make_desc(sbase + NS * A_TILE, 1024, 128);
    const uint32_t keep = fresh ? 0u : 1u;
#pragma unroll
    for (int sb = 0; sb < NS; ++sb) {
#pragma unroll
        for (int sa = 0; sa < NS; ++sa) {
            if (sa + sb < NS) {
                const uint64_t da = da0 + (uint64_t)((sa * A_TILE) >> 4);
                const uint64_t db = db0 + (uint64_t)((sb * B_TILE) >> 4);
                // level L is first touched by the pair (sa = L, sb = 0)
                mma_i8(tmem_base + (uint32_t)(sa + sb) * BN, da, db, sb == 0 ? keep : 1u);
            }
        }
    }
}

// Producer side of one stage (k-step ks of the operand images, 32 bytes): 8 A slice tiles (128 rows: the two 16-byte
// k-chunks of a 128-row block image are 4 KB contiguous) and 8 B slice tiles (64 rows = one half of a 128-row block:
// 1 KB per k-chunk).  a_blk / b_blk: byte offsets of the operands' (block, k-block 0) images inside a slice.
__device__ __forceinline__ void load_stage(uint32_t sbase, uint64_t *full, const int8_t *a_slices, size_t a_stride, size_t a_blk,
                                           const int8_t *b_slices, size_t b_stride, size_t b_blk, int b_half, int ks)
{
    mbar_expect_tx(full, STAGE_BYTES);
    const size_t a_off = a_blk + (size_t)(ks >> 1) * IMG_BLOCK + (size_t)(ks & 1) * A_TILE;
    const size_t b_off = b_blk + (size_t)(ks >> 1) * IMG_BLOCK + (size_t)(ks & 1) * 4096 + (size_t)b_half * 1024;
    for (int s = 0; s < NS; ++s) {
        bulk_g2s(sbase + s * A_TILE, a_slices + s * a_stride + a_off, A_TILE, full);
        bulk_g2s(sbase + NS * A_TILE + s * B_TILE, b_slices + s * b_stride + b_off, 1024, full);
        bulk_g2s(sbase + NS * A_TILE + s * B_TILE + 1024, b_slices + s * b_stride + b_off + 2048, 1024, full);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 128 x 128 output tiles in TWO SWEEPS over k.  An MMA reads its A tile (128 rows x 32 B) from shared memory whatever N
// is, so at N = 64 the shared-memory port (128 B/clk: 4 KB + 2 KB per 32-clk MMA) and not the tensor pipe bounds the
// kernels (tools/ozaki/mma_rate.cu: 48 clk per MMA resident, 58 with the TMA writes; N = 128: 64 clk = the pipe).  N = 128
// leaves room for 4 accumulators in the 512 TMEM columns, so a tile is computed as two "virtual tiles" over the same k
// range: sweep 0 = levels 0..3 (10 slice pairs, needs slices 0..3 of both operands only), sweep 1 = levels 4..NS-1 (18
// pairs, all slices).  Both level groups are contiguous, so each drains as one exact 64-bit integer per entry, and every
// quantity downstream is linear in the product, so the two sweeps are simply two contributions.
// ------------------------------------------------------------------------------------------------------------------
constexpr int BN2 = 128;
constexpr int STAGE2_BYTES = NS * 2 * A_TILE;          // 56 KB (NS = 7): NS A tiles + NS B tiles of 4 KB
__host__ __device__ constexpr int sweep_slices(int sw) { return sw == 0 ? 4 : NS; }       // operand slices a sweep reads
__host__ __device__ constexpr int sweep_levels(int sw) { return sw == 0 ? 4 : NS - 4; }   // accumulators it fills
constexpr uint32_t IDESC2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_i8_n128(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}" ::"r"(tmem_d),
                 "l"(da), "l"(db), "r"(IDESC2), "r"(accumulate), "r"(0u) : "memory");
}

// one 32-byte k-step of sweep SW: level L = sa + sb -> TMEM columns [128 (L - 4 SW), + 128).  a_base / b_base: shared
// addresses of the A and B slice tiles of the stage (b_base == a_base on the diagonal tiles of a Gram).
template <int SW>
__device__ __forceinline__ void issue_stage2(uint32_t a_base, uint32_t b_base, uint32_t tmem_base, bool fresh)
{
    const uint64_t da0 = make_desc(a_base, 2048, 128);
    const uint64_t db0 = make_desc(b_base, 2048, 128);
    const uint32_t keep = fresh ? 0u : 1u;
#pragma unroll
    for (int sb = 0; sb < NS; ++sb) {
#pragma unroll
        for (int sa = 0; sa < NS; ++sa) {
            const int L = sa + sb;
            if (SW == 0 ? L < 4 : (L >= 4 && L < NS)) {
                const uint64_t da = da0 + (uint64_t)((sa * A_TILE) >> 4);
                const uint64_t db = db0 + (uint64_t)((sb * A_TILE) >> 4);
                // accumulator L is first touched by the pair (sa = L, sb = 0)
                mma_i8_n128(tmem_base + (uint32_t)(L - 4 * SW) * BN2, da, db, sb == 0 ? keep : 1u);
            }
        }
    }
}

// producer side of one stage of sweep sw: slices 0 .. sweep_slices(sw) - 1 of the A block (and of the B block unless it is
// the same block); a 128-row block's two 16-byte k-chunks of a 32-byte k-step are 4 KB contiguous in the image
__device__ __forceinline__ void load_stage2(uint32_t sbase, uint64_t *full, const int8_t *a_slices, size_t a_stride, size_t a_blk,
                                            const int8_t *b_slices, size_t b_stride, size_t b_blk, bool same, int ks, int nsl)
{
    mbar_expect_tx(full, (uint32_t)(nsl * A_TILE * (same ? 1 : 2)));
    const size_t off = (size_t)(ks >> 1) * IMG_BLOCK + (size_t)(ks & 1) * A_TILE;
    for (int s = 0; s < nsl; ++s) {
        bulk_g2s(sbase + s * A_TILE, a_slices + s * a_stride + a_blk + off, A_TILE, full);
        if (!same) bulk_g2s(sbase + (NS + s) * A_TILE, b_slices + s * b_stride + b_blk + off, A_TILE, full);
    }
}

// ---- thread-block clusters: operand tiles that several CTAs need at the same time are fetched from L2 once and
// multicast into the shared memory of all of them (the row passes are bound by the L2 -> SM path, not by the tensor
// pipe or HBM: profiles/r02_km_bound.txt) ----
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// data lands at the same CTA-relative offset, and complete_tx is signalled on the same-offset mbarrier, in every CTA of
// `mask`
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar, uint16_t mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// arrives (once the issued MMAs completed) on the same-offset mbarrier of every CTA of `mask`
__device__ __forceinline__ void mma_commit_mc(uint64_t *bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask) : "memory");
}

// ---- CTA pairs (tcgen05 cta_group::2): one MMA spans two SMs -- M = 256 = 128 rows per CTA, each CTA holds its own A
// tile and HALF of the B tile (the hardware reads the peer's half), each CTA's TMEM receives its 128 rows x N columns.
// The B operand is therefore fetched from L2 once per pair instead of once per CTA (tools/ozaki/pair_rate.cu checks the
// levels bit for bit against single-CTA MMAs).  Only the leader (cluster rank 0) issues MMAs and commits. ----
constexpr uint32_t IDESC2_PAIR = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ void mma_i8_n128_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}" ::"r"(tmem_d),
                 "l"(da), "l"(db), "r"(IDESC2_PAIR), "r"(accumulate), "r"(0u) : "memory");
}
// arrives on the same-offset mbarrier of both CTAs of the pair once the MMAs issued so far completed
__device__ __forceinline__ void mma_commit_pair(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_all_pair(uint32_t *slot)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free_all_pair(uint32_t base)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank)
{
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// wait on a local mbarrier whose arrivals come from the peer CTA (acquire at cluster scope)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    uint32_t spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (++spins > (1u << 28)) __trap();
    }
}

// NL levels at stride BN2 columns: acc = sum_k lev_k 256^(NL-1-k) for 16 columns
template <int NL>
__device__ __forceinline__ void drain16_n128(uint32_t taddr, long long (&acc)[16])
{
    static_assert(NL == 3 || NL == 4, "levels per exact 64-bit group");
    uint32_t v[NL][16];
#pragma unroll
    for (int k = 0; k < NL; ++k) tmem_ld16_nowait(taddr + (uint32_t)(k * BN2), v[k]);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        long long a = (long long)(int)v[0][c];
#pragma unroll
        for (int k = 1; k < NL; ++k) a = a * 256 + (long long)(int)v[k][c];
        acc[c] = a;
    }
}

}  // namespace i8
}  // namespace srgp

// Kernel 2 (tensor-core variant) -- fused query x prototype match on tcgen05 (sm_100a).
//
// Replaces safe_norm(qry) + get_prediction_from_prototypes of the reference
// (models/alpmodule.py:14-18, 57-94, 195).  Same contract as the CUDA-core variant
// (psam_match_simt.cu): d[m,n] = 20 * <q_m, p_n> / max(|q_m|, 1e-4) against the already
// normalised prototypes, reduced over the prototype axis of each set in the epilogue
// (softmax-weighted sum + argmax for the grid modes, max for 'mask'), so the [P, HW] similarity
// tensor never exists in memory.
//
// Why tensor cores: ncu on the CUDA-core variant (profiles/r1_prof_match_simt_raw.csv) shows the
// [HW x C].[C x sum(P)] contraction compute-bound on the FMA pipe at the named shapes (DRAM 0.6 %
// of peak, arithmetic intensity sum(P)/2 = 600 FLOP/B at config 2).
//
// Why three bf16 passes instead of one TF32 pass: maps must stay within 1e-3 absolute of the fp32
// reference.  A TF32 operand keeps 10 mantissa bits; on unit vectors of C = 768 that is a
// ~3e-4 (1 sigma) error on d = 20*cos, too thin over millions of outputs.  Each fp32 operand is
// therefore split x = hi + lo with hi = bf16(x), lo = bf16(x - hi) and the product is evaluated
// as  hi_a*hi_b + lo_a*hi_b + hi_a*lo_b  (kind::f16, fp32 accumulation in TMEM); the dropped
// lo_a*lo_b term and the rounding of lo are both ~2^-17 relative, i.e. the result is fp32-grade
// (measured max |d - d_fp32| ~ 1e-5) at 1.5x the tensor time of a single TF32 pass.
//
// Structure (one persistent CTA per SM):
//   k_pack_protos  prototype rows of all sets, concatenated (each set padded to 16 columns) so one B matrix serves
//                  every set -> bf16 hi/lo planes already in the swizzled K-major layout the MMA reads.
//   k_pack_query   [algo 2] the same for the query rows, per (128-row tile, 32-channel block), plus the per-row
//                  scale 20/max(|q|,1e-4).
//   k_match_tc     [algo 2] warp 0: one thread streams operand blocks global -> shared with cp.async.bulk
//                  (TMA engine, mbarrier complete_tx) through a ring of 48 KB stages;
//                  warp 1: one thread issues tcgen05.mma (M=128, N<=256, K=16, 6 per 32-channel k-block)
//                  into one of two 256-column TMEM accumulators and commits to mbarriers;
//                  warps 2-5: epilogue -- tcgen05.ld 16 columns at a time, scale, exp2, running
//                  sum(e), sum(e*d), max/argmax per set, store one float per (row, set).
//                  MMA of chunk i+1 overlaps the epilogue of chunk i.
//   k_match_ts     [algo 3] the query operand never exists in memory as bf16: warp 0 streams the RAW fp32 rows of the
//                  tile (one 128-byte cp.async.bulk per row and k-block) next to the prototype block; warps 6-9
//                  (thread = query row = TMEM lane) read their row from shared memory, split it into bf16 hi/lo,
//                  accumulate the row norm and write the halves with tcgen05.st into a TMEM operand slot; the MMAs
//                  take A from TMEM and only B from shared memory.  The query is read from HBM once by the whole
//                  path, k_pack_query and its 2 x 135 MB round trip are gone, and the MMA's shared-memory reads
//                  drop by a third (ncu: the L1 data pipe, which the MMA operand reads share with every LDG/STS,
//                  is what limited the earlier LDG -> STS converter variant; see profiles/README.md).
//
// Roofline: tensor-bound.  Algorithmic flops per (slice, set) = 2*HW*C*P (executed: 3x that in
// bf16).  Algorithmic bytes: SURVEY.md section 8(d).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <type_traits>

#include <cuda.h>
#include <cuda_bf16.h>
#include <math_constants.h>

#include "psam_match.cuh"

namespace psam {

namespace tc {

// k-block depth: 64 bf16 = 128-byte rows (SWIZZLE_128B, 2 stages of 96 KB) or 32 bf16 = 64-byte rows (SWIZZLE_64B,
// 4 stages of 48 KB: the same bytes in flight, but a stage is requested three MMA batches ahead instead of one)
#ifndef PSAM_TC_BK
#define PSAM_TC_BK 32
#endif
constexpr int BM = 128;                          // query rows per tile = TMEM lanes
constexpr int BK = PSAM_TC_BK;                   // bf16 elements per k-block = one swizzle row
static_assert(BK == 32 || BK == 64, "k-block depth");
constexpr int NCH = 256;                         // prototype columns per TMEM accumulator buffer
constexpr int MAX_STAGES = BK == 64 ? 2 : 4;
constexpr int ROW_BYTES = BK * 2;                // one operand row of a k-block
constexpr int CHUNKS = ROW_BYTES / 16;           // 16-byte chunks per row
constexpr int PLANE_BYTES = 8 * ROW_BYTES;       // 8 rows of one plane = one swizzle atom
constexpr int GROUP_BYTES = 2 * PLANE_BYTES;     // 8 rows: hi plane then lo plane
constexpr uint64_t UMMA_LAYOUT = BK == 64 ? 2 : 4;   // cute::UMMA::LayoutType SWIZZLE_128B / SWIZZLE_64B

// physical 16-byte chunk of logical chunk c in row r of an 8-row atom (Swizzle<3,4,3> / Swizzle<2,4,3>)
__host__ __device__ __forceinline__ int swz(int r, int c) { return BK == 64 ? (c ^ r) : (c ^ ((r >> 1) & 3)); }
constexpr int A_STAGE_BYTES = BM / 8 * GROUP_BYTES;    // 16 KB (BK = 32)
constexpr int B_STAGE_BYTES = NCH / 8 * GROUP_BYTES;   // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
__host__ __device__ constexpr int smem_bytes(int stages) { return stages * STAGE_BYTES + 1024; }  // + slack for the 1024-byte alignment
constexpr int MAX_SETS = 128;
constexpr int MAX_SPLIT = 8;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;
// ---- TMEM-sourced fused variant (k_match_ts) ----
// TMEM budget: 2 accumulators of kNch columns + operand slots of 32 columns (one k-block each: 16 hi + 16 lo) = 512:
// kNch = 224 -> 2 slots, 208 -> 3, 192 -> 4
__host__ __device__ constexpr int ts_slots(int nch) { return (TMEM_COLS - 2 * nch) / 32; }
constexpr int TS_CONV_WARPS = 4;                 // warps 6..9: thread = row of the tile = TMEM lane
constexpr int TS_THREADS = THREADS + TS_CONV_WARPS * 32;
constexpr int TS_RAW_PITCH = BK * 4;             // raw fp32 row of a k-block: 128 bytes, 16-byte chunks XOR-swizzled by the TMA unit
constexpr int TS_RAW_BYTES = BM * TS_RAW_PITCH;  // 16 KB = one tensor-map box (32 channels x 128 rows)
__host__ __device__ constexpr int ts_b_bytes(int nch) { return nch / 8 * GROUP_BYTES; }          // 28 KB at 224 columns
__host__ __device__ constexpr int ts_stage_bytes(int nch) { return ts_b_bytes(nch) + TS_RAW_BYTES; }
__host__ __device__ constexpr int ts_smem_bytes(int stages, int nch) { return stages * ts_stage_bytes(nch) + 1024; }
static_assert(BK == 32, "k_match_ts is written for 32-channel k-blocks");
constexpr float LOG2E = 1.4426950408889634f;

__host__ __device__ __forceinline__ int pad16(int x) { return (x + 15) & ~15; }

// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// The same for waiters with slack (producer, converters, epilogue): the try_wait carries a suspend-time hint, so the warp
// sleeps in hardware until the phase completes instead of re-issuing the try (ncu: 30 % of the instructions the fused
// kernel executed were YIELD/TRYWAIT/BRA spins, issue slots the co-resident prompt kernels of other volumes want).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(1000000u)
        : "memory");
}

// Named barriers for plain shared-memory hand-offs between warp groups (the row scales of k_match_ts): the producers
// arrive without waiting, the consumers sync; `threads` = producers + consumers.  (An mbarrier would do as well, but
// compute-sanitizer's racecheck does not see mbarrier waits as ordering ordinary st.shared / ld.shared pairs.)
__device__ __forceinline__ void named_arrive(int id, int threads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ void named_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// global -> shared bulk copy on the TMA engine, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// one box of a 2-D tensor map (coordinates: channel, row) global -> shared; out-of-range rows / channels arrive as zeros
__device__ __forceinline__ void tensor_g2s(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, M=128, N from idesc, K=16
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// the same with A read from TMEM (lane = row, 32-bit column = two consecutive bf16 of the K axis)
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// 16 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major operand, 128- or 64-byte swizzle, 8-row groups GROUP_BYTES apart (cute::UMMA::SmemDescriptor:
// start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64))
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(GROUP_BYTES >> 4) << 32) |
           ((uint64_t)1 << 46) | (UMMA_LAYOUT << 61);
}

// cute::UMMA::InstrDescriptor: D=f32 [4,6), A=bf16 [7,10), B=bf16 [10,13), K-major A/B, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc(int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ float ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ------------------------------------------------------------------------------ operand packing
// 2 fp32 -> packed bf16 hi pair + packed bf16 lo pair, three instructions per element: one F2FP for the hi pair, a shift /
// mask to read the two halves back as fp32, one subtraction each, one F2FP for the lo pair
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo)
{
    const __nv_bfloat162 hb = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<const uint32_t*>(&hb);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xFFFF0000u);
    const __nv_bfloat162 lb = __floats2bfloat162_rn(x0 - h0, x1 - h1);
    lo = *reinterpret_cast<const uint32_t*>(&lb);
}

// 8 fp32 -> 8 bf16 hi + 8 bf16 lo (x = hi + lo up to 2^-17 relative)
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo)
{
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hb = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const float2 hf = __bfloat1622float2(hb);
        const __nv_bfloat162 lb = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hb);
        l[i] = *reinterpret_cast<const uint32_t*>(&lb);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// One warp per operand row: reads the row's channels 8 at a time (1 KB per warp iteration, coalesced),
// writes 16-byte chunks of the hi and lo planes at their swizzled position inside the row's 8-row group
// (chunk c of row r lands at r*128 + ((c ^ r) << 4)), returns the row's sum of squares.
__device__ __forceinline__ float pack_row(const float* __restrict__ src, bool valid, int C, int KB, int lane,
                                          uint8_t* __restrict__ group0, size_t kb_stride, int r)
{
    float ssq = 0.f;
    for (int ch = lane; ch < KB * CHUNKS; ch += 32) {
        const int k0 = ch * 8;
        float v[8];
        if (valid && k0 < C) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(src + k0));
            const float4 y = __ldg(reinterpret_cast<const float4*>(src + k0 + 4));
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) ssq = fmaf(v[i], v[i], ssq);
        uint4 hi, lo;
        split8(v, hi, lo);
        const int kb = ch / CHUNKS, c = ch % CHUNKS;
        uint8_t* dst = group0 + (size_t)kb * kb_stride + r * ROW_BYTES + (swz(r, c) << 4);
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + PLANE_BYTES) = lo;
    }
    return warp_sum(ssq);
}

// a_img[tile][kb][16 groups][2048 B]; rows beyond R are zero-filled.  grid = ntiles*128/8 blocks of 256.
__global__ void __launch_bounds__(256) k_pack_query(const float* __restrict__ qry, int64_t slice_stride,
                                                    int64_t row_stride, int HW, int R, int C, int KB,
                                                    uint8_t* __restrict__ a_img, float* __restrict__ scale)
{
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int tile = g >> 7, grp = (g & 127) >> 3, r = g & 7;
    TraceRec* tr = (threadIdx.x == 0 && (blockIdx.x & 15) == 0) ? trace_begin(2) : nullptr;   // every 16th CTA
    const bool valid = g < R;
    const float* src = valid ? qry + (size_t)(g / HW) * slice_stride + (size_t)(g % HW) * row_stride : qry;
    uint8_t* group0 = a_img + ((size_t)tile * KB * 16 + grp) * GROUP_BYTES;
    const float ssq = pack_row(src, valid, C, KB, lane, group0, (size_t)16 * GROUP_BYTES, r);
    if (lane == 0) scale[g] = valid ? 20.0f / fmaxf(sqrtf(ssq), 1e-4f) : 0.f;
    trace_end(tr);
}

// b_img[kb][G groups][2048 B] over the concatenated, 16-padded prototype columns of all sets.
// grid = (ceil(pad16(cap_rows)/8), nsets), block = 256 (one warp per row of the group).
__global__ void __launch_bounds__(256) k_pack_protos(const float* __restrict__ protos, int cap_rows,
                                                     const int32_t* __restrict__ counts, int C, int KB, int G,
                                                     uint8_t* __restrict__ b_img)
{
    __shared__ int s_base;
    const int set = blockIdx.y, count = counts[set];
    const int row0 = blockIdx.x * 8;
    if (row0 >= pad16(count)) return;
    if (threadIdx.x == 0) {
        int b = 0;
        for (int s = 0; s < set; ++s) b += pad16(counts[s]);
        s_base = b;
    }
    __syncthreads();
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = row0 + wid, col = s_base + n;
    const bool valid = n < count;
    const float* src = protos + ((size_t)set * cap_rows + (valid ? n : 0)) * C;
    uint8_t* group0 = b_img + (size_t)(col >> 3) * GROUP_BYTES;
    pack_row(src, valid, C, KB, lane, group0, (size_t)G * GROUP_BYTES, col & 7);
}

// ------------------------------------------------------------------------------ the GEMM + reduction
struct TcParams {
    const uint8_t* a_img;
    const uint8_t* b_img;
    const float* scale;
    const int32_t* counts;
    const int32_t* eff_modes;
    float* scores;
    float* assign;
    int32_t* status;
    int nsets, HW, R, ntiles, KB, G, nsplit;
    // fused variant: the query rows themselves (fp32, channels-last)
    const float* qry;
    int64_t slice_stride, row_stride;
    int C;
};

struct RowAcc {
    float se, sed, best;
    int bi;
};

// One 16-column group of a row folded into the running reductions of its set.
//   kGrid : softmax-weighted sum: se += e, sed += e*x with e = exp(20*cos - 20) (x is the raw dot product; the row scale
//           is applied to sed once per set)
//   kMax  : running maximum ('mask' mode, and the argmax of `assign`); kIndex additionally tracks its column
template <bool kFull, bool kGrid, bool kMax, bool kIndex>
__device__ __forceinline__ void fold16(const float (&v)[16], int nvalid, int nbase, float sc2, RowAcc& a)
{
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (kFull || j < nvalid) {
            const float x = v[j];
            if (kGrid) {
                const float e = ex2(fmaf(x, sc2, -20.0f * LOG2E));   // exp(d - 20), d = x*sc in [-20, 20]
                a.se += e;
                a.sed = fmaf(e, x, a.sed);
            }
            if (kIndex) {
                if (x > a.best) { a.best = x; a.bi = nbase + j; }
            } else if (kMax) {
                a.best = fmaxf(a.best, x);
            }
        }
    }
}

template <bool kGrid, bool kMax, bool kIndex>
__device__ __forceinline__ void fold16_any(const float (&v)[16], int nvalid, int nbase, float sc2, RowAcc& a)
{
    if (nvalid >= 16) fold16<true, kGrid, kMax, kIndex>(v, 16, nbase, sc2, a);
    else fold16<false, kGrid, kMax, kIndex>(v, nvalid, nbase, sc2, a);
}

// Column schedule of the persistent CTAs, derived on the device from the prototype counts (no host synchronisation):
// column splits at set boundaries, as even as the set sizes allow.
struct Schedule {
    int count[MAX_SETS];
    int split_set[MAX_SPLIT + 1], split_col[MAX_SPLIT + 1];
};

__device__ __forceinline__ void make_schedule(const TcParams& p, Schedule& sch)
{
    int T = 0;
    for (int s = 0; s < p.nsets; ++s) {
        const int c = p.counts[s];
        sch.count[s] = c;
        T += pad16(c);
    }
    sch.split_set[0] = 0;
    sch.split_col[0] = 0;
    int acc = 0, k = 1;
    for (int s = 0; s < p.nsets && k < p.nsplit; ++s) {
        acc += pad16(sch.count[s]);
        while (k < p.nsplit && (long long)acc * p.nsplit >= (long long)k * T) {
            sch.split_set[k] = s + 1;
            sch.split_col[k] = acc;
            ++k;
        }
    }
    for (; k < p.nsplit; ++k) { sch.split_set[k] = p.nsets; sch.split_col[k] = T; }
    sch.split_set[p.nsplit] = p.nsets;
    sch.split_col[p.nsplit] = T;
}

// The columns [c0, c1) of a work item are processed in accumulator-sized chunks.  Equal widths (in units of 16 columns)
// instead of "full chunks + a remainder": the MMA's cost per column rises as N shrinks, and a 60-column tail chunk after
// ten full ones (config 3: 2300 columns) costs a whole pass over the query tile for a quarter of the work.
struct Chunks {
    int n, base, rem;       // n chunks; chunk j is 16 * (base + (j < rem)) columns wide
};

__device__ __forceinline__ Chunks make_chunks(int cols, int nch_max)
{
    const int n = (cols + nch_max - 1) / nch_max, units = cols / 16;
    return n > 0 ? Chunks{n, units / n, units % n} : Chunks{0, 0, 0};
}

__device__ __forceinline__ int chunk_start(const Chunks& c, int j) { return 16 * (j * c.base + min(j, c.rem)); }
__device__ __forceinline__ int chunk_width(const Chunks& c, int j) { return 16 * (c.base + (j < c.rem ? 1 : 0)); }

// Epilogue of both GEMM kernels (warps 2-5): TMEM accumulator chunks of kNch columns -> per-set reductions -> global.
// kScaleFromSmem: the row scales 20/max(|q|,1e-4) come from the converter warps of the same CTA (k_match_ts) instead of
// the scale[] array k_pack_query wrote.
constexpr int SCALE_FULL_BAR = 1, SCALE_FREE_BAR = 3;     // named barriers 1-2 and 3-4 (0 is __syncthreads)

template <int kNch, bool kScaleFromSmem>
__device__ __forceinline__ void epilogue_loop(const TcParams& p, const Schedule& sch, uint32_t tmem_base, int warp, int lane,
                                              uint64_t* s_tfull, uint64_t* s_tempty, const float (*s_scale)[BM])
{
    const int nitems = p.ntiles * p.nsplit;
    const int lg = warp & 3;                       // TMEM lane group this warp may read
    const int row_in_tile = lg * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    const bool want_index = p.assign != nullptr;
    uint32_t cit = 0, nz = 0;
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
        const int tile = it / p.nsplit, split = it - tile * p.nsplit;
        const int c0 = sch.split_col[split];
        const int g = tile * BM + row_in_tile;
        const bool valid = g < p.R;
        const int q = valid ? g / p.HW : 0, pix = valid ? g - q * p.HW : 0;
        float sc;
        if (kScaleFromSmem) {
            sc = 0.f;
            if (c0 < sch.split_col[split + 1]) {     // the converters publish the row scales of every non-empty item
                named_sync(SCALE_FULL_BAR + (nz & 1), 256);              // 4 converter warps arrive, 4 epilogue warps wait
                sc = s_scale[nz & 1][row_in_tile];
                named_arrive(SCALE_FREE_BAR + (nz & 1), 256);            // the slot may be rewritten (item nz + 2)
                ++nz;
            }
        } else {
            sc = valid ? __ldg(p.scale + g) : 0.f;
        }
        const float sc2 = sc * LOG2E;
        const Chunks ck = make_chunks(sch.split_col[split + 1] - c0, kNch);
        int col = 0;                                // column cursor relative to c0
        int cur_chunk = -1, chunk_begin = 0, chunk_end = 0;
        uint32_t b = 0;
        for (int set = sch.split_set[split]; set < sch.split_set[split + 1]; ++set) {
            const int cnt = sch.count[set];
            const size_t o = ((size_t)q * p.nsets + set) * p.HW + pix;
            if (cnt <= 0) {   // empty grid set: the reference raises (alpmodule.py:68); report, write NaN
                if (valid) {
                    p.scores[o] = CUDART_NAN_F;
                    if (p.assign) p.assign[o] = CUDART_NAN_F;
                }
                if (g == 0) atomicOr(p.status + set, PSAM_SET_EMPTY);
                continue;
            }
            RowAcc a{0.f, 0.f, -CUDART_INF_F, 0};
            const bool is_mask = p.eff_modes[set] == PSAM_MODE_MASK;
            for (int nb = 0; nb < cnt; nb += 16, col += 16) {
                if (col >= chunk_end) {             // next accumulator chunk (16-column groups never straddle chunks)
                    if (cur_chunk >= 0) {           // done with the previous accumulator buffer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&s_tempty[b]);
                        ++cit;
                    }
                    ++cur_chunk;
                    chunk_begin = chunk_end;
                    chunk_end += chunk_width(ck, cur_chunk);
                    b = cit & 1;
                    mbar_wait_relaxed(&s_tfull[b], (cit >> 1) & 1);
                    tc_fence_after();
                }
                float v[16];
                tc_ld16(lane_addr + b * kNch + (col - chunk_begin), v);
                const int nvalid = cnt - nb;
                // the engine path asks for no `assign`: grid sets then need no maximum at all, 'mask' sets no exponential
                if (is_mask) fold16_any<false, true, false>(v, nvalid, nb, sc2, a);
                else if (want_index) fold16_any<true, true, true>(v, nvalid, nb, sc2, a);
                else fold16_any<true, false, false>(v, nvalid, nb, sc2, a);
            }
            if (valid) {
                if (is_mask) {
                    const float d = a.best * sc;
                    p.scores[o] = d;
                    if (p.assign) p.assign[o] = d;
                } else {
                    p.scores[o] = a.sed * sc / a.se;
                    if (p.assign) p.assign[o] = (float)a.bi;
                }
            }
        }
        if (cur_chunk >= 0) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_tempty[b]);
            ++cit;
        }
    }
}

// ---- algo 2: both operands arrive as pre-packed images through the TMA engine (k_pack_query ran before) ----
// STAGES operand stages of 48 KB: 4 fill the SM's shared memory (fastest GEMM when it runs alone); 3 leave ~80 KB, which
// is what the ALU-bound prompt kernels of ANOTHER volume (k_blocks_warp, k_components) need to be resident on the same SM
// while the tensor pipe works -- the step is then max(tensor, ALU) instead of their sum.
template <int STAGES>
__global__ void __launch_bounds__(THREADS, 1) k_match_tc(const TcParams p)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t s_full[STAGES], s_empty[STAGES], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem;
    __shared__ Schedule sch;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    TraceRec* tr = threadIdx.x == 0 ? trace_begin(1) : nullptr;

    if (threadIdx.x == 0) {
        make_schedule(p, sch);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_tfull[i], 1); mbar_init(&s_tempty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    // item = (tile, split), splits of one tile adjacent: they run on neighbouring CTAs at the same time, so the
    // second read of the tile's operand block hits L2
    const int nitems = p.ntiles * p.nsplit;

    if (warp == 0) {
        // ===== producer: operand blocks global -> shared =====
        if (lane == 0) {
            uint32_t kit = 0;
            for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
                const int tile = it / p.nsplit, split = it - tile * p.nsplit;
                const int c0 = sch.split_col[split], c1 = sch.split_col[split + 1];
                const uint8_t* a_tile = p.a_img + (size_t)tile * p.KB * A_STAGE_BYTES;
                const Chunks ck = make_chunks(c1 - c0, NCH);
                for (int j = 0; j < ck.n; ++j) {
                    const int n0 = c0 + chunk_start(ck, j);
                    const uint32_t bytes_b = (uint32_t)chunk_width(ck, j) * (GROUP_BYTES / 8);
                    for (int kb = 0; kb < p.KB; ++kb, ++kit) {
                        const uint32_t s = kit % STAGES, ph = (kit / STAGES) & 1;
                        mbar_wait_relaxed(&s_empty[s], ph ^ 1);
                        uint8_t* sa = smem + s * STAGE_BYTES;
                        mbar_expect_tx(&s_full[s], A_STAGE_BYTES + bytes_b);
                        bulk_g2s(sa, a_tile + (size_t)kb * A_STAGE_BYTES, A_STAGE_BYTES, &s_full[s]);
                        bulk_g2s(sa + A_STAGE_BYTES, p.b_img + ((size_t)kb * p.G + (n0 >> 3)) * GROUP_BYTES, bytes_b,
                                 &s_full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t kit = 0, cit = 0;
            for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
                const int split = it % p.nsplit;
                const int c0 = sch.split_col[split], c1 = sch.split_col[split + 1];
                const Chunks ck = make_chunks(c1 - c0, NCH);
                for (int j = 0; j < ck.n; ++j, ++cit) {
                    const uint32_t b = cit & 1, tph = (cit >> 1) & 1;
                    mbar_wait(&s_tempty[b], tph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + b * NCH;
                    const uint32_t idesc = make_idesc(chunk_width(ck, j));
                    for (int kb = 0; kb < p.KB; ++kb, ++kit) {
                        const uint32_t s = kit % STAGES, ph = (kit / STAGES) & 1;
                        mbar_wait(&s_full[s], ph);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES), b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t a_hi = make_sdesc(a_addr + k * 32), a_lo = make_sdesc(a_addr + PLANE_BYTES + k * 32);
                            const uint64_t b_hi = make_sdesc(b_addr + k * 32), b_lo = make_sdesc(b_addr + PLANE_BYTES + k * 32);
                            tc_mma(d_tmem, a_hi, b_hi, idesc, (kb | k) != 0);
                            tc_mma(d_tmem, a_lo, b_hi, idesc, 1);
                            tc_mma(d_tmem, a_hi, b_lo, idesc, 1);
                        }
                        tc_commit(&s_empty[s]);        // frees the smem stage when these MMAs retire
                    }
                    tc_commit(&s_tfull[b]);            // accumulator complete
                }
            }
        }
    } else {
        epilogue_loop<NCH, false>(p, sch, tmem_base, warp, lane, s_tfull, s_tempty, nullptr);
    }

    tc_fence_before();
    __syncthreads();
    trace_end(tr);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- algo 3: the fp32 -> bf16 hi/lo split of the query happens inside the GEMM, the A operand lives in TMEM ----
// Stage s of the ring = [prototype block, TS_NCH columns, swizzled bf16 hi/lo | RAW fp32 rows of the query tile, one
// 32-channel k-block, as the TMA unit delivers a [32 channels x 128 rows] box of the 2-D tensor map over the query matrix
// with the 128-byte swizzle].  (One cp.async.bulk per ROW was measured first: 128 small copies per k-block cost ~33 ns
// each and made the kernel 8x slower; the tensor map needs dense slices, slice_stride == HW * row_stride.)  Per k-block:
//   warp 0 (1 thread)  waits empty[s]; one cp.async.bulk for the prototype block + one cp.async.bulk.tensor for the
//                      query box, both counted in bytes on full[s];
//   warps 6-9          thread = row: wait full[s], 8 x LDS.128 of its row (conflict-free thanks to the swizzle), row
//                      norm, bf16 hi/lo split in registers, wait a_empty[t], tcgen05.st of 16 hi + 16 lo columns into
//                      TMEM operand slot t, tcgen05.wait::st, arrive a_full[t];
//   warp 1 (1 thread)  wait full[s] + a_full[t]; 6 MMAs (hi*hi, lo*hi, hi*lo for both 16-deep halves) with A from
//                      TMEM and B from shared memory; commit -> empty[s], commit -> a_empty[t];
//   warps 2-5          epilogue as in k_match_tc, on 224-column accumulators, scales from the converters.
template <int STAGES, int TS_NCH>
__global__ void __launch_bounds__(TS_THREADS, 1) k_match_ts(const TcParams p, const __grid_constant__ CUtensorMap qmap)
{
    constexpr int TS_A_SLOTS = ts_slots(TS_NCH), TS_A_COL0 = 2 * TS_NCH;
    constexpr int TS_B_BYTES = ts_b_bytes(TS_NCH), TS_STAGE_BYTES = ts_stage_bytes(TS_NCH);
    static_assert(TS_NCH % 16 == 0 && TS_A_SLOTS >= 2 && TS_STAGE_BYTES % 1024 == 0, "stage / TMEM layout");
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t s_full[STAGES], s_empty[STAGES], s_tfull[2], s_tempty[2];
    __shared__ uint64_t s_afull[TS_A_SLOTS], s_aempty[TS_A_SLOTS];
    __shared__ float s_scale[2][BM];
    __shared__ uint32_t s_tmem;
    __shared__ Schedule sch;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    TraceRec* tr = threadIdx.x == 0 ? trace_begin(1) : nullptr;

    if (threadIdx.x == 0) {
        make_schedule(p, sch);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_tfull[i], 1);
            mbar_init(&s_tempty[i], 4);
        }
        for (int i = 0; i < TS_A_SLOTS; ++i) { mbar_init(&s_afull[i], TS_CONV_WARPS); mbar_init(&s_aempty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const int nitems = p.ntiles * p.nsplit;

    if (warp == 0) {
        // ===== producer: prototype block + raw query rows, global -> shared =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&qmap) : "memory");
            uint32_t kit = 0;
            for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
                const int tile = it / p.nsplit, split = it - tile * p.nsplit;
                const int c0 = sch.split_col[split], c1 = sch.split_col[split + 1];
                const Chunks ck = make_chunks(c1 - c0, TS_NCH);
                for (int j = 0; j < ck.n; ++j) {
                    const int n0 = c0 + chunk_start(ck, j);
                    const uint32_t bytes_b = (uint32_t)chunk_width(ck, j) * (GROUP_BYTES / 8);
                    for (int kb = 0; kb < p.KB; ++kb, ++kit) {
                        const uint32_t s = kit % STAGES, ph = (kit / STAGES) & 1;
                        mbar_wait_relaxed(&s_empty[s], ph ^ 1);
                        uint8_t* sb = smem + s * TS_STAGE_BYTES;
                        mbar_expect_tx(&s_full[s], bytes_b + TS_RAW_BYTES);      // a box always delivers all its bytes
                        tensor_g2s(sb + TS_B_BYTES, &qmap, kb * BK, tile * BM, &s_full[s]);
                        bulk_g2s(sb, p.b_img + ((size_t)kb * p.G + (n0 >> 3)) * GROUP_BYTES, bytes_b, &s_full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t kit = 0, cit = 0;
            for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
                const int split = it % p.nsplit;
                const int c0 = sch.split_col[split], c1 = sch.split_col[split + 1];
                const Chunks ck = make_chunks(c1 - c0, TS_NCH);
                for (int j = 0; j < ck.n; ++j, ++cit) {
                    const uint32_t b = cit & 1, tph = (cit >> 1) & 1;
                    mbar_wait(&s_tempty[b], tph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + b * TS_NCH;
                    const uint32_t idesc = make_idesc(chunk_width(ck, j));
                    for (int kb = 0; kb < p.KB; ++kb, ++kit) {
                        const uint32_t s = kit % STAGES, ph = (kit / STAGES) & 1;
                        const uint32_t t = kit % TS_A_SLOTS, aph = (kit / TS_A_SLOTS) & 1;
                        mbar_wait(&s_full[s], ph);
                        mbar_wait(&s_afull[t], aph);
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(smem + s * TS_STAGE_BYTES);
                        const uint32_t a_tmem = tmem_base + TS_A_COL0 + t * 32;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint32_t a_hi = a_tmem + k * 8, a_lo = a_tmem + 16 + k * 8;
                            const uint64_t b_hi = make_sdesc(b_addr + k * 32), b_lo = make_sdesc(b_addr + PLANE_BYTES + k * 32);
                            tc_mma_ts(d_tmem, a_hi, b_hi, idesc, (kb | k) != 0);
                            tc_mma_ts(d_tmem, a_lo, b_hi, idesc, 1);
                            tc_mma_ts(d_tmem, a_hi, b_lo, idesc, 1);
                        }
                        tc_commit(&s_empty[s]);        // frees the smem stage ...
                        tc_commit(&s_aempty[t]);       // ... and the TMEM operand slot when these MMAs retire
                    }
                    tc_commit(&s_tfull[b]);            // accumulator complete
                }
            }
        }
    } else if (warp < 6) {
        epilogue_loop<TS_NCH, true>(p, sch, tmem_base, warp, lane, s_tfull, s_tempty, s_scale);
    } else {
        // ===== converters: raw fp32 row (shared memory) -> bf16 hi/lo -> TMEM operand slot =====
        const int lg = warp & 3, row = lg * 32 + lane;               // a warp may only touch its own TMEM lane quarter
        const uint32_t a_lane = tmem_base + ((uint32_t)(lg * 32) << 16) + TS_A_COL0;
        uint32_t kit = 0, nz = 0;
        for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
            const int tile = it / p.nsplit, split = it - tile * p.nsplit;
            const int c0 = sch.split_col[split], c1 = sch.split_col[split + 1];
            if (c0 >= c1) continue;
            const bool valid = tile * BM + row < p.R;
            const int nchunks = make_chunks(c1 - c0, TS_NCH).n;
            float ssq = 0.f;
            // one k-block: this thread's row of the box -> 16 hi + 16 lo packed columns in TMEM operand slot t
            auto convert = [&](auto first_chunk) {
                constexpr bool kNorm = decltype(first_chunk)::value;
                const uint32_t s = kit % STAGES, ph = (kit / STAGES) & 1;
                const uint32_t t = kit % TS_A_SLOTS, aph = (kit / TS_A_SLOTS) & 1;
                ++kit;
                mbar_wait_relaxed(&s_full[s], ph);
                // rows beyond R and channels beyond C arrive as zeros; chunk c of row r sits at chunk c ^ (r & 7)
                const float4* raw = reinterpret_cast<const float4*>(smem + s * TS_STAGE_BYTES + TS_B_BYTES + row * TS_RAW_PITCH);
                // two halves of 16 channels: 8 hi + 8 lo columns each (TMEM slot: hi columns 0-15, lo columns 16-31), so
                // that at most 16 packed values + 16 raw floats are live
                float4 x[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) x[c] = raw[c ^ (row & 7)];
                mbar_wait_relaxed(&s_aempty[t], aph ^ 1);
                tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (kNorm) ssq = fmaf(x[c].x, x[c].x, fmaf(x[c].y, x[c].y, fmaf(x[c].z, x[c].z, fmaf(x[c].w, x[c].w, ssq))));
                        split2(x[c].x, x[c].y, hi[2 * c], lo[2 * c]);
                        split2(x[c].z, x[c].w, hi[2 * c + 1], lo[2 * c + 1]);
                    }
                    if (half == 0) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) x[c] = raw[(4 + c) ^ (row & 7)];
                    }
                    tc_st8(a_lane + t * 32 + half * 8, hi);
                    tc_st8(a_lane + t * 32 + 16 + half * 8, lo);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_afull[t]);
            };
            for (int kb = 0; kb < p.KB; ++kb) convert(std::true_type{});
            if (nz >= 2) named_sync(SCALE_FREE_BAR + (nz & 1), 256);    // the epilogue has read item nz - 2's scales
            s_scale[nz & 1][row] = valid ? 20.0f / fmaxf(sqrtf(ssq), 1e-4f) : 0.f;
            named_arrive(SCALE_FULL_BAR + (nz & 1), 256);
            ++nz;
            for (int rest = (nchunks - 1) * p.KB; rest > 0; --rest) convert(std::false_type{});
        }
    }

    tc_fence_before();
    __syncthreads();
    trace_end(tr);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

static int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

struct Layout {
    int R, ntiles, KB, G;
    size_t a_bytes, b_bytes, scale_bytes;
};

static Layout make_layout(int Q, int HW, int C, int nsets, int cap_rows)
{
    Layout L;
    L.R = Q * HW;
    L.ntiles = (L.R + BM - 1) / BM;
    L.KB = (C + BK - 1) / BK;
    L.G = nsets * pad16(cap_rows) / 8;
    L.a_bytes = (size_t)L.ntiles * L.KB * A_STAGE_BYTES;
    L.b_bytes = (size_t)L.KB * L.G * GROUP_BYTES;
    L.scale_bytes = (size_t)L.ntiles * BM * sizeof(float);
    return L;
}

}  // namespace tc

// SMs the persistent GEMM grids leave free (psam_match_reserve_sms); n < 0 queries
int match_reserve_sms(int n)
{
    static std::atomic<int> v{-1};
    int cur = v.load(std::memory_order_relaxed);
    if (cur < 0) {
        cur = 0;
        if (const char* ov = getenv("PSAM_TC_RESERVE_SMS")) cur = std::max(0, atoi(ov));
        v.store(cur, std::memory_order_relaxed);
    }
    if (n >= 0) v.store(n, std::memory_order_relaxed);
    return cur;
}

PSAM_TRACE_TU();
bool match_tc_supported(int Q, int HW, int C, int nsets, int cap_rows, bool want_sims)
{
    (void)cap_rows;
    return !want_sims && C % 8 == 0 && nsets <= tc::MAX_SETS && (long long)Q * HW < (1ll << 30);
}

bool match_ts_supported(const MatchParams& p)
{
    return p.Q == 1 || p.slice_stride == (int64_t)p.HW * p.row_stride;
}

// Which tensor-core variant algo 0 takes.  The fused kernel converts the query tile once per column chunk of the work item,
// the packed-operand path once per launch (k_pack_query): with many chunks per item the conversions, the narrower
// accumulators (224 instead of 256 columns) and the fused kernel's larger footprint beside the prompt kernels outweigh the
// saved pass.  Measured on one B200 (ms per volume, fused / packed): config 2 (1.3 k columns, 2 x 3 chunks) 0.31 / 0.35,
// config 5 (0.6 k, 3 chunks) 7.7 / 10.1, config 4 (1.3 k, 6 chunks) 2.01 / 2.13, config 3 (2.3 k, 11 chunks) 3.57 / 3.22.
// The live column count is device-side data; the host only knows the table's capacity, of which the grid sets typically
// fill about half (background sets most of their windows, foreground sets few): capacity >= 4096 columns -> packed.
bool match_ts_preferred(const MatchParams& p)
{
    static int thresh = 0;
    if (thresh == 0) {
        thresh = 4096;
        if (const char* ov = getenv("PSAM_TS_MAX_COLUMNS")) thresh = std::max(1, atoi(ov));
    }
    return (long long)p.nsets * tc::pad16(p.cap_rows) < thresh;
}

size_t match_tc_workspace(int Q, int HW, int C, int nsets, int cap_rows, bool fused)
{
    const tc::Layout L = tc::make_layout(Q, HW, C, nsets, cap_rows);
    if (fused) return align_up(L.b_bytes, 1024) + 1024;
    return align_up(L.a_bytes, 1024) + align_up(L.b_bytes, 1024) + align_up(L.scale_bytes, 1024) + 1024;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [R rows x C channels] fp32, row pitch row_stride floats, boxes of 32 channels x 128 rows, 128-byte swizzle
static int make_query_map(const MatchParams& p, int R, CUtensorMap* map)
{
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        set_error("psam_alp_match: cuTensorMapEncodeTiled is not available from this driver");
        return PSAM_ERR_UNSUPPORTED;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)p.C, (cuuint64_t)R};
    const cuuint64_t gstride[1] = {(cuuint64_t)p.row_stride * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)tc::BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.qry), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("psam_alp_match: cuTensorMapEncodeTiled failed (%d) for C=%d R=%d row_stride=%lld", (int)r, p.C, R,
                  (long long)p.row_stride);
        return PSAM_ERR_LAUNCH;
    }
    return PSAM_OK;
}

template <typename K, typename... Extra>
static int launch_gemm(K kernel, const char* name, int threads, int smem, const tc::TcParams& t, int grid, cudaStream_t stream,
                       bool* attr_set, const Extra&... extra)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {       // once per device
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("%s: cudaFuncSetAttribute(%d bytes): %s", name, smem, cudaGetErrorString(e));
            return PSAM_ERR_LAUNCH;
        }
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    PSAM_PROF_BEGIN(stream);
    kernel<<<grid, threads, smem, stream>>>(t, extra...);
    PSAM_CHECK_LAUNCH(name);
    return PSAM_OK;
}

// operand stages of the GEMMs: 3 by default (co-residency, see k_match_tc); PSAM_TC_STAGES=4 is the experiment knob for
// the stand-alone optimum
static int gemm_stages()
{
    using tc::MAX_STAGES;
    static int v = 0;
    if (v == 0) {
        v = MAX_STAGES >= 4 ? 3 : MAX_STAGES;
        if (const char* ov = getenv("PSAM_TC_STAGES")) {
            const int x = atoi(ov);
            if (x == 3 || x == 4) v = x <= MAX_STAGES ? x : MAX_STAGES;
        }
    }
    return v;
}

int launch_match_tc(const MatchParams& p, void* workspace, size_t workspace_bytes, bool fused, cudaStream_t stream)
{
    using namespace tc;
    if (!match_tc_supported(p.Q, p.HW, p.C, p.nsets, p.cap_rows, p.sims != nullptr)) {
        set_error("psam_alp_match: the tensor-core variant needs C %% 8 == 0, nsets <= %d and sims == NULL", MAX_SETS);
        return PSAM_ERR_UNSUPPORTED;
    }
    if (fused && !match_ts_supported(p)) {
        set_error("psam_alp_match: algo 3 reads the query through a 2-D tensor map and needs dense slices "
                  "(slice_stride == HW * row_stride)");
        return PSAM_ERR_UNSUPPORTED;
    }
    const size_t need = match_tc_workspace(p.Q, p.HW, p.C, p.nsets, p.cap_rows, fused);
    if (!workspace || workspace_bytes < need) {
        set_error("psam_alp_match: workspace too small for the tensor-core variant (%zu < %zu)", workspace_bytes, need);
        return PSAM_ERR_WORKSPACE;
    }
    const Layout L = make_layout(p.Q, p.HW, p.C, p.nsets, p.cap_rows);
    uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
    uint8_t* b_img = base;
    uint8_t* a_img = fused ? nullptr : b_img + align_up(L.b_bytes, 1024);
    float* scale = fused ? nullptr : reinterpret_cast<float*>(a_img + align_up(L.a_bytes, 1024));
    // the tensor map of the fused kernel is host work: built before anything is enqueued, so that a failure (driver
    // without cuTensorMapEncodeTiled, a shape the encoder rejects) leaves no trace and auto can take the packed path
    CUtensorMap qmap;
    if (fused) {
        if (int rc = make_query_map(p, L.R, &qmap)) return rc == PSAM_ERR_LAUNCH ? PSAM_ERR_UNSUPPORTED : rc;
    }

    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_pack_protos);
    k_pack_protos<<<dim3(pad16(p.cap_rows) / 8, p.nsets), 256, 0, stream>>>(p.protos, p.cap_rows, p.counts, p.C, L.KB, L.G,
                                                                           b_img);
    PSAM_CHECK_LAUNCH("k_pack_protos");
    if (!fused) {
        PSAM_PROF_BEGIN(stream);
        PSAM_MAX_CARVEOUT(k_pack_query);
        k_pack_query<<<L.ntiles * BM / 8, 256, 0, stream>>>(p.qry, p.slice_stride, p.row_stride, p.HW, L.R, p.C, L.KB,
                                                           a_img, scale);
        PSAM_CHECK_LAUNCH("k_pack_query");
    }
    const int sms = sm_count();
    int nsplit = (4 * sms + L.ntiles - 1) / L.ntiles;
    nsplit = max(1, min(min(nsplit, p.nsets), MAX_SPLIT));
    if (const char* ov = getenv("PSAM_TC_NSPLIT")) {      // experiment knob: column splits per row tile
        const int v = atoi(ov);
        if (v >= 1) nsplit = min(min(v, p.nsets), MAX_SPLIT);
    }
    TcParams t{a_img, b_img, scale, p.counts, p.eff_modes, p.scores, p.assign, p.status,
               p.nsets, p.HW, L.R, L.ntiles, L.KB, L.G, nsplit, p.qry, p.slice_stride, p.row_stride, p.C};
    // Grid: the items are equal-sized, so a launch takes ceil(items / CTAs) rounds whatever the CTA count inside a round
    // bracket; take the FEWEST CTAs that keep the round count (config 2: 686 items = 5 rounds on 138..148 CTAs -> 138).
    // The SMs left over carry only short-lived CTAs: room for the prompt kernels and the NCCL kernels of the other
    // volumes in flight at no cost to the GEMM.  psam_match_reserve_sms() additionally caps the CTA count.
    const int reserve = max(0, min(match_reserve_sms(-1), sms - 1));
    const int nitems = L.ntiles * nsplit, cap = min(sms - reserve, nitems);
    const int rounds = (nitems + cap - 1) / cap;
    int grid = (nitems + rounds - 1) / rounds;
    if (getenv("PSAM_TC_FULL_GRID")) grid = cap;          // experiment knob
    constexpr int S3 = MAX_STAGES >= 4 ? 3 : MAX_STAGES;
    const bool three = MAX_STAGES >= 4 && gemm_stages() == 3;
    static bool set_tc3[64] = {}, set_tc4[64] = {};
    if (fused) {
        // 4 stages of 44 KB by default: unlike the packed kernel, a stage's cycle includes the conversion, and the fourth
        // stage is worth 12 % to the kernel and 4 % to the pipelined step (measured); PSAM_TS_STAGES / PSAM_TS_NCH are
        // experiment knobs
        static int st = 0, nch = 0;
        if (st == 0) {
            st = 4; nch = 224;
            if (const char* ov = getenv("PSAM_TS_STAGES")) st = atoi(ov);
            if (const char* ov = getenv("PSAM_TS_NCH")) nch = atoi(ov);
        }
#define PSAM_TS_CASE(ST, NC)                                                                                      \
    if (st == ST && nch == NC) {                                                                                  \
        static bool done[64] = {};                                                                                \
        PSAM_MAX_CARVEOUT((k_match_ts<ST, NC>));                                                                  \
        return launch_gemm(k_match_ts<ST, NC>, "k_match_ts", TS_THREADS, ts_smem_bytes(ST, NC), t, grid, stream, done, qmap); \
    }
        PSAM_TS_CASE(3, 224) PSAM_TS_CASE(5, 224)
        PSAM_TS_CASE(3, 208) PSAM_TS_CASE(4, 208) PSAM_TS_CASE(5, 208)
        PSAM_TS_CASE(3, 192) PSAM_TS_CASE(4, 192) PSAM_TS_CASE(5, 192)
        st = 4; nch = 224;
        PSAM_TS_CASE(4, 224)
#undef PSAM_TS_CASE
    }
    if (three) {
        PSAM_MAX_CARVEOUT(k_match_tc<S3>);
        return launch_gemm(k_match_tc<S3>, "k_match_tc", THREADS, smem_bytes(S3), t, grid, stream, set_tc3);
    }
    PSAM_MAX_CARVEOUT(k_match_tc<MAX_STAGES>);
    return launch_gemm(k_match_tc<MAX_STAGES>, "k_match_tc", THREADS, smem_bytes(MAX_STAGES), t, grid, stream, set_tc4);
}

}  // namespace psam

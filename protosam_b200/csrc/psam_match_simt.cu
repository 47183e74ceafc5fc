// Kernel 2 (CUDA-core variant) -- fused query x prototype match (sm_100a).
//
// Replaces safe_norm(qry) + get_prediction_from_prototypes of the reference
// (models/alpmodule.py:57-94, 195): L2-normalise every query pixel (clamp 1e-4), contract
// [HW x C] . [C x P] against the normalised prototypes of a set, scale by 20, and reduce over
// the prototype axis in the epilogue -- softmax-weighted sum + argmax for the grid modes, max
// for 'mask' -- so the [P, HW] similarity tensor the reference materialises three times
// never leaves the registers (unless `sims` is requested for visualisation).
//
// This is the exact-fp32 variant: classic 128x64x16 register-tiled SGEMM, 256 threads, 8x4
// micro-tile, register-prefetch double buffering.  It is the parity reference on the device
// and the path for shapes the tensor-core variant (psam_match_tc.cu) does not take.
//
// Roofline: compute-bound for all named configs (arithmetic intensity = sum(P)/2 FLOP/B,
// SURVEY.md section 8(d)); fp32 FFMA peak on B200 = 148 SM * 128 lanes * 2 * 1.965 GHz = 74.4
// TFLOP/s.  Algorithmic flops per (slice, set) = 2*HW*C*P.
//
// Because |d| <= 20 (both operands are unit vectors or shorter), softmax uses the fixed
// reference point 20 instead of a running maximum: e = exp(d - 20) >= exp(-40) stays normal
// in fp32 and no rescaling pass is needed.
#include <math_constants.h>

#include "psam_match.cuh"

// algo = 0 (auto) takes the fused tensor-core variant (query converted inside the GEMM, A operand in TMEM) whenever the
// slices are dense, else the packed-operand variant; -DPSAM_AUTO_FUSED=0 restores the packed variant as the default
#ifndef PSAM_AUTO_FUSED
#define PSAM_AUTO_FUSED 1
#endif

namespace psam {

constexpr int BM = 128, BN = 64, BK = 16, APAD = 4;

__global__ void __launch_bounds__(256) k_match_simt(MatchParams p)
{
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN];
    __shared__ float s_scale[BM];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, set = blockIdx.y, q = blockIdx.z;
    const int count = p.counts[set], mode = p.eff_modes[set];
    const int HW = p.HW, C = p.C;
    float* out_scores = p.scores + ((size_t)q * p.nsets + set) * HW;
    float* out_assign = p.assign ? p.assign + ((size_t)q * p.nsets + set) * HW : nullptr;

    if (count <= 0) {  // empty grid set: the reference raises (alpmodule.py:68); report, write NaN
        for (int i = tid; i < BM; i += 256)
            if (m0 + i < HW) {
                out_scores[m0 + i] = CUDART_NAN_F;
                if (out_assign) out_assign[m0 + i] = CUDART_NAN_F;
            }
        if (tid == 0 && blockIdx.x == 0 && q == 0) atomicOr(p.status + set, PSAM_SET_EMPTY);
        return;
    }

    const float* A = p.qry + (size_t)q * p.slice_stride;
    const float* B = p.protos + (size_t)set * p.cap_rows * C;
    float* out_sims = p.sims ? p.sims + ((size_t)q * p.nsets + set) * p.cap_rows * HW : nullptr;

    // global -> register staging assignments
    const int a_r0 = tid >> 2, a_kc = tid & 3;  // rows a_r0 and a_r0 + 64, k offset a_kc*4
    const int b_n = tid >> 2, b_kc = tid & 3;   // proto row b_n, k offset b_kc*4

    float se[8], sed[8], bv[8];
    int bi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { se[i] = 0.f; sed[i] = 0.f; bv[i] = -CUDART_INF_F; bi[i] = 0; }
    float ssq0 = 0.f, ssq1 = 0.f;
    float scale[8];

    const int ktiles = (C + BK - 1) / BK;
    const int ntiles = (count + BN - 1) / BN;

    for (int nt = 0; nt < ntiles; ++nt) {
        const int n0 = nt * BN;
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

        float4 ra0, ra1, rb;
        auto gload = [&](int kt) {
            const int k = kt * BK + a_kc * 4;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            ra0 = (m0 + a_r0 < HW && k < C) ? __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + a_r0) * p.row_stride + k)) : z;
            ra1 = (m0 + a_r0 + 64 < HW && k < C) ? __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + a_r0 + 64) * p.row_stride + k)) : z;
            const int kb = kt * BK + b_kc * 4;
            rb = (n0 + b_n < count && kb < C) ? __ldg(reinterpret_cast<const float4*>(B + (size_t)(n0 + b_n) * C + kb)) : z;
        };
        auto sstore = [&](int buf) {
            As[buf][a_kc * 4 + 0][a_r0] = ra0.x; As[buf][a_kc * 4 + 1][a_r0] = ra0.y;
            As[buf][a_kc * 4 + 2][a_r0] = ra0.z; As[buf][a_kc * 4 + 3][a_r0] = ra0.w;
            As[buf][a_kc * 4 + 0][a_r0 + 64] = ra1.x; As[buf][a_kc * 4 + 1][a_r0 + 64] = ra1.y;
            As[buf][a_kc * 4 + 2][a_r0 + 64] = ra1.z; As[buf][a_kc * 4 + 3][a_r0 + 64] = ra1.w;
            Bs[buf][b_kc * 4 + 0][b_n] = rb.x; Bs[buf][b_kc * 4 + 1][b_n] = rb.y;
            Bs[buf][b_kc * 4 + 2][b_n] = rb.z; Bs[buf][b_kc * 4 + 3][b_n] = rb.w;
            if (nt == 0) {
                ssq0 += ra0.x * ra0.x + ra0.y * ra0.y + ra0.z * ra0.z + ra0.w * ra0.w;
                ssq1 += ra1.x * ra1.x + ra1.y * ra1.y + ra1.z * ra1.z + ra1.w * ra1.w;
            }
        };

        gload(0);
        sstore(0);
        __syncthreads();
        for (int kt = 0; kt < ktiles; ++kt) {
            const int buf = kt & 1;
            if (kt + 1 < ktiles) gload(kt + 1);
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bw[j], acc[i][j]);
            }
            if (kt + 1 < ktiles) sstore(buf ^ 1);
            __syncthreads();
        }

        if (nt == 0) {
            // row norms: 4 consecutive lanes hold the partial sums of one row
            ssq0 += __shfl_xor_sync(0xffffffffu, ssq0, 1);
            ssq0 += __shfl_xor_sync(0xffffffffu, ssq0, 2);
            ssq1 += __shfl_xor_sync(0xffffffffu, ssq1, 1);
            ssq1 += __shfl_xor_sync(0xffffffffu, ssq1, 2);
            if (a_kc == 0) {
                s_scale[a_r0] = 20.0f / fmaxf(sqrtf(ssq0), 1e-4f);
                s_scale[a_r0 + 64] = 20.0f / fmaxf(sqrtf(ssq1), 1e-4f);
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 8; ++i) scale[i] = s_scale[ty * 8 + i];
        }

        // epilogue: fold this tile's columns into the per-row reductions
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < count) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float d = acc[i][j] * scale[i];
                    const float e = __expf(d - 20.0f);
                    se[i] += e;
                    sed[i] = fmaf(e, d, sed[i]);
                    if (d > bv[i]) { bv[i] = d; bi[i] = n; }
                    if (out_sims && m0 + ty * 8 + i < HW) out_sims[(size_t)n * HW + m0 + ty * 8 + i] = d;
                }
            }
        }
        __syncthreads();
    }

    // combine the 16 column groups of each row (lanes tx = 0..15 of a half warp)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            se[i] += __shfl_xor_sync(0xffffffffu, se[i], o);
            sed[i] += __shfl_xor_sync(0xffffffffu, sed[i], o);
            const float ov = __shfl_xor_sync(0xffffffffu, bv[i], o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi[i], o);
            if (ov > bv[i] || (ov == bv[i] && oi < bi[i])) { bv[i] = ov; bi[i] = oi; }
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + ty * 8 + i;
            if (m < HW) {
                if (mode == PSAM_MODE_MASK) {
                    out_scores[m] = bv[i];
                    if (out_assign) out_assign[m] = bv[i];
                } else {
                    out_scores[m] = sed[i] / se[i];
                    if (out_assign) out_assign[m] = (float)bi[i];
                }
            }
        }
    }
}

int launch_match_simt(const MatchParams& p, cudaStream_t stream)
{
    dim3 grid((p.HW + BM - 1) / BM, p.nsets, p.Q);
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_match_simt);
    k_match_simt<<<grid, 256, 0, stream>>>(p);
    PSAM_CHECK_LAUNCH("k_match_simt");
    return PSAM_OK;
}

}  // namespace psam

using namespace psam;

extern "C" size_t psam_alp_match_workspace(int Q, int HW, int C, int nsets, int cap_rows, int algo)
{
    // algo 1 needs no scratch; 2 / 3 stage bf16 operand images (3 = fused: of the prototypes only); 0 (auto) may run
    // either tensor-core variant, so it asks for the larger one
    if (algo == 1 || Q < 1 || HW < 1 || C < 1 || nsets < 1 || cap_rows < 1) return 256;
    if (!match_tc_supported(Q, HW, C, nsets, cap_rows, false)) return 256;
    return match_tc_workspace(Q, HW, C, nsets, cap_rows, algo == 3);
}

extern "C" int psam_match_reserve_sms(int n) { return match_reserve_sms(n); }

extern "C" int psam_alp_match(const float* qry, int64_t slice_stride, int64_t row_stride, int Q, int HW, int C,
                              const float* protos, int cap_rows, const int32_t* counts, const int32_t* eff_modes,
                              int nsets, float* scores, float* assign, float* sims, int32_t* status, void* workspace,
                              size_t workspace_bytes, int algo, psam_stream_t stream_)
{
    PSAM_TRACE("psam_alp_match");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(qry && protos && counts && eff_modes && scores && status, "psam_alp_match: null pointer");
    PSAM_CHECK_ARG(Q >= 1 && Q <= 65535 && HW >= 1 && C >= 1 && nsets >= 1 && nsets <= 65535 && cap_rows >= 1,
                   "psam_alp_match: bad shape Q=%d HW=%d C=%d nsets=%d", Q, HW, C, nsets);
    PSAM_CHECK_ARG(C % 4 == 0 && row_stride % 4 == 0 && slice_stride % 4 == 0 && row_stride >= C,
                   "psam_alp_match: C, row_stride and slice_stride must be multiples of 4 floats (16-byte rows)");
    PSAM_CHECK_ARG((reinterpret_cast<uintptr_t>(qry) & 15) == 0 && (reinterpret_cast<uintptr_t>(protos) & 15) == 0,
                   "psam_alp_match: qry/protos must be 16-byte aligned");
    PSAM_CHECK_ARG(algo >= 0 && algo <= 3, "psam_alp_match: algo %d", algo);
    MatchParams p{qry, slice_stride, row_stride, Q, HW, C, protos, cap_rows, counts, eff_modes,
                  nsets, scores, assign, sims, status};
    if (algo == 2 || algo == 3) return launch_match_tc(p, workspace, workspace_bytes, algo == 3, stream);
    // auto: tensor cores whenever the variant applies (the raw-similarity dump for visualisation and channel
    // counts that are not a multiple of 8 stay on the CUDA-core kernel) and the caller sized the workspace for it
    if (algo == 0 && match_tc_supported(Q, HW, C, nsets, cap_rows, sims != nullptr) && workspace) {
        if (PSAM_AUTO_FUSED && match_ts_supported(p) && match_ts_preferred(p) &&
            workspace_bytes >= match_tc_workspace(Q, HW, C, nsets, cap_rows, true)) {
            const int rc = launch_match_tc(p, workspace, workspace_bytes, true, stream);
            if (rc != PSAM_ERR_UNSUPPORTED) return rc;          // no tensor map for this shape / driver: nothing was enqueued
        }
        if (workspace_bytes >= match_tc_workspace(Q, HW, C, nsets, cap_rows, false))
            return launch_match_tc(p, workspace, workspace_bytes, false, stream);
    }
    return launch_match_simt(p, stream);
}

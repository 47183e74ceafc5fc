// Kernel 1 -- ALP prototype pass (sm_100a).
//
// Replaces MultiProtoAsConv.get_prototypes + safe_norm of the reference
// (models/alpmodule.py:14-18, 97-159): masked grid average pooling, the foreground-fraction
// threshold, ORDERED compaction of the surviving local prototypes ((shot,gy,gx) row-major,
// which feeds debug_assign / proto_grid numbering), the global masked-average prototype and
// the L2 normalisation, for many prototype sets that share one support feature tensor.
//
// Roofline: HBM-bound and tiny -- one read of the S*h*w*C support features per set
// (4.2 MB at ViT-B/14 37x37) once per VOLUME, against Q reads of the same size in kernel 2.
// Algorithmic bytes per set: 4*S*h*w*(C+1) read + 4*P*C written.  What matters here is
// (a) no host synchronisation (the reference syncs four times: nonzero x2, boolean index,
// .max() >= thresh) and (b) coalesced channels-last reads.
#include "psam_common.cuh"

namespace psam {

constexpr int kMaxSets = 128;
struct SetModes {
    int8_t m[kMaxSets];
    int8_t shot[kMaxSets];   // -1: the set uses every shot; k: only shot k (FewShotSeg's per-shot foreground calls)
};

// ------------------------------------------------------------------------------------
// K1a: per set -- window mask fractions (exact ATen order: sequential (dy,dx) sum, true
// division), AUTO_FG decision, survive flags, ordered row indices, per-shot mask sums.
// One CTA of 256 threads per set (CTA 0 of each set's row in k_proto_stage1).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void pool_mask_body(int set, const float* __restrict__ sup_y, SetModes modes, int S, int h,
                                                   int w, int kh, int kw, int akh, int akw, float thresh,
                                                   float* __restrict__ pooled, uint8_t* __restrict__ survive,
                                                   int32_t* __restrict__ rowidx, float* __restrict__ ysum,
                                                   int32_t* __restrict__ counts, int32_t* __restrict__ eff_modes,
                                                   int32_t* __restrict__ status, int32_t* __restrict__ plocal)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gh = h / kh, gw = w / kw, N = S * gh * gw;
    const float* y = sup_y + (size_t)set * S * h * w;
    __shared__ int s_warp[8];
    __shared__ int s_flag;
    __shared__ float s_wsum[8];

    int mode = modes.m[set];
    const int sel = modes.shot[set];
    if (mode == PSAM_MODE_AUTO_FG) {
        // F.avg_pool2d(mask, kernel_size).max() >= thresh   (grid_proto_fewshot.py:254-256)
        const int agh = h / akh, agw = w / akw, AN = S * agh * agw;
        const float adiv = (float)(akh * akw);
        int hit = 0;
        for (int n = tid; n < AN; n += 256) {
            int s = n / (agh * agw), r = n % (agh * agw), gy = r / agw, gx = r % agw;
            if (sel >= 0 && s != sel) continue;
            const float* p = y + ((size_t)s * h + gy * akh) * w + gx * akw;
            float acc = 0.0f;
            for (int dy = 0; dy < akh; ++dy)
                for (int dx = 0; dx < akw; ++dx) acc = __fadd_rn(acc, p[dy * w + dx]);
            hit |= (__fdiv_rn(acc, adiv) >= thresh);
        }
        hit = __syncthreads_or(hit);
        mode = hit ? PSAM_MODE_GRIDCONV_PLUS : PSAM_MODE_MASK;
    }
    const bool locals = (mode != PSAM_MODE_MASK);
    const bool globals = (mode != PSAM_MODE_GRIDCONV);

    // pooled fraction + survive + ordered compaction index
    const float div = (float)(kh * kw);
    int running = 0;
    for (int base = 0; base < N; base += 256) {
        int n = base + tid, flag = 0;
        if (n < N) {
            int s = n / (gh * gw), r = n % (gh * gw), gy = r / gw, gx = r % gw;
            const float* p = y + ((size_t)s * h + gy * kh) * w + gx * kw;
            float acc = 0.0f;
            for (int dy = 0; dy < kh; ++dy)
                for (int dx = 0; dx < kw; ++dx) acc = __fadd_rn(acc, p[dy * w + dx]);
            float f = __fdiv_rn(acc, div);
            pooled[(size_t)set * N + n] = f;
            flag = locals && (f > thresh) && (sel < 0 || s == sel);
            survive[(size_t)set * N + n] = (uint8_t)(f > thresh);
        }
        int ex = warp_excl_scan_i(flag, lane);
        if (lane == 31) s_warp[wid] = ex + flag;
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int v = s_warp[k];
            if (k < wid) woff += v;
            tot += v;
        }
        if (n < N) rowidx[(size_t)set * N + n] = flag ? running + woff + ex : -1;
        running += tot;
        __syncthreads();
    }

    // per-shot mask sums (denominator of the global prototype, alpmodule.py:100,156)
    for (int s = 0; s < S; ++s) {
        float acc = 0.0f;
        for (int i = tid; i < h * w; i += 256) acc += y[(size_t)s * h * w + i];
        acc = warp_sum(acc);
        if (lane == 0) s_wsum[wid] = acc;
        __syncthreads();
        if (tid == 0) {
            float t = 0.0f;
            for (int k = 0; k < 8; ++k) t += s_wsum[k];
            ysum[set * S + s] = t;
        }
        __syncthreads();
    }
    if (tid == 0) {
        plocal[set] = running;
        counts[set] = running + (globals ? (sel < 0 ? S : 1) : 0);
        eff_modes[set] = mode;
        status[set] = (mode == PSAM_MODE_GRIDCONV && running == 0) ? PSAM_SET_EMPTY : 0;
    }
    (void)s_flag;
}

// ------------------------------------------------------------------------------------
// K1b: one CTA per (window, set): pooled feature vector of a surviving window, L2-normalised
// (clamp 1e-4), written at its compacted row.  Threads run along C (coalesced when the
// tensor is channels-last).  CTAs 0..N-1 of each set's row in k_proto_stage2, block = 256, C <= 256*16.
// ------------------------------------------------------------------------------------
constexpr int kMaxCPerThread = 16;

__device__ __forceinline__ float block_sum_256(float v, float* s_red)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) s_red[wid] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s_red[k];
    __syncthreads();
    return t;
}

// Feature tiles are staged with the TMA engine when the tensor is channels-last (the layout DINOv2 hands over): the
// kh*kw pixels of a window are kh*kw contiguous C-float rows, fetched as kh*kw bulk copies (cp.async.bulk, completion
// counted in bytes on an mbarrier) by one thread while the CTA's other warps only wait; strided layouts and windows
// that do not fit kStageBytes are read with plain coalesced loads.
constexpr int kStageBytes = 40 * 1024;

__device__ __forceinline__ void pool_feat_body(int n, int set, const float* __restrict__ sup_x, int64_t xs_s, int64_t xs_c,
                                               int64_t xs_y, int64_t xs_x, const int32_t* __restrict__ rowidx,
                                               int S, int C, int h, int w, int kh, int kw, int cap_rows,
                                               float* __restrict__ protos, float* s_stage, uint64_t* s_bar)
{
    const int tid = threadIdx.x;
    const int gh = h / kh, gw = w / kw, N = S * gh * gw;
    const int row = rowidx[(size_t)set * N + n];
    if (row < 0) return;
    __shared__ float s_red[8];
    const int s = n / (gh * gw), r = n % (gh * gw), gy = r / gw, gx = r % gw;
    const float* base = sup_x + s * xs_s + (int64_t)(gy * kh) * xs_y + (int64_t)(gx * kw) * xs_x;
    const float div = (float)(kh * kw);
    const bool staged = xs_c == 1 && (C & 3) == 0 && (size_t)kh * kw * C * sizeof(float) <= (size_t)kStageBytes &&
                        ((reinterpret_cast<uintptr_t>(sup_x) | (uintptr_t)(xs_s * 4) | (uintptr_t)(xs_y * 4) |
                          (uintptr_t)(xs_x * 4)) & 15) == 0;
    if (staged) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                         "r"((uint32_t)(kh * kw * C * sizeof(float))) : "memory");
            for (int dy = 0; dy < kh; ++dy)
                for (int dx = 0; dx < kw; ++dx)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     (uint32_t)__cvta_generic_to_shared(s_stage + (size_t)(dy * kw + dx) * C)),
                                 "l"(base + dy * xs_y + dx * xs_x), "r"((uint32_t)(C * sizeof(float))), "r"(bar)
                                 : "memory");
        }
        __syncthreads();                     // the barrier is initialised before anyone polls it
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "WAIT_STAGE:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t"
            "@P1 bra DONE_STAGE;\n\t"
            "bra WAIT_STAGE;\n\t"
            "DONE_STAGE:\n\t"
            "}" ::"r"(bar)
            : "memory");
    }
    float v[kMaxCPerThread];
    float ss = 0.0f;
#pragma unroll
    for (int j = 0; j < kMaxCPerThread; ++j) {
        int c = tid + j * 256;
        v[j] = 0.0f;
        if (c < C) {
            float acc = 0.0f;
            if (staged) {
                for (int k = 0; k < kh * kw; ++k) acc = __fadd_rn(acc, s_stage[(size_t)k * C + c]);
            } else {
                for (int dy = 0; dy < kh; ++dy)
                    for (int dx = 0; dx < kw; ++dx) acc = __fadd_rn(acc, __ldg(base + c * xs_c + dy * xs_y + dx * xs_x));
            }
            v[j] = __fdiv_rn(acc, div);
            ss += v[j] * v[j];
        }
    }
    ss = block_sum_256(ss, s_red);
    float nrm = fmaxf(sqrtf(ss), 1e-4f);
    float* dst = protos + ((size_t)set * cap_rows + row) * C;
#pragma unroll
    for (int j = 0; j < kMaxCPerThread; ++j) {
        int c = tid + j * 256;
        if (c < C) dst[c] = v[j] / nrm;
    }
}

// ------------------------------------------------------------------------------------
// K1c/K1d: global masked-average prototype  sum(x*y) / (sum(y) + 1e-5)  (alpmodule.py:99-100,
// 155-156), deterministic two-stage reduction: partial sums per feature row (CTAs 1..h*S of each set's row in
// k_proto_stage1), then a fixed order combine + safe_norm (CTAs N..N+S-1 of each set's row in k_proto_stage2).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void global_partial_body(int yy, int s, int set, const float* __restrict__ sup_x, int64_t xs_s,
                                                    int64_t xs_c, int64_t xs_y, int64_t xs_x,
                                                    const float* __restrict__ sup_y, const SetModes& modes, int S,
                                                    int C, int h, int w, float* __restrict__ partial)
{
    const int tid = threadIdx.x;
    // runs in the same launch as the AUTO_FG decision, so only sets that are 'gridconv' by request are skipped
    if (modes.m[set] == PSAM_MODE_GRIDCONV) return;
    if (modes.shot[set] >= 0 && s != modes.shot[set]) return;
    const float* m = sup_y + (((size_t)set * S + s) * h + yy) * w;
    const float* base = sup_x + s * xs_s + (int64_t)yy * xs_y;
    float* dst = partial + (((size_t)set * S + s) * h + yy) * C;
    for (int c = tid; c < C; c += 256) {
        float acc = 0.0f;
        for (int xx = 0; xx < w; ++xx) {
            float mv = m[xx];
            if (mv != 0.0f) acc += __ldg(base + c * xs_c + xx * xs_x) * mv;
        }
        dst[c] = acc;
    }
}

__device__ __forceinline__ void global_final_body(int s, int set, const float* __restrict__ partial,
                                                  const float* __restrict__ ysum,
                                                  const int32_t* __restrict__ eff_modes,
                                                  const int32_t* __restrict__ plocal, const SetModes& modes, int S, int C,
                                                  int h, int cap_rows, float* __restrict__ protos)
{
    const int tid = threadIdx.x;
    if (eff_modes[set] == PSAM_MODE_GRIDCONV) return;
    const int sel = modes.shot[set];
    if (sel >= 0 && s != sel) return;
    __shared__ float s_red[8];
    const float* src = partial + ((size_t)set * S + s) * h * C;
    const float den = ysum[set * S + s] + 1e-5f;
    float v[kMaxCPerThread];
    float ss = 0.0f;
#pragma unroll
    for (int j = 0; j < kMaxCPerThread; ++j) {
        int c = tid + j * 256;
        v[j] = 0.0f;
        if (c < C) {
            float acc = 0.0f;
            for (int yy = 0; yy < h; ++yy) acc += src[(size_t)yy * C + c];
            v[j] = acc / den;
            ss += v[j] * v[j];
        }
    }
    ss = block_sum_256(ss, s_red);
    // safe_norm for the grid modes (:158); cosine_similarity's clamp_min(eps=1e-4) for 'mask'
    // (:59) -- the same arithmetic.
    float nrm = fmaxf(sqrtf(ss), 1e-4f);
    float* dst = protos + ((size_t)set * cap_rows + plocal[set] + (sel >= 0 ? 0 : s)) * C;
#pragma unroll
    for (int j = 0; j < kMaxCPerThread; ++j) {
        int c = tid + j * 256;
        if (c < C) dst[c] = v[j] / nrm;
    }
}

// ------------------------------------------------------------------------------------
// The two launches of kernel 1.  get_prototypes (alpmodule.py:108-158) is one logical pass with one true dependency:
// a window's output row (its rank among the survivors) and the AUTO_FG decision need the whole mask first.
//   stage 1  grid (1 + h*S, nsets): CTA 0 of a set = the mask pass (K1a), the others = per-row partial sums of the
//            global prototype (K1c), which do not depend on it;
//   stage 2  grid (N + S, nsets): CTAs < N = surviving windows (K1b), the others = the global rows (K1d).
// ------------------------------------------------------------------------------------
struct ProtoArgs {
    const float* sup_x;
    int64_t xs_s, xs_c, xs_y, xs_x;
    const float* sup_y;
    int S, C, h, w, kh, kw, akh, akw, cap_rows;
    float thresh;
    float *pooled, *ysum, *partial, *protos;
    uint8_t* survive;
    int32_t *rowidx, *counts, *eff_modes, *status, *plocal;
};

__global__ void __launch_bounds__(256) k_proto_stage1(ProtoArgs a, SetModes modes)
{
    const int set = blockIdx.y;
    TraceRec* tr = (threadIdx.x == 0 && (blockIdx.x & 7) == 0) ? trace_begin(6) : nullptr;
    if (blockIdx.x == 0)
        pool_mask_body(set, a.sup_y, modes, a.S, a.h, a.w, a.kh, a.kw, a.akh, a.akw, a.thresh, a.pooled, a.survive, a.rowidx,
                       a.ysum, a.counts, a.eff_modes, a.status, a.plocal);
    else
        global_partial_body((blockIdx.x - 1) % a.h, (blockIdx.x - 1) / a.h, set, a.sup_x, a.xs_s, a.xs_c, a.xs_y, a.xs_x,
                            a.sup_y, modes, a.S, a.C, a.h, a.w, a.partial);
    trace_end(tr);
}

__global__ void __launch_bounds__(256) k_proto_stage2(ProtoArgs a, SetModes modes, int N)
{
    extern __shared__ __align__(16) float s_stage[];
    __shared__ __align__(8) uint64_t s_bar;
    const int set = blockIdx.y;
    TraceRec* tr = (threadIdx.x == 0 && (blockIdx.x & 15) == 0) ? trace_begin(7) : nullptr;
    if ((int)blockIdx.x < N)
        pool_feat_body(blockIdx.x, set, a.sup_x, a.xs_s, a.xs_c, a.xs_y, a.xs_x, a.rowidx, a.S, a.C, a.h, a.w, a.kh, a.kw,
                       a.cap_rows, a.protos, s_stage, &s_bar);
    else
        global_final_body(blockIdx.x - N, set, a.partial, a.ysum, a.eff_modes, a.plocal, modes, a.S, a.C, a.h, a.cap_rows,
                          a.protos);
    trace_end(tr);
}

// ------------------------------------------------------------------------------------
// Viz grid (`resized_proto_grid`, alpmodule.py:120-128 / 142-150).  One CTA.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_proto_grid(const float* __restrict__ pooled, int S, int gh, int gw, int vw,
                                                    float thresh, int mode, float* __restrict__ out,
                                                    int32_t* __restrict__ ord /* [S*gh*gw] scratch in out tail? */)
{
    // ord[n] = ordinal of entry n among the non-zero entries of the thresholded grid, or -1
    extern __shared__ int32_t s_ord[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = S * gh * gw;
    __shared__ int s_warp[8];
    int running = 0;
    for (int base = 0; base < N; base += 256) {
        int n = base + tid, flag = 0;
        if (n < N) {
            float v = pooled[n];
            flag = (!(v < thresh)) && (v != 0.0f);
        }
        int ex = warp_excl_scan_i(flag, lane);
        if (lane == 31) s_warp[wid] = ex + flag;
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int v = s_warp[k];
            if (k < wid) woff += v;
            tot += v;
        }
        if (n < N) s_ord[n] = flag ? running + woff + ex : -1;
        running += tot;
        __syncthreads();
    }
    const int OH = gh * vw, OW = gw * vw;
    for (int i = tid; i < OH * OW; i += 256) {
        int Y = i / OW, X = i % OW, gy = Y / vw;
        // candidate cells whose 2-wide stripe covers column X
        int best_s = -1, best_gx = -1;
        for (int cand = 0; cand < 2; ++cand) {
            int gx;
            if (cand == 0) {
                gx = X / vw;
                if (X - gx * vw >= 2) continue;
            } else {
                if (vw != 1 || X == 0) continue;
                gx = X - 1;
            }
            for (int s = S - 1; s >= 0; --s)
                if (s_ord[(s * gh + gy) * gw + gx] >= 0) {
                    if (s > best_s || (s == best_s && gx > best_gx)) { best_s = s; best_gx = gx; }
                    break;
                }
        }
        float val = 0.0f;
        if (best_s >= 0) {
            int cell = gy * gw + best_gx;   // value always read from shot 0's plane (:128,150)
            if (mode == PSAM_MODE_GRIDCONV_PLUS) {
                int last = -1;              // renumbering (:146-147): last non-zero entry at this cell wins
                for (int s = S - 1; s >= 0 && last < 0; --s) last = s_ord[s * gh * gw + cell];
                val = (float)(last + 1);
            } else {
                float v0 = pooled[cell];
                val = (s_ord[cell] >= 0) ? v0 : 0.0f;
            }
        }
        out[i] = val;
    }
    (void)ord;
}

// ------------------------------------------------------------------------------------
// F.interpolate(mask, (h,w), mode='nearest') of the support masks (models/grid_proto_fewshot.py:228-231):
// ATen's nearest index = min(floor(dst * scale), in - 1) with scale = (float)in / out.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mask_nearest(const float* __restrict__ src, int n, int H, int W, int h, int w,
                                                      float* __restrict__ dst)
{
    const size_t total = (size_t)n * h * w;
    const float sy = __fdiv_rn((float)H, (float)h), sx = __fdiv_rn((float)W, (float)w);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int x = (int)(i % w), y = (int)((i / w) % h);
        const size_t img = i / ((size_t)w * h);
        const int iy = min((int)floorf(__fmul_rn((float)y, sy)), H - 1);
        const int ix = min((int)floorf(__fmul_rn((float)x, sx)), W - 1);
        dst[i] = src[(img * H + iy) * W + ix];
    }
}

// ------------------------------------------------------------------------------------
// FewShotSeg.forward with several shots (models/grid_proto_fewshot.py:244-270): the background set uses all shots,
// the foreground is matched once per shot and the scores are combined with an element-wise max.  Sets of label l
// are (bg_l, fg_l shot 0, ..., fg_l shot S-1); the result is the [Q*L, 2, HW] logits tensor of the caller.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_combine_shots(const float* __restrict__ scores, int Q, int L, int S, int HW,
                                                       float* __restrict__ logits)
{
    const size_t total = (size_t)Q * L * HW;
    const int nsets = L * (1 + S);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int p = (int)(i % HW), l = (int)((i / HW) % L);
        const size_t q = i / ((size_t)HW * L);
        const float* src = scores + (q * nsets + (size_t)l * (1 + S)) * HW + p;
        float m = src[(size_t)HW];
        for (int s = 1; s < S; ++s) {
            const float v = src[(size_t)(1 + s) * HW];
            m = (v > m || v != v) ? v : m;          // torch.max propagates NaN
        }
        float* dst = logits + ((q * L + l) * 2) * HW + p;
        dst[0] = src[0];
        dst[HW] = m;
    }
}

// ------------------------------------------------------------------------------------
// DINOv2 token hand-off (models/grid_proto_fewshot.py:90-98): x_norm_patchtokens [B, h*w, C] is already the
// channels-last feature map; when it has fewer than 32x32 tokens the reference resizes it bilinearly
// (align_corners=False) to 32x32.  Same arithmetic as ATen's CPU kernel (source index = scale*(dst+0.5)-0.5 clamped at
// 0, horizontal lerps then the vertical one, each fma(a, w0, b*w1)); threads run along C, so every load and store is
// coalesced.  grid = (ow, oh, B).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void bil_src(int in, int out, int o, int& i0, int& i1, float& w0, float& w1)
{
    if (in == out) { i0 = i1 = o; w0 = 1.f; w1 = 0.f; return; }
    const float scale = __fdiv_rn((float)in, (float)out);
    float r = __fmaf_rn(scale, __fadd_rn((float)o, 0.5f), -0.5f);
    if (r < 0.f) r = 0.f;
    i0 = min((int)floorf(r), in - 1);
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    w1 = fminf(fmaxf(__fsub_rn(r, (float)i0), 0.f), 1.f);
    w0 = __fsub_rn(1.f, w1);
}

__global__ void __launch_bounds__(256) k_tokens_bilinear(const float* __restrict__ tok, int h, int w, int C, int oh, int ow,
                                                         float* __restrict__ out)
{
    const int ox = blockIdx.x, oy = blockIdx.y, b = blockIdx.z;
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    bil_src(h, oh, oy, y0, y1, wy0, wy1);
    bil_src(w, ow, ox, x0, x1, wx0, wx1);
    const float* base = tok + (size_t)b * h * w * C;
    const float* p00 = base + ((size_t)y0 * w + x0) * C;
    const float* p01 = base + ((size_t)y0 * w + x1) * C;
    const float* p10 = base + ((size_t)y1 * w + x0) * C;
    const float* p11 = base + ((size_t)y1 * w + x1) * C;
    float* dst = out + (((size_t)b * oh + oy) * ow + ox) * C;
    for (int c = threadIdx.x; c < C; c += 256) {
        const float r0 = __fmaf_rn(p00[c], wx0, __fmul_rn(p01[c], wx1));
        const float r1 = __fmaf_rn(p10[c], wx0, __fmul_rn(p11[c], wx1));
        dst[c] = __fmaf_rn(r0, wy0, __fmul_rn(r1, wy1));
    }
}

}  // namespace psam

using namespace psam;

PSAM_TRACE_TU();
extern "C" int psam_tokens_to_features(const float* tokens, int B, int h, int w, int C, int oh, int ow, float* out,
                                       psam_stream_t stream_)
{
    PSAM_TRACE("psam_tokens_to_features");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(tokens && out, "psam_tokens_to_features: null pointer");
    PSAM_CHECK_ARG(B >= 1 && B <= 65535 && h >= 1 && w >= 1 && C >= 1 && oh >= h && ow >= w && oh <= 65535,
                   "psam_tokens_to_features: bad shape (upsampling only)");
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_tokens_bilinear);
    k_tokens_bilinear<<<dim3(ow, oh, B), 256, 0, stream>>>(tokens, h, w, C, oh, ow, out);
    PSAM_CHECK_LAUNCH("k_tokens_bilinear");
    return PSAM_OK;
}

extern "C" int psam_combine_shots(const float* scores, int Q, int L, int S, int HW, float* logits, psam_stream_t stream_)
{
    PSAM_TRACE("psam_combine_shots");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(scores && logits, "psam_combine_shots: null pointer");
    PSAM_CHECK_ARG(Q >= 1 && L >= 1 && S >= 1 && HW >= 1, "psam_combine_shots: bad shape");
    const size_t total = (size_t)Q * L * HW;
    const int grid = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_combine_shots);
    k_combine_shots<<<grid, 256, 0, stream>>>(scores, Q, L, S, HW, logits);
    PSAM_CHECK_LAUNCH("k_combine_shots");
    return PSAM_OK;
}

extern "C" int psam_mask_nearest(const float* src, int n, int H, int W, int h, int w, float* dst, psam_stream_t stream_)
{
    PSAM_TRACE("psam_mask_nearest");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(src && dst, "psam_mask_nearest: null pointer");
    PSAM_CHECK_ARG(n >= 1 && H >= 1 && W >= 1 && h >= 1 && w >= 1, "psam_mask_nearest: bad shape");
    const size_t total = (size_t)n * h * w;
    const int grid = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_mask_nearest);
    k_mask_nearest<<<grid, 256, 0, stream>>>(src, n, H, W, h, w, dst);
    PSAM_CHECK_LAUNCH("k_mask_nearest");
    return PSAM_OK;
}

extern "C" size_t psam_alp_prototypes_workspace(int nsets, int S, int C, int h, int w, int kh, int kw)
{
    if (nsets <= 0 || S <= 0 || C <= 0 || h <= 0 || w <= 0 || kh <= 0 || kw <= 0) return 0;
    size_t N = (size_t)S * (h / kh) * (w / kw);
    size_t b = 0;
    b += align_up(sizeof(int32_t) * nsets * (N ? N : 1), 256);      // rowidx
    b += align_up(sizeof(float) * nsets * S, 256);                  // ysum
    b += align_up(sizeof(int32_t) * nsets, 256);                    // plocal
    b += align_up(sizeof(float) * (size_t)nsets * S * h * C, 256);  // partial
    return b + 256;
}

static int alp_prototypes_impl(const float* sup_x, const int64_t* xs, const float* sup_y, int nsets,
                               const int32_t* set_modes, const int32_t* set_shots, int S, int C, int h, int w, int kh,
                               int kw, int auto_kh, int auto_kw, float thresh, float* protos, int32_t* counts,
                               int32_t* eff_modes, int32_t* status, uint8_t* survive, float* pooled,
                               void* workspace, size_t workspace_bytes, psam_stream_t stream_);

extern "C" int psam_alp_prototypes(const float* sup_x, const int64_t* xs, const float* sup_y, int nsets,
                                   const int32_t* set_modes, int S, int C, int h, int w, int kh, int kw,
                                   int auto_kh, int auto_kw, float thresh, float* protos, int32_t* counts,
                                   int32_t* eff_modes, int32_t* status, uint8_t* survive, float* pooled,
                                   void* workspace, size_t workspace_bytes, psam_stream_t stream)
{
    PSAM_TRACE("psam_alp_prototypes");
    return alp_prototypes_impl(sup_x, xs, sup_y, nsets, set_modes, nullptr, S, C, h, w, kh, kw, auto_kh, auto_kw, thresh,
                               protos, counts, eff_modes, status, survive, pooled, workspace, workspace_bytes, stream);
}

extern "C" int psam_alp_prototypes_shots(const float* sup_x, const int64_t* xs, const float* sup_y, int nsets,
                                         const int32_t* set_modes, const int32_t* set_shots, int S, int C, int h, int w,
                                         int kh, int kw, int auto_kh, int auto_kw, float thresh, float* protos,
                                         int32_t* counts, int32_t* eff_modes, int32_t* status, uint8_t* survive,
                                         float* pooled, void* workspace, size_t workspace_bytes, psam_stream_t stream)
{
    PSAM_TRACE("psam_alp_prototypes_shots");
    PSAM_CHECK_ARG(set_shots, "psam_alp_prototypes_shots: null set_shots");
    return alp_prototypes_impl(sup_x, xs, sup_y, nsets, set_modes, set_shots, S, C, h, w, kh, kw, auto_kh, auto_kw, thresh,
                               protos, counts, eff_modes, status, survive, pooled, workspace, workspace_bytes, stream);
}

static int alp_prototypes_impl(const float* sup_x, const int64_t* xs, const float* sup_y, int nsets,
                               const int32_t* set_modes, const int32_t* set_shots, int S, int C, int h, int w, int kh,
                               int kw, int auto_kh, int auto_kw, float thresh, float* protos, int32_t* counts,
                               int32_t* eff_modes, int32_t* status, uint8_t* survive, float* pooled,
                               void* workspace, size_t workspace_bytes, psam_stream_t stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(sup_x && xs && sup_y && set_modes && protos && counts && eff_modes && status && survive && pooled,
                   "psam_alp_prototypes: null pointer");
    PSAM_CHECK_ARG(nsets >= 1 && nsets <= kMaxSets, "psam_alp_prototypes: nsets %d not in [1,%d]", nsets, kMaxSets);
    PSAM_CHECK_ARG(S >= 1 && C >= 1 && h >= 1 && w >= 1, "psam_alp_prototypes: bad shape");
    PSAM_CHECK_ARG(C <= 256 * kMaxCPerThread, "psam_alp_prototypes: C=%d exceeds %d", C, 256 * kMaxCPerThread);
    PSAM_CHECK_ARG(kh >= 1 && kw >= 1 && kh <= h && kw <= w, "psam_alp_prototypes: window %dx%d vs map %dx%d", kh, kw, h, w);
    SetModes modes;
    bool any_auto = false;
    for (int i = 0; i < nsets; ++i) {
        PSAM_CHECK_ARG(set_modes[i] >= 0 && set_modes[i] <= 3, "psam_alp_prototypes: invalid mode %d", set_modes[i]);
        modes.m[i] = (int8_t)set_modes[i];
        const int shot = set_shots ? set_shots[i] : -1;
        PSAM_CHECK_ARG(shot >= -1 && shot < S && shot < 127, "psam_alp_prototypes: set %d selects shot %d of %d", i, shot, S);
        modes.shot[i] = (int8_t)shot;
        any_auto |= set_modes[i] == PSAM_MODE_AUTO_FG;
    }
    if (any_auto)
        PSAM_CHECK_ARG(auto_kh >= 1 && auto_kw >= 1 && auto_kh <= h && auto_kw <= w,
                       "psam_alp_prototypes: AUTO_FG needs a valid auto window");
    else
        auto_kh = auto_kw = 1;
    if (workspace_bytes < psam_alp_prototypes_workspace(nsets, S, C, h, w, kh, kw) || !workspace) {
        set_error("psam_alp_prototypes: workspace too small");
        return PSAM_ERR_WORKSPACE;
    }
    const int gh = h / kh, gw = w / kw, N = S * gh * gw, cap_rows = N + S;
    Carver cv(workspace);
    int32_t* rowidx = cv.take<int32_t>((size_t)nsets * (N ? N : 1));
    float* ysum = cv.take<float>((size_t)nsets * S);
    int32_t* plocal = cv.take<int32_t>(nsets);
    float* partial = cv.take<float>((size_t)nsets * S * h * C);

    ProtoArgs a;
    a.sup_x = sup_x; a.xs_s = xs[0]; a.xs_c = xs[1]; a.xs_y = xs[2]; a.xs_x = xs[3]; a.sup_y = sup_y;
    a.S = S; a.C = C; a.h = h; a.w = w; a.kh = kh; a.kw = kw; a.akh = auto_kh; a.akw = auto_kw; a.cap_rows = cap_rows;
    a.thresh = thresh; a.pooled = pooled; a.ysum = ysum; a.partial = partial; a.protos = protos; a.survive = survive;
    a.rowidx = rowidx; a.counts = counts; a.eff_modes = eff_modes; a.status = status; a.plocal = plocal;
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_proto_stage1);
    k_proto_stage1<<<dim3(1 + h * S, nsets), 256, 0, stream>>>(a, modes);
    PSAM_CHECK_LAUNCH("k_proto_stage1");
    // feature tiles staged through shared memory by the TMA engine when they fit (see pool_feat_body)
    const size_t tile = (size_t)kh * kw * C * sizeof(float);
    const size_t smem = tile <= (size_t)kStageBytes ? tile : 0;
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_proto_stage2);
    k_proto_stage2<<<dim3(N + S, nsets), 256, smem, stream>>>(a, modes, N);
    PSAM_CHECK_LAUNCH("k_proto_stage2");
    return PSAM_OK;
}

extern "C" int psam_alp_proto_grid(const float* pooled, int S, int gh, int gw, int vw, float thresh, int mode,
                                   float* out, psam_stream_t stream_)
{
    PSAM_TRACE("psam_alp_proto_grid");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(pooled && out, "psam_alp_proto_grid: null pointer");
    PSAM_CHECK_ARG(S >= 1 && gh >= 0 && gw >= 0 && vw >= 1, "psam_alp_proto_grid: bad shape");
    PSAM_CHECK_ARG(mode == PSAM_MODE_GRIDCONV || mode == PSAM_MODE_GRIDCONV_PLUS, "psam_alp_proto_grid: grid modes only");
    size_t N = (size_t)S * gh * gw;
    if (N == 0 || gh * vw * gw * vw == 0) return PSAM_OK;
    PSAM_CHECK_ARG(N * sizeof(int32_t) <= 200 * 1024, "psam_alp_proto_grid: grid too large");
    size_t smem = N * sizeof(int32_t);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(k_proto_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_proto_grid);
    k_proto_grid<<<1, 256, smem, stream>>>(pooled, S, gh, gw, vw, thresh, mode, out, nullptr);
    PSAM_CHECK_LAUNCH("k_proto_grid");
    return PSAM_OK;
}

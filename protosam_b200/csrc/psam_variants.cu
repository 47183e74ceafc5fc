// Optional prompt variants of ProtoSAM.forward on the device (off by default in the
// reference, config_ssl_upload.py:93,102):
//   negative points   get_sam_input_points(..., get_neg_points=True)      models/ProtoSAM.py:361-434
//   mask prompts      get_sam_input_mask + predict_w_masks                models/ProtoSAM.py:452-498
//   coarse_pred_only  get_confidence_from_logits                          util/utils.py:429-434
// They work on what the every-pixel variant of kernel 3a and kernel 3b leave on the device: the background
// probability map, the cv2-numbered label image and the prompt records.  Integer / selection work, bit-exact with the
// reference's CPU results (torch.topk tie rules included); none of it is on the default hot path.
#include "psam_common.cuh"
#include "psam_topk.cuh"

namespace psam {

constexpr int NT = 1024;          // threads per CTA = one image row (out <= 1024)
constexpr int MAX_RING = 15;      // dilation iterations supported (the reference uses 10)
constexpr int WIN = 2 * MAX_RING + 2;

struct NegParams {
    const int32_t* labels;        // [n_img,out,out] cv2 labels (0/1 image with use_cca)
    const float* p_bg;            // background probability of image i at p_bg + i * p_bg_stride
    int64_t p_bg_stride;
    const psam_image_hdr* hdr;
    const psam_prompt_rec* recs;
    int n_img, out, max_cc, ring, use_cca, host_aliasing;
    float thresh;
    psam_neg_point* neg;          // [n_img, max_cc + 1]; entry max_cc = the image's global point
};

// torch.topk(values[mask], 1) over a raster scan done by the whole CTA: every thread feeds its pixels through add();
// finish() returns the raster index torch reports (-1: empty mask).  n >= 64 -> first maximum (partial_sort); n < 64 ->
// the first 64 masked pixels are kept in raster order and libstdc++'s nth_element is replayed on them.
struct TopK1Scan {
    unsigned long long best = 0;   // (value bits << 32 | ~index): largest value, then first index
    TK* q;
    int* cnt;                      // [32] per-warp counts of the current row
    int* total;                    // masked pixels seen so far (saturates the collection at 64)

    __device__ void add_row(bool in, float v, int idx, int lane, int wid)
    {
        if (in) best = max(best, ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx));
        const int seen = *total;                                    // uniform: written behind a barrier
        if (seen >= 64) return;
        const uint32_t rb = __ballot_sync(0xffffffffu, in);
        if (lane == 0) cnt[wid] = __popc(rb);
        __syncthreads();
        int base = seen, row_total = 0;
        for (int w2 = 0; w2 < NT / 32; ++w2) {
            const int c = cnt[w2];
            if (w2 < wid) base += c;
            row_total += c;
        }
        if (in) {
            const int pos = base + __popc(rb & ((1u << lane) - 1u));
            if (pos < 64) { q[pos].v = v; q[pos].i = idx; }
        }
        __syncthreads();
        if (threadIdx.x == 0) *total = seen + row_total;
        __syncthreads();
    }

    // every thread gets the result; red = 32 x u64 of shared scratch
    __device__ int finish(unsigned long long* red, int* result, float* value)
    {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        unsigned long long b = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned int lo = __shfl_xor_sync(0xffffffffu, (unsigned int)b, o), hi = __shfl_xor_sync(0xffffffffu, (unsigned int)(b >> 32), o);
            b = max(b, ((unsigned long long)hi << 32) | lo);
        }
        if (lane == 0) red[wid] = b;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long m = 0;
            for (int k = 0; k < NT / 32; ++k) m = max(m, red[k]);
            const int n = *total;
            if (n == 0) { *result = -1; *value = 0.f; }
            else if (n >= 64) { *result = (int)(0xFFFFFFFFu - (unsigned)(m & 0xFFFFFFFFu)); *value = __uint_as_float((unsigned)(m >> 32)); }
            else {
                float vals[64];
                for (int k = 0; k < n; ++k) vals[k] = q[k].v;
                const int idx = topk1_small(q, n);               // permutes q; the value is looked up in the copy
                float v = 0.f;
                for (int k = 0; k < n; ++k) if (q[k].i == idx) v = q[k].v;
                (void)vals;
                *result = idx; *value = v;
            }
        }
        __syncthreads();
        return *result;
    }
};

// grid = (max_cc + 1, n_img), block = 1024.  CTA (slot, img): slot < n_rec -> the ring of component `slot`;
// slot == max_cc -> the global point.  Thread = column.
__global__ void __launch_bounds__(NT) k_neg_points(NegParams P)
{
    __shared__ uint32_t s_C[WIN][32], s_H[WIN][32];
    __shared__ unsigned long long s_red[32];
    __shared__ int s_cnt[32], s_total, s_result;
    __shared__ float s_value;
    __shared__ TK s_q[64];
    const int img = blockIdx.y, slot = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int out = P.out, R = P.ring;
    const psam_image_hdr H = P.hdr[img];
    const bool glob = slot == P.max_cc;
    psam_neg_point* dst = P.neg + (size_t)img * (P.max_cc + 1) + slot;
    if (!glob && slot >= H.n_rec) return;
    const float* pbg = P.p_bg + (size_t)img * P.p_bg_stride;
    if (tid == 0) s_total = 0;
    __syncthreads();
    TopK1Scan scan;
    scan.q = s_q; scan.cnt = s_cnt; scan.total = &s_total;
    const bool col = tid < out;

    if (glob) {
        // bg_p[bg_p < 0.95] = 0; get_most_conf_points(bg_p, bg_p > 0, 1)          models/ProtoSAM.py:363-366
        for (int y = 0; y < out; ++y) {
            float v = col ? pbg[(size_t)y * out + tid] : 0.f;
            if (v < P.thresh) v = 0.f;
            scan.add_row(col && v > 0.f, v, y * out + tid, lane, wid);
        }
    } else {
        const psam_prompt_rec rec = P.recs[(size_t)img * P.max_cc + slot];
        const int label = P.use_cca ? 1 : rec.label;
        const int miny = (int)rec.box[1], maxy = (int)rec.box[3];
        const int ylo = max(0, miny - R), yhi = min(out - 1, maxy + R);
        const int32_t* lab = P.labels + (size_t)img * out * out;
        int next_row = miny;                                         // next component row to turn into bit rows
        for (int y = ylo; y <= yhi; ++y) {
            // bit rows of the component (C) and their horizontal dilation by R (H) up to row y + R
            const int need_to = min(maxy, y + R);
            while (next_row <= need_to) {
                const int r = next_row, sl = r % WIN;
                const uint32_t wbits = __ballot_sync(0xffffffffu, col && lab[(size_t)r * out + tid] == label);
                if (lane == 0) s_C[sl][wid] = wbits;
                __syncthreads();
                if (tid < 32) {
                    const uint32_t lo = tid > 0 ? s_C[sl][tid - 1] : 0u, mid = s_C[sl][tid], hi = tid < 31 ? s_C[sl][tid + 1] : 0u;
                    uint32_t acc = mid;
                    for (int sft = 1; sft <= R; ++sft) acc |= (mid << sft) | (lo >> (32 - sft)) | (mid >> sft) | (hi << (32 - sft));
                    s_H[sl][tid] = acc;
                }
                __syncthreads();
                ++next_row;
            }
            uint32_t dw = 0u;                                        // cv2.dilate(3x3, iterations=R): rows y-R .. y+R
            for (int r = max(miny, y - R); r <= min(maxy, y + R); ++r) dw |= s_H[r % WIN][wid];
            const bool comp = y >= miny && y <= maxy && ((s_C[y % WIN][wid] >> lane) & 1u);
            const bool ring = col && ((dw >> lane) & 1u) && !comp;   // dilated_mask - pred_uint8      :401-405
            float v = 0.f;
            if (ring) {
                v = pbg[(size_t)y * out + tid];
                if (P.host_aliasing && v < P.thresh) v = 0.f;        // the CPU tensor was thresholded in place (:364)
            }
            scan.add_row(ring, v, y * out + tid, lane, wid);
        }
    }
    const int idx = scan.finish(s_red, &s_result, &s_value);
    if (tid == 0) {
        psam_neg_point o;
        o.pt[0] = idx >= 0 ? idx % out : -1;
        o.pt[1] = idx >= 0 ? idx / out : -1;
        o.p = s_value;
        o.has = idx >= 0;
        o.n = s_total;
        o.reserved = 0;
        *dst = o;
    }
}

// get_sam_input_mask + predict_w_masks' input (models/ProtoSAM.py:452-476): per component, the 0/1 mask resized to
// size x size with cv2.INTER_NEAREST (source index = min(floor(dst * (1 / (size / out))), out - 1)), 1 -> 10, 0 -> -8,
// cast to uint8 (-8 -> 248).  Masks are packed in image order: component `slot` of image i lands at
// offsets[i] + slot, offsets = exclusive prefix of n_rec.  grid = (max_cc, n_img), block = 256.
__global__ void __launch_bounds__(256) k_mask_prompts(const int32_t* __restrict__ labels, const psam_image_hdr* __restrict__ hdr,
                                                      const psam_prompt_rec* __restrict__ recs, int n_img, int out, int max_cc,
                                                      int use_cca, int size, int cap, uint8_t* __restrict__ masks,
                                                      int32_t* __restrict__ offsets)
{
    __shared__ int s_part[8];
    const int img = blockIdx.y, slot = blockIdx.x, tid = threadIdx.x;
    int part = 0;
    for (int i = tid; i < img; i += 256) part += hdr[i].n_rec;
    part = warp_sum_i(part);
    if ((tid & 31) == 0) s_part[tid >> 5] = part;
    __syncthreads();
    int base = 0;
    for (int k = 0; k < 8; ++k) base += s_part[k];
    const int n_rec = hdr[img].n_rec;
    if (slot == 0 && tid == 0) {
        offsets[img] = base;
        if (img == n_img - 1) offsets[n_img] = base + n_rec;
    }
    if (slot >= n_rec || base + slot >= cap) return;
    const int label = use_cca ? 1 : recs[(size_t)img * max_cc + slot].label;
    const int32_t* lab = labels + (size_t)img * out * out;
    uint8_t* dst = masks + (size_t)(base + slot) * size * size;
    const double inv = 1.0 / ((double)size / (double)out);
    for (int i = tid; i < size * size; i += 256) {
        const int yy = i / size, xx = i - yy * size;
        const int sy = min((int)floor((double)yy * inv), out - 1), sx = min((int)floor((double)xx * inv), out - 1);
        dst[i] = lab[(size_t)sy * out + sx] == label ? (uint8_t)10 : (uint8_t)248;
    }
}

// get_confidence_from_logits (util/utils.py:429-434) from the foreground probabilities: sum of p over p >= 0.5 divided
// by (count + 1e-6).  p >= 0.5 is a multiple of 2^-24, so the sum is accumulated exactly in integers (the reference sums
// in float32: agreement ~1e-7 relative).  grid = n_img, block = 1024.
__global__ void __launch_bounds__(NT) k_confidence(const float* __restrict__ p_fg, size_t npx, double* __restrict__ conf)
{
    __shared__ unsigned long long s_sum[32];
    __shared__ unsigned int s_cnt[32];
    const float* p = p_fg + (size_t)blockIdx.x * npx;
    unsigned long long sum = 0;
    unsigned int cnt = 0;
    for (size_t i = threadIdx.x; i < npx; i += NT) {
        const float v = p[i];
        if (v >= 0.5f) { sum += (unsigned long long)(v * 16777216.0f); ++cnt; }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned int lo = __shfl_xor_sync(0xffffffffu, (unsigned int)sum, o), hi = __shfl_xor_sync(0xffffffffu, (unsigned int)(sum >> 32), o);
        sum += ((unsigned long long)hi << 32) | lo;
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) { s_sum[wid] = sum; s_cnt[wid] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        unsigned int c = 0;
        for (int k = 0; k < NT / 32; ++k) { t += s_sum[k]; c += s_cnt[k]; }
        conf[blockIdx.x] = ((double)t * (1.0 / 16777216.0)) / ((double)c + 1e-6);
    }
}

}  // namespace psam

using namespace psam;

// get_most_conf_points for any k: one thread per component replays torch.topk over the component's pixels in raster order
// (they are read through the component's box).  Selection work for a rarely used variant: correctness, not speed.
__global__ void __launch_bounds__(32) k_topk_points(const int32_t* __restrict__ labels, const float* __restrict__ p_fg,
                                                    int64_t p_stride, const psam_image_hdr* __restrict__ hdr,
                                                    const psam_prompt_rec* __restrict__ recs, int n_img, int out, int max_cc,
                                                    int use_cca, int k, int64_t* __restrict__ pts, float* __restrict__ conf,
                                                    TK* __restrict__ scratch)
{
    const int slot = blockIdx.x * 32 + threadIdx.x;
    if (slot >= n_img * max_cc) return;
    const int img = slot / max_cc, r = slot - img * max_cc;
    int64_t* o_pts = pts + (size_t)slot * k * 2;
    float* o_conf = conf + (size_t)slot * k;
    for (int j = 0; j < k; ++j) { o_pts[2 * j] = -1; o_pts[2 * j + 1] = -1; o_conf[j] = 0.f; }
    if (r >= hdr[img].n_rec) return;
    const psam_prompt_rec rec = recs[slot];
    const int n = rec.area;
    if (n < k) return;                                   // torch.topk raises: the host wrapper does
    const int want = use_cca ? 1 : rec.label;
    const int32_t* lab = labels + (size_t)img * out * out;
    const float* pf = p_fg + (size_t)img * p_stride;
    const int x0 = (int)rec.box[0], y0 = (int)rec.box[1], x1 = (int)rec.box[2], y1 = (int)rec.box[3];
    TK heap[TOPK_MAX_K];
    TK* q;
    if ((long long)k * 64 <= n) {
        // partial_sort: heap of the first k pixels, every later pixel replaces the top only if strictly larger
        q = heap;
        int c = 0;
        for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) {
                const int idx = y * out + x;
                if (lab[idx] != want) continue;
                TK e{pf[idx], idx};
                if (c < k) {
                    q[c] = e;
                    if (++c == k) tk_make_heap(q, k);
                } else if (tk_gt(e, q[0])) {
                    tk_adjust_heap(q, 0, k, e);          // __pop_heap: the old top leaves the range
                }
            }
        tk_sort_heap(q, k);
    } else {
        // n < 64 k: nth_element + sort over all pixels of the component, staged in this slot's scratch
        q = scratch + (size_t)slot * 64 * k;
        int c = 0;
        for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) {
                const int idx = y * out + x;
                if (lab[idx] == want) q[c++] = TK{pf[idx], idx};
            }
        tk_introselect(q, 0, k - 1, n);
        tk_sort(q, k - 1);
    }
    for (int j = 0; j < k; ++j) {
        o_pts[2 * j] = q[j].i % out;
        o_pts[2 * j + 1] = q[j].i / out;
        o_conf[j] = q[j].v;
    }
}

extern "C" size_t psam_topk_points_workspace(int n_img, int max_cc, int k)
{
    if (n_img < 1 || max_cc < 1 || k < 1) return 256;
    return align_up((size_t)n_img * max_cc * 64 * k * sizeof(TK), 256);
}

extern "C" int psam_topk_points(const int32_t* labels, const float* p_fg, int64_t p_fg_image_stride, const psam_image_hdr* hdr,
                                const psam_prompt_rec* recs, int n_img, int out, int max_cc, int use_cca, int k, int64_t* pts,
                                float* conf, void* workspace, size_t workspace_bytes, psam_stream_t stream_)
{
    PSAM_TRACE("psam_topk_points");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(labels && p_fg && hdr && recs && pts && conf, "psam_topk_points: null pointer");
    PSAM_CHECK_ARG(n_img >= 1 && out >= 1 && max_cc >= 1 && (long long)n_img * max_cc < (1ll << 30), "psam_topk_points: bad shape");
    PSAM_CHECK_ARG(k >= 1 && k <= TOPK_MAX_K, "psam_topk_points: k = %d not in [1, %d]", k, TOPK_MAX_K);
    if (!workspace || workspace_bytes < psam_topk_points_workspace(n_img, max_cc, k)) {
        set_error("psam_topk_points: workspace too small (%zu < %zu)", workspace_bytes, psam_topk_points_workspace(n_img, max_cc, k));
        return PSAM_ERR_WORKSPACE;
    }
    PSAM_PROF_BEGIN(stream);
    k_topk_points<<<(n_img * max_cc + 31) / 32, 32, 0, stream>>>(labels, p_fg, p_fg_image_stride, hdr, recs, n_img, out, max_cc,
                                                               use_cca ? 1 : 0, k, pts, conf, static_cast<TK*>(workspace));
    PSAM_CHECK_LAUNCH("k_topk_points");
    return PSAM_OK;
}

extern "C" int psam_neg_points(const int32_t* labels, const float* p_bg, int64_t p_bg_image_stride, const psam_image_hdr* hdr,
                               const psam_prompt_rec* recs, int n_img, int out, int max_cc, int use_cca, int ring_width,
                               float thresh, int host_aliasing, psam_neg_point* neg, psam_stream_t stream_)
{
    PSAM_TRACE("psam_neg_points");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(labels && p_bg && hdr && recs && neg, "psam_neg_points: null pointer");
    PSAM_CHECK_ARG(n_img >= 1 && n_img <= 65535 && out >= 1 && out <= NT && max_cc >= 1, "psam_neg_points: bad shape (out <= %d)", NT);
    PSAM_CHECK_ARG(ring_width >= 1 && ring_width <= MAX_RING, "psam_neg_points: ring width %d not in [1,%d]", ring_width, MAX_RING);
    NegParams P;
    P.labels = labels; P.p_bg = p_bg; P.p_bg_stride = p_bg_image_stride; P.hdr = hdr; P.recs = recs; P.n_img = n_img; P.out = out;
    P.max_cc = max_cc; P.ring = ring_width; P.use_cca = use_cca ? 1 : 0; P.host_aliasing = host_aliasing ? 1 : 0; P.thresh = thresh;
    P.neg = neg;
    cudaError_t e = cudaMemsetAsync(neg, 0, sizeof(psam_neg_point) * (size_t)n_img * (max_cc + 1), stream);
    if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return PSAM_ERR_LAUNCH; }
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_neg_points);
    k_neg_points<<<dim3(max_cc + 1, n_img), NT, 0, stream>>>(P);
    PSAM_CHECK_LAUNCH("k_neg_points");
    return PSAM_OK;
}

extern "C" int psam_mask_prompts(const int32_t* labels, const psam_image_hdr* hdr, const psam_prompt_rec* recs, int n_img, int out,
                                 int max_cc, int use_cca, int size, int capacity, uint8_t* masks, int32_t* offsets,
                                 psam_stream_t stream_)
{
    PSAM_TRACE("psam_mask_prompts");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(labels && hdr && recs && masks && offsets, "psam_mask_prompts: null pointer");
    PSAM_CHECK_ARG(n_img >= 1 && n_img <= 65535 && out >= 1 && max_cc >= 1 && size >= 1 && capacity >= 1, "psam_mask_prompts: bad shape");
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_mask_prompts);
    k_mask_prompts<<<dim3(max_cc, n_img), 256, 0, stream>>>(labels, hdr, recs, n_img, out, max_cc, use_cca ? 1 : 0, size, capacity, masks,
                                                           offsets);
    PSAM_CHECK_LAUNCH("k_mask_prompts");
    return PSAM_OK;
}

extern "C" int psam_confidence(const float* p_fg, int n_img, int64_t pixels_per_image, double* conf, psam_stream_t stream_)
{
    PSAM_TRACE("psam_confidence");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(p_fg && conf, "psam_confidence: null pointer");
    PSAM_CHECK_ARG(n_img >= 1 && pixels_per_image >= 1, "psam_confidence: bad shape");
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_confidence);
    k_confidence<<<n_img, NT, 0, stream>>>(p_fg, (size_t)pixels_per_image, conf);
    PSAM_CHECK_LAUNCH("k_confidence");
    return PSAM_OK;
}

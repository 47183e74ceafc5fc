/*
 * psam_oracle.c -- CPU restatement of the ProtoSAM coarse-segmentation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path
 * in protosam_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * Parity status: PINNED BY EXECUTION.  The reference (levayz/ProtoSAM) holds no
 * tests or golden vectors for this path, so every function
 * here is pinned against the reference's own Python code run in the build
 * container (oracle/make_golden.py -> tests/golden/ (npz files), produced with
 * torch 2.11.0+cu128 CPU / numpy 2.3.5 / opencv 4.13.0).
 *
 * Third-party arithmetic the reference path relies on and that is restated
 * here (none of it lives under /root/reference):
 *   - ATen CPU avg_pool2d            (call sites models/alpmodule.py:114,118,136,140)
 *   - ATen CPU upsample_bilinear2d   (models/grid_proto_fewshot.py:272-273, models/ProtoSAM.py:594)
 *   - ATen CPU softmax, dim=1        (models/ProtoSAM.py:599) incl. SLEEF expf_u10
 *   - numpy float32 pairwise sum     (util/utils.py:490)
 *   - OpenCV connectedComponentsWithStats, 8-connectivity (util/utils.py:478)
 * Each restatement below was checked bit-for-bit against the library it
 * restates (tests/test_oracle_vs_libs.py).
 *
 * Compile with -ffp-contract=off: every fused multiply-add is written as an
 * explicit fmaf() so the operation sequence is the same on any host and can be
 * replayed with __fmaf_rn / __fmul_rn / __fadd_rn on the GPU.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PSAMO_MODE_MASK 0
#define PSAMO_MODE_GRIDCONV 1
#define PSAMO_MODE_GRIDCONV_PLUS 2

/* ------------------------------------------------------------------------- */
/* A. prototypes: models/alpmodule.py:97-159 (get_prototypes)                 */
/* ------------------------------------------------------------------------- */

/* F.avg_pool2d(y, (kh,kw)) on one [h,w] plane: kernel = stride, no padding,
 * floor.  ATen CPU sums the window sequentially in (dy,dx) order in fp32 and
 * divides by kh*kw (true division).  models/alpmodule.py:118,140. */
static void pool_plane(const float* y, int64_t sy, int64_t sx, int h, int w, int kh, int kw, float* out)
{
    int gh = h / kh, gw = w / kw;
    float div = (float)(kh * kw);
    for (int gy = 0; gy < gh; ++gy)
        for (int gx = 0; gx < gw; ++gx) {
            float acc = 0.0f;
            for (int dy = 0; dy < kh; ++dy)
                for (int dx = 0; dx < kw; ++dx)
                    acc = acc + y[(int64_t)(gy * kh + dy) * sy + (int64_t)(gx * kw + dx) * sx];
            out[gy * gw + gx] = acc / div;
        }
}

/* pooled foreground fraction for all shots: [S,h,w] contiguous -> [S,gh,gw] */
void psamo_pool_mask(const float* y, int S, int h, int w, int kh, int kw, float* pooled)
{
    int gh = h / kh, gw = w / kw;
    for (int s = 0; s < S; ++s)
        pool_plane(y + (int64_t)s * h * w, w, 1, h, w, kh, kw, pooled + (int64_t)s * gh * gw);
}

/* safe_norm over one row: x / max(||x||_2, 1e-4).  models/alpmodule.py:14-18 */
static void safe_norm_row(float* v, int C)
{
    float ss = 0.0f;
    for (int c = 0; c < C; ++c) ss += v[c] * v[c];
    float n = sqrtf(ss);
    if (n < 1e-4f) n = 1e-4f;
    for (int c = 0; c < C; ++c) v[c] = v[c] / n;
}

/* get_prototypes.  x: [S,C,h,w] addressed through element strides xs[4]
 * (shot, channel, row, col) so the physically channels-last tensor the
 * reference receives is read in place; y: [S,h,w]
 * contiguous.  Outputs:
 *   pooled  [S*gh*gw]   window foreground fraction (sup_y_g)
 *   survive [S*gh*gw]   pooled > thresh              (alpmodule.py:131,153)
 *   protos  [(P + (mode==2 ? S : 0)) * C] rows in (shot, gy, gx) order, then the
 *           S global prototypes for gridconv+; L2-normalised with safe_norm.
 *           mode 'mask': [S*C] global masked averages, NOT normalised (:99-103).
 * Returns the number of rows written to protos. */
int psamo_prototypes(const float* x, const int64_t* xs, const float* y,
                     int S, int C, int h, int w, int kh, int kw, float thresh, int mode,
                     float* protos, uint8_t* survive, float* pooled)
{
    int gh = h / kh, gw = w / kw, rows = 0;
    if (mode != PSAMO_MODE_MASK) {
        psamo_pool_mask(y, S, h, w, kh, kw, pooled);
        float div = (float)(kh * kw);
        for (int s = 0; s < S; ++s)
            for (int gy = 0; gy < gh; ++gy)
                for (int gx = 0; gx < gw; ++gx) {
                    int n = (s * gh + gy) * gw + gx;
                    survive[n] = pooled[n] > thresh;
                    if (!survive[n]) continue;
                    float* dst = protos + (int64_t)rows * C;
                    for (int c = 0; c < C; ++c) {
                        float acc = 0.0f;
                        const float* base = x + s * xs[0] + c * xs[1];
                        for (int dy = 0; dy < kh; ++dy)
                            for (int dx = 0; dx < kw; ++dx)
                                acc = acc + base[(gy * kh + dy) * xs[2] + (gx * kw + dx) * xs[3]];
                        dst[c] = acc / div;
                    }
                    ++rows;
                }
    }
    if (mode == PSAMO_MODE_MASK || mode == PSAMO_MODE_GRIDCONV_PLUS) {
        /* sum(x*y, hw) / (sum(y, hw) + 1e-5)   alpmodule.py:99-100,155-156 */
        for (int s = 0; s < S; ++s) {
            float ysum = 0.0f;
            for (int i = 0; i < h * w; ++i) ysum += y[(int64_t)s * h * w + i];
            float* dst = protos + (int64_t)rows * C;
            for (int c = 0; c < C; ++c) {
                float acc = 0.0f;
                const float* base = x + s * xs[0] + c * xs[1];
                for (int yy = 0; yy < h; ++yy)
                    for (int xx = 0; xx < w; ++xx)
                        acc += base[yy * xs[2] + xx * xs[3]] * y[((int64_t)s * h + yy) * w + xx];
                dst[c] = acc / (ysum + 1e-5f);
            }
            ++rows;
        }
    }
    if (mode != PSAMO_MODE_MASK)
        for (int r = 0; r < rows; ++r) safe_norm_row(protos + (int64_t)r * C, C);
    return rows;
}

/* ------------------------------------------------------------------------- */
/* B. match: models/alpmodule.py:57-94 (get_prediction_from_prototypes)       */
/* ------------------------------------------------------------------------- */

/* q: [C,h,w] addressed through element strides qs[3] (channel,row,col).
 * grid modes (:67-91): qn = safe_norm(q) (:195); d_p = 20 * <qn, pro_p>;
 *   pred = sum_p softmax_p(d) * d; assign = argmax_p d (first maximum).
 * mask mode (:58-65): d_s = 20 * cos(q, pro_s) with both norms clamped at
 *   eps=1e-4; pred = max_s d_s; assign = pred.
 * sims (optional): [P,h*w] raw d for grid modes; [h*w] = pred in mask mode.
 * Returns 0, or -1 when P == 0 in a grid mode (the reference raises inside
 * F.conv2d with a [0,C,1,1] weight). */
int psamo_match(const float* q, const int64_t* qs, int C, int h, int w,
                const float* protos, int P, int mode,
                float* pred, float* assign, float* sims)
{
    int HW = h * w;
    if (P <= 0) return -1;
    float* qn = (float*)malloc(sizeof(float) * C);
    float* pn = (float*)malloc(sizeof(float) * (size_t)P * C);
    float* d = (float*)malloc(sizeof(float) * P);
    memcpy(pn, protos, sizeof(float) * (size_t)P * C);
    if (mode == PSAMO_MODE_MASK)
        for (int p = 0; p < P; ++p) safe_norm_row(pn + (int64_t)p * C, C);
    for (int yy = 0; yy < h; ++yy)
        for (int xx = 0; xx < w; ++xx) {
            int i = yy * w + xx;
            for (int c = 0; c < C; ++c) qn[c] = q[c * qs[0] + yy * qs[1] + xx * qs[2]];
            safe_norm_row(qn, C);
            for (int p = 0; p < P; ++p) {
                const float* pr = pn + (int64_t)p * C;
                float acc = 0.0f;
                for (int c = 0; c < C; ++c) acc += qn[c] * pr[c];
                d[p] = acc * 20.0f;
            }
            int am = 0;
            float mx = d[0];
            for (int p = 1; p < P; ++p)
                if (d[p] > mx) { mx = d[p]; am = p; }
            if (mode == PSAMO_MODE_MASK) {
                pred[i] = mx;
                assign[i] = mx;
                if (sims) sims[i] = mx;
            } else {
                float se = 0.0f, sed = 0.0f;
                for (int p = 0; p < P; ++p) {
                    float e = expf(d[p] - mx);
                    se += e;
                    sed += e * d[p];
                }
                pred[i] = sed / se;
                assign[i] = (float)am;
                if (sims)
                    for (int p = 0; p < P; ++p) sims[(int64_t)p * HW + i] = d[p];
            }
        }
    free(qn); free(pn); free(d);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* C. coarse map -> probabilities: ProtoSAM.py:594-600                        */
/* ------------------------------------------------------------------------- */

/* ATen CPU upsample_bilinear2d (align_corners=False, no scale factors) on one
 * plane, restated from its observed arithmetic in torch 2.11 (AVX2/AVX512
 * builds, gcc -ffp-contract=fast):
 *   scale = (float)in / (float)out
 *   src   = max(fmaf(scale, dst + 0.5f, -0.5f), 0)
 *   i0 = min((int)src, in-1); i1 = i0 + (i0 < in-1); l1 = src - i0; l0 = 1 - l1
 *   row_k = fmaf(v[k][x0], wx0, v[k][x1] * wx1)
 *   out   = fmaf(row_0, wy0, row_1 * wy1)
 * in == out along an axis copies (index identity, weights 1/0). */
static void bil_axis(int in, int out, int* i0, int* i1, float* l0, float* l1)
{
    float scale = (float)in / (float)out;
    for (int o = 0; o < out; ++o) {
        if (in == out) { i0[o] = o; i1[o] = o; l0[o] = 1.0f; l1[o] = 0.0f; continue; }
        float r = fmaf(scale, (float)o + 0.5f, -0.5f);
        if (r < 0.0f) r = 0.0f;
        int a = (int)floorf(r);
        if (a > in - 1) a = in - 1;
        float lam = r - (float)a;
        if (lam < 0.0f) lam = 0.0f;
        if (lam > 1.0f) lam = 1.0f;
        i0[o] = a;
        i1[o] = a + (a < in - 1 ? 1 : 0);
        l1[o] = lam;
        l0[o] = 1.0f - lam;
    }
}

void psamo_bilinear(const float* in, int ih, int iw, float* out, int oh, int ow)
{
    int* y0 = (int*)malloc(sizeof(int) * oh); int* y1 = (int*)malloc(sizeof(int) * oh);
    int* x0 = (int*)malloc(sizeof(int) * ow); int* x1 = (int*)malloc(sizeof(int) * ow);
    float* wy0 = (float*)malloc(sizeof(float) * oh); float* wy1 = (float*)malloc(sizeof(float) * oh);
    float* wx0 = (float*)malloc(sizeof(float) * ow); float* wx1 = (float*)malloc(sizeof(float) * ow);
    bil_axis(ih, oh, y0, y1, wy0, wy1);
    bil_axis(iw, ow, x0, x1, wx0, wx1);
    for (int y = 0; y < oh; ++y) {
        const float* ra = in + (int64_t)y0[y] * iw;
        const float* rb = in + (int64_t)y1[y] * iw;
        for (int x = 0; x < ow; ++x) {
            float r0 = fmaf(ra[x0[x]], wx0[x], ra[x1[x]] * wx1[x]);
            float r1 = fmaf(rb[x0[x]], wx0[x], rb[x1[x]] * wx1[x]);
            out[(int64_t)y * ow + x] = fmaf(r0, wy0[y], r1 * wy1[y]);
        }
    }
    free(y0); free(y1); free(x0); free(x1); free(wy0); free(wy1); free(wx0); free(wx1);
}

/* SLEEF 3.x Sleef_expf{8,16}_u10 (FMA build) -- the exp ATen's vectorised CPU
 * softmax calls through Vectorized<float>::exp(). */
static inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float pow2i(int q) { return as_float((uint32_t)(q + 0x7f) << 23); }

float psamo_expf_u10(float d)
{
    float qf = nearbyintf(d * 1.442695040888963407359924681001892137426645954152985934135449406931f);
    int q = (int)qf;
    float s = fmaf(qf, -0.693145751953125f, d);
    s = fmaf(qf, -1.428606765330187045e-06f, s);
    float u = 0.000198527617612853646278381f;
    u = fmaf(u, s, 0.00139304355252534151077271f);
    u = fmaf(u, s, 0.00833336077630519866943359f);
    u = fmaf(u, s, 0.0416664853692054748535156f);
    u = fmaf(u, s, 0.166666671633720397949219f);
    u = fmaf(u, s, 0.5f);
    u = 1.0f + fmaf(s * s, u, s);
    u = u * pow2i(q >> 1) * pow2i(q - (q >> 1));
    if (d < -104.0f) u = 0.0f;
    if (d > 100.0f) u = INFINITY;
    return u;
}

/* softmax over the 2-channel dim of [1,2,H,W] (ProtoSAM.py:599), ATen's
 * vectorised inner-dim kernel: m = max; e_k = exp(l_k - m); s = (0 + e_0) + e_1;
 * p_k = e_k / s.  Valid where ATen takes the vector path (H*W a multiple of
 * the SIMD width; 1024*1024 is). */
void psamo_softmax2(const float* l0, const float* l1, float* p0, float* p1, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        float m = l0[i] > l1[i] ? l0[i] : l1[i];
        float e0 = psamo_expf_u10(l0[i] - m), e1 = psamo_expf_u10(l1[i] - m);
        float s = 0.0f + e0;
        s = s + e1;
        p0[i] = e0 / s;
        p1[i] = e1 / s;
    }
}

/* ------------------------------------------------------------------------- */
/* D. connected components: util/utils.py:474-494                             */
/* ------------------------------------------------------------------------- */

static int32_t uf_find(int32_t* parent, int32_t a)
{
    while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
    return a;
}
static void uf_union(int32_t* parent, int32_t a, int32_t b)
{
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a < b) parent[b] = a; else if (b < a) parent[a] = b;
}

typedef struct { int64_t key; int32_t root; } keyroot_t;
static int cmp_keyroot(const void* a, const void* b)
{
    int64_t ka = ((const keyroot_t*)a)->key, kb = ((const keyroot_t*)b)->key;
    return (ka > kb) - (ka < kb);
}

/* cv2.connectedComponentsWithStats(mask, connectivity=8) restated.
 * Label numbering: OpenCV's default 8-connectivity labeller scans 2x2 blocks in
 * raster order and numbers components by their first block, i.e. by
 * min over pixels of (y/2)*ceil(W/2) + x/2 (pinned against cv2
 * in tests/test_oracle_vs_libs.py).
 * stats rows: [left, top, width, height, area] int32; centroids: (sum_x/area,
 * sum_y/area) in double from integer sums; row 0 is the background.
 * Returns the label count including background, or -(count) if count > cap
 * (labels are still complete; stats/centroids hold the first cap rows). */
int psamo_ccl8(const uint8_t* mask, int H, int W, int32_t* labels,
               int32_t* stats, double* cent, int cap)
{
    int64_t n = (int64_t)H * W;
    int32_t* parent = (int32_t*)malloc(sizeof(int32_t) * n);
    for (int64_t i = 0; i < n; ++i) parent[i] = (int32_t)i;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int64_t i = (int64_t)y * W + x;
            if (!mask[i]) continue;
            if (x > 0 && mask[i - 1]) uf_union(parent, (int32_t)i, (int32_t)(i - 1));
            if (y > 0) {
                if (mask[i - W]) uf_union(parent, (int32_t)i, (int32_t)(i - W));
                if (x > 0 && mask[i - W - 1]) uf_union(parent, (int32_t)i, (int32_t)(i - W - 1));
                if (x < W - 1 && mask[i - W + 1]) uf_union(parent, (int32_t)i, (int32_t)(i - W + 1));
            }
        }
    int64_t bw = (W + 1) / 2;
    int64_t* minkey = (int64_t*)malloc(sizeof(int64_t) * n);
    for (int64_t i = 0; i < n; ++i) minkey[i] = INT64_MAX;
    int ncomp = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int64_t i = (int64_t)y * W + x;
            if (!mask[i]) continue;
            int32_t r = uf_find(parent, (int32_t)i);
            int64_t key = (int64_t)(y / 2) * bw + (x / 2);
            if (minkey[r] == INT64_MAX) ++ncomp;
            if (key < minkey[r]) minkey[r] = key;
        }
    keyroot_t* kr = (keyroot_t*)malloc(sizeof(keyroot_t) * (ncomp > 0 ? ncomp : 1));
    int k = 0;
    for (int64_t i = 0; i < n; ++i)
        if (mask[i] && parent[i] == (int32_t)i) { kr[k].key = minkey[i]; kr[k].root = (int32_t)i; ++k; }
    qsort(kr, ncomp, sizeof(keyroot_t), cmp_keyroot);
    int32_t* newlab = (int32_t*)minkey; /* reuse storage: root index -> label */
    for (int j = 0; j < ncomp; ++j) newlab[kr[j].root] = j + 1;
    int nlab = ncomp + 1;
    int rows = nlab < cap ? nlab : cap;
    int64_t* sx = (int64_t*)calloc(nlab, sizeof(int64_t));
    int64_t* sy = (int64_t*)calloc(nlab, sizeof(int64_t));
    int32_t* st = (int32_t*)malloc(sizeof(int32_t) * 5 * nlab);
    for (int j = 0; j < nlab; ++j) { st[5*j] = INT32_MAX; st[5*j+1] = INT32_MAX; st[5*j+2] = INT32_MIN; st[5*j+3] = INT32_MIN; st[5*j+4] = 0; }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int64_t i = (int64_t)y * W + x;
            int32_t l = mask[i] ? newlab[uf_find(parent, (int32_t)i)] : 0;
            labels[i] = l;
            int32_t* s5 = st + 5 * l;
            if (x < s5[0]) s5[0] = x;
            if (y < s5[1]) s5[1] = y;
            if (x > s5[2]) s5[2] = x;
            if (y > s5[3]) s5[3] = y;
            s5[4] += 1; sx[l] += x; sy[l] += y;
        }
    for (int j = 0; j < rows; ++j) {
        int32_t* s5 = st + 5 * j;
        if (s5[4] == 0) { /* empty background (all-foreground image): cv2 reports INT_MAX/INT_MIN arithmetic; keep zeros */
            stats[5*j] = 0; stats[5*j+1] = 0; stats[5*j+2] = 0; stats[5*j+3] = 0; stats[5*j+4] = 0;
            cent[2*j] = NAN; cent[2*j+1] = NAN;
            continue;
        }
        stats[5*j] = s5[0]; stats[5*j+1] = s5[1];
        stats[5*j+2] = s5[2] - s5[0] + 1; stats[5*j+3] = s5[3] - s5[1] + 1; stats[5*j+4] = s5[4];
        cent[2*j] = (double)sx[j] / (double)s5[4];
        cent[2*j+1] = (double)sy[j] / (double)s5[4];
    }
    free(parent); free(minkey); free(kr); free(sx); free(sy); free(st);
    return nlab <= cap ? nlab : -nlab;
}

/* ------------------------------------------------------------------------- */
/* E. per-component confidence: util/utils.py:485-490                         */
/* ------------------------------------------------------------------------- */

/* numpy's float32 pairwise summation over a contiguous array
 * (numpy/_core/src/umath/loops_utils.h.src, FLOAT_pairwise_sum). */
float psamo_pairwise_sum_f32(const float* a, int64_t n)
{
    if (n < 8) {
        float res = 0.0f;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return psamo_pairwise_sum_f32(a, n2) + psamo_pairwise_sum_f32(a + n2, n - n2);
    }
}

/* Same tree, evaluated for every component at once.  The reference sums
 * (p_fg.flatten() * (labels == j).flatten()) for each j (util/utils.py:490):
 * all terms outside component j are +0.0 and x + 0.0 == x exactly, so the sum
 * for j is the pairwise tree restricted to j's pixels.  acc[j] must hold nlab
 * zeros on entry of the top-level call; scratch tmp has nlab floats; touched
 * is a list of labels seen in the current leaf. */
static void pw_leaf(const float* a, const int32_t* lab, int64_t n, float* out, int32_t* seen, int* nseen, float* r8)
{
    /* out[l] receives the leaf partial for each label present; seen lists them */
    *nseen = 0;
    for (int64_t i = 0; i < n; ++i) {
        int32_t l = lab[i];
        if (l == 0) continue;
        int found = 0;
        for (int k = 0; k < *nseen; ++k) if (seen[k] == l) { found = 1; break; }
        if (!found) seen[(*nseen)++] = l;
    }
    for (int k = 0; k < *nseen; ++k) {
        int32_t l = seen[k];
        float res;
        if (n < 8) {
            res = 0.0f;
            for (int64_t i = 0; i < n; ++i) res += (lab[i] == l ? a[i] : 0.0f);
        } else {
            for (int j = 0; j < 8; ++j) r8[j] = (lab[j] == l ? a[j] : 0.0f);
            int64_t i;
            for (i = 8; i < n - (n % 8); i += 8)
                for (int j = 0; j < 8; ++j) r8[j] += (lab[i + j] == l ? a[i + j] : 0.0f);
            res = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]));
            for (; i < n; ++i) res += (lab[i] == l ? a[i] : 0.0f);
        }
        out[l] = res;
    }
}

typedef struct { int32_t* labs; float* vals; int n; } sparse_t;

static sparse_t pw_rec(const float* a, const int32_t* lab, int64_t n, float* dense, int32_t* seen, float* r8)
{
    sparse_t s;
    if (n <= 128) {
        int ns;
        pw_leaf(a, lab, n, dense, seen, &ns, r8);
        s.n = ns;
        s.labs = (int32_t*)malloc(sizeof(int32_t) * (ns > 0 ? ns : 1));
        s.vals = (float*)malloc(sizeof(float) * (ns > 0 ? ns : 1));
        /* keep sorted by label for merging */
        for (int i = 0; i < ns; ++i) {
            int32_t l = seen[i]; int j = i;
            while (j > 0 && s.labs[j - 1] > l) { s.labs[j] = s.labs[j - 1]; s.vals[j] = s.vals[j - 1]; --j; }
            s.labs[j] = l; s.vals[j] = dense[l];
        }
        return s;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    sparse_t L = pw_rec(a, lab, n2, dense, seen, r8);
    sparse_t R = pw_rec(a + n2, lab + n2, n - n2, dense, seen, r8);
    s.labs = (int32_t*)malloc(sizeof(int32_t) * (L.n + R.n > 0 ? L.n + R.n : 1));
    s.vals = (float*)malloc(sizeof(float) * (L.n + R.n > 0 ? L.n + R.n : 1));
    int i = 0, j = 0, k = 0;
    while (i < L.n || j < R.n) {
        if (j >= R.n || (i < L.n && L.labs[i] < R.labs[j])) { s.labs[k] = L.labs[i]; s.vals[k] = L.vals[i] + 0.0f; ++i; }
        else if (i >= L.n || R.labs[j] < L.labs[i]) { s.labs[k] = R.labs[j]; s.vals[k] = 0.0f + R.vals[j]; ++j; }
        else { s.labs[k] = L.labs[i]; s.vals[k] = L.vals[i] + R.vals[j]; ++i; ++j; }
        ++k;
    }
    s.n = k;
    free(L.labs); free(L.vals); free(R.labs); free(R.vals);
    return s;
}

/* sums[j] = float32 pairwise sum of p_fg over component j (sums[0] = 0). */
void psamo_cc_sums(const float* pfg, const int32_t* labels, int64_t n, int nlab, float* sums)
{
    float* dense = (float*)calloc(nlab > 0 ? nlab : 1, sizeof(float));
    int32_t seen[128];
    float r8[8];
    for (int j = 0; j < nlab; ++j) sums[j] = 0.0f;
    sparse_t s = pw_rec(pfg, labels, n, dense, seen, r8);
    for (int i = 0; i < s.n; ++i) sums[s.labs[i]] = s.vals[i];
    free(s.labs); free(s.vals); free(dense);
}

/* ------------------------------------------------------------------------- */
/* F. per-component prompt primitives: ProtoSAM.py:242-289                    */
/* ------------------------------------------------------------------------- */

/* torch.topk(v, 1) on a 1-D CPU tensor, restated (ATen TopKImpl.h, topk_impl_loop):
 *   n >= 64  -> std::partial_sort(begin, begin+1, end, greater): a heap-select that
 *               replaces the current top only on a strictly larger value, so the
 *               FIRST maximum wins;
 *   n <  64  -> std::nth_element(begin, begin, end, greater): libstdc++ introselect
 *               (median-of-3 pivot, unguarded partition, insertion sort below 4
 *               elements).  Which of several equal maxima ends at position 0 is a
 *               function of that algorithm, so it is replayed here move for move.
 * Returns the position (in the masked, raster-ordered value list) of the element
 * torch reports.  Values are finite probabilities; NaN ordering is not needed. */
typedef struct { float v; int32_t i; } tk_t;
#define TK_GT(a, b) ((a).v > (b).v)
static void tk_swap(tk_t* a, tk_t* b) { tk_t t = *a; *a = *b; *b = t; }

static void tk_move_median_to_first(tk_t* result, tk_t* a, tk_t* b, tk_t* c)
{
    if (TK_GT(*a, *b)) {
        if (TK_GT(*b, *c)) tk_swap(result, b);
        else if (TK_GT(*a, *c)) tk_swap(result, c);
        else tk_swap(result, a);
    } else if (TK_GT(*a, *c)) tk_swap(result, a);
    else if (TK_GT(*b, *c)) tk_swap(result, c);
    else tk_swap(result, b);
}

static tk_t* tk_unguarded_partition(tk_t* first, tk_t* last, tk_t* pivot)
{
    for (;;) {
        while (TK_GT(*first, *pivot)) ++first;
        --last;
        while (TK_GT(*pivot, *last)) --last;
        if (!(first < last)) return first;
        tk_swap(first, last);
        ++first;
    }
}

static void tk_insertion_sort(tk_t* first, tk_t* last)
{
    if (first == last) return;
    for (tk_t* i = first + 1; i != last; ++i) {
        tk_t val = *i;
        if (TK_GT(val, *first)) {
            for (tk_t* p = i; p != first; --p) *p = *(p - 1);
            *first = val;
        } else {
            tk_t* cur = i; tk_t* next = i - 1;
            while (TK_GT(val, *next)) { *cur = *next; cur = next; --next; }
            *cur = val;
        }
    }
}

int psamo_topk1_pos(const float* vals, int n)
{
    if (n <= 0) return -1;
    if (n >= 64) {                       /* partial_sort / heap-select with a 1-element heap */
        int best = 0;
        for (int i = 1; i < n; ++i) if (vals[i] > vals[best]) best = i;
        return best;
    }
    tk_t q[64];
    for (int i = 0; i < n; ++i) { q[i].v = vals[i]; q[i].i = i; }
    tk_t* first = q; tk_t* last = q + n; tk_t* nth = q;
    int depth = 0;
    for (int m = n; m > 1; m >>= 1) ++depth;   /* std::__lg(n) */
    depth *= 2;
    while (last - first > 3) {
        if (depth == 0) {                /* __heap_select(first, nth+1, last) + iter_swap(first, nth) */
            for (tk_t* i = first + 1; i < last; ++i) if (TK_GT(*i, *first)) tk_swap(i, first);
            return q[0].i;
        }
        --depth;
        tk_t* mid = first + (last - first) / 2;
        tk_move_median_to_first(first, first + 1, mid, last - 1);
        tk_t* cut = tk_unguarded_partition(first + 1, last, first);
        if (cut <= nth) first = cut; else last = cut;
    }
    tk_insertion_sort(first, last);
    return q[0].i;
}

/* torch.topk(v, k) (largest, sorted) on a 1-D CPU tensor for any k, restated from the same ATen loop
 * (get_most_conf_points with k > 1, models/ProtoSAM.py:266-289):
 *   k * 64 <= n -> std::partial_sort(begin, begin + k, end, greater)  = __heap_select + __sort_heap
 *   else        -> std::nth_element(begin, begin + k - 1, end, greater) then std::sort(begin, begin + k - 1, greater)
 * libstdc++'s algorithms replayed move for move (only values are compared, so the order of equal values is whatever
 * these algorithms leave).  Writes the k positions in torch's output order; returns 0, or -1 if k > n. */
static void tk_push_heap(tk_t* first, long hole, long top, tk_t value)
{
    long parent = (hole - 1) / 2;
    while (hole > top && TK_GT(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

static void tk_adjust_heap(tk_t* first, long hole, long len, tk_t value)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (TK_GT(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    tk_push_heap(first, hole, top, value);
}

static void tk_make_heap(tk_t* first, tk_t* last)
{
    const long len = last - first;
    if (len < 2) return;
    long parent = (len - 2) / 2;
    for (;;) {
        tk_t value = first[parent];
        tk_adjust_heap(first, parent, len, value);
        if (parent == 0) return;
        parent--;
    }
}

static void tk_pop_heap(tk_t* first, tk_t* last, tk_t* result)
{
    tk_t value = *result;
    *result = *first;
    tk_adjust_heap(first, 0, last - first, value);
}

static void tk_heap_select(tk_t* first, tk_t* middle, tk_t* last)
{
    tk_make_heap(first, middle);
    for (tk_t* i = middle; i < last; ++i)
        if (TK_GT(*i, *first)) tk_pop_heap(first, middle, i);
}

static void tk_sort_heap(tk_t* first, tk_t* last)
{
    while (last - first > 1) {
        --last;
        tk_pop_heap(first, last, last);
    }
}

static tk_t* tk_partition_pivot(tk_t* first, tk_t* last)
{
    tk_t* mid = first + (last - first) / 2;
    tk_move_median_to_first(first, first + 1, mid, last - 1);
    return tk_unguarded_partition(first + 1, last, first);
}

static int tk_lg(long n) { int d = 0; for (; n > 1; n >>= 1) ++d; return d; }

static void tk_introselect(tk_t* first, tk_t* nth, tk_t* last, int depth)
{
    while (last - first > 3) {
        if (depth == 0) {
            tk_heap_select(first, nth + 1, last);
            tk_swap(first, nth);
            return;
        }
        --depth;
        tk_t* cut = tk_partition_pivot(first, last);
        if (cut <= nth) first = cut; else last = cut;
    }
    tk_insertion_sort(first, last);
}

static void tk_introsort_loop(tk_t* first, tk_t* last, int depth)
{
    while (last - first > 16) {
        if (depth == 0) {
            tk_heap_select(first, last, last);
            tk_sort_heap(first, last);
            return;
        }
        --depth;
        tk_t* cut = tk_partition_pivot(first, last);
        tk_introsort_loop(cut, last, depth);
        last = cut;
    }
}

static void tk_unguarded_linear_insert(tk_t* last)
{
    tk_t val = *last;
    tk_t* next = last - 1;
    while (TK_GT(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}

static void tk_sort(tk_t* first, tk_t* last)
{
    if (first == last) return;
    tk_introsort_loop(first, last, tk_lg(last - first) * 2);
    if (last - first > 16) {
        tk_insertion_sort(first, first + 16);
        for (tk_t* i = first + 16; i != last; ++i) tk_unguarded_linear_insert(i);
    } else {
        tk_insertion_sort(first, last);
    }
}

int psamo_topk_pos(const float* vals, int n, int k, int32_t* out_pos)
{
    if (k < 1 || k > n) return -1;
    tk_t* q = (tk_t*)malloc((size_t)n * sizeof(tk_t));
    for (int i = 0; i < n; ++i) { q[i].v = vals[i]; q[i].i = i; }
    if ((long)k * 64 <= n) {
        tk_heap_select(q, q + k, q + n);
        tk_sort_heap(q, q + k);
    } else {
        tk_t* nth = q + k - 1;
        if (nth != q + n) tk_introselect(q, nth, q + n, tk_lg(n) * 2);
        tk_sort(q, q + k - 1);
    }
    for (int i = 0; i < k; ++i) out_pos[i] = q[i].i;
    free(q);
    return 0;
}

/* For every label j in 1..nlab-1: bbox [min_x,min_y,max_x,max_y] (inclusive,
 * get_bbox_per_cc :242-264) and the most confident pixel (x,y) =
 * torch.nonzero(mask)[torch.topk(p_fg[mask], 1).indices] (get_most_conf_points
 * :266-289 with k=1), tie behaviour per psamo_topk1_pos. */
void psamo_cc_prompts(const float* pfg, const int32_t* labels, int H, int W, int nlab,
                      int64_t* bbox /*[nlab,4]*/, int64_t* confpt /*[nlab,2]*/, float* confval /*[nlab]*/)
{
    int32_t* area = (int32_t*)calloc(nlab > 0 ? nlab : 1, sizeof(int32_t));
    for (int j = 0; j < nlab; ++j) {
        bbox[4*j] = INT64_MAX; bbox[4*j+1] = INT64_MAX; bbox[4*j+2] = -1; bbox[4*j+3] = -1;
        confpt[2*j] = -1; confpt[2*j+1] = -1; confval[j] = -INFINITY;
    }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int64_t i = (int64_t)y * W + x;
            int32_t l = labels[i];
            if (l <= 0 || l >= nlab) continue;
            area[l] += 1;
            if (x < bbox[4*l]) bbox[4*l] = x;
            if (y < bbox[4*l+1]) bbox[4*l+1] = y;
            if (x > bbox[4*l+2]) bbox[4*l+2] = x;
            if (y > bbox[4*l+3]) bbox[4*l+3] = y;
            if (pfg[i] > confval[l]) { confval[l] = pfg[i]; confpt[2*l] = x; confpt[2*l+1] = y; }
        }
    for (int l = 1; l < nlab; ++l) {
        if (area[l] == 0 || area[l] >= 64) continue;
        float vals[64]; int64_t px[64], py[64]; int n = 0;
        for (int64_t y = bbox[4*l+1]; y <= bbox[4*l+3]; ++y)
            for (int64_t x = bbox[4*l]; x <= bbox[4*l+2]; ++x)
                if (labels[y * W + x] == l) { vals[n] = pfg[y * W + x]; px[n] = x; py[n] = y; ++n; }
        int pos = psamo_topk1_pos(vals, n);
        confval[l] = vals[pos]; confpt[2*l] = px[pos]; confpt[2*l+1] = py[pos];
    }
    free(area);
}

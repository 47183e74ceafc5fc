// Kernel 3b -- mask bits + p_fg -> connected components -> SAM prompt records (sm_100a).
//
// Replaces, per image (one query slice x one label), the CPU tail of ProtoSAM.forward:
//   cv2.connectedComponentsWithStats(pred, connectivity=8)       util/utils.py:478
//   per-component confidence and `cca` selection                 util/utils.py:485-541
//   get_bbox_per_cc, get_most_conf_points, centroids             models/ProtoSAM.py:242-289, 349-450
// The reference moves the 1024^2 maps to the host and makes O(ncc) full-image numpy passes;
// here one persistent CTA per image works on the 128 KB bit mask:
//   rows -> foreground runs (ballot/popc) -> lock-free union-find over RUNS (8-connectivity =
//   runs of adjacent rows overlapping within one pixel) -> component order = OpenCV's label
//   order (rank of the component's first 2x2 block in block-raster order, from a 32 KB bitmap
//   + popcount prefix, no sort) -> exact integer statistics with atomics.
// p_fg is read only under foreground runs.  Everything is integer arithmetic, so results do
// not depend on scheduling: p_fg of a foreground pixel lies in [0.5,1], i.e. is a multiple of
// 2^-24, and its sums are accumulated exactly in 64-bit integers.
//
// Roofline: latency/HBM bound, small: algorithmic bytes per image = out^2/8 (bits) + 4*n_fg
// (p_fg under the mask) read + 96*ncc written.
#include <cstdlib>
#include <cstring>

#include "psam_common.cuh"
#include "psam_topk.cuh"

namespace psam {

constexpr int CT = 512;          // threads per CTA: the kernel is latency-bound, so per image 512 threads cost little
                                 // time, and a 512-thread, 24 K-register, ~13 KB-shared-memory CTA fits on an SM
                                 // beside a GEMM CTA and two block-kernel CTAs of other volumes (1024 threads did not)
constexpr int MAXR = 2048;       // components ranked through the shared-memory key list (more: global bitmap path)
constexpr int MAX_OUT = 1024;    // image side supported by the static tables below
constexpr int KEY_WORDS = (MAX_OUT / 2) * (MAX_OUT / 2) / 32;  // bitmap over 2x2 blocks

struct Acc {
    unsigned long long sumx, sumy, sump, best;
    unsigned int area, minx, miny, maxx, maxy;
    int root;
};

struct CompParams {
    const uint32_t* maskbits;
    const float* p_fg;
    const uint2* wstat;      // per-word (sum, best) from kernel 3a, or NULL: read p_fg pixel by pixel
    int n_img, out, use_cca, max_cc, max_runs;
    psam_image_hdr* hdr;
    psam_prompt_rec* recs;
    int32_t* labels_out;
    // per-CTA scratch (index = blockIdx.x)
    uint16_t *run_s, *run_e, *run_y;
    int32_t* parent;
    uint32_t* minkey;
    uint32_t* gkeys;         // per CTA: 2 * KEY_WORDS words (bitmap + prefix of the > MAXR-components path)
    unsigned long long* sump;
    int32_t* rank;
    Acc* acc;
};

__device__ __forceinline__ int uf_find(volatile int32_t* parent, int x)
{
    int p = parent[x];
    while (p != x) { x = p; p = parent[x]; }
    return x;
}

__device__ __forceinline__ void uf_union(int32_t* parent, int a, int b)
{
    volatile int32_t* vp = parent;
    for (;;) {
        a = uf_find(vp, a);
        b = uf_find(vp, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }   // attach the larger root under the smaller
        const int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

// sum of the bit positions of the set bits of w
__device__ __forceinline__ int bitpos_sum(uint32_t w)
{
    return __popc(w & 0xAAAAAAAAu) + 2 * __popc(w & 0xCCCCCCCCu) + 4 * __popc(w & 0xF0F0F0F0u) +
           8 * __popc(w & 0xFF00FF00u) + 16 * __popc(w & 0xFFFF0000u);
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int o)
{
    unsigned int lo = (unsigned int)v, hi = (unsigned int)(v >> 32);
    lo = __shfl_xor_sync(0xffffffffu, lo, o);
    hi = __shfl_xor_sync(0xffffffffu, hi, o);
    return ((unsigned long long)hi << 32) | lo;
}

// Exact statistics of one foreground run [s,e] of row y, computed by a whole warp:
//   sum  = sum of p_fg * 2^24 over the run (integer, exact)
//   best = max over the run of (p bits << 32 | ~raster index): highest p_fg, then first pixel.
// A word whose foreground bits all belong to this run contributes through kernel 3a's per-word
// statistics (8 bytes instead of 128); words shared with another run fall back to the pixels.
__device__ __forceinline__ void run_stats(const uint32_t* __restrict__ bits, const float* __restrict__ pfg,
                                          const uint2* __restrict__ wstat, int out, int wpr, int y, int s, int e,
                                          int lane, unsigned long long& sum, unsigned long long& best)
{
    sum = 0; best = 0;
    const int wa = s >> 5, wb = e >> 5;
    for (int j = wa + lane; j <= wb; j += 32) {
        const int lo = (j == wa) ? (s & 31) : 0, hi = (j == wb) ? (e & 31) : 31;
        const uint32_t fm = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
        const uint32_t m = bits[(size_t)y * wpr + j];
        if (wstat && (m & ~fm) == 0u) {
            const uint2 st = wstat[(size_t)y * wpr + j];
            sum += st.x;
            const uint32_t k = st.y >> 5, x = j * 32 + 31 - (st.y & 31);
            // k = p * 2^24 in [2^23, 2^24]  ->  float bits of p
            const uint32_t pb = (k == 16777216u) ? 0x3f800000u : (0x3f000000u + ((k - 8388608u)));
            best = max(best, ((unsigned long long)pb << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(y * out + x)));
        } else {
            for (int b = lo; b <= hi; ++b) {
                const int x = j * 32 + b;
                const float pv = pfg[(size_t)y * out + x];
                sum += (unsigned long long)(pv * 16777216.0f);
                best = max(best, ((unsigned long long)__float_as_uint(pv) << 32) |
                                     (unsigned long long)(0xFFFFFFFFu - (unsigned)(y * out + x)));
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += shfl_xor_u64(sum, o);
        best = max(best, shfl_xor_u64(best, o));
    }
}

// The same statistics computed by ONE thread (used where every thread of the CTA owns a different run: thousands
// of independent loads in flight instead of one run per warp).  Words are taken four at a time so that the
// mask-word and per-word-statistics loads of a group are issued before any of them is consumed.
__device__ __forceinline__ void run_stats_thread(const uint32_t* __restrict__ bits, const float* __restrict__ pfg,
                                                 const uint2* __restrict__ wstat, int out, int wpr, int y, int s, int e,
                                                 unsigned long long& sum, unsigned long long& best)
{
    sum = 0; best = 0;
    const int wa = s >> 5, wb = e >> 5;
    const uint32_t* brow = bits + (size_t)y * wpr;
    const uint2* srow = wstat ? wstat + (size_t)y * wpr : nullptr;
    for (int j0 = wa; j0 <= wb; j0 += 4) {
        uint32_t m[4];
        uint2 st[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = min(j0 + u, wb);
            m[u] = brow[j];
            st[u] = srow ? srow[j] : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            if (j > wb) break;
            const int lo = (j == wa) ? (s & 31) : 0, hi = (j == wb) ? (e & 31) : 31;
            const uint32_t fm = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
            if (srow && (m[u] & ~fm) == 0u) {
                sum += st[u].x;
                const uint32_t k = st[u].y >> 5, x = j * 32 + 31 - (st[u].y & 31);
                const uint32_t pb = (k == 16777216u) ? 0x3f800000u : (0x3f000000u + ((k - 8388608u)));
                best = max(best, ((unsigned long long)pb << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(y * out + x)));
            } else {
                for (int b = lo; b <= hi; ++b) {
                    const int x = j * 32 + b;
                    const float pv = pfg[(size_t)y * out + x];
                    sum += (unsigned long long)(pv * 16777216.0f);
                    best = max(best, ((unsigned long long)__float_as_uint(pv) << 32) |
                                         (unsigned long long)(0xFFFFFFFFu - (unsigned)(y * out + x)));
                }
            }
        }
    }
}

__global__ void __maxnreg__(48) k_components(CompParams P)
{
    __shared__ int s_rowstart[MAX_OUT + 1];
    __shared__ __align__(8) uint32_t s_keys[MAXR];        // first-block keys of the components; later the topk scratch
    uint32_t* s_bitmap = P.gkeys + (size_t)blockIdx.x * 2 * KEY_WORDS;   // global fallback for > MAXR components
    uint32_t* s_prefix = s_bitmap + KEY_WORDS;
    __shared__ int s_scan[32];
    __shared__ unsigned long long s_red64[32];
    __shared__ int s_misc[16];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int out = P.out, wpr = (out + 31) >> 5, bw = (out + 1) >> 1;
    const int key_words = (bw * bw + 31) >> 5;
    const size_t so = (size_t)blockIdx.x * P.max_runs;
    uint16_t* run_s = P.run_s + so;
    uint16_t* run_e = P.run_e + so;
    uint16_t* run_y = P.run_y + so;
    int32_t* parent = P.parent + so;
    uint32_t* minkey = P.minkey + so;
    unsigned long long* sump = P.sump + so;
    int32_t* rank = P.rank + so;
    Acc* acc = P.acc + (size_t)blockIdx.x * P.max_cc;
    TraceRec* tr = tid == 0 ? trace_begin(4) : nullptr;

    for (int img = blockIdx.x; img < P.n_img; img += gridDim.x) {
        const uint32_t* bits = P.maskbits + (size_t)img * out * wpr;
        const float* pfg = P.p_fg + (size_t)img * out * out;
        const uint2* wst = P.wstat ? P.wstat + (size_t)img * out * wpr : nullptr;
        psam_image_hdr* hdr = P.hdr + img;
        psam_prompt_rec* recs = P.recs + (size_t)img * P.max_cc;
        int32_t* labels = P.labels_out ? P.labels_out + (size_t)img * out * out : nullptr;

        // ---- S1: runs per row, foreground count, background box/sums --------------------
        int npix = 0;
        unsigned int bminx = 0xffffffffu, bminy = 0xffffffffu, bmaxx = 0, bmaxy = 0;
        unsigned long long bsx = 0, bsy = 0;
        int bany = 0;
        for (int i = tid; i <= MAX_OUT; i += CT) s_rowstart[i] = 0;
        __syncthreads();
        for (int yb = wid; yb < out; yb += 4 * (CT / 32)) {
          uint32_t wq[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {     // four independent row loads in flight per warp
              const int y = yb + u * (CT / 32);
              wq[u] = (lane < wpr && y < out) ? bits[(size_t)y * wpr + lane] : 0u;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int y = yb + u * (CT / 32);
            if (y >= out) break;
            const uint32_t word = wq[u];
            uint32_t prev_msb = __shfl_up_sync(0xffffffffu, word >> 31, 1);
            if (lane == 0) prev_msb = 0;
            const uint32_t starts = word & ~((word << 1) | prev_msb);
            const int tot = __reduce_add_sync(0xffffffffu, __popc(starts));
            npix += __popc(word);
            if (lane < wpr) {
                // bits beyond the last column (out % 32 != 0) are neither foreground nor background
                const uint32_t inv = ~word & ((lane == wpr - 1 && (out & 31)) ? (0xffffffffu >> (32 - (out & 31))) : 0xffffffffu);
                if (inv) {
                    bany = 1;
                    const unsigned int x0 = lane * 32 + (__ffs(inv) - 1), x1 = lane * 32 + 31 - __clz(inv);
                    bminx = min(bminx, x0); bmaxx = max(bmaxx, x1);
                    bminy = min(bminy, (unsigned)y); bmaxy = max(bmaxy, (unsigned)y);
                    const int c = __popc(inv);
                    bsx += (unsigned long long)(lane * 32) * c + bitpos_sum(inv);
                    bsy += (unsigned long long)y * c;
                }
            }
            if (lane == 0) s_rowstart[y] = tot;
          }
        }
        // block reductions of the S1 scalars
        npix = warp_sum_i(npix);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bminx = min(bminx, __shfl_xor_sync(0xffffffffu, bminx, o));
            bminy = min(bminy, __shfl_xor_sync(0xffffffffu, bminy, o));
            bmaxx = max(bmaxx, __shfl_xor_sync(0xffffffffu, bmaxx, o));
            bmaxy = max(bmaxy, __shfl_xor_sync(0xffffffffu, bmaxy, o));
            bsx += shfl_xor_u64(bsx, o);
            bsy += shfl_xor_u64(bsy, o);
            bany |= __shfl_xor_sync(0xffffffffu, bany, o);
        }
        if (tid == 0) { s_misc[0] = 0; s_misc[1] = 0x7fffffff; s_misc[2] = 0x7fffffff; s_misc[3] = 0; s_misc[4] = 0; s_misc[5] = 0; }
        if (tid < 2) s_red64[tid] = 0;
        __syncthreads();
        if (lane == 0) {
            atomicAdd(&s_misc[0], npix);
            if (bany) {
                atomicMin(&s_misc[1], (int)bminx); atomicMin(&s_misc[2], (int)bminy);
                atomicMax(&s_misc[3], (int)bmaxx); atomicMax(&s_misc[4], (int)bmaxy);
                atomicAdd(&s_red64[0], bsx); atomicAdd(&s_red64[1], bsy);
                s_misc[5] = 1;
            }
        }
        __syncthreads();
        const int n_fg = s_misc[0];
        const int n_bg = out * out - n_fg;

        // ---- S2: exclusive scan of the per-row run counts --------------------------------
        {
            constexpr int RP = (MAX_OUT + CT - 1) / CT;          // consecutive rows per thread
            int cnt[RP], v = 0;
#pragma unroll
            for (int k = 0; k < RP; ++k) {
                const int y = tid * RP + k;
                cnt[k] = y < out ? s_rowstart[y] : 0;
                v += cnt[k];
            }
            const int ex = warp_excl_scan_i(v, lane);
            if (lane == 31) s_scan[wid] = ex + v;
            __syncthreads();
            if (wid == 0) {
                const int t = lane < CT / 32 ? s_scan[lane] : 0;
                const int e2 = warp_excl_scan_i(t, lane);
                s_scan[lane] = e2;
                if (lane == 31) s_misc[6] = e2 + t;
            }
            __syncthreads();
            int run = s_scan[wid] + ex;
#pragma unroll
            for (int k = 0; k < RP; ++k) {
                const int y = tid * RP + k;
                if (y < out) s_rowstart[y] = run;
                run += cnt[k];
            }
            if (tid == 0) s_rowstart[out] = s_misc[6];
            __syncthreads();
        }
        const int total = s_rowstart[out];

        if (tid == 0) {
            hdr->n_fg = n_fg;
            hdr->n_runs = total;
            hdr->bg_stats[0] = s_misc[5] ? s_misc[1] : 0;
            hdr->bg_stats[1] = s_misc[5] ? s_misc[2] : 0;
            hdr->bg_stats[2] = s_misc[5] ? s_misc[3] - s_misc[1] + 1 : 0;
            hdr->bg_stats[3] = s_misc[5] ? s_misc[4] - s_misc[2] + 1 : 0;
            hdr->bg_stats[4] = n_bg;
            hdr->bg_centroid[0] = n_bg ? (double)s_red64[0] / (double)n_bg : 0.0;
            hdr->bg_centroid[1] = n_bg ? (double)s_red64[1] / (double)n_bg : 0.0;
            hdr->selected = 0;
            hdr->reserved = 0;
        }
        if (total == 0 || total > P.max_runs) {
            if (tid == 0) {
                hdr->ncc = 0;
                hdr->n_rec = 0;
                hdr->flags = total == 0 ? PSAM_IMG_EMPTY : PSAM_IMG_RUN_OVERFLOW;
            }
            __syncthreads();
            continue;
        }

        // ---- S3: materialise the runs -----------------------------------------------------
        for (int yb = wid; yb < out; yb += 4 * (CT / 32)) {
          uint32_t wq[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {     // four independent row loads in flight per warp
              const int y = yb + u * (CT / 32);
              wq[u] = (lane < wpr && y < out) ? bits[(size_t)y * wpr + lane] : 0u;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int y = yb + u * (CT / 32);
            if (y >= out) break;
            const uint32_t word = wq[u];
            uint32_t prev_msb = __shfl_up_sync(0xffffffffu, word >> 31, 1);
            if (lane == 0) prev_msb = 0;
            uint32_t next_lsb = __shfl_down_sync(0xffffffffu, word & 1u, 1);
            if (lane == 31) next_lsb = 0;
            uint32_t starts = word & ~((word << 1) | prev_msb);
            uint32_t ends = word & ~((word >> 1) | (next_lsb << 31));
            // ends before this lane = starts before this lane - (a run is open on entry to this lane)
            int si = s_rowstart[y] + warp_excl_scan_i(__popc(starts), lane);
            int ei = si - (int)(prev_msb & word & 1u);
            while (starts) {
                const int b = __ffs(starts) - 1;
                starts &= starts - 1;
                run_s[si] = (uint16_t)(lane * 32 + b);
                run_y[si] = (uint16_t)y;
                parent[si] = si;
                minkey[si] = 0xffffffffu;
                sump[si] = 0ull;
                ++si;
            }
            while (ends) {
                const int b = __ffs(ends) - 1;
                ends &= ends - 1;
                run_e[ei] = (uint16_t)(lane * 32 + b);
                ++ei;
            }
          }
        }
        if (tid == 0) s_misc[10] = 0;
        __syncthreads();

        // ---- S4: connect runs of adjacent rows that touch (8-connectivity) ----------------------
        // A blob is a chain of ~1000 vertically stacked runs, so plain union-find degenerates into
        // O(rows) pointer chasing per run.  Instead: (a) link every run to its FIRST neighbour in the
        // row above (a forest, no atomics), (b) pointer-jump until flat (log2(depth) rounds), (c) merge
        // the trees joined by the remaining neighbours with lock-free unions on now-shallow trees,
        // (d) pointer-jump again.
        for (int i = tid; i < total; i += CT) {
            const int y = run_y[i];
            int link = i;
            if (y > 0) {
                const int s = run_s[i], e = run_e[i];
                int l = s_rowstart[y - 1], r = s_rowstart[y];
                const int hi = r;           // first j with run_e[j] >= s - 1
                while (l < r) {
                    const int m = (l + r) >> 1;
                    if ((int)run_e[m] < s - 1) l = m + 1; else r = m;
                }
                if (l < hi && (int)run_s[l] <= e + 1) link = l;
                rank[i] = l;                // remembered for (c)
            }
            parent[i] = link;
        }
        __syncthreads();
        for (int pass = 0; pass < 2; ++pass) {
            for (;;) {                      // (b)/(d) pointer jumping
                int changed = 0;
                for (int i = tid; i < total; i += CT) {
                    const int pp = ((volatile int32_t*)parent)[i];
                    const int gp = ((volatile int32_t*)parent)[pp];
                    if (gp != pp) { parent[i] = gp; changed = 1; }
                }
                if (!__syncthreads_or(changed)) break;
            }
            if (pass == 1) break;
            for (int i = tid; i < total; i += CT) {   // (c) remaining neighbours
                const int y = run_y[i];
                if (y == 0) continue;
                const int e = run_e[i], hi = s_rowstart[y];
                for (int j = rank[i] + 1; j < hi && (int)run_s[j] <= e + 1; ++j) uf_union(parent, i, j);
            }
            __syncthreads();
        }
        // ---- S5/S6: first 2x2 block (block-raster order) of every component -------------------
        for (int i = tid; i < total; i += CT) {
            const uint32_t key = (uint32_t)(run_y[i] >> 1) * bw + (run_s[i] >> 1);
            atomicMin(&minkey[parent[i]], key);
        }
        __syncthreads();
        // ---- S7-S9: OpenCV label = 1 + rank of that block among all components' first blocks.  A 2x2 block holds pixels
        // of one component only, so the keys are distinct.  Up to MAXR components: their keys are collected in shared
        // memory and every root counts the smaller ones; beyond that (speckle): bitmap over the blocks + popcount prefix
        // in this CTA's global scratch.
        for (int i = tid; i < total; i += CT)
            if (parent[i] == i) {
                const int slot = atomicAdd(&s_misc[10], 1);
                if (slot < MAXR) s_keys[slot] = minkey[i];
            }
        __syncthreads();
        const int ncc = s_misc[10];
        if (ncc <= MAXR) {
            for (int i = tid; i < total; i += CT)
                if (parent[i] == i) {
                    const uint32_t k = minkey[i];
                    int r = 0;
                    for (int j = 0; j < ncc; ++j) r += s_keys[j] < k;
                    rank[i] = r;
                }
        } else {
            for (int i = tid; i < key_words; i += CT) s_bitmap[i] = 0u;
            __syncthreads();
            for (int i = tid; i < total; i += CT)
                if (parent[i] == i) atomicOr(&s_bitmap[minkey[i] >> 5], 1u << (minkey[i] & 31));
            __syncthreads();
            const int per = (key_words + CT - 1) / CT;
            const int w0 = tid * per;
            int local = 0;
            for (int k = 0; k < per; ++k)
                if (w0 + k < key_words) local += __popc(s_bitmap[w0 + k]);
            const int ex = warp_excl_scan_i(local, lane);
            if (lane == 31) s_scan[wid] = ex + local;
            __syncthreads();
            if (wid == 0) {
                const int t = lane < CT / 32 ? s_scan[lane] : 0;
                const int e2 = warp_excl_scan_i(t, lane);
                s_scan[lane] = e2;
            }
            __syncthreads();
            int run = s_scan[wid] + ex;
            for (int k = 0; k < per; ++k)
                if (w0 + k < key_words) { s_prefix[w0 + k] = run; run += __popc(s_bitmap[w0 + k]); }
            __syncthreads();
            for (int i = tid; i < total; i += CT)
                if (parent[i] == i) {
                    const uint32_t k = minkey[i];
                    rank[i] = (int)s_prefix[k >> 5] + __popc(s_bitmap[k >> 5] & ((1u << (k & 31)) - 1u));
                }
        }
        __syncthreads();

        // ---- S10: use_cca -- exact per-component sums of p_fg, keep the largest ------------
        int sel_root = -1, flags = 0;
        if (P.use_cca) {
            for (int i = wid; i < total; i += CT / 32) {
                unsigned long long acc_p, b_unused;
                run_stats(bits, pfg, wst, out, wpr, run_y[i], run_s[i], run_e[i], lane, acc_p, b_unused);
                if (lane == 0) atomicAdd(&sump[parent[i]], acc_p);
            }
            __syncthreads();
            // strictly largest confidence, first label on ties (util/utils.py:511-515): the largest exact sum, then the
            // smallest rank among the components that reach it (two reductions: a packed (sum << 20 | rank) key would
            // overflow 64 bits for a full-frame component, sum = 2^44)
            unsigned long long best = 0;
            for (int i = tid; i < total; i += CT)
                if (parent[i] == i) best = max(best, sump[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = max(best, shfl_xor_u64(best, o));
            if (lane == 0) s_red64[wid] = best;
            if (tid == 0) s_misc[9] = 0x7fffffff;
            __syncthreads();
            best = 0;
            for (int k = 0; k < CT / 32; ++k) best = max(best, s_red64[k]);
            const unsigned long long best_sum = best;
            for (int i = tid; i < total; i += CT)
                if (parent[i] == i && sump[i] == best_sum) atomicMin(&s_misc[9], rank[i]);
            __syncthreads();
            const int best_rank = s_misc[9];
            unsigned long long second = 0;
            for (int i = tid; i < total; i += CT)
                if (parent[i] == i) {
                    if (rank[i] == best_rank) s_misc[8] = i;
                    else second = max(second, sump[i]);
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) second = max(second, shfl_xor_u64(second, o));
            if (lane == 0) s_red64[wid] = second;
            __syncthreads();
            second = 0;
            for (int k = 0; k < CT / 32; ++k) second = max(second, s_red64[k]);
            sel_root = s_misc[8];
            // the reference compares fp32 pairwise sums (relative error < 4e-6): flag near-ties
            if ((double)second * (1.0 + 1e-5) >= (double)best_sum) flags |= PSAM_IMG_CCA_AMBIGUOUS;
            __syncthreads();
        }

        // ---- S11: statistics of the emitted components -------------------------------------
        const int n_rec = P.use_cca ? 1 : min(ncc, P.max_cc);
        if (!P.use_cca && ncc > P.max_cc) flags |= PSAM_IMG_CC_TRUNCATED;
        for (int r = tid; r < n_rec; r += CT) {
            Acc a;
            a.sumx = 0; a.sumy = 0; a.sump = 0; a.best = 0;
            a.area = 0; a.minx = 0xffffffffu; a.miny = 0xffffffffu; a.maxx = 0; a.maxy = 0; a.root = -1;
            acc[r] = a;
        }
        __syncthreads();
        for (int i = tid; i < total; i += CT) {          // one thread per run
            const int root = parent[i];
            const int slot = P.use_cca ? (root == sel_root ? 0 : -1) : (rank[root] < P.max_cc ? rank[root] : -1);
            if (slot < 0) continue;
            const int y = run_y[i], s = run_s[i], e = run_e[i];
            unsigned long long acc_p, best;
            run_stats_thread(bits, pfg, wst, out, wpr, y, s, e, acc_p, best);
            Acc* a = acc + slot;
            const unsigned int len = e - s + 1;
            atomicAdd(&a->area, len);
            atomicAdd(&a->sumx, (unsigned long long)(s + e) * len / 2);
            atomicAdd(&a->sumy, (unsigned long long)y * len);
            atomicAdd(&a->sump, acc_p);
            atomicMax(&a->best, best);
            atomicMin(&a->minx, (unsigned)s); atomicMax(&a->maxx, (unsigned)e);
            atomicMin(&a->miny, (unsigned)y); atomicMax(&a->maxy, (unsigned)y);
            if (i == root) a->root = root;
        }
        __syncthreads();

        // ---- S12: records (label order) ------------------------------------------------------------
        for (int r = tid; r < n_rec; r += CT) {
            const Acc a = acc[r];
            psam_prompt_rec rec;
            rec.box[0] = a.minx; rec.box[1] = a.miny; rec.box[2] = a.maxx; rec.box[3] = a.maxy;
            const unsigned int idx = 0xFFFFFFFFu - (unsigned int)(a.best & 0xFFFFFFFFu);
            rec.conf_pt[0] = idx % out;
            rec.conf_pt[1] = idx / out;
            rec.centroid[0] = (double)a.sumx / (double)a.area;
            rec.centroid[1] = (double)a.sumy / (double)a.area;
            rec.conf = ((double)a.sump * (1.0 / 16777216.0)) / ((double)n_fg + 1e-6);
            rec.conf_pt_p = __uint_as_float((unsigned int)(a.best >> 32));
            rec.area = (int)a.area;
            rec.label = rank[a.root] + 1;
            rec.flags = P.use_cca ? PSAM_REC_SELECTED : 0;
            rec.reserved[0] = rec.reserved[1] = 0;
            recs[r] = rec;
            if (P.use_cca) hdr->selected = rec.label;
        }
        if (tid == 0) s_misc[9] = 0;
        __syncthreads();
        // ---- S13: components of < 64 pixels: torch.topk(v, 1) takes its nth_element path there, which on TIES of
        // the maximum does not return the first pixel in raster order.  One warp per such component gathers the
        // component's pixels in raster order (lane = candidate pixel of the bounding box, membership = the run
        // under it belongs to the component) and, only if the maximum is attained more than once, replays
        // libstdc++'s nth_element on them.  A unique maximum needs no replay: every algorithm returns it.
        {
            TK* q = reinterpret_cast<TK*>(s_keys) + wid * 64;           // the key list is free again: 16 warps x 64 x 8 B = 8 KB
            for (int r = wid; r < n_rec; r += CT / 32) {
                const Acc a = acc[r];
                if (a.area >= 64) continue;
                const int bwid = (int)(a.maxx - a.minx) + 1, bh = (int)(a.maxy - a.miny) + 1, ncand = bwid * bh;
                int n = 0;
                for (int c0 = 0; c0 < ncand && n < (int)a.area; c0 += 32) {
                    const int c = c0 + lane;
                    bool in = false;
                    int x = 0, y = 0;
                    if (c < ncand) {
                        y = (int)a.miny + c / bwid;
                        x = (int)a.minx + c % bwid;
                        if ((bits[(size_t)y * wpr + (x >> 5)] >> (x & 31)) & 1u) {
                            int l = s_rowstart[y], rr = s_rowstart[y + 1];   // the run of row y that ends at or after x
                            while (l < rr) {
                                const int m = (l + rr) >> 1;
                                if ((int)run_e[m] < x) l = m + 1; else rr = m;
                            }
                            in = parent[l] == a.root;
                        }
                    }
                    const uint32_t inb = __ballot_sync(0xffffffffu, in);
                    if (in) {
                        const int pos = n + __popc(inb & ((1u << lane) - 1u));
                        q[pos].v = pfg[(size_t)y * out + x];
                        q[pos].i = y * out + x;
                    }
                    n += __popc(inb);
                }
                __syncwarp();
                float vmax = -1.0f;
                for (int i = lane; i < n; i += 32) vmax = fmaxf(vmax, q[i].v);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
                int ties = 0;
                for (int i = lane; i < n; i += 32) ties += q[i].v == vmax;
                ties = warp_sum_i(ties);
                if (ties > 1 && lane == 0) {
                    const int idx = topk1_small(q, n);
                    recs[r].conf_pt[0] = idx % out;
                    recs[r].conf_pt[1] = idx / out;
                    recs[r].conf_pt_p = pfg[idx];
                }
                __syncwarp();
            }
        }
        if (tid == 0) {
            hdr->ncc = ncc;
            hdr->n_rec = n_rec;
            hdr->flags = flags;
        }
        // ---- optional label image (function-level API: cv2-style tuple) ---------------------
        if (labels) {
            for (int i = wid; i < total; i += CT / 32) {
                const int root = parent[i];
                const int lab = P.use_cca ? (root == sel_root ? 1 : 0) : rank[root] + 1;
                const int y = run_y[i], s = run_s[i], e = run_e[i];
                for (int x = s + lane; x <= e; x += 32) labels[(size_t)y * out + x] = lab;
            }
        }
        __syncthreads();
    }
    trace_end(tr);
}

// ------------------------------------------------------------------------------------------------
// Records -> the tensors SamPredictor.predict_torch takes (models/segment_anything/predictor.py:136-167):
// ResizeLongestSide.apply_coords / apply_boxes (models/segment_anything/utils/transforms.py:40-62) scale x by
// new_w/old_w and y by new_h/old_h in double precision, torch.as_tensor(..., dtype=torch.float) rounds to fp32.
// One thread per (image, component slot); slots beyond n_rec are zero-filled.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_records_to_sam(const psam_image_hdr* __restrict__ hdr,
                                                        const psam_prompt_rec* __restrict__ recs, int n_img, int max_cc,
                                                        int point_mode, double sx, double sy, float* __restrict__ points,
                                                        int32_t* __restrict__ labels, float* __restrict__ boxes)
{
    const int npts = point_mode == 2 ? 2 : 1;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_img * max_cc) return;
    const int img = i / max_cc, r = i - img * max_cc;
    float* pt = points + (size_t)i * npts * 2;
    int32_t* lb = labels + (size_t)i * npts;
    float* bx = boxes + (size_t)i * 4;
    if (r >= hdr[img].n_rec) {
        for (int k = 0; k < npts * 2; ++k) pt[k] = 0.f;
        for (int k = 0; k < npts; ++k) lb[k] = 0;
        for (int k = 0; k < 4; ++k) bx[k] = 0.f;
        return;
    }
    const psam_prompt_rec rec = recs[i];
    int k = 0;
    if (point_mode != 1) {          // most confident point
        pt[k++] = (float)((double)rec.conf_pt[0] * sx);
        pt[k++] = (float)((double)rec.conf_pt[1] * sy);
    }
    if (point_mode != 0) {          // centroid
        pt[k++] = (float)(rec.centroid[0] * sx);
        pt[k++] = (float)(rec.centroid[1] * sy);
    }
    for (int j = 0; j < npts; ++j) lb[j] = 1;
    bx[0] = (float)((double)rec.box[0] * sx);
    bx[1] = (float)((double)rec.box[1] * sy);
    bx[2] = (float)((double)rec.box[2] * sx);
    bx[3] = (float)((double)rec.box[3] * sy);
}

// ------------------------------------------------------------------------------------------------
// Records of a batch, compacted: what leaves the GPU (gather to rank 0, device->host copy).  The dense layout
// [n_img, max_cc] is >95 % empty slots (1-3 components per image, max_cc = 256); here the headers are followed by the
// live records only, in image order.  One CTA: exclusive scan of n_rec over the images, then one warp per image copies
// its records (96 B = 6 x 16 B per lane).
//   packed = [n_alloc headers (64 B each; hdr.reserved = index of the image's first record)]
//            [psam_packed_tail: total records, capacity, flags]  [capacity records]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_compact_records(const psam_image_hdr* __restrict__ hdr,
                                                          const psam_prompt_rec* __restrict__ recs, int n_img, int n_alloc,
                                                          int max_cc, int capacity, uint8_t* __restrict__ packed)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    psam_image_hdr* ohdr = reinterpret_cast<psam_image_hdr*>(packed);
    psam_packed_tail* tail = reinterpret_cast<psam_packed_tail*>(packed + (size_t)n_alloc * sizeof(psam_image_hdr));
    psam_prompt_rec* orec = reinterpret_cast<psam_prompt_rec*>(tail + 1);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    TraceRec* tr = tid == 0 ? trace_begin(9) : nullptr;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n_alloc; i0 += 1024) {
        const int i = i0 + tid;
        psam_image_hdr h;
        if (i < n_img) h = hdr[i];
        else memset(&h, 0, sizeof(h));
        const int n = i < n_img ? h.n_rec : 0;
        const int ex = warp_excl_scan_i(n, lane);
        if (lane == 31) s_warp[wid] = ex + n;
        __syncthreads();
        int woff = 0, tot = 0;
        for (int k = 0; k < 32; ++k) {
            const int v = s_warp[k];
            if (k < wid) woff += v;
            tot += v;
        }
        const int first = s_base + woff + ex;
        if (i < n_alloc) {
            h.reserved = first;
            ohdr[i] = h;
        }
        // images with a few records (the usual 1-3): their thread copies them, all loads of the CTA in flight at once;
        // images with many: one warp per image, 96 B = 6 x 16 B per lane
        if (n > 0 && n <= 8) {
            const uint4* src = reinterpret_cast<const uint4*>(recs + (size_t)i * max_cc);
            uint4* dst = reinterpret_cast<uint4*>(orec + first);
            const int nk = min(n, max(capacity - first, 0)) * 6;
            for (int k = 0; k < nk; ++k) dst[k] = src[k];
        }
        for (int j = 0; j < 32; ++j) {
            const int nj = __shfl_sync(0xffffffffu, n, j), fj = __shfl_sync(0xffffffffu, first, j);
            if (nj <= 8) continue;
            const int img = i0 + wid * 32 + j;
            const uint4* src = reinterpret_cast<const uint4*>(recs + (size_t)img * max_cc);
            uint4* dst = reinterpret_cast<uint4*>(orec + fj);
            for (int k = lane; k < nj * 6; k += 32)
                if (fj + k / 6 < capacity) dst[k] = src[k];
        }
        __syncthreads();
        if (tid == 0) s_base += tot;
        __syncthreads();
    }
    if (tid == 0) {
        tail->total = s_base;
        tail->capacity = capacity;
        tail->flags = s_base > capacity ? PSAM_PACKED_OVERFLOW : 0;
        tail->n_img = n_img;
        for (int k = 0; k < 12; ++k) tail->reserved[k] = 0;
    }
    trace_end(tr);
}

}  // namespace psam

using namespace psam;

PSAM_TRACE_TU();
extern "C" size_t psam_packed_bytes(int n_alloc, int capacity)
{
    if (n_alloc <= 0 || capacity < 0) return 0;
    return (size_t)n_alloc * sizeof(psam_image_hdr) + sizeof(psam_packed_tail) + (size_t)capacity * sizeof(psam_prompt_rec);
}

extern "C" int psam_compact_records(const psam_image_hdr* hdr, const psam_prompt_rec* recs, int n_img, int n_alloc, int max_cc,
                                    int capacity, void* packed, psam_stream_t stream_)
{
    PSAM_TRACE("psam_compact_records");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(hdr && recs && packed, "psam_compact_records: null pointer");
    PSAM_CHECK_ARG(n_img >= 1 && n_alloc >= n_img && max_cc >= 1 && capacity >= 1, "psam_compact_records: bad shape");
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_compact_records);
    k_compact_records<<<1, 1024, 0, stream>>>(hdr, recs, n_img, n_alloc, max_cc, capacity, static_cast<uint8_t*>(packed));
    PSAM_CHECK_LAUNCH("k_compact_records");
    return PSAM_OK;
}

static int comp_grid(int n_img)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return n_img < 2 * sms ? n_img : 2 * sms;      // up to two 512-thread CTAs per SM
}

static size_t comp_scratch_bytes(int ctas, int max_runs, int max_cc)
{
    // arrays are carved per kind (all CTAs contiguous), see psam_components
    size_t b = 0;
    b += align_up(sizeof(uint16_t) * (size_t)ctas * max_runs, 256) * 3;
    b += align_up(sizeof(int32_t) * (size_t)ctas * max_runs, 256) * 2;
    b += align_up(sizeof(uint32_t) * (size_t)ctas * max_runs, 256);
    b += align_up(sizeof(unsigned long long) * (size_t)ctas * max_runs, 256);
    b += align_up(sizeof(Acc) * (size_t)ctas * max_cc, 256);
    b += align_up(sizeof(uint32_t) * (size_t)ctas * 2 * KEY_WORDS, 256);
    return b + 256;
}

extern "C" size_t psam_prompts_workspace(int n_img, int out, int max_runs, int max_cc)
{
    (void)out;
    if (n_img <= 0 || max_runs <= 0 || max_cc <= 0) return 0;
    return comp_scratch_bytes(comp_grid(n_img), max_runs, max_cc);
}

extern "C" int psam_components(const uint32_t* maskbits, const float* p_fg, const uint64_t* wstat, int n_img, int out, int use_cca,
                               int max_cc, int max_runs, psam_image_hdr* hdr, psam_prompt_rec* recs,
                               int32_t* labels_out, void* workspace, size_t workspace_bytes, psam_stream_t stream_)
{
    PSAM_TRACE("psam_components");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(maskbits && p_fg && hdr && recs && workspace, "psam_components: null pointer");
    PSAM_CHECK_ARG(n_img >= 1, "psam_components: n_img %d", n_img);
    PSAM_CHECK_ARG(out >= 1 && out <= MAX_OUT, "psam_components: out=%d must be in [1,%d]", out, MAX_OUT);
    PSAM_CHECK_ARG(max_cc >= 1 && max_cc <= (1 << 18), "psam_components: max_cc %d", max_cc);
    PSAM_CHECK_ARG(max_runs >= 1 && max_runs <= (1 << 24), "psam_components: max_runs %d", max_runs);
    const int ctas = comp_grid(n_img);
    if (workspace_bytes < comp_scratch_bytes(ctas, max_runs, max_cc)) {
        set_error("psam_components: workspace too small (%zu < %zu)", workspace_bytes, comp_scratch_bytes(ctas, max_runs, max_cc));
        return PSAM_ERR_WORKSPACE;
    }
    CompParams P;
    P.maskbits = maskbits; P.p_fg = p_fg; P.wstat = reinterpret_cast<const uint2*>(wstat); P.n_img = n_img; P.out = out; P.use_cca = use_cca ? 1 : 0;
    P.max_cc = max_cc; P.max_runs = max_runs; P.hdr = hdr; P.recs = recs; P.labels_out = labels_out;
    Carver cv(workspace);
    const size_t n = (size_t)ctas * max_runs;
    P.run_s = cv.take<uint16_t>(n);
    P.run_e = cv.take<uint16_t>(n);
    P.run_y = cv.take<uint16_t>(n);
    P.parent = cv.take<int32_t>(n);
    P.rank = cv.take<int32_t>(n);
    P.minkey = cv.take<uint32_t>(n);
    P.sump = cv.take<unsigned long long>(n);
    P.acc = cv.take<Acc>((size_t)ctas * max_cc);
    P.gkeys = cv.take<uint32_t>((size_t)ctas * 2 * KEY_WORDS);
    if (labels_out) {
        cudaError_t e = cudaMemsetAsync(labels_out, 0, sizeof(int32_t) * (size_t)n_img * out * out, stream);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return PSAM_ERR_LAUNCH; }
    }
    const int dyn = 0;
    static const bool skip = getenv("PSAM_EXPERIMENT_SKIP_COMPONENTS") != nullptr;   // timing experiments only: no records
    if (skip) return PSAM_OK;
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_components);
    k_components<<<ctas, CT, dyn, stream>>>(P);
    PSAM_CHECK_LAUNCH("k_components");
    return PSAM_OK;
}

extern "C" size_t psam_coarse_to_prompts_workspace(int n_img, int out, int max_runs, int max_cc)
{
    if (n_img <= 0 || out <= 0) return 0;
    size_t b = 0;
    b += align_up(sizeof(float) * (size_t)n_img * out * out, 256);            // p_fg
    b += align_up(sizeof(uint32_t) * (size_t)n_img * out * ((out + 31) / 32), 256);  // mask bits
    b += align_up(sizeof(uint2) * (size_t)n_img * out * ((out + 31) / 32), 256);     // per-word statistics
    b += align_up(psam_upsample_workspace(n_img, out), 256);                  // block work list
    b += psam_prompts_workspace(n_img, out, max_runs, max_cc);
    return b + 256;
}

extern "C" int psam_coarse_to_prompts(const float* logits, int n_img, int h, int w, int mid, int out, int use_cca,
                                      int prob_mode, int max_cc, int max_runs, psam_image_hdr* hdr, psam_prompt_rec* recs,
                                      void* workspace, size_t workspace_bytes, psam_stream_t stream)
{
    PSAM_TRACE("psam_coarse_to_prompts");
    PSAM_CHECK_ARG(workspace, "psam_coarse_to_prompts: null workspace");
    if (workspace_bytes < psam_coarse_to_prompts_workspace(n_img, out, max_runs, max_cc)) {
        set_error("psam_coarse_to_prompts: workspace too small");
        return PSAM_ERR_WORKSPACE;
    }
    Carver cv(workspace);
    float* p_fg = cv.take<float>((size_t)n_img * out * out);
    uint32_t* bits = cv.take<uint32_t>((size_t)n_img * out * ((out + 31) / 32));
    uint64_t* wstat = cv.take<uint64_t>((size_t)n_img * out * ((out + 31) / 32));
    const size_t upw = psam_upsample_workspace(n_img, out);
    char* upws = cv.take<char>(upw);
    char* rest = static_cast<char*>(workspace) + cv.used();
    int rc = psam_upsample_softmax(logits, n_img, h, w, mid, out, p_fg, bits, nullptr, wstat, 1, prob_mode, upws, upw, stream);
    if (rc) return rc;
    return psam_components(bits, p_fg, wstat, n_img, out, use_cca, max_cc, max_runs, hdr, recs, nullptr, rest,
                           workspace_bytes - cv.used(), stream);
}

extern "C" int psam_records_to_sam(const psam_image_hdr* hdr, const psam_prompt_rec* recs, int n_img, int max_cc,
                                   int point_mode, int old_h, int old_w, int target_length, float* points,
                                   int32_t* labels, float* boxes, psam_stream_t stream_)
{
    PSAM_TRACE("psam_records_to_sam");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PSAM_CHECK_ARG(hdr && recs && points && labels && boxes, "psam_records_to_sam: null pointer");
    PSAM_CHECK_ARG(n_img >= 1 && max_cc >= 1 && point_mode >= 0 && point_mode <= 2 && old_h >= 1 && old_w >= 1 &&
                       target_length >= 1, "psam_records_to_sam: bad argument");
    // ResizeLongestSide.get_preprocess_shape (models/segment_anything/utils/transforms.py:140-148)
    const double scale = (double)target_length * 1.0 / (double)(old_h > old_w ? old_h : old_w);
    const int new_h = (int)((double)old_h * scale + 0.5), new_w = (int)((double)old_w * scale + 0.5);
    const double sx = (double)new_w / (double)old_w, sy = (double)new_h / (double)old_h;
    PSAM_PROF_BEGIN(stream);
    PSAM_MAX_CARVEOUT(k_records_to_sam);
    k_records_to_sam<<<(n_img * max_cc + 255) / 256, 256, 0, stream>>>(hdr, recs, n_img, max_cc, point_mode, sx, sy, points,
                                                                       labels, boxes);
    PSAM_CHECK_LAUNCH("k_records_to_sam");
    return PSAM_OK;
}

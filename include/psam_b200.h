/*
 * psam_b200.h -- C ABI of libpsam_b200.so: the ProtoSAM coarse-segmentation hot path
 * (ALP prototype pooling, cosine matching, coarse-map -> SAM prompts) as sm_100a CUDA.
 *
 * The reference (levayz/ProtoSAM) is pure Python: it has no FFI, plugin table or
 * operator registry for this path.  The "drop-in boundary" is the
 * nn.Module / helper-function surface listed below; each entry point here names the
 * reference interface it replaces (file:line relative to the reference tree) and is what
 * a ctypes binding in the reference would call (INTEGRATION.md shows that stub).
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller unless marked "host".
 *     The library allocates nothing; scratch comes in through `workspace`.
 *   - All work is enqueued on `stream` (a cudaStream_t); no entry point synchronises the
 *     host.  Nothing is cached between calls; calls on different streams are independent.
 *   - Return value: 0 = enqueued; negative = argument/launch error (see psam_last_error()).
 *     Data-dependent conditions (empty prototype set, table overflow) are reported in
 *     device-side status words so that no host sync is needed to launch the next stage.
 *   - fp32 throughout; integer outputs are bit-exact w.r.t. the reference's CPU code,
 *     maps are within 1e-3 absolute (BASELINE.json north_star).
 */
#ifndef PSAM_B200_H
#define PSAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSAM_ABI_VERSION 3

typedef void* psam_stream_t; /* cudaStream_t */

/* Prototype modes of MultiProtoAsConv (models/alpmodule.py:57-94, 97-159). */
#define PSAM_MODE_MASK          0 /* 'mask'      : one global masked-average prototype per shot  */
#define PSAM_MODE_GRIDCONV      1 /* 'gridconv'  : local grid prototypes only                   */
#define PSAM_MODE_GRIDCONV_PLUS 2 /* 'gridconv+' : local grid prototypes + the global prototype */
/* Caller-side decision of FewShotSeg.forward (models/grid_proto_fewshot.py:250-256), taken on
 * the device: 'gridconv+' if max(avg_pool2d(mask, auto_k)) >= thresh else 'mask'. */
#define PSAM_MODE_AUTO_FG       3

/* Bits of the per-set status word written by psam_alp_prototypes / psam_alp_match. */
#define PSAM_SET_EMPTY 1 /* grid mode with zero prototypes: the reference raises inside F.conv2d
                            (models/alpmodule.py:68); scores are NaN */

/* Error codes. */
#define PSAM_OK              0
#define PSAM_ERR_ARG        -1
#define PSAM_ERR_WORKSPACE  -2
#define PSAM_ERR_LAUNCH     -3
#define PSAM_ERR_UNSUPPORTED -4

int psam_abi_version(void);
/* Message of the last error on the calling thread ("" if none). */
const char* psam_last_error(void);
/* Number of kernels this library has launched on the calling process so far (bench.py's
 * gpu_launches claim is read from here). */
uint64_t psam_launch_count(void);
/* Optional per-kernel device timing for bench.py's roofline: while enabled, every kernel launch of the
 * library is bracketed by a CUDA event pair on its stream.  psam_profile_collect synchronises on those
 * events, returns the number of distinct kernels seen since the last collect and fills, in first-launch
 * order, their newline-separated names, accumulated milliseconds and launch counts. */
/* Optional CTA-level trace for timeline analysis (tools/trace_timeline.py): install a zero-initialised device buffer and
 * the big kernels record, per CTA, {start ns, end ns, SM id, kernel id, CTA index} (32-byte records; record 0 is the
 * header: t0 = records taken, t1 = capacity).  NULL uninstalls.  Needs a library built with `make TRACE=1`
 * (PSAM_ERR_UNSUPPORTED otherwise).  Synchronises the device; not for use inside graphs
 * capture.  Kernel ids: 1 k_match_tc, 2 k_pack_query, 3 k_blocks_warp, 4 k_components, 5 k_classify_blocks,
 * 6 k_proto_stage1, 7 k_proto_stage2, 8 k_pack_protos, 9 k_compact_records. */
int psam_trace_install(void* device_buffer, size_t bytes);
void psam_profile_enable(int on);
int psam_profile_collect(char* names, size_t names_bytes, float* total_ms, int32_t* launches, int max_kernels);

/* --------------------------------------------------------------------------------------
 * Kernel 1 -- prototypes.  Replaces MultiProtoAsConv.get_prototypes + safe_norm
 * (models/alpmodule.py:14-18, 97-159) for `nsets` prototype sets that share one support
 * feature tensor (e.g. background + foreground set of every label of a volume).
 *
 *   sup_x         [S,C,h,w] logical, addressed through `sup_x_strides` (host array of 4
 *                 element strides: shot, channel, row, col) -- the reference hands over a
 *                 physically channels-last tensor (stride_c == 1), which is the fast case.
 *   sup_y         [nsets,S,h,w] contiguous masks (float, usually {0,1}).
 *   set_modes     host [nsets] PSAM_MODE_*.
 *   kh,kw         pooling window = stride (val_wsize when isval, else module kernel_size).
 *   auto_kh/kw    window of the AUTO_FG decision (module kernel_size); ignored otherwise.
 *   thresh        survivors are windows with mean(mask) > thresh (:131,153).
 * Outputs (N = S*(h/kh)*(w/kw); cap_rows = N + S):
 *   protos        [nsets,cap_rows,C] set i holds counts[i] L2-normalised rows (norm clamped
 *                 at 1e-4): surviving windows in (shot,gy,gx) order, then S global rows for
 *                 'gridconv+'; 'mask' sets hold their S global rows only.
 *   counts        [nsets] int32 rows in use;  eff_modes [nsets] resolved mode (AUTO_FG).
 *   status        [nsets] int32 PSAM_SET_* bits.
 *   survive       [nsets,N] uint8 (bit-exact gate);  pooled [nsets,N] window mask fraction.
 * ------------------------------------------------------------------------------------ */
size_t psam_alp_prototypes_workspace(int nsets, int S, int C, int h, int w, int kh, int kw);

int psam_alp_prototypes(const float* sup_x, const int64_t* sup_x_strides, const float* sup_y,
                        int nsets, const int32_t* set_modes, int S, int C, int h, int w,
                        int kh, int kw, int auto_kh, int auto_kw, float thresh,
                        float* protos, int32_t* counts, int32_t* eff_modes, int32_t* status,
                        uint8_t* survive, float* pooled,
                        void* workspace, size_t workspace_bytes, psam_stream_t stream);

/* The same with a shot selector per set (host [nsets]: -1 = every shot, k = shot k only).  FewShotSeg.forward
 * matches the foreground once per shot (models/grid_proto_fewshot.py:244-262): a set restricted to shot k takes its
 * local windows, its AUTO_FG decision and its single global row from that shot alone, exactly like calling the
 * module with supp_fts[:, k] and that shot's mask. */
int psam_alp_prototypes_shots(const float* sup_x, const int64_t* sup_x_strides, const float* sup_y,
                              int nsets, const int32_t* set_modes, const int32_t* set_shots, int S, int C, int h, int w,
                              int kh, int kw, int auto_kh, int auto_kw, float thresh,
                              float* protos, int32_t* counts, int32_t* eff_modes, int32_t* status,
                              uint8_t* survive, float* pooled,
                              void* workspace, size_t workspace_bytes, psam_stream_t stream);

/* Element-wise max over the per-shot foreground scores (grid_proto_fewshot.py:263-270).  scores [Q, L*(1+S), HW] with
 * the sets of label l ordered (bg_l, fg_l shot 0, ..., fg_l shot S-1) -> logits [Q*L, 2, HW] = (bg_l, max_s fg_l,s). */
int psam_combine_shots(const float* scores, int Q, int L, int S, int HW, float* logits, psam_stream_t stream);

/* DINOv2 token hand-off of FewShotSeg.get_features (models/grid_proto_fewshot.py:90-98): tokens [B, h*w, C]
 * (= the channels-last feature map) -> out [B, oh, ow, C], bilinear, align_corners=False; the reference does this when
 * the map has fewer than 32x32 tokens (BASELINE config 1: 18x18 -> 32x32).  Upsampling only. */
int psam_tokens_to_features(const float* tokens, int B, int h, int w, int C, int oh, int ow, float* out,
                            psam_stream_t stream);

/* Nearest-neighbour resize of the support masks to feature resolution, as FewShotSeg.forward does before calling
 * the ALP module (F.interpolate(mask, fts_size, mode='nearest'), models/grid_proto_fewshot.py:228-231):
 * src [n,H,W] -> dst [n,h,w], source index = min(floor(dst * (float)in / out), in - 1). */
int psam_mask_nearest(const float* src, int n, int H, int W, int h, int w, float* dst, psam_stream_t stream);

/* Viz grid returned as 4th output of MultiProtoAsConv.forward for the grid modes
 * (`resized_proto_grid`, models/alpmodule.py:120-128, 142-150) from the `pooled` values of one
 * set: out [gh*vw, gw*vw] float32 (the reference builds it on the CPU in a Python loop). */
int psam_alp_proto_grid(const float* pooled, int S, int gh, int gw, int vw, float thresh,
                        int mode, float* out, psam_stream_t stream);

/* --------------------------------------------------------------------------------------
 * Kernel 2 -- fused match.  Replaces safe_norm(qry) + get_prediction_from_prototypes
 * (models/alpmodule.py:57-94, 195) for Q query slices against all `nsets` sets at once.
 *
 *   qry        [Q,HW,C] channels-last rows: element (q,p,c) at qry[q*slice_stride + p*row_stride + c].
 *   protos/counts/eff_modes   as written by psam_alp_prototypes (cap_rows rows per set).
 *   scores     [Q,nsets,HW]  grid modes: sum_p softmax_p(d) * d with d = 20*cos;  mask: max_s d.
 *              With sets ordered (bg_0, fg_0, bg_1, fg_1, ...) this buffer *is* the
 *              [Q*L,2,h,w] logits tensor FewShotSeg concatenates (grid_proto_fewshot.py:270).
 *   assign     [Q,nsets,HW] or NULL: argmax_p d as float (grid modes) / the score (mask mode).
 *   sims       [Q,nsets,cap_rows,HW] or NULL: raw d ('raw_local_sims', vis_sim=True).
 *   status     [nsets] int32, PSAM_SET_EMPTY is OR-ed in for empty grid sets.
 *   algo       0 = auto, 1 = fp32 CUDA-core kernel, 2 = tcgen05 split-bf16 tensor-core kernel fed from packed
 *              operand images (a pack kernel writes the query's bf16 hi/lo image first), 3 = the fused tensor-core
 *              kernel: the fp32 query is read through a 2-D tensor map, split into bf16 hi/lo inside the GEMM and used
 *              as the A operand from TMEM (needs dense slices: slice_stride == HW * row_stride, or Q == 1).
 *              auto = 3 for dense slices and tables of < 4096 columns capacity (nsets * pad16(cap_rows)), 2 for wider
 *              tables / strided slices, 1 when sims != NULL, C % 8 != 0 or the workspace is too small.
 * ------------------------------------------------------------------------------------ */
size_t psam_alp_match_workspace(int Q, int HW, int C, int nsets, int cap_rows, int algo);

/* The tensor-core match kernels are persistent: one CTA per SM for the whole launch.  psam_match_reserve_sms(n) makes them
 * leave n SMs without such a CTA (process-wide; n < 0 only queries; returns the previous value; default 0, or
 * PSAM_TC_RESERVE_SMS).  For multi-GPU runs with several volumes in flight: the NCCL kernels of the other volumes'
 * prototype broadcast / record gather (no counterpart in the single-GPU reference) then find an
 * SM whose CTAs are all short-lived instead of waiting for a GEMM CTA to retire (2 GPUs, config 2: 0.346 -> 0.328 ms per
 * volume with n = 8). */
int psam_match_reserve_sms(int n);

int psam_alp_match(const float* qry, int64_t slice_stride, int64_t row_stride, int Q, int HW, int C,
                   const float* protos, int cap_rows, const int32_t* counts, const int32_t* eff_modes,
                   int nsets, float* scores, float* assign, float* sims, int32_t* status,
                   void* workspace, size_t workspace_bytes, int algo, psam_stream_t stream);

/* --------------------------------------------------------------------------------------
 * Kernel 3 -- coarse map -> prompts.  Replaces, per image (= one query slice x one label):
 *   F.interpolate(pred, img_size, 'bilinear')              models/grid_proto_fewshot.py:270-273
 *   F.interpolate(logits, 1024, 'bilinear'), softmax, argmax    models/ProtoSAM.py:592-602
 *   get_connected_components / cca                         util/utils.py:474-541
 *   get_bbox_per_cc, get_most_conf_points, centroids       models/ProtoSAM.py:242-289, 349-450
 * ------------------------------------------------------------------------------------ */

/* One connected component, in OpenCV label order.  96 bytes. */
typedef struct psam_prompt_rec {
    int64_t box[4];      /* min_x, min_y, max_x, max_y (inclusive)   ProtoSAM.py:242-264      */
    int64_t conf_pt[2];  /* x, y of torch.topk(p_fg[mask], 1)        ProtoSAM.py:266-289      */
    double  centroid[2]; /* x, y = integer sums / area in double     cv2 centroids            */
    double  conf;        /* sum(p_fg over component) / (n_fg + 1e-6) util/utils.py:490        */
    float   conf_pt_p;   /* p_fg at conf_pt                                                   */
    int32_t area;
    int32_t label;       /* label cv2.connectedComponentsWithStats gives this component       */
    int32_t flags;       /* PSAM_REC_* */
    int32_t reserved[2];
} psam_prompt_rec;

#define PSAM_REC_SELECTED 1 /* the component `cca` keeps (use_cca) */

/* Per-image header.  64 bytes. */
typedef struct psam_image_hdr {
    int32_t ncc;         /* foreground components found (cv2 count - 1)                        */
    int32_t n_rec;       /* records written: min(ncc, max_cc), or 0/1 with use_cca             */
    int32_t n_fg;        /* foreground pixels (= _pred.sum())                                  */
    int32_t flags;       /* PSAM_IMG_*                                                         */
    int32_t bg_stats[5]; /* cv2 stats row 0: left, top, width, height, area of the background  */
    int32_t n_runs;      /* foreground runs found (diagnostic)                                 */
    double  bg_centroid[2];
    int32_t selected;    /* use_cca: cv2 label of the kept component, 0 if none                */
    int32_t reserved;
} psam_image_hdr;

#define PSAM_IMG_EMPTY          1 /* no foreground pixel: ProtoSAM.forward returns early (:612-613) */
#define PSAM_IMG_RUN_OVERFLOW   2 /* more runs than the workspace was sized for: nothing emitted     */
#define PSAM_IMG_CC_TRUNCATED   4 /* ncc > max_cc: only the first max_cc labels were emitted         */
#define PSAM_IMG_CCA_AMBIGUOUS  8 /* use_cca: two components' confidences are closer than fp32 sum
                                     rounding; the exact-sum winner was kept                         */

/* logits [n_img,2,h,w] -> two-stage bilinear (h,w)->(mid,mid)->(out,out) (single stage when
 * mid == out), 2-way softmax, foreground bit.  ATen-CPU operation order, so bit-exact with
 * the reference's CPU run.  Outputs, each optional (NULL to skip) except maskbits:
 *   p_fg     [n_img,out,out] float  probability of class 1 (written at every pixel)
 *   maskbits [n_img,out,out/32] uint32, bit (x&31) of word (y, x>>5) = argmax == 1
 *   probs2   [n_img,2,out,out] float  full softmax (`output_p`), for function-level parity
 *   wstat    [n_img,out,out/32] uint64 per-32-pixel-word statistics consumed by psam_components:
 *            low word = sum over the word's foreground pixels of p_fg * 2^24 (exact integer),
 *            high word = max of (p_fg * 2^24) << 5 | (31 - lane): best probability, leftmost pixel
 * fg_only != 0 (engine path, needs p_fg, wstat and a workspace): 32x32-pixel blocks whose result is
 * known exactly from the low-resolution cells they depend on (all background / all saturated) are not
 * evaluated, p_fg is written at foreground pixels only, and exp/division run only where class 1 can
 * win (probs2 must be NULL).  The mask, p_fg at foreground pixels and wstat are identical to fg_only=0. */
size_t psam_upsample_workspace(int n_img, int out);

/* prob_mode: which probability p_fg / wstat (and through them the component confidences and the most confident
 * points) carry.  The mask is argmax of the 2-way softmax in both modes. */
#define PSAM_PROB_SOFTMAX        0 /* ProtoSAM: softmax(logits)[1]                        models/ProtoSAM.py:599        */
#define PSAM_PROB_SOFTMAX_TWICE  1 /* ProtoMedSAM: need_softmax turns the logits into probabilities first and
                                      cca / get_connected_components apply softmax(1) to them AGAIN
                                      (models/ProtoMedSAM.py:178-187, util/utils.py:62-63, 486): softmax(softmax(logits))[1] */

int psam_upsample_softmax(const float* logits, int n_img, int h, int w, int mid, int out,
                          float* p_fg, uint32_t* maskbits, float* probs2, uint64_t* wstat, int fg_only, int prob_mode,
                          void* workspace, size_t workspace_bytes, psam_stream_t stream);

/* max_runs: capacity of the per-CTA run table (foreground row segments per image). */
size_t psam_prompts_workspace(int n_img, int out, int max_runs, int max_cc);

/* maskbits + p_fg (+ optional wstat, NULL = derive everything from p_fg) -> headers and records.  use_cca != 0 keeps only the most confident
 * component (util/utils.py:496-541).  labels_out (optional) [n_img,out,out] int32 receives
 * the cv2-numbered label image (0/1 image of the kept component with use_cca). */
int psam_components(const uint32_t* maskbits, const float* p_fg, const uint64_t* wstat, int n_img, int out,
                    int use_cca, int max_cc, int max_runs,
                    psam_image_hdr* hdr, psam_prompt_rec* recs, int32_t* labels_out,
                    void* workspace, size_t workspace_bytes, psam_stream_t stream);

/* Records -> the prompt tensors of SamPredictor.predict_torch (models/segment_anything/predictor.py:136-167), in
 * SAM's input frame: ResizeLongestSide.apply_coords / apply_boxes (models/segment_anything/utils/transforms.py:40-62,
 * 140-148) scale x by new_w/old_w and y by new_h/old_h in double precision, torch.as_tensor(dtype=float) rounds to
 * fp32.  point_mode 0 = 'conf', 1 = 'centroid', 2 = 'both' (npts = 1, 1, 2; models/ProtoSAM.py:349-450).
 *   points [n_img,max_cc,npts,2] float, labels [n_img,max_cc,npts] int32 (1 = foreground), boxes [n_img,max_cc,4]
 *   float XYXY; slots beyond the image's n_rec are zero. */
int psam_records_to_sam(const psam_image_hdr* hdr, const psam_prompt_rec* recs, int n_img, int max_cc,
                        int point_mode, int old_h, int old_w, int target_length,
                        float* points, int32_t* labels, float* boxes, psam_stream_t stream);

/* Compact form of a batch's headers + records: what leaves the GPU (device->host copy, gather to rank 0).  The dense
 * [n_img, max_cc] record array is almost all empty slots; `packed` holds
 *   n_alloc headers (image i's hdr.reserved = index of its first record; images >= n_img zeroed)
 *   one psam_packed_tail
 *   `capacity` records, the live ones of all images in image order.
 * psam_packed_bytes gives the size.  More than `capacity` live records set PSAM_PACKED_OVERFLOW (the surplus is
 * dropped; the dense arrays are untouched, so the caller can fall back to them). */
typedef struct psam_packed_tail {
    int32_t total;       /* live records of the batch                    */
    int32_t capacity;
    int32_t flags;       /* PSAM_PACKED_* */
    int32_t n_img;
    int32_t reserved[12];
} psam_packed_tail;     /* 64 bytes */

#define PSAM_PACKED_OVERFLOW 1

size_t psam_packed_bytes(int n_alloc, int capacity);

int psam_compact_records(const psam_image_hdr* hdr, const psam_prompt_rec* recs, int n_img, int n_alloc, int max_cc,
                         int capacity, void* packed, psam_stream_t stream);

/* Both stages for a batch of images: what the volume engine calls.  p_fg / maskbits live in
 * the workspace. */
size_t psam_coarse_to_prompts_workspace(int n_img, int out, int max_runs, int max_cc);

int psam_coarse_to_prompts(const float* logits, int n_img, int h, int w, int mid, int out,
                           int use_cca, int prob_mode, int max_cc, int max_runs,
                           psam_image_hdr* hdr, psam_prompt_rec* recs,
                           void* workspace, size_t workspace_bytes, psam_stream_t stream);

/* --------------------------------------------------------------------------------------
 * Multi-GPU exchanges over NVLink peer memory (the reference is single-GPU; north_star shards the
 * query slices of a volume over the GPUs of one box, sends the support prototypes to every rank and brings the prompt
 * records back).  One-sided: a sender stores straight into the receiver's memory and raises a signal word there; no
 * collective library on the path.  The caller provides, per channel, one REGION per rank -- a symmetric allocation that
 * every peer has mapped (e.g. torch.distributed._symmetric_memory.empty + rendezvous; `regions` is the device array of
 * the world's region addresses as THIS rank sees them, its `buffer_ptrs_dev`) -- zero-initialised before first use, of
 * psam_peer_region_bytes(payload) bytes: PSAM_PEER_SIGNAL_BYTES of signal words, then the payload ("mailbox").  `ctr` is
 * 16 zero-initialised bytes of local device memory: one for a table channel (push and recv share it), two for a record
 * channel (put, collect); epochs live there, so the calls can be captured into CUDA graphs.  Source and destination
 * ranks may change from one exchange to the next.  Every rank must issue the matching calls in the same order; a peer that never answers
 * traps the waiting kernel after 120 s instead of hanging the GPU.
 *
 *   psam_peer_push_table  (source rank) waits until every peer acknowledged the previous table, writes the LIVE rows
 *                         (counts[s] of cap_rows per set) + the integer arrays of its prototype table into every peer's
 *                         mailbox, signals.  table = the packed buffer of psam_alp_prototypes' outputs: rows at 0, the
 *                         integer arrays (counts first) at ints_offset.
 *   psam_peer_recv_table  (other ranks) waits for the signal, copies mailbox -> its private table, acknowledges.
 *   psam_peer_put         (every rank) waits for dst's acknowledgement of the previous round, writes nbytes into slot
 *                         `rank` (slot_bytes apart) of dst's mailbox, signals.
 *   psam_peer_collect     (dst, after its own psam_peer_put on the same stream; put_ctr = that call's ctr) waits for every
 *                         rank's signal, copies the world slots -> out, acknowledges to all.
 * ------------------------------------------------------------------------------------ */
#define PSAM_PEER_SIGNAL_BYTES 4096

size_t psam_peer_region_bytes(size_t payload_bytes);

int psam_peer_push_table(const void* table, int nsets, int cap_rows, int C, size_t ints_offset, size_t ints_bytes,
                         void* const* regions, int world, int rank, void* ctr, psam_stream_t stream);

int psam_peer_recv_table(void* table, int nsets, int cap_rows, int C, size_t ints_offset, size_t ints_bytes,
                         void* const* regions, int world, int rank, int src, void* ctr, psam_stream_t stream);

int psam_peer_put(const void* src, size_t nbytes, size_t slot_bytes, void* const* regions, int world, int rank, int dst,
                  void* ctr, psam_stream_t stream);

int psam_peer_collect(void* out, size_t slot_bytes, void* const* regions, int world, int rank, const void* put_ctr, void* ctr,
                      psam_stream_t stream);

/* --------------------------------------------------------------------------------------
 * Optional prompt variants of ProtoSAM.forward (off in the reference's configs, config_ssl_upload.py:93,102).  They
 * consume what psam_upsample_softmax (fg_only = 0, probs2) and psam_components (labels_out) leave on the device.
 * ------------------------------------------------------------------------------------ */

/* One negative-point candidate.  32 bytes. */
typedef struct psam_neg_point {
    int64_t pt[2];      /* x, y; -1 when there is none                                              */
    float   p;          /* the (possibly thresholded) background probability torch.topk picked      */
    int32_t has;        /* 0: the search mask was empty                                             */
    int32_t n;          /* pixels in the search mask, saturating once >= 64 were seen               */
    int32_t reserved;
} psam_neg_point;

/* get_sam_input_points(..., get_neg_points=True, l=1) (models/ProtoSAM.py:361-434).  For image i, neg[i][r] (r < n_rec)
 * is the most confident background pixel in the ring cv2.dilate(component r, 3x3, iterations=ring_width) minus the
 * component, and neg[i][max_cc] the most confident pixel of the whole image with p_bg >= thresh; "most confident" is
 * torch.topk(values[mask], 1) with its tie rules.  The caller stacks [ring point, global point] per component.
 *   labels    [n_img,out,out] as written by psam_components (0/1 image of the kept component with use_cca)
 *   p_bg      background probabilities, image i at p_bg + i * p_bg_image_stride (= channel 0 of probs2)
 *   host_aliasing != 0: the ring search reads the map thresholded in place (p_bg < thresh -> 0), which is what the
 *   reference computes when its tensors live on the CPU (.cpu() returns a view, models/ProtoSAM.py:363-364 vs :414);
 *   0 = the raw map, as on the reference's CUDA path.  out <= 1024. */
int psam_neg_points(const int32_t* labels, const float* p_bg, int64_t p_bg_image_stride, const psam_image_hdr* hdr,
                    const psam_prompt_rec* recs, int n_img, int out, int max_cc, int use_cca, int ring_width, float thresh,
                    int host_aliasing, psam_neg_point* neg /* [n_img, max_cc + 1] */, psam_stream_t stream);

/* get_most_conf_points(output_p_fg, pred, k) for any k (models/ProtoSAM.py:266-289; production uses k = 1, which the
 * records already carry): for component r of image i, pts[i][r][j] = (x, y) and conf[i][r][j] = p_fg of the j-th entry of
 * torch.topk(p_fg[component], k) -- ATen's CPU algorithm (partial_sort for k * 64 <= area, nth_element + sort below)
 * replayed, so equal probabilities come out in the reference's order.  Components with fewer than k pixels (torch.topk
 * raises) and slots beyond n_rec get pts = -1.  k <= 64.
 *   p_fg   foreground probabilities at every pixel, image i at p_fg + i * p_fg_image_stride (= channel 1 of probs2) */
size_t psam_topk_points_workspace(int n_img, int max_cc, int k);

int psam_topk_points(const int32_t* labels, const float* p_fg, int64_t p_fg_image_stride, const psam_image_hdr* hdr,
                     const psam_prompt_rec* recs, int n_img, int out, int max_cc, int use_cca, int k,
                     int64_t* pts /* [n_img,max_cc,k,2] */, float* conf /* [n_img,max_cc,k] */,
                     void* workspace, size_t workspace_bytes, psam_stream_t stream);

/* get_sam_input_mask + the mask_input of predict_w_masks (models/ProtoSAM.py:452-476): per component the 0/1 mask
 * resized to size x size (cv2.INTER_NEAREST), foreground 10, background -8 cast to uint8 (= 248).  Masks are packed in
 * image order: masks[offsets[i] + r] belongs to component r of image i; offsets [n_img + 1] is written here; masks
 * beyond `capacity` are dropped (offsets[n_img] tells). */
int psam_mask_prompts(const int32_t* labels, const psam_image_hdr* hdr, const psam_prompt_rec* recs, int n_img, int out,
                      int max_cc, int use_cca, int size, int capacity, uint8_t* masks /* [capacity,size,size] */,
                      int32_t* offsets, psam_stream_t stream);

/* get_confidence_from_logits (util/utils.py:429-434, ProtoSAM coarse_pred_only :580-590) from foreground probabilities
 * p_fg [n_img, pixels_per_image] (psam_upsample_softmax, fg_only = 0): conf[i] = sum(p[p >= 0.5]) / (count + 1e-6). */
int psam_confidence(const float* p_fg, int n_img, int64_t pixels_per_image, double* conf, psam_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PSAM_B200_H */

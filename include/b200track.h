/* b200track - C-ABI of the B200-native multi-stream tracking-by-detection hot path.
 *
 * Drop-in boundary for BoxMOT's per-frame loop (reference: /root/reference, BoxMOT 10.0.51).
 * The reference is pure Python, so there is no reference FFI to mirror; every entry point
 * below names the reference Python interface it replaces (file:line).  The Python binding a
 * maintainer adds is a ctypes stub; yolo_tracking_b200/_lib.py is that stub.
 *
 * Conventions: plain pointers and sizes, no torch types.  Functions return 0 on success and
 * a negative b200track_status otherwise; b200track_last_error() gives the message of the last
 * failure on the calling thread.  Pointers named d_* are DEVICE pointers, h_* are HOST
 * pointers.  `stream` is a cudaStream_t passed as void* (NULL = default stream).  A context
 * is bound to one device and is not thread-safe (like a reference tracker object).
 */
#ifndef B200TRACK_H
#define B200TRACK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200TRACK_ABI_VERSION 3

typedef enum {
    B200TRACK_OK = 0,
    B200TRACK_ERR_ARG = -1,       /* bad argument                                        */
    B200TRACK_ERR_CUDA = -2,      /* CUDA runtime error (no device, launch failure ...)  */
    B200TRACK_ERR_CAPACITY = -3,  /* more tracks / detections than max_tracks / max_dets */
    B200TRACK_ERR_STATE = -4      /* call not valid for this context kind                */
} b200track_status;

typedef enum {
    B200TRACK_BYTETRACK = 0,      /* boxmot/trackers/bytetrack/byte_tracker.py:114 BYTETracker */
    B200TRACK_OCSORT = 1,         /* boxmot/trackers/ocsort/ocsort.py:190 OCSort               */
    B200TRACK_BOTSORT = 2,        /* boxmot/trackers/botsort/bot_sort.py:184 BoTSORT           */
    B200TRACK_DEEPOCSORT = 3,     /* boxmot/trackers/deepocsort/deep_ocsort.py:308 DeepOCSort  */
    B200TRACK_STRONGSORT = 4,     /* boxmot/trackers/strongsort/strong_sort.py:13 StrongSORT   */
    B200TRACK_HYBRIDSORT = 5      /* boxmot/trackers/hybridsort/hybridsort.py:336 HybridSORT   */
} b200track_kind;

typedef enum { B200TRACK_KF_XYAH = 0, B200TRACK_KF_XYWH = 1, B200TRACK_KF_XYAH_CONF = 2 } b200track_kf_kind;
typedef enum {
    B200TRACK_SIM_IOU = 0, B200TRACK_SIM_GIOU = 1, B200TRACK_SIM_DIOU = 2,
    B200TRACK_SIM_CIOU = 3, B200TRACK_SIM_CENTROID = 4
} b200track_sim;

/* Constructor arguments of the reference trackers (tracker_zoo.py:43-81), plus capacities. */
typedef struct {
    int32_t kind;            /* b200track_kind                                              */
    int32_t n_streams;       /* independent video streams tracked by this context           */
    int32_t max_tracks;      /* slots per stream (tracked + lost lists), multiple of 32     */
    int32_t max_dets;        /* detections per stream per frame, multiple of 32             */
    int32_t feat_dim;        /* appearance embedding size (BoTSORT), else 0                 */
    int32_t device;          /* CUDA device ordinal                                         */
    /* ByteTrack (byte_tracker.py:115-130) / BoTSORT (bot_sort.py:185-229) */
    double track_thresh;     /* track_thresh | track_high_thresh                            */
    double track_low_thresh; /* 0.1 (hard-coded in ByteTrack) | track_low_thresh            */
    double new_track_thresh; /* det_thresh = track_thresh | new_track_thresh                */
    double match_thresh;
    double proximity_thresh;
    double appearance_thresh;
    int32_t track_buffer;
    int32_t frame_rate;
    /* OCSORT (ocsort.py:191-216) */
    double det_thresh;
    double iou_thresh;
    double inertia;
    int32_t max_age;
    int32_t min_hits;
    int32_t delta_t;
    int32_t asso_func;       /* b200track_sim                                               */
    int32_t use_byte;
    int32_t with_reid;       /* BoTSORT: use the embedding cost                             */
    int32_t fuse_first_associate; /* BoTSORT: fuse_score on the first association (bot_sort.py:199,300-301); default 0 */
    /* DeepOCSORT (deep_ocsort.py:309-330); it also reads det_thresh, iou_thresh, inertia, max_age, min_hits, delta_t,
     * asso_func and feat_dim above.  embedding_off != 0: no appearance term (feat_dim may be 0). */
    double w_association_emb;
    double alpha_fixed_emb;
    double aw_param;
    int32_t embedding_off;
    int32_t aw_off;
    /* BoTSORT: camera_motion != 0 keeps the filter in the form a camera warp needs (two 4x4 covariance blocks instead of
     * four 2x2) so that per-stream warps can be applied (bot_sort.py:293-295); DeepOCSORT contexts always accept warps */
    int32_t camera_motion;
    int32_t reserved;
    /* StrongSORT (strong_sort.py:14-41; it also reads max_age above): feat_dim a multiple of 64, nn_budget in [1, 128],
     * max_tracks and max_dets <= 256 (the shapes of the tensor-core gallery distance), mc_lambda > 0.  Padded interface
     * only (b200track_step / _step_cam / _step_host / _submit_host); every context accepts warps (Track.camera_update). */
    double max_dist;
    double max_iou_dist;
    double mc_lambda;
    double ema_alpha;
    int32_t n_init;
    int32_t nn_budget;
    /* HybridSORT (hybridsort.py:337-364) reads det_thresh, iou_thresh, inertia, max_age, min_hits, delta_t, asso_func and
     * feat_dim (a multiple of 4) above; everything else the reference's constructor fixes is compiled in.  use_byte must
     * be 0 (tracker_zoo.py:100-115 never forwards it; the branch cannot produce a result row).  Padded interface only;
     * d_feats holds one appearance row per detection row (get_features of every box, hybridsort.py:394).  The last column
     * of a result row is what the reference writes there: the score of the input row the match indexes (:396-404). */
} b200track_config;

typedef struct b200track_ctx b200track_ctx;

int b200track_abi_version(void);
const char* b200track_last_error(void);

/* ---- tracker contexts: create_tracker(...) / tracker.update(dets, img) ------------------
 * b200track_create      <- tracker_zoo.py:18-118 create_tracker (one ctx = n_streams trackers)
 * b200track_step        <- BYTETracker.update byte_tracker.py:132-281 / OCSort.update
 *                          ocsort.py:218-379 / BoTSORT.update bot_sort.py:231-420, for every
 *                          stream at once; everything stays on the device, asynchronous.
 *   d_dets  [n_streams, max_dets, 6]  (x1, y1, x2, y2, conf, cls), rows >= d_ndets[s] ignored
 *   d_feats [n_streams, max_dets, feat_dim] fp32 appearance rows (BoTSORT) or NULL
 *   img_h, img_w: frame size (OCSORT reads only img.shape[:2]; ocsort.py:239)
 *   d_out   [n_streams, max_tracks, 8] (x1, y1, x2, y2, id, conf, cls, det_ind) in the
 *           reference's row order; d_nout[s] rows are valid.
 *   d_dets, d_feats and d_out must be 16-byte aligned (rows are moved with 16-byte accesses).
 *   StrongSORT contexts (StrongSORT.update, strong_sort.py:43-99): d_feats holds one appearance row per detection row
 *           (what the ReID seam returns); one step is a fixed sequence of eight launches for all streams (strongsort_step.cu).
 * b200track_step_cam    the same with one externally estimated 2x3 camera-motion warp per stream, d_warps
 *                          [n_streams, 6] row-major (NULL = identity): BoT-SORT contexts created with camera_motion
 *                          (STrack.multi_gmc, bot_sort.py:95-111, :293-295) and DeepOCSORT contexts
 *                          (apply_affine_correction, deep_ocsort.py:222-241, :393-396).  For DeepOCSORT d_feats holds the
 *                          embedding of every detection row with conf > det_thresh (deep_ocsort.py:382-390).
 * b200track_step_host   same call with HOST buffers: copies in, steps, copies out, waits.
 * b200track_submit_host / b200track_wait_host: pipelined variant - up to
 *   b200track_host_slots() frames in flight (copy-in of frame k+1 and copy-out of frame k-1
 *   overlap the step of frame k).  Host buffers must stay valid until the wait returns.
 */
int b200track_create(const b200track_config* cfg, b200track_ctx** out_ctx);
void b200track_destroy(b200track_ctx* ctx);
int b200track_reset(b200track_ctx* ctx);
int b200track_step(b200track_ctx* ctx, const double* d_dets, const int32_t* d_ndets,
                   const float* d_feats, int32_t img_h, int32_t img_w,
                   double* d_out, int32_t* d_nout, void* stream);
int b200track_step_cam(b200track_ctx* ctx, const double* d_dets, const int32_t* d_ndets,
                       const float* d_feats, const double* d_warps, int32_t img_h, int32_t img_w,
                       double* d_out, int32_t* d_nout, void* stream);
int b200track_step_host(b200track_ctx* ctx, const double* h_dets, const int32_t* h_ndets,
                        const float* h_feats, int32_t img_h, int32_t img_w,
                        double* h_out, int32_t* h_nout);
int b200track_host_slots(b200track_ctx* ctx);
int b200track_submit_host(b200track_ctx* ctx, int32_t slot, const double* h_dets,
                          const int32_t* h_ndets, const float* h_feats, int32_t img_h,
                          int32_t img_w, double* h_out, int32_t* h_nout);
int b200track_wait_host(b200track_ctx* ctx, int32_t slot);   /* B200TRACK_ERR_CAPACITY if that step overflowed */

/* ---- packed frames: the same step with the bytes a detector produces and a consumer needs --------------------------
 * The padded interface above moves max_dets-row blocks and 64-byte result rows of which conf / cls (and, for OC-SORT,
 * the box) are copies of the caller's own detection row det_ind (byte_tracker.py:270-279, ocsort.py:354-363).  Here a
 * frame is ONE input block and ONE result block, so the host interface is one linear copy per direction:
 *   input block  (b200track_layout.in_*):  int32 offsets[n_streams + 1] (stream s owns detection rows
 *       [offsets[s], offsets[s + 1]) - the exclusive scan of the per-stream counts), double warps[n_streams][6]
 *       (BoT-SORT contexts only; read when flags has B200TRACK_FRAME_HAS_WARPS: the 2x3 camera-motion warp per stream,
 *       bot_sort.py:293-295; DeepOCSORT contexts likewise), the detection rows [n_rows][6] (x1, y1, x2, y2, conf, cls) back to back as fp32
 *       (B200TRACK_F32: what detectors emit; widened on the device, which is exact, so the result equals the padded
 *       interface's on dets.astype(float64)) or fp64, and for BoT-SORT with_reid / DeepOCSORT the embeddings [n_rows][feat_dim] fp32;
 *   result block (b200track_layout.out_*): int32 header[4] ([0] = capacity bits of this step, 0 = fine),
 *       int32 nout[n_streams], then compact rows: stream s owns rows [offsets[s], offsets[s] + nout[s]) (a result row
 *       carries a distinct detection of its frame, so nout[s] <= its detection count), in the reference's row order:
 *         ByteTrack, DeepOCSORT b200track_row    40 B  x1, y1, x2, y2 (fp64), id, det_ind
 *         BoT-SORT  b200track_row_bot 48 B  + cls (the voted class, bot_sort.py:50-67), conf as fp32
 *         OC-SORT   b200track_row_oc   8 B  id, det_ind; det_ind bit 30 set = tracker created this frame (its box is
 *                   convert_x_to_bbox(convert_bbox_to_z(det)), otherwise the detection's own box); bit 29 set = the row
 *                   reports the filter's box (a matched tracker whose observation sums below zero, ocsort.py:355-358):
 *                   header[1] such rows have an entry b200track_exc_oc {row, box} in the exception area that follows
 *                   the rows (out_off_exc, exc_capacity entries; more than that sets a capacity bit)
 * b200track_frame_layout   offsets / sizes of both blocks for n_rows detection rows.
 * b200track_step_packed    device blocks, asynchronous on `stream`.
 * b200track_submit_packed / b200track_wait_packed   HOST blocks through the 3-deep copy / step / copy pipeline (slots are
 *   shared with b200track_submit_host); wait returns B200TRACK_ERR_CAPACITY if that step overflowed. */
typedef enum { B200TRACK_F32 = 0, B200TRACK_F64 = 1 } b200track_dtype;
#define B200TRACK_FRAME_HAS_WARPS 1
#define B200TRACK_ROW_OC_NEW (1 << 30)
#define B200TRACK_ROW_OC_STATE (1 << 29)
typedef struct { double x1, y1, x2, y2; int32_t id; int32_t det_ind; } b200track_row;
typedef struct { double x1, y1, x2, y2; int32_t id; int32_t det_ind; float cls; float conf; } b200track_row_bot;
typedef struct { int32_t id; int32_t det_ind; } b200track_row_oc;
typedef struct { int32_t row; int32_t reserved; double x1, y1, x2, y2; } b200track_exc_oc;   /* row = index into the rows */
typedef struct {
    uint64_t in_bytes, in_off_offsets, in_off_warps, in_off_dets, in_off_feats;
    uint64_t out_bytes, out_off_nout, out_off_rows, out_off_exc;
    int32_t row_bytes, exc_capacity;
} b200track_layout;
int b200track_frame_layout(b200track_ctx* ctx, int64_t n_rows, int32_t det_dtype, b200track_layout* out);
int b200track_step_packed(b200track_ctx* ctx, const void* d_in, int64_t n_rows, int32_t det_dtype, int32_t flags,
                          int32_t img_h, int32_t img_w, void* d_result, void* stream);
int b200track_submit_packed(b200track_ctx* ctx, int32_t slot, const void* h_in, int32_t det_dtype, int32_t flags,
                            int32_t img_h, int32_t img_w, void* h_result);
int b200track_wait_packed(b200track_ctx* ctx, int32_t slot);
/* Waits for all work of the context, then reports capacity overflows seen since the last call. */
int b200track_sync(b200track_ctx* ctx);
/* Sum over streams of the tracks that entered association so far (SURVEY.md 8(d) metric). */
int b200track_track_updates(b200track_ctx* ctx, uint64_t* h_total);
/* Kernel launches issued by this context so far (bench.py's gpu_launches). */
int b200track_launch_count(b200track_ctx* ctx, uint64_t* h_launches);
/* Profiling aid: the first call switches on per-phase cycle counters inside the step kernel (thread 0 of
 * every CTA, clock64 deltas between block barriers); later calls read them.  h_out16[0] = CTAs counted,
 * h_out16[k] = cycles spent before barrier k.  reset != 0 zeroes the counters after reading. */
int b200track_phase_cycles(b200track_ctx* ctx, uint64_t* h_out16, int32_t reset);
/* Bytes one stream's state occupies on the device / dynamic shared memory of the step kernel. */
int b200track_footprint(b200track_ctx* ctx, uint64_t* h_state_bytes_per_stream, uint64_t* h_smem_bytes);

/* Parity probe: the track records of one stream in list order (tracked list, then lost
 * list), expanded to the reference's dense form.  Buffers hold max_tracks rows.
 *   h_counts[4] = n_tracked, n_lost, id counter, frame_id
 *   h_rec [max_tracks, 6] = track_id, state, is_activated, frame_id, start_frame, tracklet_len
 *   h_mean[max_tracks, 8], h_cov[max_tracks, 64], h_aux[max_tracks, 3] = score, cls, det_ind
 * StrongSORT contexts (Track fields, strongsort/sort/track.py:72-99): h_counts = list length, 0, next id, frame count;
 *   h_rec = track_id, state (1 tentative, 2 confirmed), hits, age, time_since_update, stored gallery rows; h_aux = conf,
 *   cls, det_ind; b200track_get_features gives the smoothed feature (Track.features[-1]) of every listed track.
 * OC-SORT contexts (KalmanBoxTracker fields, ocsort.py:65-128): h_counts[0] = live trackers;
 *   h_rec = id, age, time_since_update, hits, hit_streak, kf.observed; h_mean = kf.x[7], has-observation;
 *   h_cov[0:49] = dense 7x7 kf.P, [49:51] = velocity, [51:56] = last_observation. */
int b200track_get_state(b200track_ctx* ctx, int32_t stream_index, int32_t* h_counts,
                        int32_t* h_rec, double* h_mean, double* h_cov, double* h_aux);
/* BoT-SORT contexts (with_reid) and StrongSORT contexts: STrack.smooth_feat (bot_sort.py:40-48) / Track.features[-1] of every listed track of one
 * stream, in the same list order as b200track_get_state; h_feat[max_tracks, feat_dim] fp32. */
int b200track_get_features(b200track_ctx* ctx, int32_t stream_index, float* h_feat);
/* DeepOCSORT contexts: KalmanBoxTracker fields like the OC-SORT form of b200track_get_state (h_rec[.., 5] = kf.observed +
 * 2 * frozen; h_mean = x[8]; h_cov = dense 8x8 kf.P; h_aux = conf, cls, det_ind) plus, in b200track_get_track_extras,
 * h_extra[max_tracks, 8] = velocity[2], last_observation[5] (placeholder -1), unused; h_emb[max_tracks, feat_dim] = the
 * smoothed embedding (fp64), NULL to skip. */
int b200track_get_track_extras(b200track_ctx* ctx, int32_t stream_index, double* h_extra, double* h_emb);
/* HybridSORT contexts: the KalmanBoxTracker list of one stream (hybridsort.py:106-334) in list order.  h_counts[4] as
 * b200track_get_state; h_rec[max_tracks, 6] = id, age, time_since_update, hits, hit_streak, kf.observed; h_x[max_tracks, 9]
 * = kf.x; h_P[max_tracks, 81] = dense kf.P; h_vel[max_tracks, 8] = velocity_lt, _rt, _lb, _rb as (dy, dx), zeros while
 * None; h_last[max_tracks, 5] = last_observation (-1 placeholder); h_aux[max_tracks, 3] = conf, cls, matched input row.
 * Any output pointer may be NULL.  The smoothed embeddings come from b200track_get_features (same order). */
int b200track_get_state_hybridsort(b200track_ctx* ctx, int32_t stream_index, int32_t* h_counts, int32_t* h_rec, double* h_x,
                                   double* h_P, double* h_vel, double* h_last, double* h_aux);
/* Event counters summed over all streams since create / reset (DeepOCSORT contexts): h_out8[0] = first associations solved
 * as an assignment problem (not by the permutation shortcut, association.py:157-159), [1] = recovery rounds that ran an
 * assignment (deep_ocsort.py:466), [2] = observation-centric re-updates (deepocsort_kf.py:433-478); StrongSORT contexts: [3] = stored gallery rows the
 * distance metric compared against (sum over frames and confirmed tracks); the rest reserved. */
int b200track_counters(b200track_ctx* ctx, uint64_t* h_out8);

/* ---- operator level: the reference's functional API, batched, device pointers -----------
 * Dense [n, 8] means / [n, 8, 8] covariances in the reference's memory layout.
 * b200track_kf_initiate        <- KalmanFilter.initiate        bytetrack_kf.py:55 / botsort_kf.py:55
 * b200track_kf_predict         <- KalmanFilter.multi_predict   bytetrack_kf.py:155 / botsort_kf.py:154
 * b200track_kf_project         <- KalmanFilter.project         bytetrack_kf.py:126 / strongsort_kf.py:124
 * b200track_kf_update          <- KalmanFilter.update          bytetrack_kf.py:194 / strongsort_kf.py:157
 * b200track_kf_gating_distance <- KalmanFilter.gating_distance bytetrack_kf.py:228 / strongsort_kf.py:191
 *     (T tracks x D measurements -> d2[T, D]; metric 0 = 'maha', 1 = 'gaussian')
 * d_conf (per row, [n]) is only read for B200TRACK_KF_XYAH_CONF and may be NULL otherwise. */
int b200track_kf_initiate(int32_t kf_kind, int32_t n, const double* d_z, double* d_mean, double* d_cov, void* stream);
int b200track_kf_predict(int32_t kf_kind, int32_t n, double* d_mean, double* d_cov, void* stream);
int b200track_kf_project(int32_t kf_kind, int32_t n, const double* d_mean, const double* d_cov,
                         const double* d_conf, double* d_pmean, double* d_pcov, void* stream);
int b200track_kf_update(int32_t kf_kind, int32_t n, double* d_mean, double* d_cov, const double* d_z,
                        const double* d_conf, void* stream);
/* b200track_kf_apply_warp <- STrack.multi_gmc (bot_sort.py:95-111) / Track.camera_update's matrix form: applies an
 * externally estimated 2x3 camera-motion warp to dense states: mean <- kron(I4, R) mean (+ t on x, y), cov <- R8 cov R8^T.
 * d_warp [n_warps, 6] row-major 2x3; d_warp_index [n] picks the warp of each track (NULL: every track uses warp 0). */
int b200track_kf_apply_warp(int32_t n, double* d_mean, double* d_cov, const double* d_warp, const int32_t* d_warp_index, void* stream);
/* b200track_aw_max_metric <- compute_aw_max_metric (association.py:79-108, DeepOCSORT's adaptive appearance weight):
 * emb [batch, rows, cols] similarity matrices -> out = w_assoc * w_row * w_col * emb, the weights from the two largest
 * entries of every row / column. */
int b200track_aw_max_metric(int32_t batch, int32_t rows, int32_t cols, const double* d_emb, double w_assoc, double bottom,
                            double* d_out, void* stream);
int b200track_kf_gating_distance(int32_t kf_kind, int32_t n_tracks, int32_t n_meas, const double* d_mean,
                                 const double* d_cov, const double* d_meas, int32_t only_position,
                                 int32_t metric, const double* d_conf, double* d_out, void* stream);
/* The same for `batch` independent (tracks, measurements) problems - one per stream - in one launch:
 * mean [batch, T, 8], cov [batch, T, 8, 8], meas [batch, D, 4], conf [batch, T] -> out [batch, T, D]. */
int b200track_kf_gating_distance_batched(int32_t kf_kind, int32_t batch, int32_t n_tracks, int32_t n_meas, const double* d_mean,
                                         const double* d_cov, const double* d_meas, int32_t only_position,
                                         int32_t metric, const double* d_conf, double* d_out, void* stream);
/* b200track_gate_cost <- gate_cost_matrix (matching.py:170-181; fuse = 0) and fuse_motion (:184-196; fuse = 1) for `batch`
 * streams: cost[b, t, d] = inf where the squared Mahalanobis distance exceeds chi2inv95[4] (only_position: [2]), then
 * (fuse) cost = lambda * cost + (1 - lambda) * distance.  In place on d_cost [batch, T, D]; the distance matrix is
 * never materialised.  d_meas are the detections' xyah (to_xyah()). */
int b200track_gate_cost(int32_t kf_kind, int32_t batch, int32_t n_tracks, int32_t n_meas, const double* d_mean,
                        const double* d_cov, const double* d_meas, int32_t only_position, int32_t fuse, double lambda,
                        const double* d_conf, double* d_cost, void* stream);
/* b200track_box_similarity <- iou_batch / giou_batch / diou_batch / ciou_batch / centroid_batch
 *     boxmot/utils/iou.py:6-188 ; a[n,4] x b[m,4] -> out[n,m] (img_w, img_h only for centroid)
 * b200track_iou_distance   <- matching.py:94-119 (1 - iou), optional fuse_score :213-221 when
 *     d_score != NULL */
int b200track_box_similarity(int32_t sim, int32_t n, int32_t m, const double* d_a, const double* d_b,
                             double img_w, double img_h, double* d_out, void* stream);
int b200track_iou_distance(int32_t n, int32_t m, const double* d_a, const double* d_b,
                           const double* d_score, double* d_out, void* stream);
/* b200track_embedding_distance <- matching.py:145-167: fp32 features, max(0, cosine distance)
 *     in double on the fp32 values; a[n, dim] x b[m, dim] -> out[n, m] fp64 */
int b200track_embedding_distance(int32_t n, int32_t m, int32_t dim, const float* d_a, const float* d_b,
                                 double* d_out, void* stream);
/* b200track_appearance_cost <- the thresholded appearance cost its callers build from embedding_distance
 *     (matching.py:145-167): BoT-SORT bot_sort.py:304-306 / :363-365 (emb / 2; emb[emb > appearance_thresh] = 1;
 *     emb[iou mask] = 1), StrongSORT strongsort/sort/linear_assignment.py:59-78 (cost > max_distance ->
 *     max_distance + 1e-5), for `batch` streams at once:
 *       out[b, t, d] = fill  if d_gate && d_gate[b, t, d]  or  scale * max(0, cosine distance) > thresh
 *                    = scale * max(0, cosine distance)   otherwise, exact (fp64 on the fp32 values).
 *     d_trk [batch, n_tracks, dim], d_det [batch, n_dets, dim] fp32, dim a multiple of 64; d_gate uint8
 *     [batch, n_tracks, n_dets] or NULL.  A bf16 tcgen05 GEMM pre-filters, survivors are re-evaluated in
 *     fp64 in the same kernel.  d_workspace: b200track_appearance_cost_workspace() bytes of device memory.
 *     d_stats (device, may be NULL): [0] += entries evaluated exactly, [1] += internal protocol errors. */
int b200track_appearance_cost_workspace(int32_t batch, int32_t n_tracks, int32_t n_dets, int32_t dim, uint64_t* h_bytes);
int b200track_appearance_cost(int32_t batch, int32_t n_tracks, int32_t n_dets, int32_t dim, const float* d_trk,
                              const float* d_det, const uint8_t* d_gate, double scale, double thresh, double fill,
                              double* d_out, void* d_workspace, uint64_t workspace_bytes, uint64_t* d_stats, void* stream);
/* b200track_lapjv <- lap.lapjv(cost, extend_cost=True, cost_limit=L) as called from
 *     matching.py:64 (finite limit) and association.py:23 (cost_limit = +inf): `batch`
 *     independent problems cost[batch, rows, cols] -> x[batch, rows], y[batch, cols] (-1 = unmatched) */
/* b200track_gallery_cost <- NearestNeighborDistanceMetric.distance with the cosine metric (matching.py:247-308, :360-378)
 * followed by the clip of min_cost_matching (strongsort/sort/linear_assignment.py:59-78), for `batch` streams in one launch:
 *     out[b, t, d] = min over g < count[b, t] of 1 - gallery[b, t, g]^ . det[b, d]^   (float32 like the reference)  if <= thresh
 *                  = fill                                                                                            otherwise
 * d_gallery [batch, n_tracks, budget, dim] fp32 (budget <= 128 rows per track, rows past count[b, t] are ignored),
 * d_det [batch, n_dets, dim] fp32 (n_dets <= 256), dim a multiple of 64.  Tensor cores (bf16 tcgen05) pre-filter; every
 * value that is written was recomputed exactly.  d_gallery_bf16: optional unit-norm bf16 copy of the gallery kept by the
 * caller (b200track_unit_bf16 on the rows that changed); NULL = converted here into the workspace on every call.
 * d_stats (3 x uint64, may be NULL): exact row evaluations, protocol errors, surviving pairs. */
int b200track_gallery_cost_workspace(int32_t batch, int32_t n_tracks, int32_t budget, int32_t n_dets, int32_t dim,
                                     int32_t with_gallery, uint64_t* h_bytes);
int b200track_gallery_cost(int32_t batch, int32_t n_tracks, int32_t budget, int32_t n_dets, int32_t dim, const float* d_gallery,
                           const void* d_gallery_bf16, const int32_t* d_count, const float* d_det, double thresh, double fill,
                           double* d_out, void* d_workspace, uint64_t workspace_bytes, uint64_t* d_stats, void* stream);
/* b200track_gallery_append <- NearestNeighborDistanceMetric.partial_fit (matching.py:343-358) on a device-resident gallery
 * [n_slots, budget, dim]: row r goes to gallery[d_slot[r], d_pos[r]] as fp32 and, unit-normalised, as bf16 (the operand
 * formats of b200track_gallery_cost); the caller keeps the ring position of every slot.
 * b200track_ema_unit_features <- Track.update's feature smoothing (strongsort/sort/track.py:166-172), float32 in place:
 * trk <- unit(alpha * trk + (1 - alpha) * unit(det)) for n rows.
 * b200track_camera_update_xyah <- Track.camera_update (track.py:129-138) on n xyah means [n, 8] (in place): the box corners
 * through the 2x3 warp d_warp[6] (NULL: the identity, which is still not an exact no-op in floating point). */
int b200track_gallery_append(int32_t n, int32_t dim, int32_t budget, const float* d_rows, const int32_t* d_slot, const int32_t* d_pos,
                             float* d_gallery, void* d_gallery_bf16, void* stream);
int b200track_ema_unit_features(int32_t n, int32_t dim, float* d_trk, const float* d_det, double alpha, void* stream);
/* rows /= |row| in float32, in place (the first feature of a new StrongSORT track, strongsort/sort/tracker.py:170-172) */
int b200track_unit_features(int32_t n, int32_t dim, float* d_rows, void* stream);
int b200track_camera_update_xyah(int32_t n, double* d_mean, const double* d_warp, void* stream);
/* rows of fp32 -> unit-norm bf16 rows (the operand format of the two tensor-core operators) */
int b200track_unit_bf16(int64_t rows, int32_t dim, const float* d_src, void* d_dst, void* stream);
int b200track_lapjv(int32_t batch, int32_t rows, int32_t cols, const double* d_cost, double cost_limit,
                    int32_t* d_x, int32_t* d_y, void* stream);

/* b200track_nn_cosine_distance <- NearestNeighborDistanceMetric.distance, cosine metric (matching.py:247-308, :360-378):
 *     out[t, d] = min over the gallery rows [d_seg[t], d_seg[t+1]) of track t of 1 - a_hat . b_hat, float32 arithmetic
 *     like the reference; d_gallery [d_seg[n_tracks], dim], d_det [n_dets, dim] fp32 -> d_out [n_tracks, n_dets] fp64. */
int b200track_nn_cosine_distance(int32_t n_tracks, int32_t n_dets, int32_t dim, const float* d_gallery, const int32_t* d_seg,
                                 const float* d_det, double* d_out, void* stream);
/* b200track_linear_sum_assignment <- scipy.optimize.linear_sum_assignment as called by StrongSORT
 *     (strongsort/sort/linear_assignment.py:59-61), bit-faithful including exactly tied costs (the clipped matrix is
 *     mostly one repeated value and the tie behaviour decides the order of `unmatched_detections`): `batch` problems
 *     cost[batch, rows, cols] -> d_col4row[batch, min(rows, cols)]: the assignment of the internal problem, which is
 *     the transposed one when rows > cols exactly like scipy; (row_ind, col_ind) = (arange, col4row) or, transposed,
 *     (col4row[argsort(col4row)], argsort(col4row)).  *d_err |= 1 when a problem is infeasible (inf / nan). */
int b200track_linear_sum_assignment(int32_t batch, int32_t rows, int32_t cols, const double* d_cost,
                                    int32_t* d_col4row, int32_t* d_err, void* stream);

/* ---- OC-SORT's XYSR filter at operator level (csrc/kf_xysr.cu): the 7-d [x, y, s, r, vx, vy, vs] filter KalmanBoxTracker
 * configures (ocsort.py:79-106) on dense d_x [n, 7] / d_P [n, 7, 7] arrays, in place, with the same device functions the fused
 * OC-SORT step runs.  *d_err |= 1 (may be NULL) if a covariance does not have the structure of this filter (three
 * (position, velocity) 2x2 blocks + P_rr).
 * b200track_kf_xysr_predict         <- KalmanFilter.predict (ocsort_kf.py:339-379); the tracker-level guards of
 *                                      KalmanBoxTracker.predict (ocsort.py:168-181) stay with the caller
 * b200track_kf_xysr_update          <- KalmanFilter.update(z) (ocsort_kf.py:437-526: Joseph form), d_z [n, 4]
 * b200track_kf_xysr_unfreeze_update <- KalmanFilter.update(z) on a frozen filter: unfreeze() (:383-434) replays a
 *                                      straight-line virtual trajectory from the last measurement d_last_z [n, 4] to d_z over
 *                                      d_gap [n] frames starting from the state saved by freeze() (freeze is a copy: the
 *                                      caller passes that copy as d_x / d_P), then the real measurement is applied on top;
 *                                      d_virtual_last [n, 4] (may be NULL) receives the last virtual box, which the
 *                                      reference leaves at the end of history_obs. */
int b200track_kf_xysr_predict(int32_t n, double* d_x, double* d_P, int32_t* d_err, void* stream);
int b200track_kf_xysr_update(int32_t n, double* d_x, double* d_P, const double* d_z, int32_t* d_err, void* stream);
int b200track_kf_xysr_unfreeze_update(int32_t n, double* d_x, double* d_P, const double* d_last_z, const int32_t* d_gap,
                                      const double* d_z, double* d_virtual_last, int32_t* d_err, void* stream);

/* ---- HybridSORT's score-carrying filter at operator level (csrc/kf_hybrid.cu): the 9-d [u, v, s, c, r, du, dv, ds, dc]
 * filter KalmanBoxTracker configures (hybridsort.py:126-150) on dense d_x [n, 9] / d_P [n, 9, 9] arrays, in place, with the
 * same device functions the fused HybridSORT step runs.  Measurements are [x, y, s, score, r] (convert_bbox_to_z,
 * hybridsort.py:33-49).  *d_err |= 1 (may be NULL) if a covariance does not have the structure of this filter (four
 * (position, velocity) 2x2 blocks + P_rr).
 * b200track_kf_xyscr_predict         <- KalmanFilter.predict (hybridsort_kf.py:339-379); the tracker-level guard of
 *                                       KalmanBoxTracker.predict (hybridsort.py:303-304) stays with the caller
 * b200track_kf_xyscr_update          <- KalmanFilter.update(z) (hybridsort_kf.py:439-528: Joseph form), d_z [n, 5]
 * b200track_kf_xyscr_unfreeze_update <- KalmanFilter.update(z) on a frozen filter: unfreeze() (:390-436, which unpacks the
 *                                       five-vector as x, y, s, r, c - the score read as the aspect ratio, kept) replays the
 *                                       virtual trajectory from d_last_z [n, 5] to d_z over d_gap [n] frames starting from the
 *                                       state saved by freeze() (passed as d_x / d_P), then applies the real measurement;
 *                                       d_virtual_last [n, 5] (may be NULL) receives the last virtual box. */
int b200track_kf_xyscr_predict(int32_t n, double* d_x, double* d_P, int32_t* d_err, void* stream);
int b200track_kf_xyscr_update(int32_t n, double* d_x, double* d_P, const double* d_z, int32_t* d_err, void* stream);
int b200track_kf_xyscr_unfreeze_update(int32_t n, double* d_x, double* d_P, const double* d_last_z, const int32_t* d_gap,
                                       const double* d_z, double* d_virtual_last, int32_t* d_err, void* stream);

/* ---- DeepOCSORT operators (csrc/kf8.cu): the 8-d [x, y, w, h, vx, vy, vw, vh] filter the reference configures in
 * KalmanBoxTracker.__init__ (deep_ocsort.py:103-138), dense [n, 8] / [n, 8, 8] arrays.
 * b200track_kf8_predict <- KalmanBoxTracker.predict's kf.predict(Q=new_kf_process_noise(w, h)) (deep_ocsort.py:76-80, :263-266,
 *     deepocsort_kf.py:340-381); unit_q != 0: Q = I, the filter default that unfreeze's own predicts use.
 * b200track_kf8_update  <- kf.update(z, R=new_kf_measurement_noise(w, h)) (deep_ocsort.py:83-87, :217-218,
 *     deepocsort_kf.py:549-563: Joseph form, explicit inverse of S); d_wh [n, 2] = the w, h the caller read from the
 *     state BEFORE a possible unfreeze (as the reference does); NULL: R = I.
 * b200track_kf8_oru     <- KalmanFilter.unfreeze (deepocsort_kf.py:433-478) on the restored states: d_box1 [n, 4] =
 *     last_measurement, d_box2 [n, 4] = the new measurement (both read as [x, y, s, r], the reference's quirk), d_gap [n]
 *     = index2 - index1; the whole virtual trajectory of every track in one launch; d_last_virtual [n, 4] = the last
 *     virtual box (the final entry of history_obs afterwards).
 * b200track_ocm_cost    <- associate's cost (association.py:130-172): d_cost[d, t] = -(d_sim[d, t] + angle[d, t] + d_emb[d, t])
 *     with the velocity-direction term from d_dets5 [D, 5], d_vel [T, 2] (dy, dx), d_prev5 [T, 5] (k_previous_obs; col 4 < 0
 *     = none); d_emb may be NULL; d_dets5 NULL: no direction term (the OCR round's -iou, deep_ocsort.py:478).  The canonical
 *     tie-break of the no-limit assignment, + 2^-50 * (d * T + t), is part of the cost.
 * b200track_dot_matrix  <- dets_embs @ trk_embs.T (deep_ocsort.py:433, :464): a [n, dim], b [m, dim] -> out [n, m], fp64. */
int b200track_kf8_predict(int32_t n, double* d_mean, double* d_cov, int32_t unit_q, void* stream);
int b200track_kf8_update(int32_t n, double* d_mean, double* d_cov, const double* d_z, const double* d_wh, void* stream);
int b200track_kf8_oru(int32_t n, double* d_mean, double* d_cov, const double* d_box1, const double* d_box2, const int32_t* d_gap,
                      double* d_last_virtual, void* stream);
int b200track_ocm_cost(int32_t n_dets, int32_t n_tracks, const double* d_dets5, const double* d_vel, const double* d_prev5,
                       double inertia, const double* d_sim, const double* d_emb, double* d_cost, void* stream);
int b200track_dot_matrix(int32_t n, int32_t m, int32_t dim, const double* d_a, const double* d_b, double* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200TRACK_H */

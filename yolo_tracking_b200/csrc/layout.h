// Device-resident track-state layout shared by the kernels and the C-ABI (api.cu).
//
// Per stream s, slot t (capacity Tmax), "stream-major, component-planar":
//   state_f[(s * NF + c) * Tmax + t]   fp64 components
//   state_i[(s * NI + c) * Tmax + t]   int32 components
//   counts [s * 4 + {0: n_tracked, 1: n_lost, 2: id counter, 3: frame_id}]
// Slots [0, n_tracked) are the reference's tracked_stracks list IN LIST ORDER, slots
// [n_tracked, n_tracked + n_lost) its lost_stracks list in list order; every frame step
// rewrites the slots in the new list order, so list order never needs an indirection.
// One CTA owns one stream, so a warp reads/writes contiguous 8-byte runs per component
// (coalesced) and the whole stream state is one contiguous block.
#pragma once

#define B200_NF 22          // 8 mean + 4 x (pp, pv, vv) + score + cls
#define B200_NI 6
#define B200_TF_MEAN 0
#define B200_TF_COV 8       // + 3 * axis + {0: pp, 1: pv, 2: vv}
#define B200_TF_SCORE 20
#define B200_TF_CLS 21
#define B200_TI_ID 0
#define B200_TI_FRAME 1
#define B200_TI_START 2
#define B200_TI_LEN 3
#define B200_TI_DET 4
#define B200_TI_FLAGS 5     // bits 0-1 TrackState, bit 2 is_activated, bit 3 "id is in removed_stracks"

#define B200_ST_NEW 0
#define B200_ST_TRACKED 1
#define B200_ST_LOST 2
#define B200_ST_REMOVED 3
#define B200_FLAG_ACTIVATED 4
#define B200_FLAG_STICKY 8

#define B200_ERR_DET_OVERFLOW 1
#define B200_ERR_TRACK_OVERFLOW 2
#define B200_ERR_BOT_CAPACITY 4     // BoT-SORT: candidate graph overflow or more than 4 classes voted on one track
#define B200_ERR_LSA 16             // StrongSORT: an assignment problem with nan / inf costs (scipy raises ValueError there)
#define B200_ERR_PACKED_ROW 8       // OC-SORT compact rows: more filter-box rows than the exception area of the result block holds
#define B200_ERR_PIPELINE 32        // HybridSORT: a bulk-copy / mbarrier wait of the cosine pass ran into its spin limit (protocol error)

// BoT-SORT contexts created with camera_motion keep the covariance as the two 4x4 blocks a camera warp leaves (kf44.cuh):
// 8 mean + group A (x, y, vx, vy) 10 + group B (w, h, vw, vh) 10 + score + cls
#define B200_NF_CAM 30
#define B200_TFC_COVA 8
#define B200_TFC_COVB 18
#define B200_TFC_SCORE 28
#define B200_TFC_CLS 29

// BoT-SORT contexts carry one more int32 component per slot: the row of the track in the stream's
// embedding pool feat_pool[(s * Tmax + row) * feat_dim] (fp32) and class-vote table
// cls_hist[(s * Tmax + row) * 9] = {cls[4], score sum[4], n}.  Rows never move; slots do.
#define B200_NI_BOT 7
#define B200_TI_FROW 6

// Compact result rows of the packed frame interface (b200track_step_packed):
//   ByteTrack 40 B: double x1, y1, x2, y2; int32 id; int32 det_ind    (conf / cls are the caller's own dets[det_ind, 4:6])
//   BoT-SORT  48 B: double x1, y1, x2, y2; int32 id; int32 det_ind; float cls (the voted class); float conf
//   OC-SORT    8 B: int32 id; int32 det_ind                            (the box is the caller's own dets[det_ind, 0:4])
#define B200_ROW_BYTE 40
#define B200_ROW_BOT 48
#define B200_ROW_OC 8
// OC-SORT compact rows: det_ind bit 30 = row of a tracker created this frame (its box is the detection's round trip
// convert_x_to_bbox(convert_bbox_to_z(det)), ocsort.py:354-360 with the placeholder last_observation)
#define B200_ROW_OC_NEW (1 << 30)
// det_ind bit 29 = the row reports the filter's box, carried by an entry {int32 row, int32 0, double box[4]} of the
// exception area behind the rows (ocsort.py:355-358: last_observation sums below zero)
#define B200_ROW_OC_STATE (1 << 29)
#define B200_EXC_OC_BYTES 40

// bytes a track slot occupies in HBM (one direction)
#define B200_SLOT_BYTES (B200_NF * 8 + B200_NI * 4)

// ---- OC-SORT slot (ocsort_step.cu) -------------------------------------------------------
// Same stream-major, component-planar arrangement; slots [0, n) are the reference's
// `self.trackers` list in list order.  fp64 components:
#define B200_OC_NF 58
#define B200_OC_X 0          // x[7] = x, y, s, r, vx, vy, vs
#define B200_OC_P 7          // 3 x (pp, pv, vv) for (x,vx) (y,vy) (s,vs), then P_rr            (10)
#define B200_OC_LAST 17      // last_observation box (valid when the has-observation flag is set)
#define B200_OC_CONF 21
#define B200_OC_CLS 22
#define B200_OC_VEL 23       // velocity = (dy, dx) / norm, zeros while None
#define B200_OC_RING 25      // observations of the last 3 ages: ring[age % 3][4]
#define B200_OC_SX 37        // frozen x   (attr_saved, written at the first missed frame)
#define B200_OC_SP 44        // frozen P   (10, same order as B200_OC_P)
#define B200_OC_LASTZ 54     // history_obs[index1]: last measurement [x, y, s, r] before a gap
// int32 components:
#define B200_OC_NI 10
#define B200_OCI_ID 0
#define B200_OCI_AGE 1
#define B200_OCI_TSU 2       // time_since_update
#define B200_OCI_HITS 3
#define B200_OCI_STREAK 4
#define B200_OCI_DET 5
#define B200_OCI_RINGAGE 6   // age key of ring[0..2], -1 = empty
#define B200_OCI_FLAGS 9     // bit 0 kf.observed, bit 1 kf.attr_saved is not None, bit 2 last_observation is real
#define B200_OCF_OBSERVED 1
#define B200_OCF_SAVED 2
#define B200_OCF_HASOBS 4
// regular per-step traffic of a slot: x, P, last, conf, cls, vel, ring (37 doubles) + 10 ints
#define B200_OC_SLOT_BYTES (37 * 8 + B200_OC_NI * 4)

// ---- DeepOCSORT slot (deepocsort_step.cu) ---------------------------------------------------------
// Same arrangement as the OC-SORT slot; the 8-d filter [x, y, w, h, vx, vy, vw, vh] keeps its covariance as the two
// independent 4x4 blocks a camera warp leaves (kf44.cuh): group A = (x, y, vx, vy), group B = (w, h, vw, vh), each 10
// numbers in upper-triangle order 00 01 02 03 11 12 13 22 23 33.  fp64 components:
#define B200_DO_NF 80
#define B200_DO_X 0          // x[8]
#define B200_DO_PA 8         // group A covariance (10)
#define B200_DO_PB 18        // group B covariance (10)
#define B200_DO_LAST 28      // last_observation box
#define B200_DO_CONF 32
#define B200_DO_CLS 33
#define B200_DO_VEL 34       // (dy, dx) / norm, zeros while None
#define B200_DO_RING 36      // observations of the last 3 ages: ring[age % 3][4]
#define B200_DO_HOT 48       // components touched every frame
#define B200_DO_SX 48        // frozen x (8), frozen PA (10), frozen PB (10): attr_saved of the filter
#define B200_DO_SPA 56
#define B200_DO_SPB 66
#define B200_DO_LASTZ 76     // last entry of the filter's observation history [4] (a measurement or a virtual box)
// int32 components: the OC-SORT ones (B200_OCI_*) plus the row of the track in the stream's embedding pool
// emb_pool[(s * Tmax + row) * feat_dim] (fp64: the reference's smoothed embedding is float64 after its first blend)
#define B200_DO_NI 11
#define B200_DOI_EROW 10
#define B200_DOF_FROZEN 16   // flag bit 4: KalmanBoxTracker.frozen
#define B200_DO_SLOT_BYTES (B200_DO_HOT * 8 + B200_DO_NI * 4)

// ---- HybridSORT slot (hybridsort_step.cu) ---------------------------------------------------------
// Same arrangement as the OC-SORT slot; the 9-d filter [u, v, s, c, r, du, dv, ds, dc] keeps four (position, velocity)
// 2x2 blocks and P_rr (kf_hybrid.cuh).  fp64 components:
#define B200_HY_NF 75
#define B200_HY_X 0          // x[9]
#define B200_HY_P 9          // 4 x (pp, pv, vv) for (u,du) (v,dv) (s,ds) (c,dc), then P_rr                        (13)
#define B200_HY_LAST 22      // last_observation box (valid when the has-observation flag is set)
#define B200_HY_CONF 26
#define B200_HY_CLS 27
#define B200_HY_VEL 28       // velocity_lt, _rt, _lb, _rb as (dy, dx) sums of unit vectors, zeros while None       (8)
#define B200_HY_RING 36      // observations of the last 3 ages: ring[age % 3][4]
#define B200_HY_HOT 48       // components touched every frame
#define B200_HY_SX 48        // frozen x (9)
#define B200_HY_SP 57        // frozen P (13, same order as B200_HY_P)
#define B200_HY_LASTZ 70     // last entry of the filter's observation history [x, y, s, score, r] (measurement or virtual box)
// int32 components: the OC-SORT ones (B200_OCI_*) plus the row of the track in the stream's embedding pool
// feat_pool[(s * Tmax + row) * feat_dim] (fp32: every embedding operation of the reference is float32, hybridsort.py:188-205)
#define B200_HY_NI 11
#define B200_HYI_FROW 10
#define B200_HY_SLOT_BYTES (B200_HY_HOT * 8 + B200_HY_NI * 4)

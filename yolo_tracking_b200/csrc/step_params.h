// Kernel-side parameter block of one batched frame step (filled by api.cu from b200track_config).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200 {

struct StepParams {
    int n_streams, max_tracks, max_dets, feat_dim;
    // ByteTrack / BoTSORT thresholds (bytetrack.yaml / botsort.yaml keys)
    double track_thresh;      // track_thresh | track_high_thresh
    double low_thresh;        // 0.1          | track_low_thresh
    double new_thresh;        // det_thresh (= track_thresh) | new_track_thresh
    double match_thresh;      // first association cost limit
    double second_thresh;     // 0.5
    double unconf_thresh;     // 0.7
    double dup_thresh;        // 0.15
    double proximity_thresh, appearance_thresh;   // BoTSORT
    int max_time_lost;
    // OC-SORT (ocsort.yaml keys)
    double det_thresh, iou_thresh, inertia, img_w, img_h;
    int max_age, min_hits, delta_t, asso_func, use_byte;
    double* scratch;          // [S, Dcap, Tcap] dense cost / appearance matrices (DeepOCSORT with a dense similarity, HybridSORT)
    // BoT-SORT
    int with_reid;
    int fuse_first;           // fuse_first_associate (bot_sort.py:300-301)
    float* feat_pool;         // [S, Tcap, feat_dim] smoothed track embeddings, row-indexed (layout.h)
    double* cls_hist;         // [S, Tcap, 9] class-vote tables, row-indexed
    // DeepOCSORT (deepocsort.yaml keys)
    double w_assoc_emb, alpha_fixed_emb, aw_param;
    int embedding_off, aw_off;
    double* emb_pool;         // [S, Tcap, feat_dim] fp64 smoothed track embeddings, row-indexed (layout.h)
    // device state (layout.h)
    double* state_f;
    int* state_i;
    int* counts;
    unsigned long long* track_updates;
    int* err;
    int* err_slot;            // [host slots] per-step capacity words of the padded host interface
    unsigned long long* dbg;  // [16] optional phase cycle counters, null = off
    unsigned long long* stats;   // [8] event counters over all streams (b200track_counters): [0] first associations solved as an
                                 // assignment problem, [1] recovery rounds that ran theirs, [2] observation-centric re-updates
    // per-step inputs / outputs (device)
    const double* dets;       // [S, max_dets, 6]
    const int* ndets;         // [S]
    const float* feats;       // [S, max_dets, feat_dim] or null
    double* out;              // [S, max_tracks, 8]
    int* nout;                // [S]
    // packed frame interface (b200track_step_packed, include/b200track.h): det_off != null selects it.  The detection
    // rows of all streams lie back to back (stream s owns rows [det_off[s], det_off[s + 1])), as fp32 (dets32) or fp64
    // (dets) rows of 6; embeddings, if any, are packed the same way (feats[row]); the result rows of stream s are
    // written compactly (layout.h: B200_ROW_*) at the same row offsets - every result row carries a distinct detection
    // of its frame, so a stream never has more result rows than detections.
    const float* dets32;      // [R, 6] or null
    const int* det_off;       // [S + 1]
    unsigned char* rows;      // [R] compact result rows
    int* err_out;             // header of the result block: [0] capacity overflow bits of this step, [1] exception entries
    unsigned char* exc;       // OC-SORT: exception area of the result block (layout.h: B200_ROW_OC_STATE)
    int exc_cap;
    const double* warps;      // [S, 6] row-major 2x3 camera-motion warp per stream (BoT-SORT), null = identity
};

// The step kernel is compiled for a few (slot capacity, detection capacity) pairs; a context
// uses the smallest one that covers its max_tracks / max_dets.  The device state is laid out
// with the variant's slot capacity as stride.
int bytetrack_step_variant(int max_tracks, int max_dets);   // -1: nothing large enough
int bytetrack_step_tmax(int variant);
size_t bytetrack_step_smem(int variant, bool botsort = false, bool cam = false);
cudaError_t launch_bytetrack_step(const StepParams& p, int kf_kind, int variant, cudaStream_t stream);
cudaError_t launch_botsort_step(const StepParams& p, int variant, cudaStream_t stream, bool cam = false);
// the same steps on packed frames (bytetrack_step_packed.cu)
cudaError_t launch_bytetrack_step_packed(const StepParams& p, int kf_kind, int variant, cudaStream_t stream);
cudaError_t launch_botsort_step_packed(const StepParams& p, int variant, cudaStream_t stream, bool cam = false);
size_t ocsort_step_smem(int variant);
cudaError_t launch_ocsort_step(const StepParams& p, int variant, cudaStream_t stream);
int step_variant_dmax(int variant);
size_t deepocsort_step_smem(int variant);
cudaError_t launch_deepocsort_step(const StepParams& p, int variant, cudaStream_t stream);
size_t hybridsort_step_smem(int variant);
cudaError_t launch_hybridsort_step(const StepParams& p, int variant, cudaStream_t stream);

}  // namespace b200

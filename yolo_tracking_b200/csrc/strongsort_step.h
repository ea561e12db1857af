// Batched StrongSORT frame step (strongsort_step.cu): device state and per-frame scratch of one context.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200 {

// int32 components of a slot, planar per stream: ti[(s * SS_NI + c) * T + slot]
constexpr int SS_NI = 7;
constexpr int SSI_ID = 0, SSI_STATE = 1, SSI_HITS = 2, SSI_AGE = 3, SSI_TSU = 4, SSI_DET = 5, SSI_APPENDED = 6;
constexpr int SS_FREE = 0, SS_TENTATIVE = 1, SS_CONFIRMED = 2;   // TrackState (strongsort/sort/track.py:7-18); 0 = free slot

struct SSParams {
    int n_streams, T, D, F, budget;      // T: slots per stream (= max_tracks), D: detection rows per stream (= max_dets)
    double max_dist, max_iou_dist, mc_lambda;
    float ema_alpha, ema_beta;           // float32(alpha), float32(1 - alpha): numpy's float32 array times a Python float
    int max_age, n_init;
    // ---- state: slots never move, `order` is the reference's self.tracks list ----
    double* mean;                        // [S, T, 8]   dense, the layout of the Kalman operators (ops.cu)
    double* cov;                         // [S, T, 8, 8]
    double* conf;                        // [S, T]
    double* cls;                         // [S, T]
    int* ti;                             // [S, SS_NI, T]
    int* order;                          // [S, T] list position -> slot
    int* counts;                         // [S, 4] = list length, next id, frame count, unused
    float* feat;                         // [S, T, F] the track's smoothed feature (Track.features[-1])
    float* gal32;                        // [S, T, budget, F] gallery ring of every slot (NearestNeighborDistanceMetric.samples)
    void* gal16;                         // the same rows, unit-norm bf16: the A operand of the tensor-core distance
    // ---- per-frame scratch ----
    double* cost;                        // [S, T, D] gallery distance, then gated + fused (rows = slots)
    int* gcount;                         // [S, T] gallery rows of a confirmed slot, 0 otherwise
    double* meas;                        // [S, D, 4] detections as xyah
    double* tlwh;                        // [S, D, 4]
    double* dconf;                       // [S, D]
    int* match;                          // [S, T] slot -> matched detection, -1 = none
    double* iou;                         // [S, T * D] cost matrix of the IoU round
    int* ud;                             // [S, D] unmatched detections in the reference's order
    int* nud;                            // [S]
    void* ws;                            // workspace of the gallery distance
    uint64_t ws_bytes;
    unsigned long long* gstats;          // [3] counters of the gallery distance ([1] = protocol errors)
    unsigned long long* stats;           // context counters (b200track_counters): [3] += gallery rows compared this step
    unsigned long long* track_updates;   // [S]
    int* err;
};

size_t strongsort_match_smem();
// one frame for all streams: dets [S, D, 6], ndets [S], feats [S, D, F] fp32, warps [S, 6] or null -> out [S, T, 8], nout [S]
int launch_strongsort_step(const SSParams& p, const double* dets, const int* ndets, const float* feats, const double* warps,
                           double* out, int* nout, int* err_step, cudaStream_t st);
cudaError_t launch_kf_update_masked(int kind, int n_streams, int T, int D, double* mean, double* cov, const double* meas,
                                    const double* conf, const int* sel, cudaStream_t st);

}  // namespace b200

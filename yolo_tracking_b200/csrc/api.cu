// C-ABI of the tracker contexts (include/b200track.h): device state, the batched frame step,
// the host-buffer step with a 3-deep copy/compute/copy pipeline, and the parity probe.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200track.h"
#include "api_util.h"
#include "layout.h"
#include "step_params.h"
#include "strongsort_step.h"

namespace b200 {
thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace b200

using b200::set_error;

#define CU_TRY(expr)                                                                       \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
            return B200TRACK_ERR_CUDA;                                                     \
        }                                                                                  \
    } while (0)

namespace {
constexpr int NSLOT = 3;

// every entry point runs on the context's device and leaves the caller's current device as it found it
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev); else if (err == cudaSuccess) prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(ctx) DeviceGuard _guard((ctx)->cfg.device); CU_TRY(_guard.err)

struct HostSlot {
    double* d_dets = nullptr;
    int32_t* d_ndets = nullptr;
    float* d_feats = nullptr;
    double* d_out = nullptr;
    int32_t* d_nout = nullptr;
    cudaEvent_t in_ready = nullptr, done = nullptr, out_ready = nullptr;
    bool used = false;
    int32_t* h_err = nullptr;           // pinned: capacity bits of the step this slot carried (b200track_wait_host)
    // packed frames: one input block, one result block (b200track_frame_layout)
    unsigned char* d_in = nullptr;
    unsigned char* d_res = nullptr;
    const int32_t* h_res = nullptr;     // the caller's result block of the step in flight (header read by wait_packed)
    bool packed = false;
};
}  // namespace

struct b200track_ctx {
    b200track_config cfg;
    b200::StepParams p;
    int kf_kind = 0;
    int variant = 0;
    int tcap = 0;               // slot stride of the device state (the kernel variant's capacity)
    int nf = B200_NF, ni = B200_NI;   // fp64 / int32 components per slot (depends on the tracker kind)
    cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    HostSlot slot[NSLOT];
    uint64_t launches = 0;
    int32_t* h_err = nullptr;   // pinned
    b200::SSParams ss{};         // StrongSORT contexts: state + scratch of the multi-launch step (strongsort_step.cu)
};

static int check_cfg(const b200track_config* c) {
    if (!c) { set_error("cfg is NULL"); return B200TRACK_ERR_ARG; }
    if (c->n_streams <= 0) { set_error("n_streams must be > 0"); return B200TRACK_ERR_ARG; }
    if (c->max_tracks <= 0 || c->max_tracks % 32 || c->max_tracks > 512) {
        set_error("max_tracks must be a multiple of 32 in [32, 512]"); return B200TRACK_ERR_ARG; }
    if (c->max_dets <= 0 || c->max_dets % 32 || c->max_dets > 512) {
        set_error("max_dets must be a multiple of 32 in [32, 512]"); return B200TRACK_ERR_ARG; }
    if (c->kind != B200TRACK_BYTETRACK && c->kind != B200TRACK_OCSORT && c->kind != B200TRACK_BOTSORT && c->kind != B200TRACK_DEEPOCSORT &&
        c->kind != B200TRACK_STRONGSORT && c->kind != B200TRACK_HYBRIDSORT) {
        set_error("unknown tracker kind"); return B200TRACK_ERR_ARG; }
    return 0;
}

extern "C" int b200track_abi_version(void) { return B200TRACK_ABI_VERSION; }
extern "C" const char* b200track_last_error(void) { return b200::g_last_error.c_str(); }

extern "C" void b200track_destroy(b200track_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard _guard(ctx->cfg.device);
    cudaDeviceSynchronize();
    cudaFree(ctx->p.state_f); cudaFree(ctx->p.state_i); cudaFree(ctx->p.counts);
    cudaFree(ctx->p.track_updates); cudaFree(ctx->p.err); cudaFree(ctx->p.stats); cudaFree(ctx->p.err_slot); cudaFree(ctx->p.dbg); cudaFree(ctx->p.scratch);
    cudaFree(ctx->p.feat_pool); cudaFree(ctx->p.cls_hist); cudaFree(ctx->p.emb_pool);
    {
        b200::SSParams& q = ctx->ss;
        cudaFree(q.mean); cudaFree(q.cov); cudaFree(q.conf); cudaFree(q.cls); cudaFree(q.ti); cudaFree(q.order); cudaFree(q.feat);
        cudaFree(q.gal32); cudaFree(q.gal16); cudaFree(q.cost); cudaFree(q.gcount); cudaFree(q.meas); cudaFree(q.tlwh); cudaFree(q.dconf);
        cudaFree(q.match); cudaFree(q.iou); cudaFree(q.ud); cudaFree(q.nud); cudaFree(q.ws); cudaFree(q.gstats);
    }
    for (auto& s : ctx->slot) {
        cudaFree(s.d_dets); cudaFree(s.d_ndets); cudaFree(s.d_feats); cudaFree(s.d_out); cudaFree(s.d_nout);
        cudaFree(s.d_in); cudaFree(s.d_res);
        if (s.h_err) cudaFreeHost(s.h_err);
        if (s.in_ready) cudaEventDestroy(s.in_ready);
        if (s.done) cudaEventDestroy(s.done);
        if (s.out_ready) cudaEventDestroy(s.out_ready);
    }
    if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    delete ctx;
}

extern "C" int b200track_reset(b200track_ctx* ctx) {
    if (!ctx) { set_error("ctx is NULL"); return B200TRACK_ERR_ARG; }
    ON_DEVICE(ctx);
    CU_TRY(cudaDeviceSynchronize());
    const size_t S = ctx->cfg.n_streams, T = ctx->tcap;
    if (ctx->cfg.kind == B200TRACK_STRONGSORT) {
        b200::SSParams& q = ctx->ss;
        CU_TRY(cudaMemset(q.mean, 0, S * T * 8 * sizeof(double)));
        CU_TRY(cudaMemset(q.cov, 0, S * T * 64 * sizeof(double)));
        CU_TRY(cudaMemset(q.ti, 0, S * b200::SS_NI * T * sizeof(int)));
        CU_TRY(cudaMemset(q.order, 0, S * T * sizeof(int)));
        CU_TRY(cudaMemset(q.feat, 0, S * T * (size_t)q.F * sizeof(float)));
        CU_TRY(cudaMemset(q.gstats, 0, 3 * sizeof(unsigned long long)));
        std::vector<int> c(S * 4, 0);
        for (size_t i = 0; i < S; ++i) c[4 * i + 1] = 1;                // Tracker._next_id = 1 (tracker.py:57)
        CU_TRY(cudaMemcpy(ctx->p.counts, c.data(), c.size() * sizeof(int), cudaMemcpyHostToDevice));
    } else {
        CU_TRY(cudaMemset(ctx->p.state_f, 0, S * ctx->nf * T * sizeof(double)));
        CU_TRY(cudaMemset(ctx->p.state_i, 0, S * ctx->ni * T * sizeof(int)));
        CU_TRY(cudaMemset(ctx->p.counts, 0, S * 4 * sizeof(int)));
    }
    CU_TRY(cudaMemset(ctx->p.track_updates, 0, S * sizeof(unsigned long long)));
    CU_TRY(cudaMemset(ctx->p.err, 0, sizeof(int)));
    CU_TRY(cudaMemset(ctx->p.stats, 0, 8 * sizeof(unsigned long long)));
    return 0;
}

extern "C" int b200track_create(const b200track_config* cfg, b200track_ctx** out_ctx) {
    if (!out_ctx) { set_error("out_ctx is NULL"); return B200TRACK_ERR_ARG; }
    *out_ctx = nullptr;
    if (int rc = check_cfg(cfg)) return rc;
    if (cfg->kind == B200TRACK_BOTSORT && cfg->with_reid) {
        if (cfg->feat_dim <= 0 || cfg->feat_dim % 128 || cfg->feat_dim > 4096) {
            set_error("BoT-SORT with_reid needs feat_dim to be a multiple of 128 in [128, 4096]"); return B200TRACK_ERR_ARG; }
    }
    if (cfg->kind == B200TRACK_OCSORT || cfg->kind == B200TRACK_DEEPOCSORT || cfg->kind == B200TRACK_HYBRIDSORT) {
        if (cfg->delta_t < 1 || cfg->delta_t > 3) { set_error("OC-SORT / DeepOCSORT / HybridSORT delta_t must be in [1, 3]"); return B200TRACK_ERR_ARG; }
        if (cfg->asso_func < 0 || cfg->asso_func > B200TRACK_SIM_CENTROID) { set_error("unknown asso_func"); return B200TRACK_ERR_ARG; }
    }
    if (cfg->kind == B200TRACK_DEEPOCSORT && !cfg->embedding_off) {
        if (cfg->feat_dim <= 0 || cfg->feat_dim % 4 || cfg->feat_dim > 4096) {
            set_error("DeepOCSORT needs feat_dim to be a multiple of 4 in [4, 4096] (or embedding_off)"); return B200TRACK_ERR_ARG; }
    }
    if (cfg->kind == B200TRACK_HYBRIDSORT) {
        if (cfg->feat_dim <= 0 || cfg->feat_dim % 4 || cfg->feat_dim > 4096) {
            set_error("HybridSORT needs feat_dim to be a multiple of 4 in [4, 4096]"); return B200TRACK_ERR_ARG; }
        if (cfg->use_byte) { set_error("HybridSORT: use_byte is not supported (the reference's branch cannot produce a result row)"); return B200TRACK_ERR_ARG; }
    }
    if (cfg->kind == B200TRACK_STRONGSORT) {
        if (cfg->feat_dim <= 0 || cfg->feat_dim % 64 || cfg->feat_dim > 4096) {
            set_error("StrongSORT needs feat_dim to be a multiple of 64 in [64, 4096]"); return B200TRACK_ERR_ARG; }
        if (cfg->nn_budget < 1 || cfg->nn_budget > 128) { set_error("StrongSORT nn_budget must be in [1, 128]"); return B200TRACK_ERR_ARG; }
        if (cfg->max_tracks > 256 || cfg->max_dets > 256) { set_error("StrongSORT contexts take at most 256 tracks / detections per stream"); return B200TRACK_ERR_CAPACITY; }
        if (!(cfg->mc_lambda > 0.0) || cfg->n_streams > 65535) { set_error("StrongSORT needs mc_lambda > 0 and at most 65535 streams"); return B200TRACK_ERR_ARG; }
    }
    int ndev = 0;
    CU_TRY(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) { set_error("no such CUDA device"); return B200TRACK_ERR_CUDA; }
    DeviceGuard _guard(cfg->device);
    CU_TRY(_guard.err);
    b200track_ctx* ctx = new b200track_ctx();
    ctx->cfg = *cfg;
    b200::StepParams& p = ctx->p;
    memset(&p, 0, sizeof(p));
    p.n_streams = cfg->n_streams; p.max_tracks = cfg->max_tracks; p.max_dets = cfg->max_dets; p.feat_dim = cfg->feat_dim;
    p.track_thresh = cfg->track_thresh;
    p.low_thresh = cfg->track_low_thresh;
    p.new_thresh = cfg->new_track_thresh;
    p.match_thresh = cfg->match_thresh;
    p.second_thresh = 0.5;      // byte_tracker.py:211
    p.unconf_thresh = 0.7;      // byte_tracker.py:233
    p.dup_thresh = 0.15;        // byte_tracker.py:314
    p.proximity_thresh = cfg->proximity_thresh;
    p.appearance_thresh = cfg->appearance_thresh;
    p.max_time_lost = (int)(cfg->frame_rate / 30.0 * cfg->track_buffer);   // byte_tracker.py:128-129
    p.det_thresh = cfg->det_thresh; p.iou_thresh = cfg->iou_thresh; p.inertia = cfg->inertia;
    p.max_age = cfg->max_age; p.min_hits = cfg->min_hits; p.delta_t = cfg->delta_t; p.asso_func = cfg->asso_func; p.use_byte = cfg->use_byte ? 1 : 0;
    if (cfg->kind == B200TRACK_OCSORT) { ctx->nf = B200_OC_NF; ctx->ni = B200_OC_NI; }
    if (cfg->kind == B200TRACK_HYBRIDSORT) { ctx->nf = B200_HY_NF; ctx->ni = B200_HY_NI; }
    if (cfg->kind == B200TRACK_DEEPOCSORT) {
        ctx->nf = B200_DO_NF; ctx->ni = B200_DO_NI;
        p.w_assoc_emb = cfg->w_association_emb; p.alpha_fixed_emb = cfg->alpha_fixed_emb; p.aw_param = cfg->aw_param;
        p.embedding_off = cfg->embedding_off ? 1 : 0; p.aw_off = cfg->aw_off ? 1 : 0;
        if (p.embedding_off) p.feat_dim = 0;
    }
    if (cfg->kind == B200TRACK_BOTSORT) {
        ctx->ni = B200_NI_BOT; p.with_reid = cfg->with_reid ? 1 : 0; p.fuse_first = cfg->fuse_first_associate ? 1 : 0;
        if (cfg->camera_motion) ctx->nf = B200_NF_CAM;
    }
    ctx->kf_kind = cfg->kind == B200TRACK_BOTSORT ? B200TRACK_KF_XYWH : B200TRACK_KF_XYAH;
    ctx->variant = b200::bytetrack_step_variant(cfg->max_tracks, cfg->max_dets);
    if (ctx->variant < 0) { set_error("no kernel variant covers max_tracks / max_dets"); delete ctx; return B200TRACK_ERR_CAPACITY; }
    ctx->tcap = b200::bytetrack_step_tmax(ctx->variant);
    const size_t S = cfg->n_streams, T = ctx->tcap, D = cfg->max_dets;
    int rc = 0;
    auto fail = [&](int code) { b200track_destroy(ctx); return code; };
#define CU_TRY_CTX(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); return fail(B200TRACK_ERR_CUDA); } } while (0)
    if (cfg->kind == B200TRACK_STRONGSORT) {
        ctx->tcap = cfg->max_tracks;
        const size_t Ts = cfg->max_tracks, F = cfg->feat_dim, G = cfg->nn_budget;
        b200::SSParams& q = ctx->ss;
        q.n_streams = cfg->n_streams; q.T = (int)Ts; q.D = cfg->max_dets; q.F = cfg->feat_dim; q.budget = cfg->nn_budget;
        q.max_dist = cfg->max_dist; q.max_iou_dist = cfg->max_iou_dist; q.mc_lambda = cfg->mc_lambda;
        q.ema_alpha = (float)cfg->ema_alpha; q.ema_beta = (float)(1.0 - cfg->ema_alpha);
        q.max_age = cfg->max_age; q.n_init = cfg->n_init;
        CU_TRY_CTX(cudaMalloc(&q.mean, S * Ts * 8 * sizeof(double)));
        CU_TRY_CTX(cudaMalloc(&q.cov, S * Ts * 64 * sizeof(double)));
        CU_TRY_CTX(cudaMalloc(&q.conf, S * Ts * sizeof(double)));
        CU_TRY_CTX(cudaMalloc(&q.cls, S * Ts * sizeof(double)));
        CU_TRY_CTX(cudaMalloc(&q.ti, S * b200::SS_NI * Ts * sizeof(int)));
        CU_TRY_CTX(cudaMalloc(&q.order, S * Ts * sizeof(int)));
        CU_TRY_CTX(cudaMalloc(&q.feat, S * Ts * F * sizeof(float)));
        CU_TRY_CTX(cudaMalloc(&q.gal32, S * Ts * G * F * sizeof(float)));
        CU_TRY_CTX(cudaMalloc(&q.gal16, S * Ts * G * F * 2));
        CU_TRY_CTX(cudaMemset(q.gal32, 0, S * Ts * G * F * sizeof(float)));
        CU_TRY_CTX(cudaMemset(q.gal16, 0, S * Ts * G * F * 2));
        CU_TRY_CTX(cudaMalloc(&q.cost, S * Ts * D * sizeof(double)));
        CU_TRY_CTX(cudaMalloc(&q.iou, S * Ts * D * sizeof(double)));
        CU_TRY_CTX(cudaMalloc(&q.gcount, S * Ts * sizeof(int)));
        CU_TRY_CTX(cudaMalloc(&q.match, S * Ts * sizeof(int)));
        CU_TRY_CTX(cudaMalloc(&q.meas, S * D * 4 * sizeof(double)));
        CU_TRY_CTX(cudaMalloc(&q.tlwh, S * D * 4 * sizeof(double)));
        CU_TRY_CTX(cudaMalloc(&q.dconf, S * D * sizeof(double)));
        CU_TRY_CTX(cudaMalloc(&q.ud, S * D * sizeof(int)));
        CU_TRY_CTX(cudaMalloc(&q.nud, S * sizeof(int)));
        CU_TRY_CTX(cudaMalloc(&q.gstats, 3 * sizeof(unsigned long long)));
        uint64_t wsb = 0;
        if (b200track_gallery_cost_workspace(cfg->n_streams, (int)Ts, cfg->nn_budget, cfg->max_dets, cfg->feat_dim, 0, &wsb)) return fail(B200TRACK_ERR_ARG);
        q.ws_bytes = wsb;
        CU_TRY_CTX(cudaMalloc(&q.ws, wsb ? wsb : 16));
    } else {
    CU_TRY_CTX(cudaMalloc(&p.state_f, S * ctx->nf * T * sizeof(double)));
    CU_TRY_CTX(cudaMalloc(&p.state_i, S * ctx->ni * T * sizeof(int)));
    }
    if (cfg->kind == B200TRACK_DEEPOCSORT && !p.embedding_off) {
        CU_TRY_CTX(cudaMalloc(&p.emb_pool, S * T * (size_t)cfg->feat_dim * sizeof(double)));
        // diou / ciou / centroid have a similarity - hence an appearance term - for every pair: per-stream scratch matrix
        if (!(cfg->asso_func <= B200TRACK_SIM_GIOU && cfg->iou_thresh >= 0.0))
            CU_TRY_CTX(cudaMalloc(&p.scratch, S * T * (size_t)b200::step_variant_dmax(ctx->variant) * sizeof(double)));
    }
    if (cfg->kind == B200TRACK_HYBRIDSORT) {
        CU_TRY_CTX(cudaMalloc(&p.feat_pool, S * T * (size_t)cfg->feat_dim * sizeof(float)));
        CU_TRY_CTX(cudaMemset(p.feat_pool, 0, S * T * (size_t)cfg->feat_dim * sizeof(float)));
        // every pair has an appearance term: per-stream cost matrix [detection capacity of the variant][slot capacity]
        CU_TRY_CTX(cudaMalloc(&p.scratch, S * T * (size_t)b200::step_variant_dmax(ctx->variant) * sizeof(double)));
    }
    if (cfg->kind == B200TRACK_BOTSORT) {
        CU_TRY_CTX(cudaMalloc(&p.cls_hist, S * T * 9 * sizeof(double)));
        if (cfg->with_reid) {
            CU_TRY_CTX(cudaMalloc(&p.feat_pool, S * T * (size_t)cfg->feat_dim * sizeof(float)));
        }
    }
    CU_TRY_CTX(cudaMalloc(&p.counts, S * 4 * sizeof(int)));
    CU_TRY_CTX(cudaMalloc(&p.track_updates, S * sizeof(unsigned long long)));
    CU_TRY_CTX(cudaMalloc(&p.err, sizeof(int)));
    CU_TRY_CTX(cudaMalloc(&p.stats, 8 * sizeof(unsigned long long)));
    CU_TRY_CTX(cudaMalloc(&p.err_slot, NSLOT * sizeof(int)));
    CU_TRY_CTX(cudaMemset(p.err_slot, 0, NSLOT * sizeof(int)));
    CU_TRY_CTX(cudaMallocHost(&ctx->h_err, sizeof(int32_t)));
    CU_TRY_CTX(cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
    CU_TRY_CTX(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    CU_TRY_CTX(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    for (auto& s : ctx->slot) {
        // the device staging buffers of a slot are allocated on first use (padded or packed form)
        CU_TRY_CTX(cudaMallocHost(&s.h_err, sizeof(int32_t)));
        *s.h_err = 0;
        CU_TRY_CTX(cudaEventCreateWithFlags(&s.in_ready, cudaEventDisableTiming));
        CU_TRY_CTX(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        CU_TRY_CTX(cudaEventCreateWithFlags(&s.out_ready, cudaEventDisableTiming));
    }
    if (cfg->kind == B200TRACK_STRONGSORT) {
        ctx->ss.counts = p.counts; ctx->ss.track_updates = p.track_updates; ctx->ss.err = p.err; ctx->ss.stats = p.stats;
    }
    const size_t smem = cfg->kind == B200TRACK_STRONGSORT ? b200::strongsort_match_smem()
                        : cfg->kind == B200TRACK_OCSORT ? b200::ocsort_step_smem(ctx->variant)
                        : cfg->kind == B200TRACK_DEEPOCSORT ? b200::deepocsort_step_smem(ctx->variant)
                        : cfg->kind == B200TRACK_HYBRIDSORT ? b200::hybridsort_step_smem(ctx->variant)
                        : b200::bytetrack_step_smem(ctx->variant, cfg->kind == B200TRACK_BOTSORT, cfg->kind == B200TRACK_BOTSORT && cfg->camera_motion);
    int max_smem = 0;
    CU_TRY_CTX(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
    if (smem > (size_t)max_smem) {
        set_error("max_tracks / max_dets need more shared memory than one CTA can have");
        return fail(B200TRACK_ERR_CAPACITY);
    }
    rc = b200track_reset(ctx);
    if (rc) return fail(rc);
    *out_ctx = ctx;
    return 0;
}

static int launch_step(b200track_ctx* ctx, const double* d_dets, const int32_t* d_ndets, const float* d_feats,
                       int32_t img_h, int32_t img_w, double* d_out, int32_t* d_nout, cudaStream_t st, int* d_err_step = nullptr,
                       const double* d_warps = nullptr) {
    b200::StepParams p = ctx->p;
    p.dets = d_dets; p.ndets = d_ndets; p.feats = d_feats; p.out = d_out; p.nout = d_nout; p.err_out = d_err_step;
    p.img_h = img_h; p.img_w = img_w; p.warps = d_warps;
    if (ctx->cfg.kind == B200TRACK_STRONGSORT) {
        if (!d_feats) { set_error("StrongSORT: the embedding buffer is NULL"); return B200TRACK_ERR_ARG; }
        if (int rc = b200::launch_strongsort_step(ctx->ss, d_dets, d_ndets, d_feats, d_warps, d_out, d_nout, d_err_step, st)) return rc;
        ctx->launches += 8;
        return 0;
    }
    if (d_warps && !(ctx->cfg.kind == B200TRACK_DEEPOCSORT || (ctx->cfg.kind == B200TRACK_BOTSORT && ctx->cfg.camera_motion))) {
        set_error("camera-motion warps need a DeepOCSORT context or a BoT-SORT context created with camera_motion"); return B200TRACK_ERR_STATE; }
    if (ctx->cfg.kind == B200TRACK_DEEPOCSORT) {
        if (!ctx->p.embedding_off && !d_feats) { set_error("DeepOCSORT: the embedding buffer is NULL"); return B200TRACK_ERR_ARG; }
        CU_TRY(b200::launch_deepocsort_step(p, ctx->variant, st));
    } else if (ctx->cfg.kind == B200TRACK_OCSORT) CU_TRY(b200::launch_ocsort_step(p, ctx->variant, st));
    else if (ctx->cfg.kind == B200TRACK_HYBRIDSORT) {
        if (!d_feats) { set_error("HybridSORT: the embedding buffer is NULL"); return B200TRACK_ERR_ARG; }
        CU_TRY(b200::launch_hybridsort_step(p, ctx->variant, st));
    } else if (ctx->cfg.kind == B200TRACK_BOTSORT) {
        if (p.with_reid && !d_feats) { set_error("BoT-SORT with_reid: the embedding buffer is NULL"); return B200TRACK_ERR_ARG; }
        CU_TRY(b200::launch_botsort_step(p, ctx->variant, st, ctx->cfg.camera_motion != 0));
    } else CU_TRY(b200::launch_bytetrack_step(p, ctx->kf_kind, ctx->variant, st));
    ctx->launches += 1;
    return 0;
}

extern "C" int b200track_step(b200track_ctx* ctx, const double* d_dets, const int32_t* d_ndets,
                              const float* d_feats, int32_t img_h, int32_t img_w,
                              double* d_out, int32_t* d_nout, void* stream) {
    if (!ctx || !d_dets || !d_ndets || !d_out || !d_nout) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    // detection rows are fetched and output rows stored with 16-byte accesses
    if ((reinterpret_cast<uintptr_t>(d_dets) | reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(d_feats)) & 15) {
        set_error("d_dets / d_out / d_feats must be 16-byte aligned"); return B200TRACK_ERR_ARG; }
    ON_DEVICE(ctx);
    return launch_step(ctx, d_dets, d_ndets, d_feats, img_h, img_w, d_out, d_nout, (cudaStream_t)stream);
}

extern "C" int b200track_step_cam(b200track_ctx* ctx, const double* d_dets, const int32_t* d_ndets,
                                  const float* d_feats, const double* d_warps, int32_t img_h, int32_t img_w,
                                  double* d_out, int32_t* d_nout, void* stream) {
    if (!ctx || !d_dets || !d_ndets || !d_out || !d_nout) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    if ((reinterpret_cast<uintptr_t>(d_dets) | reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(d_feats)) & 15) {
        set_error("d_dets / d_out / d_feats must be 16-byte aligned"); return B200TRACK_ERR_ARG; }
    ON_DEVICE(ctx);
    return launch_step(ctx, d_dets, d_ndets, d_feats, img_h, img_w, d_out, d_nout, (cudaStream_t)stream, nullptr, d_warps);
}

extern "C" int b200track_host_slots(b200track_ctx* ctx) { return ctx ? NSLOT : B200TRACK_ERR_ARG; }

static std::string capacity_message(int e) {
    return std::string("capacity overflow:") + ((e & B200_ERR_DET_OVERFLOW) ? " detections > max_dets" : "") +
           ((e & B200_ERR_TRACK_OVERFLOW) ? " tracks > max_tracks" : "") +
           ((e & B200_ERR_BOT_CAPACITY) ? " BoT-SORT candidate graph or class history (> 4 classes on a track)" : "") +
           ((e & B200_ERR_LSA) ? " StrongSORT: cost matrix contains invalid numeric entries" : "") +
           ((e & B200_ERR_PIPELINE) ? " HybridSORT: bulk-copy pipeline protocol error in the cosine pass" : "") +
           ((e & B200_ERR_PACKED_ROW) ? " OC-SORT exception area of the result block is full (rows that report the filter's box)" : "") +
           "; the context's state is truncated - b200track_reset before reuse";
}

static int ensure_padded_slot(b200track_ctx* ctx, HostSlot& s) {
    if (s.d_dets) return 0;
    const size_t S = ctx->cfg.n_streams, D = ctx->cfg.max_dets;
    CU_TRY(cudaMalloc(&s.d_dets, S * D * 6 * sizeof(double)));
    CU_TRY(cudaMalloc(&s.d_ndets, S * sizeof(int32_t)));
    if (ctx->p.feat_dim > 0) CU_TRY(cudaMalloc(&s.d_feats, S * D * (size_t)ctx->p.feat_dim * sizeof(float)));
    CU_TRY(cudaMalloc(&s.d_out, S * (size_t)ctx->cfg.max_tracks * 8 * sizeof(double)));
    CU_TRY(cudaMalloc(&s.d_nout, S * sizeof(int32_t)));
    return 0;
}

extern "C" int b200track_submit_host(b200track_ctx* ctx, int32_t slot, const double* h_dets,
                                     const int32_t* h_ndets, const float* h_feats, int32_t img_h,
                                     int32_t img_w, double* h_out, int32_t* h_nout) {
    if (!ctx || !h_dets || !h_ndets || !h_out || !h_nout) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    if (slot < 0 || slot >= NSLOT) { set_error("slot out of range"); return B200TRACK_ERR_ARG; }
    if (((ctx->cfg.kind == B200TRACK_BOTSORT && ctx->cfg.with_reid) || (ctx->cfg.kind == B200TRACK_DEEPOCSORT && !ctx->p.embedding_off) ||
         ctx->cfg.kind == B200TRACK_STRONGSORT || ctx->cfg.kind == B200TRACK_HYBRIDSORT) && !h_feats) {
        set_error("this context needs embeddings: h_feats is NULL"); return B200TRACK_ERR_ARG; }
    ON_DEVICE(ctx);
    HostSlot& s = ctx->slot[slot];
    if (int rc = ensure_padded_slot(ctx, s)) return rc;
    const size_t S = ctx->cfg.n_streams, T = ctx->cfg.max_tracks, D = ctx->cfg.max_dets;
    if (s.used) {
        CU_TRY(cudaStreamWaitEvent(ctx->s_h2d, s.done, 0));        // previous step on this slot consumed its inputs
        CU_TRY(cudaStreamWaitEvent(ctx->s_compute, s.out_ready, 0)); // ... and its outputs have been read back
    }
    // only the rows any stream actually uses are copied (pitch = one stream's padded block)
    int maxnd = 0;
    for (size_t i = 0; i < S; ++i) maxnd = h_ndets[i] > maxnd ? h_ndets[i] : maxnd;
    if (maxnd > (int)D) maxnd = (int)D;
    CU_TRY(cudaMemcpyAsync(s.d_ndets, h_ndets, S * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_h2d));
    if (maxnd > 0)
        CU_TRY(cudaMemcpy2DAsync(s.d_dets, D * 6 * sizeof(double), h_dets, D * 6 * sizeof(double),
                                 (size_t)maxnd * 6 * sizeof(double), S, cudaMemcpyHostToDevice, ctx->s_h2d));
    if (ctx->p.feat_dim > 0 && h_feats && maxnd > 0) {
        const size_t row = (size_t)ctx->p.feat_dim * sizeof(float);
        CU_TRY(cudaMemcpy2DAsync(s.d_feats, D * row, h_feats, D * row, maxnd * row, S, cudaMemcpyHostToDevice, ctx->s_h2d));
    }
    CU_TRY(cudaEventRecord(s.in_ready, ctx->s_h2d));
    CU_TRY(cudaStreamWaitEvent(ctx->s_compute, s.in_ready, 0));
    // capacity overflows of THIS step come back with its outputs (b200track_wait_host reports them)
    CU_TRY(cudaMemsetAsync(ctx->p.err_slot + slot, 0, sizeof(int), ctx->s_compute));
    if (int rc = launch_step(ctx, s.d_dets, s.d_ndets, ctx->p.feat_dim > 0 && h_feats ? s.d_feats : nullptr, img_h, img_w,
                             s.d_out, s.d_nout, ctx->s_compute, ctx->p.err_slot + slot)) return rc;
    CU_TRY(cudaEventRecord(s.done, ctx->s_compute));
    CU_TRY(cudaStreamWaitEvent(ctx->s_d2h, s.done, 0));
    CU_TRY(cudaMemcpyAsync(h_nout, s.d_nout, S * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_d2h));
    CU_TRY(cudaMemcpyAsync(s.h_err, ctx->p.err_slot + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_d2h));
    // every output row carries a distinct detection of this frame (a track is only listed as tracked in the frame it
    // was matched or born), so at most maxnd rows per stream are valid: copy that prefix of each stream's block
    if (maxnd > 0)
        CU_TRY(cudaMemcpy2DAsync(h_out, T * 8 * sizeof(double), s.d_out, T * 8 * sizeof(double),
                                 (size_t)(maxnd < (int)T ? maxnd : (int)T) * 8 * sizeof(double), S, cudaMemcpyDeviceToHost, ctx->s_d2h));
    CU_TRY(cudaEventRecord(s.out_ready, ctx->s_d2h));
    s.used = true; s.packed = false;
    return 0;
}

static int finish_slot(HostSlot& s) {
    if (!s.used) return 0;
    CU_TRY(cudaEventSynchronize(s.out_ready));
    const int e = s.packed ? (s.h_res ? s.h_res[0] : 0) : *s.h_err;
    if (e) { set_error(capacity_message(e)); return B200TRACK_ERR_CAPACITY; }
    return 0;
}

extern "C" int b200track_wait_host(b200track_ctx* ctx, int32_t slot) {
    if (!ctx || slot < 0 || slot >= NSLOT) { set_error("bad ctx / slot"); return B200TRACK_ERR_ARG; }
    return finish_slot(ctx->slot[slot]);
}

extern "C" int b200track_step_host(b200track_ctx* ctx, const double* h_dets, const int32_t* h_ndets,
                                   const float* h_feats, int32_t img_h, int32_t img_w,
                                   double* h_out, int32_t* h_nout) {
    if (int rc = b200track_submit_host(ctx, 0, h_dets, h_ndets, h_feats, img_h, img_w, h_out, h_nout)) return rc;
    return b200track_wait_host(ctx, 0);
}

// ---- packed frames ------------------------------------------------------------------------------------------------
static inline uint64_t align16(uint64_t v) { return (v + 15) & ~(uint64_t)15; }

static int row_bytes_of(const b200track_ctx* ctx) {
    return ctx->cfg.kind == B200TRACK_OCSORT ? B200_ROW_OC : (ctx->cfg.kind == B200TRACK_BOTSORT ? B200_ROW_BOT : B200_ROW_BYTE);   // DeepOCSORT: 40-byte rows
}

extern "C" int b200track_frame_layout(b200track_ctx* ctx, int64_t n_rows, int32_t det_dtype, b200track_layout* out) {
    if (!ctx || !out) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    if (ctx->cfg.kind == B200TRACK_STRONGSORT || ctx->cfg.kind == B200TRACK_HYBRIDSORT) {
        set_error("StrongSORT / HybridSORT contexts use the padded interface (b200track_step / _submit_host)"); return B200TRACK_ERR_STATE; }
    if (n_rows < 0 || (det_dtype != B200TRACK_F32 && det_dtype != B200TRACK_F64)) { set_error("bad n_rows / det_dtype"); return B200TRACK_ERR_ARG; }
    const uint64_t S = ctx->cfg.n_streams, R = (uint64_t)n_rows;
    const uint64_t det_row = det_dtype == B200TRACK_F32 ? 24 : 48;
    const bool feats = (ctx->cfg.kind == B200TRACK_BOTSORT && ctx->cfg.with_reid) || (ctx->cfg.kind == B200TRACK_DEEPOCSORT && !ctx->p.embedding_off);
    const bool warps = ctx->cfg.kind == B200TRACK_BOTSORT || ctx->cfg.kind == B200TRACK_DEEPOCSORT;
    out->in_off_offsets = 0;
    out->in_off_warps = align16(4 * (S + 1));
    out->in_off_dets = align16(out->in_off_warps + (warps ? 48 * S : 0));
    out->in_off_feats = align16(out->in_off_dets + det_row * R);
    out->in_bytes = align16(out->in_off_feats + (feats ? R * (uint64_t)ctx->p.feat_dim * 4 : 0));
    out->row_bytes = row_bytes_of(ctx);
    out->out_off_nout = 16;
    out->out_off_rows = align16(16 + 4 * S);
    out->out_off_exc = align16(out->out_off_rows + R * (uint64_t)out->row_bytes);
    out->exc_capacity = ctx->cfg.kind == B200TRACK_OCSORT ? (int32_t)(64 + R / 128) : 0;
    out->out_bytes = align16(out->out_off_exc + (uint64_t)out->exc_capacity * B200_EXC_OC_BYTES);
    return 0;
}

// validates the offsets of a host-side input block and returns its row count
static int64_t packed_rows(const b200track_ctx* ctx, const int32_t* off) {
    const int S = ctx->cfg.n_streams;
    if (off[0] != 0) return -1;
    for (int i = 0; i < S; ++i) if (off[i + 1] < off[i]) return -1;
    if ((int64_t)off[S] > (int64_t)S * ctx->cfg.max_dets) return -2;
    return off[S];
}

static int launch_packed(b200track_ctx* ctx, const unsigned char* d_in, int64_t R, int32_t det_dtype, int32_t flags,
                         int32_t img_h, int32_t img_w, unsigned char* d_res, cudaStream_t st) {
    b200track_layout L;
    if (int rc = b200track_frame_layout(ctx, R, det_dtype, &L)) return rc;
    b200::StepParams p = ctx->p;
    p.det_off = reinterpret_cast<const int*>(d_in + L.in_off_offsets);
    p.dets32 = det_dtype == B200TRACK_F32 ? reinterpret_cast<const float*>(d_in + L.in_off_dets) : nullptr;
    p.dets = det_dtype == B200TRACK_F64 ? reinterpret_cast<const double*>(d_in + L.in_off_dets) : nullptr;
    p.ndets = nullptr;
    const bool has_feats = (ctx->cfg.kind == B200TRACK_BOTSORT && ctx->cfg.with_reid) || (ctx->cfg.kind == B200TRACK_DEEPOCSORT && !ctx->p.embedding_off);
    p.feats = has_feats ? reinterpret_cast<const float*>(d_in + L.in_off_feats) : nullptr;
    p.warps = (flags & B200TRACK_FRAME_HAS_WARPS) ? reinterpret_cast<const double*>(d_in + L.in_off_warps) : nullptr;
    if (p.warps && !(ctx->cfg.kind == B200TRACK_DEEPOCSORT || (ctx->cfg.kind == B200TRACK_BOTSORT && ctx->cfg.camera_motion))) {
        set_error("camera-motion warps need a DeepOCSORT context or a BoT-SORT context created with camera_motion"); return B200TRACK_ERR_STATE; }
    p.out = nullptr;
    p.nout = reinterpret_cast<int*>(d_res + L.out_off_nout);
    p.rows = d_res + L.out_off_rows;
    p.err_out = reinterpret_cast<int*>(d_res);
    p.exc = d_res + L.out_off_exc;
    p.exc_cap = L.exc_capacity;
    p.img_h = img_h; p.img_w = img_w;
    CU_TRY(cudaMemsetAsync(d_res, 0, 16, st));                  // header: [0] capacity bits of this step
    if (ctx->cfg.kind == B200TRACK_OCSORT) CU_TRY(b200::launch_ocsort_step(p, ctx->variant, st));
    else if (ctx->cfg.kind == B200TRACK_DEEPOCSORT) CU_TRY(b200::launch_deepocsort_step(p, ctx->variant, st));
    else if (ctx->cfg.kind == B200TRACK_BOTSORT) CU_TRY(b200::launch_botsort_step_packed(p, ctx->variant, st, ctx->cfg.camera_motion != 0));
    else CU_TRY(b200::launch_bytetrack_step_packed(p, ctx->kf_kind, ctx->variant, st));
    ctx->launches += 1;
    return 0;
}

extern "C" int b200track_step_packed(b200track_ctx* ctx, const void* d_in, int64_t n_rows, int32_t det_dtype, int32_t flags,
                                     int32_t img_h, int32_t img_w, void* d_result, void* stream) {
    if (!ctx || !d_in || !d_result) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    if ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_result)) & 15) { set_error("blocks must be 16-byte aligned"); return B200TRACK_ERR_ARG; }
    if (n_rows < 0 || n_rows > (int64_t)ctx->cfg.n_streams * ctx->cfg.max_dets) { set_error("n_rows out of range"); return B200TRACK_ERR_ARG; }
    ON_DEVICE(ctx);
    return launch_packed(ctx, static_cast<const unsigned char*>(d_in), n_rows, det_dtype, flags, img_h, img_w,
                         static_cast<unsigned char*>(d_result), (cudaStream_t)stream);
}

extern "C" int b200track_submit_packed(b200track_ctx* ctx, int32_t slot, const void* h_in, int32_t det_dtype, int32_t flags,
                                       int32_t img_h, int32_t img_w, void* h_result) {
    if (!ctx || !h_in || !h_result) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    if (slot < 0 || slot >= NSLOT) { set_error("slot out of range"); return B200TRACK_ERR_ARG; }
    const int64_t R = packed_rows(ctx, static_cast<const int32_t*>(h_in));
    if (R == -1) { set_error("offsets must start at 0 and be non-decreasing"); return B200TRACK_ERR_ARG; }
    if (R == -2) { set_error("more detection rows than n_streams * max_dets"); return B200TRACK_ERR_CAPACITY; }
    ON_DEVICE(ctx);
    HostSlot& s = ctx->slot[slot];
    b200track_layout L, Lmax;
    if (int rc = b200track_frame_layout(ctx, R, det_dtype, &L)) return rc;
    if (!s.d_in) {
        if (int rc = b200track_frame_layout(ctx, (int64_t)ctx->cfg.n_streams * ctx->cfg.max_dets, B200TRACK_F64, &Lmax)) return rc;
        CU_TRY(cudaMalloc(&s.d_in, Lmax.in_bytes));
        CU_TRY(cudaMalloc(&s.d_res, Lmax.out_bytes));
    }
    if (s.used) {
        CU_TRY(cudaStreamWaitEvent(ctx->s_h2d, s.done, 0));
        CU_TRY(cudaStreamWaitEvent(ctx->s_compute, s.out_ready, 0));
    }
    // ONE linear copy per direction: the input block in, the result block (header, counts, compact rows) out
    CU_TRY(cudaMemcpyAsync(s.d_in, h_in, L.in_bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
    CU_TRY(cudaEventRecord(s.in_ready, ctx->s_h2d));
    CU_TRY(cudaStreamWaitEvent(ctx->s_compute, s.in_ready, 0));
    if (int rc = launch_packed(ctx, s.d_in, R, det_dtype, flags, img_h, img_w, s.d_res, ctx->s_compute)) return rc;
    CU_TRY(cudaEventRecord(s.done, ctx->s_compute));
    CU_TRY(cudaStreamWaitEvent(ctx->s_d2h, s.done, 0));
    CU_TRY(cudaMemcpyAsync(h_result, s.d_res, L.out_bytes, cudaMemcpyDeviceToHost, ctx->s_d2h));
    CU_TRY(cudaEventRecord(s.out_ready, ctx->s_d2h));
    s.used = true; s.packed = true; s.h_res = static_cast<const int32_t*>(h_result);
    return 0;
}

extern "C" int b200track_wait_packed(b200track_ctx* ctx, int32_t slot) {
    if (!ctx || slot < 0 || slot >= NSLOT) { set_error("bad ctx / slot"); return B200TRACK_ERR_ARG; }
    return finish_slot(ctx->slot[slot]);
}

extern "C" int b200track_sync(b200track_ctx* ctx) {
    if (!ctx) { set_error("ctx is NULL"); return B200TRACK_ERR_ARG; }
    ON_DEVICE(ctx);
    CU_TRY(cudaDeviceSynchronize());
    if (ctx->cfg.kind == B200TRACK_STRONGSORT) {
        unsigned long long g[3];
        CU_TRY(cudaMemcpy(g, ctx->ss.gstats, sizeof(g), cudaMemcpyDeviceToHost));
        if (g[1]) { set_error("StrongSORT: tensor-core pipeline protocol error in the gallery distance"); return B200TRACK_ERR_CUDA; }
    }
    CU_TRY(cudaMemcpy(ctx->h_err, ctx->p.err, sizeof(int), cudaMemcpyDeviceToHost));
    const int e = *ctx->h_err;
    if (e) {
        CU_TRY(cudaMemset(ctx->p.err, 0, sizeof(int)));
        set_error(capacity_message(e));
        return B200TRACK_ERR_CAPACITY;
    }
    return 0;
}

extern "C" int b200track_track_updates(b200track_ctx* ctx, uint64_t* h_total) {
    if (!ctx || !h_total) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    ON_DEVICE(ctx);
    CU_TRY(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(ctx->cfg.n_streams);
    CU_TRY(cudaMemcpy(h.data(), ctx->p.track_updates, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    uint64_t tot = 0;
    for (auto v : h) tot += v;
    *h_total = tot;
    return 0;
}

extern "C" int b200track_launch_count(b200track_ctx* ctx, uint64_t* h_launches) {
    if (!ctx || !h_launches) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    *h_launches = ctx->launches;
    return 0;
}

extern "C" int b200track_phase_cycles(b200track_ctx* ctx, uint64_t* h_out16, int32_t reset) {
    if (!ctx) { set_error("ctx is NULL"); return B200TRACK_ERR_ARG; }
    ON_DEVICE(ctx);
    CU_TRY(cudaDeviceSynchronize());
    if (!ctx->p.dbg) {
        CU_TRY(cudaMalloc(&ctx->p.dbg, 16 * sizeof(unsigned long long)));
        CU_TRY(cudaMemset(ctx->p.dbg, 0, 16 * sizeof(unsigned long long)));
    }
    if (h_out16) CU_TRY(cudaMemcpy(h_out16, ctx->p.dbg, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (reset) CU_TRY(cudaMemset(ctx->p.dbg, 0, 16 * sizeof(unsigned long long)));
    return 0;
}

extern "C" int b200track_footprint(b200track_ctx* ctx, uint64_t* h_state, uint64_t* h_smem) {
    if (!ctx) { set_error("ctx is NULL"); return B200TRACK_ERR_ARG; }
    if (ctx->cfg.kind == B200TRACK_STRONGSORT) {
        const uint64_t F = ctx->ss.F, G = ctx->ss.budget;
        if (h_state) *h_state = (uint64_t)ctx->tcap * (72 * 8 + 2 * 8 + b200::SS_NI * 4 + 4 + F * 4 + G * F * 6) + 4 * sizeof(int) + sizeof(unsigned long long);
        if (h_smem) *h_smem = b200::strongsort_match_smem();
        return 0;
    }
    if (h_state) *h_state = (uint64_t)ctx->tcap * (ctx->nf * 8 + ctx->ni * 4) + 4 * sizeof(int) + sizeof(unsigned long long);
    if (h_smem) *h_smem = ctx->cfg.kind == B200TRACK_OCSORT ? b200::ocsort_step_smem(ctx->variant)
                          : ctx->cfg.kind == B200TRACK_DEEPOCSORT ? b200::deepocsort_step_smem(ctx->variant)
                          : ctx->cfg.kind == B200TRACK_HYBRIDSORT ? b200::hybridsort_step_smem(ctx->variant)
                          : b200::bytetrack_step_smem(ctx->variant, ctx->cfg.kind == B200TRACK_BOTSORT, ctx->cfg.kind == B200TRACK_BOTSORT && ctx->cfg.camera_motion);
    return 0;
}

extern "C" int b200track_get_state(b200track_ctx* ctx, int32_t stream_index, int32_t* h_counts,
                                   int32_t* h_rec, double* h_mean, double* h_cov, double* h_aux) {
    if (!ctx || !h_counts) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    if (stream_index < 0 || stream_index >= ctx->cfg.n_streams) { set_error("stream_index out of range"); return B200TRACK_ERR_ARG; }
    ON_DEVICE(ctx);
    CU_TRY(cudaDeviceSynchronize());
    const size_t T = ctx->tcap, s = stream_index;
    if (ctx->cfg.kind == B200TRACK_HYBRIDSORT) { set_error("HybridSORT contexts are read with b200track_get_state_hybridsort"); return B200TRACK_ERR_STATE; }
    if (ctx->cfg.kind == B200TRACK_STRONGSORT) {
        const b200::SSParams& q = ctx->ss;
        int c4[4];
        CU_TRY(cudaMemcpy(c4, ctx->p.counts + 4 * s, sizeof(c4), cudaMemcpyDeviceToHost));
        std::vector<int> iv((size_t)b200::SS_NI * T), order(T);
        std::vector<double> mean(T * 8), cov(T * 64), conf(T), cls(T);
        CU_TRY(cudaMemcpy(iv.data(), q.ti + s * b200::SS_NI * T, iv.size() * sizeof(int), cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(order.data(), q.order + s * T, T * sizeof(int), cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(mean.data(), q.mean + s * T * 8, mean.size() * sizeof(double), cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(cov.data(), q.cov + s * T * 64, cov.size() * sizeof(double), cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(conf.data(), q.conf + s * T, T * sizeof(double), cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(cls.data(), q.cls + s * T, T * sizeof(double), cudaMemcpyDeviceToHost));
        const int n = c4[0];
        for (int k = 0; k < n && k < (int)T; ++k) {
            const int t = order[k];
            if (h_rec) {
                int32_t* r = h_rec + 6 * k;
                r[0] = iv[b200::SSI_ID * T + t]; r[1] = iv[b200::SSI_STATE * T + t]; r[2] = iv[b200::SSI_HITS * T + t];
                r[3] = iv[b200::SSI_AGE * T + t]; r[4] = iv[b200::SSI_TSU * T + t];
                r[5] = iv[b200::SSI_STATE * T + t] == b200::SS_CONFIRMED ? std::min(iv[b200::SSI_APPENDED * T + t], q.budget) : 0;
            }
            if (h_mean) for (int c = 0; c < 8; ++c) h_mean[8 * k + c] = mean[(size_t)t * 8 + c];
            if (h_cov) for (int c = 0; c < 64; ++c) h_cov[64 * k + c] = cov[(size_t)t * 64 + c];
            if (h_aux) { h_aux[3 * k] = conf[t]; h_aux[3 * k + 1] = cls[t]; h_aux[3 * k + 2] = (double)iv[b200::SSI_DET * T + t]; }
        }
        h_counts[0] = n; h_counts[1] = 0; h_counts[2] = c4[1]; h_counts[3] = c4[2];
        return 0;
    }
    std::vector<double> f((size_t)ctx->nf * T);
    std::vector<int> iv((size_t)ctx->ni * T);
    CU_TRY(cudaMemcpy(h_counts, ctx->p.counts + 4 * s, 4 * sizeof(int), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(f.data(), ctx->p.state_f + s * ctx->nf * T, f.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(iv.data(), ctx->p.state_i + s * ctx->ni * T, iv.size() * sizeof(int), cudaMemcpyDeviceToHost));
    if (ctx->cfg.kind == B200TRACK_DEEPOCSORT) {
        // alive slots in list order; h_rec = id (1-based like KalmanBoxTracker.count), age, time_since_update, hits, hit_streak,
        // observed + 2 * frozen; h_mean = x[8]; h_cov = dense 8x8 P from the two 4x4 groups; h_aux = conf, cls, det_ind
        static const int gidx[2][4] = {{0, 1, 4, 5}, {2, 3, 6, 7}};
        int k = 0;
        for (int t = 0; t < h_counts[0] && t < ctx->cfg.max_tracks; ++t) {
            const int fl = iv[B200_OCI_FLAGS * T + t];
            if (!(fl & 8)) continue;
            if (h_rec) {
                int32_t* r = h_rec + 6 * k;
                r[0] = iv[B200_OCI_ID * T + t] + 1; r[1] = iv[B200_OCI_AGE * T + t]; r[2] = iv[B200_OCI_TSU * T + t];
                r[3] = iv[B200_OCI_HITS * T + t]; r[4] = iv[B200_OCI_STREAK * T + t];
                r[5] = (fl & B200_OCF_OBSERVED) | ((fl & B200_DOF_FROZEN) ? 2 : 0);
            }
            if (h_mean) for (int c = 0; c < 8; ++c) h_mean[8 * k + c] = f[(B200_DO_X + c) * T + t];
            if (h_cov) {
                double* c = h_cov + 64 * k;
                for (int q = 0; q < 64; ++q) c[q] = 0.0;
                for (int g = 0; g < 2; ++g) {
                    int q = 0;
                    for (int a = 0; a < 4; ++a)
                        for (int b = a; b < 4; ++b) {
                            const double v = f[((g ? B200_DO_PB : B200_DO_PA) + q) * T + t];
                            c[gidx[g][a] * 8 + gidx[g][b]] = v; c[gidx[g][b] * 8 + gidx[g][a]] = v;
                            ++q;
                        }
                }
            }
            if (h_aux) {
                h_aux[3 * k + 0] = f[B200_DO_CONF * T + t];
                h_aux[3 * k + 1] = f[B200_DO_CLS * T + t];
                h_aux[3 * k + 2] = (double)iv[B200_OCI_DET * T + t];
            }
            ++k;
        }
        h_counts[0] = k; h_counts[1] = 0;
        return 0;
    }
    if (ctx->cfg.kind == B200TRACK_OCSORT) {
        // alive slots in list order; h_rec = id, age, time_since_update, hits, hit_streak, observed;
        // h_mean[8] = x[7], has-observation flag; h_cov = dense 7x7 P in the first 49 entries;
        // h_aux = conf, cls, det_ind
        int k = 0;
        for (int t = 0; t < h_counts[0] && t < ctx->cfg.max_tracks; ++t) {
            const int fl = iv[B200_OCI_FLAGS * T + t];
            if (!(fl & 8)) continue;
            if (h_rec) {
                int32_t* r = h_rec + 6 * k;
                r[0] = iv[B200_OCI_ID * T + t]; r[1] = iv[B200_OCI_AGE * T + t]; r[2] = iv[B200_OCI_TSU * T + t];
                r[3] = iv[B200_OCI_HITS * T + t]; r[4] = iv[B200_OCI_STREAK * T + t]; r[5] = fl & B200_OCF_OBSERVED;
            }
            if (h_mean) {
                for (int c = 0; c < 7; ++c) h_mean[8 * k + c] = f[(B200_OC_X + c) * T + t];
                h_mean[8 * k + 7] = (fl & B200_OCF_HASOBS) ? 1.0 : 0.0;
            }
            if (h_cov) {
                double* c = h_cov + 64 * k;
                for (int q = 0; q < 64; ++q) c[q] = 0.0;
                for (int a = 0; a < 3; ++a) {
                    c[a * 7 + a] = f[(B200_OC_P + 3 * a + 0) * T + t];
                    c[a * 7 + a + 4] = c[(a + 4) * 7 + a] = f[(B200_OC_P + 3 * a + 1) * T + t];
                    c[(a + 4) * 7 + a + 4] = f[(B200_OC_P + 3 * a + 2) * T + t];
                }
                c[3 * 7 + 3] = f[(B200_OC_P + 9) * T + t];
                // velocity and last observation ride in the unused tail of the 64-entry row
                c[49] = f[(B200_OC_VEL + 0) * T + t]; c[50] = f[(B200_OC_VEL + 1) * T + t];
                for (int q = 0; q < 4; ++q) c[51 + q] = (fl & B200_OCF_HASOBS) ? f[(B200_OC_LAST + q) * T + t] : -1.0;
                c[55] = (fl & B200_OCF_HASOBS) ? f[B200_OC_CONF * T + t] : -1.0;
            }
            if (h_aux) {
                h_aux[3 * k + 0] = f[B200_OC_CONF * T + t];
                h_aux[3 * k + 1] = f[B200_OC_CLS * T + t];
                h_aux[3 * k + 2] = (double)iv[B200_OCI_DET * T + t];
            }
            ++k;
        }
        h_counts[0] = k; h_counts[1] = 0;
        return 0;
    }
    const int n = h_counts[0] + h_counts[1];
    for (int t = 0; t < n && t < ctx->cfg.max_tracks; ++t) {
        const int fl = iv[B200_TI_FLAGS * T + t];
        if (h_rec) {
            int32_t* r = h_rec + 6 * t;
            r[0] = iv[B200_TI_ID * T + t]; r[1] = fl & 3; r[2] = (fl & B200_FLAG_ACTIVATED) ? 1 : 0;
            if (ctx->cfg.kind == B200TRACK_BOTSORT && r[1] == 3) r[1] = 4;      // botsort/basetrack.py:7-12: Removed = 4
            r[3] = iv[B200_TI_FRAME * T + t]; r[4] = iv[B200_TI_START * T + t]; r[5] = iv[B200_TI_LEN * T + t];
        }
        if (h_mean) for (int c = 0; c < 8; ++c) h_mean[8 * t + c] = f[(B200_TF_MEAN + c) * T + t];
        const bool cam = ctx->nf == B200_NF_CAM;
        if (h_cov) {
            double* c = h_cov + 64 * t;
            for (int k = 0; k < 64; ++k) c[k] = 0.0;
            if (cam) {
                static const int gidx[2][4] = {{0, 1, 4, 5}, {2, 3, 6, 7}};
                for (int g = 0; g < 2; ++g) {
                    int q = 0;
                    for (int a = 0; a < 4; ++a)
                        for (int b = a; b < 4; ++b) {
                            const double v = f[((g ? B200_TFC_COVB : B200_TFC_COVA) + q) * T + t];
                            c[gidx[g][a] * 8 + gidx[g][b]] = v; c[gidx[g][b] * 8 + gidx[g][a]] = v;
                            ++q;
                        }
                }
            } else
            for (int a = 0; a < 4; ++a) {
                c[a * 8 + a] = f[(B200_TF_COV + 3 * a + 0) * T + t];
                c[a * 8 + a + 4] = c[(a + 4) * 8 + a] = f[(B200_TF_COV + 3 * a + 1) * T + t];
                c[(a + 4) * 8 + a + 4] = f[(B200_TF_COV + 3 * a + 2) * T + t];
            }
        }
        if (h_aux) {
            h_aux[3 * t + 0] = f[(cam ? B200_TFC_SCORE : B200_TF_SCORE) * T + t];
            h_aux[3 * t + 1] = f[(cam ? B200_TFC_CLS : B200_TF_CLS) * T + t];
            h_aux[3 * t + 2] = (double)iv[B200_TI_DET * T + t];
        }
    }
    return 0;
}

extern "C" int b200track_get_features(b200track_ctx* ctx, int32_t stream_index, float* h_feat) {
    if (!ctx || !h_feat) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    if (stream_index < 0 || stream_index >= ctx->cfg.n_streams) { set_error("stream_index out of range"); return B200TRACK_ERR_ARG; }
    if (ctx->cfg.kind == B200TRACK_STRONGSORT) {
        ON_DEVICE(ctx);
        CU_TRY(cudaDeviceSynchronize());
        const b200::SSParams& q = ctx->ss;
        const size_t T = ctx->tcap, s = stream_index, F = q.F;
        int c4[4];
        CU_TRY(cudaMemcpy(c4, ctx->p.counts + 4 * s, sizeof(c4), cudaMemcpyDeviceToHost));
        std::vector<int> order(T);
        CU_TRY(cudaMemcpy(order.data(), q.order + s * T, T * sizeof(int), cudaMemcpyDeviceToHost));
        for (int k = 0; k < c4[0] && k < (int)T; ++k)
            CU_TRY(cudaMemcpy(h_feat + (size_t)k * F, q.feat + (s * T + order[k]) * F, F * sizeof(float), cudaMemcpyDeviceToHost));
        return 0;
    }
    if (ctx->cfg.kind == B200TRACK_HYBRIDSORT) {
        ON_DEVICE(ctx);
        CU_TRY(cudaDeviceSynchronize());
        const size_t T = ctx->tcap, s = stream_index, F = ctx->cfg.feat_dim;
        int counts[4];
        CU_TRY(cudaMemcpy(counts, ctx->p.counts + 4 * s, sizeof(counts), cudaMemcpyDeviceToHost));
        std::vector<int> iv((size_t)ctx->ni * T);
        std::vector<float> pool(T * F);                              // the stream's whole pool in one copy, rows picked on the host
        CU_TRY(cudaMemcpy(iv.data(), ctx->p.state_i + s * ctx->ni * T, iv.size() * sizeof(int), cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(pool.data(), ctx->p.feat_pool + s * T * F, pool.size() * sizeof(float), cudaMemcpyDeviceToHost));
        int k = 0;
        for (int t = 0; t < counts[0] && t < ctx->cfg.max_tracks; ++t) {
            if (!(iv[B200_OCI_FLAGS * T + t] & 8)) continue;
            memcpy(h_feat + (size_t)k * F, pool.data() + (size_t)iv[B200_HYI_FROW * T + t] * F, F * sizeof(float));
            ++k;
        }
        return 0;
    }
    if (ctx->cfg.kind != B200TRACK_BOTSORT || !ctx->p.feat_pool) { set_error("context holds no embeddings"); return B200TRACK_ERR_STATE; }
    ON_DEVICE(ctx);
    CU_TRY(cudaDeviceSynchronize());
    const size_t T = ctx->tcap, s = stream_index, F = ctx->cfg.feat_dim;
    int counts[4];
    CU_TRY(cudaMemcpy(counts, ctx->p.counts + 4 * s, sizeof(counts), cudaMemcpyDeviceToHost));
    std::vector<int> rows(T);
    CU_TRY(cudaMemcpy(rows.data(), ctx->p.state_i + (s * ctx->ni + B200_TI_FROW) * T, T * sizeof(int), cudaMemcpyDeviceToHost));
    const int n = counts[0] + counts[1];
    for (int t = 0; t < n && t < ctx->cfg.max_tracks; ++t)
        CU_TRY(cudaMemcpy(h_feat + (size_t)t * F, ctx->p.feat_pool + (s * T + rows[t]) * F, F * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int b200track_get_track_extras(b200track_ctx* ctx, int32_t stream_index, double* h_extra, double* h_emb) {
    if (!ctx) { set_error("ctx is NULL"); return B200TRACK_ERR_ARG; }
    if (ctx->cfg.kind != B200TRACK_DEEPOCSORT) { set_error("not a DeepOCSORT context"); return B200TRACK_ERR_STATE; }
    if (stream_index < 0 || stream_index >= ctx->cfg.n_streams) { set_error("stream_index out of range"); return B200TRACK_ERR_ARG; }
    ON_DEVICE(ctx);
    CU_TRY(cudaDeviceSynchronize());
    const size_t T = ctx->tcap, s = stream_index, F = ctx->p.feat_dim;
    int counts[4];
    CU_TRY(cudaMemcpy(counts, ctx->p.counts + 4 * s, sizeof(counts), cudaMemcpyDeviceToHost));
    std::vector<double> f((size_t)ctx->nf * T);
    std::vector<int> iv((size_t)ctx->ni * T);
    CU_TRY(cudaMemcpy(f.data(), ctx->p.state_f + s * ctx->nf * T, f.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(iv.data(), ctx->p.state_i + s * ctx->ni * T, iv.size() * sizeof(int), cudaMemcpyDeviceToHost));
    int k = 0;
    for (int t = 0; t < counts[0] && t < ctx->cfg.max_tracks; ++t) {
        const int fl = iv[B200_OCI_FLAGS * T + t];
        if (!(fl & 8)) continue;
        if (h_extra) {
            double* e = h_extra + 8 * k;
            e[0] = f[(B200_DO_VEL + 0) * T + t]; e[1] = f[(B200_DO_VEL + 1) * T + t];
            const bool has = fl & B200_OCF_HASOBS;
            for (int q = 0; q < 4; ++q) e[2 + q] = has ? f[(B200_DO_LAST + q) * T + t] : -1.0;
            e[6] = has ? f[B200_DO_CONF * T + t] : -1.0;
            e[7] = 0.0;
        }
        if (h_emb && F > 0 && ctx->p.emb_pool)
            CU_TRY(cudaMemcpy(h_emb + (size_t)k * F, ctx->p.emb_pool + (s * T + iv[B200_DOI_EROW * T + t]) * F, F * sizeof(double), cudaMemcpyDeviceToHost));
        ++k;
    }
    return 0;
}

extern "C" int b200track_get_state_hybridsort(b200track_ctx* ctx, int32_t stream_index, int32_t* h_counts, int32_t* h_rec, double* h_x,
                                              double* h_P, double* h_vel, double* h_last, double* h_aux) {
    if (!ctx || !h_counts) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    if (ctx->cfg.kind != B200TRACK_HYBRIDSORT) { set_error("not a HybridSORT context"); return B200TRACK_ERR_STATE; }
    if (stream_index < 0 || stream_index >= ctx->cfg.n_streams) { set_error("stream_index out of range"); return B200TRACK_ERR_ARG; }
    ON_DEVICE(ctx);
    CU_TRY(cudaDeviceSynchronize());
    const size_t T = ctx->tcap, s = stream_index;
    std::vector<double> f((size_t)ctx->nf * T);
    std::vector<int> iv((size_t)ctx->ni * T);
    CU_TRY(cudaMemcpy(h_counts, ctx->p.counts + 4 * s, 4 * sizeof(int), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(f.data(), ctx->p.state_f + s * ctx->nf * T, f.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(iv.data(), ctx->p.state_i + s * ctx->ni * T, iv.size() * sizeof(int), cudaMemcpyDeviceToHost));
    int k = 0;
    for (int t = 0; t < h_counts[0] && t < ctx->cfg.max_tracks; ++t) {
        const int fl = iv[B200_OCI_FLAGS * T + t];
        if (!(fl & 8)) continue;
        const bool has = fl & B200_OCF_HASOBS;
        if (h_rec) {
            int32_t* r = h_rec + 6 * k;
            r[0] = iv[B200_OCI_ID * T + t]; r[1] = iv[B200_OCI_AGE * T + t]; r[2] = iv[B200_OCI_TSU * T + t];
            r[3] = iv[B200_OCI_HITS * T + t]; r[4] = iv[B200_OCI_STREAK * T + t]; r[5] = fl & B200_OCF_OBSERVED;
        }
        if (h_x) for (int c = 0; c < 9; ++c) h_x[9 * k + c] = f[(B200_HY_X + c) * T + t];
        if (h_P) {
            double* c = h_P + 81 * k;
            for (int q = 0; q < 81; ++q) c[q] = 0.0;
            for (int a = 0; a < 4; ++a) {
                c[a * 9 + a] = f[(B200_HY_P + 3 * a + 0) * T + t];
                c[a * 9 + a + 5] = c[(a + 5) * 9 + a] = f[(B200_HY_P + 3 * a + 1) * T + t];
                c[(a + 5) * 9 + a + 5] = f[(B200_HY_P + 3 * a + 2) * T + t];
            }
            c[4 * 9 + 4] = f[(B200_HY_P + 12) * T + t];
        }
        if (h_vel) for (int c = 0; c < 8; ++c) h_vel[8 * k + c] = f[(B200_HY_VEL + c) * T + t];
        if (h_last) {
            for (int c = 0; c < 4; ++c) h_last[5 * k + c] = has ? f[(B200_HY_LAST + c) * T + t] : -1.0;
            h_last[5 * k + 4] = has ? f[B200_HY_CONF * T + t] : -1.0;
        }
        if (h_aux) {
            h_aux[3 * k + 0] = f[B200_HY_CONF * T + t];
            h_aux[3 * k + 1] = f[B200_HY_CLS * T + t];
            h_aux[3 * k + 2] = (double)iv[B200_OCI_DET * T + t];
        }
        ++k;
    }
    h_counts[0] = k; h_counts[1] = 0;
    return 0;
}

extern "C" int b200track_counters(b200track_ctx* ctx, uint64_t* h_out8) {
    if (!ctx || !h_out8) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    ON_DEVICE(ctx);
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(h_out8, ctx->p.stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}

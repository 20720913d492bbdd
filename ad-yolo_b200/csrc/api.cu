// extern "C" surface of libadyolo_b200.so (declared in include/adyolo_b200.h).
#include <math.h>
#include <stddef.h>

#include "../../include/adyolo_b200.h"
#include "assign_host.h"
#include "common.cuh"
#include "frontend_core.cuh"
#include "frontend_host.h"

namespace ady {
int launch_scaler_partials(const float* feats, int B, int C, long long T, double* sum, double* sumsq,
                           double* maxv, double* minv, cudaStream_t stream);

struct OutStrides { long long sb, sc, st, sj; };
int launch_stft(const void* audio, int dtype, int B, long long N, float dc, float2* out, cudaStream_t stream);
int launch_logmel_from_stft(const float2* spec, int B, long long T, int C, int Cs, const float* mean, const float* istd,
                            float* out, OutStrides os, uint32_t* gmax_ws, float top_db, int apply_topdb, cudaStream_t stream);
int launch_iv_from_stft(const float2* spec, int B, long long T, const float* mean, const float* istd, float* out,
                        OutStrides os, int* flags, cudaStream_t stream);
int launch_spec_mask(float* feat, int B, int C, long long T, int F, const int* rects, int G, const int* bounds,
                     cudaStream_t stream);
int launch_gcc_from_stft(const float2* spec, int B, long long T, const float* mean, const float* istd, float* out,
                         OutStrides os, cudaStream_t stream);
int launch_gcc_from_phasors(const void* phasors, int B, long long T, const float* mean, const float* istd, float* out,
                            OutStrides os, cudaStream_t stream);

static int check_frontend_cfg(const adyolo_frontend_cfg* c) {
    if (!c) return set_error(ADY_ERR_INVALID, "frontend cfg is NULL");
    if (c->sr != 24000 || c->n_fft != NFFT || c->hop_length != HOP || c->win_length != NFFT ||
        c->mel_bins != NMEL || c->n_channels != NCH_IN)
        return set_error(ADY_ERR_UNSUPPORTED,
                         "front end is compiled for sr=24000 n_fft=win=1200 hop=600 mel=64 ch=4 "
                         "(hyp_data_DCASE20xx.yaml); got sr=%d n_fft=%d hop=%d win=%d mel=%d ch=%d",
                         c->sr, c->n_fft, c->hop_length, c->win_length, c->mel_bins, c->n_channels);
    return ADY_OK;
}

// grid constants with the FP32 rounding steps of loss.py:162-171 / float64 of datasets.py:219-235
static int make_cfgs(const adyolo_grid_cfg* g, AssignCfg* a, CellCfg* cc) {
    if (!g) return set_error(ADY_ERR_INVALID, "grid cfg is NULL");
    if (!(g->grid_size[0] > 0) || !(g->grid_size[1] > 0)) return set_error(ADY_ERR_INVALID, "grid_size must be > 0");
    const double gs0 = g->grid_size[0], gs1 = g->grid_size[1];
    int ga = (int)floor(360.0 / gs0), ge = (int)floor(180.0 / gs1);
    if (fmod(360.0, gs0) != 0.0) ga += 1;   // np.divmod remainder != 0
    if (fmod(180.0, gs1) != 0.0) ge += 1;
    if (ga > ADY_MAX_GRID || ge > ADY_MAX_GRID) return set_error(ADY_ERR_UNSUPPORTED, "grid %dx%d too large", ga, ge);
    if (a) {
        a->nb_classes = g->nb_classes; a->nb_anchors = g->nb_anchors; a->ga = ga; a->ge = ge; a->n_thr = g->n_thr;
        a->gs_u = (float)g->grid_size[0]; a->gs_v = (float)g->grid_size[1];   // torch.Tensor(grid_size): float32
        a->ovl_scale = (float)(0.5 + g->g_overlap);   // python float (0.5 + g_overlap) cast by the f32 tensor multiply
        for (int i = 0; i < ADY_MAX_GRID; ++i) {
            // torch: arange(int64)*grid_size(f32) - Tensor([180,90]) + grid_size*0.5, all float32 ops
            a->off_u[i] = ((float)i * a->gs_u - 180.0f) + a->gs_u * 0.5f;
            a->off_v[i] = ((float)i * a->gs_v - 90.0f) + a->gs_v * 0.5f;
        }
        for (int i = 0; i < ADY_MAX_THR; ++i) a->thr[i] = i < g->n_thr ? g->train_unify[i] : 0.f;
        a->deg2rad = (float)0.017453292519943295769236907684886;   // at::deg2rad M_PI_180
        a->rad2deg = (float)57.295779513082320876798154814105;     // at::rad2deg M_180_PI
        a->clip_lo = (float)(-1 + 1e-7);
        a->clip_hi = (float)(1 - 1e-7);
        a->gain_ang = g->angular_gain; a->gain_obj = g->object_gain;
        a->gain_nonobj = g->nonobj_gain; a->gain_cls = g->class_gain;
    }
    if (cc) {
        cc->ga = ga; cc->ge = ge;
        const double ov = g->g_overlap;
        for (int i = 0; i < ADY_MAX_GRID; ++i) {
            const double ca = i * gs0 - 180.0 + gs0 * 0.5, ce = i * gs1 - 90.0 + gs1 * 0.5;
            cc->lb_a[i] = ca - gs0 * (0.5 + ov);
            cc->ub_a[i] = ca + gs0 * (0.5 + ov);
            double l = ce - gs1 * (0.5 + ov), u = ce + gs1 * (0.5 + ov);
            cc->lb_e[i] = l < -90 ? -90 : (l > 90 ? 90 : l);   // np.clip(..., -90, 90)
            cc->ub_e[i] = u < -90 ? -90 : (u > 90 ? 90 : u);
        }
    }
    return ADY_OK;
}
}  // namespace ady

using namespace ady;

extern "C" {

const char* adyolo_last_error(void) { return last_error_buf(); }
int adyolo_version(void) { return 200; }
long long adyolo_launch_count(void) { return launch_counter().load(); }
size_t adyolo_loss_bad_rows_offset(void) { return offsetof(LossAccum, bad_rows); }

int adyolo_mel_filterbank(int sr, int n_fft, int n_mels, float* out_host) {
    if (!out_host || sr <= 0 || n_fft <= 0 || n_mels <= 0) return set_error(ADY_ERR_INVALID, "mel_filterbank: bad args");
    mel_filterbank_host(sr, n_fft, n_mels, out_host);
    return ADY_OK;
}

size_t adyolo_frontend_workspace_bytes(const adyolo_frontend_cfg* cfg, int B, int64_t N) {
    if (check_frontend_cfg(cfg)) return 0;
    return frontend_workspace_bytes(B, (long long)N);
}

int adyolo_features_foa_views(const int16_t* audio, const int64_t* clip_offsets, int B, int64_t N,
                              const adyolo_frontend_cfg* cfg, const float* mean, const float* inv_std,
                              const int8_t* rot_comb, float* out, void* workspace, int apply_topdb, void* stream) {
    int rc = check_frontend_cfg(cfg);
    if (rc) return rc;
    if (!audio || !out || !workspace) return set_error(ADY_ERR_INVALID, "features_foa: NULL pointer");
    if ((mean == nullptr) != (inv_std == nullptr)) return set_error(ADY_ERR_INVALID, "features_foa: mean and inv_std must both be given or both NULL");
    return launch_features_foa(audio, B, (long long)N, mean, inv_std, cfg->dc_offset, cfg->top_db, apply_topdb, rot_comb,
                               (const long long*)clip_offsets, out, workspace, (cudaStream_t)stream);
}

int adyolo_features_foa_rot(const int16_t* audio, int B, int64_t N, const adyolo_frontend_cfg* cfg,
                            const float* mean, const float* inv_std, const int8_t* rot_comb, float* out,
                            void* workspace, int apply_topdb, void* stream) {
    return adyolo_features_foa_views(audio, nullptr, B, N, cfg, mean, inv_std, rot_comb, out, workspace, apply_topdb, stream);
}

int adyolo_features_foa(const int16_t* audio, int B, int64_t N, const adyolo_frontend_cfg* cfg,
                        const float* mean, const float* inv_std, float* out, void* workspace,
                        int apply_topdb, void* stream) {
    return adyolo_features_foa_rot(audio, B, N, cfg, mean, inv_std, nullptr, out, workspace, apply_topdb, stream);
}

int adyolo_features_mic_logmel(const int16_t* audio, int B, int64_t N, const adyolo_frontend_cfg* cfg, const float* mean,
                               const float* inv_std, float* out, void* spec_c64, void* workspace, int apply_topdb,
                               void* stream) {
    int rc = check_frontend_cfg(cfg);
    if (rc) return rc;
    if (!audio || !out || !spec_c64 || !workspace) return set_error(ADY_ERR_INVALID, "features_mic_logmel: NULL pointer");
    if ((mean == nullptr) != (inv_std == nullptr)) return set_error(ADY_ERR_INVALID, "features_mic_logmel: mean and inv_std must both be given or both NULL");
    return launch_features_mic_logmel(audio, B, (long long)N, mean, inv_std, cfg->dc_offset, cfg->top_db, apply_topdb, out,
                                      (float2*)spec_c64, workspace, (cudaStream_t)stream);
}

size_t adyolo_mic_spec_bytes(int B, int64_t N) {
    if (B <= 0 || N < HOP) return 0;
    return (size_t)B * (size_t)(N / HOP) * NBIN * 4 * sizeof(float2);
}

int adyolo_features_mic_gcc(const int16_t* audio, int B, int64_t N, const adyolo_frontend_cfg* cfg, const float* mean,
                            const float* inv_std, float* out, void* spec_c64, void* workspace, int apply_topdb,
                            void* stream) {
    // product path: fe2 kernel (log-mel + half2 unit phasors, 16 bytes per bin through the scratch) -> tcgen05 lag
    // transform.  The scratch is the caller's `spec_c64` buffer (adyolo_mic_spec_bytes is sized for the larger
    // complex64 layout of adyolo_features_mic_logmel, which remains as the materialising compatibility entry).
    int rc = check_frontend_cfg(cfg);
    if (rc) return rc;
    if (!audio || !out || !spec_c64 || !workspace) return set_error(ADY_ERR_INVALID, "features_mic_gcc: NULL pointer");
    if ((mean == nullptr) != (inv_std == nullptr)) return set_error(ADY_ERR_INVALID, "features_mic_gcc: mean and inv_std must both be given or both NULL");
    rc = launch_features_mic_fe2(audio, B, (long long)N, mean, inv_std, cfg->dc_offset, out, spec_c64, workspace, (cudaStream_t)stream);
    if (rc) return rc;
    if (apply_topdb) {
        rc = launch_features_clamp_nch(out, B, (long long)N, mean, inv_std, cfg->top_db, 10, workspace, (cudaStream_t)stream);
        if (rc) return rc;
    }
    const long long T = (long long)N / HOP;
    return launch_gcc_from_phasors(spec_c64, B, T, mean ? mean + 4 * NMEL : nullptr,
                                   inv_std ? inv_std + 4 * NMEL : nullptr, out + 4 * T * NMEL,
                                   OutStrides{10 * T * NMEL, T * NMEL, NMEL, 1}, (cudaStream_t)stream);
}

int adyolo_features_foa_clamp(float* out, int B, int64_t N, const adyolo_frontend_cfg* cfg, const float* mean,
                              const float* inv_std, void* workspace, void* stream) {
    int rc = check_frontend_cfg(cfg);
    if (rc) return rc;
    if (!out || !workspace) return set_error(ADY_ERR_INVALID, "features_foa_clamp: NULL pointer");
    return launch_features_foa_clamp(out, B, (long long)N, mean, inv_std, cfg->top_db, workspace, (cudaStream_t)stream);
}

int adyolo_stft(const void* audio, int audio_dtype, int B, int64_t N, const adyolo_frontend_cfg* cfg,
                void* out_c64, void* stream) {
    int rc = check_frontend_cfg(cfg);
    if (rc) return rc;
    if (!audio || !out_c64) return set_error(ADY_ERR_INVALID, "stft: NULL pointer");
    return launch_stft(audio, audio_dtype, B, (long long)N, cfg->dc_offset, (float2*)out_c64, (cudaStream_t)stream);
}

int adyolo_logmel_from_stft(const void* spec, int B, int64_t T, int C, int Cs, const adyolo_frontend_cfg* cfg,
                            const float* mean, const float* inv_std, float* out, const int64_t* st,
                            void* gmax_ws, int apply_topdb, void* stream) {
    int rc = check_frontend_cfg(cfg);
    if (rc) return rc;
    if (!spec || !out || !st || !gmax_ws || B <= 0 || T <= 0) return set_error(ADY_ERR_INVALID, "logmel_from_stft: bad args");
    return launch_logmel_from_stft((const float2*)spec, B, (long long)T, C, Cs, mean, inv_std, out,
                                   OutStrides{st[0], st[1], st[2], st[3]}, (uint32_t*)gmax_ws, cfg->top_db,
                                   apply_topdb, (cudaStream_t)stream);
}

int adyolo_iv_from_stft(const void* spec, int B, int64_t T, const adyolo_frontend_cfg* cfg, const float* mean,
                        const float* inv_std, float* out, const int64_t* st, int32_t* flags, void* stream) {
    int rc = check_frontend_cfg(cfg);
    if (rc) return rc;
    if (!spec || !out || !st || !flags || B <= 0 || T <= 0) return set_error(ADY_ERR_INVALID, "iv_from_stft: bad args");
    return launch_iv_from_stft((const float2*)spec, B, (long long)T, mean, inv_std, out,
                               OutStrides{st[0], st[1], st[2], st[3]}, flags, (cudaStream_t)stream);
}

int adyolo_gcc_from_stft(const void* spec, int B, int64_t T, const adyolo_frontend_cfg* cfg, const float* mean,
                         const float* inv_std, float* out, const int64_t* st, void* stream) {
    int rc = check_frontend_cfg(cfg);
    if (rc) return rc;
    if (!spec || !out || !st || B <= 0 || T <= 0) return set_error(ADY_ERR_INVALID, "gcc_from_stft: bad args");
    return launch_gcc_from_stft((const float2*)spec, B, (long long)T, mean, inv_std, out,
                                OutStrides{st[0], st[1], st[2], st[3]}, (cudaStream_t)stream);
}

int adyolo_scaler_partials(const float* feats, int B, int C, int64_t T, double* sum, double* sumsq,
                           double* maxv, double* minv, void* stream) {
    if (!feats || !sum || !sumsq || !maxv || !minv) return set_error(ADY_ERR_INVALID, "scaler_partials: NULL pointer");
    return launch_scaler_partials(feats, B, C, (long long)T, sum, sumsq, maxv, minv, (cudaStream_t)stream);
}

size_t adyolo_label_workspace_bytes(int64_t E) { return label_workspace_bytes((long long)E); }

int adyolo_label_cells(const double* events, int64_t E, int nb_label_frames, const adyolo_grid_cfg* cfg,
                       const int8_t* rot_comb, int64_t n_rot, uint32_t* cellmask, int64_t* total_rows, void* workspace,
                       void* stream) {
    CellCfg cc;
    int rc = make_cfgs(cfg, nullptr, &cc);
    if (rc) return rc;
    if (E > 0 && (!events || !cellmask || !workspace)) return set_error(ADY_ERR_INVALID, "label_cells: NULL pointer");
    if (!total_rows) return set_error(ADY_ERR_INVALID, "label_cells: total_rows is NULL");
    return launch_label_cells(events, (long long)E, nb_label_frames, cc, rot_comb, (long long)n_rot, cellmask, (long long*)total_rows,
                              workspace, (cudaStream_t)stream);
}

int adyolo_label_rows(const double* events, int64_t E, const adyolo_grid_cfg* cfg, const int8_t* rot_comb, int64_t n_rot,
                      const uint32_t* cellmask, const void* workspace, float* rows, int64_t max_rows, void* stream) {
    CellCfg cc;
    int rc = make_cfgs(cfg, nullptr, &cc);
    if (rc) return rc;
    if (E > 0 && max_rows > 0 && (!events || !cellmask || !workspace || !rows)) return set_error(ADY_ERR_INVALID, "label_rows: NULL pointer");
    return launch_label_rows(events, (long long)E, cc, rot_comb, (long long)n_rot, cellmask, workspace, rows, (long long)max_rows,
                             (cudaStream_t)stream);
}

int adyolo_label_cells_rows(const double* events, int64_t E, int nb_label_frames, const adyolo_grid_cfg* cfg,
                            const int8_t* rot_comb, int64_t n_rot, uint32_t* cellmask, int64_t* total_rows,
                            void* workspace, float* rows, int64_t max_rows, void* stream) {
    CellCfg cc;
    int rc = make_cfgs(cfg, nullptr, &cc);
    if (rc) return rc;
    if (!total_rows) return set_error(ADY_ERR_INVALID, "label_cells_rows: NULL total_rows");
    if (E > 0 && (!events || !cellmask || !workspace || (max_rows > 0 && !rows))) return set_error(ADY_ERR_INVALID, "label_cells_rows: NULL pointer");
    return launch_label_cells_rows(events, (long long)E, nb_label_frames, cc, rot_comb, (long long)n_rot, cellmask,
                                   reinterpret_cast<long long*>(total_rows), workspace, rows, (long long)max_rows, (cudaStream_t)stream);
}

int adyolo_assign(const float* logit, const float* target, int64_t M, int B, int T,
                  const adyolo_grid_cfg* cfg, float* D, uint8_t* mask, int32_t* argmin, void* stream) {
    AssignCfg a;
    int rc = make_cfgs(cfg, &a, nullptr);
    if (rc) return rc;
    if (M > 0 && (!logit || !target)) return set_error(ADY_ERR_INVALID, "assign: NULL pointer");
    return launch_assign(logit, target, (long long)M, B, T, a, D, mask, argmin, (cudaStream_t)stream);
}

size_t adyolo_loss_workspace_bytes(int B, int T, const adyolo_grid_cfg* cfg) {
    AssignCfg a;
    if (make_cfgs(cfg, &a, nullptr)) return 0;
    return loss_workspace_bytes(B, T, a);
}

int adyolo_loss(const float* logit, const float* target, int64_t M, int B, int T,
                const adyolo_grid_cfg* cfg, float* loss_out, float* grad_out, float* D, uint8_t* mask,
                int32_t* argmin, void* workspace, void* stream) {
    AssignCfg a;
    int rc = make_cfgs(cfg, &a, nullptr);
    if (rc) return rc;
    if (!logit || !loss_out || !workspace || (M > 0 && !target)) return set_error(ADY_ERR_INVALID, "loss: NULL pointer");
    return launch_loss(logit, target, (long long)M, nullptr, B, T, a, loss_out, grad_out, D, mask, argmin, workspace,
                       (cudaStream_t)stream);
}

int adyolo_loss_devcount(const float* logit, const float* target, int64_t max_rows, const int64_t* n_rows_dev, int B, int T,
                         const adyolo_grid_cfg* cfg, float* loss_out, float* grad_out, void* workspace, void* stream) {
    AssignCfg a;
    int rc = make_cfgs(cfg, &a, nullptr);
    if (rc) return rc;
    if (!logit || !loss_out || !workspace || !target || !n_rows_dev) return set_error(ADY_ERR_INVALID, "loss_devcount: NULL pointer");
    return launch_loss(logit, target, (long long)max_rows, (const long long*)n_rows_dev, B, T, a, loss_out, grad_out,
                       nullptr, nullptr, nullptr, workspace, (cudaStream_t)stream);
}

int adyolo_loss_backward(const float* logit, int B, int T, const adyolo_grid_cfg* cfg, const void* workspace,
                         const float* grad_output, float* grad_out, void* stream) {
    AssignCfg a;
    int rc = make_cfgs(cfg, &a, nullptr);
    if (rc) return rc;
    if (!logit || !workspace || !grad_out) return set_error(ADY_ERR_INVALID, "loss_backward: NULL pointer");
    return launch_loss_backward(logit, B, T, a, workspace, grad_output, grad_out, (cudaStream_t)stream);
}

int adyolo_spec_mask(float* feat, int B, int C, int64_t T, int F, const int32_t* rects, int n_groups,
                     const int32_t* group_bounds, void* stream) {
    if (!feat || !rects || !group_bounds || B < 0 || C <= 0 || T <= 0 || F <= 0 || n_groups < 0)
        return set_error(ADY_ERR_INVALID, "spec_mask: bad args");
    return launch_spec_mask(feat, B, C, (long long)T, F, rects, n_groups, group_bounds, (cudaStream_t)stream);
}

int adyolo_loss_grad_scale(float* grad, int64_t n, const float* grad_output, void* stream) {
    if (!grad || !grad_output || n < 0) return set_error(ADY_ERR_INVALID, "loss_grad_scale: bad args");
    return launch_grad_scale(grad, (long long)n, grad_output, (cudaStream_t)stream);
}

int adyolo_yolo_post(const float* logit, int64_t n_frames, const adyolo_grid_cfg* cfg, float conf_thresh,
                     float clss_thresh, float unify_thresh, int nms_mode, int max_det, float* det, int32_t* count,
                     int32_t* overflow, void* stream) {
    AssignCfg a;
    int rc = make_cfgs(cfg, &a, nullptr);
    if (rc) return rc;
    if (!logit || !det || !count || !overflow) return set_error(ADY_ERR_INVALID, "yolo_post: NULL pointer");
    return launch_yolo_post(logit, (long long)n_frames, a, conf_thresh, clss_thresh, unify_thresh, nms_mode, max_det, det,
                            count, overflow, (cudaStream_t)stream);
}

}  // extern "C"

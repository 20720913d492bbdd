// Host-side declarations for the AD-YOLO assignment / loss / label kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#define ADY_MAX_THR 4
#define ADY_MAX_ANCHORS 8
#define ADY_MAX_GRID 16
#define ADY_LOSS_NCTR 16

namespace ady {

// Everything the kernels need from params['train_config'] / ['data_config'], precomputed on the
// host with the same FP32 rounding steps torch applies (loss.py:157-180).
struct AssignCfg {
    int nb_classes, nb_anchors, ga, ge, n_thr;
    float gs_u, gs_v;            // grid_size
    float ovl_scale;             // float(0.5 + g_overlap)
    float off_u[ADY_MAX_GRID];   // grid_offset[gi][.][0]
    float off_v[ADY_MAX_GRID];   // grid_offset[.][gj][1]
    float thr[ADY_MAX_THR];      // train_unify
    float deg2rad, rad2deg;      // float(pi/180), float(180/pi)  (torch.deg2rad / rad2deg scalars)
    float clip_lo, clip_hi;      // float(-1+1e-7), float(1-1e-7)
    double gain_ang, gain_obj, gain_nonobj, gain_cls;
};

struct LossAccum {
    double ang_sum;
    unsigned long long ang_cnt;
    unsigned long long n_pos[ADY_MAX_THR];
    double s_pos[ADY_MAX_THR], s_neg[ADY_MAX_THR], s_cls[ADY_MAX_THR];
    float w_pos[ADY_MAX_THR], w_neg[ADY_MAX_THR], w_cls[ADY_MAX_THR];   // gain / (n_thr * count), by loss_weights_kernel
    float w_ang, pad_f;
    unsigned int pad_u[2];
    int bad_rows;
    unsigned int done_blocks;    // ticket of loss_anchor_kernel's last-block finalisation
    // loss_anchor_kernel's group counters (dynamic tail of the work split), one per 128-byte line: same-address atomics
    // serialise in one L2 slice
    unsigned int group_ctr[ADY_LOSS_NCTR][32];
};

size_t loss_workspace_bytes(int B, int T, const AssignCfg& cfg);
int launch_assign(const float* logit, const float* target, long long M, int B, int T, const AssignCfg& cfg,
                  float* D, uint8_t* mask, int32_t* argmin, cudaStream_t stream);
int launch_loss(const float* logit, const float* target, long long M, const long long* M_dev, int B, int T,
                const AssignCfg& cfg, float* loss_out, float* grad_out, float* D, uint8_t* mask, int32_t* argmin,
                void* ws, cudaStream_t stream);

int launch_loss_backward(const float* logit, int B, int T, const AssignCfg& cfg, const void* ws,
                         const float* grad_output, float* grad_out, cudaStream_t stream);
int launch_grad_scale(float* grad, long long n, const float* gscale, cudaStream_t stream);

// decode + conn-merge NMS (datasets.py:741-857)
int launch_yolo_post(const float* logit, long long n_frames, const AssignCfg& cfg, float conf_thresh, float clss_thresh,
                     float unify_thresh, int nms_mode, int max_det, float* det, int32_t* count, int* overflow,
                     cudaStream_t stream);

// grid-cell responsibility (datasets.py:457-482): events -> 32-bit cell mask + count, rows
struct CellCfg {
    int ga, ge;
    double lb_a[ADY_MAX_GRID], ub_a[ADY_MAX_GRID], lb_e[ADY_MAX_GRID], ub_e[ADY_MAX_GRID];
};
size_t label_workspace_bytes(long long E);
int launch_label_cells(const double* events, long long E, int nb_label_frames, const CellCfg& cfg, const int8_t* rot, long long n_rot,
                       uint32_t* cellmask, long long* total_rows_dev, void* ws, cudaStream_t stream);
int launch_label_cells_rows(const double* events, long long E, int nb_label_frames, const CellCfg& cfg, const int8_t* rot, long long n_rot,
                            uint32_t* cellmask, long long* total_rows_dev, void* ws, float* rows, long long max_rows, cudaStream_t stream);
int launch_label_rows(const double* events, long long E, const CellCfg& cfg, const int8_t* rot, long long n_rot, const uint32_t* cellmask,
                      const void* ws, float* rows, long long max_rows, cudaStream_t stream);

}  // namespace ady

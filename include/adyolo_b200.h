/* adyolo_b200 — C ABI of the B200-native AD-YOLO hot path (sm_100a CUDA kernels).
 *
 * The reference (sadPororo/AD-YOLO) has no FFI: its hot path is a Python call surface
 * (SURVEY.md §8(b)).  Each entry point below names the reference function it replaces; the
 * Python mirror of that surface lives in ad-yolo_b200/ and binds these symbols with ctypes.
 *
 * Conventions
 *  - every pointer marked "device" is a CUDA device pointer owned by the caller
 *    (e.g. torch.Tensor.data_ptr()); the library allocates nothing per call except the constant
 *    tables it builds once per device (window, sparse mel matrix);
 *  - `stream` is a cudaStream_t passed as void*; all work is asynchronous and stream-ordered;
 *  - return value: 0 = OK, negative = error (ADYOLO_ERR_*); adyolo_last_error() gives the
 *    message of the last failing call on the calling thread.  Nothing throws across the ABI;
 *  - inputs are never modified (the reference's in-place azi 180 -> -180 rewrite of the
 *    caller's label dict, datasets.py:470, is NOT reproduced; outputs carry the rewritten value).
 */
#ifndef ADYOLO_B200_H
#define ADYOLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADYOLO_OK 0
#define ADYOLO_ERR_INVALID (-1)
#define ADYOLO_ERR_CUDA (-2)
#define ADYOLO_ERR_UNSUPPORTED (-3)

#define ADYOLO_MAX_THR 4
#define ADYOLO_MAX_GRID 16

const char* adyolo_last_error(void);
int adyolo_version(void);
/* Number of kernels this library has launched in the calling process so far (every launch site
 * counts itself): what bench.py reports as `gpu_launches`.                                       */
long long adyolo_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Feature front end.  Geometry of src/configs/hyp_data_DCASE20xx.yaml:5-12 (the only one the
 * reference ships): sr 24000, n_fft = win_length = 1200, hop_length 600, 64 mel bins, 4 channels,
 * periodic Hann window.  Other geometries return ADYOLO_ERR_UNSUPPORTED.                        */
typedef struct adyolo_frontend_cfg {
    int32_t sr, n_fft, hop_length, win_length, mel_bins, n_channels;
    float dc_offset; /* the "+ 1e-8" of datasets.py:147 / preprocess.py:105 (int16 input only) */
    float top_db;    /* librosa.power_to_db default 80.0 (datasets.py:265)                      */
} adyolo_frontend_cfg;

/* librosa.filters.mel(sr, n_fft, n_mels) with librosa-0.8.1 defaults -> host buffer
 * (n_mels x (1+n_fft/2)) float32.  Replaces datasets.py:203 / utility.py:183,204.  Host only. */
int adyolo_mel_filterbank(int sr, int n_fft, int n_mels, float* out_host);

/* bytes of device scratch needed by adyolo_features_foa / adyolo_features_mic_gcc */
size_t adyolo_frontend_workspace_bytes(const adyolo_frontend_cfg* cfg, int B, int64_t N);

/* FeatureLabelProcessor.get_feature (datasets.py:281-292) + normalisation (:147) + the (C,T,F)
 * stacking of Dataset.__getitem__ (:158-160), batched:
 *   audio  device int16 (B, N, 4) interleaved PCM (W,Y,Z,X), as scipy.io.wavfile.read returns it
 *   mean, std  device float32 (7, 64) channel-major [mel W,Y,Z,X | iv Y,Z,X] or NULL (no
 *              standardisation: the output of utility.stft2melscale / stft2iv)
 *   out    device float32 (B, 7, T, 64), T = N / 600
 *   apply_topdb  1: power_to_db's global per-(clip,channel) top_db clamp; 0: leave unclamped
 *   workspace    device scratch of adyolo_frontend_workspace_bytes(); after the call its first
 *                int32 has bit0 set if an intensity value was NaN (the reference prints and
 *                exit()s, datasets.py:277-278)                                                  */
int adyolo_features_foa(const int16_t* audio, int B, int64_t N, const adyolo_frontend_cfg* cfg,
                        const float* mean, const float* inv_std, float* out, void* workspace,
                        int apply_topdb, void* stream);

/* adyolo_features_foa with the FOA rotation augmentation (RotationAug._rotate,
 * utils/augmentations.py:84-111) fused in: rot_comb = device int8 (B) combination number 0..15
 * per clip (NULL = none).  Equivalent to sign-flipping / swapping the int16 channels before
 * extraction (except that -1 * -32768 does not wrap as numpy int16 does).                      */
int adyolo_features_foa_rot(const int16_t* audio, int B, int64_t N, const adyolo_frontend_cfg* cfg,
                            const float* mean, const float* inv_std, const int8_t* rot_comb, float* out,
                            void* workspace, int apply_topdb, void* stream);

/* adyolo_features_foa over VIEWS of one resident buffer: clip b = N samples starting at sample
 * clip_offsets[b] (device int64 (B); even offsets) of `audio` (device int16 (S, 4)).  Serves the
 * reference's `chunking` action (preprocess.py:13-48: 20-s windows every 1 s) without writing the
 * 41x-duplicated chunk files: each chunk is still extracted as an independent clip (own reflect
 * padding and top_db maximum).  rot_comb as above (may be NULL).                     */
int adyolo_features_foa_views(const int16_t* audio, const int64_t* clip_offsets, int B, int64_t N,
                              const adyolo_frontend_cfg* cfg, const float* mean, const float* inv_std,
                              const int8_t* rot_comb, float* out, void* workspace, int apply_topdb, void* stream);

/* Second half of adyolo_features_foa when it was called with apply_topdb = 0: applies the
 * power_to_db top_db clamp (datasets.py:265) from the per-(clip,channel) maxima and minima of the
 * un-clamped dB values that call left in `workspace` (the fused kernel's epilogue reduces them): a
 * plane whose minimum is within top_db of its maximum is left alone, the others are rewritten in one
 * pass.  Pass the SAME workspace, untouched in between.  Exposed separately so the two kernels can
 * be timed individually.                                                                        */
int adyolo_features_foa_clamp(float* out, int B, int64_t N, const adyolo_frontend_cfg* cfg, const float* mean,
                              const float* inv_std, void* workspace, void* stream);

/* MIC (microphone array) format, first half: the fused front end writes the 4 log-mel channels
 * into channels 0..3 of out (B, 10, T, 64) (mean / inv_std: device (>=4, 64) or NULL) and the four
 * channel spectra into spec_c64 (B, T, 601, 4) complex64; adyolo_gcc_from_stft then fills
 * channels 4..9.  The reference has no MIC path: parity unpinned.                  */
int adyolo_features_mic_logmel(const int16_t* audio, int B, int64_t N, const adyolo_frontend_cfg* cfg, const float* mean,
                               const float* inv_std, float* out, void* spec_c64, void* workspace, int apply_topdb,
                               void* stream);

/* MIC format in one call: adyolo_features_mic_logmel followed by adyolo_gcc_from_stft on the same
 * stream.  out (B, 10, T, 64) = 4 log-mel + 6 GCC-PHAT (64 lags, pairs (0,1) (0,2) (0,3) (1,2) (1,3)
 * (2,3)); mean / inv_std: device (10, 64) or NULL; spec_c64: caller-owned scratch of
 * adyolo_mic_spec_bytes(B, N) bytes (the channel spectra travel through it).                     */
size_t adyolo_mic_spec_bytes(int B, int64_t N);
int adyolo_features_mic_gcc(const int16_t* audio, int B, int64_t N, const adyolo_frontend_cfg* cfg, const float* mean,
                            const float* inv_std, float* out, void* spec_c64, void* workspace, int apply_topdb,
                            void* stream);

/* Un-fused stages behind the reference's per-function surface (they materialise the STFT):
 *
 * adyolo_stft  — utility.audio2stft (utility.py:142-165) == get_stft_spectrogram (datasets.py:252-258)
 *   audio_dtype 0: device int16 (B,N,4), normalised as x/32768 + cfg->dc_offset; 1: device float32
 *   (B,N,4) already normalised.  out: device complex64 (B, T, 601, 4), T = N/600.
 * adyolo_logmel_from_stft — utility.stft2melscale (utility.py:168-191) == datasets.py:260-267
 *   spec device complex64 (B, T, 601, Cs); the first C (<= 4) channels are used.
 * adyolo_iv_from_stft — utility.stft2iv (utility.py:194-215) == datasets.py:269-279 (Cs = 4)
 * adyolo_gcc_from_stft — GCC-PHAT of the 6 microphone pairs, 64 lags.  NOT in the reference
 *   (SURVEY F1): upstream seld-dcase2022 `_get_gcc` semantics, parity unpinned.
 * Output element (b, c, t, j) is written to out[b*sb + c*sc + t*st + j*sj] (strides in elements),
 * so both the reference's (T, 64, C) layout and the batched (B, C, T, 64) layout are served.
 * mean / inv_std: device float32 (C, 64) or NULL.  gmax_ws: device uint32 (B*C) scratch.       */
int adyolo_stft(const void* audio, int audio_dtype, int B, int64_t N, const adyolo_frontend_cfg* cfg,
                void* out_c64, void* stream);
int adyolo_logmel_from_stft(const void* spec, int B, int64_t T, int C, int Cs, const adyolo_frontend_cfg* cfg,
                            const float* mean, const float* inv_std, float* out, const int64_t* strides4,
                            void* gmax_ws, int apply_topdb, void* stream);
int adyolo_iv_from_stft(const void* spec, int B, int64_t T, const adyolo_frontend_cfg* cfg, const float* mean,
                        const float* inv_std, float* out, const int64_t* strides4, int32_t* flags, void* stream);
int adyolo_gcc_from_stft(const void* spec, int B, int64_t T, const adyolo_frontend_cfg* cfg, const float* mean,
                         const float* inv_std, float* out, const int64_t* strides4, void* stream);

/* preprocess_scaler accumulation (preprocess.py:116-127): FP64 partial statistics over frames of
 * un-standardised features (B, C, T, 64) -> sum, sumsq (C,64) double (accumulated into: zero them
 * first), maxv/minv (C,64) double (initialise to -inf/+inf).  count += B*T.                    */
int adyolo_scaler_partials(const float* feats, int B, int C, int64_t T, double* sum, double* sumsq,
                           double* maxv, double* minv, void* stream);

/* ---------------------------------------------------------------------------------------------
 * AD-YOLO labels and loss.                                                                      */
typedef struct adyolo_grid_cfg {
    int32_t nb_classes;                 /* data_config.nb_classes (<= 15)                         */
    int32_t nb_anchors;                 /* train_config.nb_anchors (<= 8)                         */
    double grid_size[2];                /* train_config.grid_size -- double: the label path        */
    double g_overlap;                   /* train_config.g_overlap    (datasets.py:229-233) is float64 */
    int32_t n_thr;                      /* len(train_config.train_unify) (<= 4)                   */
    float train_unify[ADYOLO_MAX_THR];
    float angular_gain, object_gain, nonobj_gain, class_gain; /* train_config.loss_gains          */
} adyolo_grid_cfg;

/* FeatureLabelProcessor.get_yolo_label (datasets.py:457-482) + label half of collate_fn (:175-184)
 *   events   device float64 (E, 5) rows [batch, frame, class, azi, ele] in dataset order
 *   cellmask device uint32 (E): bit (Gi*Ge + Gj) set for every responsible cell -- hence grids of at most
 *            32 cells (the reference's 45-degree grid is 8 x 4 = 32; 60 / 90 degrees are smaller); a finer grid
 *            returns ADYOLO_ERR_UNSUPPORTED here (the loss entries take up to 16 x 16 cells)
 *   total_rows device int64[1]: M = total number of rows
 * then adyolo_label_rows writes rows device float32 (M, 7) [batch, frame, Gi, Gj, class, U, V].  */
size_t adyolo_label_workspace_bytes(int64_t E);
/* rot_comb: device int8 (n_clips) rotation combination per batch index, or NULL: the label half
 * of RotationAug._rotate (augmentations.py:98-109) is applied to (azi, ele) before the cell test. */
int adyolo_label_cells(const double* events, int64_t E, int nb_label_frames, const adyolo_grid_cfg* cfg,
                       const int8_t* rot_comb, int64_t n_rot, uint32_t* cellmask, int64_t* total_rows,
                       void* workspace, void* stream);
/* Rows past max_rows are not written; *total_rows (from adyolo_label_cells) still holds the full
 * count, so `*total_rows > max_rows` is the overflow test (DeviceRows.materialize raises on it).
 * n_rot = number of entries of rot_comb: events whose batch id lies outside [0, n_rot) are left
 * unrotated rather than read out of bounds.                                                      */
int adyolo_label_rows(const double* events, int64_t E, const adyolo_grid_cfg* cfg, const int8_t* rot_comb,
                      int64_t n_rot, const uint32_t* cellmask, const void* workspace, float* rows,
                      int64_t max_rows, void* stream);
/* adyolo_label_cells + adyolo_label_rows in one call for a caller that sizes `rows` up front (max_rows): two launches,
 * the block offsets are derived inside the row kernel and *total_rows is written by it (same overflow contract).    */
int adyolo_label_cells_rows(const double* events, int64_t E, int nb_label_frames, const adyolo_grid_cfg* cfg,
                            const int8_t* rot_comb, int64_t n_rot, uint32_t* cellmask, int64_t* total_rows,
                            void* workspace, float* rows, int64_t max_rows, void* stream);

/* ADYOLOloss decode + distance_between_polar_coordinates + responsibility (loss.py:193-226):
 *   logit  device float32 (B, T, Ga*Ge*A*(C+3));  target device float32 (M, 7)
 *   D      device float32 (M, A)            (may be NULL)
 *   mask   device uint8   (n_thr, M, A)     (may be NULL)   1 = responsible
 *   argmin device int32   (M)               (may be NULL)                                       */
int adyolo_assign(const float* logit, const float* target, int64_t M, int B, int T,
                  const adyolo_grid_cfg* cfg, float* D, uint8_t* mask, int32_t* argmin, void* stream);

/* ADYOLOloss.__call__ (loss.py:189-251).
 *   loss_out device float32[1]
 *   grad_out device float32 like logit: d loss / d logit written in the same pass (upstream
 *            gradient 1), or NULL for forward only — then adyolo_loss_backward produces
 *            grad_output[0] * d loss / d logit later from the same `workspace` (which keeps the
 *            label bits, counts and angular partials of the forward call; do not reuse it in
 *            between).  D / mask / argmin as in adyolo_assign, may be NULL.                     */
size_t adyolo_loss_workspace_bytes(int B, int T, const adyolo_grid_cfg* cfg);
/* Byte offset, inside the loss workspace, of an int32 that counts the target rows the last
 * adyolo_loss / adyolo_loss_devcount call skipped because (batch, frame, Gi, Gj, class) fell outside
 * the logit tensor.  The reference raises IndexError / a device assert on such rows
 * (loss.py:216,228-232); read it back (ADYOLOloss(check_rows=True) does) to get the same error. */
size_t adyolo_loss_bad_rows_offset(void);
int adyolo_loss(const float* logit, const float* target, int64_t M, int B, int T,
                const adyolo_grid_cfg* cfg, float* loss_out, float* grad_out, float* D, uint8_t* mask,
                int32_t* argmin, void* workspace, void* stream);

/* Same as adyolo_loss, but the number of valid target rows is read from device memory
 * (the int64 written by adyolo_label_cells), `target` having capacity max_rows: lets label rows
 * and loss be enqueued back to back with no host synchronisation in between.                 */
int adyolo_loss_devcount(const float* logit, const float* target, int64_t max_rows, const int64_t* n_rows_dev,
                         int B, int T, const adyolo_grid_cfg* cfg, float* loss_out, float* grad_out,
                         void* workspace, void* stream);
int adyolo_loss_backward(const float* logit, int B, int T, const adyolo_grid_cfg* cfg, const void* workspace,
                         const float* grad_output /* device float32[1] or NULL (= 1) */, float* grad_out,
                         void* stream);

/* Autograd glue for the fused forward (adyolo_loss with grad_out != NULL already holds
 * d loss / d logit): grad[0..n) *= grad_output[0], in place, with the scalar read on the device
 * -- when it is exactly 1 (the plain `loss.backward()` of train.py:54) the kernel returns without
 * touching the buffer, so backward costs one launch and no memory traffic.                     */
int adyolo_loss_grad_scale(float* grad, int64_t n, const float* grad_output /* device float32[1] */, void* stream);

/* SpecAug masks (augmentations.py:6-33 as applied by datasets.py:158-160 to each feature group
 * permuted to (C, T, F): torchaudio "TimeMasking" therefore zeroes a band of MEL bins and
 * "FrequencyMasking" a run of FRAMES, shared by the channels of a group; mask value 0).
 *   feat        device float32 (B, C, T, F) contiguous, masked in place (write-only)
 *   rects       device int32 (B, n_groups, 4) = [mel0, mel1, frame0, frame1), empty = no mask
 *   group_bounds device int32 (n_groups, 2) = [c0, c1) channel range of each group
 * The intervals are drawn on the host with the reference's RNG sequence (adyolo_b200.SpecAug).  */
int adyolo_spec_mask(float* feat, int B, int C, int64_t T, int F, const int32_t* rects, int n_groups,
                     const int32_t* group_bounds, void* stream);

/* LabelPostProcessor.get_yolo_output (datasets.py:741-857) for n_frames frames of logits
 * (n_frames, Ga*Ge*A*(C+3)): decode, class-confidence thresholding, then per class
 *   nms_mode 0  'conn-merge' (reference default, :786-814): connected components of D < unify_thresh,
 *               softmax-weighted Cartesian vote per component
 *   nms_mode 1  'soft-merge' (:817-831): best remaining candidate voted with every candidate of the
 *               class within D <= unify_thresh, then it and its neighbours are suppressed
 *   nms_mode 2  any other string (:834-846): plain NMS, best kept as is, D <= unify_thresh suppressed
 *   det   device float32 (n_frames, max_det, 4) rows [class, x, y, z] in the reference's order
 *   count device int32 (n_frames); overflow device int32[1] set when a frame had > max_det rows */
int adyolo_yolo_post(const float* logit, int64_t n_frames, const adyolo_grid_cfg* cfg, float conf_thresh,
                     float clss_thresh, float unify_thresh, int nms_mode, int max_det, float* det, int32_t* count,
                     int32_t* overflow, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADYOLO_B200_H */

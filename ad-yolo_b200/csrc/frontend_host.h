// Host-side declarations shared by the front-end translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ady {

struct FrontendTables {
    float melw[1216];        // concatenated non-zero runs of the (64 x 601) Slaney mel matrix
    int16_t melidx[3 * 64];  // [start | len | offset into melw] per mel filter
    float hann[1200];        // plain periodic Hann (for the float-input / stft paths)
    float wcs[25 * 2];       // fused kernel: cos / sin (2 pi n2 / 25) of the window factorisation
    int mel_hdr[8];          // fused kernel: static mel schedule, it0[4] | nit[4]
    int mel_pos[56 * 32];    // fused kernel: schedule entries [row][lane]: V position ...
    float mel_w[56 * 32];    //               ... and weight (0 = padding)
};

// librosa.filters.mel(sr, n_fft, n_mels) defaults (Slaney scale + norm, float32), row-major
// (n_mels, 1 + n_fft/2).  Host arithmetic in double, replicating librosa's rounding steps.
void mel_filterbank_host(int sr, int n_fft, int n_mels, float* out);

// Device-resident constant tables for the fixed DCASE geometry (built once per device).
int get_frontend_tables(const FrontendTables** dev_tables);

size_t frontend_workspace_bytes(int B, long long N);
int launch_features_foa(const int16_t* audio, int B, long long N, const float* mean, const float* istd,
                        float dc_offset, float top_db, int apply_topdb, const int8_t* rot, const long long* clip_off,
                        float* out, void* ws, cudaStream_t stream);

// second-generation fused kernel (fe2.cu): FOA log-mel + IV, un-clamped; same arguments as launch_features_foa minus the clamp
int launch_features_foa_fe2(const int16_t* audio, int B, long long N, const float* mean, const float* istd, float dc_offset,
                            const int8_t* rot, const long long* clip_off, float* out, void* ws, cudaStream_t stream);

// MIC format on the fe2 kernel: log-mel channels 0..3 of out (B, 10, T, 64), un-clamped, + half2 unit phasors
// (B, T, 608, 4) for launch_gcc_from_phasors (gcc_tc.cu)
int launch_features_mic_fe2(const int16_t* audio, int B, long long N, const float* mean, const float* istd, float dc_offset,
                            float* out, void* phasor, void* ws, cudaStream_t stream);

int launch_features_mic_logmel(const int16_t* audio, int B, long long N, const float* mean, const float* istd,
                               float dc_offset, float top_db, int apply_topdb, float* out, float2* spec, void* ws,
                               cudaStream_t stream);
int launch_features_clamp_nch(float* out, int B, long long N, const float* mean, const float* istd, float top_db, int nch,
                               const void* ws, cudaStream_t stream);
int launch_features_foa_clamp(float* out, int B, long long N, const float* mean, const float* istd, float top_db,
                              void* ws, cudaStream_t stream);

}  // namespace ady

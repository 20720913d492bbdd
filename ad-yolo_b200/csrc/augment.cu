// SpecAug masks: augmentations.py:6-33 applied per feature group by
// datasets.py:158-160.  The reference permutes each group to (C, T, F) before torchaudio's
// TimeMasking / FrequencyMasking, so the "time" mask zeroes a band of MEL bins and the "frequency"
// mask a run of FRAMES, shared by all channels of the group; the mask value is 0 (= the mean after
// standardisation).  The intervals are drawn on the host (same RNG sequence as the reference, see
// augment.py); this kernel only writes zeros into the masked regions of the feature tensor that the
// front end produced -- no reads, traffic = the masked fraction.
#include <stdint.h>

#include "common.cuh"

namespace ady {

// feat (B, C, T, F) contiguous; rects (B, G, 4) = [mel0, mel1, frame0, frame1]; bounds (G, 2) = [c0, c1)
__global__ void __launch_bounds__(256)
spec_mask_kernel(float* __restrict__ feat, int C, long long T, int F, const int* __restrict__ rects, int G,
                 const int* __restrict__ bounds) {
    const int b = blockIdx.x / G, g = blockIdx.x - b * G;
    const int* r = rects + ((long long)b * G + g) * 4;
    const int m0 = max(r[0], 0), m1 = min(r[1], F);
    const long long f0 = max(r[2], 0), f1 = min((long long)r[3], T);
    const int c0 = bounds[2 * g], c1 = bounds[2 * g + 1];
    // frame mask: (f1 - f0) * F contiguous floats per channel
    if (f1 > f0) {
        const long long run = (f1 - f0) * F;
        for (int c = c0; c < c1; ++c) {
            float* p = feat + (((long long)b * C + c) * T + f0) * F;
            if ((F & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                float4* p4 = reinterpret_cast<float4*>(p);
                for (long long i = threadIdx.x; i < run / 4; i += blockDim.x) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                for (long long i = threadIdx.x; i < run; i += blockDim.x) p[i] = 0.f;
            }
        }
    }
    // mel mask: (m1 - m0) floats in every row of every channel
    if (m1 > m0) {
        const int w = m1 - m0;
        const long long rows = (long long)(c1 - c0) * T;
        for (long long i = threadIdx.x; i < rows * w; i += blockDim.x) {
            const long long row = i / w;
            const int j = (int)(i - row * w);
            feat[((long long)b * C + c0) * T * F + row * F + m0 + j] = 0.f;
        }
    }
}

int launch_spec_mask(float* feat, int B, int C, long long T, int F, const int* rects, int G, const int* bounds,
                     cudaStream_t stream) {
    if (B <= 0 || G <= 0) return ADY_OK;
    if ((long long)B * G > 0x7fffffffLL) return set_error(ADY_ERR_INVALID, "spec_mask: too many (clip, group) pairs");
    spec_mask_kernel<<<(unsigned)(B * G), 256, 0, stream>>>(feat, C, T, F, rects, G, bounds);
    ADY_LAUNCH_CHECK("spec_mask_kernel");
    return ADY_OK;
}

}  // namespace ady

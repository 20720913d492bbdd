// AD-YOLO decode + NMS post-processing on the GPU: the reference's default
// connectivity-merge and its two alternatives, soft-merge and plain NMS.
//
// Reference behaviour replaced: /root/reference/src/datasets.py:741-857 (LabelPostProcessor.
// get_yolo_output, all three `nms` branches) and its helpers :863-919.  The reference walks the
// frames of one clip in Python on the CPU (it dominates validation time); here one block handles
// one (clip, frame):
//   1. thread = anchor: decode (sigmoid / tanh, scale, offset, clamp V, wrap U) with the same
//      individually rounded FP32 op sequence as the loss kernel; class score = sigmoid(cls) *
//      sigmoid(obj); (anchor, class) candidates above the two thresholds are appended per class
//   2. per class: rank-sort by score (descending), pairwise great-circle distances < unify_thresh
//      -> adjacency bit rows, connected components grown from the highest-scoring remaining
//      candidate (the reference's iterative closure), softmax(exp(score^2/thr))-weighted Cartesian
//      vote per component; a class with a single candidate is converted directly
// Output: det (B*T, max_det, 4) float32 rows [class, x, y, z] in the reference's order (classes
// ascending, components by their best score) and count (B*T) int32.
// The threshold decisions (score > thr, D < unify) use the ATen-equivalent FP32 sequence so that
// they are identical to torch executing the reference's ops on the same GPU.
#include "assign_host.h"
#include "common.cuh"

namespace ady {

constexpr int NMS_MAXA = 256;   // max anchors per frame (Ga*Ge*A)

__device__ __forceinline__ float nms_sigmoid(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }
__device__ __forceinline__ float nms_clamp(float v, float lo, float hi) { return (v != v) ? v : fminf(fmaxf(v, lo), hi); }

struct NmsCfg {
    float conf_thresh, clss_thresh, unify_thresh, v_hi;   // v_hi = float(90 - 1e-7)
    float inv_clss;                                       // ATen divides by a scalar as a * (1/b)
    int max_det;
    int mode;                                             // 0 conn-merge, 1 soft-merge, 2 plain NMS (datasets.py:786-848)
};

__global__ void __launch_bounds__(NMS_MAXA)
yolo_post_kernel(const float* __restrict__ logit, AssignCfg cfg, NmsCfg nc, float* __restrict__ det,
                 int32_t* __restrict__ count, int* __restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int NA = cfg.ga * cfg.ge * cfg.nb_anchors, C = cfg.nb_classes, CH = C + 3;
    // per class candidate lists (unsorted), then one working set for the class being processed
    float* c_score = reinterpret_cast<float*>(sm);                 // [C][NA]
    float* c_u = c_score + C * NA;                                  // [C][NA]
    float* c_v = c_u + C * NA;                                      // [C][NA]
    short* c_anchor = reinterpret_cast<short*>(c_v + C * NA);       // [C][NA]
    int* c_cnt = reinterpret_cast<int*>(c_anchor + C * NA + (C * NA & 1));   // [C]
    float* w_score = reinterpret_cast<float*>(c_cnt + ((C + 3) & ~3));      // [NA] sorted working set
    float* w_ur = w_score + NA;                                     // [NA] azimuth in rad
    float* w_sv = w_ur + NA;                                        // [NA] sin(elevation)
    float* w_cv = w_sv + NA;                                        // [NA] cos(elevation)
    float* w_cx = w_cv + NA;                                        // [NA] cartesian x, y, z
    float* w_cy = w_cx + NA;
    float* w_cz = w_cy + NA;
    unsigned* adj = reinterpret_cast<unsigned*>(w_cz + NA);         // [NA][8] adjacency bit rows
    __shared__ int s_ndet;

    const int tid = threadIdx.x;
    const long long frame = blockIdx.x;
    if (tid < C) c_cnt[tid] = 0;
    if (tid == 0) s_ndet = 0;
    __syncthreads();

    // ---- 1. decode + candidates
    if (tid < NA) {
        const float* x = logit + (frame * NA + tid) * CH;
        const float p0 = nms_sigmoid(x[0]);
        if (p0 > nc.conf_thresh) {
            const int cell = tid / cfg.nb_anchors, gi = cell / cfg.ge, gj = cell - gi * cfg.ge;
            const float thu = tanhf(x[C + 1]), thv = tanhf(x[C + 2]);
            float U = __fadd_rn(__fmul_rn(__fmul_rn(thu, cfg.ovl_scale), cfg.gs_u), cfg.off_u[gi]);
            float V = __fadd_rn(__fmul_rn(__fmul_rn(thv, cfg.ovl_scale), cfg.gs_v), cfg.off_v[gj]);
            V = nms_clamp(V, -90.f, nc.v_hi);                         // datasets.py:765
            if (U >= 180.f) U = __fadd_rn(U, -360.f);
            if (U < -180.f) U = __fadd_rn(U, 360.f);
            for (int c = 0; c < C; ++c) {
                const float s = __fmul_rn(nms_sigmoid(x[1 + c]), p0);   // :772
                if (s > nc.clss_thresh) {                             // :783
                    const int i = atomicAdd(&c_cnt[c], 1);
                    c_score[c * NA + i] = s; c_u[c * NA + i] = U; c_v[c * NA + i] = V; c_anchor[c * NA + i] = (short)tid;
                }
            }
        }
    }
    __syncthreads();

    float* out = det + frame * nc.max_det * 4;
    // ---- 2. classes in ascending order (torch.unique)
    for (int c = 0; c < C; ++c) {
        const int n = c_cnt[c];
        if (n == 0) continue;                                         // block-uniform
        // rank sort: descending score, ties by (anchor, i.e. nonzero order)
        if (tid < n) {
            const float s = c_score[c * NA + tid];
            const int a = c_anchor[c * NA + tid];
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const float sj = c_score[c * NA + j];
                rank += (sj > s) || (sj == s && c_anchor[c * NA + j] < a);
            }
            const float ur = __fmul_rn(c_u[c * NA + tid], cfg.deg2rad), vr = __fmul_rn(c_v[c * NA + tid], cfg.deg2rad);
            const float sv = sinf(vr), cv = cosf(vr);
            w_score[rank] = s; w_ur[rank] = ur; w_sv[rank] = sv; w_cv[rank] = cv;
            w_cx[rank] = __fmul_rn(cosf(ur), cv);                     // :889-891 / :913-915
            w_cy[rank] = __fmul_rn(sinf(ur), cv);
            w_cz[rank] = sv;
        }
        __syncthreads();
        if (n == 1) {
            if (tid == 0) {                                           // :791-793 direct conversion
                const int d = s_ndet;
                if (d < nc.max_det) { out[d * 4] = (float)c; out[d * 4 + 1] = w_cx[0]; out[d * 4 + 2] = w_cy[0]; out[d * 4 + 3] = w_cz[0]; }
                else *overflow = 1;
                s_ndet = d + 1;
            }
            __syncthreads();
            continue;
        }
        // adjacency rows: D(i, j) < unify_thresh  (:795-797, distance of :863-876 with clip(-1, 1))
        if (tid < n) {
            unsigned row[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const float ui = w_ur[tid], svi = w_sv[tid], cvi = w_cv[tid];
            for (int j = 0; j < n; ++j) {
                const float du = fabsf(__fadd_rn(w_ur[j], -ui));
                const float dist = __fadd_rn(__fmul_rn(w_sv[j], svi), __fmul_rn(__fmul_rn(w_cv[j], cvi), cosf(du)));
                const float D = __fmul_rn(acosf(nms_clamp(dist, -1.f, 1.f)), cfg.rad2deg);
                // conn-merge links with D < thr (:801); soft-merge / plain NMS unify and suppress with D <= thr (:826-831, :845-846)
                if (nc.mode == 0 ? D < nc.unify_thresh : D <= nc.unify_thresh) row[j >> 5] |= 1u << (j & 31);
            }
#pragma unroll
            for (int w = 0; w < 8; ++w) adj[tid * 8 + w] = row[w];
        }
        __syncthreads();
        // connected components + votes: one thread walks the (small) bit sets
        if (tid == 0) {
            unsigned rem[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) rem[w] = (n > 32 * w) ? (n >= 32 * (w + 1) ? 0xffffffffu : ((1u << (n - 32 * w)) - 1u)) : 0u;
            int left = n;
            while (left > 0) {
                int seed = 0;
#pragma unroll
                for (int w = 7; w >= 0; --w) if (rem[w]) seed = 32 * w + __ffs(rem[w]) - 1;
                unsigned cur[8], pre[8];
                if (nc.mode != 0) {
                    // soft-merge (:818-831): the best remaining candidate is voted together with ALL candidates
                    // of the class within the threshold (also already suppressed ones); plain NMS (:834-846)
                    // emits it as it is.  Both then drop it and every remaining candidate within the threshold.
                    const int d = s_ndet;
                    if (nc.mode == 2) {
                        if (d < nc.max_det) { out[d * 4] = (float)c; out[d * 4 + 1] = w_cx[seed]; out[d * 4 + 2] = w_cy[seed]; out[d * 4 + 3] = w_cz[seed]; }
                        else *overflow = 1;
                    } else {
                        float emax = -INFINITY;
                        for (int w = 0; w < 8; ++w) for (unsigned b = adj[seed * 8 + w]; b; b &= b - 1) {
                            const int m = 32 * w + __ffs(b) - 1;
                            emax = fmaxf(emax, expf(__fmul_rn(__fmul_rn(w_score[m], w_score[m]), nc.inv_clss)));
                        }
                        float den = 0.f;
                        for (int w = 0; w < 8; ++w) for (unsigned b = adj[seed * 8 + w]; b; b &= b - 1) {
                            const int m = 32 * w + __ffs(b) - 1;
                            den += expf(expf(__fmul_rn(__fmul_rn(w_score[m], w_score[m]), nc.inv_clss)) - emax);
                        }
                        float vx = 0.f, vy = 0.f, vz = 0.f;
                        for (int w = 0; w < 8; ++w) for (unsigned b = adj[seed * 8 + w]; b; b &= b - 1) {
                            const int m = 32 * w + __ffs(b) - 1;
                            const float wt = expf(expf(__fmul_rn(__fmul_rn(w_score[m], w_score[m]), nc.inv_clss)) - emax) / den;
                            vx += w_cx[m] * wt; vy += w_cy[m] * wt; vz += w_cz[m] * wt;
                        }
                        const float nrm = sqrtf(vx * vx + vy * vy + vz * vz);
                        if (d < nc.max_det) { out[d * 4] = (float)c; out[d * 4 + 1] = vx / nrm; out[d * 4 + 2] = vy / nrm; out[d * 4 + 3] = vz / nrm; }
                        else *overflow = 1;
                    }
                    s_ndet = d + 1;
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        unsigned drop = adj[seed * 8 + w] & rem[w];
                        if (w == (seed >> 5)) drop |= 1u << (seed & 31);
                        left -= __popc(drop & rem[w]);
                        rem[w] &= ~drop;
                    }
                    continue;
                }
#pragma unroll
                for (int w = 0; w < 8; ++w) { cur[w] = adj[seed * 8 + w] & rem[w]; pre[w] = 0; }
                cur[seed >> 5] |= 1u << (seed & 31);
                bool changed = true;
                while (changed) {                                     // :802-806 closure
                    changed = false;
                    for (int w = 0; w < 8; ++w) {
                        unsigned fresh = cur[w] & ~pre[w];
                        pre[w] = cur[w];
                        while (fresh) {
                            const int m = 32 * w + __ffs(fresh) - 1;
                            fresh &= fresh - 1;
                            for (int w2 = 0; w2 < 8; ++w2) {
                                const unsigned add = adj[m * 8 + w2] & rem[w2] & ~cur[w2];
                                if (add) { cur[w2] |= add; changed = true; }
                            }
                        }
                    }
                }
                // vote (:911-919): weights softmax(exp(score^2 / clss_thresh)) over the members
                float emax = -INFINITY;
                for (int w = 0; w < 8; ++w) for (unsigned b = cur[w]; b; b &= b - 1) {
                    const int m = 32 * w + __ffs(b) - 1;
                    emax = fmaxf(emax, expf(__fmul_rn(__fmul_rn(w_score[m], w_score[m]), nc.inv_clss)));
                }
                float den = 0.f;
                for (int w = 0; w < 8; ++w) for (unsigned b = cur[w]; b; b &= b - 1) {
                    const int m = 32 * w + __ffs(b) - 1;
                    den += expf(expf(__fmul_rn(__fmul_rn(w_score[m], w_score[m]), nc.inv_clss)) - emax);
                }
                float vx = 0.f, vy = 0.f, vz = 0.f;
                int members = 0;
                for (int w = 0; w < 8; ++w) for (unsigned b = cur[w]; b; b &= b - 1) {
                    const int m = 32 * w + __ffs(b) - 1;
                    const float wt = expf(expf(__fmul_rn(__fmul_rn(w_score[m], w_score[m]), nc.inv_clss)) - emax) / den;
                    vx += w_cx[m] * wt; vy += w_cy[m] * wt; vz += w_cz[m] * wt;
                    ++members;
                }
                const float nrm = sqrtf(vx * vx + vy * vy + vz * vz);
                const int d = s_ndet;
                if (d < nc.max_det) { out[d * 4] = (float)c; out[d * 4 + 1] = vx / nrm; out[d * 4 + 2] = vy / nrm; out[d * 4 + 3] = vz / nrm; }
                else *overflow = 1;
                s_ndet = d + 1;
#pragma unroll
                for (int w = 0; w < 8; ++w) rem[w] &= ~cur[w];
                left -= members;
            }
        }
        __syncthreads();
    }
    if (tid == 0) count[frame] = min(s_ndet, nc.max_det);
}

size_t nms_shared_bytes(const AssignCfg& c) {
    const size_t NA = (size_t)c.ga * c.ge * c.nb_anchors, C = c.nb_classes;
    return C * NA * 4 * 3 + (C * NA + (C * NA & 1)) * 2 + ((C + 3) & ~3) * 4 + NA * 4 * 7 + NA * 8 * 4 + 64;
}

int launch_yolo_post(const float* logit, long long n_frames, const AssignCfg& cfg, float conf_thresh, float clss_thresh,
                     float unify_thresh, int nms_mode, int max_det, float* det, int32_t* count, int* overflow,
                     cudaStream_t stream) {
    const int NA = cfg.ga * cfg.ge * cfg.nb_anchors;
    if (NA > NMS_MAXA) return set_error(ADY_ERR_UNSUPPORTED, "yolo_post: %d anchors per frame exceed %d", NA, NMS_MAXA);
    if (n_frames <= 0 || max_det <= 0) return set_error(ADY_ERR_INVALID, "yolo_post: empty input");
    if (n_frames > 0x7fffffffLL) return set_error(ADY_ERR_INVALID, "yolo_post: too many frames");
    if (nms_mode < 0 || nms_mode > 2) return set_error(ADY_ERR_INVALID, "yolo_post: nms_mode must be 0 (conn-merge), 1 (soft-merge) or 2 (nms)");
    NmsCfg nc{conf_thresh, clss_thresh, unify_thresh, (float)(90 - 1e-7), 1.0f / clss_thresh, max_det, nms_mode};
    const size_t shmem = nms_shared_bytes(cfg);
    static int configured_dev = -1;
    int dev = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        ADY_CUDA_CHECK(cudaFuncSetAttribute(yolo_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        configured_dev = dev;
    }
    if (shmem > 100 * 1024) return set_error(ADY_ERR_UNSUPPORTED, "yolo_post: configuration needs %zu bytes of shared memory", shmem);
    ADY_CUDA_CHECK(cudaMemsetAsync(overflow, 0, sizeof(int), stream));
    const int threads = ((NA + 31) / 32) * 32;
    yolo_post_kernel<<<(unsigned)n_frames, threads, shmem, stream>>>(logit, cfg, nc, det, count, overflow);
    ADY_LAUNCH_CHECK("yolo_post_kernel");
    return ADY_OK;
}

}  // namespace ady

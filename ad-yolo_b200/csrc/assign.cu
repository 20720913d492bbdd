// AD-YOLO angular-distance responsibility assignment + loss (forward and backward) on sm_100a.
//
// Reference behaviour replaced: /root/reference/src/models/loss.py:156-251 (ADYOLOloss):
//   :193-213 decode (sigmoid / tanh, scale, offset, clamp V, wrap U)
//   :182-187 distance_between_polar_coordinates (great-circle distance in degrees)
//   :222-226 responsibility masks D < {45,25,10} + forced argmin
//   :227-249 obj / class label scatter, 3 x BCE(mean) x 3 thresholds, angular term
//
// Bit-exactness: the mask decisions sit on FP32 rounding noise, so the decode and
// distance chain reproduces torch's *eager op sequence* with individually rounded FP32 ops
// (__fmul_rn/__fadd_rn, never contracted into FMAs) and the same libdevice functions torch's CUDA
// kernels call (tanhf, sinf, cosf, acosf, expf).  Everything downstream of the masks (BCE sums,
// gradients) only has to match to ~1e-5 and is accumulated in FP64.
//
// Launches per call, no host synchronisation:
//   assign_rows_kernel  : 1 thread / target row  -> D, masks, argmin; atomicOr label bits into a
//                         per-anchor 64-bit state word; unnormalised angular gradients
//   loss_anchor_kernel  : warp-autonomous pass over the anchors -> BCE sums, (when asked) d loss / d logit, and the
//                         scalar loss by the last block to finish
//   grad_scale_kernel   : backward of the autograd wrapper: grad *= grad_output unless it is exactly 1
#include "assign_host.h"
#include "common.cuh"

namespace ady {

__device__ __forceinline__ float sigmoid_torch(float x) {
    // ATen sigmoid_kernel_cuda: 1 / (1 + exp(-x)) in float
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}
__device__ __forceinline__ float clamp_torch(float v, float lo, float hi) {
    // ATen clamp: NaN-propagating min(max(v, lo), hi)
    return (v != v) ? v : fminf(fmaxf(v, lo), hi);
}

struct RowGeom {
    int b, t, gi, gj, cls;
    float tu, tv;
    bool ok;
};

__device__ __forceinline__ RowGeom read_row(const float* __restrict__ target, long long m, const AssignCfg& c,
                                            int B, int T) {
    const float* r = target + m * 7;
    RowGeom g;
    g.b = (int)r[0]; g.t = (int)r[1]; g.gi = (int)r[2]; g.gj = (int)r[3]; g.cls = (int)r[4];
    g.tu = r[5]; g.tv = r[6];
    g.ok = g.b >= 0 && g.b < B && g.t >= 0 && g.t < T && g.gi >= 0 && g.gi < c.ga && g.gj >= 0 && g.gj < c.ge &&
           g.cls >= 0 && g.cls < c.nb_classes;
    return g;
}

// Decode + great-circle distance of one anchor (loss.py:196-213, :184-187) with the intermediates
// its backward needs.  noinline: one copy of the libdevice chains in the kernel.
struct AnchorEval {
    float D, thu, thv, Vp, sov, cov, du, cdu, dist, cl;
};
struct ChainK {   // the constants of the chain, by value (a reference to the kernel's AssignCfg would spill it to local memory)
    float ovl_scale, gs_u, gs_v, deg2rad, rad2deg, clip_lo, clip_hi;
};

__device__ __noinline__ AnchorEval eval_anchor(float xu, float xv, const ChainK cfg, float off_u, float off_v,
                                               float tur, float stv, float ctv) {
    AnchorEval e;
    // loss.py:196-198 tanh; :203-205 three separately rounded ops
    e.thu = tanhf(xu); e.thv = tanhf(xv);
    float U = __fadd_rn(__fmul_rn(__fmul_rn(e.thu, cfg.ovl_scale), cfg.gs_u), off_u);
    e.Vp = __fadd_rn(__fmul_rn(__fmul_rn(e.thv, cfg.ovl_scale), cfg.gs_v), off_v);
    const float V = clamp_torch(e.Vp, -90.f, 90.f);       // :208
    if (U >= 180.f) U = __fadd_rn(U, -360.f);              // :209-210
    if (U < -180.f) U = __fadd_rn(U, 360.f);               // :211-212
    // :184-187
    const float our = __fmul_rn(U, cfg.deg2rad), ovr = __fmul_rn(V, cfg.deg2rad);
    e.sov = sinf(ovr); e.cov = cosf(ovr);
    e.du = __fadd_rn(our, -tur);
    e.cdu = cosf(fabsf(e.du));
    e.dist = __fadd_rn(__fmul_rn(e.sov, stv), __fmul_rn(__fmul_rn(e.cov, ctv), e.cdu));
    e.cl = clamp_torch(e.dist, cfg.clip_lo, cfg.clip_hi);
    e.D = __fmul_rn(acosf(e.cl), cfg.rad2deg);
    return e;
}

// backward of the chain (only needs ~1e-5): dD/d(xu), dD/d(xv)
__device__ __forceinline__ float anchor_dD_ddist(const AnchorEval& e, const ChainK& cfg) {
    const float pass = (e.dist >= cfg.clip_lo && e.dist <= cfg.clip_hi) ? 1.f : 0.f;
    // acos backward as ATen evaluates it: grad * -rsqrt(-x*x + 1), two roundings (no FMA):
    // 1 - x^2 cancels badly for small angles, so the op order matters at the 1e-5 level
    const float om = __fadd_rn(1.f, -__fmul_rn(e.cl, e.cl));
    return -cfg.rad2deg * rsqrtf(fmaxf(om, 1e-30f)) * pass;
}
__device__ __forceinline__ float anchor_grad_u(const AnchorEval& e, const ChainK& cfg, float ctv) {
    const float sgn = e.du > 0.f ? 1.f : (e.du < 0.f ? -1.f : 0.f);
    const float ddist_dou = -(e.cov * ctv) * sinf(fabsf(e.du)) * sgn;
    return anchor_dD_ddist(e, cfg) * ddist_dou * (cfg.deg2rad * cfg.ovl_scale) * cfg.gs_u * (1.f - e.thu * e.thu);
}
__device__ __forceinline__ float anchor_grad_v(const AnchorEval& e, const ChainK& cfg, float stv, float ctv) {
    const float ddist_dov = e.cov * stv - e.sov * ctv * e.cdu;
    const float vpass = (e.Vp >= -90.f && e.Vp <= 90.f) ? 1.f : 0.f;
    return anchor_dD_ddist(e, cfg) * ddist_dov * (cfg.deg2rad * cfg.ovl_scale) * cfg.gs_v * vpass * (1.f - e.thv * e.thv);
}

// Block reduction of the few counters of the assignment kernels: warp shuffles, shared memory across the warps, then ONE
// set of atomics per block (all blocks add to the same seven addresses; one set per warp queued 12 k - 80 k same-address
// atomics in L2).
__device__ __forceinline__ void flush_counters(double ang_sum, int ang_cnt, int (&newpos)[ADY_MAX_THR], int bad, int n_thr,
                                               LossAccum* __restrict__ acc) {
    __shared__ double s_sum[32];
    __shared__ int s_cnt[32][2 + ADY_MAX_THR];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        ang_sum += __shfl_xor_sync(0xffffffffu, ang_sum, o);
        ang_cnt += __shfl_xor_sync(0xffffffffu, ang_cnt, o);
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
#pragma unroll
        for (int i = 0; i < ADY_MAX_THR; ++i) newpos[i] += __shfl_xor_sync(0xffffffffu, newpos[i], o);
    }
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) {
        s_sum[w] = ang_sum; s_cnt[w][0] = ang_cnt; s_cnt[w][1] = bad;
#pragma unroll
        for (int i = 0; i < ADY_MAX_THR; ++i) s_cnt[w][2 + i] = newpos[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < nw; ++k) {
            ang_sum += s_sum[k]; ang_cnt += s_cnt[k][0]; bad += s_cnt[k][1];
#pragma unroll
            for (int i = 0; i < ADY_MAX_THR; ++i) newpos[i] += s_cnt[k][2 + i];
        }
        if (ang_cnt) { atomicAdd(&acc->ang_sum, ang_sum); atomicAdd(&acc->ang_cnt, (unsigned long long)ang_cnt); }
        if (bad) atomicAdd(&acc->bad_rows, bad);
        for (int i = 0; i < n_thr; ++i)
            if (newpos[i]) atomicAdd(&acc->n_pos[i], (unsigned long long)newpos[i]);
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int AR_THREADS = 256;
__global__ void __launch_bounds__(AR_THREADS)
assign_rows_kernel(const float* __restrict__ logit, const float* __restrict__ target, long long M_host,
                   const long long* __restrict__ M_dev, int B, int T,
                   AssignCfg cfg, float* __restrict__ D_out, uint8_t* __restrict__ mask_out,
                   int32_t* __restrict__ argmin_out, unsigned long long* __restrict__ state,
                   float2* __restrict__ ang_grad, LossAccum* __restrict__ acc) {
    // M_dev (optional): the true row count lives on the device (label_rows wrote it); the launch
    // then covers the caller's row capacity M_host and surplus threads fall through
    const long long M = M_dev ? min(M_host, *M_dev) : M_host;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int A = cfg.nb_anchors, CH = cfg.nb_classes + 3;
    double ang_sum = 0.0;
    int ang_cnt = 0, newpos[ADY_MAX_THR] = {0, 0, 0, 0}, bad = 0;
    if (m < M) {
        const RowGeom g = read_row(target, m, cfg, B, T);
        if (!g.ok) {
            bad = 1;
            for (int a = 0; a < A; ++a) {
                if (D_out) D_out[m * A + a] = __int_as_float(0x7fc00000);
                if (mask_out) for (int i = 0; i < cfg.n_thr; ++i) mask_out[((long long)i * M + m) * A + a] = 0;
            }
            if (argmin_out) argmin_out[m] = -1;
        } else {
            const long long cell = (((long long)g.b * T + g.t) * cfg.ga + g.gi) * cfg.ge + g.gj;
            const float* lp = logit + cell * A * CH + (cfg.nb_classes + 1);
            const float off_u = cfg.off_u[g.gi], off_v = cfg.off_v[g.gj];
            // target side of loss.py:184-186 (deg2rad, sin, cos) — computed once per row
            const float tur = __fmul_rn(g.tu, cfg.deg2rad), tvr = __fmul_rn(g.tv, cfg.deg2rad);
            const float stv = sinf(tvr), ctv = cosf(tvr);
            // Rolled loops around one non-inlined copy of the chain keep the code small (the 5x unrolled chain with its
            // inlined libdevice calls was 8.2 k SASS instructions and the kernel spent its time on instruction fetch).
            // The logits of all anchors are requested before the first chain starts (one DRAM latency instead of A),
            // and the chain runs once per anchor: its gradient is taken right away (cheap next to the chain itself)
            // and kept for the anchors that turn out to enter the angular term.
            const ChainK ck{cfg.ovl_scale, cfg.gs_u, cfg.gs_v, cfg.deg2rad, cfg.rad2deg, cfg.clip_lo, cfg.clip_hi};
            float xu[ADY_MAX_ANCHORS], xv[ADY_MAX_ANCHORS];
#pragma unroll
            for (int a = 0; a < ADY_MAX_ANCHORS; ++a)
                if (a < A) { xu[a] = lp[a * CH]; xv[a] = lp[a * CH + 1]; }
            float Dv[ADY_MAX_ANCHORS], gu[ADY_MAX_ANCHORS], gv[ADY_MAX_ANCHORS];
            float dmin = 0.f;
            int amin = 0;
#pragma unroll 1
            for (int a = 0; a < A; ++a) {
                const AnchorEval e = eval_anchor(xu[a], xv[a], ck, off_u, off_v, tur, stv, ctv);
                Dv[a] = e.D;
                if (ang_grad) { gu[a] = anchor_grad_u(e, ck, ctv); gv[a] = anchor_grad_v(e, ck, stv, ctv); }
                if (a == 0 || e.D < dmin) { dmin = e.D; amin = a; }   // first minimum (torch.min)
            }
            if (argmin_out) argmin_out[m] = amin;
            unsigned long long bits_a[ADY_MAX_ANCHORS];
#pragma unroll 1
            for (int a = 0; a < A; ++a) {
                const float Da = Dv[a];
                if (D_out) D_out[m * A + a] = Da;
                unsigned long long bits = 0ull;
                for (int i = 0; i < cfg.n_thr; ++i) {
                    const bool resp = (Da < cfg.thr[i]) || (a == amin);   // :223-224
                    if (mask_out) mask_out[((long long)i * M + m) * A + a] = resp ? 1 : 0;
                    if (resp) bits |= (1ull | (2ull << g.cls)) << (16 * i);
                    if (i == 0 && resp) {                                    // :244-246 angular term
                        ang_sum += (double)Da;
                        ang_cnt += 1;
                        if (ang_grad) {
                            atomicAdd(&ang_grad[cell * A + a].x, gu[a]);
                            atomicAdd(&ang_grad[cell * A + a].y, gv[a]);
                        }
                    }
                }
                bits_a[a] = bits;
            }
            if (state) {
                // the label words of all anchors are updated before the first returned value is looked at: A atomics
                // in flight instead of A round trips in a row (27 % of the kernel's stall samples)
                unsigned long long old_a[ADY_MAX_ANCHORS];
#pragma unroll
                for (int a = 0; a < ADY_MAX_ANCHORS; ++a)
                    if (a < A) old_a[a] = bits_a[a] ? atomicOr(&state[cell * A + a], bits_a[a]) : 0ull;
#pragma unroll
                for (int a = 0; a < ADY_MAX_ANCHORS; ++a)
                    if (a < A) {
                        const unsigned long long fresh = bits_a[a] & ~old_a[a];
                        for (int i = 0; i < cfg.n_thr; ++i) newpos[i] += (int)((fresh >> (16 * i)) & 1ull);
                    }
            }
        }
    }
    if (!acc) return;
    flush_counters(ang_sum, ang_cnt, newpos, bad, cfg.n_thr, acc);
}

// (Measured and rejected, round 2: one LANE per (row, anchor) -- groups of 8 lanes, the first-minimum scan replayed on the
//  gathered distances, bit-identical results -- runs the five libdevice chains of a row side by side on 2.2 waves of
//  threads instead of 0.45, but issues 60 % more warp-level chain evaluations (20 of 32 lanes active) and repeats the
//  target-side sin / cos per lane: 30.6 us against 21.4 us for the kernel above.)
static void launch_assign_kernel(const float* logit, const float* target, long long M, const long long* M_dev, int B, int T,
                                 const AssignCfg& cfg, float* D, uint8_t* mask, int32_t* argmin, unsigned long long* state,
                                 float2* ang, LossAccum* acc, cudaStream_t stream) {
    const int blocks = (int)((M + AR_THREADS - 1) / AR_THREADS);
    assign_rows_kernel<<<blocks, AR_THREADS, 0, stream>>>(logit, target, M, M_dev, B, T, cfg, D, mask, argmin, state, ang, acc);
}

// ------------------------------------------------------------------------------------------------
// BCE pieces as ATen defines them (Loss.cu binary_cross_entropy_out_cuda / _backward):
//   loss = (t-1) * max(log1p(-p), -100) - t * max(log(p), -100)
//   dL/dp = (p - t) / max((1-p) p, 1e-12);   dp/dx = (1-p) p
// evaluated with the fast MUFU paths (ex2 / rcp / lg2): these terms only enter sums and
// gradients gated at 1e-5 relative (the bit-exact part is the assignment above); saturation
// behaves identically (p == 1 -> log(0) = -inf -> clamped to -100).
// plain division: __fdividef returns 0 for denominators above 2^126 (logits below about -87), which would turn the
// BCE term into the -100 clamp instead of ATen's ~-88
__device__ __forceinline__ float sigmoid_fast(float x) { return 1.0f / (1.0f + __expf(-x)); }
// log for the BCE: MUFU.LG2 flushes subnormal arguments to zero, ATen's logf does not -- sigmoid of a logit in
// (-103, -87) is subnormal and ATen's term is ~-88..-100 rather than the -100 clamp (never taken in practice)
__device__ __forceinline__ float log_bce(float p) {
    return (p > 0.f && p < 1.17549435e-38f) ? logf(p) : __logf(p);
}
__device__ __forceinline__ float bce_fwd(float p, float t) {
    const float lp = fmaxf(log_bce(p), -100.f), l1 = fmaxf(__logf(1.0f - p), -100.f);
    return (t - 1.f) * l1 - t * lp;
}
__device__ __forceinline__ float bce_bwd_logit(float p, float t) {
    const float q = (1.f - p) * p;
    return q >= 1e-12f ? (p - t) : (p - t) * q * 1e12f;
}

__device__ __forceinline__ float bce_fwd_pos(float p) { return -fmaxf(log_bce(p), -100.f); }          // bce_fwd(p, 1)
__device__ __forceinline__ float bce_fwd_neg(float p) { return -fmaxf(__logf(1.0f - p), -100.f); }   // bce_fwd(p, 0)

constexpr int LA_THREADS = 256;
#ifndef ADY_LA_MINB
#define ADY_LA_MINB 4
#endif

struct LossW {
    float w_pos[ADY_MAX_THR], w_neg[ADY_MAX_THR], w_cls[ADY_MAX_THR], w_ang, w_neg_sum;
};

// Normalisers of the means of loss.py:236-249 from the final counts, once per block (FP64 divisions by every thread
// were 20 % of the kernel's stall samples; a separate one-block launch cost 3 us of the step), times the upstream
// gradient (device scalar).
__device__ __forceinline__ void loss_weights_block(long long n_anchor, const AssignCfg& cfg, const LossAccum* acc, float gs,
                                                   LossW* sw) {
    const int i = threadIdx.x;
    if (i < ADY_MAX_THR) {
        const double np_ = i < cfg.n_thr ? (double)acc->n_pos[i] : 0.0;
        const double nn_ = (double)n_anchor - np_;
        sw->w_pos[i] = i < cfg.n_thr ? gs * (float)(cfg.gain_obj / (cfg.n_thr * np_)) : 0.f;
        sw->w_neg[i] = i < cfg.n_thr ? gs * (float)(cfg.gain_nonobj / (cfg.n_thr * nn_)) : 0.f;
        sw->w_cls[i] = i < cfg.n_thr ? gs * (float)(cfg.gain_cls / (cfg.n_thr * np_ * cfg.nb_classes)) : 0.f;
    }
    if (i == 32) {
        sw->w_ang = gs * (float)(cfg.gain_ang / (180.0 * (double)acc->ang_cnt));
        float t = 0.f;                                  // gradient weight of an anchor that is negative at every threshold
        for (int k = 0; k < cfg.n_thr; ++k) {
            const double nn_ = (double)n_anchor - (double)acc->n_pos[k];
            t += gs * (float)(cfg.gain_nonobj / (cfg.n_thr * nn_));
        }
        sw->w_neg_sum = t;
    }
}

__device__ __forceinline__ double ld_fresh_f64(const double* p) {     // a value other blocks of this launch accumulated
    return __longlong_as_double((long long)atomicAdd(reinterpret_cast<unsigned long long*>(const_cast<double*>(p)), 0ull));
}

// loss.py:219, 244-249 (means of empty sets are NaN exactly like torch)
__device__ __forceinline__ float loss_total(long long n_anchor, const AssignCfg& cfg, const LossAccum* acc) {
    double total = cfg.gain_ang * (acc->ang_sum / 180.0) / (double)acc->ang_cnt;
    for (int i = 0; i < cfg.n_thr; ++i) {
        const double np_ = (double)acc->n_pos[i], nn_ = (double)n_anchor - np_;
        const double pos = ld_fresh_f64(&acc->s_pos[i]) / np_, neg = ld_fresh_f64(&acc->s_neg[i]) / nn_;
        const double cls = ld_fresh_f64(&acc->s_cls[i]) / (np_ * cfg.nb_classes);
        total += (pos * cfg.gain_obj + neg * cfg.gain_nonobj + cls * cfg.gain_cls) / cfg.n_thr;
    }
    return (float)total;
}

// Warp-autonomous pass over the anchors (no block barriers in the loop); a warp owns groups of 32 consecutive anchors
// (the group base is 16-byte aligned for any channel count):
//   1. lane = anchor: objectness logit -> sigmoid -> BCE sums and the objectness gradient.  The label word and the
//      logit of the next group are in flight while the current one is processed.  An anchor without any positive bit
//      (94 % of them) needs one log and no per-threshold loop: its BCE(p, 0) goes to one accumulator that is added to
//      every threshold's negative sum at the end.
//   2. positive anchors of the group, two per iteration (peeled off the ballot mask with ffs), 16 lanes each, lane =
//      class: class BCE sums / gradients.  No rank -> lane search, no division.
//   3. (GRAD) the group's gradient rows are assembled in a per-warp shared-memory tile that is zero except for the
//      objectness / (u, v) / positive-class entries; it is copied out with coalesced streaming float4 stores and every
//      float4 is zeroed again right after it was read.
// The last block to finish (ticket) turns the sums into the scalar loss: no finalisation launch.
// do_sums: accumulate the BCE sums (forward); GRAD: write gscale * d loss / d logit.
template <bool GRAD>
__global__ void __launch_bounds__(LA_THREADS, ADY_LA_MINB)
loss_anchor_kernel(const float* __restrict__ logit, long long n_anchor, AssignCfg cfg,
                   const unsigned long long* __restrict__ state, const float2* __restrict__ ang_grad,
                   LossAccum* __restrict__ acc, float* __restrict__ grad, const float* __restrict__ gscale,
                   int do_sums, float* __restrict__ loss_out) {
    const unsigned CH = cfg.nb_classes + 3, C = cfg.nb_classes;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const unsigned long long OBJ_ANY = 0x0001000100010001ull;

    __shared__ LossW sw;
    extern __shared__ __align__(16) float s_tile_all[];
    float* tile = s_tile_all + (threadIdx.x >> 5) * (32 * CH);
    if (GRAD) {
        loss_weights_block(n_anchor, cfg, acc, gscale ? gscale[0] : 1.0f, &sw);
        for (unsigned i = lane; i < 32 * CH; i += 32) tile[i] = 0.f;
        __syncthreads();
    }
    const float w_neg_sum = GRAD ? sw.w_neg_sum : 0.f, w_ang = GRAD ? sw.w_ang : 0.f;

    // per-thread partial sums stay in FP32 (a thread sees only tens of anchors); they are widened
    // to FP64 for the warp reduction and the global accumulation
    float s_pos[ADY_MAX_THR] = {0, 0, 0, 0}, s_neg[ADY_MAX_THR] = {0, 0, 0, 0}, s_cls[ADY_MAX_THR] = {0, 0, 0, 0};
    float s_neg_all = 0.f;

    // Work distribution: a warp takes its first rounds of groups statically (group = warp + k * n_warps) and claims the
    // last quarter from a counter, one group at a time: the time of a group depends on how many positives it holds, and
    // with a purely static split a quarter of all warp time was spent waiting for the slowest warp at the final barrier.
    // The group after the current one is always known (and its label words / objectness logits in flight) while the
    // current one is processed; the claim for the one after that is issued a whole group ahead of its use.
    const long long n_groups = (n_anchor + 31) / 32;
    const long long warp0 = ((long long)blockIdx.x * LA_THREADS + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * LA_THREADS) >> 5;
    const long long k_static = (n_groups / n_warps) * 3 / 4;
    const long long dyn_base = k_static * n_warps;
    // the dynamic range is cut into ADY_LOSS_NCTR contiguous parts, each with its own counter (in its own 128-byte line)
    // and served by the blocks with blockIdx % ADY_LOSS_NCTR == part: one counter for all 4.7 k warps queued the claims
    // behind each other in one L2 slice (20 % of the kernel's stall samples)
    const int n_parts = min((int)gridDim.x, ADY_LOSS_NCTR);
    const long long dyn_part = (n_groups - dyn_base + n_parts - 1) / n_parts;
    const int my_part = blockIdx.x % n_parts;
    const long long part_base = dyn_base + my_part * dyn_part;
    long long k_round = -1;
    auto issue_claim = [&]() -> long long {              // the value is meaningful in lane 0 (static rounds: in every lane)
        ++k_round;
        if (k_round < k_static) return warp0 + k_round * n_warps;
        long long v = 0;
        if (lane == 0) {
            const long long t = (long long)atomicAdd(&acc->group_ctr[my_part][0], 1u);
            v = (t < dyn_part && part_base + t < n_groups) ? part_base + t : n_groups;
        }
        return v;
    };
    unsigned long long st_n = 0ull;
    float x_n = 0.f;
    auto prefetch = [&](long long g) {
        const long long an = g * 32 + lane;
        if (g < n_groups && an < n_anchor) { st_n = state[an]; x_n = logit[an * CH]; }
        else { st_n = 0ull; x_n = 0.f; }
    };
    long long grp = __shfl_sync(FULL, issue_claim(), 0);
    long long raw_next = issue_claim();
    prefetch(grp);
    while (grp < n_groups) {
        const long long a0 = grp * 32;
        const unsigned na = (unsigned)min(32LL, n_anchor - a0);
        const unsigned nel = na * CH;
        const float* src = logit + a0 * CH;
        const bool valid = (unsigned)lane < na;
        const unsigned long long st = st_n;
        const float x_obj = x_n;
        const long long nxt = __shfl_sync(FULL, raw_next, 0);
        prefetch(nxt);
        raw_next = issue_claim();

        // everything this group will wait for is requested before the arithmetic starts: the angular partials of
        // the first threshold's positives and the class logits of the first two positive anchors
        const bool positive = valid && (st & OBJ_ANY);
        unsigned m = __ballot_sync(FULL, positive);
        const unsigned c = lane & 15;
        auto peel = [&]() -> int {                       // next two positive anchors off the mask -> this half-warp's anchor
            const int al0 = __ffs(m) - 1;
            m &= m - 1;
            const int al1 = m ? __ffs(m) - 1 : -1;
            m &= m - 1;                                  // (0 & anything = 0)
            return (lane & 16) ? al1 : al0;
        };
        int al = -1;
        float xc = 0.f;
        bool have = m != 0;                              // warp-uniform: a pair (or a single anchor) is pending in (al, xc)
        if (have) {
            al = peel();
            if (al >= 0 && c < C) xc = src[(unsigned)al * CH + 1 + c];
        }
        float2 ag = make_float2(0.f, 0.f);
        if (GRAD && positive && (st & 1ull)) ag = ang_grad[a0 + lane];   // the angular term lives on the first threshold's positives

        // ---- pass 1: objectness, lane = anchor
        if (valid) {
            const float p = sigmoid_fast(x_obj);
            const float g_neg = bce_bwd_logit(p, 0.f);
            float l_neg = 0.f;
            if (do_sums) l_neg = bce_fwd_neg(p);
            float go;
            if (!positive) {
                s_neg_all += l_neg;
                go = w_neg_sum * g_neg;
            } else {
                const float g_pos = bce_bwd_logit(p, 1.f);
                float l_pos = 0.f;
                if (do_sums) l_pos = bce_fwd_pos(p);
                go = 0.f;
#pragma unroll
                for (int i = 0; i < ADY_MAX_THR; ++i) {
                    if (i >= cfg.n_thr) break;
                    if ((st >> (16 * i)) & 1ull) { s_pos[i] += l_pos; if (GRAD) go += sw.w_pos[i] * g_pos; }
                    else                         { s_neg[i] += l_neg; if (GRAD) go += sw.w_neg[i] * g_neg; }
                }
                if (GRAD && (st & 1ull)) {
                    tile[(unsigned)lane * CH + C + 1] = ag.x * w_ang;
                    tile[(unsigned)lane * CH + C + 2] = ag.y * w_ang;
                }
            }
            if (GRAD) tile[(unsigned)lane * CH] = go;
        }

        // ---- pass 2: the positive anchors, two per iteration, lane = (which of the two, class); the logits of the
        // next pair are requested before the current pair is evaluated
        while (have) {
            const int al_c = al;
            const float x_c = xc;
            al = -1;
            have = m != 0;
            if (have) {
                al = peel();
                if (al >= 0 && c < C) xc = src[(unsigned)al * CH + 1 + c];
            }
            const unsigned long long sta = __shfl_sync(FULL, st, al_c & 31);
            if (al_c >= 0 && c < C) {
                const float pc = sigmoid_fast(x_c);
                float l1 = 0.f, l0 = 0.f;
                if (do_sums) { l1 = bce_fwd_pos(pc); l0 = bce_fwd_neg(pc); }
                const float g1 = bce_bwd_logit(pc, 1.f), g0 = bce_bwd_logit(pc, 0.f);
                float gc = 0.f;
#pragma unroll
                for (int i = 0; i < ADY_MAX_THR; ++i) {
                    if (i >= cfg.n_thr) break;
                    const unsigned f = (unsigned)(sta >> (16 * i)) & 0xffffu;   // this threshold's (class bits | obj bit)
                    if (f & 1u) {
                        const bool t = (f >> (1 + c)) & 1u;
                        s_cls[i] += t ? l1 : l0;
                        if (GRAD) gc += sw.w_cls[i] * (t ? g1 : g0);
                    }
                }
                if (GRAD) tile[(unsigned)al_c * CH + 1 + c] = gc;
            }
        }

        // ---- pass 3 (GRAD): coalesced float4 copy-out of the tile; each float4 is zeroed again once it is in registers
        if (GRAD) {
            float* dst = grad + a0 * CH;
            __syncwarp();
            if (na == 32) {                              // nel = 32 CH: a whole number of float4s, at most 5 per lane (CH <= 18)
#pragma unroll
                for (int it = 0; it < 5; ++it) {
                    const unsigned el = (unsigned)lane * 4 + it * 128;
                    if (el < nel) {
                        const float4 v = *reinterpret_cast<const float4*>(tile + el);
                        *reinterpret_cast<float4*>(tile + el) = make_float4(0.f, 0.f, 0.f, 0.f);
                        __stcs(reinterpret_cast<float4*>(dst + el), v);
                    }
                }
            } else {                                     // last, partial group
                for (unsigned el = lane; el < nel; el += 32) { dst[el] = tile[el]; tile[el] = 0.f; }
            }
            __syncwarp();
        }
        grp = nxt;
    }
    // warp reduce (FP64) -> block reduce through shared memory -> one atomic per block and sum:
    // per-warp atomics on a handful of addresses serialise in L2 (10^5 of them cost ~10^2 us)
    __shared__ double s_red[LA_THREADS / 32][3 * ADY_MAX_THR];
    __shared__ bool s_last;
    if (do_sums) {
        double d[3 * ADY_MAX_THR];
#pragma unroll
        for (int i = 0; i < ADY_MAX_THR; ++i) {
            d[i] = s_pos[i];
            d[ADY_MAX_THR + i] = i < cfg.n_thr ? (double)s_neg[i] + (double)s_neg_all : 0.0;
            d[2 * ADY_MAX_THR + i] = s_cls[i];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1)
#pragma unroll
            for (int i = 0; i < 3 * ADY_MAX_THR; ++i) d[i] += __shfl_xor_sync(FULL, d[i], o);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 3 * ADY_MAX_THR; ++i) s_red[threadIdx.x >> 5][i] = d[i];
        }
        __syncthreads();
        if (threadIdx.x < 3 * ADY_MAX_THR) {
            double t = 0.0;
            for (int w = 0; w < LA_THREADS / 32; ++w) t += s_red[w][threadIdx.x];
            const int kind = threadIdx.x / ADY_MAX_THR, i = threadIdx.x % ADY_MAX_THR;
            if (i < cfg.n_thr && t != 0.0) {
                double* dstp = kind == 0 ? acc->s_pos : (kind == 1 ? acc->s_neg : acc->s_cls);
                atomicAdd(&dstp[i], t);
            }
            __threadfence();
        }
    }
    // the last block to finish (ticket) writes the scalar loss and leaves the two scheduling counters at zero for a
    // later backward launch over the same workspace
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&acc->done_blocks, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        if (do_sums && loss_out) loss_out[0] = loss_total(n_anchor, cfg, acc);
        for (int j = 0; j < ADY_LOSS_NCTR; ++j) acc->group_ctr[j][0] = 0u;
        acc->done_blocks = 0u;
    }
}

template <bool GRAD>
static int loss_anchor_blocks(long long n_anchor, size_t shmem, int* blocks_out) {
    static int per_sm = 0;
    int dev = 0, sms = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    ADY_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (!per_sm) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, loss_anchor_kernel<GRAD>, LA_THREADS, shmem) != cudaSuccess || n < 1) n = 2;
        per_sm = n;
    }
    long long blocks = (n_anchor + LA_THREADS - 1) / LA_THREADS;
    if (blocks > (long long)sms * per_sm) blocks = (long long)sms * per_sm;   // one resident wave
    *blocks_out = (int)blocks;
    return ADY_OK;
}

// ------------------------------------------------------------------------------------------------
size_t loss_workspace_bytes(int B, int T, const AssignCfg& cfg) {
    const size_t n_anchor = (size_t)B * T * cfg.ga * cfg.ge * cfg.nb_anchors;
    return sizeof(LossAccum) + n_anchor * (sizeof(unsigned long long) + sizeof(float2));
}

static int check_cfg(const AssignCfg& c) {
    if (c.nb_anchors < 1 || c.nb_anchors > ADY_MAX_ANCHORS) return set_error(ADY_ERR_INVALID, "nb_anchors must be 1..%d", ADY_MAX_ANCHORS);
    if (c.n_thr < 1 || c.n_thr > ADY_MAX_THR) return set_error(ADY_ERR_INVALID, "train_unify must have 1..%d entries", ADY_MAX_THR);
    if (c.nb_classes < 1 || c.nb_classes > 15) return set_error(ADY_ERR_INVALID, "nb_classes must be 1..15");
    if (c.ga < 1 || c.ga > ADY_MAX_GRID || c.ge < 1 || c.ge > ADY_MAX_GRID) return set_error(ADY_ERR_INVALID, "grid too large");
    return ADY_OK;
}

int launch_assign(const float* logit, const float* target, long long M, int B, int T, const AssignCfg& cfg,
                  float* D, uint8_t* mask, int32_t* argmin, cudaStream_t stream) {
    // (mask layout (n_thr, M, A) uses the host M: no device-count variant here)
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (M <= 0) return ADY_OK;
    if (M / AR_THREADS + 1 > 0x7fffffffLL) return set_error(ADY_ERR_INVALID, "adyolo_assign: too many rows");
    launch_assign_kernel(logit, target, M, nullptr, B, T, cfg, D, mask, argmin, nullptr, nullptr, nullptr, stream);
    ADY_LAUNCH_CHECK("assign_rows_kernel");
    return ADY_OK;
}

int launch_loss(const float* logit, const float* target, long long M, const long long* M_dev, int B, int T,
                const AssignCfg& cfg, float* loss_out, float* grad_out, float* D, uint8_t* mask, int32_t* argmin,
                void* ws, cudaStream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    const long long n_anchor = (long long)B * T * cfg.ga * cfg.ge * cfg.nb_anchors;
    if (n_anchor <= 0) return set_error(ADY_ERR_INVALID, "adyolo_loss: empty logit");
    LossAccum* acc = reinterpret_cast<LossAccum*>(ws);
    unsigned long long* state = reinterpret_cast<unsigned long long*>(acc + 1);
    float2* ang = reinterpret_cast<float2*>(state + n_anchor);
    ADY_CUDA_CHECK(cudaMemsetAsync(ws, 0, loss_workspace_bytes(B, T, cfg), stream));
    if (M > 0) {
        if (M / AR_THREADS + 1 > 0x7fffffffLL) return set_error(ADY_ERR_INVALID, "adyolo_loss: too many rows");
        launch_assign_kernel(logit, target, M, M_dev, B, T, cfg, D, mask, argmin, state, ang, acc, stream);
        ADY_LAUNCH_CHECK("assign_rows_kernel");
    }
    int blocks = 0;
    if (grad_out) {      // training: gradient, sums and the scalar loss in one launch
        if (reinterpret_cast<uintptr_t>(grad_out) & 15) return set_error(ADY_ERR_INVALID, "adyolo_loss: grad must be 16-byte aligned");
        const size_t shmem = (size_t)(LA_THREADS / 32) * 32 * (cfg.nb_classes + 3) * sizeof(float);
        rc = loss_anchor_blocks<true>(n_anchor, shmem, &blocks);
        if (rc) return rc;
        loss_anchor_kernel<true><<<blocks, LA_THREADS, shmem, stream>>>(logit, n_anchor, cfg, state, ang, acc, grad_out, nullptr, 1, loss_out);
    } else {
        rc = loss_anchor_blocks<false>(n_anchor, 0, &blocks);
        if (rc) return rc;
        loss_anchor_kernel<false><<<blocks, LA_THREADS, 0, stream>>>(logit, n_anchor, cfg, state, ang, acc, nullptr, nullptr, 1, loss_out);
    }
    ADY_LAUNCH_CHECK("loss_anchor_kernel");
    return ADY_OK;
}

// Backward from the workspace a forward call left behind (label bits, counts, angular partials):
// grad_out = grad_output[0] * d loss / d logit, one pass over the logits.
int launch_loss_backward(const float* logit, int B, int T, const AssignCfg& cfg, const void* ws,
                         const float* grad_output, float* grad_out, cudaStream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    const long long n_anchor = (long long)B * T * cfg.ga * cfg.ge * cfg.nb_anchors;
    if (n_anchor <= 0) return set_error(ADY_ERR_INVALID, "adyolo_loss_backward: empty logit");
    LossAccum* acc = const_cast<LossAccum*>(reinterpret_cast<const LossAccum*>(ws));
    const unsigned long long* state = reinterpret_cast<const unsigned long long*>(acc + 1);
    const float2* ang = reinterpret_cast<const float2*>(state + n_anchor);
    if (reinterpret_cast<uintptr_t>(grad_out) & 15) return set_error(ADY_ERR_INVALID, "adyolo_loss_backward: grad must be 16-byte aligned");
    const size_t shmem = (size_t)(LA_THREADS / 32) * 32 * (cfg.nb_classes + 3) * sizeof(float);
    int blocks = 0;
    rc = loss_anchor_blocks<true>(n_anchor, shmem, &blocks);
    if (rc) return rc;
    loss_anchor_kernel<true><<<blocks, LA_THREADS, shmem, stream>>>(logit, n_anchor, cfg, state, ang, acc, grad_out, grad_output, 0, nullptr);
    ADY_LAUNCH_CHECK("loss_anchor_kernel(backward)");
    return ADY_OK;
}

// grad *= gscale[0] in place; nothing to do (and nothing touched) for the usual upstream gradient 1
__global__ void __launch_bounds__(256) grad_scale_kernel(float* __restrict__ g, long long n, const float* __restrict__ gscale) {
    const float s = gscale[0];
    if (s == 1.0f) return;
    const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
    float4* g4 = reinterpret_cast<float4*>(g);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = g4[i];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        g4[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) g[(n4 << 2) + threadIdx.x] *= s;
}

int launch_grad_scale(float* grad, long long n, const float* gscale, cudaStream_t stream) {
    if (n <= 0) return ADY_OK;
    if (reinterpret_cast<uintptr_t>(grad) & 15) return set_error(ADY_ERR_INVALID, "adyolo_loss_grad_scale: grad must be 16-byte aligned");
    int dev = 0, sms = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    ADY_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long blocks = ((n >> 2) + 255) / 256;
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    if (blocks < 1) blocks = 1;
    grad_scale_kernel<<<(int)blocks, 256, 0, stream>>>(grad, n, gscale);
    ADY_LAUNCH_CHECK("grad_scale_kernel");
    return ADY_OK;
}

}  // namespace ady

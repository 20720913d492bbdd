// Grid-cell responsibility assignment of AD-YOLO labels on the GPU (integer/exact part 1).
//
// Reference behaviour replaced: /root/reference/src/datasets.py:457-482 (get_yolo_label) with the
// grid constants of :219-235, plus the label half of collate_fn (:175-184): one output row
// [batch, frame, Gi, Gj, class, U, V] (float32) per responsible cell, events in input order and
// cells in row-major (Gi, Gj) order — the order numpy.where produces.
//
// All comparisons are done in float64 exactly as numpy does them (bounds are multiples of 22.5,
// the +-360 wrap terms are exact), so the result is bit-identical by construction.
//   cells kernel : per event -> 32-bit cell mask (Ga*Ge <= 32) and block-local exclusive scan
//   scan kernel  : exclusive scan of the block totals (single block)
//   rows kernel  : expands masks into rows at the scanned offsets
#include "assign_host.h"
#include "common.cuh"

namespace ady {

constexpr int LB = 256;  // threads per block in the cells / rows kernels

// Label half of the rotation augmentation (utils/augmentations.py:98-109): pi' = pi*pi_weight + d_pi
// wrapped once into [-180, 180], theta' = theta*theta_weight; combination per clip (batch index).
__device__ __forceinline__ void rotate_label(int comb, double& azi, double& ele) {
    // pi_weight: -1 for combinations 2,3,6,7,10,11,14,15; d_pi: 0,0,0,0,180,180,180,180,90,90,90,90,-90,...
    // theta_weight: -1 for odd combinations
    const double pw = (comb & 2) ? -1.0 : 1.0;
    const int q = (comb >> 2) & 3;
    const double dpi = q == 0 ? 0.0 : (q == 1 ? 180.0 : (q == 2 ? 90.0 : -90.0));
    const double tw = (comb & 1) ? -1.0 : 1.0;
    double pi = azi * pw + dpi;
    if (pi < -180.0) pi += 360.0;
    else if (pi > 180.0) pi -= 360.0;
    azi = pi;
    ele = ele * tw;
}

__global__ void __launch_bounds__(LB)
label_cells_kernel(const double* __restrict__ ev, long long E, int nlf, CellCfg cfg, const int8_t* __restrict__ rot, long long n_rot,
                   uint32_t* __restrict__ cellmask, uint32_t* __restrict__ local_off, long long* __restrict__ block_tot) {
    __shared__ uint32_t wsum[LB / 32];
    const long long e = (long long)blockIdx.x * LB + threadIdx.x;
    uint32_t mask = 0;
    if (e < E) {
        const double frame = ev[e * 5 + 1];
        double azi = ev[e * 5 + 3];
        double ele = ev[e * 5 + 4];
        if (rot) {   // batch ids outside rot_comb (caller bug) are left unrotated instead of read out of bounds
            const long long bi = (long long)ev[e * 5];
            if (bi >= 0 && bi < n_rot) rotate_label(rot[bi], azi, ele);
        }
        if (azi == 180.0) azi = -180.0;                                  // :470
        if (frame < (double)nlf) {                                       // :468
            uint32_t am = 0, em = 0;
            for (int i = 0; i < cfg.ga; ++i) {
                const bool r = ((cfg.lb_a[i] <= azi) && (azi < cfg.ub_a[i])) ||   // :472
                               (azi + 360.0 < cfg.ub_a[i]) ||                      // :475
                               (cfg.lb_a[i] < azi - 360.0);                        // :476
                am |= (uint32_t)r << i;
            }
            for (int j = 0; j < cfg.ge; ++j)
                em |= (uint32_t)((cfg.lb_e[j] <= ele) && (ele < cfg.ub_e[j])) << j;  // :473
            for (int i = 0; i < cfg.ga; ++i)
                if ((am >> i) & 1u) mask |= em << (i * cfg.ge);
        }
        cellmask[e] = mask;
    }
    // block exclusive scan of popcounts
    const uint32_t cnt = __popc(mask);
    uint32_t inc = cnt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    if (e < E) local_off[e] = base + inc - cnt;
    if (threadIdx.x == LB - 1) block_tot[blockIdx.x] = (long long)(base + inc);
}

__global__ void label_scan_kernel(long long* __restrict__ block_tot, long long nblocks, long long* __restrict__ total) {
    // single block, sequential chunks of 1024 (nblocks = E/256: tiny)
    __shared__ long long sh[1024];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < nblocks; base += 1024) {
        const long long i = base + threadIdx.x;
        const long long v = i < nblocks ? block_tot[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblocks) block_tot[i] = carry + sh[threadIdx.x] - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// INLINE_SCAN: block_off holds the per-block totals as label_cells_kernel left them; every block adds up the totals of
// the blocks before it (nblocks = E / 256 is small: ~75 values at config[1]) and the last block publishes the row count,
// which saves the one-block scan launch between the two kernels.
template <bool INLINE_SCAN>
__global__ void __launch_bounds__(LB)
label_rows_kernel(const double* __restrict__ ev, long long E, CellCfg cfg, const int8_t* __restrict__ rot, long long n_rot,
                  const uint32_t* __restrict__ cellmask, const uint32_t* __restrict__ local_off,
                  const long long* __restrict__ block_off, float* __restrict__ rows, long long max_rows,
                  long long* __restrict__ total) {
    long long base;
    if (INLINE_SCAN) {
        __shared__ long long s_part[LB / 32];
        __shared__ long long s_base;
        long long t = 0;
        for (long long i = threadIdx.x; i < (long long)blockIdx.x; i += LB) t += block_off[i];
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long b = 0;
            for (int w = 0; w < LB / 32; ++w) b += s_part[w];
            s_base = b;
            if (blockIdx.x == gridDim.x - 1) *total = b + block_off[blockIdx.x];
        }
        __syncthreads();
        base = s_base;
    } else {
        base = block_off[blockIdx.x];
    }
    const long long e = (long long)blockIdx.x * LB + threadIdx.x;
    if (e >= E) return;
    uint32_t mask = cellmask[e];
    if (!mask) return;
    long long o = base + local_off[e];
    double azi = ev[e * 5 + 3], ele = ev[e * 5 + 4];
    if (rot) {
        const long long bi = (long long)ev[e * 5];
        if (bi >= 0 && bi < n_rot) rotate_label(rot[bi], azi, ele);
    }
    if (azi == 180.0) azi = -180.0;
    const float fb = (float)ev[e * 5 + 0], ff = (float)ev[e * 5 + 1], fc = (float)ev[e * 5 + 2];
    const float fu = (float)azi, fv = (float)ele;
    while (mask) {
        const int cell = __ffs(mask) - 1;   // ascending cell index == row-major (Gi, Gj)
        mask &= mask - 1;
        if (o < max_rows) {
            float* r = rows + o * 7;
            r[0] = fb; r[1] = ff; r[2] = (float)(cell / cfg.ge); r[3] = (float)(cell % cfg.ge);
            r[4] = fc; r[5] = fu; r[6] = fv;
        }
        ++o;
    }
}

size_t label_workspace_bytes(long long E) {
    const long long nb = (E + LB - 1) / LB;
    return (size_t)(E * 4 + (nb + 1) * 8 + 64);
}

int launch_label_cells(const double* events, long long E, int nb_label_frames, const CellCfg& cfg, const int8_t* rot, long long n_rot,
                       uint32_t* cellmask, long long* total_rows_dev, void* ws, cudaStream_t stream) {
    if (cfg.ga * cfg.ge > 32 || cfg.ga > ADY_MAX_GRID || cfg.ge > ADY_MAX_GRID)
        return set_error(ADY_ERR_UNSUPPORTED, "label cells: grid %dx%d exceeds the 32-cell mask", cfg.ga, cfg.ge);
    if (E <= 0) {
        ADY_CUDA_CHECK(cudaMemsetAsync(total_rows_dev, 0, 8, stream));
        return ADY_OK;
    }
    const long long nb = (E + LB - 1) / LB;
    uint32_t* local_off = reinterpret_cast<uint32_t*>(ws);
    long long* block_tot = reinterpret_cast<long long*>(reinterpret_cast<char*>(ws) + ((E * 4 + 7) / 8) * 8);
    label_cells_kernel<<<(int)nb, LB, 0, stream>>>(events, E, nb_label_frames, cfg, rot, n_rot, cellmask, local_off, block_tot);
    ADY_LAUNCH_CHECK("label_cells_kernel");
    label_scan_kernel<<<1, 1024, 0, stream>>>(block_tot, nb, total_rows_dev);
    ADY_LAUNCH_CHECK("label_scan_kernel");
    return ADY_OK;
}

int launch_label_rows(const double* events, long long E, const CellCfg& cfg, const int8_t* rot, long long n_rot, const uint32_t* cellmask,
                      const void* ws, float* rows, long long max_rows, cudaStream_t stream) {
    if (E <= 0) return ADY_OK;
    const long long nb = (E + LB - 1) / LB;
    const uint32_t* local_off = reinterpret_cast<const uint32_t*>(ws);
    const long long* block_off = reinterpret_cast<const long long*>(reinterpret_cast<const char*>(ws) + ((E * 4 + 7) / 8) * 8);
    label_rows_kernel<false><<<(int)nb, LB, 0, stream>>>(events, E, cfg, rot, n_rot, cellmask, local_off, block_off, rows, max_rows, nullptr);
    ADY_LAUNCH_CHECK("label_rows_kernel");
    return ADY_OK;
}

// cells + rows back to back for a caller that has the row capacity up front (no host read of the count in between)
int launch_label_cells_rows(const double* events, long long E, int nb_label_frames, const CellCfg& cfg, const int8_t* rot, long long n_rot,
                            uint32_t* cellmask, long long* total_rows_dev, void* ws, float* rows, long long max_rows, cudaStream_t stream) {
    const long long nb = (E + LB - 1) / LB;
    if (E <= 0 || nb > 4096) {       // empty input, or so many blocks that the quadratic inline scan stops being free
        int rc = launch_label_cells(events, E, nb_label_frames, cfg, rot, n_rot, cellmask, total_rows_dev, ws, stream);
        if (rc) return rc;
        return launch_label_rows(events, E, cfg, rot, n_rot, cellmask, ws, rows, max_rows, stream);
    }
    if (cfg.ga * cfg.ge > 32 || cfg.ga > ADY_MAX_GRID || cfg.ge > ADY_MAX_GRID)
        return set_error(ADY_ERR_UNSUPPORTED, "label cells: grid %dx%d exceeds the 32-cell mask", cfg.ga, cfg.ge);
    uint32_t* local_off = reinterpret_cast<uint32_t*>(ws);
    long long* block_tot = reinterpret_cast<long long*>(reinterpret_cast<char*>(ws) + ((E * 4 + 7) / 8) * 8);
    label_cells_kernel<<<(int)nb, LB, 0, stream>>>(events, E, nb_label_frames, cfg, rot, n_rot, cellmask, local_off, block_tot);
    ADY_LAUNCH_CHECK("label_cells_kernel");
    label_rows_kernel<true><<<(int)nb, LB, 0, stream>>>(events, E, cfg, rot, n_rot, cellmask, local_off, block_tot, rows, max_rows, total_rows_dev);
    ADY_LAUNCH_CHECK("label_rows_kernel");
    return ADY_OK;
}

}  // namespace ady

// Constant tables of the front end (window, sparse mel matrix) and the library error string.
#include <math.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "frontend_core.cuh"
#include "frontend_host.h"
#include "frontend_tables.h"
#include "fe2_tables.h"

namespace ady {

static thread_local char g_err[512] = "";
char* last_error_buf() { return g_err; }
std::atomic<long long>& launch_counter() {
    static std::atomic<long long> n{0};
    return n;
}
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// ---- librosa 0.8.1 filters.mel (htk=False, norm='slaney'), reference call sites:
//      /root/reference/src/datasets.py:203, src/utils/utility.py:183,204
static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

void mel_filterbank_host(int sr, int n_fft, int n_mels, float* out) {
    const int nb = 1 + n_fft / 2;
    std::vector<double> fftf(nb), melf(n_mels + 2);
    for (int k = 0; k < nb; ++k) fftf[k] = (double)k * ((double)sr / 2) / (nb - 1);
    const double m0 = hz_to_mel(0.0), m1 = hz_to_mel((double)sr / 2);
    for (int i = 0; i < n_mels + 2; ++i) {
        // numpy.linspace: start + i*step, last element exactly stop
        double m = (i == n_mels + 1) ? m1 : m0 + i * ((m1 - m0) / (n_mels + 1));
        melf[i] = mel_to_hz(m);
    }
    for (int i = 0; i < n_mels; ++i) {
        const double fd0 = melf[i + 1] - melf[i], fd1 = melf[i + 2] - melf[i + 1];
        const double enorm = 2.0 / (melf[i + 2] - melf[i]);
        for (int k = 0; k < nb; ++k) {
            const double lower = -(melf[i] - fftf[k]) / fd0;
            const double upper = (melf[i + 2] - fftf[k]) / fd1;
            double w = lower < upper ? lower : upper;
            if (!(w > 0)) w = 0;
            const float w32 = (float)w;                       // weights[i] = ... (float32 array)
            out[(size_t)i * nb + k] = (float)((double)w32 * enorm);  // weights *= enorm[:, None]
        }
    }
}

static int build_tables(FrontendTables& t) {
    memset(&t, 0, sizeof(t));
    for (int n = 0; n < NFFT; ++n) t.hann[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / NFFT));
    std::vector<float> mel((size_t)NMEL * NBIN);
    mel_filterbank_host(24000, NFFT, NMEL, mel.data());
    int off = 0;
    for (int j = 0; j < NMEL; ++j) {
        int first = -1, last = -1;
        for (int k = 0; k < NBIN; ++k)
            if (mel[(size_t)j * NBIN + k] != 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        if (first < 0) return set_error(ADY_ERR_INVALID, "empty mel filter %d", j);
        const int len = last - first + 1;
        if (off + len > MEL_MAXNNZ) return set_error(ADY_ERR_INVALID, "mel matrix denser than expected");
        t.melidx[j] = (int16_t)first;
        t.melidx[NMEL + j] = (int16_t)len;
        t.melidx[2 * NMEL + j] = (int16_t)off;
        for (int i = 0; i < len; ++i) t.melw[off + i] = mel[(size_t)j * NBIN + first + i];
        off += len;
    }
    for (int n2 = 0; n2 < 25; ++n2) {
        t.wcs[2 * n2] = (float)cos(2.0 * M_PI * n2 / 25.0);
        t.wcs[2 * n2 + 1] = (float)sin(2.0 * M_PI * n2 / 25.0);
    }
    MelSchedule sch;
    if (!build_mel_schedule(mel.data(), sch)) return set_error(ADY_ERR_INVALID, "mel schedule does not fit (%d rows)", sch.total_rows);
    for (int i = 0; i < 4; ++i) { t.mel_hdr[i] = sch.it0[i]; t.mel_hdr[4 + i] = sch.nit[i]; }
    for (size_t i = 0; i < sch.ent.size(); ++i) { t.mel_pos[i] = sch.ent[i].pos; t.mel_w[i] = sch.ent[i].w; }
    return ADY_OK;
}

int get_frontend_tables(const FrontendTables** dev_tables) {
    static std::mutex mu;
    static FrontendTables* cache[64] = {nullptr};
    int dev = 0;
    ADY_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return set_error(ADY_ERR_INVALID, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(mu);
    if (!cache[dev]) {
        static FrontendTables host;
        int rc = build_tables(host);
        if (rc) return rc;
        FrontendTables* d = nullptr;
        ADY_CUDA_CHECK(cudaMalloc(&d, sizeof(FrontendTables)));
        ADY_CUDA_CHECK(cudaMemcpy(d, &host, sizeof(FrontendTables), cudaMemcpyHostToDevice));
        cache[dev] = d;
    }
    *dev_tables = cache[dev];
    return ADY_OK;
}

// fe2 kernel tables (window, stage-C twiddles, balanced mel schedule): plain host code in fe2_tables.h
int build_fe2_tables_host(fe2::Tables* t) {
    std::vector<float> mel((size_t)fe2::NMEL * fe2::NBIN);
    mel_filterbank_host(24000, fe2::NFFT, fe2::NMEL, mel.data());
    fe2::MelPlan plan;
    bool ok = false;
    fe2::fill_tables(mel.data(), *t, plan, ok);
    if (!ok) return set_error(ADY_ERR_INVALID, "fe2: the mel matrix does not fit %d lane-jobs of %d entries", fe2::NJOBS, fe2::MEL_L);
    return ADY_OK;
}

}  // namespace ady

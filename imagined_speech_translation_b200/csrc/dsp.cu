// Plan management and dispatch of the fused DSP chain (C ABI, include/eegx.h).
#include <math.h>
#include <stdlib.h>
#include <new>
#include <vector>

#include "dsp_plan.h"

extern "C" int eegx_dsp_plan_create(eegx_dsp_plan** out_plan, int C, int T, int n_fft, int hop,
                                    const float* fir, int numtaps, float log_eps, float z_eps) {
    EEGX_REQUIRE(out_plan && fir, EEGX_ERR_ARG, "plan/fir must not be NULL");
    *out_plan = nullptr;
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(C > 0 && T > 0 && hop > 0, EEGX_ERR_SHAPE, "bad sizes C=%d T=%d hop=%d", C, T, hop);
    EEGX_REQUIRE(n_fft >= 32 && n_fft <= 2048 && (n_fft & (n_fft - 1)) == 0, EEGX_ERR_SHAPE,
                 "n_fft=%d must be a power of two in [32, 2048]", n_fft);
    EEGX_REQUIRE(numtaps >= 1 && numtaps <= 129 && (numtaps & 1), EEGX_ERR_SHAPE,
                 "numtaps=%d must be odd and <= 129", numtaps);
    EEGX_REQUIRE(T > n_fft / 2, EEGX_ERR_SHAPE, "reflect padding needs T (%d) > n_fft/2 (%d)", T,
                 n_fft / 2);
    EEGX_REQUIRE(log_eps > 0.0f && z_eps >= 0.0f, EEGX_ERR_ARG, "log_eps must be > 0, z_eps >= 0");

    eegx_dsp_plan* p = new (std::nothrow) eegx_dsp_plan();
    EEGX_REQUIRE(p, EEGX_ERR_CUDA, "out of host memory");
    p->C = C; p->T = T; p->n_fft = n_fft; p->hop = hop; p->numtaps = numtaps;
    p->F = n_fft / 2 + 1;
    p->n_frames = 1 + T / hop;
    p->log_eps = log_eps; p->z_eps = z_eps;
    p->force_generic = 0;
    p->precise = 0;
    {
        const char* v = getenv("EEGX_DSP_VARIANT");
        p->tuned_variant = v ? atoi(v) : 0;
    }
    p->d_tables = nullptr;
    p->d_lane_tables = nullptr;
    for (int i = 0; i < 132; ++i) p->h_taps[i] = i < numtaps ? fir[i] : 0.0f;
    p->smem_generic = eegx::dsp_generic_smem_bytes(T, n_fft, hop, numtaps);
    if (p->smem_generic > 227 * 1024) {
        delete p;
        return eegx::set_error(EEGX_ERR_SHAPE,
                               "T=%d / n_fft=%d need %zu bytes of shared memory per CTA (> 227 KB)", T,
                               n_fft, p->smem_generic);
    }
    cudaError_t e = cudaGetDevice(&p->device);
    if (e != cudaSuccess) { delete p; return eegx::set_error(EEGX_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e)); }

    // constant tables, computed in double and rounded once
    p->off_window = 132;
    p->off_twiddle = p->off_window + n_fft;
    p->table_floats = p->off_twiddle + n_fft;  // n_fft/2 complex
    std::vector<float> host(p->table_floats, 0.0f);
    for (int i = 0; i < numtaps; ++i) host[i] = fir[i];
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < n_fft; ++n)  // torch.hann_window(n_fft, periodic=True)
        host[p->off_window + n] = (float)(0.5 - 0.5 * cos(two_pi * n / n_fft));
    for (int k = 0; k < n_fft / 2; ++k) {
        host[p->off_twiddle + 2 * k] = (float)cos(two_pi * k / n_fft);
        host[p->off_twiddle + 2 * k + 1] = (float)(-sin(two_pi * k / n_fft));
    }
    e = cudaMalloc(&p->d_tables, host.size() * sizeof(float));
    if (e != cudaSuccess) { delete p; return eegx::set_error(EEGX_ERR_CUDA, "cudaMalloc(tables): %s", cudaGetErrorString(e)); }
    e = cudaMemcpy(p->d_tables, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(p->d_tables);
        delete p;
        return eegx::set_error(EEGX_ERR_CUDA, "cudaMemcpy(tables): %s", cudaGetErrorString(e));
    }
    p->kernel = eegx::dsp_tuned_supported(p) ? 1 : (eegx::dsp_long_supported(p) ? 2 : 0);
    if (p->kernel != 0) {
        std::vector<float> lt(p->kernel == 1 ? eegx::dsp_tuned_table_floats() : eegx::dsp_long_table_floats());
        if (p->kernel == 1) eegx::dsp_tuned_fill_tables(lt.data());
        else eegx::dsp_long_fill_tables(lt.data());
        e = cudaMalloc(&p->d_lane_tables, lt.size() * sizeof(float));
        if (e == cudaSuccess)
            e = cudaMemcpy(p->d_lane_tables, lt.data(), lt.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            if (p->d_lane_tables) cudaFree(p->d_lane_tables);
            cudaFree(p->d_tables);
            delete p;
            return eegx::set_error(EEGX_ERR_CUDA, "tuned tables: %s", cudaGetErrorString(e));
        }
    }
    *out_plan = p;
    return EEGX_OK;
}

extern "C" int eegx_dsp_plan_destroy(eegx_dsp_plan* plan) {
    if (!plan) return EEGX_OK;
    if (plan->d_tables) cudaFree(plan->d_tables);
    if (plan->d_lane_tables) cudaFree(plan->d_lane_tables);
    delete plan;
    return EEGX_OK;
}

extern "C" int eegx_dsp_plan_dims(const eegx_dsp_plan* plan, int* F, int* N_f) {
    EEGX_REQUIRE(plan, EEGX_ERR_ARG, "plan is NULL");
    if (F) *F = plan->F;
    if (N_f) *N_f = plan->n_frames;
    return EEGX_OK;
}

extern "C" int eegx_dsp_plan_kernel(const eegx_dsp_plan* plan) {
    EEGX_REQUIRE(plan, EEGX_ERR_ARG, "plan is NULL");
    return plan->precise ? 3 : (plan->force_generic ? 0 : plan->kernel);
}

extern "C" int eegx_dsp_plan_force_generic(eegx_dsp_plan* plan, int on) {
    EEGX_REQUIRE(plan, EEGX_ERR_ARG, "plan is NULL");
    plan->force_generic = on ? 1 : 0;
    return EEGX_OK;
}

extern "C" int eegx_dsp_plan_set_precise(eegx_dsp_plan* plan, int on) {
    EEGX_REQUIRE(plan, EEGX_ERR_ARG, "plan is NULL");
    if (on) {
        const size_t smem = eegx::dsp_precise_smem_bytes(plan->T, plan->n_fft, plan->hop, plan->numtaps);
        EEGX_REQUIRE(smem <= 227 * 1024, EEGX_ERR_SHAPE, "float64 DSP kernel: T=%d / n_fft=%d need %zu bytes of "
                     "shared memory per CTA (> 227 KB)", plan->T, plan->n_fft, smem);
    }
    plan->precise = on ? 1 : 0;
    return EEGX_OK;
}

extern "C" int eegx_dsp_forward(const eegx_dsp_plan* plan, const float* x, const int64_t* onsets,
                                int64_t rec_len, float* out, int64_t B, void* stream) {
    EEGX_REQUIRE(plan, EEGX_ERR_ARG, "plan is NULL");
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(B >= 0, EEGX_ERR_SHAPE, "B=%lld", (long long)B);
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(x && out, EEGX_ERR_ARG, "x/out must not be NULL");
    EEGX_REQUIRE(onsets == nullptr || rec_len >= plan->T, EEGX_ERR_SHAPE,
                 "windowed mode needs rec_len (%lld) >= T (%d)", (long long)rec_len, plan->T);
    EEGX_REQUIRE(eegx::aligned16(x) && eegx::aligned16(out), EEGX_ERR_ALIGN,
                 "x and out must be 16-byte aligned");
    if (B == 0) return EEGX_OK;
    int dev = -1;
    EEGX_CUDA_CHECK(cudaGetDevice(&dev));
    EEGX_REQUIRE(dev == plan->device, EEGX_ERR_ARG, "plan was created on device %d, current is %d",
                 plan->device, dev);

    eegx::DspArgs a;
    a.x = x; a.onsets = onsets; a.rec_len = rec_len; a.out = out;
    a.rows = B * plan->C;
    a.C = plan->C; a.T = plan->T; a.n_fft = plan->n_fft; a.hop = plan->hop;
    a.numtaps = plan->numtaps; a.F = plan->F; a.n_frames = plan->n_frames;
    int l2 = 0;
    while ((1 << l2) < plan->n_fft / 2) ++l2;
    a.log2_m = l2;
    a.log_eps = plan->log_eps; a.z_eps = plan->z_eps;
    a.taps = plan->d_tables;
    a.window = plan->d_tables + plan->off_window;
    a.twiddle = reinterpret_cast<const float2*>(plan->d_tables + plan->off_twiddle);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (plan->precise) return eegx::launch_dsp_precise(plan, a, st);
    // windowed mode (arbitrary, possibly unaligned onsets) always takes the generic kernel
    if (plan->kernel == 1 && !plan->force_generic && onsets == nullptr)
        return eegx::launch_dsp_tuned(plan, a, st);
    if (plan->kernel == 2 && !plan->force_generic && onsets == nullptr)
        return eegx::launch_dsp_long(plan, a, st);
    return eegx::launch_dsp_generic(plan, a, st);
}

// Plan object of the fused DSP chain (opaque to C callers, see include/eegx.h).
#pragma once
#include "eegx_common.h"

struct eegx_dsp_plan {
    int C, T, n_fft, hop, numtaps;
    int F, n_frames;
    float log_eps, z_eps;
    int device;
    int kernel;          // 0 generic, 1 tuned (T 2048 / n_fft 256 / hop 64 / 65 taps), 2 long (T 4096 / n_fft 1024 / hop 256 / 65 taps)
    int force_generic;
    int precise;         // float64 arithmetic (dsp_precise.cu); overrides every other choice
    int tuned_variant;   // tile shape of the tuned kernel (EEGX_DSP_VARIANT, default 0 = auto)
    // device tables, one allocation: taps[numtaps] | pad to 4 | window[n_fft] | twiddle float2[n_fft/2]
    float* d_tables;
    float* d_lane_tables;  // tuned kernels only: per-lane window / twiddle constants
    int off_window, off_twiddle, table_floats;
    float h_taps[132];
    size_t smem_generic;
};

namespace eegx {

struct DspArgs {
    const float* x;
    const int64_t* onsets;
    int64_t rec_len;
    float* out;
    int64_t rows;  // B * C
    int C, T, n_fft, hop, numtaps, F, n_frames, log2_m;
    float log_eps, z_eps;
    const float* taps;     // device
    const float* window;   // device
    const float2* twiddle; // device, exp(-2*pi*i*k/n_fft), k < n_fft/2
};

int launch_dsp_generic(const eegx_dsp_plan* plan, const DspArgs& a, cudaStream_t st);
size_t dsp_generic_smem_bytes(int T, int n_fft, int hop, int numtaps);

// Tuned kernel for n_fft = 256, hop = 64, 65 taps (BASELINE config 2).
bool dsp_tuned_supported(const eegx_dsp_plan* plan);
int launch_dsp_tuned(const eegx_dsp_plan* plan, const DspArgs& a, cudaStream_t st);
int launch_dsp_pair(const eegx_dsp_plan* plan, const DspArgs& a, cudaStream_t st);
int launch_dsp_umma(const eegx_dsp_plan* plan, const DspArgs& a, cudaStream_t st);   // FIR on tcgen05 (dsp_umma.cu)
int dsp_tuned_table_floats();
void dsp_tuned_fill_tables(float* host);

// Float64 variant of the generic kernel (eegx_dsp_plan_set_precise).
int launch_dsp_precise(const eegx_dsp_plan* plan, const DspArgs& a, cudaStream_t st);
size_t dsp_precise_smem_bytes(int T, int n_fft, int hop, int numtaps);

// Tuned kernel for T = 4096, n_fft = 1024, hop = 256, 65 taps (BASELINE config 4).
bool dsp_long_supported(const eegx_dsp_plan* plan);
int launch_dsp_long(const eegx_dsp_plan* plan, const DspArgs& a, cudaStream_t st);
int dsp_long_table_floats();
void dsp_long_fill_tables(float* host);

}  // namespace eegx

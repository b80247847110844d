// Tuned DSP kernel for BASELINE config 2 (n_fft 256 / hop 64 / 65 taps).
#include "dsp_plan.h"

namespace eegx {

bool dsp_tuned_supported(const eegx_dsp_plan*) { return false; }

int launch_dsp_tuned(const eegx_dsp_plan*, const DspArgs&, cudaStream_t) {
    return set_error(EEGX_ERR_SHAPE, "tuned DSP kernel not available for this plan");
}

}  // namespace eegx

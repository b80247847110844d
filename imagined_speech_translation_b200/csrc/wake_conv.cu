// wake_model's convolution / max-pool front (wake_model/train.cpp:26-33) in fp64 on the GPU.
//
//   Convolution::forward   wake_model/layers/convolution.cpp:36-57   -> conv_fwd_kernel
//   Convolution::backward  convolution.cpp:60-112                    -> conv_dx_kernel, conv_update_kernel
//   MaxPool::forward       wake_model/layers/maxpool.cpp:6-43        -> pool_fwd_kernel
//   MaxPool::backward      maxpool.cpp:46-69                         -> pool_bwd_kernel
//
// Integer / index results (the argmax pairs) and every fp64 result are BIT-EXACT against the reference: each output
// element is owned by one thread that adds its terms in the reference's loop order with separately rounded
// multiplies and adds (__dmul_rn / __dadd_rn: no FMA contraction, as the reference is compiled).  The reference's
// quirks are kept as written: the constructor's activation is never applied, the "input gradient" uses the flipped
// kernel at input position (y + ky, x + kx), and the pool's row bound is the member input_height that maxpool.h:15
// sets to input_width.  The grids are tiny (2 x a few thousand samples, kernels of 32 .. 128 taps): this path is
// bound by launch latency, not by the machine; it exists so that the wake-word network runs end to end on the device.
#include <math.h>

#include "eegx_common.h"

namespace {

struct ConvArgs {
    const double* x;
    double* kernel;
    double* bias;
    const double* dout;
    double* y;
    double* dx;
    int H, W, kh, kw, OH, OW;
    double lr;
};

__global__ void conv_fwd_kernel(const ConvArgs a) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)a.OH * a.OW) return;
    const int oy = (int)(idx / a.OW), ox = (int)(idx % a.OW);
    double sum = 0.0;
    for (int ky = 0; ky < a.kh; ++ky)
        for (int kx = 0; kx < a.kw; ++kx)
            sum = __dadd_rn(sum, __dmul_rn(a.x[(long long)(oy + ky) * a.W + ox + kx], a.kernel[ky * a.kw + kx]));
    a.y[idx] = __dadd_rn(sum, a.bias[0]);
}

// dx[iy][ix] = sum over the output positions (y, x) that touch it, in the reference's order (y ascending, then x
// ascending; for a fixed output position ky / kx are determined), of kernel[kh-1-ky][kw-1-kx] * dout[y][x]
__global__ void conv_dx_kernel(const ConvArgs a) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)a.H * a.W) return;
    const int iy = (int)(idx / a.W), ix = (int)(idx % a.W);
    const int y0 = max(0, iy - a.kh + 1), y1 = min(a.OH - 1, iy);
    const int x0 = max(0, ix - a.kw + 1), x1 = min(a.OW - 1, ix);
    double sum = 0.0;
    for (int y = y0; y <= y1; ++y) {
        const int ky = iy - y;
        for (int x = x0; x <= x1; ++x) {
            const int kx = ix - x;
            sum = __dadd_rn(sum, __dmul_rn(a.kernel[(a.kh - ky - 1) * a.kw + (a.kw - kx - 1)], a.dout[(long long)y * a.OW + x]));
        }
    }
    a.dx[idx] = sum;
}

// thread t < kh*kw: kernel gradient of tap t (sequential over the output positions), then the SGD update;
// thread kh*kw: the bias gradient and its update.  Runs after conv_dx_kernel (which needs the old kernel).
__global__ void conv_update_kernel(const ConvArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int taps = a.kh * a.kw;
    if (t < taps) {
        const int ky = t / a.kw, kx = t % a.kw;
        double g = 0.0;
        for (int y = 0; y < a.OH; ++y)
            for (int x = 0; x < a.OW; ++x)
                g = __dadd_rn(g, __dmul_rn(a.x[(long long)(y + ky) * a.W + x + kx], a.dout[(long long)y * a.OW + x]));
        a.kernel[t] = __dsub_rn(a.kernel[t], __dmul_rn(a.lr, g));
    } else if (t == taps) {
        double g = 0.0;
        for (long long i = 0; i < (long long)a.OH * a.OW; ++i) g = __dadd_rn(g, a.dout[i]);
        a.bias[0] = __dsub_rn(a.bias[0], __dmul_rn(a.lr, g));
    }
}

struct PoolArgs {
    const double* x;
    const double* dout;
    double* y;
    int32_t* argmax;
    double* dx;
    int H, W, pw, ph, stride, OH, OW;
};

__global__ void pool_fwd_kernel(const PoolArgs a) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)a.OH * a.OW) return;
    const int i = (int)(idx / a.OW), j = (int)(idx % a.OW);
    const int bound_h = a.W;                       // maxpool.h:15: this->input_height = input_width
    double best = -INFINITY;
    int bi = -1, bj = -1;
    for (int m = 0; m < a.ph; ++m)
        for (int n = 0; n < a.pw; ++n) {
            const int ii = i * a.stride + m, jj = j * a.stride + n;
            if (ii < bound_h && jj < a.W) {
                const double v = a.x[(long long)ii * a.W + jj];
                if (v > best) { best = v; bi = ii; bj = jj; }
            }
        }
    if (a.y) a.y[idx] = best;
    a.argmax[2 * idx] = bi;
    a.argmax[2 * idx + 1] = bj;
}

// dx[iy][ix] = sum, in (i, j) order, of dout[i][j] over the windows whose recorded maximum is (iy, ix)
__global__ void pool_bwd_kernel(const PoolArgs a) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)a.H * a.W) return;
    const int iy = (int)(idx / a.W), ix = (int)(idx % a.W);
    const int i0 = max(0, (iy - a.ph + a.stride) / a.stride), i1 = min(a.OH - 1, iy / a.stride);
    const int j0 = max(0, (ix - a.pw + a.stride) / a.stride), j1 = min(a.OW - 1, ix / a.stride);
    double sum = 0.0;
    for (int i = i0; i <= i1; ++i)
        for (int j = j0; j <= j1; ++j) {
            const long long o = (long long)i * a.OW + j;
            if (a.argmax[2 * o] == iy && a.argmax[2 * o + 1] == ix) sum = __dadd_rn(sum, a.dout[o]);
        }
    a.dx[idx] = sum;
}

inline unsigned blocks_for(long long n, int nt) { return (unsigned)((n + nt - 1) / nt); }

}  // namespace

extern "C" int eegx_wake_conv2d_f64(double* kernel, double* bias, const double* x, int64_t H, int64_t W, int64_t kh,
                                    int64_t kw, const double* dout, double lr, double* y, double* dx, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(kernel && bias && x, EEGX_ERR_ARG, "kernel / bias / x must not be NULL");
    EEGX_REQUIRE(H > 0 && W > 0 && kh > 0 && kw > 0 && kh <= H && kw <= W && H * W < (1ll << 31), EEGX_ERR_SHAPE,
                 "bad shape: input %lld x %lld, kernel %lld x %lld", (long long)H, (long long)W, (long long)kh, (long long)kw);
    EEGX_REQUIRE(y || dout, EEGX_ERR_ARG, "nothing to do: y and dout are both NULL");
    EEGX_REQUIRE(!dout || dx, EEGX_ERR_ARG, "backward needs dx");
    ConvArgs a;
    a.x = x; a.kernel = kernel; a.bias = bias; a.dout = dout; a.y = y; a.dx = dx;
    a.H = (int)H; a.W = (int)W; a.kh = (int)kh; a.kw = (int)kw; a.OH = (int)(H - kh + 1); a.OW = (int)(W - kw + 1);
    a.lr = lr;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (y) conv_fwd_kernel<<<blocks_for((long long)a.OH * a.OW, 128), 128, 0, st>>>(a);
    if (dout) {
        conv_dx_kernel<<<blocks_for((long long)a.H * a.W, 128), 128, 0, st>>>(a);
        conv_update_kernel<<<blocks_for(a.kh * a.kw + 1, 64), 64, 0, st>>>(a);
    }
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

extern "C" int eegx_wake_maxpool_f64(const double* x, int64_t H, int64_t W, int64_t pool_w, int64_t pool_h, int64_t stride,
                                     const double* dout, double* y, int32_t* argmax, double* dx, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(x && argmax, EEGX_ERR_ARG, "x / argmax must not be NULL");
    EEGX_REQUIRE(H > 0 && W > 0 && pool_w > 0 && pool_h > 0 && stride > 0 && pool_h <= H && pool_w <= W &&
                 H * W < (1ll << 31), EEGX_ERR_SHAPE, "bad shape: input %lld x %lld, pool %lld x %lld, stride %lld",
                 (long long)H, (long long)W, (long long)pool_h, (long long)pool_w, (long long)stride);
    EEGX_REQUIRE(!dout || dx, EEGX_ERR_ARG, "backward needs dx");
    PoolArgs a;
    a.x = x; a.dout = dout; a.y = y; a.argmax = argmax; a.dx = dx;
    a.H = (int)H; a.W = (int)W; a.pw = (int)pool_w; a.ph = (int)pool_h; a.stride = (int)stride;
    a.OH = (int)((H - pool_h) / stride + 1); a.OW = (int)((W - pool_w) / stride + 1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    pool_fwd_kernel<<<blocks_for((long long)a.OH * a.OW, 128), 128, 0, st>>>(a);
    if (dout) pool_bwd_kernel<<<blocks_for((long long)a.H * a.W, 128), 128, 0, st>>>(a);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

/* TEST INFRASTRUCTURE -- CPU restatement of wake_model's convolution / max-pool layers (the front of the wake-word
 * network, wake_model/train.cpp:26-33), in plain C with the reference's statement and rounding order, so that the
 * CUDA kernels (csrc/wake_conv.cu) can be checked bit for bit.  Pinned against the reference's own classes compiled
 * from /root/reference (oracle/_ref/libwake_ref.so, oracle/wake_ref_harness.cpp) by tests/test_wake.py.
 * Only tests/, __graft_entry__.smoke() and CPU-baseline legs may call this.  Built with -ffp-contract=off.
 *
 *   Convolution::forward   wake_model/layers/convolution.cpp:36-57   valid cross-correlation + one bias; the
 *                          constructor's "activation" string is stored and never applied
 *   Convolution::backward  convolution.cpp:60-112   kernel gradient, the layer's "input gradient" (flipped kernel at
 *                          input position (y + ky, x + kx) -- restated as written), plain SGD on kernel and bias
 *   MaxPool::forward       maxpool.cpp:6-43    first strict maximum of the window; the row bound is the member
 *                          input_height, which maxpool.h:15 sets to input_WIDTH
 *   MaxPool::backward      maxpool.cpp:46-69   scatter-add of the output gradient to the recorded maxima
 */
#include <math.h>
#include <stddef.h>

/* y = conv(x); when dout != NULL also dx = Convolution::backward(dout, lr) and the SGD update of kernel / bias.
 * x (H, W), kernel (kh, kw), y and dout (H - kh + 1, W - kw + 1), dx (H, W); all row-major doubles. */
int wake_conv2d_oracle(double* kernel, double* bias, const double* x, int H, int W, int kh, int kw, const double* dout,
                       double lr, double* y, double* dx) {
    const int OH = H - kh + 1, OW = W - kw + 1;
    if (OH <= 0 || OW <= 0) return -1;
    if (y) {
        for (int oy = 0; oy < OH; ++oy)
            for (int ox = 0; ox < OW; ++ox) {
                double sum = 0.0;
                for (int ky = 0; ky < kh; ++ky)
                    for (int kx = 0; kx < kw; ++kx) sum += x[(size_t)(oy + ky) * W + ox + kx] * kernel[ky * kw + kx];
                y[(size_t)oy * OW + ox] = sum + bias[0];
            }
    }
    if (!dout) return 0;
    /* convolution.cpp:61-97: one pass over the output positions feeds both gradients, in this order */
    double kgrad[64 * 64];
    if (kh * kw > 64 * 64) return -2;
    for (int i = 0; i < kh * kw; ++i) kgrad[i] = 0.0;
    if (dx)
        for (size_t i = 0; i < (size_t)H * W; ++i) dx[i] = 0.0;
    for (int oy = 0; oy < OH; ++oy)
        for (int ox = 0; ox < OW; ++ox) {
            const double d = dout[(size_t)oy * OW + ox];
            for (int ky = 0; ky < kh; ++ky)
                for (int kx = 0; kx < kw; ++kx) kgrad[ky * kw + kx] += x[(size_t)(oy + ky) * W + ox + kx] * d;
            if (dx)
                for (int ky = 0; ky < kh; ++ky)
                    for (int kx = 0; kx < kw; ++kx) {
                        const int iy = oy + ky, ix = ox + kx;
                        if (iy < H && ix < W) dx[(size_t)iy * W + ix] += kernel[(kh - ky - 1) * kw + (kw - kx - 1)] * d;
                    }
        }
    for (int i = 0; i < kh * kw; ++i) kernel[i] -= lr * kgrad[i];
    double bgrad = 0.0;
    for (int oy = 0; oy < OH; ++oy)
        for (int ox = 0; ox < OW; ++ox) bgrad += dout[(size_t)oy * OW + ox];
    bias[0] -= lr * bgrad;
    return 0;
}

/* y = maxpool(x) with the argmax pairs (row, col; -1 when the window saw no element); when dout != NULL also
 * dx = MaxPool::backward(dout).  x (H, W); y, dout (OH, OW) with OH = (H - ph) / stride + 1, OW = (W - pw) / stride + 1. */
int wake_maxpool_oracle(const double* x, int H, int W, int pw, int ph, int stride, const double* dout, double* y,
                        int* argmax, double* dx) {
    const int OH = (H - ph) / stride + 1, OW = (W - pw) / stride + 1;
    if (OH <= 0 || OW <= 0 || stride <= 0) return -1;
    const int bound_h = W;   /* maxpool.h:15: this->input_height = input_width */
    for (int i = 0; i < OH; ++i)
        for (int j = 0; j < OW; ++j) {
            double best = -INFINITY;
            int bi = -1, bj = -1;
            for (int m = 0; m < ph; ++m)
                for (int n = 0; n < pw; ++n) {
                    const int ii = i * stride + m, jj = j * stride + n;
                    if (ii < bound_h && jj < W) {
                        const double v = x[(size_t)ii * W + jj];
                        if (v > best) { best = v; bi = ii; bj = jj; }
                    }
                }
            if (y) y[(size_t)i * OW + j] = best;
            argmax[2 * ((size_t)i * OW + j)] = bi;
            argmax[2 * ((size_t)i * OW + j) + 1] = bj;
        }
    if (!dout || !dx) return 0;
    for (size_t i = 0; i < (size_t)H * W; ++i) dx[i] = 0.0;
    for (int i = 0; i < OH; ++i)
        for (int j = 0; j < OW; ++j) {
            const int bi = argmax[2 * ((size_t)i * OW + j)], bj = argmax[2 * ((size_t)i * OW + j) + 1];
            if (bi >= 0 && bj >= 0) dx[(size_t)bi * W + bj] += dout[(size_t)i * OW + j];
        }
    return 0;
}

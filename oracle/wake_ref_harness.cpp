// TEST INFRASTRUCTURE -- drives the REFERENCE's own wake_model code (compiled from /root/reference where it lies,
// never copied) through the same C signature as oracle/wake_dense_oracle.c so the restatement can be pinned bit-exactly.
//
// Links wake_model/layers/linear.cpp, convolution.cpp and maxpool.cpp and includes their headers, layers/activations.h
// and layers/losses.h; the loop
// below is the Linear-only part of wake_model/train.cpp:68-117 (forward through both layers, loss, delta, backward).
// The full program cannot serve as an oracle: it needs a dataset that is not shipped and reads out of bounds
// (SURVEY.md section 2), so only these leaf classes are exercised.
#include <string>
#include <vector>

#include "layers/convolution.h"
#include "layers/linear.h"
#include "layers/losses.h"
#include "layers/maxpool.h"

static const char* act_name(int act) {
    switch (act) {
        case 1: return "relu";
        case 2: return "sigmoid";
        case 3: return "tanh";
        default: return "";
    }
}

extern "C" int wake_dense_ref(double* w1, double* b1, double* w2, double* b2, const double* x, const int* label, long n,
                              int in, int hidden, int ncls, double lr, int act, int train, double* loss, double* probs,
                              double* dx) {
    Linear l1(in, hidden, act_name(act));
    Linear l2(hidden, ncls, "softmax", true);
    for (int i = 0; i < hidden; ++i) {
        for (int j = 0; j < in; ++j) l1.weights[i][j] = w1[(long)i * in + j];
        l1.biases[i] = b1[i];
    }
    for (int k = 0; k < ncls; ++k) {
        for (int j = 0; j < hidden; ++j) l2.weights[k][j] = w2[(long)k * hidden + j];
        l2.biases[k] = b2[k];
    }
    std::vector<Neuron> inp(in);
    for (long s = 0; s < n; ++s) {
        for (int j = 0; j < in; ++j) inp[j].output = x[s * (long)in + j];
        std::vector<Neuron> hid = l1.forward(inp);
        std::vector<Neuron> out = l2.forward(hid);
        std::vector<double> res(ncls), truth(ncls, 0.0);
        for (int k = 0; k < ncls; ++k) res[k] = out[k].output;
        if (label[s] >= 0 && label[s] < ncls) truth[label[s]] = 1.0;
        if (probs) for (int k = 0; k < ncls; ++k) probs[s * (long)ncls + k] = res[k];
        if (loss) loss[s] = categorical_cross_entropy_loss(res, truth);
        if (!train) continue;
        std::vector<double> delta = derivative_categorical_cross_entropy(res, truth);
        delta = l2.backward(delta, lr);
        delta = l1.backward(delta, lr);
        if (dx) for (int j = 0; j < in; ++j) dx[s * (long)in + j] = delta[j];
    }
    for (int i = 0; i < hidden; ++i) {
        for (int j = 0; j < in; ++j) w1[(long)i * in + j] = l1.weights[i][j];
        b1[i] = l1.biases[i];
    }
    for (int k = 0; k < ncls; ++k) {
        for (int j = 0; j < hidden; ++j) w2[(long)k * hidden + j] = l2.weights[k][j];
        b2[k] = l2.biases[k];
    }
    return 0;
}


// ---- the convolution / max-pool front (wake_model/train.cpp:26-33): the reference's own classes, same C signatures as
// oracle/wake_conv_oracle.c ----
static std::vector<std::vector<Neuron>> to_grid(const double* x, int H, int W) {
    std::vector<std::vector<Neuron>> g(H, std::vector<Neuron>(W));
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) g[i][j].output = x[(long)i * W + j];
    return g;
}

extern "C" int wake_conv2d_ref(double* kernel, double* bias, const double* x, int H, int W, int kh, int kw,
                               const double* dout, double lr, double* y, double* dx) {
    Convolution conv(W, H, kw, kh, "relu");
    for (int a = 0; a < kh; ++a)
        for (int b = 0; b < kw; ++b) conv.kernel[a][b] = kernel[a * kw + b];
    conv.biases[0] = bias[0];
    const int OH = conv.output_height, OW = conv.output_width;
    std::vector<std::vector<Neuron>> out = conv.forward(to_grid(x, H, W));
    if (y)
        for (int i = 0; i < OH; ++i)
            for (int j = 0; j < OW; ++j) y[(long)i * OW + j] = out[i][j].output;
    if (!dout) return 0;
    std::vector<std::vector<Neuron>> gin = conv.backward(to_grid(dout, OH, OW), lr);
    if (dx)
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) dx[(long)i * W + j] = gin[i][j].output;
    for (int a = 0; a < kh; ++a)
        for (int b = 0; b < kw; ++b) kernel[a * kw + b] = conv.kernel[a][b];
    bias[0] = conv.biases[0];
    return 0;
}

extern "C" int wake_maxpool_ref(const double* x, int H, int W, int pw, int ph, int stride, const double* dout, double* y,
                                int* argmax, double* dx) {
    MaxPool pool(W, H, pw, ph, stride);
    const int OH = pool.output_height, OW = pool.output_width;
    std::vector<std::vector<Neuron>> out = pool.forward(to_grid(x, H, W));
    for (int i = 0; i < OH; ++i)
        for (int j = 0; j < OW; ++j) {
            if (y) y[(long)i * OW + j] = out[i][j].output;
            argmax[2 * ((long)i * OW + j)] = pool.max_indices[i][j].first;
            argmax[2 * ((long)i * OW + j) + 1] = pool.max_indices[i][j].second;
        }
    if (!dout || !dx) return 0;
    std::vector<std::vector<Neuron>> gin = pool.backward(to_grid(dout, OH, OW));
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) dx[(long)i * W + j] = gin[i][j].output;
    return 0;
}

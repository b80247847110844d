// TEST INFRASTRUCTURE -- drives the REFERENCE's own wake_model code (compiled from /root/reference where it lies,
// never copied) through the same C signature as oracle/wake_dense_oracle.c so the restatement can be pinned bit-exactly.
//
// Links wake_model/layers/linear.cpp and includes layers/linear.h, layers/activations.h, layers/losses.h; the loop
// below is the Linear-only part of wake_model/train.cpp:68-117 (forward through both layers, loss, delta, backward).
// The full program cannot serve as an oracle: it needs a dataset that is not shipped and reads out of bounds
// (SURVEY.md section 2), so only these leaf classes are exercised.
#include <string>
#include <vector>

#include "layers/linear.h"
#include "layers/losses.h"

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

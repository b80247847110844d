/* TEST INFRASTRUCTURE -- CPU restatement of wake_model's dense head, per-sample SGD, fp64.
 *
 * Only tests/, __graft_entry__.smoke() and bench/tools CPU-baseline legs may load this; the product path
 * (imagined_speech_translation_b200/wake.py -> eegx_wake_dense_f64) never does.
 *
 * Follows, statement by statement and in the same floating-point operation order:
 *   wake_model/layers/linear.cpp:5-44     Linear::forward   (out = 0; out += in_j * w_ij for j = 0..; out += bias; activation)
 *   wake_model/layers/linear.cpp:47-72    Linear::backward  (dout_i *= act'(output_i); dinput_j += w_ij * dout_i BEFORE
 *                                         w_ij -= lr * (in_j * dout_i); b_i -= lr * dout_i)
 *   wake_model/layers/activations.h:9-41  sigmoid / tanh written with std::pow(e, x), softmax with max subtraction
 *   wake_model/layers/activations.h:64-95 apply_activation / derivative_of_activation (derivative taken at the OUTPUT)
 *   wake_model/layers/losses.h:8-22       loss = -sum labels_i * log(p_i + 1e-15), delta = p - labels
 *   wake_model/train.cpp:98-117           forward through both layers, loss, delta, backward output layer then hidden layer
 *
 * Pinned: bit-exact against oracle/_ref/libwake_ref.so (the reference's own linear.cpp / activations.h / losses.h compiled
 * here, see oracle/Makefile) and against tests/golden/wake_dense_ref.npz generated from it.
 * Compile WITHOUT -march=native / -ffast-math so no FMA contraction changes the rounding.
 */
#include <math.h>
#include <stdlib.h>

static const double e_const = 2.718281828459045235360287471352;

static double act_apply(double v, int act) {
    switch (act) {
        case 1: return v > 0.0 ? v : 0.0;                                         /* std::max(0.0, x) */
        case 2: return 1 / (1 + pow(e_const, -v));
        case 3: return (pow(e_const, v) - pow(e_const, -v)) / (pow(e_const, v) + pow(e_const, -v));
        default: return v;
    }
}

static double act_derivative(double out, int act) {
    switch (act) {
        case 1: return out > 0.0 ? 1.0 : 0.0;
        case 2: { double s = act_apply(out, 2); return s * (1 - act_apply(out, 2)); }
        case 3: { double t = act_apply(out, 3); return 1 - t * t; }
        default: return 1.0;
    }
}

/* activation: 0 none, 1 relu, 2 sigmoid, 3 tanh (hidden layer); the output layer is softmax + CCE.
 * x (n, in) row-major; w1 (hidden, in); w2 (ncls, hidden); loss (n) / probs (n, ncls) / dx (n, in) may be NULL. */
int wake_dense_oracle(double* w1, double* b1, double* w2, double* b2, const double* x, const int* label, long n, int in,
                      int hidden, int ncls, double lr, int act, int train, double* loss, double* probs, double* dx) {
    double* h = (double*)malloc(sizeof(double) * hidden);
    double* p = (double*)malloc(sizeof(double) * ncls);
    double* d2 = (double*)malloc(sizeof(double) * ncls);
    double* dh = (double*)malloc(sizeof(double) * hidden);
    double* dxs = (double*)malloc(sizeof(double) * in);
    if (!h || !p || !d2 || !dh || !dxs) return -1;
    for (long s = 0; s < n; ++s) {
        const double* xs = x + s * (long)in;
        for (int i = 0; i < hidden; ++i) {
            double out = 0.0;
            for (int j = 0; j < in; ++j) out += xs[j] * w1[(long)i * in + j];
            out += b1[i];
            h[i] = act ? act_apply(out, act) : out;
        }
        for (int k = 0; k < ncls; ++k) {
            double out = 0.0;
            for (int j = 0; j < hidden; ++j) out += h[j] * w2[(long)k * hidden + j];
            out += b2[k];
            p[k] = out;
        }
        double mx = p[0];
        for (int k = 1; k < ncls; ++k) if (p[k] > mx) mx = p[k];
        double sum = 0.0;
        for (int k = 0; k < ncls; ++k) { p[k] = exp(p[k] - mx); sum += p[k]; }
        for (int k = 0; k < ncls; ++k) p[k] /= sum;
        if (probs) for (int k = 0; k < ncls; ++k) probs[s * (long)ncls + k] = p[k];
        const int y = label[s];
        if (loss) {
            double l = 0.0;
            for (int k = 0; k < ncls; ++k) l -= (k == y ? 1.0 : 0.0) * log(p[k] + 1e-15);
            loss[s] = l;
        }
        if (!train) continue;
        for (int k = 0; k < ncls; ++k) d2[k] = p[k] - (k == y ? 1.0 : 0.0);
        /* output layer backward (softmax: no derivative factor) */
        for (int j = 0; j < hidden; ++j) dh[j] = 0.0;
        for (int k = 0; k < ncls; ++k) {
            for (int j = 0; j < hidden; ++j) {
                dh[j] += w2[(long)k * hidden + j] * d2[k];
                w2[(long)k * hidden + j] -= lr * (h[j] * d2[k]);
            }
            b2[k] -= lr * d2[k];
        }
        /* hidden layer backward */
        if (act) for (int i = 0; i < hidden; ++i) dh[i] *= act_derivative(h[i], act);
        for (int j = 0; j < in; ++j) dxs[j] = 0.0;
        for (int i = 0; i < hidden; ++i) {
            for (int j = 0; j < in; ++j) {
                dxs[j] += w1[(long)i * in + j] * dh[i];
                w1[(long)i * in + j] -= lr * (xs[j] * dh[i]);
            }
            b1[i] -= lr * dh[i];
        }
        if (dx) for (int j = 0; j < in; ++j) dx[s * (long)in + j] = dxs[j];
    }
    free(h); free(p); free(d2); free(dh); free(dxs);
    return 0;
}

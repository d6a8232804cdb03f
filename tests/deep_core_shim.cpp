// C shim around openrl_b200/csrc/orl_deep_core.h for the CPU test (g++ -O2 -shared -fPIC).
#include "orl_deep_core.h"
using namespace orl_deep;

extern "C" {
int shim_deep_param_count(int d, int n) { return deep_offsets(d, n).total; }
int shim_deep_tape_width() { return TAPE; }

// rows independent forwards (value, logits) and, when dvalue / dlogits are given, the backward tape rows
void shim_deep_rows(const float* P, int d, int n, int act, int rows, const float* X, float* values, float* logits,
                    const float* dvalue, const float* dlogits, float* tape) {
    const Offsets o = deep_offsets(d, n);
    for (int r = 0; r < rows; ++r) {
        Save sv;
        float* tp = tape ? tape + (size_t)r * TAPE : nullptr;
        deep_forward(P, o, act, X + r * d, values + r, logits + r * n, tape ? &sv : nullptr, tp);
        if (tape) deep_backward(P, o, act, sv, dvalue[r], dlogits + r * n, tp);
    }
}
}

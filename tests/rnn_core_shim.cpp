// C shim around openrl_b200/csrc/orl_rnn_core.h for the CPU test (g++ -O2 -shared -fPIC).
#include "orl_rnn_core.h"
#include <vector>
using namespace orl_rnn;

extern "C" {
int shim_param_count(int d, int n) { return rnn_offsets(d, n).total; }
int shim_tape_width() { return TAPE; }

// rows independent single steps
void shim_forward_rows(const float* P, int d, int n, int act, int rows, const float* X, const float* Hin, const float* mask,
                       float* Hout, float* Out) {
    const Offsets o = rnn_offsets(d, n);
    for (int r = 0; r < rows; ++r) rnn_step_forward(P, o, act, X + r * d, Hin + r * H, mask[r], Hout + r * H, Out + r * n, nullptr, nullptr);
}

// chunks: time-major rows (row = l*nchunk + c); forward L steps with saves, backward with the dh chain
void shim_chunk_fwdbwd(const float* P, int d, int n, int act, int L, int nchunk, const float* X, const float* H0, const float* masks,
                       const float* dlogits, float* Out, float* tape) {
    const Offsets o = rnn_offsets(d, n);
    std::vector<StepSave> sv(L);
    for (int c = 0; c < nchunk; ++c) {
        float h[H], h2[H];
        for (int j = 0; j < H; ++j) h[j] = H0[c * H + j];
        for (int l = 0; l < L; ++l) {
            const int row = l * nchunk + c;
            rnn_step_forward(P, o, act, X + row * d, h, masks[row], h2, Out + row * n, &sv[l], tape + (size_t)row * TAPE);
            for (int j = 0; j < H; ++j) h[j] = h2[j];
        }
        float dh[H], dhp[H];
        for (int j = 0; j < H; ++j) dh[j] = 0.f;
        for (int l = L - 1; l >= 0; --l) {
            const int row = l * nchunk + c;
            rnn_step_backward(P, o, act, sv[l], masks[row], dlogits + row * n, dh, dhp, tape + (size_t)row * TAPE);
            for (int j = 0; j < H; ++j) dh[j] = dhp[j];
        }
    }
}
}

// synth.cpp — see synth.h for the specification.
#include "synth.h"
#include <random>
#include <thread>
#include <vector>
#include <cstring>
#include <algorithm>

static uint32_t gen_window(const cg_synth_spec* s, uint32_t w, char* out, uint32_t* lens) {
    static const char B[4] = {'A', 'C', 'G', 'T'};
    std::mt19937_64 rng(s->seed * 1000003ULL + w);
    std::vector<uint8_t> truth(s->truth_len);
    for (uint32_t i = 0; i < s->truth_len; ++i) truth[i] = (uint8_t)(rng() & 3);
    uint32_t total = 0;
    for (uint32_t q = 0; q < s->n_seqs; ++q) {
        uint32_t n = 0;
        for (uint32_t i = 0; i < s->truth_len; ++i) {
            double u = (double)(rng() >> 11) * (1.0 / 9007199254740992.0);
            if (u >= s->err) {
                out[total + n++] = B[truth[i]];
            } else {
                double v = u / s->err;
                if (v < s->p_sub) {
                    out[total + n++] = B[(truth[i] + 1 + (uint32_t)(rng() % 3)) & 3];
                } else if (v < s->p_sub + s->p_ins) {
                    out[total + n++] = B[rng() & 3];
                    out[total + n++] = B[truth[i]];
                }
            }
        }
        lens[q] = n;
        total += n;
    }
    return total;
}

extern "C" uint64_t cg_synth_max_bases(const cg_synth_spec* s) {
    return (uint64_t)s->n_windows * s->n_seqs * (2ULL * s->truth_len);
}

extern "C" uint64_t cg_synth_windows(const cg_synth_spec* s, uint32_t* win_seq_begin, uint64_t* seq_off,
                                     char* bases, int threads) {
    const uint32_t W = s->n_windows, N = s->n_seqs;
    const uint64_t stride = (uint64_t)N * 2 * s->truth_len;
    // Pass 1: every window into a private slot (parallel), pass 2: compact (serial memmove).
    std::vector<uint32_t> lens((size_t)W * N);
    std::vector<uint64_t> wbytes(W);
    if (threads < 1) threads = 1;
    auto work = [&](int t) {
        for (uint32_t w = t; w < W; w += threads)
            wbytes[w] = gen_window(s, s->first_window + w, bases + (uint64_t)w * stride, &lens[(size_t)w * N]);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    uint64_t off = 0;
    for (uint32_t w = 0; w < W; ++w) {
        win_seq_begin[w] = w * N;
        if (off != (uint64_t)w * stride) memmove(bases + off, bases + (uint64_t)w * stride, wbytes[w]);
        uint64_t o = off;
        for (uint32_t q = 0; q < N; ++q) { seq_off[(size_t)w * N + q] = o; o += lens[(size_t)w * N + q]; }
        off += wbytes[w];
    }
    win_seq_begin[W] = W * N;
    seq_off[(size_t)W * N] = off;
    return off;
}

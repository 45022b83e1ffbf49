// synth.cpp — see synth.h for the specification.
#include "synth.h"
#include <random>
#include <thread>
#include <vector>
#include <cstring>
#include <algorithm>
#include <string>

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

// ---- reads with their windows (re-anchoring path) ----------------------------------------------
namespace {
const char kB[4] = {'A', 'C', 'G', 'T'};
// truth[t0..t1] through the channel, appended to out; if map != nullptr also records the truth index
// every emitted base belongs to.
void channel(const cg_synth_read_spec* s, std::mt19937_64& rng, const std::vector<uint8_t>& truth,
             uint32_t t0, uint32_t t1, std::string& out, std::vector<uint32_t>* map) {
    for (uint32_t i = t0; i <= t1; ++i) {
        double u = (double)(rng() >> 11) * (1.0 / 9007199254740992.0);
        if (u >= s->err) {
            out.push_back(kB[truth[i]]); if (map) map->push_back(i);
        } else {
            double v = u / s->err;
            if (v < s->p_sub) {
                out.push_back(kB[(truth[i] + 1 + (uint32_t)(rng() % 3)) & 3]); if (map) map->push_back(i);
            } else if (v < s->p_sub + s->p_ins) {
                out.push_back(kB[rng() & 3]); if (map) map->push_back(i);
                out.push_back(kB[truth[i]]); if (map) map->push_back(i);
            }
        }
    }
}
}  // namespace

extern "C" void cg_synth_reads_bounds(const cg_synth_read_spec* s, uint64_t* max_windows, uint64_t* max_seqs,
                                      uint64_t* max_bases, uint64_t* max_read_bases) {
    uint64_t step = s->window_size > s->window_overlap ? s->window_size - s->window_overlap : 1;
    uint64_t per_read = (2ULL * s->truth_len) / step + 2;
    *max_windows = per_read * s->n_reads;
    *max_seqs = *max_windows * s->n_seqs;
    *max_bases = *max_seqs * (2ULL * s->window_size + 2);
    *max_read_bases = 2ULL * s->truth_len * s->n_reads;
}

extern "C" uint64_t cg_synth_reads(const cg_synth_read_spec* s, uint32_t* win_seq_begin, uint64_t* seq_off, char* bases,
                                   uint32_t* read_win_begin, uint64_t* read_off, char* read_bases, uint32_t* win_pos) {
    uint64_t W = 0, S = 0, B = 0, RB = 0;
    win_seq_begin[0] = 0; seq_off[0] = 0; read_win_begin[0] = 0; read_off[0] = 0;
    const uint32_t ws = s->window_size, ov = s->window_overlap;
    for (uint32_t r = 0; r < s->n_reads; ++r) {
        std::mt19937_64 rng(s->seed * 7000003ULL + s->first_read + r);
        std::vector<uint8_t> truth(s->truth_len);
        for (auto& t : truth) t = (uint8_t)(rng() & 3);
        std::string read; std::vector<uint32_t> map;
        channel(s, rng, truth, 0, s->truth_len - 1, read, &map);
        const uint32_t len = (uint32_t)read.size();
        std::vector<uint32_t> starts;
        if (len >= ws) {
            for (uint32_t b = 0; b + ws <= len; b += ws - ov) starts.push_back(b);
            if (len > ws) starts.push_back(len - ws);
        }
        for (size_t i = 0; i < starts.size(); ++i) {
            uint32_t a = starts[i], b = a + ws - 1;
            uint32_t n = (s->thin_every && (W % s->thin_every) == s->thin_every - 1) ? s->thin_seqs : s->n_seqs;
            if (n < 1) n = 1;
            memcpy(bases + B, read.data() + a, ws); B += ws; seq_off[++S] = B;
            for (uint32_t q = 1; q < n; ++q) {
                std::string o;
                channel(s, rng, truth, map[a], map[b], o, nullptr);
                memcpy(bases + B, o.data(), o.size()); B += o.size(); seq_off[++S] = B;
            }
            win_pos[W] = a;
            win_seq_begin[++W] = (uint32_t)S;
        }
        memcpy(read_bases + RB, read.data(), len); RB += len;
        read_off[r + 1] = RB;
        read_win_begin[r + 1] = (uint32_t)W;
    }
    return W;
}

// ---- read piles over a random genome (window-extraction path) ------------------------------------------------------
namespace {
struct PileSet {
    std::vector<uint64_t> store_off;
    std::string store;
    std::vector<uint32_t> pile_read, pile_qlen, pile_ov_begin, ov7;
};
}  // namespace

extern "C" void* cg_synth_piles_build(const cg_synth_pile_spec* s, uint64_t* store_bases, uint64_t* n_overlaps) {
    PileSet* ps = new PileSet();
    std::mt19937_64 rng(s->seed * 9000011ULL);
    std::vector<uint8_t> genome(s->genome_len);
    for (auto& g : genome) g = (uint8_t)(rng() & 3);
    const uint32_t R = s->n_reads, L = s->read_len;
    std::vector<uint32_t> start(R), strand(R);
    std::vector<std::vector<uint32_t>> gpos(R);               // genome position of every read base
    ps->store_off.push_back(0);
    cg_synth_read_spec ch{};
    ch.err = s->err; ch.p_sub = s->p_sub; ch.p_ins = s->p_ins;
    for (uint32_t r = 0; r < R; ++r) {
        start[r] = (uint32_t)(rng() % (s->genome_len - L + 1));
        strand[r] = (uint32_t)(rng() & 1);
        std::vector<uint8_t> seg(L);
        for (uint32_t i = 0; i < L; ++i) seg[i] = strand[r] ? (uint8_t)(3 - genome[start[r] + L - 1 - i]) : genome[start[r] + i];
        std::string read; std::vector<uint32_t> map;
        channel(&ch, rng, seg, 0, L - 1, read, &map);
        gpos[r].resize(map.size());
        for (size_t i = 0; i < map.size(); ++i) gpos[r][i] = strand[r] ? start[r] + L - 1 - map[i] : start[r] + map[i];
        ps->store += read;
        ps->store_off.push_back(ps->store.size());
    }
    // read positions whose genome coordinate lies in [g0, g1]: a contiguous range because gpos is monotone
    auto range = [&](uint32_t r, uint32_t g0, uint32_t g1, uint32_t& a, uint32_t& b) {
        const std::vector<uint32_t>& gp = gpos[r];
        if (!strand[r]) {
            a = (uint32_t)(std::lower_bound(gp.begin(), gp.end(), g0) - gp.begin());
            b = (uint32_t)(std::upper_bound(gp.begin(), gp.end(), g1) - gp.begin());
        } else {                                               // descending
            a = (uint32_t)(std::lower_bound(gp.begin(), gp.end(), g1, [](uint32_t x, uint32_t v) { return x > v; }) - gp.begin());
            b = (uint32_t)(std::upper_bound(gp.begin(), gp.end(), g0, [](uint32_t v, uint32_t x) { return v > x; }) - gp.begin());
        }
        return b > a;                                          // [a, b)
    };
    std::vector<uint32_t> by_start(R);
    for (uint32_t r = 0; r < R; ++r) by_start[r] = r;
    std::sort(by_start.begin(), by_start.end(), [&](uint32_t x, uint32_t y) { return start[x] < start[y] || (start[x] == start[y] && x < y); });
    ps->pile_ov_begin.push_back(0);
    const uint32_t P = std::min(s->n_piles, R);
    for (uint32_t q = 0; q < P; ++q) {
        struct Cand { uint32_t len, t, g0, g1; };
        std::vector<Cand> cand;
        const uint32_t q0 = start[q], q1 = start[q] + L - 1;
        auto lo = std::lower_bound(by_start.begin(), by_start.end(), q0 > L ? q0 - L : 0u, [&](uint32_t r, uint32_t v) { return start[r] < v; });
        for (auto it = lo; it != by_start.end() && start[*it] <= q1; ++it) {
            const uint32_t t = *it;
            if (t == q) continue;
            const uint32_t g0 = std::max(q0, start[t]), g1 = std::min(q1, start[t] + L - 1);
            if (g1 < g0 || g1 - g0 + 1 < s->min_overlap) continue;
            cand.push_back({g1 - g0 + 1, t, g0, g1});
        }
        std::sort(cand.begin(), cand.end(), [](const Cand& a, const Cand& b) { return a.len > b.len || (a.len == b.len && a.t < b.t); });
        if (cand.size() > s->max_support) cand.resize(s->max_support);
        for (const Cand& c : cand) {
            uint32_t qa, qb, ta, tb;
            if (!range(q, c.g0, c.g1, qa, qb) || !range(c.t, c.g0, c.g1, ta, tb)) continue;
            const uint32_t rec[7] = {c.t, strand[q] ^ strand[c.t], qa, qb - 1, ta, tb - 1, (uint32_t)gpos[c.t].size()};
            ps->ov7.insert(ps->ov7.end(), rec, rec + 7);
        }
        ps->pile_read.push_back(q);
        ps->pile_qlen.push_back((uint32_t)gpos[q].size());
        ps->pile_ov_begin.push_back((uint32_t)(ps->ov7.size() / 7));
    }
    *store_bases = ps->store.size();
    *n_overlaps = ps->ov7.size() / 7;
    return ps;
}

extern "C" void cg_synth_piles_fetch(void* handle, uint64_t* store_off, char* store, uint32_t* pile_read, uint32_t* pile_qlen,
                                     uint32_t* pile_ov_begin, uint32_t* overlaps7) {
    PileSet* ps = static_cast<PileSet*>(handle);
    memcpy(store_off, ps->store_off.data(), ps->store_off.size() * 8);
    memcpy(store, ps->store.data(), ps->store.size());
    if (!ps->pile_read.empty()) {
        memcpy(pile_read, ps->pile_read.data(), ps->pile_read.size() * 4);
        memcpy(pile_qlen, ps->pile_qlen.data(), ps->pile_qlen.size() * 4);
    }
    memcpy(pile_ov_begin, ps->pile_ov_begin.data(), ps->pile_ov_begin.size() * 4);
    if (!ps->ov7.empty()) memcpy(overlaps7, ps->ov7.data(), ps->ov7.size() * 4);
    delete ps;
}

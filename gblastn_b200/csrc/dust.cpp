// dust.cpp — symmetric DUST low-complexity masking of query sequences (host side).
//
// What the blastn command line does to its queries before the lookup table is built (`-dust yes`, the megablast /
// blastn task default "20 64 1"): Blast_FindDustFilterLoc (c++/src/algo/blast/api/dust_filter.cpp:65-151) runs
// CSymDustMasker (c++/src/algo/dustmask/symdust.cpp:213-319, include/algo/dustmask/symdust.hpp) over each query and
// the intervals become the query's masked locations (mask-at-hash: they only keep words out of the lookup table).
// SURVEY.md 8(f) rank 4; the masks feed bn_setup_create / lookup_segments.
//
// The algorithm (Morgulis, Gertz, Schaffer, Agarwala 2006): slide a window of at most `window` bases over the
// sequence as a queue of triplets (3-mers, 64 values).  A stretch of triplets scores sum_t c_t (c_t - 1) / 2 over its
// triplet counts; it is "perfect" when 10 * score exceeds level * (its length in triplets) and no sub-stretch of it
// scores proportionally higher.  Perfect intervals are collected per window position and written out, merged when
// they lie within `linker` bases of each other, as soon as the window has moved past their left end.
// The state below mirrors the reference's so that the intervals are identical, including the special handling of
// windows that hold one distinct triplet (homopolymers) and the restart of the scan that follows such a stretch:
//   w_count / w_sum : triplet counts and running score of the whole window
//   v_count / v_sum : the same for the window's longest suffix in which no triplet occurs more than level / 5 times
//                     (shorter suffixes cannot reach the threshold, so candidate stretches start left of it)
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/gblastn_b200.h"

namespace {

struct Perfect { uint32_t first, last, score, len; };      // bases [first, last], score, length in triplets

class Window {
public:
    Window(uint32_t window, uint32_t low_k, std::vector<Perfect> &perfect, const std::vector<uint32_t> &thresholds)
        : max_size_(window - 2), low_k_(low_k), P(perfect), thr_(thresholds)
    {
        ring_.assign(64, 0);
        memset(w_count_, 0, sizeof w_count_);
        memset(v_count_, 0, sizeof v_count_);
    }
    uint32_t start() const { return start_; }
    uint32_t size() const { return stop_ - start_; }

    // push triplet t at the right end; false when the (full) window then holds a single distinct triplet
    bool shift(uint8_t t)
    {
        if (size() >= max_size_) {
            if (distinct_ <= 1) return shift_uniform(t);
            const uint8_t s = at(start_);
            remove(w_sum_, w_count_, s);
            if (w_count_[s] == 0) --distinct_;
            if (suffix_ == start_) { ++suffix_; remove(v_sum_, v_count_, s); }
            ++start_;
        }
        put(stop_, t);
        if (w_count_[t] == 0) ++distinct_;
        add(w_sum_, w_count_, t);
        add(v_sum_, v_count_, t);
        if (v_count_[t] > low_k_) {
            // shrink the suffix from the left until one occurrence of t has left it
            uint8_t x;
            do {
                x = at(suffix_);
                remove(v_sum_, v_count_, x);
                ++suffix_;
            } while (x != t);
        }
        ++stop_;
        if (size() >= max_size_ && distinct_ <= 1) {
            P.clear();
            P.insert(P.begin(), Perfect{start_, stop_ + 1, 0, 0});
            return false;
        }
        return true;
    }

    // Proposition 2 of the paper: the suffix scan can be skipped when the whole window is below the threshold
    bool needs_processing() const
    {
        const uint32_t count = stop_ - suffix_;
        return count < size() && 10 * w_sum_ > thr_[count];
    }

    // extend the low-count suffix leftwards one triplet at a time and record every stretch that is perfect
    void find_perfect()
    {
        uint32_t count = stop_ - suffix_;
        uint8_t counts[64];
        memcpy(counts, v_count_, sizeof counts);
        uint32_t score = v_sum_;
        size_t pi = 0;                           // position in P (ordered by descending left end)
        uint32_t best_score = 0, best_len = 0;
        uint32_t pos = suffix_ - 1;
        for (uint32_t k = count; k < size(); ++k, ++count, --pos) {
            const uint8_t t = at(stop_ - 1 - k);
            const uint8_t before = counts[t];
            score += counts[t]; ++counts[t];
            if (before > 0 && score * 10 > thr_[count]) {
                // the best score-per-length among the recorded intervals inside the current stretch
                while (pi < P.size() && pos <= P[pi].first) {
                    if (best_score == 0 || (uint64_t)best_len * P[pi].score > (uint64_t)best_score * P[pi].len) {
                        best_score = P[pi].score;
                        best_len = P[pi].len;
                    }
                    ++pi;
                }
                if (best_score == 0 || (uint64_t)score * best_len >= (uint64_t)best_score * count) {
                    best_score = score;
                    best_len = count;
                    P.insert(P.begin() + (long)pi, Perfect{pos, stop_ + 1, best_score, count});
                }
            }
        }
    }

private:
    // a full window of one repeated triplet slides without touching the suffix bookkeeping
    bool shift_uniform(uint8_t t)
    {
        const uint8_t s = at(start_);
        remove(w_sum_, w_count_, s);
        if (w_count_[s] == 0) --distinct_;
        ++start_;
        put(stop_, t);
        if (w_count_[t] == 0) ++distinct_;
        add(w_sum_, w_count_, t);
        ++stop_;
        if (distinct_ <= 1) {
            P.insert(P.begin(), Perfect{start_, stop_ + 1, 0, 0});
            return false;
        }
        return true;
    }
    static void add(uint32_t &sum, uint8_t *c, uint8_t t) { sum += c[t]; ++c[t]; }
    static void remove(uint32_t &sum, uint8_t *c, uint8_t t) { --c[t]; sum -= c[t]; }
    uint8_t at(uint32_t pos) const { return ring_[pos & 63u]; }
    void put(uint32_t pos, uint8_t t) { ring_[pos & 63u] = t; }

    std::vector<uint8_t> ring_;          // triplet at window position p lives at p mod 64 (a window holds <= 62)
    uint32_t start_ = 0, stop_ = 0;      // positions of the oldest triplet / one past the newest
    uint32_t max_size_, low_k_;
    uint32_t suffix_ = 0;                // position where the low-count suffix starts
    std::vector<Perfect> &P;
    const std::vector<uint32_t> &thr_;
    uint8_t w_count_[64], v_count_[64];
    uint32_t w_sum_ = 0, v_sum_ = 0, distinct_ = 0;
};

struct Masker {
    uint32_t level, window, linker;
    std::vector<uint32_t> thresholds;
    std::vector<Perfect> P;
    std::vector<int32_t> out;            // flat [from, to] pairs

    Masker(uint32_t lv, uint32_t w, uint32_t lk)
        : level((lv >= 2 && lv <= 64) ? lv : 20), window((w >= 8 && w <= 64) ? w : 64), linker((lk >= 1 && lk <= 32) ? lk : 1)
    {
        thresholds.push_back(1);
        for (uint32_t i = 1; i < window - 2; ++i) thresholds.push_back(i * level);
    }

    // write out the perfect intervals whose left end the window has passed
    void flush_passed(uint32_t window_start, uint32_t offset)
    {
        if (P.empty()) return;
        const Perfect b = P.back();
        if (b.first >= window_start) return;
        const int32_t from = (int32_t)(b.first + offset), to = (int32_t)(b.last + offset);
        if (!out.empty() && (uint32_t)out.back() + linker >= (uint32_t)from) out.back() = std::max(out.back(), to);
        else { out.push_back(from); out.push_back(to); }
        while (!P.empty() && P.back().first < window_start) P.pop_back();
    }

    void run(const uint8_t *seq, uint32_t len)
    {
        if (len == 0) return;
        auto base = [&](uint32_t p) -> uint8_t { return (p < len && seq[p] < 4) ? seq[p] : 0; };   // non-ACGT reads as A
        uint32_t start = 0;
        const uint32_t stop = len - 1;
        while (stop > 2 + start) {
            P.clear();
            Window w(window, level / 5, P, thresholds);
            uint8_t t = (uint8_t)((base(start) << 2) + base(start + 1));
            uint32_t pos = start + 2;
            bool done = false;
            while (!done && pos <= stop) {
                flush_passed(w.start(), start);
                t = (uint8_t)(((t << 2) & 0x3F) + base(pos));
                ++pos;
                if (w.shift(t)) {
                    if (w.needs_processing()) w.find_perfect();
                } else {
                    // inside a run of one repeated triplet: slide until it ends, then start over behind it
                    while (pos <= stop) {
                        flush_passed(w.start(), start);
                        t = (uint8_t)(((t << 2) & 0x3F) + base(pos));
                        if (w.shift(t)) { done = true; break; }
                        ++pos;
                    }
                }
            }
            uint32_t ws = w.start();
            while (!P.empty()) { flush_passed(ws, start); ++ws; }
            if (w.start() > 0) start += w.start();
            else break;
        }
    }
};

}  // namespace

extern "C" int bn_dust_mask(const uint8_t *seq, int32_t len, int32_t level, int32_t window, int32_t linker,
                            int32_t **intervals, int32_t *n_intervals)
{
    if (!intervals || !n_intervals || len < 0 || (len > 0 && !seq)) return BN_ERR_INVALID;
    Masker m((uint32_t)std::max(level, 0), (uint32_t)std::max(window, 0), (uint32_t)std::max(linker, 0));
    m.run(seq, (uint32_t)len);
    *n_intervals = (int32_t)(m.out.size() / 2);
    *intervals = (int32_t *)malloc(sizeof(int32_t) * std::max<size_t>(m.out.size(), 2));
    if (!*intervals) return BN_ERR_MEMORY;
    if (!m.out.empty()) memcpy(*intervals, m.out.data(), m.out.size() * sizeof(int32_t));
    return BN_OK;
}

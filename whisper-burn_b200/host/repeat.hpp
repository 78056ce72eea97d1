// Repetition heuristics of the reference's transcribe.rs (:385-447).  The reference compiles these three functions but only
// its commented-out greedy loop (:314-380) calls them; they are kept here with the same results, including the inputs on
// which the reference panics (reported as `invalid` instead of aborting).  Host-side integer work, no device code.
#pragma once
#include <cstdint>

namespace wb {
namespace repeat {

inline bool same(const int64_t* a, const int64_t* b, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        if (a[i] != b[i]) return false;
    return true;
}

// transcribe.rs:385-393.  Walks back from the end over positions whose preceding `period` tokens equal the following ones;
// returns the position after the first mismatch, or `period`.  period > n underflows in the reference: invalid (-1).
inline int64_t first_repetition_end(const int64_t* tokens, int64_t n, int64_t period) {
    if (period < 0 || period > n) return -1;
    for (int64_t i = n - period - 1; i >= period; --i)
        if (!same(tokens + i - period, tokens + i, period)) return i + 1;
    return period;
}

// transcribe.rs:395-419.  The shortest suffix length `period` such that the `min_repetitions` blocks before the suffix all
// equal it; 0 when the search runs out of room (the reference's None).
inline int64_t repetition_period(const int64_t* tokens, int64_t n, int64_t min_repetitions) {
    if (min_repetitions < 0) return -1;
    for (int64_t i = n - 1; i >= 0; --i) {
        const int64_t period = n - i;
        if (i / period < min_repetitions) return 0;
        bool all = true;
        for (int64_t j = 0; j < min_repetitions && all; ++j) {
            const int64_t e = i - period * j, s = e - period;
            all = same(tokens + s, tokens + i, period);
        }
        if (all) return period;
    }
    return 0;
}

// transcribe.rs:421-447.  Windows of `window_size` tokens equal to the last one (the last window itself and windows overlapping
// it are not candidates): 1 and (index of the first such window, index of the second) when at least `min_repeat_count` exist,
// 0 otherwise; -1 where the reference unwraps a missing second repeat (min_repeat_count < 2 with fewer than two repeats).
inline int find_repeated_tokens_index(const int64_t* tokens, int64_t n, int64_t window_size, int64_t min_repeat_count, int64_t* first_repeat_index,
                                      int64_t* end) {
    if (window_size < 0 || min_repeat_count < 0) return -1;
    if (2 * window_size > n) return 0;
    const int64_t last_index = n - window_size;
    int64_t n_repeats = 0, r0 = -1, r1 = -1;
    for (int64_t i = 0; i + window_size <= last_index; ++i) {
        if (same(tokens + i, tokens + last_index, window_size)) {
            if (n_repeats == 0) r0 = i;
            else if (n_repeats == 1) r1 = i;
            ++n_repeats;
        }
    }
    if (n_repeats < min_repeat_count) return 0;
    if (n_repeats < 2) return -1;
    *first_repeat_index = r0;
    *end = r1;
    return 1;
}

}  // namespace repeat
}  // namespace wb

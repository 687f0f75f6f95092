// take_sampler.h -- sampling without replacement as lib/mnist_csv2.c:41-62 (get_random_data_take) does it, minus the scan.
//
// The reference draws r = floor((examples - num_sampled) * rand() / RAND_MAX) and walks the `sampled` flags from the start
// until it has passed r unsampled examples: O(examples) per draw, 1.8e9 flag reads per 60,000-example epoch.  It then takes the
// example it stopped AT -- the one after the r-th unsampled example (example 0 when r == 0), which may already be sampled --
// marks it, and counts the draw.  Here the flags are mirrored by a Fenwick tree of "not sampled yet" counts, so the r-th
// unsampled example is found in O(log examples); the libc rand() stream, the float arithmetic of r and the index rule are the
// reference's, hence the same index sequence (tests/test_data_cpu.py, tests/test_host_io_cpu.py against the compiled reference).
// Host-only; shared by the device data pipeline (data.cu) and the drop-in lib/mnist_csv2.h (host_io.cu).
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace bla {

struct TakeSampler {
    int n = 0;
    std::vector<int> tree;   // 1-based Fenwick tree over flags "example i - 1 has not been sampled"

    void reset_all_unsampled(int examples) {
        n = examples;
        tree.assign((size_t)n + 1, 0);
        for (int i = 1; i <= n; ++i) {
            tree[i] += 1;
            const int up = i + (i & -i);
            if (up <= n) tree[up] += tree[i];
        }
    }
    // mirror flags somebody else maintains (the caller of lib/mnist_csv2.h owns `sampled` and may rewrite it)
    void rebuild(const char* sampled, int examples) {
        n = examples;
        tree.assign((size_t)n + 1, 0);
        for (int i = 1; i <= n; ++i) {
            tree[i] += sampled[i - 1] ? 0 : 1;
            const int up = i + (i & -i);
            if (up <= n) tree[up] += tree[i];
        }
    }
    // 0-based index of the k-th (k >= 1) unsampled example; n if there are fewer than k
    int find(int k) const {
        int pos = 0, step = 1;
        while (step * 2 <= n) step *= 2;
        for (; step > 0; step >>= 1)
            if (pos + step <= n && tree[pos + step] < k) { pos += step; k -= tree[pos]; }
        return pos;
    }
    void mark_sampled(int i) {
        for (int j = i + 1; j <= n; j += j & -j) tree[j] -= 1;
    }
    // One draw.  `sampled` (n flags) and *num_sampled are the reference's public state and are updated as it updates them.
    // Where the reference runs off the end of its arrays (rand() == RAND_MAX on the last draws; an empty data set) the last
    // example is taken instead.
    int take(char* sampled, int* num_sampled) {
        if (n <= 0) return 0;
        if (*num_sampled == n) {                                                        // :43-46 start over
            memset(sampled, 0, (size_t)n);
            *num_sampled = 0;
            reset_all_unsampled(n);
        }
        const int r = (int)floor((float)(n - *num_sampled) * (float)rand() / (float)RAND_MAX);   // :49
        int i = r > 0 ? find(r) + 1 : 0;                                                // :52-57 one past the r-th unsampled
        if (i >= n) i = n - 1;
        if (!sampled[i]) { sampled[i] = 1; mark_sampled(i); }
        ++*num_sampled;                                                                 // :59 counted even when i had been taken
        return i;
    }
};

}  // namespace bla

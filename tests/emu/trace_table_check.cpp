// test-only: psb_result_extra::trace_table() (csrc/psb_internal.h) against the layout of the device's flag-byte block,
// [strip][step][lane][K] with cell (i, j) at step j + lane -- a synthetic block, no GPU.  Also times a large table.
// usage: trace_table_check qlen rlen K
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "../../parasail_rs_b200/csrc/psb_internal.h"

static inline uint8_t cell_value(int i, int j) { return (uint8_t)((i * 131 + j * 7 + (i ^ j)) & 0x7f); }

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const int qlen = std::atoi(argv[1]), rlen = std::atoi(argv[2]), K = std::atoi(argv[3]);
    psb_result_extra x;
    x.qlen = qlen; x.rlen = rlen; x.trace_K = K;
    const int rows = 32 * K, nsteps = rlen + 31, strips = (qlen + rows - 1) / rows;
    x.trace_blob.assign((size_t)strips * nsteps * 32 * K, 0xee);
    for (int i = 0; i < qlen; ++i) {
        const int strip = i / rows, rem = i % rows, lane = rem / K, k = rem % K;
        for (int j = 0; j < rlen; ++j) x.trace_blob[(((size_t)strip * nsteps + (j + lane)) * 32 + lane) * K + k] = cell_value(i, j);
    }
    const auto t0 = std::chrono::steady_clock::now();
    const int8_t *t = x.trace_table();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (!t) { std::puts("no table"); return 1; }
    for (int i = 0; i < qlen; ++i)
        for (int j = 0; j < rlen; ++j)
            if ((uint8_t)t[(size_t)i * rlen + j] != cell_value(i, j)) { std::printf("mismatch at (%d, %d)\n", i, j); return 1; }
    if (t != x.trace_table()) { std::puts("second call returned another table"); return 1; }
    std::printf("ok %d x %d K %d: %.1f ms\n", qlen, rlen, K, ms);
    return 0;
}

// dpx_bench.cu -- measures the issue rate of the integer instructions the Gotoh kernels are made
// of ("replace 64 lane-ops/clk/SM by the result of a first-day microbenchmark").
// Each test runs NCHAIN independent dependency chains per thread, enough warps to fill every
// SMSP, and reports warp-instructions per clock per SM (x32 = lane-ops/clk/SM).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int NCHAIN = 8;
constexpr int ITERS = 4096;

enum Op { OP_VIADDMNMX16 = 0, OP_VIMNMX3_16, OP_VIADDMNMX32, OP_VIMNMX3_32, OP_PRMT, OP_IMAD, OP_LEA, OP_IADD3,
          OP_MIX_DPX_IMAD, OP_MIX_CELL, OP_IMADHI, OP_MIX_CELL2, OP_COUNT };
static const char *kNames[] = {"VIADDMNMX.S16x2", "VIMNMX3.S16x2", "VIADDMNMX(s32)", "VIMNMX3(s32)", "PRMT", "IMAD", "LEA",
                               "IADD3", "VIADDMNMX.S16x2+IMAD(1:1)", "cell(4DPX+PRMT+2IMAD)", "IMAD.HI.U32",
                               "cellpair x2 (8DPX+VIMNMX3 | 5IMAD+2IMAD.HI)"};

template <int OP>
__global__ void bench(unsigned *out, unsigned a0, unsigned b0, unsigned c0, unsigned m1) {
    unsigned x[NCHAIN], y[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) { x[i] = a0 + i * 3 + threadIdx.x; y[i] = b0 + i + 7 * threadIdx.x; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            if (OP == OP_VIADDMNMX16) x[i] = __viaddmax_s16x2(x[i], b0, c0);
            if (OP == OP_VIMNMX3_16) x[i] = __vimax3_s16x2(x[i], y[i], c0 + it);
            if (OP == OP_VIADDMNMX32) x[i] = (unsigned)__viaddmax_s32((int)x[i], (int)b0, (int)c0);
            if (OP == OP_VIMNMX3_32) x[i] = (unsigned)__vimax3_s32((int)x[i], (int)y[i], (int)(c0 + it));
            if (OP == OP_PRMT) { unsigned d; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x[i]), "r"(y[i]), "r"(0xC480u)); x[i] = d; }
            if (OP == OP_IMAD) x[i] = x[i] * m1 + c0;
            if (OP == OP_LEA) { unsigned d; asm volatile("shl.b32 %0, %1, 4;" : "=r"(d) : "r"(x[i])); x[i] = d + y[i]; }
            if (OP == OP_IADD3) { unsigned d; asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(x[i]), "r"(y[i])); x[i] = d; }
            if (OP == OP_MIX_DPX_IMAD) { x[i] = __viaddmax_s16x2(x[i], b0, c0); y[i] = y[i] * m1 + c0; }
            if (OP == OP_IMADHI) x[i] = __umulhi(x[i], m1 + 65535u) + c0;
            if (OP == OP_MIX_CELL2) {
                // two cell-pairs of the scan kernel with the score interleave moved to the FMA pipe:
                // ALU: 8 DPX + 1 VIMNMX3; FMA: 3 IMAD + 2 IMAD.HI (interleave) + 2 IMAD (T = X - o)
                const unsigned m64k = m1 + 65535u;
                unsigned xa = y[i], yb = y[(i + 1) % NCHAIN];
                unsigned a1 = __umulhi(xa, m64k);
                unsigned a0 = xa - a1 * m64k;
                unsigned s0 = yb * m64k + a0;
                unsigned b1 = __umulhi(yb, m64k);
                unsigned s1 = b1 * m64k + a1;
                unsigned e0 = __viaddmax_s16x2(y[i], b0, x[i]);
                unsigned f0 = __viaddmax_s16x2(x[(i + 1) % NCHAIN], b0, y[i]);
                unsigned h0 = __viaddmax_s16x2(x[i], s0, e0);
                unsigned X0 = __vimax3_s16x2(h0, f0, c0);
                unsigned t0 = X0 * m1 + b0;
                unsigned e1 = __viaddmax_s16x2(e0, b0, t0);
                unsigned f1 = __viaddmax_s16x2(f0, b0, t0);
                unsigned h1 = __viaddmax_s16x2(t0, s1, e1);
                unsigned X1 = __vimax3_s16x2(h1, f1, c0);
                unsigned t1 = X1 * m1 + b0;
                x[i] = __vimax3_s16x2(x[i], t0, t1);
                y[i] = t1 + e1;
            }
            if (OP == OP_MIX_CELL) {
                unsigned s; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(s) : "r"(y[i]), "r"(c0), "r"(0xC480u));
                unsigned e = __viaddmax_s16x2(y[i], b0, x[i]);
                unsigned f = __viaddmax_s16x2(x[(i + 1) % NCHAIN], b0, y[i]);
                unsigned h = __viaddmax_s16x2(x[i], s, e);
                unsigned X = __vimax3_s16x2(h, f, c0);
                x[i] = X * m1 + b0;
                y[i] = x[i] * m1 + e;
            }
        }
    }
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) acc ^= x[i] ^ y[i];
    if (acc == 0x12345678u) out[0] = acc;
}

template <int OP> static double run(int sms, int warps_per_sm, double *clk_ghz) {
    unsigned *d; CHECK(cudaMalloc(&d, 64));
    const int threads = 256, blocks = sms * warps_per_sm * 32 / threads;
    cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    bench<OP><<<blocks, threads>>>(d, 1, 2, 3, 1);  // warm-up
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CHECK(cudaEventRecord(e0));
        bench<OP><<<blocks, threads>>>(d, 1, 2, 3, 1);
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    int per_iter = (OP == OP_MIX_DPX_IMAD) ? 2 : (OP == OP_MIX_CELL ? 7 : (OP == OP_MIX_CELL2 ? 17 : 1));
    double warp_instr = (double)blocks * (threads / 32) * ITERS * NCHAIN * per_iter;
    double per_sm_per_s = warp_instr / sms / (best * 1e-3);
    cudaFree(d);
    return per_sm_per_s / (*clk_ghz * 1e9);
}

int main(int argc, char **argv) {
    const char *path = argc > 1 ? argv[1] : "gpurun_out/dpx_peak.json";
    cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    int clk_khz = 0; CHECK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    double ghz = clk_khz * 1e-6;  // nominal max clock; the bench reports per-clock rates against it
    double r[OP_COUNT];
    r[0] = run<0>(sms, 32, &ghz); r[1] = run<1>(sms, 32, &ghz); r[2] = run<2>(sms, 32, &ghz); r[3] = run<3>(sms, 32, &ghz);
    r[4] = run<4>(sms, 32, &ghz); r[5] = run<5>(sms, 32, &ghz); r[6] = run<6>(sms, 32, &ghz); r[7] = run<7>(sms, 32, &ghz);
    r[8] = run<8>(sms, 32, &ghz); r[9] = run<9>(sms, 32, &ghz); r[10] = run<10>(sms, 32, &ghz); r[11] = run<11>(sms, 16, &ghz);
    FILE *f = fopen(path, "w");
    if (!f) f = stdout;
    fprintf(f, "{\"gpu\": \"%s\", \"sms\": %d, \"assumed_clock_ghz\": %.3f, \"unit\": \"warp-instr/clk/SM at the assumed clock (x32 = lane-ops)\", \"rates\": {", prop.name, sms, ghz);
    for (int i = 0; i < OP_COUNT; ++i) fprintf(f, "%s\"%s\": %.3f", i ? ", " : "", kNames[i], r[i]);
    fprintf(f, "}}\n");
    if (f != stdout) fclose(f);
    for (int i = 0; i < OP_COUNT; ++i) printf("%-28s %.3f warp-instr/clk/SM  (%.1f lane-ops/clk/SM)\n", kNames[i], r[i], r[i] * 32);
    return 0;
}

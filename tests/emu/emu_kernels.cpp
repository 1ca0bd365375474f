// emu_kernels.cpp -- TEST ONLY.  Compiles the product's kernel sources with -DPSB_EMULATE
// (every lane an OS thread, see csrc/psb_simt.h) so that the warp programs can be stepped
// on a CPU-only box and compared with the oracle before any GPU time is spent.
#define PSB_EMULATE 1
#ifdef EMU_NO_PINGPONG
#define SW16_PINGPONG 0
#endif
#include "../../parasail_rs_b200/csrc/kern_gotoh32.cuh"
#include <cstdio>

using namespace psb;

static bool g_prof = false;
template <int K, bool STATS, bool TRACE, typename W>
static void run32(const Gotoh32Params &p, int nblocks) {
    size_t smem = gotoh32_smem_bytes(p.size, 1, STATS, sizeof(W), g_prof);
    if (p.tabH) emu::launch(nblocks, smem, [&]() { gotoh32_kernel<K, STATS, TRACE, true, W>(p); });
    else if (g_prof) emu::launch(nblocks, smem, [&]() { gotoh32_kernel<K, STATS, TRACE, false, W, true>(p); });
    else emu::launch(nblocks, smem, [&]() { gotoh32_kernel<K, STATS, TRACE, false, W>(p); });
}

extern "C" void emu_gotoh32_use_profile(int on) { g_prof = on != 0; }

extern "C" int emu_gotoh32(int K, int stats, int trace, int wide_stats, const Gotoh32Params *pp, int nblocks) {
    Gotoh32Params p = *pp;
#define CASE(KK)                                                                              \
    case KK:                                                                                  \
        if (trace) run32<KK, false, true, unsigned>(p, nblocks);                              \
        else if (stats && wide_stats) run32<KK, true, false, unsigned long long>(p, nblocks); \
        else if (stats) run32<KK, true, false, unsigned>(p, nblocks);                         \
        else run32<KK, false, false, unsigned>(p, nblocks);                                   \
        return 0;
    switch (K) {
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(8) CASE(16)
    }
    return -1;
}
extern "C" int emu_sizeof_params() { return (int)sizeof(Gotoh32Params); }

// ---- packed 16-bit scan kernel ------------------------------------------------------------------
#include "../../parasail_rs_b200/csrc/kern_sw16.cuh"

extern "C" int emu_sw16_build(const uint8_t *mapped_query, int lq, const int *table, int size, int open,
                              int8_t *out, int cap, int *K, int *max_score, int *chunks) {
    Sw16Profile pr;
    std::vector<int8_t> host;
    if (!sw16_build_profile(mapped_query, lq, table, size, open, &pr, &host)) return -1;
    if ((int)host.size() > cap) return -2;
    std::memcpy(out, host.data(), host.size());
    *K = pr.K; *max_score = pr.max_score; *chunks = pr.chunks;
    return (int)host.size();
}

extern "C" int emu_sw16(int K, const Sw16Params *pp, int nblocks) {
    Sw16Params p = *pp;
    size_t smem = sw16_smem_bytes(p.nletters, K, 1);
#define SCASE(KK) case KK: emu::launch(nblocks, smem, [&]() { sw16_scan_kernel<KK>(p); }); return 0;
    switch (K) { SCASE(4) SCASE(8) SCASE(12) SCASE(16) SCASE(20) SCASE(25) SCASE(28) SCASE(32) }
    return -1;
}
extern "C" int emu_sizeof_sw16() { return (int)sizeof(Sw16Params); }

// experimental split-column variant (csrc/kern_sw16x.cuh)
#include "../../parasail_rs_b200/csrc/kern_sw16x.cuh"
extern "C" int emu_sw16x(int K, const Sw16Params *pp, int nblocks) {
    Sw16Params p = *pp;
    size_t smem = sw16_smem_bytes(p.nletters, K, 1);
#define SXCASE(KK) case KK: emu::launch(nblocks, smem, [&]() { sw16x_scan_kernel<KK>(p); }); return 0;
    switch (K) { SXCASE(8) SXCASE(12) SXCASE(16) SXCASE(20) SXCASE(25) SXCASE(28) SXCASE(32) }
    return -1;
}

// ---- long-pair wavefront kernel -------------------------------------------------------------------
#include "../../parasail_rs_b200/csrc/kern_wave32.cuh"

static bool g_wave_v2 = false;
static int g_wave_gen = 1;   // 1, 2, or 3 (column-blocked, local alignment only)
extern "C" void emu_wave32_use_v2(int on) { g_wave_v2 = on == 1; g_wave_gen = on == 2 ? 3 : (on == 1 ? 2 : 1); }
extern "C" int emu_wave32(int K, const Wave32Params *pp, const WaveReduceParams *rp, int nblocks) {
    Wave32Params p = *pp;
    if (g_wave_gen == 3) {
        size_t smem3 = wave32v3_smem_bytes(p.size, 1);
        switch (K) {
#define W3CASE(KK) case KK: if (p.mode == MODE_SW) emu::launch(nblocks, smem3, [&]() { wave32v3_kernel<KK, 4, true>(p); }); else emu::launch(nblocks, smem3, [&]() { wave32v3_kernel<KK, 4, false>(p); }); break;
            W3CASE(4) W3CASE(8) W3CASE(16)
            default: return -1;
        }
        WaveReduceParams r3 = *rp;
        emu::launch((r3.multi_n + 31) / 32 + (r3.multi_n == 0), 64, [&]() { wave32_reduce_kernel(r3); });
        return 0;
    }
    size_t smem = g_wave_v2 ? wave32v2_smem_bytes(p.size, 1) : wave32_smem_bytes(p.size, 1);
#define WCASE(KK) case KK: if (g_wave_v2) emu::launch(nblocks, smem, [&]() { wave32v2_kernel<KK>(p); }); else emu::launch(nblocks, smem, [&]() { wave32_kernel<KK>(p); }); break;
    switch (K) { WCASE(1) WCASE(2) WCASE(4) WCASE(8) default: return -1; }
    WaveReduceParams r = *rp;
    emu::launch((r.multi_n + 31) / 32 + (r.multi_n == 0), 64, [&]() { wave32_reduce_kernel(r); });
    return 0;
}
extern "C" int emu_sizeof_wave32() { return (int)sizeof(Wave32Params); }
extern "C" int emu_sizeof_wavereduce() { return (int)sizeof(WaveReduceParams); }

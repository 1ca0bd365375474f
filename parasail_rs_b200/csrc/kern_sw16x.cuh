// kern_sw16x.cuh -- EXPERIMENT (not part of the default build): the packed 16-bit scan kernel of
// kern_sw16.cuh with every lane's column split into two half-columns that run one column apart.
//
// Why: the shipped kernel keeps four warps per scheduler (120 registers) and each warp's step is one
// serial chain through its 25 rows (X -> T -> F', three dependent instructions per row), so the
// schedulers find nothing to issue in a third of the cycles (ncu: issue 68 %, ALU pipe 77 %, `wait` the
// top stall).  Here lane t fills rows [0, KU) of column s - 2t and rows [KU, K) of column s - 2t - 1 in
// the same step: two independent chains of half the length, the same registers, the same instructions
// per cell.  The group becomes a 32-stage systolic array (two stages per lane), so a subject pair
// costs 16 more steps of fill/drain (~4 % at 360 columns).
//
// MEASURED (B200, config C2, build with -DPSB_SW16X, tools/variant_bench.sh): 4 718 GCUPS against the
// shipped kernel's 5 140 (30.3 ms vs 27.85 ms per scan), results equal.  The extra parallelism inside
// a warp does not buy issue slots; the file stays as the record of that experiment.
//
// Bit-exactness is checked on the CPU emulation (tests/test_emu_sw16.py::test_split_column_variant).
#pragma once
#include "kern_sw16.cuh"

namespace psb {

template <int K>
PSB_KERNEL void SW16_BOUNDS sw16x_scan_kernel(Sw16Params p) {
    static_assert(SW16_PROF32 == 1, "the split-column variant is written for the 32-bit profile layout");
    static_assert(K >= 8, "needs at least two 4-row chunks");
    constexpr int G = SW16_G;
    constexpr int KU = ((K / 2 + 3) / 4) * 4;          // rows of the upper half-column: whole 4-row profile chunks
    constexpr int CH = (K + 3) / 4;                    // 16-byte profile loads per lane and letter
    constexpr int CHU = KU / 4;                        // ... of which the upper half-column takes the first CHU
    static_assert(KU > 0 && KU < K, "both half-columns need rows");
    constexpr unsigned LSTRIDE = CH * G * 16;          // bytes per letter
    PSB_SHARED_DECL(smem_raw);
    const int lane = lane_id();
    const int lg = lane & (G - 1);
    const int grp = lane >> 4;
    {
        const uint4 *src = (const uint4 *)p.prof;
        uint4 *dst = (uint4 *)smem_raw;
        const int n16 = p.nletters * (int)(LSTRIDE / 16);
        for (int x = thread_in_block(); x < n16; x += threads_per_block()) dst[x] = src[x];
    }
    sync_block();
    const unsigned char *lane_base = smem_raw + lg * 16;
    constexpr int PARKW = ((K + 3) / 4) * 4;
    unsigned char *wsm = smem_raw + (size_t)p.nletters * LSTRIDE + (size_t)warp_in_block() * (2 * 64 * 4 + 16 + 2 * 32 * PARKW * 4);
    unsigned *ring = (unsigned *)wsm + grp * 64;
    volatile unsigned *gpub = (volatile unsigned *)(wsm + 2 * 64 * 4) + grp;
    uint4 *park = (uint4 *)(wsm + 2 * 64 * 4 + 16);
    const int pad_code = p.nletters - 1;
    const unsigned O2 = (unsigned)p.open * 0x10001u;
    const unsigned NEGE = ((unsigned)(-p.gap) & 0xffffu) * 0x10001u;
    const unsigned NEGO = 0u - O2;
    const unsigned one = p.mul_one;
    const unsigned m64k = p.mul_64k;
    const long long nitems = (p.n + 1) >> 1;
    const long long nslots = (nitems + 1) >> 1;
    const int limit = 32767 - p.open - p.max_score;

    for (;;) {
        int slot = 0;
        if (lane == 0) slot = atomic_add(p.counter, 1);
        slot = shfl(slot, 0);
        if (slot >= nslots) break;
        const long long item = 2ll * slot + grp;
        const long long sa = 2 * item, sb = sa + 1;
        const int lenA = sa < p.n ? p.len[sa] : 0;
        const int lenB = sb < p.n ? p.len[sb] : 0;
        const long long wA = sa < p.n ? p.word_off[sa] : 0, wB = sb < p.n ? p.word_off[sb] : 0;
        const int Lmax = lenA > lenB ? lenA : lenB;
        const int Lother = shfl_xor(Lmax, 16);
        const int nsteps = (Lmax > Lother ? Lmax : Lother) + 2 * G - 1;   // 32 stages per group

        unsigned T[K], T2[K], E[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { T[k] = 0; T2[k] = 0; E[k] = 0; }
        unsigned Tdiag_in = 0;                 // T above-left of the upper half-column's first row
        unsigned Tout = 0, Fout = 0;           // bottom row of the lower half-column, previous step
        unsigned Tmid = 0, Fmid = 0;           // bottom row of the upper half-column, previous step (= above the lower one now)
        unsigned Tmid_d = 0;                   // ... two steps ago (= above-left of the lower one now)
        unsigned wprev = 0;                    // profile offsets of the column the upper half filled in the previous step
        unsigned thr = 0;
        unsigned best = 0, bestj = 0, bestpart = 0;   // per half: score, column, 0xffff if the best lies in the lower half-column
        sync_warp();
        if (lg == 0) *gpub = 0;

        for (int s0 = 0; s0 < nsteps; s0 += 32) {
            sync_warp();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int c = s0 + lg + u * G;
                const unsigned ca = sw16_fetch_code(p.words, wA, lenA, c, p.bits, pad_code);
                const unsigned cb = sw16_fetch_code(p.words, wB, lenB, c, p.bits, pad_code);
                ring[c & 63] = (ca * LSTRIDE) | ((cb * LSTRIDE) << 16);
            }
            sync_warp();
            auto step = [&](const int s, unsigned (&Tin)[K], unsigned (&Tnew)[K]) {
                const int jA = s - 2 * lg;         // column of the upper half-column
                const int jB = jA - 1;             // column of the lower half-column
                unsigned Tup = shfl_up(Tout, 1);
                unsigned Fup = shfl_up(Fout, 1);
                if (lg == 0) { Tup = 0; Fup = 0; }
                unsigned cmaxL = 0, cmaxU = 0;
                // ---- lower half-column first: it needs the upper one's previous results -------------
                if (jB >= 0 && jB < Lmax) {
                    const unsigned char *pa = lane_base + (wprev & 0xffffu);
                    const unsigned char *pb = lane_base + (wprev >> 16);
                    unsigned was[(CH - CHU) * 4], wbs[(CH - CHU) * 4];
#pragma unroll
                    for (int c = CHU; c < CH; ++c) {
                        const uint4 a = *(const uint4 *)(pa + c * (G * 16));
                        const uint4 b = *(const uint4 *)(pb + c * (G * 16));
                        was[4 * (c - CHU)] = a.x; was[4 * (c - CHU) + 1] = a.y; was[4 * (c - CHU) + 2] = a.z; was[4 * (c - CHU) + 3] = a.w;
                        wbs[4 * (c - CHU)] = b.x; wbs[4 * (c - CHU) + 1] = b.y; wbs[4 * (c - CHU) + 2] = b.z; wbs[4 * (c - CHU) + 3] = b.w;
                    }
                    unsigned Td = Tmid_d, Tu = Tmid, Fu = Fmid;
                    unsigned tprev = 0;
#pragma unroll
                    for (int k = KU; k < K; ++k) {
                        const unsigned So = wbs[k - KU] * m64k + was[k - KU];
                        const unsigned Tl = Tin[k];
                        const unsigned En = viaddmax2(E[k], NEGE, Tl);
                        const unsigned Fn = viaddmax2(Fu, NEGE, Tu);
                        const unsigned h = viaddmax2(Td, So, En);
                        const unsigned X = vimax3_2(h, Fn, O2);
                        const unsigned Tn = X * one + NEGO;
                        if ((k - KU) & 1) cmaxL = vimax3_2(cmaxL, tprev, Tn);
                        else tprev = Tn;
                        Td = Tl; Tnew[k] = Tn; E[k] = En; Tu = Tn; Fu = Fn;
                    }
                    if ((K - KU) & 1) cmaxL = vimax2(cmaxL, tprev);
                    Tout = Tu; Fout = Fu;
                }
                // ---- upper half-column -------------------------------------------------------------
                if (jA >= 0 && jA < Lmax) {
                    const unsigned w = ring[jA & 63];
                    const unsigned char *pa = lane_base + (w & 0xffffu);
                    const unsigned char *pb = lane_base + (w >> 16);
                    unsigned was[CHU * 4], wbs[CHU * 4];
#pragma unroll
                    for (int c = 0; c < CHU; ++c) {
                        const uint4 a = *(const uint4 *)(pa + c * (G * 16));
                        const uint4 b = *(const uint4 *)(pb + c * (G * 16));
                        was[4 * c] = a.x; was[4 * c + 1] = a.y; was[4 * c + 2] = a.z; was[4 * c + 3] = a.w;
                        wbs[4 * c] = b.x; wbs[4 * c + 1] = b.y; wbs[4 * c + 2] = b.z; wbs[4 * c + 3] = b.w;
                    }
                    unsigned Td = Tdiag_in, Tu = Tup, Fu = Fup;
                    unsigned tprev = 0;
#pragma unroll
                    for (int k = 0; k < KU; ++k) {
                        const unsigned So = wbs[k] * m64k + was[k];
                        const unsigned Tl = Tin[k];
                        const unsigned En = viaddmax2(E[k], NEGE, Tl);
                        const unsigned Fn = viaddmax2(Fu, NEGE, Tu);
                        const unsigned h = viaddmax2(Td, So, En);
                        const unsigned X = vimax3_2(h, Fn, O2);
                        const unsigned Tn = X * one + NEGO;
                        if (k & 1) cmaxU = vimax3_2(cmaxU, tprev, Tn);
                        else tprev = Tn;
                        Td = Tl; Tnew[k] = Tn; E[k] = En; Tu = Tn; Fu = Fn;
                    }
                    if (KU & 1) cmaxU = vimax2(cmaxU, tprev);
                    Tdiag_in = Tup;
                    Tmid_d = Tmid; Tmid = Tu; Fmid = Fu;
                    wprev = w;
                }
                thr = vimax2(thr, *gpub);
                const unsigned m = vimax2(thr, vimax2(cmaxL, cmaxU));
                if (m != thr) {
                    // cold: a half beat the threshold in one of the half-columns.  The lower one holds the
                    // smaller column, so it is looked at first; the upper one then needs a strictly larger score.
#pragma unroll
                    for (int part = 1; part >= 0; --part) {
                        const unsigned cm = part ? cmaxL : cmaxU;
                        const unsigned mp = vimax2(thr, cm);
                        if (mp == thr) continue;
                        const int jp = part ? jB : jA;
                        const unsigned diff = mp ^ thr;
                        const unsigned mask = ((diff & 0xffffu) ? 0xffffu : 0u) | ((diff >> 16) ? 0xffff0000u : 0u);
                        best = (best & ~mask) | (cm & mask);
                        bestj = (bestj & ~mask) | (((unsigned)jp * 0x10001u) & mask);
                        bestpart = (bestpart & ~mask) | (part ? mask : 0u);
                        thr = mp;
                        const int c4lo = part ? CHU : 0, c4hi = part ? CH : CHU;
#pragma unroll
                        for (int c4 = 0; c4 < CH; ++c4) {
                            if (c4 < c4lo || c4 >= c4hi) continue;
                            uint4 v;
                            v.x = Tnew[4 * c4];
                            v.y = 4 * c4 + 1 < K ? Tnew[4 * c4 + 1] : 0u;
                            v.z = 4 * c4 + 2 < K ? Tnew[4 * c4 + 2] : 0u;
                            v.w = 4 * c4 + 3 < K ? Tnew[4 * c4 + 3] : 0u;
                            if (mask & 0xffffu) park[(0 * (PARKW / 4) + c4) * 32 + lane] = v;
                            if (mask >> 16) park[(1 * (PARKW / 4) + c4) * 32 + lane] = v;
                        }
                    }
                    const unsigned pub = vimax2(*gpub, vimax2(best, 0x00010001u) - 0x00010001u);
                    *gpub = pub;
                }
            };
            const int send = (s0 + 32 < nsteps) ? s0 + 32 : nsteps;
            // even steps read T and write T2, odd steps the reverse; a trailing odd step past the end has no active lane
            for (int s = s0; s < send; s += 2) {
                step(s, T, T2);
                step(s + 1, T2, T);
            }
        }
        // ---- merge the group's lanes: (score desc, end_ref asc, end_query asc), each half separately --
        sync_warp();
        unsigned long long comps[2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const unsigned sc = half ? (best >> 16) : (best & 0xffffu);
            const unsigned col = half ? (bestj >> 16) : (bestj & 0xffffu);
            const bool lower = (half ? (bestpart >> 16) : (bestpart & 0xffffu)) != 0;
            // first row of the parked half-column that holds the lane's best
            unsigned rowk = 0;
            if (sc != 0) {
                const unsigned *pk = (const unsigned *)park;
                const int k0 = lower ? KU : 0, k1 = lower ? K : KU;
                for (int k = K - 1; k >= 0; --k) {
                    if (k < k0 || k >= k1) continue;
                    const unsigned wv = pk[(((half * (PARKW / 4) + (k >> 2)) * 32 + lane) << 2) + (k & 3)];
                    if ((half ? (wv >> 16) : (wv & 0xffffu)) == sc) rowk = (unsigned)k;
                }
            }
            const unsigned row = (unsigned)(lg * K) + rowk;
            unsigned long long comp = ((unsigned long long)sc << 26) | ((unsigned long long)(0xffffu - col) << 10) |
                                      (unsigned long long)(1023u - row);
#pragma unroll
            for (int m = G / 2; m >= 1; m >>= 1) {
                const unsigned long long other = (unsigned long long)shfl_xor((long long)comp, m);
                comp = other > comp ? other : comp;
            }
            comps[half] = comp;
        }
        if (lg == 0) {
            const bool overflow = (int)(comps[0] >> 26) >= limit || (int)(comps[1] >> 26) >= limit;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const long long sid = half ? sb : sa;
                if (sid >= p.n) continue;
                const unsigned long long comp = comps[half];
                const int sc = (int)(comp >> 26);
                const int oid = p.out_map ? p.out_map[sid] : (int)sid;
                if (overflow) {
                    const int slot_r = atomic_add(p.retry_count, 1);
                    p.retry[slot_r] = (int)sid + p.sid_base;
                } else if (sc == 0) {
                    p.score[oid] = 0; p.end_query[oid] = 0; p.end_ref[oid] = 0;
                } else {
                    p.score[oid] = sc;
                    p.end_ref[oid] = (int)(0xffffu - (unsigned)((comp >> 10) & 0xffffu));
                    p.end_query[oid] = 1023 - (int)(comp & 1023u);
                }
            }
        }
    }
}

}  // namespace psb

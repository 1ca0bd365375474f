// kern_pairs16.cuh -- many independent pairs in packed 16-bit lanes (DPX VIADDMNMX.S16x2 and friends):
// every 32-bit register holds the same cell of TWO different pairs, a group of G lanes is a G-stage
// systolic array over the query rows, and a warp runs 32/G groups -- 4 pairs per warp for G = 16.
//
// Replaces upstream parasail_{nw,sg*,sw}[_trace|_stats]_{striped,scan,diag}_{16,sat} for the many-pairs
// entry point psb_align_pairs (SURVEY configs C1, C3, C4) [REF src/aligner/mod.rs:411-429 is the per-pair
// call it batches]; results follow SURVEY.md Appendix A through the rule constants of psb_defs.h.
//
// Design notes (nothing here mirrors parasail's striped CPU layout):
//   * true (unshifted) score space with the vertical gap carried as Fh = F + open, so that the only
//     loop-carried dependency between the K rows of a lane is ONE instruction (needs open >= extend):
//        E  = max(E - e, Tleft)           Tleft = H[i][j-1] - o
//        h  = max(Tdiag + (S+o), E)       = max(H[i-1][j-1] + S, E)
//        H  = max(Fh - o, h [, 0])        VIADDMNMX.S16x2[.RELU]
//        Fh'= max(Fh - e, h)              equals max(Fh - e, H) wherever it can matter
//        T  = H - o                       VIADD.16x2
//   * every pair has its own query, so each group builds two int8 query profiles (one per half) of
//     (S + open) in shared memory, [letter][lane][K rows]; a step fetches the two letters' slots and
//     PRMT (sign-replicating selectors) interleaves them into one s16x2 score word per row.
//   * the query is aligned to the BOTTOM of the G*K rows: the padding rows sit above row 0 and are made
//     transparent (they reproduce the top boundary exactly: score -128 with a zero left boundary for
//     nw-style tops, score 0 for free tops), so the last query row is always row K-1 of lane G-1 and the
//     last row / corner candidates of nw and sg come out of a register that is live anyway.
//   * 16-bit safety is decided on the host by a static bound on the pair's lengths (pairs16_fits);
//     anything that does not fit goes to the 32-bit kernel, so there is no saturation to detect here.
//   * TRACE: the low byte of H of every cell -- nothing else.  Neighbouring cells of a Gotoh table differ by
//     less than 2*(max|S| + open) <= 127 (pairs16_trace_ok), so a walk that knows H exactly at one cell
//     recovers it exactly at any neighbour from that byte, and every decision of the traceback follows from
//     exact H values: the diagonal was taken iff H[i-1][j-1] + S == H[i][j] (it wins ties); otherwise a
//     vertical gap of length k iff H[i-k][j] - o - (k-1)e == H[i][j] for some k (F wins ties over E), with the
//     LARGEST such k, which is what "extend on ties" means for a flag-driven walk; otherwise the same test
//     along the row.  The fill therefore pays ONE PRMT per two rows for the trace (the low bytes of two rows
//     x two pairs make one word) instead of a dozen instructions of per-cell decision bits, and stores
//     2*ceil(K/2) bytes per lane and step, one coalesced run per group and step (1 byte per cell).
//     walk16_kernel turns the bytes into the CIGAR run list or into (matches, similar, length) -- the
//     statistics are those of the traceback path, which is how `_stats` results are produced for batches
//     (same source priorities).
#pragma once
#include "psb_defs.h"
#include "psb_simt.h"

namespace psb {

struct Pairs16Params {
    const uint8_t *q;           // residues mapped to matrix indices
    const long long *q_off;     // n+1 (shared_query: [0],[1])
    const uint8_t *r;
    const long long *r_off;
    int shared_query;
    const int *items;           // 2 pair ids per work item; the second is -1 when the item holds one pair
    int nitems;
    const int8_t *mat8;         // [33][32] int8, mat8[a*32 + x] = S(x, a) + open for query letter x and reference
                                // letter a; column 31 = the padding row above the query, row `size` = the pad letter
    int size;
    int open, gap;
    int mode, s1_beg, s1_end, s2_beg, s2_end;
    int *score, *end_query, *end_ref;     // indexed by pair id
    unsigned *trace;            // TRACE: low bytes of H, [step][lane][row pair][pair A, pair B]
    const long long *trace_off; // first word of each item's block
    int *counter;               // dynamic work queue over warp slots
};

inline constexpr int pairs16_slot(int K) { return ((K + 3) / 4) * 4; }        // profile bytes per lane and letter
inline constexpr int pairs16_trace_words(int K) { return (((K + 1) / 2) + 1) & ~1; } // H-byte words per lane and step (even)
inline size_t pairs16_warp_smem(int K, int nletters, bool sw) {
    const size_t prof = (size_t)64 * nletters * pairs16_slot(K);   // (32/G groups) x 2 halves x letters x G lanes x slot
    const size_t park = sw ? (size_t)2 * 32 * pairs16_slot(K) * 4 : 0;
    return 2 * 64 * 4 + 16 + park + prof;
}
inline size_t pairs16_smem_bytes(int K, int nletters, bool sw, int warps) {
    return 33 * 32 + (size_t)warps * pairs16_warp_smem(K, nletters, sw);
}
// words of decision bits of one item
inline long long pairs16_item_trace_words(int G, int K, int lr_max) {
    return (((long long)(lr_max + G - 1) * G * pairs16_trace_words(K)) + 3) / 4 * 4;
}
// static 16-bit bound: every intermediate of the fill stays inside int16 for this pair in a G*K-row frame
// (global / semi-global values are kept biased by +16384: half the range on either side)
inline bool pairs16_fits(int rows, int lq, int lr, int smax, int smin, int open, int gap, bool sw) {
    const long long pos = (long long)(lq < lr ? lq : lr) * (smax > 0 ? smax : 0) + 2ll * open + 256;
    const long long neg = 3ll * open + (long long)(rows + lr + 4) * gap + 256 + (smin < 0 ? -smin : 0);
    const long long lim = sw ? 32000 : 16000;
    return pos < lim && neg < lim && lr <= 65000;   // end columns travel as 16-bit halves
}
// trace / stats: neighbouring cells must differ by less than 128 so that one byte of H per cell is enough
inline bool pairs16_trace_ok(int mat_min, int mat_max, int open) {
    const int m = (mat_max > -mat_min ? mat_max : -mat_min);
    return 2 * (m + open) <= 127;
}
// the kernel's preconditions on the scoring scheme (the int8 profile and the one-instruction F chain)
inline bool pairs16_scheme_ok(int size, int mat_min, int mat_max, int open, int gap, bool pssm) {
    return !pssm && size <= 30 && gap >= 0 && open >= gap && mat_max + open <= 127 && mat_min + open >= -127 && open <= 127;
}
// host-side build of the [33][32] score table; pad_top = score byte of the padding rows above the query
inline void pairs16_build_mat8(const int *table, int size, int open, bool top_free, int8_t *out) {
    for (int x = 0; x < 33 * 32; ++x) out[x] = (int8_t)-128;
    for (int a = 0; a < size; ++a) {
        for (int x = 0; x < size; ++x) out[a * 32 + x] = (int8_t)(table[(size_t)x * size + a] + open);
        out[a * 32 + 31] = (int8_t)(top_free ? open : -128);
    }
}

PSB_DEV unsigned p16_pack(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
PSB_DEV int p16_lo(unsigned w) { return (int)(short)(w & 0xffffu); }
PSB_DEV int p16_hi(unsigned w) { return (int)(short)(w >> 16); }

template <int G, int K, bool SW, bool TRACE>
PSB_KERNEL void pairs16_kernel(Pairs16Params p) {
    constexpr int NG = 32 / G;
    constexpr int SLOT = ((K + 3) / 4) * 4;
    constexpr int NW4 = SLOT / 4;
    constexpr unsigned LSTRIDE = (unsigned)G * SLOT;
    constexpr int TW = (((K + 1) / 2) + 1) & ~1;   // = pairs16_trace_words(K)
    constexpr int ROWS = G * K;
    PSB_SHARED_DECL(smem_raw);
    const int lane = lane_id();
    const int lg = lane & (G - 1);
    const int grp = lane / G;
    const int NL = p.size + 1;
    const int pad_code = p.size;
    // ---- shared memory: the byte score table, then per warp: rings, published bests, parked columns, profiles
    for (int x = thread_in_block(); x < 33 * 32 / 4; x += threads_per_block()) ((unsigned *)smem_raw)[x] = ((const unsigned *)p.mat8)[x];
    sync_block();
    const unsigned char *mot = smem_raw;
    const size_t park_bytes = SW ? (size_t)2 * 32 * SLOT * 4 : 0;
    const size_t warp_bytes = 2 * 64 * 4 + 16 + park_bytes + (size_t)64 * NL * SLOT;
    unsigned char *wsm = smem_raw + 33 * 32 + (size_t)warp_in_block() * warp_bytes;
    unsigned *ring = (unsigned *)wsm + (NG == 2 ? grp * 64 : 0);
    volatile unsigned *gpub = (volatile unsigned *)(wsm + 2 * 64 * 4) + grp;
    uint4 *park = (uint4 *)(wsm + 2 * 64 * 4 + 16);
    unsigned char *prof = wsm + 2 * 64 * 4 + 16 + park_bytes + (size_t)grp * 2 * NL * LSTRIDE;   // this group's two profiles
    const unsigned char *pa_base = prof + lg * SLOT;
    const unsigned char *pb_base = prof + (size_t)NL * LSTRIDE + lg * SLOT;

    const int o = p.open, e = p.gap;
    const unsigned NEGE = p16_pack(-e, -e), NEGO = p16_pack(-o, -o);
    // global / semi-global kernels keep every value biased by +16384 in both halves: max and add are
    // translation invariant, the trace byte is untouched (the bias is a multiple of 256), and a biased half
    // never drops below o, so T = H - o is a plain 32-bit subtraction that cannot borrow across the halves
    // -- it leaves the DPX/ALU pipe, which is what bounds these kernels.  (Local alignment keeps the floor at
    // 0 for VIADDMNMX.RELU and pays the packed add.)
    constexpr int BIAS = SW ? 0 : 16384;
    const unsigned O32 = (unsigned)o * 0x10001u, E32 = (unsigned)e * 0x10001u;
    const int mode = p.mode;
    const bool top_free = SW || (mode == MODE_SG && p.s1_beg);
    const bool left_free = SW || (mode == MODE_SG && p.s2_beg);
    const bool row_ends = !SW && mode == MODE_SG && p.s1_end;
    const bool col_ends = !SW && mode == MODE_SG && p.s2_end;
    const int nslots = (p.nitems + NG - 1) / NG;

    for (;;) {
        int slot = 0;
        if (lane == 0) slot = atomic_add(p.counter, 1);
        slot = shfl(slot, 0);
        if (slot >= nslots) break;
        int item = slot * NG + grp;
        const bool item_ok = item < p.nitems;
        if (!item_ok) item = p.nitems - 1;
        const int pidA = p.items[2 * item];
        int pidB = p.items[2 * item + 1];
        const bool hasB = pidB >= 0;
        if (!hasB) pidB = pidA;
        const long long qoA = p.shared_query ? p.q_off[0] : p.q_off[pidA], qoB = p.shared_query ? p.q_off[0] : p.q_off[pidB];
        const int LqA = (int)((p.shared_query ? p.q_off[1] : p.q_off[pidA + 1]) - qoA);
        const int LqB = (int)((p.shared_query ? p.q_off[1] : p.q_off[pidB + 1]) - qoB);
        const long long roA = p.r_off[pidA], roB = p.r_off[pidB];
        const int LrA = (int)(p.r_off[pidA + 1] - roA), LrB = (int)(p.r_off[pidB + 1] - roB);
        const int padA = ROWS - LqA, padB = ROWS - LqB;     // padding rows above each query
        const int Lmax = LrA > LrB ? LrA : LrB;
        int nsteps = Lmax + G - 1;
        if (NG == 2) { const int other = shfl_xor(nsteps, 16); nsteps = nsteps > other ? nsteps : other; }

        // ---- the two query profiles of this group: lane t writes its own K rows for every letter ----------
        sync_warp();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint8_t *qs = p.q + (half ? qoB : qoA);
            const int pad = half ? padB : padA;
            int rowq[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int i = lg * K + k - pad;
                rowq[k] = i >= 0 ? (int)qs[i] : 31;
            }
            unsigned char *dst = prof + (size_t)half * NL * LSTRIDE + lg * SLOT;
            for (int a = 0; a < NL; ++a) {
                const unsigned char *mrow = mot + a * 32;
#pragma unroll
                for (int c = 0; c < NW4; ++c) {
                    unsigned w = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (4 * c + b < K) w |= (unsigned)mrow[rowq[4 * c + b]] << (8 * b);
                    *(unsigned *)(dst + (size_t)a * LSTRIDE + 4 * c) = w;
                }
            }
        }

        // ---- boundaries ------------------------------------------------------------------------------------
        unsigned T[K], E[K], T2[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int il = lg * K + k;
            const int ia = il - padA, ib = il - padB;
            const int la = (ia < 0 || left_free) ? 0 : -o - ia * e;
            const int lb = (ib < 0 || left_free) ? 0 : -o - ib * e;
            T[k] = p16_pack(la - o + BIAS, lb - o + BIAS);
            E[k] = T[k];
            T2[k] = T[k];
        }
        unsigned Tdiag_in;
        {
            const int il = lg * K - 1;   // the row above this lane's first row (lane 0: the corner)
            const int ia = il - padA, ib = il - padB;
            const int la = (il < 0 || ia < 0 || left_free) ? 0 : -o - ia * e;
            const int lb = (il < 0 || ib < 0 || left_free) ? 0 : -o - ib * e;
            Tdiag_in = p16_pack(la - o + BIAS, lb - o + BIAS);
        }
        unsigned Tout = Tdiag_in, Fout = 0;
        // the table's top boundary as lane 0 sees it at step s (its column j = s): T = H[-1][j] - o and
        // Fh = F[0][j] + o = H[-1][j], advanced by one column per step by every lane alike (no divergence)
        unsigned topT = p16_pack(-(top_free ? 0 : o) - o + BIAS, -(top_free ? 0 : o) - o + BIAS);
        unsigned topF = p16_pack(-(top_free ? 0 : o) + BIAS, -(top_free ? 0 : o) + BIAS);
        const unsigned topStep = top_free ? 0u : E32;
        // local: per half, a column maximum must exceed `thr` to matter (see kern_sw16.cuh)
        unsigned thr = 0, best = 0, bestj = 0;
        // global / semi-global candidates, kept in T space (T = H - o)
        unsigned rbest = 0x80008000u, rbestj = 0;      // last row, first maximum (lane G-1 only)
        int colTa = -32768, colTb = -32768, colIa = 0, colIb = 0;   // last column
        int cornTa = -32768, cornTb = -32768;          // (Lq-1, Lr-1), lane G-1 only
        sync_warp();
        if (SW && lg == 0) *gpub = 0;
        const long long tr_base = TRACE ? p.trace_off[item] : 0;

        for (int s0 = 0; s0 < nsteps; s0 += 32) {
            sync_warp();
            // stage the next 32 columns of both references as profile byte offsets
#pragma unroll
            for (int u = 0; u < NG; ++u) {
                const int c = s0 + lg + u * G;
                const unsigned ca = c < LrA ? (unsigned)p.r[roA + c] : (unsigned)pad_code;
                const unsigned cb = c < LrB ? (unsigned)p.r[roB + c] : (unsigned)pad_code;
                ring[c & 63] = (ca * LSTRIDE) | ((cb * LSTRIDE) << 16);
            }
            sync_warp();
            const int send = (s0 + 32 < nsteps) ? s0 + 32 : nsteps;
            auto step = [&](const int s, unsigned (&Tin)[K], unsigned (&Tnew)[K]) {
                const int j = s - lg;
                unsigned Tup = shfl_up(Tout, 1);
                unsigned Fup = shfl_up(Fout, 1);
                if (lg == 0) { Tup = topT; Fup = topF; }
                if (!SW) { topT -= topStep; topF -= topStep; }   // biased halves: a 32-bit subtraction cannot borrow
                if (j >= 0 && j < Lmax) {
                    const unsigned w = ring[j & 63];
                    const unsigned char *pa = pa_base + (w & 0xffffu);
                    const unsigned char *pb = pb_base + (w >> 16);
                    unsigned wa[NW4], wb[NW4];
#pragma unroll
                    for (int c = 0; c < NW4; ++c) { wa[c] = *(const unsigned *)(pa + 4 * c); wb[c] = *(const unsigned *)(pb + 4 * c); }
                    unsigned Td = Tdiag_in, Fu = Fup;
                    unsigned cmax = 0, hprev = 0;
                    unsigned hb8[TW];   // TRACE: low bytes of H, two rows x two pairs per word
                    unsigned heven = 0;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const unsigned So = prmt(wa[k >> 2], wb[k >> 2], 0xC480u + (unsigned)(k & 3) * 0x1111u);
                        const unsigned Tl = Tin[k];
                        const unsigned En = viaddmax2(E[k], NEGE, Tl);
                        const unsigned h = viaddmax2(Td, So, En);
                        const unsigned Hn = SW ? viaddmax2_relu(Fu, NEGO, h) : viaddmax2(Fu, NEGO, h);
                        const unsigned Fn = viaddmax2(Fu, NEGE, h);
                        const unsigned Tn = SW ? vadd2(Hn, NEGO) : Hn - O32;
                        if (TRACE) {
                            if (k & 1) hb8[k >> 1] = prmt(heven, Hn, 0x6420u);
                            else heven = Hn;
                        }
                        if (SW) {
                            if (k & 1) cmax = vimax3_2(cmax, hprev, Hn);
                            else hprev = Hn;
                        }
                        Td = Tl; Tnew[k] = Tn; E[k] = En; Fu = Fn;
                    }
                    if (TRACE) {
                        if (K & 1) hb8[K >> 1] = prmt(heven, 0u, 0x6420u);
#pragma unroll
                        for (int x = (K + 1) / 2; x < TW; ++x) hb8[x] = 0;
                    }
                    if (SW && (K & 1)) cmax = vimax2(cmax, hprev);
                    Tdiag_in = Tup; Tout = Tnew[K - 1]; Fout = Fu;
                    if (TRACE) {
                        unsigned *dst = p.trace + tr_base + ((long long)s * G + lg) * TW;
                        if (TW % 4 == 0) {
#pragma unroll
                            for (int x = 0; x < TW / 4; ++x) {
                                uint4 v; v.x = hb8[4 * x]; v.y = hb8[4 * x + 1]; v.z = hb8[4 * x + 2]; v.w = hb8[4 * x + 3];
                                st_cs((uint4 *)dst + x, v);
                            }
                        } else {
#pragma unroll
                            for (int x = 0; x < TW / 2; ++x) {
                                uint2 v; v.x = hb8[2 * x]; v.y = hb8[2 * x + 1];
                                st_cs((uint2 *)dst + x, v);
                            }
                        }
                    }
                    if (SW) {
                        thr = vimax2(thr, *gpub);
                        const unsigned m = vimax2(thr, cmax);
                        if (m != thr) {
                            // cold: a half beat the threshold -- record score and column, park the column of T values
                            const unsigned diff = m ^ thr;
                            const unsigned mask = ((diff & 0xffffu) ? 0xffffu : 0u) | ((diff >> 16) ? 0xffff0000u : 0u);
                            best = (best & ~mask) | (cmax & mask);
                            bestj = (bestj & ~mask) | (((unsigned)j * 0x10001u) & mask);
                            thr = m;
#pragma unroll
                            for (int c4 = 0; c4 < NW4; ++c4) {
                                uint4 v;
                                v.x = Tnew[4 * c4];
                                v.y = 4 * c4 + 1 < K ? Tnew[4 * c4 + 1] : 0u;
                                v.z = 4 * c4 + 2 < K ? Tnew[4 * c4 + 2] : 0u;
                                v.w = 4 * c4 + 3 < K ? Tnew[4 * c4 + 3] : 0u;
                                if (mask & 0xffffu) park[(0 * NW4 + c4) * 32 + lane] = v;
                                if (mask >> 16) park[(1 * NW4 + c4) * 32 + lane] = v;
                            }
                            const unsigned pub = vimax2(*gpub, vimax2(best, 0x00010001u) - 0x00010001u);
                            *gpub = pub;
                        }
                    } else {
                        if (row_ends && lg == G - 1) {
                            // last row, left to right, the first maximum wins (strict >), real columns only
                            const unsigned m = vimax2(rbest, Tout);
                            if (m != rbest) {
                                const unsigned diff = m ^ rbest;
                                unsigned mask = ((diff & 0xffffu) ? 0xffffu : 0u) | ((diff >> 16) ? 0xffff0000u : 0u);
                                mask &= (j < LrA ? 0xffffu : 0u) | (j < LrB ? 0xffff0000u : 0u);
                                rbest = (rbest & ~mask) | (Tout & mask);
                                rbestj = (rbestj & ~mask) | (((unsigned)j * 0x10001u) & mask);
                            }
                        }
                        if (j == LrA - 1 || j == LrB - 1) {
                            // cold: the last column of a half (top to bottom, strict >) and its corner cell
                            const bool atA = j == LrA - 1, atB = j == LrB - 1;
                            if (col_ends) {
#pragma unroll
                                for (int k = 0; k < K; ++k) {
                                    const int il = lg * K + k;
                                    const int ta = p16_lo(Tnew[k]), tb = p16_hi(Tnew[k]);
                                    if (atA && il >= padA && ta > colTa) { colTa = ta; colIa = il - padA; }
                                    if (atB && il >= padB && tb > colTb) { colTb = tb; colIb = il - padB; }
                                }
                            }
                            if (lg == G - 1) {
                                if (atA) cornTa = p16_lo(Tout);
                                if (atB) cornTb = p16_hi(Tout);
                            }
                        }
                    }
                }
            };
            // ping-pong column copies (even steps read T and write T2, odd steps the reverse); a chunk starts
            // on an even step and a trailing odd step past the end has no active lane
            for (int s = s0; s < send; s += 2) {
                step(s, T, T2);
                step(s + 1, T2, T);
            }
        }

        // ---- results ---------------------------------------------------------------------------------------
        sync_warp();
        if (SW) {
            unsigned long long comps[2];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const unsigned sc = half ? (best >> 16) : (best & 0xffffu);
                const unsigned col = half ? (bestj >> 16) : (bestj & 0xffffu);
                unsigned rowk = 0;
                if (sc != 0) {
                    const unsigned *pk = (const unsigned *)park;
                    const unsigned want = (sc - (unsigned)o) & 0xffffu;   // the parked words hold T = H - o
                    for (int k = K - 1; k >= 0; --k) {
                        const unsigned wv = pk[(((half * NW4 + (k >> 2)) * 32 + lane) << 2) + (k & 3)];
                        if ((half ? (wv >> 16) : (wv & 0xffffu)) == want) rowk = (unsigned)k;
                    }
                }
                const unsigned row = (unsigned)(lg * K) + rowk;
                // score | inverted column | inverted row: the maximum is (score, smaller end_ref, smaller end_query)
                unsigned long long comp = ((unsigned long long)sc << 26) | ((unsigned long long)(0xffffu - col) << 10) |
                                          (unsigned long long)(1023u - row);
#pragma unroll
                for (int m = G / 2; m >= 1; m >>= 1) {
                    const unsigned long long other = (unsigned long long)shfl_xor((long long)comp, m);
                    comp = other > comp ? other : comp;
                }
                comps[half] = comp;
            }
            if (lg == 0 && item_ok) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half && !hasB) continue;
                    const int pid = half ? pidB : pidA;
                    const unsigned long long comp = comps[half];
                    const int sc = (int)(comp >> 26);
                    if (sc == 0) { p.score[pid] = 0; p.end_query[pid] = 0; p.end_ref[pid] = 0; }
                    else {
                        p.score[pid] = sc;
                        p.end_ref[pid] = (int)(0xffffu - (unsigned)((comp >> 10) & 0xffffu));
                        p.end_query[pid] = 1023 - (int)(comp & 1023u) - (half ? padB : padA);
                    }
                }
            }
        } else {
            // last row / corner live in lane G-1; the last column is spread over the lanes
            const int src = grp * G + (G - 1);
            const unsigned rb = shfl(rbest, src), rbj = shfl(rbestj, src);
            const int cta = shfl(cornTa, src), ctb = shfl(cornTb, src);
            if (col_ends) {
#pragma unroll
                for (int m = G / 2; m >= 1; m >>= 1) {
                    const int ta = shfl_xor(colTa, m), ia = shfl_xor(colIa, m), tb = shfl_xor(colTb, m), ib = shfl_xor(colIb, m);
                    if (ta > colTa || (ta == colTa && ia < colIa)) { colTa = ta; colIa = ia; }
                    if (tb > colTb || (tb == colTb && ib < colIb)) { colTb = tb; colIb = ib; }
                }
            }
            if (lg == 0 && item_ok) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half && !hasB) continue;
                    const int pid = half ? pidB : pidA;
                    const int Lq = half ? LqB : LqA, Lr = half ? LrB : LrA;
                    int bT, bJ, bI = Lq - 1;
                    if (row_ends) { bT = half ? p16_hi(rb) : p16_lo(rb); bJ = (int)(half ? (rbj >> 16) : (rbj & 0xffffu)); }
                    else { bT = col_ends ? -32768 : (half ? ctb : cta); bJ = Lr - 1; }
                    const int cT = half ? colTb : colTa, cI = half ? colIb : colIa;
                    // the last column beats the last row only when strictly better (rules::SG_COL_WINS_TIE)
                    if (col_ends && (!row_ends || cT > bT)) { bT = cT; bJ = Lr - 1; bI = cI; }
                    p.score[pid] = bT + o - BIAS; p.end_query[pid] = bI; p.end_ref[pid] = bJ;
                }
            }
        }
    }
}

// ---- walk over the H bytes ---------------------------------------------------------------------
// One thread per pair.  walk16_kernel<false> writes the CIGAR run-length list in reverse into the pair's
// scratch region (compact_cigar_kernel reverses it into the CSR); walk16_kernel<true> counts (matches,
// similar, length) of the path.  Every decision is re-derived from exact H values (see the header).
struct Walk16Params {
    const uint8_t *q;
    const long long *q_off;
    const uint8_t *r;
    const long long *r_off;
    int shared_query;
    const int *pair_ids;         // pairs of this launch
    const int *pair_slot;        // per pair id: item * 2 + half
    int n;
    int G, K;
    const unsigned *trace;
    const long long *trace_off;  // per item
    const int *matrix;           // size x size substitution scores (no open added)
    int size;
    int open, gap;
    int is_sw, top_free, left_free;
    const int *score, *end_query, *end_ref;   // per pair id
    // CIGAR outputs (walk16_kernel<false>)
    unsigned *rev_ops;
    const long long *rev_off;
    int *nops, *beg_query, *beg_ref;
    // statistics outputs (walk16_kernel<true>)
    int *matches, *similar, *length;
};

// dynamic shared memory: size*size ints (the substitution matrix)
inline size_t walk16_smem_bytes(int size) { return (size_t)size * size * sizeof(int); }

template <bool STATS>
PSB_KERNEL void walk16_kernel(Walk16Params p) {
    PSB_SHARED_DECL(smem_raw);
    int *smat = (int *)smem_raw;
    for (int x = thread_in_block(); x < p.size * p.size; x += threads_per_block()) smat[x] = p.matrix[x];
    sync_block();
    const long long tid = (long long)block_id() * threads_per_block() + thread_in_block();
    if (tid >= p.n) return;
    const int pid = p.pair_ids[tid];
    const int slot = p.pair_slot[pid];
    const int item = slot >> 1, half = slot & 1;
    const long long qo = p.shared_query ? p.q_off[0] : p.q_off[pid];
    const int Lq = (int)((p.shared_query ? p.q_off[1] : p.q_off[pid + 1]) - qo);
    const uint8_t *q = p.q + qo;
    const uint8_t *r = p.r + p.r_off[pid];
    const int G = p.G, K = p.K, TW = (((K + 1) / 2) + 1) & ~1;   // = pairs16_trace_words(K)
    const int TWB = TW * 4;
    const int pad = G * K - Lq;
    const int o = p.open, e = p.gap;
    const uint8_t *tr = (const uint8_t *)(p.trace + p.trace_off[item]) + half;
    unsigned *out = STATS ? nullptr : p.rev_ops + p.rev_off[pid];
    // byte of cell (i, j); exact H from a byte and the exact value `ref` of a neighbouring cell
    // il / K by multiplication: exact for il < 1024 and every K <= 32 (checked by tests/test_emu_pairs16.py)
    const unsigned rcpK = 65536u / (unsigned)K + 1u;
    auto load = [&](int i, int j) -> unsigned {
        const int il = i + pad, t = (int)(((unsigned)il * rcpK) >> 16), k = il - t * K;
        return tr[((long long)(j + t) * G + t) * TWB + 2 * k];
    };
    auto recon = [](unsigned byte, int ref) -> int { return ref + (int)(signed char)(unsigned char)(byte - (unsigned)ref); };
    auto top = [&](int j) -> int { return (j < 0 || p.top_free) ? 0 : -o - j * e; };     // H[-1][j], corner 0
    auto left = [&](int i) -> int { return (i < 0 || p.left_free) ? 0 : -o - i * e; };   // H[i][-1]
    int i = p.end_query[pid], j = p.end_ref[pid];
    int v = p.score[pid];            // exact H[i][j]
    int cur = -1, n = 0;
    unsigned len = 0;
    int nm = 0, ns = 0, nl = 0;
    auto emit = [&](int op, int count) {
        if (STATS || count <= 0) return;
        if (op == cur) len += (unsigned)count;
        else {
            if (cur >= 0) out[n++] = (len << 4) | (unsigned)cur;
            cur = op; len = (unsigned)count;
        }
    };
    constexpr int D = 8;   // loads in flight per round trip
    bool done = false;
    while (!done && (i >= 0 || j >= 0)) {
        if (i < 0 || j < 0) {
            // off the table: statistics stop here (boundary gaps are not counted); the CIGAR takes the rest of
            // the other sequence as one run (rules::CIGAR_WALK_TO_ORIGIN)
            if (i < 0) { emit((int)rules::CIGAR_OP_D, j + 1); j = -1; }
            else { emit((int)rules::CIGAR_OP_I, i + 1); i = -1; }
            break;
        }
        // ---- a run of diagonal steps: the next D cells of the diagonal are fetched together --------------
        unsigned hb[D], qa[D], rb[D];
#pragma unroll
        for (int m = 0; m < D; ++m) {
            const bool in = i - m >= 0 && j - m >= 0;
            qa[m] = in ? q[i - m] : 0u;
            rb[m] = in ? r[j - m] : 0u;
            hb[m] = (i - m - 1 >= 0 && j - m - 1 >= 0) ? load(i - m - 1, j - m - 1) : 0u;
        }
        bool gap_here = false;
#pragma unroll
        for (int m = 0; m < D; ++m) {
            if (gap_here || done) break;
            if (i < 0 || j < 0) break;
            if (p.is_sw && v <= 0) { done = true; break; }   // ZERO: the alignment starts after this cell
            const int a = (int)qa[m], b = (int)rb[m];
            const int sub = smat[a * p.size + b];
            // the diagonal wins ties: taken iff H[i-1][j-1] + S == H[i][j]
            const int hd = i == 0 ? top(j - 1) : (j == 0 ? left(i - 1) : recon(hb[m], v - sub));
            if (hd + sub == v) {
                emit((a == b) ? (int)rules::CIGAR_OP_EQ : (int)rules::CIGAR_OP_X, 1);
                nm += (a == b); ns += (sub > 0); ++nl;
                v = hd; --i; --j;
            } else gap_here = true;
        }
        if (!gap_here || done) continue;
        // ---- vertical gap (F wins ties over E): the largest k with H[i-k][j] - o - (k-1)e == H[i][j] --------
        int kbest = 0, vbest = 0, u = v, k = 1;
        while (k <= i + 1) {
            unsigned bt[D];
#pragma unroll
            for (int m = 0; m < D; ++m) bt[m] = (i - k - m >= 0) ? load(i - k - m, j) : 0u;
#pragma unroll
            for (int m = 0; m < D; ++m) {
                if (k > i + 1) break;
                u = (i - k < 0) ? top(j) : recon(bt[m], u);
                if (u - o - (k - 1) * e == v) { kbest = k; vbest = u; }
                ++k;
            }
            if (rules::GAP_OPEN_ON_TIE && kbest) break;
        }
        if (kbest) {
            emit((int)rules::CIGAR_OP_I, kbest);
            nl += kbest; i -= kbest; v = vbest;
            continue;
        }
        u = v; k = 1;
        while (k <= j + 1) {
            unsigned bt[D];
#pragma unroll
            for (int m = 0; m < D; ++m) bt[m] = (j - k - m >= 0) ? load(i, j - k - m) : 0u;
#pragma unroll
            for (int m = 0; m < D; ++m) {
                if (k > j + 1) break;
                u = (j - k < 0) ? left(i) : recon(bt[m], u);
                if (u - o - (k - 1) * e == v) { kbest = k; vbest = u; }
                ++k;
            }
            if (rules::GAP_OPEN_ON_TIE && kbest) break;
        }
        if (!kbest) break;   // cannot happen for a table the fill produced
        emit((int)rules::CIGAR_OP_D, kbest);
        nl += kbest; j -= kbest; v = vbest;
    }
    if (STATS) {
        p.matches[pid] = nm; p.similar[pid] = ns; p.length[pid] = nl;
    } else {
        if (cur >= 0) out[n++] = (len << 4) | (unsigned)cur;
        p.nops[pid] = n;
        p.beg_query[pid] = i + 1;
        p.beg_ref[pid] = j + 1;
    }
}

}  // namespace psb

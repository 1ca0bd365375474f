// kern_sw16.cuh -- the database-scan hot kernel: Smith-Waterman score + end cell of ONE resident
// query profile against the resident bit-packed database, two subjects per warp in packed
// 16-bit lanes (DPX VIADDMNMX.S16x2 / VIMNMX3.S16x2).
//
// Replaces upstream parasail_sw_{striped,scan}_profile_{16,sat} as reached from parasail-rs's
// profile aligner [REF src/aligner/mod.rs:431-450]; results follow SURVEY.md Appendix A
// (score; smallest end_ref among maxima; smallest end_query in that column).
//
// Design (B200-first; nothing here mirrors parasail's striped CPU layout):
//   * a warp is a 32-stage systolic array: lane t owns K consecutive query rows in registers
//     and fills reference column j = s - t at step s; the bottom row travels down by shfl_up.
//   * each 32-bit register holds the same cell of TWO different subjects (lo/hi half), so one
//     DPX instruction updates two cells and there is no dependency inside a word.
//   * values are kept shifted by `open`:  X = H + o,  T = X - o = H,  E' = E + o,  F' = F + o.
//     In that space the local-alignment floor is the constant o, every value stays inside
//     [0, 32767] (no -inf sentinels), and T = X - (o,o) is a plain 32-bit subtract that can
//     never borrow across the halves, so it runs on the FMA pipe (IMAD) beside the DPX ops:
//        E' = max(E' - e, Tleft)      F' = max(F'up - e, Tup)       h = max(Tdiag + (S+o), E')
//        X  = max(h, F', o)           T  = X - o
//   * query profile: int8 (S+o), laid out [letter][lane][16 rows] so that one LDS.128 per
//     subject letter fetches all of a lane's rows and the 8 lanes of a quarter-warp always hit
//     8 different 16-byte bank groups (conflict free whatever the letters are); PRMT with
//     sign-replicating selectors interleaves the two subjects' scores into one 16x2 word.
//   * end cell: key = T*16 + (15 - k) (IMAD), column maximum by VIMNMX3, compared once per
//     column against the running best with the two-predicate VIMNMX so that only a strictly
//     larger score replaces it (first column wins; the row bits make the first row win).
//   * 16-bit overflow: scores that reach 2048 - max(S) are re-run by the 32-bit kernel.
#pragma once
#include "psb_defs.h"
#include "psb_simt.h"

#include <vector>

namespace psb {

static constexpr int SW16_SCORE_LIMIT = 2048;  // keys are H*16 + row bits in a signed 16-bit half

struct Sw16Profile {
    int8_t *prof = nullptr;  // device: [nletters][32 lanes][16] bytes of (S + open); pad letter last
    int lq = 0;
    int K = 0;               // rows per lane
    int nletters = 0;        // alphabet size + 1 (the pad letter)
    int max_score = 0, min_score = 0;
    int open_baked = -1;
};

static const int kSw16K[] = {2, 4, 6, 8, 10, 12, 13, 14, 16};
inline int sw16_pick_k(int lq) {
    for (int k : kSw16K) if (32 * k >= lq) return k;
    return 0;
}
inline bool sw16_supported(const Sw16Profile &p, int open, int gap) {
    return p.K > 0 && open >= 0 && gap >= 0 && gap <= 4096 && open + p.max_score <= 127 && open + p.min_score >= -127;
}
// host-side build of the packed profile with `open` folded in
inline bool sw16_build_profile(const uint8_t *mapped_query, int lq, const int *table, int size, int open,
                               Sw16Profile *out, std::vector<int8_t> *host) {
    const int K = sw16_pick_k(lq);
    if (K == 0 || size + 1 > 32) return false;
    out->lq = lq; out->K = K; out->nletters = size + 1; out->open_baked = open;
    int mx = -1000000, mn = 1000000;
    for (int i = 0; i < size * size; ++i) { mx = table[i] > mx ? table[i] : mx; mn = table[i] < mn ? table[i] : mn; }
    out->max_score = mx; out->min_score = mn;
    if (!sw16_supported(*out, open, 0)) return false;
    host->assign((size_t)(size + 1) * 32 * 16, (int8_t)-128);
    for (int a = 0; a < size; ++a)
        for (int lane = 0; lane < 32; ++lane)
            for (int k = 0; k < K; ++k) {
                const int i = lane * K + k;
                if (i < lq) (*host)[((size_t)a * 32 + lane) * 16 + k] = (int8_t)(table[(size_t)mapped_query[i] * size + a] + open);
            }
    return true;
}

struct Sw16Params {
    const int8_t *prof;
    int nletters;
    int lq;
    int open, gap;
    int max_score;
    const unsigned *words;        // packed database (sorted by length, descending)
    const long long *word_off;    // n+1
    const int *len;               // n
    int bits;                     // 5 or 2
    long long n;                  // subjects; work item w = subjects 2w and 2w+1
    const int *out_map;           // sorted position -> caller's subject id
    int *score, *end_query, *end_ref;
    int *retry;                   // sorted positions that overflowed 16 bit
    int *retry_count;
    int sid_base;                 // added to a subject index before it is pushed to `retry`
    int *counter;                 // dynamic work queue over items
    unsigned mul_one;             // the constants 1 and 16, kept opaque so that T = X*1 - o and
    unsigned mul_16;              // key = T*16 + c stay IMADs (FMA pipe) instead of ALU-pipe VIADD/LEA
};

inline size_t sw16_smem_bytes(int nletters, int warps) { return (size_t)nletters * 512 + (size_t)warps * 64 * 4; }

PSB_DEV unsigned sw16_fetch_code(const unsigned *words, long long w0, int len, int c, int bits, int pad_code) {
    if (c >= len) return (unsigned)pad_code;
    if (bits == 2) return (words[w0 + (c >> 4)] >> (2 * (c & 15))) & 3u;
    const int w = (int)(((unsigned long long)(unsigned)c * 0xAAAAAAABull) >> 34);  // c / 6
    return (words[w0 + w] >> (5 * (c - 6 * w))) & 31u;
}

template <int K>
PSB_KERNEL void sw16_scan_kernel(Sw16Params p) {
    PSB_SHARED_DECL(smem_raw);
    const int lane = lane_id();
    // ---- stage the profile (shared by all warps of the block) ----------------------------------
    {
        const uint4 *src = (const uint4 *)p.prof;
        uint4 *dst = (uint4 *)smem_raw;
        const int n16 = p.nletters * 32;
        for (int x = thread_in_block(); x < n16; x += threads_per_block()) dst[x] = src[x];
    }
    sync_block();
    const unsigned char *lane_base = smem_raw + lane * 16;
    unsigned *ring = (unsigned *)(smem_raw + (size_t)p.nletters * 512) + warp_in_block() * 64;
    const int pad_code = p.nletters - 1;
    const unsigned O2 = (unsigned)p.open * 0x10001u;
    const unsigned NEGE = ((unsigned)(-p.gap) & 0xffffu) * 0x10001u;
    const unsigned NEGO = 0u - O2;
    const unsigned one = p.mul_one, sixteen = p.mul_16;
    const long long nitems = (p.n + 1) >> 1;

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomic_add(p.counter, 1);
        item = shfl(item, 0);
        if (item >= nitems) break;
        const long long sa = 2ll * item, sb = sa + 1;
        const int lenA = p.len[sa];
        const int lenB = sb < p.n ? p.len[sb] : 0;
        const long long wA = p.word_off[sa], wB = sb < p.n ? p.word_off[sb] : 0;
        const int Lmax = lenA > lenB ? lenA : lenB;
        const int nsteps = Lmax + 31;

        unsigned T[K], E[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { T[k] = 0; E[k] = 0; }
        unsigned Tdiag_in = 0, Tout = 0, Fout = 0;
        unsigned bestcap = 0x000f000fu;   // (bestH*16 + 15) per half: only a larger H beats it
        unsigned bestkey = 0, bestj = 0;  // key and column (u16 per half) of the current best

        for (int s0 = 0; s0 < nsteps; s0 += 32) {
            sync_warp();
            {
                const int c = s0 + lane;
                const unsigned ca = sw16_fetch_code(p.words, wA, lenA, c, p.bits, pad_code);
                const unsigned cb = sw16_fetch_code(p.words, wB, lenB, c, p.bits, pad_code);
                ring[c & 63] = (ca * 512u) | ((cb * 512u) << 16);
            }
            sync_warp();
            const int send = (s0 + 32 < nsteps) ? s0 + 32 : nsteps;
            for (int s = s0; s < send; ++s) {
                const int j = s - lane;
                unsigned Tup = shfl_up(Tout, 1);
                unsigned Fup = shfl_up(Fout, 1);
                if (lane == 0) { Tup = 0; Fup = 0; }
                if (j >= 0 && j < Lmax) {
                    const unsigned w = ring[j & 63];
                    const uint4 wa = *(const uint4 *)(lane_base + (w & 0xffffu));
                    const uint4 wb = *(const uint4 *)(lane_base + (w >> 16));
                    const unsigned was[4] = {wa.x, wa.y, wa.z, wa.w};
                    const unsigned wbs[4] = {wb.x, wb.y, wb.z, wb.w};
                    unsigned Td = Tdiag_in, Tu = Tup, Fu = Fup;
                    unsigned cmax = 0;
                    unsigned keyprev = 0;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        constexpr unsigned SEL0 = 0xC480u;
                        const unsigned sel = SEL0 + (unsigned)(k & 3) * 0x1111u;
                        const unsigned So = prmt(was[k >> 2], wbs[k >> 2], sel);
                        const unsigned Tl = T[k];
                        const unsigned En = viaddmax2(E[k], NEGE, Tl);
                        const unsigned Fn = viaddmax2(Fu, NEGE, Tu);
                        const unsigned h = viaddmax2(Td, So, En);
                        const unsigned X = vimax3_2(h, Fn, O2);
                        const unsigned Tn = X * one + NEGO;
                        const unsigned key = Tn * sixteen + (unsigned)(15 - k) * 0x10001u;
                        if (k & 1) cmax = vimax3_2(cmax, keyprev, key);
                        else keyprev = key;
                        Td = Tl; T[k] = Tn; E[k] = En; Tu = Tn; Fu = Fn;
                    }
                    if (K & 1) cmax = vimax2(cmax, keyprev);
                    Tdiag_in = Tup; Tout = Tu; Fout = Fu;
                    // strictly larger score only: bestcap carries row bits 15, so an equal score
                    // in a later column can never exceed it.  Improvements are rare after the
                    // first columns, so the bookkeeping sits behind one compare.
                    const unsigned m = vimax2(bestcap, cmax);
                    if (m != bestcap) {
                        const unsigned diff = m ^ bestcap;
                        const unsigned jp = (unsigned)j * 0x10001u;
                        if (diff & 0xffffu) { bestkey = (bestkey & 0xffff0000u) | (cmax & 0xffffu); bestj = (bestj & 0xffff0000u) | (jp & 0xffffu); }
                        if (diff >> 16) { bestkey = (bestkey & 0x0000ffffu) | (cmax & 0xffff0000u); bestj = (bestj & 0x0000ffffu) | (jp & 0xffff0000u); }
                        bestcap = m | 0x000f000fu;
                    }
                }
            }
        }
        // ---- merge the lanes: (score desc, end_ref asc, end_query asc), each half separately ------
        unsigned long long comps[2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const unsigned key = half ? (bestkey >> 16) : (bestkey & 0xffffu);
            const unsigned col = half ? (bestj >> 16) : (bestj & 0xffffu);
            const int H = (int)(key >> 4);
            const int row = lane * K + 15 - (int)(key & 15u);
            // 11 bits score | 16 bits inverted column | 9 bits inverted row
            unsigned long long comp = ((unsigned long long)H << 25) | ((unsigned long long)(0xffffu - col) << 9) |
                                      (unsigned long long)(511 - row);
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const unsigned long long other = (unsigned long long)shfl_xor((long long)comp, m);
                comp = other > comp ? other : comp;
            }
            comps[half] = comp;
        }
        if (lane == 0) {
            // a half that leaves the 16-bit range can disturb its neighbour's keys: re-run both
            const int limit = SW16_SCORE_LIMIT - p.max_score;
            const bool overflow = (int)(comps[0] >> 25) >= limit || (int)(comps[1] >> 25) >= limit;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const long long sid = half ? sb : sa;
                if (sid >= p.n) continue;
                const unsigned long long comp = comps[half];
                const int sc = (int)(comp >> 25);
                const int oid = p.out_map ? p.out_map[sid] : (int)sid;
                if (overflow) {
                    const int slot = atomic_add(p.retry_count, 1);
                    p.retry[slot] = (int)sid + p.sid_base;
                } else if (sc == 0) {
                    p.score[oid] = 0; p.end_query[oid] = 0; p.end_ref[oid] = 0;
                } else {
                    p.score[oid] = sc;
                    p.end_ref[oid] = (int)(0xffffu - (unsigned)((comp >> 9) & 0xffffu));
                    p.end_query[oid] = 511 - (int)(comp & 511u);
                }
            }
        }
    }
}

}  // namespace psb

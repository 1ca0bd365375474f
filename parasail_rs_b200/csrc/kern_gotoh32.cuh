// kern_gotoh32.cuh -- the general 32-bit Gotoh H/E/F fill: one warp per pair, every mode
// (nw / sg with the four end-gap flags / sw), optional statistics, optional trace bytes.
//
// Replaces the upstream parasail kernels that parasail-rs reaches through the function
// pointer at [REF src/aligner/mod.rs:413-421, 435-441] (parasail_{nw,sg*,sw}[_stats|_trace]
// _{striped,scan,diag}[_profile]_{32,64,sat}); results follow SURVEY.md Appendix A.
//
// Layout of the computation (nothing here is derived from parasail's CPU kernels):
//   * rows = query, columns = reference.  The query is cut into strips of 32*K rows; inside
//     a strip lane t owns K consecutive rows and, at step s, fills column j = s - t, so the
//     warp is a 32-stage systolic array sweeping the reference left to right.
//   * per row a lane keeps T = H - open and E in registers; the substitution matrix sits in
//     shared memory with `open` pre-added, so a cell is exactly five integer operations:
//       E = max(E - e, Tleft)   F = max(Fup - e, Tup)   h = max(Tdiag + (S+o), E)
//       H = max(h, F [,0])      T = H - o                       (VIADDMNMX / VIMNMX3 / VIADD)
//   * the bottom row (T, F) of a lane travels to the lane below by __shfl_up; the bottom row
//     of a strip is parked in a per-warp global scratch line and fed to lane 0 of the next
//     strip; reference residues and that boundary line are staged in a 64-entry shared ring.
//   * end cell: sw keeps a per-lane (score, first column, first row) with the key
//     H*16 + (15 - k) built on the FMA pipe; sg/nw read the last row / last column; lanes
//     are merged once per pair with parasail's tie-break (score, then smaller end_ref,
//     then smaller end_query; the last column only beats the last row when strictly better).
#pragma once
#include "psb_defs.h"
#include "psb_simt.h"

namespace psb {

struct Gotoh32Params {
    const uint8_t *q;           // residues already mapped to matrix column indices
    const long long *q_off;     // n+1 offsets (shared_query: only [0],[1] are read)
    const uint8_t *r;
    const long long *r_off;     // n+1 offsets
    const int *order;           // pair ids handled by this launch (NULL: 0..n-1)
    int n;                      // number of entries of `order`
    int shared_query;
    const int *matrix;          // square: size*size; pssm: qlen*size (global memory)
    int size;
    int is_pssm;
    int open, gap;
    int mode, s1_beg, s1_end, s2_beg, s2_end;
    int *score, *end_query, *end_ref;   // indexed by pair id
    int *matches, *similar, *length;    // STATS only
    int *bnd;                   // strip-boundary scratch, bnd_stride ints per warp
    long long bnd_stride;
    uint8_t *trace;             // TRACE only: per pair [strip][step][lane][K] bytes
    const long long *trace_off; // byte offset of each pair's trace block (indexed by pair id)
    int *counter;               // dynamic work queue
    const int *out_map;         // results are written at out_map[pair id] (NULL: pair id)
    int *tabH, *tabM, *tabS, *tabL;  // TABLE only: per-cell planes, same indexing as the trace
    const long long *tab_off;   // element offset of each pair's table block (indexed by pair id)
    // database scans read their subjects straight from the bit-packed store (r / r_off unused):
    const unsigned *r_words;    // packed residues (kern_util.cuh layout), NULL for byte subjects
    const long long *r_word_off;  // first word of each subject (indexed by pair id)
    const int *r_len;           // residues of each subject
    int r_bits;                 // 5, 3 or 2
    const int *n_dev;           // when set, the number of work items is read from device memory
    // parasail_nw_banded (coarse family only): cell (i, j) is reachable iff band_lo <= j - i <= band_hi
    int banded, band_lo, band_hi;
};

// statistics word: matches | similar | length packed so that one add updates all three
template <typename W> struct StatPack;
template <> struct StatPack<unsigned> {  // 10 | 10 | 12 bits: min(Lq,Lr) < 1024, Lq+Lr < 4096
    static constexpr int SS = 10, LS = 20;
    static constexpr unsigned MM = 0x3ffu, SM = 0x3ffu, LM = 0xfffu;
};
template <> struct StatPack<unsigned long long> {  // 21 | 21 | 22 bits
    static constexpr int SS = 21, LS = 42;
    static constexpr unsigned long long MM = 0x1fffffull, SM = 0x1fffffull, LM = 0x3fffffull;
};

// prof: each warp additionally keeps an int8 query profile [letter][lane][16 rows] of (S + open)
inline size_t gotoh32_smem_bytes(int size, int warps, bool stats, int statw, bool prof = false) {
    size_t m = (size_t)size * (size_t)size * sizeof(int);
    size_t ring = 64 /*letters*/ + 64 * 2 * sizeof(int) + (stats ? 64 * 2 * (size_t)statw : 0);
    return ((m + 15) & ~(size_t)15) + (size_t)warps * (((ring + 15) & ~(size_t)15) + (prof ? (size_t)size * 512 : 0));
}
// the per-warp profile needs S + open to fit a signed byte
inline bool gotoh32_profile_ok(int size, int mat_min, int mat_max, int open, bool pssm) {
    return !pssm && size <= 32 && open >= 0 && mat_max + open <= 127 && mat_min + open >= -127;
}

// MODESEL: 0 = alignment mode read at run time, 1 = local only, 2 = global / semi-global only
// (the specialised forms drop the other mode's bookkeeping from the inner loop)
template <int K, bool STATS, bool TRACE, bool TABLE, typename SW_, bool PROF = false, int MODESEL = 0>
PSB_KERNEL void gotoh32_kernel(Gotoh32Params p) {
    typedef SW_ SWord;
    typedef StatPack<SWord> SP;
    PSB_SHARED_DECL(smem_raw);
    const int lane = lane_id();
    const int size = p.size;
    const int o = p.open, e = p.gap;
    const bool square_in_smem = !p.is_pssm;

    // ---- shared memory: matrix (+open) then per-warp rings --------------------------------
    int *smat = (int *)smem_raw;
    const size_t mat_bytes = square_in_smem ? ((((size_t)size * size * sizeof(int)) + 15) & ~(size_t)15) : 0;
    const size_t ring_bytes = ((64 + 64 * 2 * sizeof(int) + (STATS ? 64 * 2 * sizeof(SWord) : 0)) + 15) & ~(size_t)15;
    const size_t prof_bytes = PROF ? (size_t)size * 512 : 0;
    unsigned char *ring = smem_raw + mat_bytes + (size_t)warp_in_block() * (ring_bytes + prof_bytes);
    // PROF: int8 (S + open) of this lane's K rows for every reference letter, one 16-byte slot per
    // (letter, lane): one conflict-free LDS.128 per step replaces K scattered matrix reads
    unsigned char *wprof = ring + ring_bytes;
    int *ringT = (int *)ring;                 // boundary T (= H - o) of the strip above
    int *ringF = ringT + 64;                  // boundary F
    SWord *ringHs = (SWord *)(ringF + 64);    // boundary H stats
    SWord *ringFs = ringHs + (STATS ? 64 : 0);
    uint8_t *ringL = (uint8_t *)(ringFs + (STATS ? 64 : 0));  // reference residues
    if (square_in_smem) {
        for (int x = thread_in_block(); x < size * size; x += threads_per_block()) smat[x] = p.matrix[x] + o;
    }
    sync_block();

    const int mode = p.mode;
    const bool is_sw = MODESEL == 1 ? true : (MODESEL == 2 ? false : mode == MODE_SW);
    const bool top_free = is_sw || (mode == MODE_SG && p.s1_beg);   // H[-1][j] = 0
    const bool left_free = is_sw || (mode == MODE_SG && p.s2_beg);  // H[i][-1] = 0
    const bool row_ends = mode == MODE_SG && p.s1_end;               // last row holds candidates
    const bool col_ends = mode == MODE_SG && p.s2_end;               // last column holds candidates
    const int warp_global = block_id() * warps_per_block() + warp_in_block();
    int *bndT = p.bnd ? p.bnd + (long long)warp_global * p.bnd_stride : nullptr;
    const long long bnd_cols = p.bnd_stride / (2 + (STATS ? 2 * (long long)(sizeof(SWord) / sizeof(int)) : 0));
    int *bndF = bndT ? bndT + bnd_cols : nullptr;
    SWord *bndHs = bndT ? (SWord *)(bndF + bnd_cols) : nullptr;
    SWord *bndFs = bndT ? bndHs + (STATS ? bnd_cols : 0) : nullptr;

    for (;;) {
        int slot = 0;
        if (lane == 0) slot = atomic_add(p.counter, 1);
        slot = shfl(slot, 0);
        if (slot >= (p.n_dev ? ld_cg(p.n_dev) : p.n)) break;
        const int pid = p.order ? p.order[slot] : slot;
        const long long qo = p.shared_query ? p.q_off[0] : p.q_off[pid];
        const int Lq = (int)((p.shared_query ? p.q_off[1] : p.q_off[pid + 1]) - qo);
        const bool packed = p.r_words != nullptr;
        const long long ro = packed ? p.r_word_off[pid] : p.r_off[pid];
        const int Lr = packed ? p.r_len[pid] : (int)(p.r_off[pid + 1] - ro);
        const uint8_t *q = p.q + qo;
        const uint8_t *r = packed ? nullptr : p.r + ro;
        const int rows_per_strip = 32 * K;
        const int nstrips = (Lq + rows_per_strip - 1) / rows_per_strip;
        const int nsteps = Lr + 31;
        uint8_t *tr = TRACE ? p.trace + p.trace_off[pid] : nullptr;

        // per-lane best (sw: anywhere; sg: last row) and best of the last column
        int bestH = NEG_INF32, bestJ = 0x7fffffff, bestI = 0x7fffffff;
        int colH = NEG_INF32, colI = 0x7fffffff;
        SWord bestS = 0, colS = 0;

        for (int strip = 0; strip < nstrips; ++strip) {
            const int i0 = strip * rows_per_strip + lane * K;
            const bool last_strip = strip == nstrips - 1;
            int T[K], E[K], rowbase[K];
            SWord Hs[K], Es[K];
            int rowq[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int i = i0 + k;
                const bool valid = i < Lq;
                const int qi = valid ? (int)q[i] : 0;
                rowq[k] = valid ? qi : -1;
                rowbase[k] = valid ? (p.is_pssm ? i * size : qi * size) : -1;
                // left boundary H[i][-1]
                int hl = left_free ? 0 : -o - i * e;
                if (!PROF && p.banded && -(i + 1) < p.band_lo) hl = NEG_INF32;   // column -1 of row i
                T[k] = hl - o;
                E[k] = NEG_INF32;
                Hs[k] = 0; Es[k] = 0;
            }
            if (PROF) {
                sync_warp();
                for (int a = 0; a < size; ++a) {
                    unsigned wv[4] = {0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u};  // pad rows: -128
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        if (rowq[k] >= 0) {
                            const unsigned b = (unsigned)smat[rowq[k] * size + a] & 0xffu;
                            wv[k >> 2] = (wv[k >> 2] & ~(0xffu << (8 * (k & 3)))) | (b << (8 * (k & 3)));
                        }
                    }
                    uint4 v; v.x = wv[0]; v.y = wv[1]; v.z = wv[2]; v.w = wv[3];
                    *(uint4 *)(wprof + ((size_t)a * 32 + lane) * 16) = v;
                }
                sync_warp();
            }
            // H[i0-1][-1]: corner of the whole table for i0 == 0, else the left boundary above
            int Tdiag_in = (i0 == 0) ? -o : ((left_free ? 0 : -o - (i0 - 1) * e) - o);
            if (!PROF && p.banded && i0 > 0 && -i0 < p.band_lo) Tdiag_in = NEG_INF32 - o;
            SWord Hsdiag_in = 0;
            int Tout = 0, Fout = NEG_INF32;   // bottom row of this lane, previous step
            SWord Hsout = 0, Fsout = 0;
            // row Lq-1 lives in this lane/register during the last strip
            const int klast = (Lq - 1) - i0;  // in [0,K) only for the owning lane

            for (int s = 0; s < nsteps; ++s) {
                if ((s & 31) == 0) {
                    // stage the next 32 columns: reference residues and the strip boundary
                    sync_warp();
                    const int c = s + lane;
                    if (c < Lr) {
                        if (packed) {
                            unsigned code;
                            if (p.r_bits == 2) code = (p.r_words[ro + (c >> 4)] >> (2 * (c & 15))) & 3u;
                            else if (p.r_bits == 3) {
                                const int w = (int)(((unsigned long long)(unsigned)c * 0xCCCCCCCDull) >> 35);   // c / 10
                                code = (p.r_words[ro + w] >> (3 * (c - 10 * w))) & 7u;
                            } else {
                                const int w = (int)(((unsigned long long)(unsigned)c * 0xAAAAAAABull) >> 34);   // c / 6
                                code = (p.r_words[ro + w] >> (5 * (c - 6 * w))) & 31u;
                            }
                            ringL[c & 63] = (uint8_t)code;
                        } else {
                            ringL[c & 63] = r[c];
                        }
                        if (strip > 0) {
                            ringT[c & 63] = ld_cg(bndT + c);
                            ringF[c & 63] = ld_cg(bndF + c);
                            if (STATS) { ringHs[c & 63] = ld_cg(bndHs + c); ringFs[c & 63] = ld_cg(bndFs + c); }
                        }
                    }
                    sync_warp();
                }
                const int j = s - lane;
                // bottom row of the lane above, one step old
                int Tup = shfl_up(Tout, 1);
                int Fup = shfl_up(Fout, 1);
                SWord Hsup = 0, Fsup = 0;
                if (STATS) { Hsup = shfl_up(Hsout, 1); Fsup = shfl_up(Fsout, 1); }
                const bool active = j >= 0 && j < Lr;
                if (lane == 0 && active) {
                    if (strip == 0) {
                        Tup = (top_free ? 0 : -o - j * e) - o;
                        if (!PROF && p.banded && j + 1 > p.band_hi) Tup = NEG_INF32 - o;   // row -1 of column j
                        Fup = NEG_INF32;
                        Hsup = 0; Fsup = 0;
                    } else {
                        Tup = ringT[j & 63]; Fup = ringF[j & 63];
                        if (STATS) { Hsup = ringHs[j & 63]; Fsup = ringFs[j & 63]; }
                    }
                }
                if (active) {
                    const int letter = (int)ringL[j & 63];
                    unsigned pw[4] = {0, 0, 0, 0};
                    if (PROF) {
                        const uint4 v = *(const uint4 *)(wprof + ((size_t)letter * 32 + lane) * 16);
                        pw[0] = v.x; pw[1] = v.y; pw[2] = v.z; pw[3] = v.w;
                    }
                    int Tu = Tup, Fu = Fup, Td = Tdiag_in;
                    SWord Hsu = Hsup, Fsu = Fsup, Hsd = Hsdiag_in;
                    int cmax = -0x7fffffff - 1;  // sw: column max of H*16 + (15-k)
                    unsigned long long trw = 0;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        int So;
                        if (PROF) {
                            // sign-extend byte k of the 16-byte profile slot
                            const unsigned SEL = (unsigned)(k & 3) * 0x1111u + 0x8880u;
                            So = (int)prmt(pw[k >> 2], 0u, SEL);
                        }
                        else if (rowbase[k] < 0) So = PAD_SCORE;
                        else if (square_in_smem) So = smat[rowbase[k] + letter];
                        else So = ld_ro(p.matrix + rowbase[k] + letter) + o;
                        const int Tl = T[k];
                        int H, Fn, En;
                        if (!STATS && !TRACE && !TABLE) {
                            En = viaddmax(E[k], -e, Tl);
                            Fn = viaddmax(Fu, -e, Tu);
                            const int h = viaddmax(Td, So, En);
                            H = is_sw ? vimax3(h, Fn, 0) : (h > Fn ? h : Fn);
                            if (!PROF && p.banded) {
                                const int d = j - (i0 + k);
                                if (d < p.band_lo || d > p.band_hi) { H = NEG_INF32; En = NEG_INF32; Fn = NEG_INF32; }
                            }
                        } else {
                            const int Eext = E[k] - e, Fext = Fu - e;
                            const bool eopen = Tl > Eext, fopen = Tu > Fext;
                            En = eopen ? Tl : Eext;
                            Fn = fopen ? Tu : Fext;
                            const int hd = Td + So;
                            const bool isdiag = hd >= En && hd >= Fn;
                            const bool isf = !isdiag && Fn >= En;
                            H = isdiag ? hd : (isf ? Fn : En);
                            const bool zero = is_sw && H <= 0;
                            if (zero) H = 0;
                            if (STATS) {
                                const SWord one = (SWord)1 << SP::LS;
                                const SWord es = (eopen ? Hs[k] : Es[k]) + one;
                                const SWord fs = (fopen ? Hsu : Fsu) + one;
                                SWord inc = one;
                                if (rowq[k] == letter) inc += 1;
                                if (So - o > 0) inc += (SWord)1 << SP::SS;
                                SWord hs = isdiag ? (Hsd + inc) : (isf ? fs : es);
                                if (zero) hs = 0;
                                Hsd = Hs[k];   // becomes the diagonal of the row below
                                Hs[k] = hs; Es[k] = es;
                                Hsu = hs; Fsu = fs;
                            }
                            if (TABLE) {
                                const size_t idx = (size_t)p.tab_off[pid] + (((size_t)strip * nsteps + s) * 32 + lane) * K + k;
                                p.tabH[idx] = H;
                                if (STATS) {
                                    p.tabM[idx] = (int)(Hs[k] & SP::MM);
                                    p.tabS[idx] = (int)((Hs[k] >> SP::SS) & SP::SM);
                                    p.tabL[idx] = (int)((Hs[k] >> SP::LS) & SP::LM);
                                }
                            }
                            if (TRACE) {
                                unsigned t = (eopen ? TR_DIAG_E : TR_INS_E) | (fopen ? TR_DIAG_F : TR_DEL_F);
                                if (!zero) t |= isdiag ? TR_DIAG : (isf ? TR_DEL : TR_INS);
                                trw |= (unsigned long long)t << (8 * (k & 7));
                                if ((k & 7) == 7 || k == K - 1) {
                                    // [strip][step][lane][K] bytes; K bytes of a lane are contiguous
                                    uint8_t *dst = tr + (((size_t)strip * nsteps + s) * 32 + lane) * K + (k & ~7);
                                    const int nb = (k & 7) + 1;
                                    if (nb == 8 && (K % 8) == 0) *(unsigned long long *)dst = trw;
                                    else for (int b = 0; b < nb; ++b) dst[b] = (uint8_t)(trw >> (8 * b));
                                    trw = 0;
                                }
                            }
                        }
                        Td = Tl;  // old T of this row = diagonal of the row below
                        const int Tn = H - o;
                        T[k] = Tn; E[k] = En;
                        Tu = Tn; Fu = Fn;
                        if (is_sw) {
                            const int key = H * 16 + (15 - k);
                            cmax = cmax > key ? cmax : key;
                        }
                    }
                    Tdiag_in = Tup;
                    if (STATS) Hsdiag_in = Hsup;
                    Tout = Tu; Fout = Fu;
                    if (STATS) { Hsout = Hsu; Fsout = Fsu; }
                    if (lane == 31 && !last_strip) {
                        st_cg(bndT + j, Tu); st_cg(bndF + j, Fu);
                        if (STATS) { st_cg(bndHs + j, Hsu); st_cg(bndFs + j, Fsu); }
                    }
                    // ---- end-cell bookkeeping ------------------------------------------------
                    if (is_sw) {
                        const int ch = cmax >> 4;
                        if (ch > bestH || (ch == bestH && j < bestJ)) {
                            const int kk = 15 - (cmax & 15);
                            bestH = ch; bestJ = j; bestI = i0 + kk;
                            if (STATS) {
#pragma unroll
                                for (int k = 0; k < K; ++k) if (k == kk) bestS = Hs[k];
                            }
                        }
                    } else if (last_strip && klast >= 0 && klast < K && (row_ends || (j == Lr - 1 && !col_ends))) {
                        // last row: sg scans it left to right (strict >); nw / sg-without-free-ends
                        // read the corner only
                        int hv = 0; SWord sv = 0;
#pragma unroll
                        for (int k = 0; k < K; ++k) if (k == klast) { hv = T[k] + o; if (STATS) sv = Hs[k]; }
                        if (hv > bestH) { bestH = hv; bestJ = j; bestI = Lq - 1; bestS = sv; }
                    }
                    if (col_ends && j == Lr - 1) {
                        // last column, rows top to bottom, strict >
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const int hv = T[k] + o;
                            if (i0 + k < Lq && hv > colH) { colH = hv; colI = i0 + k; if (STATS) colS = Hs[k]; }
                        }
                    }
                }
            }
        }
        // ---- merge lanes: (score desc, end_ref asc, end_query asc) -----------------------------
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const int oH = shfl_xor(bestH, m), oJ = shfl_xor(bestJ, m), oI = shfl_xor(bestI, m);
            SWord oS = 0;
            if (STATS) oS = shfl_xor(bestS, m);
            const bool take = oH > bestH || (oH == bestH && (oJ < bestJ || (oJ == bestJ && oI < bestI)));
            if (take) { bestH = oH; bestJ = oJ; bestI = oI; bestS = oS; }
            if (col_ends) {
                const int cH = shfl_xor(colH, m), cI = shfl_xor(colI, m);
                SWord cS = 0;
                if (STATS) cS = shfl_xor(colS, m);
                if (cH > colH || (cH == colH && cI < colI)) { colH = cH; colI = cI; colS = cS; }
            }
        }
        if (col_ends && (!row_ends || colH > bestH)) { bestH = colH; bestJ = Lr - 1; bestI = colI; bestS = colS; }
        if (lane == 0) {
            if (is_sw && bestH <= 0) { bestH = 0; bestJ = 0; bestI = 0; bestS = 0; }
            const int oid = p.out_map ? p.out_map[pid] : pid;
            p.score[oid] = bestH; p.end_query[oid] = bestI; p.end_ref[oid] = bestJ;
            if (STATS) {
                p.matches[oid] = (int)(bestS & SP::MM);
                p.similar[oid] = (int)((bestS >> SP::SS) & SP::SM);
                p.length[oid] = (int)((bestS >> SP::LS) & SP::LM);
            }
        }
    }
}

}  // namespace psb

// kern_wave32.cuh -- one LONG pair spread over the whole GPU: the anti-diagonal wavefront of the
// north star ("long-pair Smith-Waterman: 100 kb x 100 kb ... 32-bit scores, single-GPU intra-sequence
// parallelism", BASELINE config C5).  Replaces the same upstream kernels as kern_gotoh32.cuh
// (parasail_{nw,sg*,sw}_{striped,scan,diag}_{32,64,sat}, score + end cell) for pairs whose query is
// thousands of residues long; results follow SURVEY.md Appendix A.
//
// The table is cut into horizontal strips of 32*K query rows.  A strip is swept left to right by
// one warp exactly as in gotoh32_kernel (lane t owns K rows, column j = s - t, bottom row handed
// down by __shfl_up: the in-warp anti-diagonal).  Strips are claimed in order from an atomic
// counter by persistent warps that are all resident at once; strip b may start column chunk c as
// soon as strip b-1 has published the bottom row (T, F) of that chunk to its boundary line in
// global memory -- so the strips themselves form a second, coarser anti-diagonal wavefront across
// the SMs.  Publication is a release-store (st.release.gpu) of a column counter by the lane that wrote
// the boundary values; the consumer's lane 0 polls it with an acquire-load and the warp then reads the
// line behind a warp barrier.  Waiting is deadlock-free because a warp only ever waits for the
// strip claimed immediately before its own, which belongs to a warp that is running or finished.
#pragma once
#include "psb_defs.h"
#include "psb_simt.h"

#ifndef PSB_UNROLL
#define PSB_PRAGMA_(x) _Pragma(#x)
#define PSB_UNROLL(n) PSB_PRAGMA_(unroll n)
#endif

namespace psb {

struct Wave32Params {
    const uint8_t *q;        // mapped residues of the pair
    const uint8_t *r;
    int Lq, Lr;
    const int *matrix;       // square, size*size
    int size;
    int open, gap;
    int mode, s1_beg, s1_end, s2_beg, s2_end;
    int *bnd;                // nstrips lines of 2*Lr ints: T (= H - open) then F of the strip's bottom row
    int *progress;           // per strip: number of bottom-row columns published
    int *next_strip;         // work counter
    int *cand;               // per strip 8 ints: bestH, bestJ, bestI, colH, colI
    // multi-pair form (database scans: one query against several long subjects in one launch):
    // when multi_n > 0 the pair of item (subject s, strip b) is (q, r + r_off[s]) with Lr = r_len[s];
    // bnd / progress / cand are laid out subject-major (bnd at 2 * nstrips * r_off[s] ints).
    int multi_n;
    const long long *r_off;  // multi_n + 1 byte offsets into r
};

inline size_t wave32_smem_bytes(int size, int warps) {
    return (((size_t)size * size * sizeof(int) + 15) & ~(size_t)15) + (size_t)warps * (64 + 64 * 2 * sizeof(int));
}

#if !defined(PSB_EMULATE)
PSB_DEV int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
PSB_DEV void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
PSB_DEV void backoff() { __nanosleep(64); }
// 64-bit relaxed accesses at GPU scope: single-copy atomic, served by L2, never hoisted out of a poll loop
PSB_DEV long long ld_relaxed64(const long long *p) {
    long long v;
    asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
PSB_DEV void st_relaxed64(long long *p, long long v) { asm volatile("st.relaxed.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
PSB_DEV int wave_time_us() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (int)(t >> 10); }
#else
inline int wave_time_us() { return 0; }
inline long long ld_relaxed64(const long long *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void st_relaxed64(long long *p, long long v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
inline int ld_acquire(const int *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release(int *p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline void backoff() {}
#endif

template <int K>
PSB_KERNEL void wave32_kernel(Wave32Params p) {
    PSB_SHARED_DECL(smem_raw);
    const int lane = lane_id();
    const int size = p.size, o = p.open, e = p.gap;
    int *smat = (int *)smem_raw;
    const size_t mat_bytes = (((size_t)size * size * sizeof(int)) + 15) & ~(size_t)15;
    unsigned char *ring = smem_raw + mat_bytes + (size_t)warp_in_block() * (64 + 64 * 2 * sizeof(int));
    int *ringT = (int *)ring, *ringF = ringT + 64;
    uint8_t *ringL = (uint8_t *)(ringF + 64);
    for (int x = thread_in_block(); x < size * size; x += threads_per_block()) smat[x] = p.matrix[x] + o;
    sync_block();

    const int mode = p.mode;
    const bool is_sw = mode == MODE_SW;
    const bool top_free = is_sw || (mode == MODE_SG && p.s1_beg);
    const bool left_free = is_sw || (mode == MODE_SG && p.s2_beg);
    const bool row_ends = mode == MODE_SG && p.s1_end;
    const bool col_ends = mode == MODE_SG && p.s2_end;
    const int Lq = p.Lq;
    const int rows_per_strip = 32 * K;
    const int nstrips = (Lq + rows_per_strip - 1) / rows_per_strip;
    const int nitems = nstrips * (p.multi_n > 0 ? p.multi_n : 1);

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomic_add(p.next_strip, 1);
        item = shfl(item, 0);
        if (item >= nitems) break;
        // items are claimed in order, subject-major: strip b of a subject is always claimed after
        // its strip b-1
        const int subj = item / nstrips, strip = item - subj * nstrips;
        const long long rbase = p.multi_n > 0 ? p.r_off[subj] : 0;
        const int Lr = p.multi_n > 0 ? (int)(p.r_off[subj + 1] - rbase) : p.Lr;
        const uint8_t *rseq = p.r + rbase;
        const int nsteps = Lr + 31;
        int *bnd0 = p.bnd + 2ll * nstrips * rbase;          // this subject's boundary lines
        int *progress = p.progress + (long long)subj * nstrips;
        const bool last_strip = strip == nstrips - 1;
        const int i0 = strip * rows_per_strip + lane * K;
        int *bndT_in = bnd0 + (long long)(strip - 1) * 2 * Lr, *bndF_in = bndT_in + Lr;   // written by strip-1
        int *bndT_out = bnd0 + (long long)strip * 2 * Lr, *bndF_out = bndT_out + Lr;
        int T[K], E[K], rowbase[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = i0 + k;
            rowbase[k] = i < Lq ? (int)p.q[i] * size : -1;
            T[k] = (left_free ? 0 : -o - i * e) - o;
            E[k] = NEG_INF32;
        }
        int Tdiag_in = (i0 == 0) ? -o : ((left_free ? 0 : -o - (i0 - 1) * e) - o);
        int Tout = 0, Fout = NEG_INF32;
        int bestH = NEG_INF32, bestJ = 0x7fffffff, bestI = 0x7fffffff, colH = NEG_INF32, colI = 0x7fffffff;
        const int klast = (Lq - 1) - i0;

        for (int s = 0; s < nsteps; ++s) {
            if ((s & 31) == 0) {
                // publish what lane 31 has finished (columns < s - 31), then wait for the strip above
                sync_warp();
                if (!last_strip && lane == 31 && s >= 32) {
                    st_release(progress + strip, s - 31);
                }
                const int need = (s + 32 < Lr) ? s + 32 : Lr;   // columns [s, need) are staged now
                if (strip > 0 && s < Lr) {
                    if (lane == 0) while (ld_acquire(progress + strip - 1) < need) backoff();
                    sync_warp();
                }
                const int c = s + lane;
                if (c < Lr) {
                    ringL[c & 63] = rseq[c];
                    if (strip > 0) { ringT[c & 63] = ld_cg(bndT_in + c); ringF[c & 63] = ld_cg(bndF_in + c); }
                }
                sync_warp();
            }
            const int j = s - lane;
            int Tup = shfl_up(Tout, 1);
            int Fup = shfl_up(Fout, 1);
            const bool active = j >= 0 && j < Lr;
            if (lane == 0 && active) {
                if (strip == 0) { Tup = (top_free ? 0 : -o - j * e) - o; Fup = NEG_INF32; }
                else { Tup = ringT[j & 63]; Fup = ringF[j & 63]; }
            }
            if (active) {
                const int letter = (int)ringL[j & 63];
                int Tu = Tup, Fu = Fup, Td = Tdiag_in;
                int cmax = -0x7fffffff - 1;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int So = rowbase[k] < 0 ? PAD_SCORE : smat[rowbase[k] + letter];
                    const int Tl = T[k];
                    const int En = viaddmax(E[k], -e, Tl);
                    const int Fn = viaddmax(Fu, -e, Tu);
                    const int h = viaddmax(Td, So, En);
                    const int H = is_sw ? vimax3(h, Fn, 0) : (h > Fn ? h : Fn);
                    Td = Tl;
                    const int Tn = H - o;
                    T[k] = Tn; E[k] = En; Tu = Tn; Fu = Fn;
                    if (is_sw) { const int key = H * 16 + (15 - k); cmax = cmax > key ? cmax : key; }
                }
                Tdiag_in = Tup; Tout = Tu; Fout = Fu;
                if (lane == 31 && !last_strip) { st_cg(bndT_out + j, Tu); st_cg(bndF_out + j, Fu); }
                if (is_sw) {
                    const int ch = cmax >> 4;
                    if (ch > bestH) { bestH = ch; bestJ = j; bestI = i0 + 15 - (cmax & 15); }
                } else if (last_strip && klast >= 0 && klast < K && (row_ends || (j == Lr - 1 && !col_ends))) {
                    int hv = 0;
#pragma unroll
                    for (int k = 0; k < K; ++k) if (k == klast) hv = T[k] + o;
                    if (hv > bestH) { bestH = hv; bestJ = j; bestI = Lq - 1; }
                }
                if (col_ends && j == Lr - 1) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int hv = T[k] + o;
                        if (i0 + k < Lq && hv > colH) { colH = hv; colI = i0 + k; }
                    }
                }
            }
        }
        // the whole bottom row is out
        sync_warp();
        if (!last_strip && lane == 31) {
            st_release(progress + strip, Lr);
        }
        // merge the strip's lanes: (score desc, end_ref asc, end_query asc)
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const int oH = shfl_xor(bestH, m), oJ = shfl_xor(bestJ, m), oI = shfl_xor(bestI, m);
            if (oH > bestH || (oH == bestH && (oJ < bestJ || (oJ == bestJ && oI < bestI)))) { bestH = oH; bestJ = oJ; bestI = oI; }
            const int cH = shfl_xor(colH, m), cI = shfl_xor(colI, m);
            if (cH > colH || (cH == colH && cI < colI)) { colH = cH; colI = cI; }
        }
        if (lane == 0) {
            int *c = p.cand + (long long)item * 8;
            c[0] = bestH; c[1] = bestJ; c[2] = bestI; c[3] = colH; c[4] = colI;
        }
    }
}

// ---- generation 2 of the strip sweep: built for LATENCY ------------------------------------------
// A strip is one warp and every step hands its bottom row to the next lane, so the time of a long
// pair is (columns + strips * lag) x the latency of one step, not a throughput figure.  This form
// (needs open >= extend and matrix values + open that fit a byte; otherwise wave32_kernel runs)
//   * takes the serial part of a column down to ONE dependent instruction per row:
//       F[k] = max(F[k-1] - e, h0[k-1] - o)   with   h0 = max(Hdiag + S, E [,0])  off the chain
//     (exact because H[k-1] - o = max(h0[k-1], F[k-1]) - o and F[k-1] - o <= F[k-1] - e);
//   * reads the scores of a step with one LDS.128 from a per-warp int8 profile of the strip, whose
//     address comes from a residue fetched one step ahead;
//   * has no divergent branch in the step.
inline size_t wave32v2_smem_bytes(int size, int warps) {
    return (((size_t)size * size * sizeof(int) + 15) & ~(size_t)15) + (size_t)warps * (64 * 2 * sizeof(int) + 64 + 64 + (size_t)size * 512);
}

template <int K>
PSB_KERNEL void wave32v2_kernel(Wave32Params p) {
    PSB_SHARED_DECL(smem_raw);
    const int lane = lane_id();
    const int size = p.size, o = p.open, e = p.gap;
    int *smat = (int *)smem_raw;
    const size_t mat_bytes = (((size_t)size * size * sizeof(int)) + 15) & ~(size_t)15;
    const size_t per_warp = 64 * 2 * sizeof(int) + 64 + 64 + (size_t)size * 512;
    unsigned char *wsm = smem_raw + mat_bytes + (size_t)warp_in_block() * per_warp;
    int *ringT = (int *)wsm, *ringF = ringT + 64;
    uint8_t *ringL = (uint8_t *)(ringF + 64);           // 64 residues (+64 bytes of slack keep the profile 16-byte aligned)
    unsigned char *wprof = wsm + 64 * 2 * sizeof(int) + 128;
    for (int x = thread_in_block(); x < size * size; x += threads_per_block()) smat[x] = p.matrix[x] + o;
    sync_block();

    const int mode = p.mode;
    const bool is_sw = mode == MODE_SW;
    const bool top_free = is_sw || (mode == MODE_SG && p.s1_beg);
    const bool left_free = is_sw || (mode == MODE_SG && p.s2_beg);
    const bool row_ends = mode == MODE_SG && p.s1_end;
    const bool col_ends = mode == MODE_SG && p.s2_end;
    const int Lq = p.Lq;
    const int rows_per_strip = 32 * K;
    const int nstrips = (Lq + rows_per_strip - 1) / rows_per_strip;
    const int nitems = nstrips * (p.multi_n > 0 ? p.multi_n : 1);

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomic_add(p.next_strip, 1);
        item = shfl(item, 0);
        if (item >= nitems) break;
        const int subj = item / nstrips, strip = item - subj * nstrips;
        const long long rbase = p.multi_n > 0 ? p.r_off[subj] : 0;
        const int Lr = p.multi_n > 0 ? (int)(p.r_off[subj + 1] - rbase) : p.Lr;
        const uint8_t *rseq = p.r + rbase;
        const int nsteps = Lr + 31;
        int *bnd0 = p.bnd + 2ll * nstrips * rbase;
        int *progress = p.progress + (long long)subj * nstrips;
        const bool last_strip = strip == nstrips - 1;
        const int i0 = strip * rows_per_strip + lane * K;
        int *bndT_in = bnd0 + (long long)(strip - 1) * 2 * Lr, *bndF_in = bndT_in + Lr;
        int *bndT_out = bnd0 + (long long)strip * 2 * Lr, *bndF_out = bndT_out + Lr;

        // per-warp int8 profile of this strip: [letter][lane][16 rows] of (S + open), pad rows -128
        sync_warp();
        int T[K], E[K];
        for (int a = 0; a < size; ++a) {
            unsigned wv[4] = {0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u};
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (i0 + k < Lq) {
                    const unsigned b = (unsigned)smat[(int)p.q[i0 + k] * size + a] & 0xffu;
                    wv[k >> 2] = (wv[k >> 2] & ~(0xffu << (8 * (k & 3)))) | (b << (8 * (k & 3)));
                }
            }
            uint4 v; v.x = wv[0]; v.y = wv[1]; v.z = wv[2]; v.w = wv[3];
            *(uint4 *)(wprof + ((size_t)a * 32 + lane) * 16) = v;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            T[k] = (left_free ? 0 : -o - (i0 + k) * e) - o;
            E[k] = NEG_INF32;
        }
        int Tdiag_in = (i0 == 0) ? -o : ((left_free ? 0 : -o - (i0 - 1) * e) - o);
        int Tout = 0, Fout = NEG_INF32;
        int bestH = NEG_INF32, bestJ = 0x7fffffff, bestI = 0x7fffffff, colH = NEG_INF32, colI = 0x7fffffff;
        const int klast = (Lq - 1) - i0;
        // column 0 of the ring; every 32 steps the columns s+1 .. s+32 follow (one ahead, so that the
        // residue of the NEXT step is always staged)
        if (lane == 0) ringL[0] = rseq[0];
        if (strip > 0 && lane == 0) {
            while (ld_acquire(progress + strip - 1) < 1) backoff();
            ringT[0] = ld_cg(bndT_in); ringF[0] = ld_cg(bndF_in);
        }
        sync_warp();
        int letter_next = (int)ringL[0];   // lane t first uses it at step t (column 0)

        for (int s = 0; s < nsteps; ++s) {
            if ((s & 31) == 0) {
                sync_warp();
                if (!last_strip && lane == 31 && s >= 32) {
                    st_release(progress + strip, s - 31);
                }
                const int need = (s + 33 < Lr) ? s + 33 : Lr;   // columns [s+1, need) are staged now
                if (strip > 0 && s + 1 < Lr) {
                    if (lane == 0) while (ld_acquire(progress + strip - 1) < need) backoff();
                    sync_warp();
                }
                const int c = s + 1 + lane;
                if (c < Lr) {
                    ringL[c & 63] = rseq[c];
                    if (strip > 0) { ringT[c & 63] = ld_cg(bndT_in + c); ringF[c & 63] = ld_cg(bndF_in + c); }
                }
                sync_warp();
            }
            const int j = s - lane;
            const bool active = j >= 0 && j < Lr;
            const int jc = active ? j : 0;
            // scores of this step (address known since the previous step) and the next step's residue
            const int letter = letter_next;
            const uint4 pv = *(const uint4 *)(wprof + ((size_t)letter * 32 + lane) * 16);
            const unsigned pw[4] = {pv.x, pv.y, pv.z, pv.w};
            const int jn = j + 1;
            letter_next = (int)ringL[(jn >= 0 && jn < Lr ? jn : 0) & 63];
            int Tup = shfl_up(Tout, 1);
            int Fup = shfl_up(Fout, 1);
            if (lane == 0) {
                const int bt = ringT[jc & 63], bf = ringF[jc & 63];
                Tup = strip == 0 ? (top_free ? 0 : -o - jc * e) - o : bt;
                Fup = strip == 0 ? NEG_INF32 : bf;
            }
            if (active) {
                int Td = Tdiag_in;
                int Fk = viaddmax(Fup, -e, Tup);   // F of row 0: the one place the received T enters the chain
                int cmax = -0x7fffffff - 1;
                int Hlast = 0;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const unsigned SEL = (unsigned)(k & 3) * 0x1111u + 0x8880u;
                    const int So = (int)prmt(pw[k >> 2], 0u, SEL);
                    const int Tl = T[k];
                    const int En = viaddmax(E[k], -e, Tl);
                    const int h0 = is_sw ? viaddmax_relu(Td, So, En) : viaddmax(Td, So, En);
                    const int H = h0 > Fk ? h0 : Fk;
                    const int Fnext = viaddmax(Fk, -e, h0 - o);   // the only dependent op per row
                    Td = Tl;
                    T[k] = H - o; E[k] = En;
                    if (is_sw) { const int key = H * 16 + (15 - k); cmax = cmax > key ? cmax : key; }
                    if (k == K - 1) { Hlast = H; Fout = Fk; }
                    Fk = Fnext;
                }
                Tdiag_in = Tup; Tout = Hlast - o;
                if (lane == 31 && !last_strip) { st_cg(bndT_out + j, Tout); st_cg(bndF_out + j, Fout); }
                if (is_sw) {
                    const int ch = cmax >> 4;
                    if (ch > bestH) { bestH = ch; bestJ = j; bestI = i0 + 15 - (cmax & 15); }
                } else if (last_strip && klast >= 0 && klast < K && (row_ends || (j == Lr - 1 && !col_ends))) {
                    int hv = 0;
#pragma unroll
                    for (int k = 0; k < K; ++k) if (k == klast) hv = T[k] + o;
                    if (hv > bestH) { bestH = hv; bestJ = j; bestI = Lq - 1; }
                }
                if (col_ends && j == Lr - 1) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int hv = T[k] + o;
                        if (i0 + k < Lq && hv > colH) { colH = hv; colI = i0 + k; }
                    }
                }
            }
        }
        sync_warp();
        if (!last_strip && lane == 31) {
            st_release(progress + strip, Lr);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const int oH = shfl_xor(bestH, m), oJ = shfl_xor(bestJ, m), oI = shfl_xor(bestI, m);
            if (oH > bestH || (oH == bestH && (oJ < bestJ || (oJ == bestJ && oI < bestI)))) { bestH = oH; bestJ = oJ; bestI = oI; }
            const int cH = shfl_xor(colH, m), cI = shfl_xor(colI, m);
            if (cH > colH || (cH == colH && cI < colI)) { colH = cH; colI = cI; }
        }
        if (lane == 0) {
            int *c = p.cand + (long long)item * 8;
            c[0] = bestH; c[1] = bestJ; c[2] = bestI; c[3] = colH; c[4] = colI;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Generation 3 (open >= extend, byte-sized S + open): column-blocked strips.
//
// What bounds one long pair is the critical path, not throughput: the wavefront has to cross
// Lq/K lanes and Lr/C column blocks, one step each, and a warp that runs alone on its scheduler
// pays ~4 cycles per instruction whatever it does.  Generation 2 moved one column per step and
// spent ~100 of its ~130 instructions per step on bookkeeping (ncu, profiles/r1e_ncu_wave_*).
// Here a lane owns K rows and fills C consecutive columns per step (a K x C tile), so the
// per-step bookkeeping -- shuffles, ring reads, boundary stores, the end-cell update -- is paid once
// per C columns, the path is Lq/K + Lr/C steps long, and the tile's cells give the scheduler
// independent work (rows of column c+1 can start as soon as row 0 of column c is done).
//   * lane t fills column block b = s - t at step s; its bottom row's (T, F) of the C columns and
//     the block's four residues travel to lane t+1 by shfl_up;
//   * strip hand-over without fences or counters: lane 31 stores each bottom-row column as ONE
//     64-bit word (T with bit 30 flipped, F).  |T| < 2^30, so a written word has different top two
//     bits in its T half, which the zero-filled line can never show: every word validates itself.
//     Lane 0 of the strip below loads the four words of block s+2 straight from the line (relaxed,
//     L2) while it works on block s -- two alternating register sets, no moves -- and checks them when
//     their step comes; only a strip that is still catching up finds a word unwritten and re-polls.
//     A strip can therefore start 32 steps after the one above it (the skew of the 32 lanes) plus one
//     L2 round trip, and nothing else orders the strips;
//   * residues: a 64-column shared ring refilled every 8 steps from registers loaded one refill early;
//   * columns past the end of the subject carry a pad letter scoring -128.  Local: their H is a decayed
//     E of a real cell, which can never beat that cell (strict >), so the step has no column mask, and the
//     end cell is key = (H << BITS) + (K*C-1 - (c*K + k)) -- the maximum prefers the smaller column, then
//     the smaller row -- with a branch-free update of (best key, best block) (a per-column test of the column
//     maximum was measured: its four branches per step cost a lone warp more than the two instructions per
//     cell of the key).  Global / semi-global
//     (IS_SW = false): nothing reads a pad column's cells, the last row and last column are looked at
//     per column as in generation 2.
// `bnd` must be zero-filled before the launch; `progress` is not used by this generation.
// per warp: a 64-residue ring and a profile of 32-bit (S + open) words, [letter][4-row chunk][lane][16 B]
inline size_t wave32v3_smem_bytes(int size, int warps, int K) {
    return (((size_t)size * size * sizeof(int) + 15) & ~(size_t)15) + (size_t)warps * (64 + (size_t)(size + 1) * ((K + 3) / 4) * 512);
}
// local: scores must leave room for the tile index below them in the 32-bit key; every mode: |H - open|
// stays below 2^30 (the hand-over words keep their validity mark in the top two bits of T)
inline bool wave32v3_range_ok(int K, int C, bool is_sw, long long lq, long long lr, int max_score, int min_score, int open, int gap) {
    int bits = 0;
    while ((1 << bits) < K * C) ++bits;
    const long long up = (lq < lr ? lq : lr) * (long long)(max_score > 1 ? max_score : 1);
    if (is_sw) return up < (1ll << (30 - bits));
    long long step = -(long long)min_score;
    if (gap > step) step = gap;
    if (step < 1) step = 1;
    const long long down = (lq + lr) * step + 2ll * open + 64;
    return up < (1ll << 29) && down < (1ll << 29);
}
PSB_DEV bool wave32v3_valid(long long w) { const unsigned t = (unsigned)w; return (((t >> 30) ^ (t >> 31)) & 1u) != 0; }
PSB_DEV long long wave32v3_pack(int T, int F) { return (long long)(((unsigned long long)(unsigned)F << 32) | (unsigned)(T ^ 0x40000000)); }

template <int K, int C, bool IS_SW>
PSB_KERNEL void wave32v3_kernel(Wave32Params p) {
    static_assert(C == 4, "four residues travel in one word");
    static_assert(K <= 16 && K % 4 == 0, "rows per lane come in 16-byte profile chunks of four");
    constexpr int CH = K / 4;
    constexpr int KC = K * C;
    constexpr int BITS = KC <= 16 ? 4 : (KC <= 32 ? 5 : 6);
    PSB_SHARED_DECL(smem_raw);
    const int lane = lane_id();
    const int size = p.size, o = p.open, e = p.gap;
    int *smat = (int *)smem_raw;
    const size_t mat_bytes = (((size_t)size * size * sizeof(int)) + 15) & ~(size_t)15;
    const size_t per_warp = 64 + (size_t)(size + 1) * CH * 512;
    unsigned char *wsm = smem_raw + mat_bytes + (size_t)warp_in_block() * per_warp;
    uint8_t *ringL = (uint8_t *)wsm;                     // 64 residues
    unsigned char *wprof = wsm + 64;
    for (int x = thread_in_block(); x < size * size; x += threads_per_block()) smat[x] = p.matrix[x] + o;
    sync_block();

    const int mode = p.mode;
    const bool top_free = IS_SW || (mode == MODE_SG && p.s1_beg);
    const bool left_free = IS_SW || (mode == MODE_SG && p.s2_beg);
    const bool row_ends = !IS_SW && mode == MODE_SG && p.s1_end;
    const bool col_ends = !IS_SW && mode == MODE_SG && p.s2_end;
    const int Lq = p.Lq;
    const int rows_per_strip = 32 * K;
    const int nstrips = (Lq + rows_per_strip - 1) / rows_per_strip;
    const int nitems = nstrips * (p.multi_n > 0 ? p.multi_n : 1);
    // hand-over words carry T = H - o and Fh = F + o of the row below (the vertical gap of the next row,
    // ready to use: the one-instruction chain of kern_pairs16.cuh).  Free / local top edge: H = 0, so
    // T = -o and Fh = H = 0 (F opened from the edge).
    const long long top_edge = wave32v3_pack(-o, 0);

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomic_add(p.next_strip, 1);
        item = shfl(item, 0);
        if (item >= nitems) break;
        const int subj = item / nstrips, strip = item - subj * nstrips;
        const long long rbase = p.multi_n > 0 ? p.r_off[subj] : 0;
        const int Lr = p.multi_n > 0 ? (int)(p.r_off[subj + 1] - rbase) : p.Lr;
        const uint8_t *rseq = p.r + rbase;
        const int nblk = (Lr + C - 1) / C;
        const int nsteps = nblk + 31;
        long long *bnd0 = (long long *)p.bnd + (long long)nstrips * rbase;       // one 64-bit word per column
        const bool last_strip = strip == nstrips - 1;
        const int i0 = strip * rows_per_strip + lane * K;
        const long long *bnd_in = bnd0 + (long long)(strip - 1) * Lr;
        long long *bnd_out = bnd0 + (long long)strip * Lr;

        // per-warp profile of this strip: 32-bit (S + open) words, [letter][chunk][lane][4 rows]; pad rows and
        // the pad letter (index `size`) score -128.  One LDS.128 per four rows, no sign extension per cell.
        sync_warp();
        for (int a = 0; a <= size; ++a) {
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                int wv[4] = {-128, -128, -128, -128};
                if (a < size) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const int k = 4 * ch + k4;
                        if (i0 + k < Lq) wv[k4] = smat[(int)p.q[i0 + k] * size + a];
                    }
                }
                uint4 v; v.x = (unsigned)wv[0]; v.y = (unsigned)wv[1]; v.z = (unsigned)wv[2]; v.w = (unsigned)wv[3];
                *(uint4 *)(wprof + (((size_t)a * CH + ch) * 32 + lane) * 16) = v;
            }
        }
        int T[K], E[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            T[k] = (left_free ? 0 : -o - (i0 + k) * e) - o;   // column -1
            E[k] = NEG_INF32;
        }
        int Tdiag_in = (i0 == 0) ? -o : ((left_free ? 0 : -o - (i0 - 1) * e) - o);
        int Tout[C], Fout[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { Tout[c] = -o; Fout[c] = 0; }
        unsigned Lw_out = 0;
        int bestH = IS_SW ? 0 : NEG_INF32, bestKey = 0, bestB = 0;    // local: a score must exceed 0 to count
        int bestJ = 0x7fffffff, bestI = 0x7fffffff, colH = NEG_INF32, colI = 0x7fffffff;
        const int klast = (Lq - 1) - i0;
        const int t_claim = wave_time_us();   // debugging aid (PSB_DEBUG_TIMING): when the strip was claimed
        // residues of the refill of group g (columns [32g, 32g+32)), loaded one group early
        unsigned pre_l = lane < Lr ? (unsigned)rseq[lane] : (unsigned)size;
        // lane 0: the words of the strip above for the blocks of the next two steps
        long long WA[C], WB[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { WA[c] = top_edge; WB[c] = top_edge; }
        if (lane == 0 && strip > 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (c < Lr) WA[c] = ld_relaxed64(bnd_in + c);
                if (C + c < Lr) WB[c] = ld_relaxed64(bnd_in + C + c);
            }
        }
        // one step of the sweep; W holds lane 0's words for block s and is reloaded with block s+2
        auto step = [&](const int s, long long (&W)[C]) {
            const int b = s - lane;
            int Tup[C], Fup[C];
#pragma unroll
            for (int c = 0; c < C; ++c) { Tup[c] = shfl_up(Tout[c], 1); Fup[c] = shfl_up(Fout[c], 1); }
            unsigned Lw = shfl_up(Lw_out, 1);
            if (lane == 0) {
                if (strip > 0 && s < nblk) {
                    // words the strip above has not written yet: re-poll (only while this strip catches up)
                    for (;;) {
                        unsigned ok = 0x80000000u;
#pragma unroll
                        for (int c = 0; c < C; ++c) { const unsigned t = (unsigned)W[c]; ok &= t ^ (t << 1); }
                        if (ok & 0x80000000u) break;
#pragma unroll
                        for (int c = 0; c < C; ++c) if (C * s + c < Lr) W[c] = ld_relaxed64(bnd_in + C * s + c);
                    }
                }
#pragma unroll
                for (int c = 0; c < C; ++c) { Tup[c] = (int)((unsigned)W[c] ^ 0x40000000u); Fup[c] = (int)((unsigned long long)W[c] >> 32); }
                if (!IS_SW && strip == 0 && !top_free) {
                    // top edge of a global alignment: H(-1, j) = -o - j*e, F(0, j) opened from it
#pragma unroll
                    for (int c = 0; c < C; ++c) { Fup[c] = -o - (C * s + c) * e; Tup[c] = Fup[c] - o; }
                }
                Lw = *(const unsigned *)(ringL + ((C * s) & 63));
                if (strip > 0) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int cn = C * (s + 2) + c;
                        W[c] = cn < Lr ? ld_relaxed64(bnd_in + cn) : top_edge;
                    }
                }
            }
            Lw_out = Lw;
            if (b >= 0 && b < nblk) {
                int cmax = -0x7fffffff - 1;
                int Tdg = Tdiag_in;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const unsigned letter = (Lw >> (8 * c)) & 0xffu;
                    int So[K];
#pragma unroll
                    for (int ch = 0; ch < CH; ++ch) {
                        const uint4 pv = *(const uint4 *)(wprof + (((size_t)letter * CH + ch) * 32 + lane) * 16);
                        So[4 * ch] = (int)pv.x; So[4 * ch + 1] = (int)pv.y; So[4 * ch + 2] = (int)pv.z; So[4 * ch + 3] = (int)pv.w;
                    }
                    int Td = Tdg;
                    int Fk = Fup[c];            // Fh = F + o of this lane's first row
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int Tl = T[k];
                        const int En = viaddmax(E[k], -e, Tl);
                        const int h = viaddmax(Td, So[k], En);
                        const int H = IS_SW ? viaddmax_relu(Fk, -o, h) : viaddmax(Fk, -o, h);
                        Fk = viaddmax(Fk, -e, h);        // the only loop-carried op per row
                        Td = Tl;
                        if (IS_SW) {
                            // branch-free end cell: the maximum of (H, inverted tile index) prefers the smaller column, then the smaller row
                            const int key = (H << BITS) + (KC - 1 - (c * K + k));
                            cmax = cmax > key ? cmax : key;
                        }
                        T[k] = H - o; E[k] = En;
                    }
                    Tout[c] = T[K - 1]; Fout[c] = Fk;
                    Tdg = Tup[c];
                    const int j = C * b + c;
                    if (!IS_SW && j < Lr) {
                        if (last_strip && klast >= 0 && klast < K && (row_ends || (j == Lr - 1 && !col_ends))) {
                            // last row: sg scans it left to right (strict >); nw reads the corner only
                            int hv = 0;
#pragma unroll
                            for (int k = 0; k < K; ++k) if (k == klast) hv = T[k] + o;
                            if (hv > bestH) { bestH = hv; bestJ = j; bestI = Lq - 1; }
                        }
                        if (col_ends && j == Lr - 1) {
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                const int hv = T[k] + o;
                                if (i0 + k < Lq && hv > colH) { colH = hv; colI = i0 + k; }
                            }
                        }
                    }
                }
                Tdiag_in = Tup[C - 1];
                if (lane == 31 && !last_strip) {
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        if (C * b + c < Lr) st_relaxed64(bnd_out + C * b + c, wave32v3_pack(Tout[c], Fout[c]));
                }
                if (IS_SW) {
                    const bool upd = (cmax >> BITS) > bestH;
                    bestH = upd ? (cmax >> BITS) : bestH;
                    bestKey = upd ? cmax : bestKey;
                    bestB = upd ? b : bestB;
                }
            }
        };
        for (int s0 = 0; s0 < nsteps; s0 += 8) {
            // ---- every 8 steps: 32 residues into the ring, the next 32 into registers ----------------
            sync_warp();
            {
                const int c = C * s0 + lane;
                ringL[c & 63] = (uint8_t)pre_l;
                pre_l = c + 32 < Lr ? (unsigned)rseq[c + 32] : (unsigned)size;
                sync_warp();
            }
            PSB_UNROLL(1)
            for (int s = s0; s < s0 + 8; s += 2) {
                step(s, WA);
                step(s + 1, WB);
            }
        }
        sync_warp();
        if (IS_SW) {
            if (bestH > 0) {
                const int idx = KC - 1 - (bestKey & ((1 << BITS) - 1));
                bestJ = C * bestB + idx / K;
                bestI = i0 + idx % K;
            } else {
                bestH = NEG_INF32;
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const int oH = shfl_xor(bestH, m), oJ = shfl_xor(bestJ, m), oI = shfl_xor(bestI, m);
            if (oH > bestH || (oH == bestH && (oJ < bestJ || (oJ == bestJ && oI < bestI)))) { bestH = oH; bestJ = oJ; bestI = oI; }
            const int cH = shfl_xor(colH, m), cI = shfl_xor(colI, m);
            if (cH > colH || (cH == colH && cI < colI)) { colH = cH; colI = cI; }
        }
        if (lane == 0) {
            int *c = p.cand + (long long)item * 8;
            c[0] = bestH; c[1] = bestJ; c[2] = bestI; c[3] = colH; c[4] = colI;
            c[5] = t_claim; c[6] = 0; c[7] = wave_time_us();
        }
    }
}

// final pick over the per-strip candidates, same tie-breaks as the in-warp merge.  One thread per
// subject (a single pair is subject 0).
struct WaveReduceParams {
    const int *cand;
    int nstrips;
    int mode, s1_end, s2_end;
    int Lr;                             // single pair
    int *score, *end_query, *end_ref;   // outputs
    int multi_n;                        // > 0: subject s writes at out_map[first_id + s], Lr from r_off
    const long long *r_off;
    const int *out_map;
    int first_id;
};
PSB_KERNEL void wave32_reduce_kernel(WaveReduceParams p) {
    const int subj = block_id() * threads_per_block() + thread_in_block();
    if (subj >= (p.multi_n > 0 ? p.multi_n : 1)) return;
    const int Lr = p.multi_n > 0 ? (int)(p.r_off[subj + 1] - p.r_off[subj]) : p.Lr;
    int bestH = NEG_INF32, bestJ = 0x7fffffff, bestI = 0x7fffffff, colH = NEG_INF32, colI = 0x7fffffff;
    for (int s = 0; s < p.nstrips; ++s) {
        const int *c = p.cand + ((long long)subj * p.nstrips + s) * 8;
        if (c[0] > bestH || (c[0] == bestH && (c[1] < bestJ || (c[1] == bestJ && c[2] < bestI)))) { bestH = c[0]; bestJ = c[1]; bestI = c[2]; }
        if (c[3] > colH || (c[3] == colH && c[4] < colI)) { colH = c[3]; colI = c[4]; }
    }
    const bool row_ends = p.mode == MODE_SG && p.s1_end, col_ends = p.mode == MODE_SG && p.s2_end;
    if (col_ends && (!row_ends || colH > bestH)) { bestH = colH; bestJ = Lr - 1; bestI = colI; }
    if (p.mode == MODE_SW && bestH <= 0) { bestH = 0; bestJ = 0; bestI = 0; }
    const int o = p.multi_n > 0 ? (p.out_map ? p.out_map[p.first_id + subj] : p.first_id + subj) : 0;
    p.score[o] = bestH; p.end_query[o] = bestI; p.end_ref[o] = bestJ;
}

}  // namespace psb

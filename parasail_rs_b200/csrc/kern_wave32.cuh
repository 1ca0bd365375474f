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
    // TRACE instantiations of generation 3 (single pair, K = 8): per strip, step and lane one 32-byte record of
    // the tile's H low bytes and one 8-byte record of its gap decisions (see wave32v3_kernel / walk32_kernel)
    uint4 *trace_h;
    uint2 *trace_bits;
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
//
// TRACE = true (long pairs WITH traceback or statistics, one pair per launch): every lane also leaves, per step,
//   * the low byte of H of its K x C cells, bytes ordered [column][row] (one PRMT per 4/3 cells), and
//   * one bit per cell and gap direction: "the horizontal gap E(i,j) was OPENED from H(i,j-1)" and "the vertical
//     gap F(i+1,j) was OPENED from H(i,j)" (0 = extended; the sign bit of (gap - e) - (open candidate), shifted
//     into a word: two instructions per bit), first cell of the tile in bit 31;
// as two 16-byte and one 8-byte streaming store into records indexed [strip][step][lane], so a warp's stores of
// one step are contiguous.  walk32_kernel follows the path from the end cell: H is recovered exactly from the
// bytes along its way (neighbours differ by less than 128, pairs16_trace_ok), the diagonal is taken iff
// H(i-1,j-1) + S == H(i,j), and the length of a gap comes from the open/extend bits -- the same decisions as the
// flag bytes of gotoh32_kernel, at 1.25 bytes per cell and with a walk that is linear in the path.
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

#ifndef WAVE32_SW_KEYED
#define WAVE32_SW_KEYED 1   // measured on 100 kb x 100 kb local (profiles/r4g): score only keyed 20.4 ms / late 21.3; traced keyed 37.2 / late 33.9
#endif
#ifndef WAVE32_SW_FSTEP
#define WAVE32_SW_FSTEP 0   // 1: local launches take the all-lanes-active step too (measured slower with the keyed end cell: 22.1 ms)
#endif
template <bool B> struct WaveTag { static constexpr bool value = B; };
inline long long wave32v3_trace_records(int lq, int lr, int K) {   // records of trace_h (x 32 B) and trace_bits (x 8 B)
    return (long long)((lq + 32 * K - 1) / (32 * K)) * ((lr + 3) / 4 + 31) * 32;
}
template <int K, int C, bool IS_SW, bool TRACE = false>
PSB_KERNEL void wave32v3_kernel(Wave32Params p) {
    static_assert(C == 4, "four residues travel in one word");
    static_assert(!TRACE || K == 8, "a trace record is the 32 cells of an 8 x 4 tile");
    static_assert(K <= 16 && K % 4 == 0, "rows per lane come in 16-byte profile chunks of four");
    constexpr int CH = K / 4;
    constexpr int KC = K * C;
    // local end cell, two forms (WAVE32_SW_KEYED: bit 0 = score-only launches, bit 1 = traced launches use the key):
    //   keyed: a branch-free running maximum of (H << BITS) + inverted tile index (1.5 instructions per cell);
    //   late:  only the tile's maximum inside the tile, the cell located after it when a lane's best improves
    //          (0.5 instructions per cell, one short branch per step, the tile's H values live to its end)
    constexpr bool KEYED = IS_SW && (((WAVE32_SW_KEYED) >> (TRACE ? 1 : 0)) & 1) != 0;
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
    const int e_tie = e + (rules::GAP_OPEN_ON_TIE ? 1 : 0);
    (void)e_tie;

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
        int bestH = IS_SW ? 0 : NEG_INF32;    // local: a score must exceed 0 to count
        int bestKey = 0, bestB = 0;           // (keyed form)
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
        // ---- the K x C tile of block b (step s) and what follows it: hand-over of the bottom row, trace records, best cell.
        // ENDS = true adds the global / semi-global end-cell candidates (last query row, last reference column) after each
        // column, as selects; only the last strip and a lane's last block need them.  Either way the tile is ONE basic block
        // whose columns ptxas interleaves (per-column tests used to cost nw / sg half their speed: 100 kb x 100 kb 38.9 ms
        // against 20.4 local).
        auto tile = [&](auto ends_tag, const int b, const int s, const int (&Tup)[C], const int (&Fup)[C], const unsigned Lw) {
            constexpr bool ENDS = decltype(ends_tag)::value;
            int cmax = KEYED ? -0x7fffffff - 1 : 0;
            int Hs[IS_SW && !KEYED ? KC : 1];         // (local, late form) the tile's H values
            int Tdg = Tdiag_in;
            unsigned ebits = 0, fbits = 0, hb[8];   // (TRACE only)
            int Tc[K];                                // (ENDS only) T of reference column Lr - 1
#pragma unroll
            for (int k = 0; k < K; ++k) Tc[k] = 0;
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
                unsigned hrow[4];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int Tl = T[k];
                    const int En = viaddmax(E[k], -e, Tl);
                    const int h = viaddmax(Td, So[k], En);
                    const int H = IS_SW ? viaddmax_relu(Fk, -o, h) : viaddmax(Fk, -o, h);
                    if (TRACE) {
                        // sign bit set = the gap opens here (strictly better than extending it, rules::GAP_OPEN_ON_TIE)
                        ebits = funnel_l1((unsigned)(E[k] - e_tie - Tl), ebits);
                        fbits = funnel_l1((unsigned)(Fk - e_tie - h), fbits);
                        hrow[k & 3] = (unsigned)H;
                        if ((k & 3) == 3) hb[(2 * c + (k >> 2)) & 7] = prmt(prmt(hrow[0], hrow[1], 0x0040u), prmt(hrow[2], hrow[3], 0x0040u), 0x5410u);
                    }
                    Fk = viaddmax(Fk, -e, h);        // the only loop-carried op per row
                    Td = Tl;
                    if (KEYED) {
                        // branch-free end cell: the maximum of (H, inverted tile index) prefers the smaller column, then the smaller row
                        const int key = (H << BITS) + (KC - 1 - (c * K + k));
                        cmax = cmax > key ? cmax : key;
                    } else if (IS_SW) {
                        // end cell: only the tile's maximum is formed here (one VIMNMX3 per two cells); the H values stay
                        // in their registers until the test after the tile
                        Hs[c * K + k] = H;
                        if (k & 1) cmax = vimax3(cmax, Hs[c * K + k - 1], H);
                    }
                    T[k] = H - o; E[k] = En;
                }
                Tout[c] = T[K - 1]; Fout[c] = Fk;
                Tdg = Tup[c];
                const int j = C * b + c;
                if (ENDS) {
                    // last row: sg scans it left to right (strict >), nw reads the corner only; a lane that does not hold
                    // row Lq-1 keeps hv at -inf
                    int hv = NEG_INF32;
#pragma unroll
                    for (int k = 0; k < K; ++k) hv = (k == klast) ? T[k] + o : hv;
                    const bool upd = last_strip && j < Lr && (row_ends || (j == Lr - 1 && !col_ends)) && hv > bestH;
                    bestH = upd ? hv : bestH; bestJ = upd ? j : bestJ; bestI = upd ? Lq - 1 : bestI;
                    // last column: its cells are looked at after the tile
                    const bool lastc = j == Lr - 1;
#pragma unroll
                    for (int k = 0; k < K; ++k) Tc[k] = lastc ? T[k] : Tc[k];
                }
            }
            if (ENDS && col_ends && b == nblk - 1) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int hv = Tc[k] + o;
                    if (i0 + k < Lq && hv > colH) { colH = hv; colI = i0 + k; }
                }
            }
            Tdiag_in = Tup[C - 1];
            if (lane == 31 && !last_strip) {
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (C * b + c < Lr) st_relaxed64(bnd_out + C * b + c, wave32v3_pack(Tout[c], Fout[c]));
            }
            if (TRACE) {
                const long long rec = ((long long)strip * nsteps + s) * 32 + lane;
                uint4 v0, v1;
                v0.x = hb[0]; v0.y = hb[1]; v0.z = hb[2]; v0.w = hb[3];
                v1.x = hb[4]; v1.y = hb[5]; v1.z = hb[6]; v1.w = hb[7];
                st_cs(p.trace_h + 2 * rec, v0);
                st_cs(p.trace_h + 2 * rec + 1, v1);
                uint2 bw; bw.x = ebits; bw.y = fbits;
                st_cs(p.trace_bits + rec, bw);
            }
            if (KEYED) {
                const bool upd = (cmax >> BITS) > bestH;
                bestH = upd ? (cmax >> BITS) : bestH;
                bestKey = upd ? cmax : bestKey;
                bestB = upd ? b : bestB;
            } else if (IS_SW) {
                if (cmax > bestH) {
                    // rare (a lane's best improves in a few steps of a sweep) and after the tile, so the tile stays one
                    // basic block: the first cell in column-major order that holds the maximum -- the smaller column,
                    // then the smaller row, as the tie-break wants
                    int idx = 0;
#pragma unroll
                    for (int x = KC - 1; x >= 0; --x) idx = Hs[x] == cmax ? x : idx;
                    bestH = cmax; bestJ = C * b + idx / K; bestI = i0 + idx % K;
                }
            }
        };
        // lane 0: the words of block s must have been written by the strip above; re-poll while this strip catches up
        auto await_words = [&](const int s, long long (&W)[C]) {
            for (;;) {
                unsigned ok = 0x80000000u;
#pragma unroll
                for (int c = 0; c < C; ++c) { const unsigned t = (unsigned)W[c]; ok &= t ^ (t << 1); }
                if (ok & 0x80000000u) break;
#pragma unroll
                for (int c = 0; c < C; ++c) if (C * s + c < Lr) W[c] = ld_relaxed64(bnd_in + C * s + c);
            }
        };
        // one step of the sweep, any step: W holds lane 0's words for block s and is reloaded with block s+2
        auto step = [&](const int s, long long (&W)[C]) {
            const int b = s - lane;
            int Tup[C], Fup[C];
#pragma unroll
            for (int c = 0; c < C; ++c) { Tup[c] = shfl_up(Tout[c], 1); Fup[c] = shfl_up(Fout[c], 1); }
            unsigned Lw = shfl_up(Lw_out, 1);
            if (lane == 0) {
                if (strip > 0 && s < nblk) await_words(s, W);
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
                if (IS_SW) tile(WaveTag<false>{}, b, s, Tup, Fup, Lw);
                else if ((last_strip && (row_ends || (klast >= 0 && klast < K && b == nblk - 1))) || (col_ends && b == nblk - 1)) tile(WaveTag<true>{}, b, s, Tup, Fup, Lw);
                else tile(WaveTag<false>{}, b, s, Tup, Fup, Lw);
            }
        };
        // the same step where every lane is inside the table and no lane is at its last block (31 <= s <= nblk - 4: the
        // steady state, all but ~64 steps of a long sweep): after lane 0's validity test everything -- its unpacking of the
        // words, the ring read, the loads for block s+2 (all lanes load the same words: one sector, no branch), the tile,
        // the predicated stores -- is ONE basic block, so the latency of lane 0's part hides under the tile instead of
        // preceding it
        const long long *bnd_safe = strip > 0 ? bnd_in : bnd_out;     // (strip 0 has no line above: loads go somewhere harmless)
        auto fstep = [&](auto ends_tag, const int s, long long (&W)[C]) {
            const bool l0 = lane == 0;
            if (l0 && strip > 0) await_words(s, W);
            int Tup[C], Fup[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int tu = shfl_up(Tout[c], 1), fu = shfl_up(Fout[c], 1);
                int t0 = (int)((unsigned)W[c] ^ 0x40000000u), f0 = (int)((unsigned long long)W[c] >> 32);
                if (!IS_SW && strip == 0 && !top_free) { f0 = -o - (C * s + c) * e; t0 = f0 - o; }
                Tup[c] = l0 ? t0 : tu; Fup[c] = l0 ? f0 : fu;
            }
            const unsigned lw_up = shfl_up(Lw_out, 1);
            const unsigned lw0 = *(const unsigned *)(ringL + ((C * s) & 63));
            const unsigned Lw = l0 ? lw0 : lw_up;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const long long w = ld_relaxed64(bnd_safe + C * (s + 2) + c);
                W[c] = strip > 0 ? w : top_edge;
            }
            Lw_out = Lw;
            tile(ends_tag, s - lane, s, Tup, Fup, Lw);
        };
        const bool ends_all = !IS_SW && last_strip && row_ends;   // sg: the last strip looks at its last row in every step
        for (int s0 = 0; s0 < nsteps; s0 += 8) {
            // ---- every 8 steps: 32 residues into the ring, the next 32 into registers ----------------
            sync_warp();
            {
                const int c = C * s0 + lane;
                ringL[c & 63] = (uint8_t)pre_l;
                pre_l = c + 32 < Lr ? (unsigned)rseq[c + 32] : (unsigned)size;
                sync_warp();
            }
            // (measured on 100 kb x 100 kb: global 21.0 -> 18.4 ms, semi-global 26.3 -> 24.8 ms, but local 20.3 -> 22.1 ms:
            // the local kernel keeps the general step)
            if ((!IS_SW || WAVE32_SW_FSTEP) && s0 >= 32 && s0 + 7 <= nblk - 4) {
                if (ends_all) {
                    PSB_UNROLL(1)
                    for (int s = s0; s < s0 + 8; s += 2) { fstep(WaveTag<!IS_SW>{}, s, WA); fstep(WaveTag<!IS_SW>{}, s + 1, WB); }   // (ends_all is never set for local)
                } else {
                    PSB_UNROLL(1)
                    for (int s = s0; s < s0 + 8; s += 2) { fstep(WaveTag<false>{}, s, WA); fstep(WaveTag<false>{}, s + 1, WB); }
                }
            } else {
                PSB_UNROLL(1)
                for (int s = s0; s < s0 + 8; s += 2) {
                    step(s, WA);
                    step(s + 1, WB);
                }
            }
        }
        sync_warp();
        if (KEYED && bestH > 0) {
            const int idx = KC - 1 - (bestKey & ((1 << BITS) - 1));
            bestJ = C * bestB + idx / K;
            bestI = i0 + idx % K;
        }
        if (IS_SW && !(bestH > 0)) bestH = NEG_INF32;
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

// ---- walk over the trace of a TRACE launch of wave32v3_kernel ----------------------------------
// One warp follows the path of the one pair from its end cell.  walk32_kernel<false> writes the CIGAR run
// list in reverse into the pair's scratch region (compact_cigar_kernel reverses it into the CSR);
// walk32_kernel<true> counts (matches, similar, length) of the path -- the `_stats` result, as in walk16_kernel.
struct Walk32Params {
    const uint8_t *q, *r;        // mapped residues of the pair
    int Lq, Lr;
    int K;                       // rows per lane of the fill (8)
    const uint4 *trace_h;
    const uint2 *trace_bits;
    const int *matrix;           // size x size substitution scores (no open added)
    int size;
    int open, gap;
    int is_sw, top_free, left_free;
    int pid;                     // index of the pair in the arrays below
    const int *score, *end_query, *end_ref;
    unsigned *rev_ops;           // CIGAR outputs
    const long long *rev_off;
    int *nops, *beg_query, *beg_ref;
    int *matches, *similar, *length;   // statistics outputs
};
inline size_t walk32_smem_bytes(int size) { return (size_t)size * size * sizeof(int); }

template <bool STATS>
PSB_KERNEL void walk32_kernel(Walk32Params p) {
    PSB_SHARED_DECL(smem_raw);
    int *smat = (int *)smem_raw;
    for (int x = thread_in_block(); x < p.size * p.size; x += threads_per_block()) smat[x] = p.matrix[x];
    sync_block();
    constexpr int K = 8, RPS = 32 * K;     // the fill's tile (TRACE instantiations exist for K = 8 only)
    if (block_id() != 0 || warp_in_block() != 0 || p.K != K) return;
    // ONE WARP walks the path.  Diagonal runs go 32 cells at a time: lane m looks at cell (i-m, j-m); if all cells
    // before it were diagonal steps its value is v minus the substitution scores so far (a warp prefix sum), and the
    // step from it is diagonal iff the low byte of (that value - S) is the stored byte of its predecessor (values
    // of neighbours differ by less than 128, so equal low bytes mean equal values); the first lane that fails ends
    // the run.  Gaps (rare) are resolved by lane 0 alone.  i, j, v and the output state are warp-uniform.
    const int lane = lane_id();
    const int pid = p.pid;
    const int Lr = p.Lr;
    const long long nsteps = (Lr + 3) / 4 + 31;
    const int o = p.open, e = p.gap;
    const uint8_t *hbytes = (const uint8_t *)p.trace_h;
    unsigned *out = STATS ? nullptr : p.rev_ops + p.rev_off[pid];
    // record and position inside the record of cell (i, j)
    auto locate = [&](int i, int j, int &kc) -> long long {
        const int strip = i / RPS, rem = i - strip * RPS, t = rem / K, k = rem - t * K;
        kc = (j & 3) * K + k;
        return ((long long)strip * nsteps + ((j >> 2) + t)) * 32 + t;
    };
    auto load = [&](int i, int j) -> unsigned { int kc; const long long rec = locate(i, j, kc); return hbytes[rec * 32 + kc]; };
    // 1: the horizontal gap E(i, j) was opened from H(i, j-1)        (0: extended from E(i, j-1))
    auto e_opened = [&](int i, int j) -> unsigned { int kc; const long long rec = locate(i, j, kc); return (p.trace_bits[rec].x >> (31 - kc)) & 1u; };
    // 1: the vertical gap F(i+1, j) was opened from H(i, j)          (0: extended from F(i, j))
    auto f_opened_below = [&](int i, int j) -> unsigned { int kc; const long long rec = locate(i, j, kc); return (p.trace_bits[rec].y >> (31 - kc)) & 1u; };
    auto recon = [](unsigned byte, int ref) -> int { return ref + (int)(signed char)(unsigned char)(byte - (unsigned)ref); };
    auto top = [&](int j) -> int { return (j < 0 || p.top_free) ? 0 : -o - j * e; };     // H[-1][j], corner 0
    auto left = [&](int i) -> int { return (i < 0 || p.left_free) ? 0 : -o - i * e; };   // H[i][-1]
    int i = p.end_query[pid], j = p.end_ref[pid];
    int v = p.score[pid];            // exact H[i][j]
    int cur = -1, n = 0;
    unsigned len = 0;
    int nm = 0, ns = 0, nl = 0;
    auto emit = [&](int op, int count) {          // (uniform; lane 0 stores)
        if (STATS || count <= 0) return;
        if (op == cur) len += (unsigned)count;
        else {
            if (cur >= 0) { if (lane == 0) out[n] = (len << 4) | (unsigned)cur; ++n; }
            cur = op; len = (unsigned)count;
        }
    };
    constexpr int DG = 4;  // loads in flight per round trip along a gap
    while (i >= 0 || j >= 0) {
        if (i < 0 || j < 0) {
            // off the table: statistics stop here; the CIGAR takes the rest of the other sequence as one run
            if (i < 0) { emit((int)rules::CIGAR_OP_D, j + 1); j = -1; }
            else { emit((int)rules::CIGAR_OP_I, i + 1); i = -1; }
            break;
        }
        // ---- a run of diagonal steps, 32 candidates at a time ------------------------------------------------
        const int ci = i - lane, cj = j - lane;
        const bool in = ci >= 0 && cj >= 0;
        const int a = in ? (int)p.q[ci] : 0, b = in ? (int)p.r[cj] : 0;
        const int sub = in ? smat[a * p.size + b] : 0;
        const bool inner = ci >= 1 && cj >= 1;
        const unsigned hbyte = inner ? load(ci - 1, cj - 1) : 0u;
        int pre = sub;                               // inclusive prefix sum of the substitution scores
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int up = shfl_up(pre, d); if (lane >= d) pre += up; }
        const int vm = v - (pre - sub);              // H of cell m if every step before it was diagonal
        const int need = vm - sub;                   // what its diagonal predecessor must hold (the diagonal wins ties)
        bool diag_ok = false;
        if (in) diag_ok = inner ? (((hbyte ^ (unsigned)need) & 0xffu) == 0u) : ((ci == 0 ? top(cj - 1) : left(ci - 1)) == need);
        const bool stop = in && p.is_sw && vm <= 0;  // ZERO: the alignment starts after this cell
        const unsigned badm = ballot(!in || stop || !diag_ok);
        const int nd = badm ? find_first_set(badm) - 1 : 32;          // diagonal steps taken now
        const unsigned take = nd >= 32 ? 0xffffffffu : ((1u << nd) - 1u);
        const unsigned eqm = ballot(in && a == b) & take, posm = ballot(in && sub > 0) & take;
        if (nd > 0) {
            nm += pop_count(eqm); ns += pop_count(posm); nl += nd;
            if (!STATS) {
                int pos = 0;
                while (pos < nd) {
                    const unsigned bit = (eqm >> pos) & 1u;
                    const unsigned x = (bit ? ~eqm : eqm) >> pos;        // zero bits while the run lasts
                    int run = x ? find_first_set(x) - 1 : 32 - pos;
                    if (run > nd - pos) run = nd - pos;
                    emit(bit ? (int)rules::CIGAR_OP_EQ : (int)rules::CIGAR_OP_X, run);
                    pos += run;
                }
            }
            v = shfl(need, nd - 1);
            i -= nd; j -= nd;
        }
        if (nd >= 32) continue;
        const int why_out = shfl((int)!in, nd), why_stop = shfl((int)stop, nd);
        if (why_out) continue;        // the run left the table: handled at the top of the loop
        if (why_stop) break;
        // ---- a gap at (i, j): lane 0 follows the open / extend bits ----------------------------------------------
        int gk = 0, gu = 0, gop = 0;   // length, H where the gap was opened, 1 = vertical (I) / 2 = horizontal (D) / 0 = inconsistent
        if (lane == 0) {
            // vertical first (F wins ties over E): follow F(i, j) up to where it was opened
            int u = v, k = 0, fval = 0;
            bool opened = false;
            while (!opened) {
                unsigned bt[DG], ob[DG];
#pragma unroll
                for (int m = 0; m < DG; ++m) {
                    const int row = i - k - 1 - m;
                    bt[m] = row >= 0 ? load(row, j) : 0u;
                    ob[m] = row >= 0 ? f_opened_below(row, j) : 1u;
                }
#pragma unroll
                for (int m = 0; m < DG; ++m) {
                    if (opened) break;
                    ++k;
                    const int row = i - k;
                    u = row < 0 ? top(j) : recon(bt[m], u);
                    if (ob[m]) { opened = true; fval = u - o - (k - 1) * e; }
                }
            }
            if (fval == v) { gk = k; gu = u; gop = 1; }
            else {
                // horizontal: follow E(i, j) left to where it was opened
                u = v; k = 0; opened = false;
                int evalue = 0;
                while (!opened) {
                    unsigned bt[DG], ob[DG];
#pragma unroll
                    for (int m = 0; m < DG; ++m) {
                        const int col = j - k - 1 - m;              // the H cell; the E cell right of it carries the bit
                        bt[m] = col >= 0 ? load(i, col) : 0u;
                        ob[m] = col + 1 >= 0 ? e_opened(i, col + 1) : 1u;
                    }
#pragma unroll
                    for (int m = 0; m < DG; ++m) {
                        if (opened) break;
                        ++k;
                        const int col = j - k;
                        u = col < 0 ? left(i) : recon(bt[m], u);
                        if (ob[m] || col < 0) { opened = true; evalue = u - o - (k - 1) * e; }
                    }
                }
                if (evalue == v) { gk = k; gu = u; gop = 2; }
            }
        }
        gk = shfl(gk, 0); gu = shfl(gu, 0); gop = shfl(gop, 0);
        if (gop == 0) break;   // cannot happen for a table the fill produced
        if (gop == 1) { emit((int)rules::CIGAR_OP_I, gk); i -= gk; }
        else { emit((int)rules::CIGAR_OP_D, gk); j -= gk; }
        nl += gk; v = gu;
    }
    if (lane != 0) return;
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

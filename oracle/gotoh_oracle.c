/*
 * gotoh_oracle.c -- TEST INFRASTRUCTURE ONLY.  Scalar CPU restatement of the
 * affine-gap Gotoh H/E/F fill that parasail-rs reaches through
 * parasail_lookup_function()/parasail_lookup_pfunction()
 * [REF src/aligner/mod.rs:339-351, 397-452] and of the trace walk behind
 * Alignment::get_cigar / get_traceback_strings [REF src/alignment/mod.rs:347-419].
 *
 * The arithmetic itself lives in the third-party dependency libparasail-sys
 * 0.2.1 [REF Cargo.toml:15, Cargo.lock:150-158] -> jeffdaily/parasail (C), which
 * is NOT present under /root/reference and cannot be fetched.  This file
 * restates parasail's published scalar algorithm (upstream src/nw.c, sg.c, sw.c
 * and their _stats / _trace / _table / _rowcol variants, src/cigar.c,
 * src/traceback.c) as written out in SURVEY.md Appendix A.  It is pinned
 * against every known-answer vector the reference's own tests hold
 * [REF tests/test_parasail.rs:64-616] (tests/test_oracle_reference_vectors.py);
 * beyond those trivial vectors PARITY IS UNPINNED (no parasail binary exists
 * here to diff against) -- see DESIGN.md "Oracle".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product library never links it.
 *
 * Orientation: s1 = query = rows i, s2 = reference = columns j.
 *   E = horizontal gap (consumes a reference char; TraceFlags INS; CIGAR 'D')
 *   F = vertical gap   (consumes a query char;     TraceFlags DEL; CIGAR 'I')
 * [REF src/alignment/table.rs:127-142 for the flag values]
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ctype.h>

#define PSBO_NEG_INF (INT32_MIN / 2)

/* TraceFlags bit values [REF src/alignment/table.rs:127-142] */
#define T_ZERO 0
#define T_INS 1
#define T_DEL 2
#define T_DIAG 4
#define T_DIAG_E 8
#define T_INS_E 16
#define T_DIAG_F 32
#define T_DEL_F 64

enum { PSBO_NW = 0, PSBO_SG = 1, PSBO_SW = 2 };

typedef struct psbo_config {
    int mode;    /* PSBO_NW / PSBO_SG / PSBO_SW */
    int s1_beg;  /* sg: gaps at the beginning of s1/query free  -> top row zero   (qb/qx) */
    int s1_end;  /* sg: gaps at the end of s1/query free        -> last row ends   (qe/qx) */
    int s2_beg;  /* sg: gaps at the beginning of s2/ref free    -> left col zero  (db/dx) */
    int s2_end;  /* sg: gaps at the end of s2/ref free          -> last col ends   (de/dx) */
    int open;    /* positive penalty; a gap of length k costs open + (k-1)*gap */
    int gap;
} psbo_config_t;

typedef struct psbo_matrix {
    const int *matrix; /* row-major, (pssm ? length : size) x size */
    const int *mapper; /* 256 entries, byte -> column index */
    int size;
    int is_pssm;       /* 1: row index is the query position, query bytes ignored for scoring */
    int length;        /* pssm rows */
} psbo_matrix_t;

typedef struct psbo_out {
    int score, end_query, end_ref;
    int matches, similar, length;
    /* optional full outputs; any pointer may be NULL */
    int8_t *trace;       /* qlen*rlen TraceFlags bytes, row-major */
    int *score_table;    /* qlen*rlen */
    int *matches_table, *similar_table, *length_table;
    int *score_row, *matches_row, *similar_row, *length_row; /* rlen: last row   */
    int *score_col, *matches_col, *similar_col, *length_col; /* qlen: last column */
} psbo_out_t;

/*
 * The fill.  Follows SURVEY.md A.2 (recurrence), A.3 (boundaries and end
 * candidates), A.4 (end-cell tie-breaks), A.5 (source priority inside a cell),
 * A.6 (statistics).  Returns 0, or -1 on allocation failure / bad arguments.
 */
int psbo_align(const uint8_t *q, int qlen, const uint8_t *r, int rlen,
               const psbo_config_t *cfg, const psbo_matrix_t *mat, psbo_out_t *out)
{
    const int o = cfg->open, e = cfg->gap;
    const int mode = cfg->mode;
    int s1_beg = 0, s1_end = 0, s2_beg = 0, s2_end = 0;
    if (mode == PSBO_SG) {
        s1_beg = cfg->s1_beg; s1_end = cfg->s1_end; s2_beg = cfg->s2_beg; s2_end = cfg->s2_end;
    }
    if (mat->is_pssm) qlen = mat->length;
    if (qlen <= 0 || rlen <= 0) return -1;

    const int size = mat->size;
    int *s1 = (int *)malloc(sizeof(int) * (size_t)qlen);
    int *s2 = (int *)malloc(sizeof(int) * (size_t)rlen);
    /* column-indexed previous-row state: H, F and their stats */
    size_t n = (size_t)rlen + 1;
    int *H = (int *)malloc(sizeof(int) * n), *F = (int *)malloc(sizeof(int) * n);
    int *HM = (int *)calloc(n, sizeof(int)), *HS = (int *)calloc(n, sizeof(int)), *HL = (int *)calloc(n, sizeof(int));
    int *FM = (int *)calloc(n, sizeof(int)), *FS = (int *)calloc(n, sizeof(int)), *FL = (int *)calloc(n, sizeof(int));
    if (!s1 || !s2 || !H || !F || !HM || !HS || !HL || !FM || !FS || !FL) return -1;

    for (int i = 0; i < qlen; ++i) s1[i] = (q && !mat->is_pssm) ? mat->mapper[q[i]] : (q ? mat->mapper[q[i]] : 0);
    for (int j = 0; j < rlen; ++j) s2[j] = mat->mapper[r[j]];

    /* top row H[-1][j] (A.3): zero for sw, and for sg when query-begin gaps are free */
    H[0] = 0; F[0] = PSBO_NEG_INF;
    for (int j = 1; j <= rlen; ++j) {
        H[j] = (mode == PSBO_SW || s1_beg) ? 0 : -o - (j - 1) * e;
        F[j] = PSBO_NEG_INF;
    }

    int score = PSBO_NEG_INF, end_query = qlen - 1, end_ref = rlen - 1;
    int matches = 0, similar = 0, length = 0;
    /* sg: best of the last column, rows scanned top to bottom (strict >) */
    int col_score = PSBO_NEG_INF, col_i = 0, col_m = 0, col_s = 0, col_l = 0;
    /* sw: an all-zero table yields score 0 at (0,0): the first cell beats -inf,
     * later zeros never have a smaller column (upstream scalar sw.c behaviour) */
    if (mode == PSBO_SW) { end_query = 0; end_ref = 0; }

    for (int i = 1; i <= qlen; ++i) {
        const int *matrow = &mat->matrix[(size_t)size * (mat->is_pssm ? (i - 1) : s1[i - 1])];
        int NH = H[0], NHM = HM[0], NHS = HS[0], NHL = HL[0];
        int WH = (mode == PSBO_SW || s2_beg) ? 0 : -o - (i - 1) * e; /* left column H[i][-1] */
        int WHM = 0, WHS = 0, WHL = 0;
        int E = PSBO_NEG_INF, EM = 0, ES = 0, EL = 0;
        H[0] = WH; HM[0] = 0; HS[0] = 0; HL[0] = 0;
        for (int j = 1; j <= rlen; ++j) {
            int NWH = NH, NWM = NHM, NWS = NHS, NWL = NHL;
            int tflag = 0;
            NH = H[j]; NHM = HM[j]; NHS = HS[j]; NHL = HL[j];
            /* F: vertical, opened from the cell above iff strictly better (A.5) */
            int F_opn = NH - o, F_ext = F[j] - e;
            if (F_opn > F_ext) { F[j] = F_opn; FM[j] = NHM; FS[j] = NHS; FL[j] = NHL + 1; tflag |= T_DIAG_F; }
            else               { F[j] = F_ext;                            FL[j] = FL[j] + 1; tflag |= T_DEL_F; }
            /* E: horizontal, opened from the cell to the left iff strictly better */
            int E_opn = WH - o, E_ext = E - e;
            if (E_opn > E_ext) { E = E_opn; EM = WHM; ES = WHS; EL = WHL + 1; tflag |= T_DIAG_E; }
            else               { E = E_ext;                     EL = EL + 1;  tflag |= T_INS_E; }
            int sub = matrow[s2[j - 1]];
            int H_dag = NWH + sub;
            /* H source priority: diagonal >= both, else F if F >= E, else E */
            if (H_dag >= E && H_dag >= F[j]) {
                WH = H_dag;
                WHM = NWM + (s1[i - 1] == s2[j - 1]);
                WHS = NWS + (sub > 0);
                WHL = NWL + 1;
                tflag |= T_DIAG;
            } else if (F[j] >= E) {
                WH = F[j]; WHM = FM[j]; WHS = FS[j]; WHL = FL[j]; tflag |= T_DEL;
            } else {
                WH = E; WHM = EM; WHS = ES; WHL = EL; tflag |= T_INS;
            }
            if (mode == PSBO_SW && WH <= 0) { /* ZERO wins, stats reset */
                WH = 0; WHM = 0; WHS = 0; WHL = 0;
                tflag &= ~(T_DIAG | T_DEL | T_INS);
            }
            H[j] = WH; HM[j] = WHM; HS[j] = WHS; HL[j] = WHL;

            size_t loc = (size_t)(i - 1) * (size_t)rlen + (size_t)(j - 1);
            if (out->trace) out->trace[loc] = (int8_t)tflag;
            if (out->score_table) out->score_table[loc] = WH;
            if (out->matches_table) out->matches_table[loc] = WHM;
            if (out->similar_table) out->similar_table[loc] = WHS;
            if (out->length_table) out->length_table[loc] = WHL;
            if (i == qlen) {
                if (out->score_row) out->score_row[j - 1] = WH;
                if (out->matches_row) out->matches_row[j - 1] = WHM;
                if (out->similar_row) out->similar_row[j - 1] = WHS;
                if (out->length_row) out->length_row[j - 1] = WHL;
            }
            if (j == rlen) {
                if (out->score_col) out->score_col[i - 1] = WH;
                if (out->matches_col) out->matches_col[i - 1] = WHM;
                if (out->similar_col) out->similar_col[i - 1] = WHS;
                if (out->length_col) out->length_col[i - 1] = WHL;
            }

            if (mode == PSBO_SW) {
                /* A.4: max score, then smallest end_ref, then smallest end_query */
                if (WH > score || (WH == score && (j - 1) < end_ref)) {
                    score = WH; end_query = i - 1; end_ref = j - 1;
                    matches = WHM; similar = WHS; length = WHL;
                }
            } else if (mode == PSBO_SG) {
                if (s1_end && i == qlen && WH > score) { /* last row, left to right, strict */
                    score = WH; end_query = i - 1; end_ref = j - 1;
                    matches = WHM; similar = WHS; length = WHL;
                }
                if (s2_end && j == rlen && WH > col_score) { /* last column, top to bottom, strict */
                    col_score = WH; col_i = i - 1; col_m = WHM; col_s = WHS; col_l = WHL;
                }
            }
        }
        if (i == qlen && (mode == PSBO_NW || (mode == PSBO_SG && !s1_end && !s2_end))) {
            score = WH; end_query = qlen - 1; end_ref = rlen - 1;
            matches = WHM; similar = WHS; length = WHL;
        }
    }
    if (mode == PSBO_SG && s2_end) {
        /* the last column beats the last row only when strictly greater (A.4) */
        if (!s1_end || col_score > score) {
            score = col_score; end_query = col_i; end_ref = rlen - 1;
            matches = col_m; similar = col_s; length = col_l;
        }
    }

    out->score = score; out->end_query = end_query; out->end_ref = end_ref;
    out->matches = matches; out->similar = similar; out->length = length;
    free(s1); free(s2); free(H); free(F); free(HM); free(HS); free(HL); free(FM); free(FS); free(FL);
    return 0;
}

/*
 * Trace walk -> CIGAR (A.7; upstream src/cigar.c).  ops_out receives
 * len<<4|op with ops numbered by "MIDNSHP=X" (I=1, D=2, '='=7, X=8), already
 * in forward order.  Capacity must be >= qlen + rlen.  Returns number of ops.
 *
 * Edge rule: the walk continues while either index is >= 0; once one sequence
 * is exhausted the remainder of the other is emitted as 'D' (i<0) or 'I'
 * (j<0).  In sw the walk also stops at a ZERO cell, leaving beg_query/beg_ref
 * at the first aligned cell.
 */
int psbo_cigar(const int8_t *trace, const uint8_t *q, int qlen, const uint8_t *r, int rlen,
               const psbo_matrix_t *mat, int end_query, int end_ref,
               uint32_t *ops_out, int *beg_query, int *beg_ref)
{
    (void)qlen;
    int64_t i = end_query, j = end_ref;
    int where = T_DIAG;
    uint32_t *rev = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(end_query + end_ref + 4));
    int nrev = 0;
    int cur_op = -1; uint32_t cur_len = 0;
#define EMIT(op_) do { \
        if (cur_op == (op_)) { cur_len++; } \
        else { \
            if (cur_op >= 0) { rev[nrev++] = (cur_len << 4) | (uint32_t)cur_op; } \
            cur_op = (op_); cur_len = 1; \
        } } while (0)
    while (i >= 0 || j >= 0) {
        if (i < 0) { EMIT(2); --j; continue; }
        if (j < 0) { EMIT(1); --i; continue; }
        int t = trace[(size_t)i * (size_t)rlen + (size_t)j];
        if (where == T_DIAG) {
            if (t & T_DIAG) {
                int a = mat->mapper[q[i]], b = mat->mapper[r[j]];
                EMIT(a == b ? 7 : 8);
                --i; --j;
            } else if (t & T_INS) { where = T_INS; }
            else if (t & T_DEL) { where = T_DEL; }
            else break; /* ZERO */
        } else if (where == T_INS) { /* E: horizontal, consumes reference */
            EMIT(2);
            where = (t & T_DIAG_E) ? T_DIAG : T_INS;
            --j;
        } else { /* F: vertical, consumes query */
            EMIT(1);
            where = (t & T_DIAG_F) ? T_DIAG : T_DEL;
            --i;
        }
    }
#undef EMIT
    if (cur_op >= 0) rev[nrev++] = (cur_len << 4) | (uint32_t)cur_op;
    for (int k = 0; k < nrev; ++k) ops_out[k] = rev[nrev - 1 - k];
    *beg_query = (int)(i + 1); *beg_ref = (int)(j + 1);
    free(rev);
    return nrev;
}

/*
 * Trace walk -> three alignment strings (upstream src/traceback.c, reached by
 * Alignment::get_traceback_strings [REF src/alignment/mod.rs:347-387]).
 * Buffers need qlen + rlen + 1 bytes.  Returns the alignment length.
 */
int psbo_traceback(const int8_t *trace, const uint8_t *q, int qlen, const uint8_t *r, int rlen,
                   const psbo_matrix_t *mat, int end_query, int end_ref,
                   char match, char pos, char neg, char *qs, char *cs, char *rs)
{
    (void)qlen;
    int64_t i = end_query, j = end_ref;
    int where = T_DIAG, n = 0;
    while (i >= 0 || j >= 0) {
        if (i < 0) { qs[n] = '-'; rs[n] = (char)r[j]; cs[n] = ' '; ++n; --j; continue; }
        if (j < 0) { qs[n] = (char)q[i]; rs[n] = '-'; cs[n] = ' '; ++n; --i; continue; }
        int t = trace[(size_t)i * (size_t)rlen + (size_t)j];
        if (where == T_DIAG) {
            if (t & T_DIAG) {
                int a = mat->mapper[q[i]], b = mat->mapper[r[j]];
                int sub = mat->matrix[(size_t)mat->size * (mat->is_pssm ? (int)i : a) + b];
                qs[n] = (char)q[i]; rs[n] = (char)r[j];
                cs[n] = (a == b) ? match : (sub > 0 ? pos : neg);
                ++n; --i; --j;
            } else if (t & T_INS) where = T_INS;
            else if (t & T_DEL) where = T_DEL;
            else break;
        } else if (where == T_INS) {
            qs[n] = '-'; rs[n] = (char)r[j]; cs[n] = ' '; ++n;
            where = (t & T_DIAG_E) ? T_DIAG : T_INS; --j;
        } else {
            qs[n] = (char)q[i]; rs[n] = '-'; cs[n] = ' '; ++n;
            where = (t & T_DIAG_F) ? T_DIAG : T_DEL; --i;
        }
    }
    for (int a = 0, b = n - 1; a < b; ++a, --b) {
        char t;
        t = qs[a]; qs[a] = qs[b]; qs[b] = t;
        t = cs[a]; cs[a] = cs[b]; cs[b] = t;
        t = rs[a]; rs[a] = rs[b]; rs[b] = t;
    }
    qs[n] = cs[n] = rs[n] = '\0';
    return n;
}

/*
 * Batch driver used by the parity tests and by bench.py's scalar "port"
 * baseline: n independent pairs, concatenated residues + offsets (n+1).
 * With want_cigar the trace table is built per pair and walked; ops go to a
 * CSR (cig_off has n+1 entries, cig_ops capacity cig_cap).  Returns 0 or -1.
 */
int psbo_align_batch(const uint8_t *qcat, const int64_t *qoff, const uint8_t *rcat, const int64_t *roff,
                     int64_t n, int shared_query, const psbo_config_t *cfg, const psbo_matrix_t *mat,
                     int *score, int *end_query, int *end_ref, int *matches, int *similar, int *length,
                     int want_cigar, uint32_t *cig_ops, int64_t cig_cap, int64_t *cig_off,
                     int *beg_query, int *beg_ref)
{
    int64_t used = 0;
    if (want_cigar && cig_off) cig_off[0] = 0;
    for (int64_t p = 0; p < n; ++p) {
        const uint8_t *q = shared_query ? qcat : qcat + qoff[p];
        int qlen = (int)(shared_query ? qoff[1] - qoff[0] : qoff[p + 1] - qoff[p]);
        const uint8_t *r = rcat + roff[p];
        int rlen = (int)(roff[p + 1] - roff[p]);
        psbo_out_t o; memset(&o, 0, sizeof(o));
        int8_t *trace = NULL;
        if (want_cigar) { trace = (int8_t *)malloc((size_t)qlen * (size_t)rlen); o.trace = trace; if (!trace) return -1; }
        if (psbo_align(q, qlen, r, rlen, cfg, mat, &o) != 0) { free(trace); return -1; }
        score[p] = o.score; end_query[p] = o.end_query; end_ref[p] = o.end_ref;
        if (matches) matches[p] = o.matches;
        if (similar) similar[p] = o.similar;
        if (length) length[p] = o.length;
        if (want_cigar) {
            if (used + qlen + rlen > cig_cap) { free(trace); return -1; }
            int bq, br;
            int nops = psbo_cigar(trace, q, qlen, r, rlen, mat, o.end_query, o.end_ref, cig_ops + used, &bq, &br);
            used += nops; cig_off[p + 1] = used; beg_query[p] = bq; beg_ref[p] = br;
            free(trace);
        }
    }
    return 0;
}

/* decode len<<4|op words to "12=1X3I" text; returns bytes written (excl. NUL) */
int psbo_cigar_decode(const uint32_t *ops, int nops, char *buf, int cap)
{
    static const char tab[] = "MIDNSHP=X";
    int w = 0;
    for (int k = 0; k < nops; ++k) {
        int op = (int)(ops[k] & 0xf);
        int m = snprintf(buf + w, (size_t)(cap - w), "%u%c", ops[k] >> 4, op > 8 ? 'M' : tab[op]);
        if (m < 0 || m >= cap - w) return -1;
        w += m;
    }
    if (w < cap) buf[w] = '\0';
    return w;
}

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
#include <pthread.h>

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

/*
 * Every behaviour below that is recalled from upstream parasail rather than read
 * from a source file ([UP] in SURVEY.md) has ONE switch here; the defaults are the
 * rules the product kernels implement (psb_defs.h, namespace psb::rules, mirrors
 * them).  oracle/UP_ASSUMPTIONS.md lists each rule, its confidence and what to
 * flip; tools/diff_vs_parasail.py tries the alternatives against a real
 * libparasail.so the day one is available.
 */
typedef struct psbo_rules {
    int sw_end_tie;          /* A.4  0: max score, then smaller end_ref, then smaller end_query (default)
                                     1: first maximum in row-major order (smaller end_query, then smaller end_ref)
                                     2: last maximum in row-major order */
    int sg_col_wins_tie;     /* A.4  0: last column beats last row only when strictly greater (default); 1: on ties too */
    int sg_row_last_wins;    /* A.4  0: first maximum of the last row / column wins (strict >, default); 1: last one (>=) */
    int h_priority;          /* A.5  0: diag >= F >= E (default); 1: diag >= E >= F; 2: F >= E > diag (gaps win ties);
                                     3: E >= F > diag */
    int open_on_tie;         /* A.5  0: a gap is opened only when strictly better than extending (default); 1: ties open */
    int match_raw_bytes;     /* A.6  0: a match is equality of mapped matrix indices (default); 1: of the raw bytes */
    int count_boundary_gaps; /* A.6  0: leading boundary gaps of nw are not counted in length (default); 1: counted */
    int cigar_edge_stop;     /* A.7  0: the walk runs to (-1,-1) in every mode, emitting the rest as I/D (default);
                                     1: sg/sw stop as soon as either index leaves the table (SURVEY's reading) */
    int cigar_swap_id;       /* A.7  0: 'I' consumes query, 'D' consumes reference (default); 1: swapped */
    int sg_flag_swap;        /* A.3  0: qb/qe act on the top/last row, db/de on the left/last column (default); 1: swapped */
    int zero_beats_diag;     /* A.5  1: sw cell with H == 0 is ZERO even when the diagonal sums to 0 (default); 0: diag kept */
    int band_rule;           /* f4   0: |i - j| <= k widened by the length difference so the corner is inside (default);
                                     1: plain |i - j| <= k */
} psbo_rules_t;
static psbo_rules_t g_rules = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0};
void psbo_set_rules(const psbo_rules_t *r) { if (r) g_rules = *r; else { psbo_rules_t d = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0}; g_rules = d; } }
void psbo_get_rules(psbo_rules_t *r) { *r = g_rules; }
int psbo_rules_count(void) { return (int)(sizeof(psbo_rules_t) / sizeof(int)); }

typedef struct psbo_config {
    int mode;    /* PSBO_NW / PSBO_SG / PSBO_SW */
    int s1_beg;  /* sg: gaps at the beginning of s1/query free  -> top row zero   (qb/qx) */
    int s1_end;  /* sg: gaps at the end of s1/query free        -> last row ends   (qe/qx) */
    int s2_beg;  /* sg: gaps at the beginning of s2/ref free    -> left col zero  (db/dx) */
    int s2_end;  /* sg: gaps at the end of s2/ref free          -> last col ends   (de/dx) */
    int open;    /* positive penalty; a gap of length k costs open + (k-1)*gap */
    int gap;
    int band;    /* nw only: > 0 restricts the fill to a band of half-width `band` (parasail_nw_banded); 0 = full table */
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
    const psbo_rules_t R = g_rules;
    if (mode == PSBO_SG) {
        s1_beg = cfg->s1_beg; s1_end = cfg->s1_end; s2_beg = cfg->s2_beg; s2_end = cfg->s2_end;
        if (R.sg_flag_swap) { int t; t = s1_beg; s1_beg = s2_beg; s2_beg = t; t = s1_end; s1_end = s2_end; s2_end = t; }
    }
    if (mat->is_pssm) qlen = mat->length;
    if (qlen <= 0 || rlen <= 0) return -1;
    /* banded nw: cell (i,j) is inside iff band_lo <= j - i <= band_hi */
    int banded = mode == PSBO_NW && cfg->band > 0, band_lo = 0, band_hi = 0;
    if (banded) {
        const int d = rlen - qlen;
        band_lo = -cfg->band + ((R.band_rule == 0 && d < 0) ? d : 0);
        band_hi = cfg->band + ((R.band_rule == 0 && d > 0) ? d : 0);
    }

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
        if (R.count_boundary_gaps && !(mode == PSBO_SW || s1_beg)) HL[j] = j;
        if (banded && j > band_hi) H[j] = PSBO_NEG_INF;   /* H[-1][j-1]: row -1, column j-1 */
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
        int WHM = 0, WHS = 0, WHL = (R.count_boundary_gaps && !(mode == PSBO_SW || s2_beg)) ? i : 0;
        if (banded && -i < band_lo) WH = PSBO_NEG_INF;
        int E = PSBO_NEG_INF, EM = 0, ES = 0, EL = 0;
        H[0] = WH; HM[0] = 0; HS[0] = 0; HL[0] = WHL;
        for (int j = 1; j <= rlen; ++j) {
            int NWH = NH, NWM = NHM, NWS = NHS, NWL = NHL;
            int tflag = 0;
            NH = H[j]; NHM = HM[j]; NHS = HS[j]; NHL = HL[j];
            /* F: vertical, opened from the cell above iff strictly better (A.5) */
            int F_opn = NH - o, F_ext = F[j] - e;
            if (F_opn > F_ext || (R.open_on_tie && F_opn == F_ext)) { F[j] = F_opn; FM[j] = NHM; FS[j] = NHS; FL[j] = NHL + 1; tflag |= T_DIAG_F; }
            else               { F[j] = F_ext;                            FL[j] = FL[j] + 1; tflag |= T_DEL_F; }
            /* E: horizontal, opened from the cell to the left iff strictly better */
            int E_opn = WH - o, E_ext = E - e;
            if (E_opn > E_ext || (R.open_on_tie && E_opn == E_ext)) { E = E_opn; EM = WHM; ES = WHS; EL = WHL + 1; tflag |= T_DIAG_E; }
            else               { E = E_ext;                     EL = EL + 1;  tflag |= T_INS_E; }
            int sub = matrow[s2[j - 1]];
            int H_dag = NWH + sub;
            /* H source priority (rule h_priority): diagonal >= both, else F if F >= E, else E */
            int take_diag, take_f;
            switch (R.h_priority) {
                case 1: take_diag = H_dag >= E && H_dag >= F[j]; take_f = F[j] > E; break;
                case 2: take_diag = H_dag > E && H_dag > F[j]; take_f = F[j] >= E; break;
                case 3: take_diag = H_dag > E && H_dag > F[j]; take_f = F[j] > E; break;
                default: take_diag = H_dag >= E && H_dag >= F[j]; take_f = F[j] >= E; break;
            }
            if (take_diag) {
                WH = H_dag;
                WHM = NWM + (R.match_raw_bytes ? (q && q[i - 1] == r[j - 1]) : (s1[i - 1] == s2[j - 1]));
                WHS = NWS + (sub > 0);
                WHL = NWL + 1;
                tflag |= T_DIAG;
            } else if (take_f) {
                WH = F[j]; WHM = FM[j]; WHS = FS[j]; WHL = FL[j]; tflag |= T_DEL;
            } else {
                WH = E; WHM = EM; WHS = ES; WHL = EL; tflag |= T_INS;
            }
            if (banded && (j - i < band_lo || j - i > band_hi)) { /* outside the band: unreachable */
                WH = PSBO_NEG_INF; E = PSBO_NEG_INF; F[j] = PSBO_NEG_INF; WHM = WHS = WHL = 0; tflag = 0;
            }
            if (mode == PSBO_SW && (WH < 0 || (WH == 0 && (R.zero_beats_diag || !(tflag & T_DIAG))))) { /* ZERO wins, stats reset */
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
                /* A.4: max score, then smallest end_ref, then smallest end_query (rule sw_end_tie) */
                if (WH > score || (WH == score && (R.sw_end_tie == 0 ? (j - 1) < end_ref : R.sw_end_tie == 2))) {
                    score = WH; end_query = i - 1; end_ref = j - 1;
                    matches = WHM; similar = WHS; length = WHL;
                }
            } else if (mode == PSBO_SG) {
                if (s1_end && i == qlen && (WH > score || (R.sg_row_last_wins && WH == score))) { /* last row, left to right, strict */
                    score = WH; end_query = i - 1; end_ref = j - 1;
                    matches = WHM; similar = WHS; length = WHL;
                }
                if (s2_end && j == rlen && (WH > col_score || (R.sg_row_last_wins && WH == col_score))) { /* last column, top to bottom, strict */
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
        if (!s1_end || col_score > score || (R.sg_col_wins_tie && col_score == score)) {
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
               uint32_t *ops_out, int *beg_query, int *beg_ref, int mode)
{
    (void)qlen;
    const psbo_rules_t R = g_rules;
    const uint32_t OP_I = R.cigar_swap_id ? 2 : 1, OP_D = R.cigar_swap_id ? 1 : 2;
    int64_t i = end_query, j = end_ref;
    int where = T_DIAG;
    uint32_t *rev = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(end_query + end_ref + 4));
    int nrev = 0;
    int64_t cur_op = -1; uint32_t cur_len = 0;
#define EMIT(op_) do { \
        if (cur_op == (op_)) { cur_len++; } \
        else { \
            if (cur_op >= 0) { rev[nrev++] = (cur_len << 4) | (uint32_t)cur_op; } \
            cur_op = (op_); cur_len = 1; \
        } } while (0)
    while (i >= 0 || j >= 0) {
        if ((i < 0 || j < 0) && R.cigar_edge_stop && mode != PSBO_NW) break;
        if (i < 0) { EMIT(OP_D); --j; continue; }
        if (j < 0) { EMIT(OP_I); --i; continue; }
        int t = trace[(size_t)i * (size_t)rlen + (size_t)j];
        if (where == T_DIAG) {
            if (t & T_DIAG) {
                int a = mat->mapper[q[i]], b = mat->mapper[r[j]];
                if (R.match_raw_bytes) { a = q[i]; b = r[j]; }
                EMIT(a == b ? 7 : 8);
                --i; --j;
            } else if (t & T_INS) { where = T_INS; }
            else if (t & T_DEL) { where = T_DEL; }
            else break; /* ZERO */
        } else if (where == T_INS) { /* E: horizontal, consumes reference */
            EMIT(OP_D);
            where = (t & T_DIAG_E) ? T_DIAG : T_INS;
            --j;
        } else { /* F: vertical, consumes query */
            EMIT(OP_I);
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
                   char match, char pos, char neg, char *qs, char *cs, char *rs, int mode)
{
    (void)qlen;
    const psbo_rules_t R = g_rules;
    int64_t i = end_query, j = end_ref;
    int where = T_DIAG, n = 0;
    while (i >= 0 || j >= 0) {
        if ((i < 0 || j < 0) && R.cigar_edge_stop && mode != PSBO_NW) break;
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
typedef struct psbo_batch_job {
    const uint8_t *qcat; const int64_t *qoff; const uint8_t *rcat; const int64_t *roff;
    int64_t n; int shared_query; const psbo_config_t *cfg; const psbo_matrix_t *mat;
    int *score, *end_query, *end_ref, *matches, *similar, *length;
    int want_cigar; uint32_t *cig_ops; int64_t cig_cap; int *nops_of; int *beg_query, *beg_ref;
    int64_t next;   /* shared work counter */
    int failed;
} psbo_batch_job_t;

static int64_t batch_slot(const psbo_batch_job_t *J, int64_t p, int qlen)
{
    return (J->shared_query ? (int64_t)qlen * p : J->qoff[p] - J->qoff[0]) + (J->roff[p] - J->roff[0]);
}

static void *batch_worker(void *arg)
{
    psbo_batch_job_t *J = (psbo_batch_job_t *)arg;
    for (;;) {
        const int64_t p0 = __atomic_fetch_add(&J->next, 8, __ATOMIC_RELAXED);
        if (p0 >= J->n || __atomic_load_n(&J->failed, __ATOMIC_RELAXED)) break;
        for (int64_t p = p0; p < p0 + 8 && p < J->n; ++p) {
            const uint8_t *q = J->shared_query ? J->qcat : J->qcat + J->qoff[p];
            int qlen = (int)(J->shared_query ? J->qoff[1] - J->qoff[0] : J->qoff[p + 1] - J->qoff[p]);
            const uint8_t *r = J->rcat + J->roff[p];
            int rlen = (int)(J->roff[p + 1] - J->roff[p]);
            psbo_out_t o; memset(&o, 0, sizeof(o));
            int8_t *trace = NULL;
            if (J->want_cigar) {
                trace = (int8_t *)malloc((size_t)qlen * (size_t)rlen); o.trace = trace;
                if (!trace) { __atomic_store_n(&J->failed, 1, __ATOMIC_RELAXED); break; }
            }
            if (psbo_align(q, qlen, r, rlen, J->cfg, J->mat, &o) != 0) { free(trace); __atomic_store_n(&J->failed, 1, __ATOMIC_RELAXED); break; }
            J->score[p] = o.score; J->end_query[p] = o.end_query; J->end_ref[p] = o.end_ref;
            if (J->matches) J->matches[p] = o.matches;
            if (J->similar) J->similar[p] = o.similar;
            if (J->length) J->length[p] = o.length;
            if (J->want_cigar) {
                const int64_t slot = batch_slot(J, p, qlen);
                if (slot + qlen + rlen > J->cig_cap) { free(trace); __atomic_store_n(&J->failed, 1, __ATOMIC_RELAXED); break; }
                int bq, br;
                J->nops_of[p] = psbo_cigar(trace, q, qlen, r, rlen, J->mat, o.end_query, o.end_ref, J->cig_ops + slot, &bq, &br, J->cfg->mode);
                J->beg_query[p] = bq; J->beg_ref[p] = br;
                free(trace);
            }
        }
    }
    return NULL;
}

static int g_threads = 1;
/* worker threads of psbo_align_batch (pairs are independent; results do not depend on the count) */
void psbo_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }

int psbo_align_batch(const uint8_t *qcat, const int64_t *qoff, const uint8_t *rcat, const int64_t *roff,
                     int64_t n, int shared_query, const psbo_config_t *cfg, const psbo_matrix_t *mat,
                     int *score, int *end_query, int *end_ref, int *matches, int *similar, int *length,
                     int want_cigar, uint32_t *cig_ops, int64_t cig_cap, int64_t *cig_off,
                     int *beg_query, int *beg_ref)
{
    /* with want_cigar every pair first writes its ops at a private slot of cig_ops (slot p starts at
     * the sum of the lengths before it, so slots cannot overlap), then the slots are compacted */
    psbo_batch_job_t J;
    memset(&J, 0, sizeof(J));
    J.qcat = qcat; J.qoff = qoff; J.rcat = rcat; J.roff = roff; J.n = n; J.shared_query = shared_query;
    J.cfg = cfg; J.mat = mat; J.score = score; J.end_query = end_query; J.end_ref = end_ref;
    J.matches = matches; J.similar = similar; J.length = length;
    J.want_cigar = want_cigar; J.cig_ops = cig_ops; J.cig_cap = cig_cap; J.beg_query = beg_query; J.beg_ref = beg_ref;
    J.nops_of = want_cigar ? (int *)calloc((size_t)n + 1, sizeof(int)) : NULL;
    if (want_cigar && !J.nops_of) return -1;
    int nt = g_threads;
    if ((int64_t)nt > (n + 7) / 8) nt = (int)((n + 7) / 8);
    if (nt <= 1) batch_worker(&J);
    else {
        pthread_t th[256];
        int started = 0;
        for (int t = 0; t < nt; ++t) if (pthread_create(&th[started], NULL, batch_worker, &J) == 0) ++started;
        if (started == 0) batch_worker(&J);
        for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    }
    if (J.failed) { free(J.nops_of); return -1; }
    if (want_cigar) {
        int64_t used = 0;
        cig_off[0] = 0;
        for (int64_t p = 0; p < n; ++p) {
            const int qlen = (int)(shared_query ? qoff[1] - qoff[0] : qoff[p + 1] - qoff[p]);
            memmove(cig_ops + used, cig_ops + batch_slot(&J, p, qlen), sizeof(uint32_t) * (size_t)J.nops_of[p]);
            used += J.nops_of[p];
            cig_off[p + 1] = used;
        }
        free(J.nops_of);
    }
    return 0;
}

/*
 * Oracle-free property of a batch of CIGARs (SURVEY 8d "CIGAR re-score / recount on all pairs"): walking
 * each CIGAR from (beg_query, beg_ref) must consume exactly up to (end_query, end_ref), every '=' must sit
 * on equal mapped residues and every 'X' on unequal ones, and -- for sw and nw -- the substitution scores
 * minus the affine gap costs must add up to the reported score.  Returns the number of pairs that fail
 * (first failing pair in *first_bad), or -1 on bad arguments.  Threaded like psbo_align_batch.
 */
typedef struct psbo_check_job {
    const uint8_t *qcat; const int64_t *qoff; const uint8_t *rcat; const int64_t *roff; int64_t n;
    const psbo_config_t *cfg; const psbo_matrix_t *mat;
    const uint32_t *ops; const int64_t *off; const int *bq, *br, *eq, *er, *score;
    int64_t next, bad, first_bad;
} psbo_check_job_t;

static void *check_worker(void *arg)
{
    psbo_check_job_t *J = (psbo_check_job_t *)arg;
    const int o = J->cfg->open, e = J->cfg->gap;
    for (;;) {
        const int64_t p0 = __atomic_fetch_add(&J->next, 64, __ATOMIC_RELAXED);
        if (p0 >= J->n) break;
        for (int64_t p = p0; p < p0 + 64 && p < J->n; ++p) {
            const uint8_t *q = J->qcat + J->qoff[p], *r = J->rcat + J->roff[p];
            const int64_t qlen = J->qoff[p + 1] - J->qoff[p], rlen = J->roff[p + 1] - J->roff[p];
            int64_t i = J->bq[p], j = J->br[p], sc = 0;
            int ok = 1;
            for (int64_t k = J->off[p]; k < J->off[p + 1] && ok; ++k) {
                const int op = (int)(J->ops[k] & 0xf);
                const int64_t len = J->ops[k] >> 4;
                if (op == 7 || op == 8) {
                    for (int64_t t = 0; t < len; ++t, ++i, ++j) {
                        if (i >= qlen || j >= rlen) { ok = 0; break; }
                        const int a = J->mat->mapper[q[i]], b = J->mat->mapper[r[j]];
                        if ((a == b) != (op == 7)) { ok = 0; break; }
                        sc += J->mat->matrix[(size_t)J->mat->size * (J->mat->is_pssm ? (int)i : a) + b];
                    }
                } else if (op == 1 || op == 2) {
                    /* a local alignment that reaches the table's edge carries the rest of the other sequence as
                     * one leading I / D run (walk rule cigar_edge_stop = 0): that run is outside the score */
                    const int free_run = J->cfg->mode == PSBO_SW && k == J->off[p];
                    if (op == 1) i += len; else j += len;
                    if (!free_run) sc -= o + (len - 1) * e;
                } else ok = 0;
            }
            if (ok && (i - 1 != J->eq[p] || j - 1 != J->er[p]) && !(J->off[p] == J->off[p + 1])) ok = 0;
            if (ok && (J->cfg->mode == PSBO_SW || J->cfg->mode == PSBO_NW) && sc != J->score[p]) ok = 0;
            if (!ok) {
                __atomic_fetch_add(&J->bad, 1, __ATOMIC_RELAXED);
                int64_t cur = __atomic_load_n(&J->first_bad, __ATOMIC_RELAXED);
                while ((cur < 0 || p < cur) && !__atomic_compare_exchange_n(&J->first_bad, &cur, p, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
            }
        }
    }
    return NULL;
}

int64_t psbo_cigar_check_batch(const uint8_t *qcat, const int64_t *qoff, const uint8_t *rcat, const int64_t *roff, int64_t n,
                               const psbo_config_t *cfg, const psbo_matrix_t *mat, const uint32_t *ops, const int64_t *off,
                               const int *beg_query, const int *beg_ref, const int *end_query, const int *end_ref,
                               const int *score, int64_t *first_bad)
{
    if (!qcat || !rcat || !ops || !off || n <= 0) return -1;
    psbo_check_job_t J;
    memset(&J, 0, sizeof(J));
    J.qcat = qcat; J.qoff = qoff; J.rcat = rcat; J.roff = roff; J.n = n; J.cfg = cfg; J.mat = mat;
    J.ops = ops; J.off = off; J.bq = beg_query; J.br = beg_ref; J.eq = end_query; J.er = end_ref; J.score = score;
    J.first_bad = -1;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < g_threads; ++t) if (pthread_create(&th[started], NULL, check_worker, &J) == 0) ++started;
    if (started == 0) check_worker(&J);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    if (first_bad) *first_bad = J.first_bad;
    return J.bad;
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

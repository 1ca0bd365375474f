// psb_defs.h -- constants shared by the kernels and the host engine.
#pragma once
#include <stdint.h>

namespace psb {

enum { MODE_NW = 0, MODE_SG = 1, MODE_SW = 2 };
static constexpr int NEG_INF32 = -(1 << 30);
static constexpr int PAD_SCORE = -(1 << 28);  // substitution score of rows beyond the query

// TraceFlags bytes [REF src/alignment/table.rs:127-142]
enum { TR_INS = 1, TR_DEL = 2, TR_DIAG = 4, TR_DIAG_E = 8, TR_INS_E = 16, TR_DIAG_F = 32, TR_DEL_F = 64 };

// ---- upstream-behaviour assumptions ([UP] in SURVEY.md Appendix A) ---------------------------------
// The parasail C sources are not available here, so every rule that is recalled rather than read has
// ONE named constant; the kernels and the host walk refer to these instead of hard-coding the choice.
// oracle/gotoh_oracle.c carries the same rules as run-time switches (psbo_rules_t, same names) and
// oracle/UP_ASSUMPTIONS.md lists confidence, evidence and what flipping each one entails.
namespace rules {
// A.4: among equal local maxima the smaller end_ref wins, then the smaller end_query
constexpr bool SW_END_PREFERS_SMALLER_REF = true;
// A.4: sg -- the last column beats the last row only when strictly greater; first maximum wins inside each
constexpr bool SG_COL_WINS_TIE = false;
// A.5: H source priority  diag >= F (vertical) >= E (horizontal)
constexpr bool H_DIAG_WINS_TIE = true;
constexpr bool H_F_WINS_TIE_OVER_E = true;
// A.5: a gap is opened only when strictly better than extending it
constexpr bool GAP_OPEN_ON_TIE = false;
// A.6: a match is equality of mapped matrix indices; boundary gaps are not counted in the statistics
constexpr bool MATCH_ON_MAPPED_INDEX = true;
constexpr bool COUNT_BOUNDARY_GAPS = false;
// A.7: the walk runs to (-1,-1) in every mode (the remainder becomes a leading I / D run); 'I' consumes the
// query (F state), 'D' consumes the reference (E state); BAM op numbers
constexpr bool CIGAR_WALK_TO_ORIGIN = true;
constexpr unsigned CIGAR_OP_I = 1, CIGAR_OP_D = 2, CIGAR_OP_EQ = 7, CIGAR_OP_X = 8;
}  // namespace rules

}  // namespace psb

// psb_defs.h -- constants shared by the kernels and the host engine.
#pragma once
#include <stdint.h>

namespace psb {

enum { MODE_NW = 0, MODE_SG = 1, MODE_SW = 2 };
static constexpr int NEG_INF32 = -(1 << 30);
static constexpr int PAD_SCORE = -(1 << 28);  // substitution score of rows beyond the query

// TraceFlags bytes [REF src/alignment/table.rs:127-142]
enum { TR_INS = 1, TR_DEL = 2, TR_DIAG = 4, TR_DIAG_E = 8, TR_INS_E = 16, TR_DIAG_F = 32, TR_DEL_F = 64 };

}  // namespace psb

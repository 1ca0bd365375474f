// fn_name.cpp -- parasail's function-name grammar and the lookup tables behind
// parasail_lookup_function / parasail_lookup_pfunction [REF src/aligner/mod.rs:289-358].
//
// parasail-rs turns builder state into a name such as "sg_qb_de_stats_rowcol_scan_profile_16"
// and calls whatever pointer the lookup returns with a fixed 7-argument (or 5-argument profile)
// signature that carries no configuration.  Each valid name therefore needs its own entry
// point: they are stamped out at compile time as thunk<ID>, where ID encodes the parsed name,
// and every thunk funnels into the same GPU batch path with n = 1.
#include <array>
#include <cstring>
#include <string>
#include <utility>

#include "psb_internal.h"

namespace psb {

namespace {
// the 13 alignment classes upstream defines (sg_qx_dx is spelled "sg"; other q/d pairings
// such as qx_db do not exist upstream, so lookup fails for them just like upstream)
struct ModeDef { const char *name; int mode, s1b, s1e, s2b, s2e; };
const ModeDef kModes[] = {
    {"nw", 0, 0, 0, 0, 0},       {"sg", 1, 1, 1, 1, 1},       {"sg_qb", 1, 1, 0, 0, 0},   {"sg_qe", 1, 0, 1, 0, 0},
    {"sg_qx", 1, 1, 1, 0, 0},    {"sg_db", 1, 0, 0, 1, 0},    {"sg_de", 1, 0, 0, 0, 1},   {"sg_dx", 1, 0, 0, 1, 1},
    {"sg_qb_de", 1, 1, 0, 0, 1}, {"sg_qe_db", 1, 0, 1, 1, 0}, {"sg_qb_db", 1, 1, 0, 1, 0}, {"sg_qe_de", 1, 0, 1, 0, 1},
    {"sw", 2, 0, 0, 0, 0}};
constexpr int kNumModes = 13;
// output classes: plain, table, rowcol, stats, stats_table, stats_rowcol, trace
constexpr int kNumOut = 7;
constexpr int kNumStrategy = 3;  // striped, scan, diag
constexpr int kNumWidth = 5;     // 8, 16, 32, 64, sat
const int kWidths[kNumWidth] = {8, 16, 32, 64, 0};
constexpr int kNumIds = kNumModes * kNumOut * kNumStrategy * kNumWidth;

int mode_index(const FnConfig &c) {
    for (int i = 0; i < kNumModes; ++i) {
        const ModeDef &m = kModes[i];
        if (m.mode == c.mode && (c.mode != 1 || (m.s1b == c.s1_beg && m.s1e == c.s1_end && m.s2b == c.s2_beg && m.s2e == c.s2_end)))
            return i;
    }
    return -1;
}
int out_index(const FnConfig &c) {
    if (c.trace) return 6;
    return (c.stats ? 3 : 0) + (c.table ? 1 : (c.rowcol ? 2 : 0));
}
bool eat(const char *&p, const char *tok) {
    const size_t n = std::strlen(tok);
    if (std::strncmp(p, tok, n) == 0) { p += n; return true; }
    return false;
}
}  // namespace

int fn_id_count() { return kNumIds; }

int encode_fn(const FnConfig &c) {
    int w = 4;
    for (int i = 0; i < kNumWidth; ++i) if (kWidths[i] == c.width) w = i;
    return ((mode_index(c) * kNumOut + out_index(c)) * kNumStrategy + c.strategy) * kNumWidth + w;
}

FnConfig decode_fn(int id) {
    FnConfig c;
    c.width = kWidths[id % kNumWidth]; id /= kNumWidth;
    c.strategy = id % kNumStrategy; id /= kNumStrategy;
    const int out = id % kNumOut; id /= kNumOut;
    const ModeDef &m = kModes[id];
    c.mode = m.mode; c.s1_beg = m.s1b; c.s1_end = m.s1e; c.s2_beg = m.s2b; c.s2_end = m.s2e;
    c.trace = out == 6;
    c.stats = out >= 3 && out <= 5;
    c.table = out == 1 || out == 4;
    c.rowcol = out == 2 || out == 5;
    return c;
}

int FnConfig::flag() const {
    int f = mode == 0 ? PARASAIL_FLAG_NW : (mode == 1 ? PARASAIL_FLAG_SG : PARASAIL_FLAG_SW);
    if (mode == 1) {
        if (s1_beg) f |= PARASAIL_FLAG_SG_S1_BEG;
        if (s1_end) f |= PARASAIL_FLAG_SG_S1_END;
        if (s2_beg) f |= PARASAIL_FLAG_SG_S2_BEG;
        if (s2_end) f |= PARASAIL_FLAG_SG_S2_END;
    }
    f |= strategy == 0 ? PARASAIL_FLAG_STRIPED : (strategy == 1 ? PARASAIL_FLAG_SCAN : PARASAIL_FLAG_DIAG);
    if (stats) f |= PARASAIL_FLAG_STATS;
    if (table) f |= PARASAIL_FLAG_TABLE;
    if (rowcol) f |= PARASAIL_FLAG_ROWCOL;
    if (trace) f |= PARASAIL_FLAG_TRACE;
    // "sat" reports the narrowest width that holds the result; filled in by the caller
    if (width == 8) f |= PARASAIL_FLAG_BITS_8;
    if (width == 16) f |= PARASAIL_FLAG_BITS_16;
    if (width == 32) f |= PARASAIL_FLAG_BITS_32;
    if (width == 64) f |= PARASAIL_FLAG_BITS_64;
    return f;
}

// {mode}{_q?}{_d?}{_trace}{_stats}{_table|_rowcol}{_striped|_scan|_diag}{_profile}_{width}
bool parse_fn_name(const char *name, FnConfig *out) {
    if (!name) return false;
    const char *p = name;
    eat(p, "parasail_");
    FnConfig c;
    if (eat(p, "nw")) c.mode = 0;
    else if (eat(p, "sg")) c.mode = 1;
    else if (eat(p, "sw")) c.mode = 2;
    else return false;
    if (c.mode == 1) {
        bool q = false, d = false;
        if (eat(p, "_qb")) { c.s1_beg = 1; q = true; }
        else if (eat(p, "_qe")) { c.s1_end = 1; q = true; }
        else if (eat(p, "_qx")) { c.s1_beg = c.s1_end = 1; q = true; }
        if (eat(p, "_db")) { c.s2_beg = 1; d = true; }
        else if (eat(p, "_de")) { c.s2_end = 1; d = true; }
        else if (eat(p, "_dx")) { c.s2_beg = c.s2_end = 1; d = true; }
        if (!q && !d) c.s1_beg = c.s1_end = c.s2_beg = c.s2_end = 1;
        if (mode_index(c) < 0 || (q && d && (c.s1_beg + c.s1_end == 2 || c.s2_beg + c.s2_end == 2))) return false;
    }
    if (eat(p, "_trace")) c.trace = true;
    if (eat(p, "_stats")) c.stats = true;
    if (eat(p, "_table")) c.table = true;
    else if (eat(p, "_rowcol")) c.rowcol = true;
    if (c.trace && (c.stats || c.table || c.rowcol)) return false;  // upstream has no such kernels
    if (eat(p, "_striped")) c.strategy = 0;
    else if (eat(p, "_scan")) c.strategy = 1;
    else if (eat(p, "_diag")) c.strategy = 2;
    else return false;
    if (eat(p, "_profile")) c.profile = true;
    if (c.profile && c.strategy == 2) return false;  // no diag profile kernels upstream
    if (eat(p, "_sat")) c.width = 0;
    else if (eat(p, "_8")) c.width = 8;
    else if (eat(p, "_16")) c.width = 16;
    else if (eat(p, "_32")) c.width = 32;
    else if (eat(p, "_64")) c.width = 64;
    else return false;
    if (*p != '\0') return false;
    *out = c;
    return true;
}

namespace {

parasail_result_t *dispatch(int id, const char *s1, int s1Len, const char *s2, int s2Len, int open, int gap,
                            const parasail_matrix_t *matrix) {
    FnConfig cfg = decode_fn(id);
    if (!matrix) {
        set_error("alignment function called with a NULL matrix");
        HostMatrix none;
        return align_one(cfg, none, nullptr, 0, nullptr, 0, open, gap);
    }
    HostMatrix hm(matrix);
    return align_one(cfg, hm, (const uint8_t *)s1, s1Len, (const uint8_t *)s2, s2Len, open, gap);
}

parasail_result_t *pdispatch(int id, const parasail_profile_t *profile, const char *s2, int s2Len, int open, int gap) {
    FnConfig cfg = decode_fn(id);
    cfg.profile = true;
    if (!profile) {
        set_error("profile alignment function called with a NULL profile");
        HostMatrix none;
        return align_one(cfg, none, nullptr, 0, nullptr, 0, open, gap);
    }
    // parasail-rs takes the _stats part of the name from the profile's flag
    // [REF src/aligner/mod.rs:303-317]; upstream computes what the name says, so do we.
    return align_one(cfg, profile->matrix, profile->query.data(), (int)profile->query.size(), (const uint8_t *)s2,
                     s2Len, open, gap);
}

template <int ID>
parasail_result_t *thunk(const char *s1, int s1Len, const char *s2, int s2Len, int open, int gap,
                         const parasail_matrix_t *matrix) {
    return dispatch(ID, s1, s1Len, s2, s2Len, open, gap, matrix);
}
template <int ID>
parasail_result_t *pthunk(const parasail_profile_t *profile, const char *s2, int s2Len, int open, int gap) {
    return pdispatch(ID, profile, s2, s2Len, open, gap);
}

template <size_t... I>
constexpr std::array<parasail_function_t *, sizeof...(I)> make_table(std::index_sequence<I...>) {
    return {{&thunk<(int)I>...}};
}
template <size_t... I>
constexpr std::array<parasail_pfunction_t *, sizeof...(I)> make_ptable(std::index_sequence<I...>) {
    return {{&pthunk<(int)I>...}};
}
const auto kTable = make_table(std::make_index_sequence<kNumIds>{});
const auto kPTable = make_ptable(std::make_index_sequence<kNumIds>{});

}  // namespace
}  // namespace psb

extern "C" {

parasail_function_t *parasail_lookup_function(const char *funcname) {
    psb::FnConfig c;
    if (!psb::parse_fn_name(funcname, &c) || c.profile) return nullptr;
    return psb::kTable[psb::encode_fn(c)];
}

parasail_pfunction_t *parasail_lookup_pfunction(const char *funcname) {
    psb::FnConfig c;
    if (!psb::parse_fn_name(funcname, &c) || !c.profile) return nullptr;
    return psb::kPTable[psb::encode_fn(c)];
}

// ---- profiles [REF src/profile/mod.rs:93-103, 306-333, 384-390] ---------------------------
static parasail_profile_t *make_profile(const char *s1, int s1Len, const parasail_matrix_t *m, bool stats, int width) {
    if (!s1 || s1Len <= 0 || !m) { psb::set_error("parasail_profile_create: empty query or NULL matrix"); return nullptr; }
    parasail_profile *p = new (std::nothrow) parasail_profile();
    if (!p) return nullptr;
    p->query.assign((const uint8_t *)s1, (const uint8_t *)s1 + s1Len);
    p->matrix = psb::HostMatrix(m);
    p->stats = stats;
    p->width = width;
    return p;
}

#define PSB_DEF_PROFILE_CREATORS(ISA)                                                                            \
    parasail_profile_t *parasail_profile_create##ISA##_8(const char *s, int n, const parasail_matrix_t *m) { return make_profile(s, n, m, false, 8); }    \
    parasail_profile_t *parasail_profile_create##ISA##_16(const char *s, int n, const parasail_matrix_t *m) { return make_profile(s, n, m, false, 16); }  \
    parasail_profile_t *parasail_profile_create##ISA##_32(const char *s, int n, const parasail_matrix_t *m) { return make_profile(s, n, m, false, 32); }  \
    parasail_profile_t *parasail_profile_create##ISA##_64(const char *s, int n, const parasail_matrix_t *m) { return make_profile(s, n, m, false, 64); }  \
    parasail_profile_t *parasail_profile_create##ISA##_sat(const char *s, int n, const parasail_matrix_t *m) { return make_profile(s, n, m, false, 0); }  \
    parasail_profile_t *parasail_profile_create_stats##ISA##_8(const char *s, int n, const parasail_matrix_t *m) { return make_profile(s, n, m, true, 8); }    \
    parasail_profile_t *parasail_profile_create_stats##ISA##_16(const char *s, int n, const parasail_matrix_t *m) { return make_profile(s, n, m, true, 16); }  \
    parasail_profile_t *parasail_profile_create_stats##ISA##_32(const char *s, int n, const parasail_matrix_t *m) { return make_profile(s, n, m, true, 32); }  \
    parasail_profile_t *parasail_profile_create_stats##ISA##_64(const char *s, int n, const parasail_matrix_t *m) { return make_profile(s, n, m, true, 64); }  \
    parasail_profile_t *parasail_profile_create_stats##ISA##_sat(const char *s, int n, const parasail_matrix_t *m) { return make_profile(s, n, m, true, 0); }
PSB_DEF_PROFILE_CREATORS()
PSB_DEF_PROFILE_CREATORS(_sse_128)
PSB_DEF_PROFILE_CREATORS(_avx_256)
PSB_DEF_PROFILE_CREATORS(_neon_128)
PSB_DEF_PROFILE_CREATORS(_altivec_128)
#undef PSB_DEF_PROFILE_CREATORS

void parasail_profile_free(parasail_profile_t *profile) {
    if (!profile) return;
    psb::release_profile_resident(profile);
    delete profile;
}

}  // extern "C"

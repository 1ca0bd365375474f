// matrix.cpp -- substitution matrices behind parasail-rs's Matrix [REF src/matrix/mod.rs:25-307].
// Host-side objects only; the engine snapshots them (psb::HostMatrix) and stages the values in
// GPU shared memory per launch.  Struct layout: include/parasail_b200.h.
#include <algorithm>
#include <cctype>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "psb_internal.h"

namespace {

// BLOSUM62, order ARNDCQEGHILKMFPSTWYVBZX* (the only built-in table whose
// values exist anywhere in this build environment -- others are not fabricated).
const char kBlosum62Alphabet[] = "ARNDCQEGHILKMFPSTWYVBZX*";
const int kBlosum62[24 * 24] = {
    4,  -1, -2, -2, 0,  -1, -1, 0,  -2, -1, -1, -1, -1, -2, -1, 1,  0,  -3, -2, 0,  -2, -1, 0,  -4,
    -1, 5,  0,  -2, -3, 1,  0,  -2, 0,  -3, -2, 2,  -1, -3, -2, -1, -1, -3, -2, -3, -1, 0,  -1, -4,
    -2, 0,  6,  1,  -3, 0,  0,  0,  1,  -3, -3, 0,  -2, -3, -2, 1,  0,  -4, -2, -3, 3,  0,  -1, -4,
    -2, -2, 1,  6,  -3, 0,  2,  -1, -1, -3, -4, -1, -3, -3, -1, 0,  -1, -4, -3, -3, 4,  1,  -1, -4,
    0,  -3, -3, -3, 9,  -3, -4, -3, -3, -1, -1, -3, -1, -2, -3, -1, -1, -2, -2, -1, -3, -3, -2, -4,
    -1, 1,  0,  0,  -3, 5,  2,  -2, 0,  -3, -2, 1,  0,  -3, -1, 0,  -1, -2, -1, -2, 0,  3,  -1, -4,
    -1, 0,  0,  2,  -4, 2,  5,  -2, 0,  -3, -3, 1,  -2, -3, -1, 0,  -1, -3, -2, -2, 1,  4,  -1, -4,
    0,  -2, 0,  -1, -3, -2, -2, 6,  -2, -4, -4, -2, -3, -3, -2, 0,  -2, -2, -3, -3, -1, -2, -1, -4,
    -2, 0,  1,  -1, -3, 0,  0,  -2, 8,  -3, -3, -1, -2, -1, -2, -1, -2, -2, 2,  -3, 0,  0,  -1, -4,
    -1, -3, -3, -3, -1, -3, -3, -4, -3, 4,  2,  -3, 1,  0,  -3, -2, -1, -3, -1, 3,  -3, -3, -1, -4,
    -1, -2, -3, -4, -1, -2, -3, -4, -3, 2,  4,  -2, 2,  0,  -3, -2, -1, -2, -1, 1,  -4, -3, -1, -4,
    -1, 2,  0,  -1, -3, 1,  1,  -2, -1, -3, -2, 5,  -1, -3, -1, 0,  -1, -3, -2, -2, 0,  1,  -1, -4,
    -1, -1, -2, -3, -1, 0,  -2, -3, -2, 1,  2,  -1, 5,  0,  -2, -1, -1, -1, -1, 1,  -3, -1, -1, -4,
    -2, -3, -3, -3, -2, -3, -3, -3, -1, 0,  0,  -3, 0,  6,  -4, -2, -2, 1,  3,  -1, -3, -3, -1, -4,
    -1, -2, -2, -1, -3, -1, -1, -2, -2, -3, -3, -1, -2, -4, 7,  -1, -1, -4, -3, -2, -2, -1, -2, -4,
    1,  -1, 1,  0,  -1, 0,  0,  0,  -1, -2, -2, 0,  -1, -2, -1, 4,  1,  -3, -2, -2, 0,  0,  0,  -4,
    0,  -1, 0,  -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1, 1,  5,  -2, -2, 0,  -1, -1, 0,  -4,
    -3, -3, -4, -4, -2, -2, -3, -2, -2, -3, -2, -3, -1, 1,  -4, -3, -2, 11, 2,  -3, -4, -3, -2, -4,
    -2, -2, -2, -3, -2, -1, -2, -3, 2,  -1, -1, -2, -1, 3,  -3, -2, -2, 2,  7,  -1, -3, -2, -1, -4,
    0,  -3, -3, -3, -1, -2, -2, -3, -3, 3,  1,  -2, 1,  -1, -2, -2, 0,  -3, -1, 4,  -3, -2, -1, -4,
    -2, -1, 3,  4,  -3, 0,  1,  -1, 0,  -3, -4, 0,  -3, -3, -2, 0,  -1, -4, -3, -3, 4,  1,  -1, -4,
    -1, 0,  0,  1,  -3, 3,  4,  -2, 0,  -3, -3, 1,  -1, -3, -1, 0,  -1, -3, -2, -2, 1,  4,  -1, -4,
    0,  -1, -1, -1, -2, -1, -1, -1, -1, -1, -1, -1, -1, -1, -2, 0,  0,  -2, -1, -1, -1, -1, -1, -4,
    -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, 1};

int g_blosum62_mapper[256];
parasail_matrix_t g_blosum62;
std::once_flag g_blosum62_once;

void fill_mapper(int *mapper, const char *alphabet, int size) {
    for (int i = 0; i < 256; ++i) mapper[i] = size - 1;
    for (int i = 0; alphabet[i]; ++i) {
        const unsigned char c = (unsigned char)alphabet[i];
        mapper[(unsigned char)std::toupper(c)] = i;
        mapper[(unsigned char)std::tolower(c)] = i;
    }
}

char *dup_cstr(const std::string &s) {
    char *p = (char *)std::malloc(s.size() + 1);
    if (p) std::memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

// heap matrix: every pointer it holds is malloc'd and released by parasail_matrix_free
parasail_matrix_t *alloc_matrix(const std::string &name, const std::string &alphabet, int size, int length, int type) {
    parasail_matrix_t *m = (parasail_matrix_t *)std::calloc(1, sizeof(parasail_matrix_t));
    int *values = (int *)std::calloc((size_t)size * (size_t)length, sizeof(int));
    int *mapper = (int *)std::calloc(256, sizeof(int));
    if (!m || !values || !mapper) {
        std::free(m); std::free(values); std::free(mapper);
        return nullptr;
    }
    m->name = dup_cstr(name);
    m->matrix = values;
    m->user_matrix = values;
    m->mapper = mapper;
    m->size = size;
    m->length = length;
    m->type = type;
    m->alphabet = dup_cstr(alphabet);
    m->query = nullptr;
    return m;
}

void refresh_minmax(parasail_matrix_t *m) {
    int lo = INT_MAX, hi = INT_MIN;
    const size_t n = (size_t)m->size * (size_t)m->length;
    for (size_t i = 0; i < n; ++i) { lo = std::min(lo, m->matrix[i]); hi = std::max(hi, m->matrix[i]); }
    m->min = lo; m->max = hi;
}

bool is_integer_token(const std::string &t) {
    if (t.empty()) return false;
    size_t i = (t[0] == '-' || t[0] == '+') ? 1 : 0;
    if (i >= t.size()) return false;
    for (; i < t.size(); ++i) if (!std::isdigit((unsigned char)t[i])) return false;
    return true;
}

}  // namespace

extern "C" {

// [REF src/matrix/mod.rs:40] size = len+1; diagonal match, off-diagonal mismatch, wildcard
// row/column 0; both cases of each letter map to it, anything else to the wildcard (A.1)
parasail_matrix_t *parasail_matrix_create(const char *alphabet, int match, int mismatch) {
    if (!alphabet) return nullptr;
    const int n = (int)std::strlen(alphabet);
    const int size = n + 1;
    parasail_matrix_t *m = alloc_matrix("", std::string(alphabet) + "*", size, size, PARASAIL_MATRIX_TYPE_SQUARE);
    if (!m) return nullptr;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) m->user_matrix[i * size + j] = (i == j) ? match : mismatch;
    fill_mapper((int *)m->mapper, alphabet, size);
    m->max = std::max(match, mismatch);
    m->min = std::min(match, mismatch);
    return m;
}

// [REF src/matrix/mod.rs:62] static built-ins; the Rust side never frees them (builtin = true)
const parasail_matrix_t *parasail_matrix_lookup(const char *matrixname) {
    if (!matrixname) return nullptr;
    std::string nm(matrixname);
    for (auto &c : nm) c = (char)std::tolower((unsigned char)c);
    if (nm != "blosum62") {
        // The other built-in tables of upstream (blosum*/pam*/dnafull/nuc44 ...) exist nowhere in this
        // environment and are not fabricated from memory.  The lookup is data-driven instead:
        // a file <PSB_MATRIX_DIR>/<name>[.txt|.mat] in NCBI format (what parasail_matrix_from_file parses)
        // is loaded once and then lives as long as the process, like a built-in.
        const char *dir = std::getenv("PSB_MATRIX_DIR");
        if (!dir || !*dir || nm.find('/') != std::string::npos || nm.find("..") != std::string::npos) return nullptr;
        static std::mutex mu;
        static std::map<std::string, const parasail_matrix_t *> cache;
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(nm);
        if (it != cache.end()) return it->second;
        const parasail_matrix_t *found = nullptr;
        for (const char *ext : {"", ".txt", ".mat"}) {
            const std::string path = std::string(dir) + "/" + nm + ext;
            parasail_matrix_t *m = parasail_matrix_from_file(path.c_str());
            if (m && m->type == PARASAIL_MATRIX_TYPE_SQUARE) { found = m; break; }
            if (m) parasail_matrix_free(m);
        }
        cache[nm] = found;
        return found;
    }
    std::call_once(g_blosum62_once, []() {
        fill_mapper(g_blosum62_mapper, kBlosum62Alphabet, 24);
        g_blosum62.name = "blosum62";
        g_blosum62.matrix = kBlosum62;
        g_blosum62.mapper = g_blosum62_mapper;
        g_blosum62.size = 24;
        g_blosum62.max = 11;
        g_blosum62.min = -4;
        g_blosum62.user_matrix = nullptr;
        g_blosum62.type = PARASAIL_MATRIX_TYPE_SQUARE;
        g_blosum62.length = 24;
        g_blosum62.alphabet = kBlosum62Alphabet;
        g_blosum62.query = nullptr;
    });
    return &g_blosum62;
}

// [REF src/matrix/mod.rs:140-147; format: tests/square.txt, tests/pssm.txt].  '#' lines are
// comments; the first data line is the alphabet.  Square files repeat the alphabet as the first
// column and end with a wildcard row/column; anything else is a PSSM (one row per query
// position, optional leading letter), which gains a wildcard column holding the table minimum.
parasail_matrix_t *parasail_matrix_from_file(const char *filename) {
    if (!filename) return nullptr;
    std::ifstream in(filename);
    if (!in) return nullptr;
    std::vector<std::vector<std::string>> rows;
    std::string line;
    while (std::getline(in, line)) {
        size_t p = line.find_first_not_of(" \t\r\n");
        if (p == std::string::npos || line[p] == '#') continue;
        std::istringstream ss(line);
        std::vector<std::string> toks;
        std::string t;
        while (ss >> t) toks.push_back(t);
        if (!toks.empty()) rows.push_back(toks);
    }
    if (rows.size() < 2) return nullptr;
    std::string alphabet;
    for (auto &t : rows[0]) {
        if (t.size() != 1) return nullptr;
        alphabet.push_back(t[0]);
    }
    const int na = (int)alphabet.size();
    const int nrows = (int)rows.size() - 1;
    bool square = nrows == na;
    for (int i = 0; square && i < nrows; ++i) {
        const auto &r = rows[i + 1];
        square = (int)r.size() == na + 1 && r[0].size() == 1 && r[0][0] == alphabet[i] && !is_integer_token(r[0]);
    }
    if (square) {
        parasail_matrix_t *m = alloc_matrix(filename, alphabet, na, na, PARASAIL_MATRIX_TYPE_SQUARE);
        if (!m) return nullptr;
        for (int i = 0; i < na; ++i)
            for (int j = 0; j < na; ++j) {
                if (!is_integer_token(rows[i + 1][j + 1])) { parasail_matrix_free(m); return nullptr; }
                m->user_matrix[i * na + j] = std::atoi(rows[i + 1][j + 1].c_str());
            }
        // the last alphabet entry is the wildcard: unknown bytes map to it
        fill_mapper((int *)m->mapper, alphabet.substr(0, na - 1).c_str(), na);
        ((int *)m->mapper)[(unsigned char)alphabet[na - 1]] = na - 1;
        refresh_minmax(m);
        return m;
    }
    std::vector<int> values;
    for (int i = 0; i < nrows; ++i) {
        const auto &r = rows[i + 1];
        size_t first = ((int)r.size() == na + 1) ? 1 : 0;
        if ((int)(r.size() - first) != na) return nullptr;
        for (size_t j = first; j < r.size(); ++j) {
            if (!is_integer_token(r[j])) return nullptr;
            values.push_back(std::atoi(r[j].c_str()));
        }
    }
    return parasail_matrix_pssm_create(alphabet.c_str(), values.data(), nrows);
}

// [REF src/matrix/mod.rs:158] values: length x strlen(alphabet), row = query position
parasail_matrix_t *parasail_matrix_pssm_create(const char *alphabet, const int *values, int length) {
    if (!alphabet || !values || length <= 0) return nullptr;
    const int n = (int)std::strlen(alphabet);
    if (n == 0) return nullptr;
    const int size = n + 1;
    parasail_matrix_t *m = alloc_matrix("", std::string(alphabet) + "*", size, length, PARASAIL_MATRIX_TYPE_PSSM);
    if (!m) return nullptr;
    int lo = INT_MAX;
    for (int i = 0; i < length * n; ++i) lo = std::min(lo, values[i]);
    for (int i = 0; i < length; ++i) {
        for (int j = 0; j < n; ++j) m->user_matrix[i * size + j] = values[i * n + j];
        m->user_matrix[i * size + n] = lo;
    }
    fill_mapper((int *)m->mapper, alphabet, size);
    refresh_minmax(m);
    return m;
}

// [REF src/matrix/mod.rs:188, 281] deep copy; the copy is always heap-owned and editable
parasail_matrix_t *parasail_matrix_copy(const parasail_matrix_t *o) {
    if (!o) return nullptr;
    parasail_matrix_t *m = alloc_matrix(o->name ? o->name : "", o->alphabet ? o->alphabet : "", o->size, o->length, o->type);
    if (!m) return nullptr;
    std::memcpy(m->user_matrix, o->matrix, sizeof(int) * (size_t)o->size * (size_t)o->length);
    std::memcpy((int *)m->mapper, o->mapper, sizeof(int) * 256);
    m->max = o->max; m->min = o->min;
    if (o->query) m->query = dup_cstr(o->query);
    return m;
}

// [REF src/matrix/mod.rs:197] new PSSM whose row i is the square matrix row of s1[i]
parasail_matrix_t *parasail_matrix_convert_square_to_pssm(const parasail_matrix_t *sq, const char *s1, int s1Len) {
    if (!sq || !s1 || s1Len <= 0 || sq->type != PARASAIL_MATRIX_TYPE_SQUARE) return nullptr;
    parasail_matrix_t *m = alloc_matrix(sq->name ? sq->name : "", sq->alphabet ? sq->alphabet : "", sq->size, s1Len,
                                        PARASAIL_MATRIX_TYPE_PSSM);
    if (!m) return nullptr;
    for (int i = 0; i < s1Len; ++i) {
        const int row = sq->mapper[(unsigned char)s1[i]];
        std::memcpy(m->user_matrix + (size_t)i * sq->size, sq->matrix + (size_t)row * sq->size, sizeof(int) * (size_t)sq->size);
    }
    std::memcpy((int *)m->mapper, sq->mapper, sizeof(int) * 256);
    m->query = dup_cstr(std::string(s1, s1 + s1Len));
    refresh_minmax(m);
    return m;
}

// [REF src/matrix/mod.rs:238] the Rust side has already bounds-checked against size-2
void parasail_matrix_set_value(parasail_matrix_t *m, int row, int col, int value) {
    if (!m || !m->user_matrix) { psb::set_error("parasail_matrix_set_value: matrix is not editable"); return; }
    if (row < 0 || row >= m->length || col < 0 || col >= m->size) { psb::set_error("parasail_matrix_set_value: index out of range"); return; }
    m->user_matrix[(size_t)row * m->size + col] = value;
    refresh_minmax(m);
}

// [REF src/matrix/mod.rs:304]
void parasail_matrix_free(parasail_matrix_t *m) {
    if (!m || m == &g_blosum62) return;
    std::free((void *)m->name);
    std::free((void *)m->alphabet);
    std::free((void *)m->query);
    std::free((void *)m->mapper);
    std::free(m->user_matrix);
    std::free(m);
}

}  // extern "C"

namespace psb {
HostMatrix::HostMatrix(const parasail_matrix_t *m) {
    size = m->size; length = m->length; type = m->type; max = m->max; min = m->min;
    table.assign(m->matrix, m->matrix + (size_t)size * (size_t)length);
    for (int i = 0; i < 256; ++i) {
        int v = m->mapper[i];
        if (v < 0 || v >= size) v = size - 1;
        mapper[i] = (uint8_t)v;
    }
    if (m->query) query.assign((const uint8_t *)m->query, (const uint8_t *)m->query + std::strlen(m->query));
}
}  // namespace psb

/*
 * striped_cpu.cpp -- TEST/BENCH INFRASTRUCTURE ONLY: the CPU baseline timed beside the GPU number.
 *
 * "parasail-equivalent striped AVX2 restatement" (BASELINE.md section 4): parasail itself cannot be
 * built here (libparasail-sys 0.2.1 -> jeffdaily/parasail is not on disk, no network), so this
 * file restates its fastest CPU strategy for the benchmarked path -- Farrar's striped
 * Smith-Waterman over a query profile, 8-bit lanes first, then 16-bit, then the scalar 32-bit
 * fill, i.e. the `_sat` escalation behind `sw_striped_profile_sat`
 * [REF src/aligner/mod.rs:90, 125-126, 431-450] -- with AVX2, the widest ISA parasail-rs exposes
 * [REF src/prelude.rs:18-25].  It returns score, end_query and end_ref with parasail's rule
 * (first column whose maximum strictly exceeds the running best; smallest query index in that
 * column) and is cross-checked against oracle/gotoh_oracle.c in tests/test_striped_cpu.py before
 * it is ever timed.  One alignment per thread across the host cores, the way the reference's own
 * multi-thread test uses the library [REF tests/test_parasail.rs:689-723].
 *
 * Never linked into libparasail_b200.so.
 */
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Profile {
    int qlen = 0, size = 0, bias = 0, maxs = 0;
    int seg8 = 0, seg16 = 0;
    std::vector<__m256i> p8, p16;  // [letter][segment]
    std::vector<int> q;            // mapped query
    const int *matrix = nullptr;
};

inline __m256i shl1_epi8(__m256i a) {
    return _mm256_alignr_epi8(a, _mm256_permute2x128_si256(a, a, 0x08), 15);
}
inline __m256i shl1_epi16(__m256i a) {
    return _mm256_alignr_epi8(a, _mm256_permute2x128_si256(a, a, 0x08), 14);
}

void build_profile(Profile &P, const uint8_t *q, int qlen, const int *matrix, int size, const int *mapper) {
    P.qlen = qlen; P.size = size; P.matrix = matrix;
    P.q.resize(qlen);
    for (int i = 0; i < qlen; ++i) P.q[i] = mapper[q[i]];
    int mn = 0, mx = 0;
    for (int i = 0; i < size * size; ++i) { mn = std::min(mn, matrix[i]); mx = std::max(mx, matrix[i]); }
    P.bias = -mn; P.maxs = mx;
    P.seg8 = (qlen + 31) / 32; P.seg16 = (qlen + 15) / 16;
    P.p8.resize((size_t)size * P.seg8);
    P.p16.resize((size_t)size * P.seg16);
    for (int a = 0; a < size; ++a) {
        for (int s = 0; s < P.seg8; ++s) {
            alignas(32) uint8_t v[32];
            for (int l = 0; l < 32; ++l) {
                const int i = s + l * P.seg8;
                v[l] = (uint8_t)((i < qlen ? matrix[P.q[i] * size + a] : mn) + P.bias);
            }
            P.p8[(size_t)a * P.seg8 + s] = _mm256_load_si256((const __m256i *)v);
        }
        for (int s = 0; s < P.seg16; ++s) {
            alignas(32) int16_t v[16];
            for (int l = 0; l < 16; ++l) {
                const int i = s + l * P.seg16;
                v[l] = (int16_t)(i < qlen ? matrix[P.q[i] * size + a] : mn);
            }
            P.p16[(size_t)a * P.seg16 + s] = _mm256_load_si256((const __m256i *)v);
        }
    }
}

struct Res { int score, end_query, end_ref; bool overflow; };

// 8-bit unsigned lanes with bias (Farrar 2007)
Res sw8(const Profile &P, const int *r, int rlen, int open, int gap, __m256i *Hs, __m256i *Hl, __m256i *E, __m256i *Hmax) {
    const int seg = P.seg8;
    const __m256i vBias = _mm256_set1_epi8((char)P.bias), vO = _mm256_set1_epi8((char)open), vG = _mm256_set1_epi8((char)gap);
    const __m256i vZero = _mm256_setzero_si256();
    for (int i = 0; i < seg; ++i) { Hs[i] = vZero; E[i] = vZero; Hmax[i] = vZero; }
    int score = 0, end_ref = 0;
    for (int j = 0; j < rlen; ++j) {
        const __m256i *prof = &P.p8[(size_t)r[j] * seg];
        __m256i vF = vZero, vMaxCol = vZero;
        __m256i vH = shl1_epi8(Hs[seg - 1]);
        std::swap(Hs, Hl);
        for (int i = 0; i < seg; ++i) {
            vH = _mm256_subs_epu8(_mm256_adds_epu8(vH, prof[i]), vBias);
            __m256i e = E[i];
            vH = _mm256_max_epu8(vH, e);
            vH = _mm256_max_epu8(vH, vF);
            vMaxCol = _mm256_max_epu8(vMaxCol, vH);
            Hs[i] = vH;
            vH = _mm256_subs_epu8(vH, vO);
            e = _mm256_max_epu8(_mm256_subs_epu8(e, vG), vH);
            E[i] = e;
            vF = _mm256_max_epu8(_mm256_subs_epu8(vF, vG), vH);
            vH = Hl[i];
        }
        // lazy F: carry F across the lane boundary until it can no longer raise anything.  Only
        // the carried component F - e travels (the main pass already holds H - o of every row); the
        // loop stops when it is no better than what the main pass derived from the UNcorrected H,
        // which stays valid when open == extend (the callers route open < extend to the scalar fill).
        for (int k = 0; k < 32; ++k) {
            vF = shl1_epi8(vF);
            bool done = false;
            for (int i = 0; i < seg; ++i) {
                const __m256i hold = Hs[i];
                const __m256i told = _mm256_subs_epu8(hold, vO);
                const __m256i nh = _mm256_max_epu8(hold, vF);
                Hs[i] = nh;
                vMaxCol = _mm256_max_epu8(vMaxCol, nh);
                const __m256i t = _mm256_subs_epu8(nh, vO);
                E[i] = _mm256_max_epu8(E[i], t);  // E of the next column sees the corrected H
                vF = _mm256_subs_epu8(vF, vG);
                if (_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_subs_epu8(vF, told), vZero)) == -1) { done = true; break; }
            }
            if (done) break;
        }
        // column maximum vs running best (strictly greater => first column wins)
        __m256i m = vMaxCol;
        m = _mm256_max_epu8(m, _mm256_permute2x128_si256(m, m, 1));
        m = _mm256_max_epu8(m, _mm256_srli_si256(m, 8));
        m = _mm256_max_epu8(m, _mm256_srli_si256(m, 4));
        m = _mm256_max_epu8(m, _mm256_srli_si256(m, 2));
        m = _mm256_max_epu8(m, _mm256_srli_si256(m, 1));
        const int cm = _mm256_extract_epi8(m, 0) & 0xff;
        if (cm > score) {
            score = cm; end_ref = j;
            std::memcpy(Hmax, Hs, sizeof(__m256i) * seg);
            if (score + P.bias >= 255) return {score, 0, end_ref, true};
        }
    }
    int end_query = P.qlen;
    const uint8_t *t = (const uint8_t *)Hmax;
    for (int x = 0; x < seg * 32; ++x)
        if (t[x] == score) {
            const int i = x / 32 + (x % 32) * seg;
            if (i < end_query) end_query = i;
        }
    if (score == 0) { end_query = 0; end_ref = 0; }
    return {score, end_query, end_ref, false};
}

// 16-bit signed lanes
Res sw16(const Profile &P, const int *r, int rlen, int open, int gap, __m256i *Hs, __m256i *Hl, __m256i *E, __m256i *Hmax) {
    const int seg = P.seg16;
    const __m256i vO = _mm256_set1_epi16((short)open), vG = _mm256_set1_epi16((short)gap), vZero = _mm256_setzero_si256();
    const __m256i vNeg = _mm256_set1_epi16(-16000);
    for (int i = 0; i < seg; ++i) { Hs[i] = vZero; E[i] = vNeg; Hmax[i] = vZero; }
    int score = 0, end_ref = 0;
    for (int j = 0; j < rlen; ++j) {
        const __m256i *prof = &P.p16[(size_t)r[j] * seg];
        __m256i vF = vNeg, vMaxCol = vZero;
        __m256i vH = shl1_epi16(Hs[seg - 1]);
        std::swap(Hs, Hl);
        for (int i = 0; i < seg; ++i) {
            vH = _mm256_adds_epi16(vH, prof[i]);
            __m256i e = E[i];
            vH = _mm256_max_epi16(vH, e);
            vH = _mm256_max_epi16(vH, vF);
            vH = _mm256_max_epi16(vH, vZero);
            vMaxCol = _mm256_max_epi16(vMaxCol, vH);
            Hs[i] = vH;
            vH = _mm256_subs_epi16(vH, vO);
            e = _mm256_max_epi16(_mm256_subs_epi16(e, vG), vH);
            E[i] = e;
            vF = _mm256_max_epi16(_mm256_subs_epi16(vF, vG), vH);
            vH = Hl[i];
        }
        for (int k = 0; k < 16; ++k) {
            vF = shl1_epi16(vF);
            vF = _mm256_insert_epi16(vF, -16000, 0);
            bool done = false;
            for (int i = 0; i < seg; ++i) {
                const __m256i hold = Hs[i];
                const __m256i told = _mm256_subs_epi16(hold, vO);
                const __m256i nh = _mm256_max_epi16(hold, vF);
                Hs[i] = nh;
                vMaxCol = _mm256_max_epi16(vMaxCol, nh);
                const __m256i t = _mm256_subs_epi16(nh, vO);
                E[i] = _mm256_max_epi16(E[i], t);
                vF = _mm256_subs_epi16(vF, vG);
                if (_mm256_movemask_epi8(_mm256_cmpgt_epi16(vF, told)) == 0) { done = true; break; }
            }
            if (done) break;
        }
        __m256i m = vMaxCol;
        m = _mm256_max_epi16(m, _mm256_permute2x128_si256(m, m, 1));
        m = _mm256_max_epi16(m, _mm256_srli_si256(m, 8));
        m = _mm256_max_epi16(m, _mm256_srli_si256(m, 4));
        m = _mm256_max_epi16(m, _mm256_srli_si256(m, 2));
        const int cm = (int16_t)_mm256_extract_epi16(m, 0);
        if (cm > score) {
            score = cm; end_ref = j;
            std::memcpy(Hmax, Hs, sizeof(__m256i) * seg);
            if (score >= 32767 - P.maxs) return {score, 0, end_ref, true};
        }
    }
    int end_query = P.qlen;
    const int16_t *t = (const int16_t *)Hmax;
    for (int x = 0; x < seg * 16; ++x)
        if (t[x] == score) {
            const int i = x / 16 + (x % 16) * seg;
            if (i < end_query) end_query = i;
        }
    if (score == 0) { end_query = 0; end_ref = 0; }
    return {score, end_query, end_ref, false};
}

// scalar 32-bit last resort of the escalation (row-major scan, parasail's scalar tie-break)
Res sw32(const Profile &P, const int *r, int rlen, int open, int gap) {
    const int n = P.qlen;
    std::vector<int> H(rlen + 1, 0), F(rlen + 1, INT32_MIN / 2);
    int score = INT32_MIN / 2, eq = 0, er = 0;
    for (int i = 1; i <= n; ++i) {
        const int *row = &P.matrix[(size_t)P.size * P.q[i - 1]];
        int NH = H[0], WH = 0, E = INT32_MIN / 2;
        for (int j = 1; j <= rlen; ++j) {
            const int NWH = NH;
            NH = H[j];
            F[j] = std::max(NH - open, F[j] - gap);
            E = std::max(WH - open, E - gap);
            WH = std::max(std::max(NWH + row[r[j - 1]], E), std::max(F[j], 0));
            H[j] = WH;
            if (WH > score || (WH == score && j - 1 < er)) { score = WH; eq = i - 1; er = j - 1; }
        }
    }
    if (score <= 0) { score = 0; eq = 0; er = 0; }
    return {score, eq, er, false};
}

}  // namespace

extern "C" {

/*
 * One query profile against n subjects with `threads` worker threads (dynamic chunks).
 * widths_out (optional, n bytes) records which width produced each result (8, 16 or 32).
 * Returns elapsed seconds of the alignment phase (profile construction included).
 */
double psbs_sw_scan(const uint8_t *query, int qlen, const uint8_t *cat, const int64_t *off, int64_t n, const int *matrix,
                    int size, const int *mapper, int open, int gap, int threads, int *score, int *end_query, int *end_ref,
                    uint8_t *widths_out) {
    const auto t0 = std::chrono::steady_clock::now();
    Profile P;
    build_profile(P, query, qlen, matrix, size, mapper);
    const bool striped_ok = gap <= open;  // see the lazy-F note; parasail documents open >= extend
    const bool ok8 = striped_ok && P.bias + P.maxs < 255 && open < 255 && gap < 255;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        const int seg = std::max(P.seg8, P.seg16);
        __m256i *mem = (__m256i *)aligned_alloc(32, sizeof(__m256i) * 4 * (size_t)seg);
        std::vector<int> r;
        for (;;) {
            const int64_t lo = next.fetch_add(64);
            if (lo >= n) break;
            const int64_t hi = std::min(n, lo + 64);
            for (int64_t s = lo; s < hi; ++s) {
                const int rlen = (int)(off[s + 1] - off[s]);
                r.resize(rlen);
                for (int j = 0; j < rlen; ++j) r[j] = mapper[cat[off[s] + j]];
                Res res{0, 0, 0, true};
                int w = 8;
                if (ok8) res = sw8(P, r.data(), rlen, open, gap, mem, mem + seg, mem + 2 * seg, mem + 3 * seg);
                if (res.overflow && striped_ok) { w = 16; res = sw16(P, r.data(), rlen, open, gap, mem, mem + seg, mem + 2 * seg, mem + 3 * seg); }
                if (res.overflow) { w = 32; res = sw32(P, r.data(), rlen, open, gap); }
                score[s] = res.score; end_query[s] = res.end_query; end_ref[s] = res.end_ref;
                if (widths_out) widths_out[s] = (uint8_t)w;
            }
        }
        free(mem);
    };
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, threads); ++t) th.emplace_back(worker);
    for (auto &t : th) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int psbs_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"

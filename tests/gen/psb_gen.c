/* psb_gen.c -- TEST INFRASTRUCTURE: the synthetic generator of SURVEY.md Appendix F in C with threads, so that
 * the full-size configurations (C3: 10^7 read/window pairs, 6.5e9 residues) can be produced in seconds.
 * Bit-identical to tests/psb_data.py (rnd / protein_letters / dna_letters) and to the vectorised helpers of
 * tests/bench_configs.py (vec_random / vec_substitute); tests/test_gen.py checks that. */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline uint64_t rnd(uint64_t seed, uint64_t stream, uint64_t idx) {
    return splitmix64(seed ^ (stream * 0x9E3779B97F4A7C15ull) ^ (idx * 0xD1B54A32D192ED03ull));
}
static const char PROTEIN[] = "ARNDCQEGHILKMFPSTWYV";
static const int PROT_CUM[20] = {83, 138, 179, 234, 248, 287, 355, 426, 449, 508, 604, 662, 686, 725, 772, 838, 891, 902, 931, 1000};
static const char DNA[] = "ACGT";
static inline uint8_t protein_letter(uint64_t u) {
    const int x = (int)(u % 1000);
    int k = 0;
    while (PROT_CUM[k] <= x) ++k;   /* searchsorted(side="right") */
    return (uint8_t)PROTEIN[k];
}

typedef struct job {
    int kind;            /* 0 random, 1 substitute, 2 gather */
    uint64_t seed, stream;
    int64_t lo, hi;
    int protein;
    int rate_1e4;
    const uint8_t *src;
    uint8_t *out;
    /* gather */
    const int64_t *starts; int64_t win_len, read_len;
} job_t;

static void *work(void *arg) {
    job_t *j = (job_t *)arg;
    if (j->kind == 0) {
        for (int64_t i = j->lo; i < j->hi; ++i) {
            const uint64_t u = rnd(j->seed, j->stream, (uint64_t)i);
            j->out[i] = j->protein ? protein_letter(u) : (uint8_t)DNA[u & 3];
        }
    } else if (j->kind == 1) {
        const int nl = j->protein ? 20 : 4;
        const char *letters = j->protein ? PROTEIN : DNA;
        for (int64_t i = j->lo; i < j->hi; ++i) {
            const uint64_t u = rnd(j->seed, j->stream, (uint64_t)i);
            const int hit = (int)(u % 10000) < j->rate_1e4;
            j->out[i] = hit ? (uint8_t)letters[(u >> 20) % (uint64_t)nl] : j->src[i];
        }
    } else {
        for (int64_t p = j->lo; p < j->hi; ++p)
            memcpy(j->out + p * j->read_len, j->src + p * j->win_len + j->starts[p], (size_t)j->read_len);
    }
    return NULL;
}

static void run(job_t base, int64_t n, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    job_t jobs[256];
    int started = 0;
    for (int t = 0; t < threads; ++t) {
        jobs[t] = base;
        jobs[t].lo = n * t / threads; jobs[t].hi = n * (t + 1) / threads;
        if (pthread_create(&th[started], NULL, work, &jobs[t]) == 0) ++started; else work(&jobs[t]);
    }
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

/* out[i] = letter(rnd(seed, stream, i)) for i in [0, n) */
void psbg_random(uint64_t seed, uint64_t stream, int64_t n, int protein, uint8_t *out, int threads) {
    job_t j; memset(&j, 0, sizeof(j));
    j.kind = 0; j.seed = seed; j.stream = stream; j.protein = protein; j.out = out;
    run(j, n, threads);
}
/* out[i] = (rnd(seed, stream, i) % 10000 < rate_1e4) ? letters[(u >> 20) % nl] : src[i] */
void psbg_substitute(const uint8_t *src, int64_t n, uint64_t seed, uint64_t stream, int rate_1e4, int protein, uint8_t *out, int threads) {
    job_t j; memset(&j, 0, sizeof(j));
    j.kind = 1; j.seed = seed; j.stream = stream; j.protein = protein; j.rate_1e4 = rate_1e4; j.src = src; j.out = out;
    run(j, n, threads);
}
/* out[p*read_len .. ] = src[p*win_len + starts[p] .. + read_len] */
void psbg_gather(const uint8_t *src, int64_t npairs, int64_t win_len, const int64_t *starts, int64_t read_len, uint8_t *out, int threads) {
    job_t j; memset(&j, 0, sizeof(j));
    j.kind = 2; j.src = src; j.out = out; j.starts = starts; j.win_len = win_len; j.read_len = read_len;
    run(j, npairs, threads);
}
/* starts[p] = rnd(seed, stream, p) % mod */
void psbg_starts(uint64_t seed, uint64_t stream, int64_t n, uint64_t mod, int64_t *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = (int64_t)(rnd(seed, stream, (uint64_t)i) % mod);
}

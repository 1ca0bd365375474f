"""Synthetic inputs and matrices shared by the tests and bench.py.

Everything is a pure function of (seed, sequence id): the same corpus can be regenerated on the
GPU box without shipping data.  Integer arithmetic only, except the log-normal length draw which
is host-only by construction.
"""
import numpy as np

BLOSUM62_ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*"
BLOSUM62_TEXT = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0 -4
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1 -4
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1 -4
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1 -4
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1 -4
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1 -4
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1 -4
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1 -4
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1 -4
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1 -4
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1 -4
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1 -4
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2 -4
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0 -4
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0 -4
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2 -4
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1 -4
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1 -4
-2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1 -4
-1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1 -4
-4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1
"""


def blosum62_table():
    return np.array(BLOSUM62_TEXT.split(), dtype=np.int32).reshape(24, 24)


PROTEIN = b"ARNDCQEGHILKMFPSTWYV"
# per-mille background weights (Appendix F), same order as PROTEIN
PROTEIN_W = [83, 55, 41, 55, 14, 39, 68, 71, 23, 59, 96, 58, 24, 39, 47, 66, 53, 11, 29, 69]
_PROT_CUM = np.cumsum(PROTEIN_W)
assert _PROT_CUM[-1] == 1000
DNA = b"ACGT"

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def rnd(seed, stream, idx):
    """u64 rnd(seed, stream, idx) of Appendix F (vectorised over idx and/or stream)."""
    with np.errstate(over="ignore"):
        s = np.asarray(stream, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        i = np.asarray(idx, dtype=np.uint64) * np.uint64(0xD1B54A32D192ED03)
        return splitmix64(np.uint64(seed) ^ s ^ i)


def protein_letters(u):
    """map u64 draws to protein letters with the background weights"""
    k = np.searchsorted(_PROT_CUM, (u % np.uint64(1000)).astype(np.int64), side="right")
    return np.frombuffer(PROTEIN, dtype=np.uint8)[k]


def dna_letters(u):
    return np.frombuffer(DNA, dtype=np.uint8)[(u & np.uint64(3)).astype(np.int64)]


def random_seq(seed, seq_id, length, protein=True):
    u = rnd(seed, seq_id, np.arange(length, dtype=np.uint64))
    return protein_letters(u) if protein else dna_letters(u)


def mutate(src, seed, seq_id, p_sub, p_indel, protein=True, geometric_mean=1.0):
    """Appendix F mutate(): per position substitute / delete / insert-before; deterministic."""
    letters = np.frombuffer(PROTEIN if protein else DNA, dtype=np.uint8)
    out = []
    n = len(src)
    u = rnd(seed, seq_id, np.arange(4 * n + 8, dtype=np.uint64) + np.uint64(1 << 40))
    c = 0

    def draw():
        nonlocal c
        v = int(u[c % len(u)])
        c += 1
        return v

    scale = 1 << 30
    for pos in range(n):
        x = (draw() >> 20) % scale
        if x < p_sub * scale:
            ch = int(src[pos])
            while True:
                new = int(letters[draw() % len(letters)])
                if new != ch:
                    break
            out.append(new)
        elif x < (p_sub + p_indel / 2) * scale:
            continue  # delete (geometric extension handled by the caller's rate)
        elif x < (p_sub + p_indel) * scale:
            k = 1
            while geometric_mean > 1.0 and (draw() % 1000) < 1000 * (1 - 1 / geometric_mean):
                k += 1
            for _ in range(k):
                out.append(int(letters[draw() % len(letters)]))
            out.append(int(src[pos]))
        else:
            out.append(int(src[pos]))
    if not out:
        out.append(int(src[0]))
    return np.array(out, dtype=np.uint8)


def lognormal_lengths(seed, n, mu=5.68, sigma=0.65, lo=16, hi=35000):
    """C2 subject lengths: clip(round(exp(N(mu, sigma^2))), lo, hi); Box-Muller on two draws."""
    i = np.arange(n, dtype=np.uint64)
    u1 = (rnd(seed, 1, i) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    u2 = (rnd(seed, 2, i) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    z = np.sqrt(-2.0 * np.log(np.maximum(u1, 1e-300))) * np.cos(2.0 * np.pi * u2)
    return np.clip(np.rint(np.exp(mu + sigma * z)), lo, hi).astype(np.int64)


def concat(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    cat = np.concatenate(seqs).astype(np.uint8) if seqs else np.zeros(0, dtype=np.uint8)
    return cat, off


def protein_db(seed_len, seed_res, n, query=None, planted_frac=0.01):
    """C2 database: n proteins with log-normal lengths as one flat array + offsets.  A
    `planted_frac` of the subjects carries a mutated copy of a random query segment."""
    lens = lognormal_lengths(seed_len, n)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    # i.i.d. residues as a function of the global residue index (vectorised, chunked)
    cat = np.empty(total, dtype=np.uint8)
    step = 1 << 24
    for a in range(0, total, step):
        b = min(total, a + step)
        cat[a:b] = protein_letters(rnd(seed_res, 7, np.arange(a, b, dtype=np.uint64)))
    if query is not None and planted_frac > 0:
        nplant = max(1, int(n * planted_frac))
        ids = (rnd(seed_res, 11, np.arange(nplant, dtype=np.uint64)) % np.uint64(n)).astype(np.int64)
        for t, sid in enumerate(ids):
            L = int(lens[sid])
            u = rnd(seed_res, 13, np.arange(3, dtype=np.uint64) + np.uint64(3 * t))
            seg_len = min(L, len(query), 50 + int(u[0] % np.uint64(351)))
            qs = int(u[1] % np.uint64(len(query) - seg_len + 1))
            seg = mutate(query[qs:qs + seg_len], seed_res, 1000 + t, 0.15, 0.03)
            seg = seg[:L]
            ds = int(u[2] % np.uint64(L - len(seg) + 1))
            cat[off[sid] + ds: off[sid] + ds + len(seg)] = seg
    return cat, off


def protein_pairs(seed, n, length, related_frac=0.1, p_sub=0.15, p_indel=0.03, geometric_mean=1.0):
    """C1/C4-style many-pairs corpus: independent random proteins, a fraction related by mutate()."""
    qs, rs = [], []
    for p in range(n):
        q = random_seq(seed, 2 * p, length)
        pick = int(rnd(seed, 5, p) % np.uint64(1000))
        if pick < related_frac * 1000:
            r = mutate(q, seed, 2 * p + 1, p_sub, p_indel, True, geometric_mean)
            if len(r) >= length:
                r = r[:length]
            else:
                r = np.concatenate([r, random_seq(seed + 1, 2 * p + 1, length - len(r))])
        else:
            r = random_seq(seed, 2 * p + 1, length)
        qs.append(q)
        rs.append(r)
    return qs, rs


def dna_read_pairs(seed, n, read_len=150, win_len=500, unrelated_frac=0.05):
    """C3 corpus: 500 bp windows, 150 bp reads sampled from them with 4% subst, 0.5%+0.5% indel."""
    qs, rs = [], []
    for p in range(n):
        win = random_seq(seed, 2 * p, win_len, protein=False)
        pick = int(rnd(seed, 5, p) % np.uint64(1000))
        if pick < unrelated_frac * 1000:
            read = random_seq(seed, 2 * p + 1, read_len, protein=False)
        else:
            start = int(rnd(seed, 6, p) % np.uint64(win_len - read_len - 10))
            read = mutate(win[start:start + read_len + 10], seed, 2 * p + 1, 0.04, 0.01, protein=False)
            if len(read) >= read_len:
                read = read[:read_len]
            else:
                read = np.concatenate([read, random_seq(seed + 1, 2 * p + 1, read_len - len(read), protein=False)])
        qs.append(read)
        rs.append(win)
    return qs, rs

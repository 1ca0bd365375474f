"""ctypes binding of the test-only CPU oracle (oracle/gotoh_oracle.c).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NW, SG, SW = 0, 1, 2


class Config(C.Structure):
    _fields_ = [("mode", C.c_int), ("s1_beg", C.c_int), ("s1_end", C.c_int), ("s2_beg", C.c_int),
                ("s2_end", C.c_int), ("open", C.c_int), ("gap", C.c_int), ("band", C.c_int)]


RULE_NAMES = ("sw_end_tie", "sg_col_wins_tie", "sg_row_last_wins", "h_priority", "open_on_tie", "match_raw_bytes",
              "count_boundary_gaps", "cigar_edge_stop", "cigar_swap_id", "sg_flag_swap", "zero_beats_diag", "band_rule")
RULE_DEFAULTS = dict(zip(RULE_NAMES, (0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0)))


class Rules(C.Structure):
    """the [UP] assumptions of SURVEY Appendix A as switches (oracle/UP_ASSUMPTIONS.md)"""
    _fields_ = [(n, C.c_int) for n in RULE_NAMES]


def set_rules(**kw):
    """override some of the upstream-behaviour assumptions (process-wide); set_rules() restores the defaults"""
    vals = dict(RULE_DEFAULTS)
    for k, v in kw.items():
        if k not in vals:
            raise KeyError(k)
        vals[k] = int(v)
    r = Rules(*[vals[n] for n in RULE_NAMES])
    assert lib().psbo_rules_count() == len(RULE_NAMES)
    lib().psbo_set_rules(C.byref(r))


def set_threads(n):
    """worker threads of align_batch (0 = all hardware threads)"""
    lib().psbo_set_threads(int(n) if n > 0 else (os.cpu_count() or 1))


class CMatrix(C.Structure):
    _fields_ = [("matrix", C.POINTER(C.c_int)), ("mapper", C.POINTER(C.c_int)), ("size", C.c_int),
                ("is_pssm", C.c_int), ("length", C.c_int)]


class Out(C.Structure):
    _fields_ = [("score", C.c_int), ("end_query", C.c_int), ("end_ref", C.c_int),
                ("matches", C.c_int), ("similar", C.c_int), ("length", C.c_int),
                ("trace", C.POINTER(C.c_int8)), ("score_table", C.POINTER(C.c_int)),
                ("matches_table", C.POINTER(C.c_int)), ("similar_table", C.POINTER(C.c_int)),
                ("length_table", C.POINTER(C.c_int)),
                ("score_row", C.POINTER(C.c_int)), ("matches_row", C.POINTER(C.c_int)),
                ("similar_row", C.POINTER(C.c_int)), ("length_row", C.POINTER(C.c_int)),
                ("score_col", C.POINTER(C.c_int)), ("matches_col", C.POINTER(C.c_int)),
                ("similar_col", C.POINTER(C.c_int)), ("length_col", C.POINTER(C.c_int))]


def build(force=False):
    """Compile the checker libraries next to their sources (gcc only, seconds)."""
    targets = ["libpsb_oracle.so"]
    if os.path.exists(os.path.join(_HERE, "striped_cpu.cpp")):
        targets.append("libpsb_striped.so")
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE] + targets, check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libpsb_oracle.so")
        src = os.path.join(_HERE, "gotoh_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _lib = C.CDLL(path)
        _lib.psbo_align.restype = C.c_int
        _lib.psbo_cigar.restype = C.c_int
        _lib.psbo_traceback.restype = C.c_int
        _lib.psbo_align_batch.restype = C.c_int
        _lib.psbo_cigar_decode.restype = C.c_int
    return _lib


class Matrix:
    """Plain-data substitution matrix for the oracle (independent of the product's matrices)."""

    def __init__(self, table, mapper, is_pssm=False):
        self.table = np.ascontiguousarray(table, dtype=np.int32)
        self.mapper = np.ascontiguousarray(mapper, dtype=np.int32)
        assert self.mapper.shape == (256,)
        self.size = int(self.table.shape[1])
        self.is_pssm = bool(is_pssm)
        self.length = int(self.table.shape[0])
        self.c = CMatrix(self.table.ctypes.data_as(C.POINTER(C.c_int)),
                         self.mapper.ctypes.data_as(C.POINTER(C.c_int)), self.size, int(self.is_pssm),
                         self.length)

    @staticmethod
    def create(alphabet: bytes, match: int, mismatch: int):
        """parasail_matrix_create semantics: size = len+1, wildcard row/col 0,
        both cases of a letter map to its index, every other byte to the wildcard."""
        n = len(alphabet) + 1
        t = np.full((n, n), mismatch, dtype=np.int32)
        np.fill_diagonal(t, match)
        t[n - 1, :] = 0
        t[:, n - 1] = 0
        mapper = np.full(256, n - 1, dtype=np.int32)
        for i, ch in enumerate(alphabet):
            mapper[ord(chr(ch).upper())] = i
            mapper[ord(chr(ch).lower())] = i
        return Matrix(t, mapper)

    @staticmethod
    def from_table(alphabet: str, rows):
        """Square matrix from an explicit table whose last letter is the wildcard."""
        n = len(alphabet)
        t = np.array(rows, dtype=np.int32).reshape(n, n)
        mapper = np.full(256, n - 1, dtype=np.int32)
        for i, ch in enumerate(alphabet):
            mapper[ord(ch.upper())] = i
            mapper[ord(ch.lower())] = i
        return Matrix(t, mapper)


def _u8(b):
    return np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b, dtype=np.uint8)


def align(q, r, mat: Matrix, mode=NW, open=0, gap=0, s1_beg=True, s1_end=True, s2_beg=True, s2_end=True,
          tables=False, rowcol=False, trace=False, band=0):
    """One pair through the oracle.  Returns a dict with score/ends/stats and any requested
    tables, rows/cols, trace bytes, CIGAR (ops, text, beg_query, beg_ref) and traceback strings."""
    qa, ra = _u8(q), _u8(r)
    qlen = mat.length if mat.is_pssm else len(qa)
    rlen = len(ra)
    cfg = Config(mode, int(s1_beg), int(s1_end), int(s2_beg), int(s2_end), open, gap, band)
    out = Out()
    keep = {}

    def arr(name, n, dtype=np.int32, ctype=C.c_int):
        a = np.zeros(n, dtype=dtype)
        keep[name] = a
        setattr(out, name, a.ctypes.data_as(C.POINTER(ctype)))

    if tables:
        for nm in ("score_table", "matches_table", "similar_table", "length_table"):
            arr(nm, qlen * rlen)
    if rowcol:
        for nm in ("score_row", "matches_row", "similar_row", "length_row"):
            arr(nm, rlen)
        for nm in ("score_col", "matches_col", "similar_col", "length_col"):
            arr(nm, qlen)
    if trace:
        arr("trace", qlen * rlen, np.int8, C.c_int8)
    rc = lib().psbo_align(qa.ctypes.data_as(C.c_void_p), C.c_int(len(qa)), ra.ctypes.data_as(C.c_void_p),
                          C.c_int(rlen), C.byref(cfg), C.byref(mat.c), C.byref(out))
    if rc != 0:
        raise ValueError("oracle rejected the input (empty sequence?)")
    res = dict(score=out.score, end_query=out.end_query, end_ref=out.end_ref, matches=out.matches,
               similar=out.similar, length=out.length)
    for k, v in keep.items():
        res[k] = v.reshape(qlen, rlen) if k.endswith("table") or k == "trace" else v
    if trace:
        ops = np.zeros(qlen + rlen + 4, dtype=np.uint32)
        bq, br = C.c_int(), C.c_int()
        n = lib().psbo_cigar(keep["trace"].ctypes.data_as(C.c_void_p), qa.ctypes.data_as(C.c_void_p),
                             C.c_int(qlen), ra.ctypes.data_as(C.c_void_p), C.c_int(rlen), C.byref(mat.c),
                             C.c_int(out.end_query), C.c_int(out.end_ref), ops.ctypes.data_as(C.c_void_p),
                             C.byref(bq), C.byref(br), C.c_int(mode))
        res["cigar_ops"] = ops[:n].copy()
        res["cigar"] = decode_cigar(ops[:n])
        res["beg_query"], res["beg_ref"] = bq.value, br.value
        bufs = [C.create_string_buffer(qlen + rlen + 1) for _ in range(3)]
        lib().psbo_traceback(keep["trace"].ctypes.data_as(C.c_void_p), qa.ctypes.data_as(C.c_void_p),
                             C.c_int(qlen), ra.ctypes.data_as(C.c_void_p), C.c_int(rlen), C.byref(mat.c),
                             C.c_int(out.end_query), C.c_int(out.end_ref), C.c_char(b"|"), C.c_char(b" "),
                             C.c_char(b" "), bufs[0], bufs[1], bufs[2], C.c_int(mode))
        res["traceback"] = tuple(b.value.decode() for b in bufs)
    return res


def decode_cigar(ops):
    tab = "MIDNSHP=X"
    return "".join(f"{int(o) >> 4}{tab[int(o) & 15]}" for o in ops)


def align_batch(qcat, qoff, rcat, roff, mat: Matrix, mode=NW, open=0, gap=0, s1_beg=True, s1_end=True,
                s2_beg=True, s2_end=True, shared_query=False, stats=False, cigar=False, threads=1):
    """n pairs through the oracle (`threads` workers, 0 = all).  Returns dict of int32 arrays (+ CIGAR CSR)."""
    qcat, rcat = _u8(qcat), _u8(rcat)
    qoff = np.ascontiguousarray(qoff, dtype=np.int64)
    roff = np.ascontiguousarray(roff, dtype=np.int64)
    n = len(roff) - 1
    cfg = Config(mode, int(s1_beg), int(s1_end), int(s2_beg), int(s2_end), open, gap, 0)
    set_threads(threads)
    res = {k: np.zeros(n, dtype=np.int32) for k in ("score", "end_query", "end_ref", "matches", "similar",
                                                     "length", "beg_query", "beg_ref")}
    cap = 0
    cig_ops = np.zeros(1, dtype=np.uint32)
    cig_off = np.zeros(n + 1, dtype=np.int64)
    if cigar:
        qtot = (qoff[1] - qoff[0]) * n if shared_query else qoff[-1] - qoff[0]
        cap = int(qtot + roff[-1] - roff[0]) + 8
        cig_ops = np.zeros(cap, dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    want_stats = stats
    rc = lib().psbo_align_batch(p(qcat), p(qoff), p(rcat), p(roff), C.c_int64(n), C.c_int(int(shared_query)),
                                C.byref(cfg), C.byref(mat.c), p(res["score"]), p(res["end_query"]),
                                p(res["end_ref"]), p(res["matches"]) if want_stats else None,
                                p(res["similar"]) if want_stats else None,
                                p(res["length"]) if want_stats else None, C.c_int(int(cigar)), p(cig_ops),
                                C.c_int64(cap), p(cig_off), p(res["beg_query"]), p(res["beg_ref"]))
    if rc != 0:
        raise RuntimeError("oracle batch failed")
    if cigar:
        res["cigar_off"] = cig_off
        res["cigar_ops"] = cig_ops[: cig_off[-1]].copy()
    return res


def cigar_check_batch(qcat, qoff, rcat, roff, mat: Matrix, mode, open, gap, cigar_ops, cigar_off, beg_query, beg_ref,
                      end_query, end_ref, score, threads=0):
    """CIGAR re-score / recount property on a whole batch (no oracle fill involved): returns (number of
    pairs whose CIGAR does not re-derive its own score and end cell, index of the first one or -1)."""
    qcat, rcat = _u8(qcat), _u8(rcat)
    a64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
    a32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    qoff, roff, cigar_off = a64(qoff), a64(roff), a64(cigar_off)
    ops = np.ascontiguousarray(cigar_ops, dtype=np.uint32)
    bq, br, eq, er, sc = a32(beg_query), a32(beg_ref), a32(end_query), a32(end_ref), a32(score)
    cfg = Config(mode, 1, 1, 1, 1, open, gap, 0)
    set_threads(threads)
    first = C.c_int64(-1)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().psbo_cigar_check_batch.restype = C.c_int64
    bad = lib().psbo_cigar_check_batch(p(qcat), p(qoff), p(rcat), p(roff), C.c_int64(len(roff) - 1), C.byref(cfg), C.byref(mat.c),
                                       p(ops), p(cigar_off), p(bq), p(br), p(eq), p(er), p(sc), C.byref(first))
    return int(bad), int(first.value)


# ---- the striped AVX2 CPU baseline (oracle/striped_cpu.cpp) -----------------------------------
_slib = None


def striped_lib():
    global _slib
    if _slib is None:
        path = os.path.join(_HERE, "libpsb_striped.so")
        src = os.path.join(_HERE, "striped_cpu.cpp")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _slib = C.CDLL(path)
        _slib.psbs_sw_scan.restype = C.c_double
        _slib.psbs_hardware_threads.restype = C.c_int
    return _slib


def striped_sw_scan(query, cat, off, mat: Matrix, open, gap, threads=0):
    """parasail-equivalent striped AVX2 `sw_striped_profile_sat` over a database, `threads` host
    threads (0 = all).  Returns (dict of arrays incl. 'width', elapsed seconds)."""
    q, cat = _u8(query), _u8(cat)
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = len(off) - 1
    L = striped_lib()
    if threads <= 0:
        threads = L.psbs_hardware_threads()
    res = {k: np.zeros(n, dtype=np.int32) for k in ("score", "end_query", "end_ref")}
    width = np.zeros(n, dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    secs = L.psbs_sw_scan(p(q), C.c_int(len(q)), p(cat), p(off), C.c_int64(n), p(mat.table), C.c_int(mat.size),
                          p(mat.mapper), C.c_int(open), C.c_int(gap), C.c_int(threads), p(res["score"]),
                          p(res["end_query"]), p(res["end_ref"]), p(width))
    res["width"] = width
    res["threads"] = threads
    return res, float(secs)

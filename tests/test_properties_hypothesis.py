"""Property tests (hypothesis): the oracle against the independent pure-Python Gotoh on arbitrary
small inputs, and the function-name grammar of the C ABI against the builder that produces it."""
import itertools

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import py_gotoh

MODES = {0: "nw", 1: "sg", 2: "sw"}
dna = st.text(alphabet="ACGTN", min_size=1, max_size=24)


@settings(max_examples=150, deadline=None)
@given(q=dna, r=dna, mode=st.sampled_from([0, 1, 2]), o=st.integers(0, 12), e=st.integers(0, 6),
       flags=st.tuples(st.booleans(), st.booleans(), st.booleans(), st.booleans()),
       match=st.integers(0, 5), mismatch=st.integers(-5, 0))
def test_oracle_equals_python_gotoh(oracle, q, r, mode, o, e, flags, match, mismatch):
    mat = oracle.Matrix.create(b"ACGT", match, mismatch)
    res = oracle.align(q.encode(), r.encode(), mat, mode=mode, open=o, gap=e, s1_beg=flags[0], s1_end=flags[1],
                       s2_beg=flags[2], s2_end=flags[3])
    exp = py_gotoh.gotoh(q.encode(), r.encode(), mat.table, mat.mapper, MODES[mode], o, e, *flags)
    assert (res["score"], res["end_query"], res["end_ref"]) == exp


@pytest.fixture(scope="module")
def ps():
    import __graft_entry__ as g
    g.build()
    import parasail_rs_b200 as ps
    return ps


def test_every_builder_name_resolves_or_panics_like_upstream(ps):
    # every combination the Rust builder can express: upstream-defined names resolve to distinct
    # entry points, the q/d pairings upstream does not define make build() panic
    from parasail_rs_b200 import _lib
    L = _lib.lib()
    seen = {}
    gaps = [[], ["prefix"], ["suffix"], ["prefix", "suffix"]]
    for mode, qg, dg, out, strat, width in itertools.product(
            ["global_", "semi_global", "local"], gaps, gaps, ["", "stats", "table", "rowcol", "stats_table", "stats_rowcol", "trace"],
            ["striped", "scan", "diag"], ["sat", 8, 16, 32, 64]):
        if mode != "semi_global" and (qg or dg):
            continue
        b = getattr(ps.Aligner.new(), mode)().allow_query_gaps(qg).allow_ref_gaps(dg)
        if "stats" in out: b = b.use_stats()
        if "table" in out: b = b.use_table()
        if "rowcol" in out: b = b.use_last_rowcol()
        if out == "trace": b = b.use_trace()
        b = getattr(b, strat)()
        if width != "sat": b = b.solution_width(width)
        name = b.get_parasail_fn_name()
        ptr = L.parasail_lookup_function(name.encode())
        undefined = mode == "semi_global" and bool(qg) and bool(dg) and (len(qg) == 2) != (len(dg) == 2)
        if undefined:
            assert not ptr, name
            with pytest.raises(ps.Panic):
                b.build()
        else:
            assert ptr, name
            assert seen.setdefault(ptr, name) == name, (name, seen[ptr])   # one entry point per name
            assert L.parasail_lookup_function(("parasail_" + name).encode()) == ptr
    assert len(seen) == 13 * 7 * 3 * 5

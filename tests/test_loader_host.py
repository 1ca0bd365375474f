"""Host-side pieces in front of the hot path that need no GPU: the FASTA reader and the data-driven
matrix lookup (PSB_MATRIX_DIR)."""
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ps():
    import __graft_entry__ as g
    g.build()
    import parasail_rs_b200 as ps
    return ps


def test_fasta_reader(ps, tmp_path):
    p = tmp_path / "x.fa"
    p.write_bytes(b">sp|P1|first protein\nMKV\r\nLAA\n\n>second\nACDEFGHIKL*\n>third one\nW\n")
    cat, off, names = ps.read_fasta(p)
    assert names == ["sp|P1|first protein", "second", "third one"]
    assert list(off) == [0, 6, 16, 17]
    assert bytes(cat) == b"MKVLAAACDEFGHIKLW"


def test_fasta_without_header_and_errors(ps, tmp_path):
    p = tmp_path / "plain.fa"
    p.write_bytes(b"ACGT\nACGT\n")
    cat, off, names = ps.read_fasta(p)
    assert bytes(cat) == b"ACGTACGT" and list(off) == [0, 8] and names == [""]
    with pytest.raises(ps.Error):
        ps.read_fasta(tmp_path / "missing.fa")
    e = tmp_path / "empty.fa"
    e.write_bytes(b"\n\n")
    with pytest.raises(ps.Error):
        ps.read_fasta(e)


def test_matrix_lookup_is_data_driven(ps, tmp_path, monkeypatch):
    # upstream's other built-in tables are not fabricated here; a directory of NCBI-format files makes the
    # names resolve [REF src/matrix/mod.rs:57-73: Matrix::from(name)]
    with pytest.raises(ps.FailedLookup):
        ps.Matrix.from_name("pam_test_1")
    golden = os.path.join(os.path.dirname(__file__), "golden", "square.txt")
    (tmp_path / "pam_test_1.txt").write_bytes(open(golden, "rb").read())
    monkeypatch.setenv("PSB_MATRIX_DIR", str(tmp_path))
    m = ps.Matrix.from_name("pam_test_1")
    ref = ps.Matrix.from_file(golden)
    assert m.inner.contents.size == ref.inner.contents.size and np.array_equal(m.values(), ref.values())
    with pytest.raises(ps.FailedLookup):
        ps.Matrix.from_name("../etc/passwd")

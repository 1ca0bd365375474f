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


def test_host_scan_piece_plan(ps):
    # the pipelined host scan cuts the database so that piece k+1 lands while piece k is scanned
    MB = 1 << 20
    u = 1e3 / 52e9                      # ms per byte at 52 GB/s
    v = 27.05 / 362e6                   # C2: 27 ms of scan for 362 MB of residues
    for total in (362 * MB, 90 * MB, 45 * MB, 9 * MB, 3 * MB, 1):
        sizes = ps.host_scan_plan(total, u, v)
        assert sum(sizes) == total and all(x > 0 for x in sizes)
        r = 0.85 * v / u
        for a, b in zip(sizes, sizes[1:]):
            assert b <= a * r * 1.001 + 4 * MB          # every piece lands before its predecessor is scanned
        assert all(x >= 4 * MB for x in sizes) or len(sizes) == 1
    assert len(ps.host_scan_plan(362 * MB, u, v)) == 3 and len(ps.host_scan_plan(45 * MB, u, v)) == 2
    assert len(ps.host_scan_plan(3 * MB, u, v)) == 1
    # a slow link (28 GB/s) means smaller steps and more pieces; a link faster than the scan means one growth step
    assert len(ps.host_scan_plan(90 * MB, 1e3 / 28e9, v)) >= len(ps.host_scan_plan(90 * MB, u, v))
    # pieces are capped at 1 GB however large the database (device memory of a piece in flight)
    big = ps.host_scan_plan(40 << 30, u, v)
    assert sum(big) == 40 << 30 and max(big) <= 1 << 30
    with pytest.raises(ps.Error):
        ps.host_scan_plan(0, u, v)

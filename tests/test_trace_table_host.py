"""Host side of parasail_result_get_trace_table [REF src/alignment/mod.rs:291-303]: psb_result_extra::trace_table()
(csrc/psb_internal.h) turns the device's flag-byte block, [strip][step][lane][K] with cell (i, j) at step j + lane,
into the row-major table.  Checked on synthetic blocks by a small C++ program (tests/emu/trace_table_check.cpp); no GPU."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("ttc") / "trace_table_check")
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-o", exe, os.path.join(HERE, "emu", "trace_table_check.cpp")], check=True)
    return exe


@pytest.mark.parametrize("qlen,rlen,K", [(1, 1, 4), (300, 317, 10), (645, 63, 2), (64, 65, 1), (700, 129, 16), (2200, 2400, 8)])
def test_row_major_table_from_block(checker, qlen, rlen, K):
    r = subprocess.run([checker, str(qlen), str(rlen), str(K)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def dna_default(oracle):
    # Matrix::default() = create(b"ACGTA", 1, -1) [REF src/matrix/mod.rs:246-250]
    return oracle.Matrix.create(b"ACGTA", 1, -1)


@pytest.fixture(scope="session")
def blosum62(oracle):
    import psb_data
    return oracle.Matrix.from_table(psb_data.BLOSUM62_ALPHABET, psb_data.blosum62_table())

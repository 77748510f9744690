import lzma
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(GOLDEN, "ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ref_dir():
    return REF


@pytest.fixture(scope="session")
def words_base(tmp_path_factory):
    """Materialises words.bwt/.aux (reference fixture testdata/words.*, stored xz-compressed) in a tmp dir."""
    d = tmp_path_factory.mktemp("words")
    with lzma.open(os.path.join(REF, "words.bwt.xz"), "rb") as f:
        (d / "words.bwt").write_bytes(f.read())
    (d / "words.aux").write_bytes(open(os.path.join(REF, "words.aux"), "rb").read())
    return str(d / "words")

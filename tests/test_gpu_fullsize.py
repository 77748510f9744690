"""
Larger GPU parity runs (tens of MB; set FMX_FULLSIZE_BYTES to go bigger): GPU-built index -> count / locate / regex vs
the CPU oracle, plus the size-independent properties used at BASELINE's full sizes (located positions are real
occurrences, ascending and complete; count checksum equality between layouts).
"""
import os

import numpy as np
import pytest

from findex_b200 import fmindex as fx
from findex_b200 import synth
from oracle import fm_oracle as fo

pytestmark = pytest.mark.gpu
NBYTES = int(os.environ.get("FMX_FULLSIZE_BYTES", 48_000_000))


@pytest.fixture(scope="module")
def english(words_base, tmp_path_factory):
    o = fo.OracleIndex.load(words_base)
    sa = o.sa()
    tp = np.zeros(o.n, np.uint8)
    tp[(sa.astype(np.int64) - 1) % o.n] = o.bwt()
    vocab = [w for w in bytes(tp[:-1][::-1]).split(b"\r\n") if w]
    text = synth.english_like(vocab, NBYTES, 4).tobytes()
    base = str(tmp_path_factory.mktemp("full") / "english")
    fx.build_index_files(text, base, bigEndian=True)
    return np.frombuffer(text, np.uint8), base, fo.OracleIndex.load(base)


@pytest.mark.parametrize("layout,accel,rate", [(fx.LAYOUT_PLANES, fx.ACCEL_AUTO, 0), (fx.LAYOUT_WM, fx.ACCEL_KMER, 32), (fx.LAYOUT_PLANES, fx.ACCEL_NONE, 32)])
def test_cfg3_cfg4_scaled(english, layout, accel, rate):
    text, base, o = english
    g = fx.GpuFMSearcher(base + ".bwt", layout=layout, accel=accel, sa_sample_rate=rate)
    rng = np.random.default_rng(5)
    m, ln = 200_000, 12
    offs = rng.integers(0, len(text) - ln, m)
    pats = text[offs[:, None] + np.arange(ln - 1, -1, -1)[None, :]]
    sp, ep = g.count_fixed(pats)
    osp, oep = o.count_batch(pats.reshape(-1), np.arange(0, m * ln + 1, ln, dtype=np.int64), threads=os.cpu_count())
    assert np.array_equal(sp, osp) and np.array_equal(ep, oep)
    # locate: vs the oracle's sa on a slice, and by the text property on everything below a size cap
    occ = ep - sp
    small = np.flatnonzero(occ <= 2000)[:60_000]
    off, pos = g.locate_batch(sp[small], ep[small])
    if o.n <= 200_000_000:                                # the oracle's sa is a sequential n-step walk: skipped at full size,
        sa = o.sa().astype(np.int64)                       # where the text property below is the (complete) check
        for j in range(0, len(small), 97):
            assert np.array_equal(pos[off[j]:off[j + 1]], np.sort(sa[sp[small[j]]:ep[small[j]]]))
    n1 = g.n
    for j in range(0, len(small), 11):
        q = pos[off[j]:off[j + 1]]
        assert len(q) == occ[small[j]] and (np.diff(q) > 0).all()
        f = (n1 - 1) - q[:20] - ln
        assert (text[f[:, None] + np.arange(ln)[None, :]] == pats[small[j]][::-1][None, :]).all()
    # cfg-4 templates
    def lit(k):
        s = int(rng.integers(0, len(text) - k))
        return bytes(text[s:s + k]).replace(b"\n", b" ")
    rxs = []
    for i in range(600):
        t = i % 5
        if t == 0:
            rxs.append(lit(3) + b"[b-m]" + lit(2))
        elif t == 1:
            rxs.append(lit(3) + b"(" + lit(2) + b"|" + lit(2) + b"|" + lit(3) + b")" + lit(1))
        elif t == 2:
            x, y = lit(1), lit(1)
            rxs.append(lit(2) + x + b"?" + y + y + b"?" + y + b"?" + lit(2))
        elif t == 3:
            rxs.append(lit(3) + b"\\w" + lit(2))
        else:
            rxs.append(lit(4) + b"." + lit(2))
    got = g.regex_search_batch([fx.ReTree(r) for r in rxs])
    for rx, res in zip(rxs, got):
        assert res == o.regex_match(rx), rx
    assert sum(len(r) for r in got) > 300
    g.close()


def test_cfg2_cfg5_scaled_checksums(tmp_path):
    """random byte text and DNA text: oracle parity on a sample + full-batch checksum equality across layouts/accelerators"""
    for kind, ln in (("bytes", 16), ("dna", 32)):
        rng = np.random.default_rng(2 if kind == "bytes" else 7)
        text = rng.integers(1, 256, NBYTES, dtype=np.uint8) if kind == "bytes" else np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, NBYTES)]
        base = str(tmp_path / kind)
        fx.build_index_files(text, base, bigEndian=True)
        o = fo.OracleIndex.load(base)
        m = 400_000
        nh = int(m * 0.9)
        offs = rng.integers(0, len(text) - ln, nh)
        pats = np.empty((m, ln), np.uint8)
        pats[:nh] = text[offs[:, None] + np.arange(ln - 1, -1, -1)[None, :]]
        pats[nh:] = rng.integers(1, 256, (m - nh, ln), dtype=np.uint8) if kind == "bytes" else np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (m - nh, ln))]
        osp, oep = o.count_batch(pats[:50_000].reshape(-1), np.arange(0, 50_000 * ln + 1, ln, dtype=np.int64), threads=os.cpu_count())
        sums = set()
        for layout, accel in [(fx.LAYOUT_WM, fx.ACCEL_NONE), (fx.LAYOUT_PLANES, fx.ACCEL_NONE), (fx.LAYOUT_PLANES, fx.ACCEL_AUTO), (fx.LAYOUT_WM, fx.ACCEL_AUTO)]:
            g = fx.GpuFMSearcher(base + ".bwt", layout=layout, accel=accel)
            sp, ep = g.count_fixed(pats)
            assert np.array_equal(sp[:50_000], osp) and np.array_equal(ep[:50_000], oep)
            assert (ep[:nh] > sp[:nh]).all()
            sums.add(int((sp * 1315423911 + ep * 2654435761).sum() & 0xFFFFFFFFFFFF))
            g.close()
        assert len(sums) == 1
        o.close()

"""
GPU parity tests (run with -m gpu on a B200): every call goes through the C ABI (libfmgpu.so) and is compared
bit-exactly with the CPU oracle on the same inputs, for both rank-structure layouts and every lanes-per-query
setting.  Fixtures are the reference's own test files (tests/golden/ref) plus seeded synthetic texts.
"""
import os
import random

import numpy as np
import pytest

from findex_b200 import fmindex as fx
from oracle import fm_oracle as fo
from oracle import retree

pytestmark = pytest.mark.gpu

LAYOUTS = [fx.LAYOUT_WM, fx.LAYOUT_PLANES, fx.LAYOUT_WMX]
LANES = [1, 2, 4]
CFGS = [(l, g) for l in LAYOUTS for g in LANES]
ACCELS = [fx.ACCEL_AUTO, fx.ACCEL_NONE, fx.ACCEL_KMER, fx.ACCEL_TEXT, fx.ACCEL_CTX]


def _ids(v):
    return "%s-g%d" % ({1: "wm", 2: "planes", 3: "wmx"}[v[0]], v[1])


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    import torch
    assert torch.cuda.is_available(), "these tests need a GPU; the product has no CPU path"
    from findex_b200 import build
    build.build()


# ----------------------------------------------------------------------------------------------- fixtures
@pytest.fixture(scope="module")
def o1024(ref_dir):
    return fo.OracleIndex.load(os.path.join(ref_dir, "test1024.cmp"), big_endian=False)


def _open(path, cfg, big_endian=False, **kw):
    return fx.GpuFMSearcher(path, bigEndian=big_endian, layout=cfg[0], lanes_per_query=cfg[1], **kw)


# ----------------------------------------------------------------------------------------------- G4 on the GPU
@pytest.mark.parametrize("cfg", CFGS, ids=_ids)
def test_combined_indexing_test_on_gpu(ref_dir, cfg):
    """T/Indexer.scala:1079-1124 verbatim, against the GPU searcher."""
    sa = _open(os.path.join(ref_dir, "test1024.cmp.bwt"), cfg)
    eof = sa.eof
    assert sa.n == 1025 and eof == 462
    assert sa.getPrevI(eof) == 0
    assert sa.getNextI(eof) == 517
    assert sa.getPrevI(1) == 48
    assert sa.getPrevI(48) == 649
    assert sa.nextSubstr(1, 3) == b"haa"
    assert sa.nextSubstr(sa.getNextI(eof), 100) == \
        b"zajrtzbeqwbxdfpwjflmmsseewuudgfbtzqenjqafwzcnfanycigwsflfvxojxpqhhzekjdkhgsptqveavquuoqujbezdkarayom"
    assert sa.nextSubstr(eof, 100) == \
        b"ajrtzbeqwbxdfpwjflmmsseewuudgfbtzqenjqafwzcnfanycigwsflfvxojxpqhhzekjdkhgsptqveavquuoqujbezdkarayoml"
    assert sa.prevSubstr(1, 5) == b"bqxxa"
    assert sa.prevSubstr(eof, 5) == b"\0uexm"
    assert sa.prevSubstr(sa.getPrevI(eof), 4) == b"uexm"
    sa.close()


@pytest.mark.parametrize("cfg", CFGS, ids=_ids)
def test_g1_g2_in_memory_goldens_on_gpu(cfg):
    """T/Indexer.scala:203-351 (abracadabra, getPrevRange) through fmx_open_mem."""
    bwt, eof, cnt = fo.build_bwt(b"abracadabra")
    sa = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, layout=cfg[0], lanes_per_query=cfg[1])
    assert sa.cf(0) == 0 and sa.cf(ord("a")) == 1 and sa.cf(ord("b")) == 6
    rows = {0: [0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1], ord("a"): [1, 1, 1, 1, 1, 1, 2, 3, 4, 5, 5, 5],
            ord("b"): [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2], ord("c"): [0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1],
            ord("d"): [0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1], ord("r"): [0, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2], ord("x"): [0] * 12}
    for c, want in rows.items():
        assert sa.occ_batch([c] * 12, list(range(12))).tolist() == want
    assert sa.search(b"bra") == (6, 8)
    assert sa.getPrevI(6) == 2 and sa.getNextI(6) == 10 and sa.getNextI(10) == 1
    sa.close()
    bwt, eof, cnt = fo.build_bwt(b"mmabcacadabbbca"[::-1])
    sa = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, layout=cfg[0], lanes_per_query=cfg[1])
    assert sa.occ(ord("b"), 6) == 3
    assert sa.getPrevRange(0, 16, ord("a")) == (1, 6)
    assert sa.getPrevRange(1, 6, ord("b")) == (6, 8)
    sa.close()


# ----------------------------------------------------------------------------------------------- operator parity
@pytest.mark.parametrize("cfg", CFGS, ids=_ids)
@pytest.mark.parametrize("name", ["test1024", "test3072", "test"])
def test_operator_parity_small(ref_dir, cfg, name):
    o = fo.OracleIndex.load(os.path.join(ref_dir, name + ".cmp"), big_endian=False)
    g = _open(os.path.join(ref_dir, name + ".cmp.bwt"), cfg, sa_sample_rate=4)
    n = o.n
    assert g.n == n and g.eof == o.eof
    assert [g.cf(c) for c in range(256)] == [o.cf(c) for c in range(256)]
    # occ: every symbol of interest x every key in [-1, n-1] (+ out-of-text symbols)
    chars = sorted(set(o.bwt().tolist()) | {0, 1, 96, 123, 255})
    keys = np.arange(-1, n, dtype=np.int64)
    for c in chars:
        got = g.occ_batch(np.full(len(keys), c, np.uint8), keys)
        want = np.array([o.occ(c, int(k)) for k in keys])
        assert np.array_equal(got, want), c
    # getPrevRange on random intervals
    rng = np.random.default_rng(7)
    sp = rng.integers(0, n + 1, 4000)
    ep = rng.integers(0, n + 1, 4000)
    lo, hi = np.minimum(sp, ep), np.maximum(sp, ep)
    cc = rng.choice(chars, 4000).astype(np.uint8)
    a, b = g.prev_range_batch(lo, hi, cc)
    for i in range(4000):
        assert (int(a[i]), int(b[i])) == o.prev_range_raw(int(lo[i]), int(hi[i]), int(cc[i]))
    # getIntervalPrevRange incl. the reference's descending order
    for (s, e) in [(0, n), (10, 500), (100, 101), (5, 5)]:
        want, _ = o.getIntervalPrevRange(s, e, ord("a"), ord("z"))
        assert g.getIntervalPrevRange(s, e, ord("a"), ord("z")) == want
    # LF / FL for every row, pos2char for every key
    rows = np.arange(n, dtype=np.int64)
    assert g.get_prev_i_batch(rows).tolist() == [o.getPrevI(int(r)) for r in rows]
    assert g.get_next_i_batch(rows).tolist() == [o.getNextI(int(r)) for r in rows]
    assert [g.pos2char(int(k)) for k in range(0, n, 37)] == [o.pos2char(int(k)) for k in range(0, n, 37)]
    # extraction
    sub = rows[::17]
    assert g.prev_substr_batch(sub, 9) == [o.prevSubstr(int(r), 9) for r in sub]
    assert g.next_substr_batch(sub, 9) == [o.nextSubstr(int(r), 9) for r in sub]
    assert g.extract_batch(sub, 9, +1) == g.next_substr_batch(sub, 9) and g.extract_batch(sub, 9, -1) == g.prev_substr_batch(sub, 9)
    # locate == sorted sa[sp..ep)
    sa = o.sa().astype(np.int64)
    ivs = [(0, n), (5, 5), (17, 400), (n - 1, n), (o.eof, o.eof + 1)]
    off, pos = g.locate_batch([i[0] for i in ivs], [i[1] for i in ivs])
    for k, (s, e) in enumerate(ivs):
        assert np.array_equal(pos[off[k]:off[k + 1]], np.sort(sa[s:e]))
    g.close()
    o.close()


def _patterns(text, rng, m, maxlen):
    """hits (substrings of the file, reversed as search() wants), near-misses and random bytes"""
    pats = []
    for i in range(m):
        ln = int(rng.integers(0, maxlen + 1))
        r = i % 4
        if r < 2 and ln and len(text) > ln:
            s = int(rng.integers(0, len(text) - ln))
            p = bytes(text[s:s + ln])[::-1]
        elif r == 2 and ln and len(text) > ln:
            s = int(rng.integers(0, len(text) - ln))
            p = bytearray(bytes(text[s:s + ln])[::-1])
            p[int(rng.integers(0, ln))] = int(rng.integers(1, 256))
            p = bytes(p)
        else:
            p = bytes(rng.integers(1, 256, ln, dtype=np.uint8))
        pats.append(p)
    return pats


@pytest.mark.parametrize("accel", ACCELS, ids=lambda a: "accel%d" % a)
@pytest.mark.parametrize("cfg", CFGS, ids=_ids)
def test_count_parity_small(ref_dir, cfg, accel):
    text = open(os.path.join(ref_dir, "test.txt"), "rb").read()
    o = fo.OracleIndex.load(os.path.join(ref_dir, "test.cmp"), big_endian=False)
    g = _open(os.path.join(ref_dir, "test.cmp.bwt"), cfg, accel=accel)
    info = g.info()
    # the 16-byte isat entries (singleton shortcut) exist only where asked for: row contexts do the same in one fetch
    assert info["text_shortcut"] == (accel == fx.ACCEL_TEXT) and info["ctx_entry_bytes"] in (0, 32) and (info["kmer_k"] >= 2) == (accel in (fx.ACCEL_AUTO, fx.ACCEL_KMER))
    assert (info["ctx_depth"] > 0) == (accel in (fx.ACCEL_AUTO, fx.ACCEL_CTX))
    rng = np.random.default_rng(11)
    tprime = text[::-1]
    zero_cases = [b"", b"\0", b"a\0", bytes([200]), b"zzzzzzzzzzzzzzzzzzzzzzzzz",
                  tprime[-9:] + b"\0",                    # end of T' followed by '$': a real hit at row 0's neighbourhood
                  b"\0" + tprime[:9],                     # '$' then the start of T': the cyclic wrap the BWT allows -> row 0
                  tprime[100:106] + b"\0" + tprime[107:113], b"\0\0\0\0\0", tprime[:12], tprime[-12:], tprime[1:14]]
    pats = _patterns(text, rng, 3000, 16) + _patterns(text, rng, 1500, 40) + zero_cases
    sp, ep = g.count_batch(pats)
    for i, p in enumerate(pats):
        r = o.search(p)
        assert (int(sp[i]), int(ep[i])) == (r if r else (0, 0)), p
    assert g.search(b"") == (0, o.n)
    # fixed-length fast path, incl. ragged tail of the last CTA and odd lengths
    for ln in (1, 3, 12, 16, 21, 27, 40):
        arr = np.frombuffer(b"".join(p.ljust(ln, b"q")[:ln] for p in pats[:4499] if len(p) >= min(ln, 9)), np.uint8).reshape(-1, ln)
        sp, ep = g.count_fixed(arr)
        osp, oep = o.count_batch(arr.reshape(-1), np.arange(0, arr.size + 1, ln, dtype=np.int64))
        assert np.array_equal(sp, osp) and np.array_equal(ep, oep)
        assert np.array_equal(g.count_only_fixed(arr).astype(np.int64), oep - osp)
        s32, e32 = np.full(len(arr), -1, np.int32), np.full(len(arr), -1, np.int32)      # the reference's Int-wide results
        g.count_fixed_into(arr, s32, e32)
        assert np.array_equal(s32, osp) and np.array_equal(e32, oep)
    g.close()
    o.close()


# ----------------------------------------------------------------------------------------------- words (config 1)
@pytest.fixture(scope="module")
def words(words_base):
    o = fo.OracleIndex.load(words_base)
    sa = o.sa()
    tp = np.zeros(o.n, np.uint8)
    tp[(sa.astype(np.int64) - 1) % o.n] = o.bwt()
    return o, bytes(tp[:-1][::-1])


@pytest.mark.parametrize("cfg", [(fx.LAYOUT_WM, 4), (fx.LAYOUT_PLANES, 4), (fx.LAYOUT_PLANES, 1), (fx.LAYOUT_WM, 2), (fx.LAYOUT_WMX, 4), (fx.LAYOUT_WMX, 1)], ids=_ids)
def test_config1_words_count_and_locate(words_base, words, cfg):
    """BASELINE config 1: 1k word patterns, count + locate, vs the oracle and by brute force on words.txt."""
    o, text = words
    g = _open(words_base + ".bwt", cfg, big_endian=True, sa_sample_rate=32)
    info = g.info()
    assert info["sigma"] == 28 and (info["levels"] == {"wm": 5, "wmx": 2, "planes": 1}[info["layout"]])
    lines = text.split(b"\r\n")
    rng = np.random.default_rng(1)
    pick = rng.choice(len(lines) - 1, 1000, replace=False)
    ws = [lines[i] for i in pick]
    pats = [w[::-1] for w in ws] + ws                       # hits, and the un-reversed words as a miss/partial set
    sp, ep = g.count_batch(pats)
    osp, oep = [], []
    for p in pats:
        r = o.search(p)
        osp.append(r[0] if r else 0)
        oep.append(r[1] if r else 0)
    assert sp.tolist() == osp and ep.tolist() == oep
    for w, a, b in list(zip(ws, sp, ep))[:100]:
        assert b - a == text.count(w)
    off, pos = g.locate_batch(sp, ep)
    sa = o.sa().astype(np.int64)
    for k in range(len(pats)):
        assert np.array_equal(pos[off[k]:off[k + 1]], np.sort(sa[sp[k]:ep[k]]))
    k = 5                                                   # file offset = (n-1) - pos - len
    for q in pos[off[k]:off[k + 1]]:
        f = (o.n - 1) - int(q) - len(ws[k])
        assert text[f:f + len(ws[k])] == ws[k]
    assert g.search(b"hello"[::-1]) == (1333929, 1333938)
    assert g.search(b"ing\r\n"[::-1]) == (40318, 52882)
    assert g.search(b"zzz") is None
    g.close()


WORDS_RX = [("x(a|b|d|e)c", 90), ("ab?c[d-h]", 1668), ("q(u|a)[a-m]z?k", 22), ("th(e|a)(n|t)\\w", 363), ("b(oo|ee)+k", 158),
            ("z[aeiou][aeiou]?z", 21), ("a.*(b|c)da.*f", None), ("qu.k", None), ("q\\d", None), ("ab(cd)*ef", None), ("x[a-c]d", None)]


@pytest.mark.parametrize("cfg", [(fx.LAYOUT_WM, 4), (fx.LAYOUT_PLANES, 4), (fx.LAYOUT_PLANES, 1)], ids=_ids)
def test_regex_parity_words(words_base, words, cfg):
    o, _ = words
    g = _open(words_base + ".bwt", cfg, big_endian=True)
    trees = [fx.ReTree(rx, lineOnly=(rx == "a.*(b|c)da.*f")) for rx, _ in WORDS_RX]
    got = g.regex_search_batch(trees, cap_total=16)         # forces the FMX_E_CAPACITY retry path
    for (rx, total), res in zip(WORDS_RX, got):
        want = o.regex_match(rx, line_only=(rx == "a.*(b|c)da.*f"), max_expansions=50_000_000)
        assert res == want, rx
        if total is not None:
            assert sum(e - s for _, s, e in res) == total
    assert trees[0].matchSA(g) == got[0]
    g.close()


@pytest.mark.parametrize("cfg", CFGS, ids=_ids)
def test_regex_parity_small(ref_dir, o1024, cfg):
    g = _open(os.path.join(ref_dir, "test1024.cmp.bwt"), cfg)
    rxs = ["a.*(b|c)d.*f", "a(a|b|d|e)c", "ab", "q[a-z]q", "a.b", "x\\wy", "a(bc)*d", "k(l|m)+n", "zz?z?", "a+b", "(ab|cd)", "\\w\\w\\wq"]
    trees = [fx.ReTree(rx) for rx in rxs]
    got = g.regex_search_batch(trees)
    for rx, res in zip(rxs, got):
        assert res == o1024.regex_match(rx), rx
    # G10: .*(a|b)ca over the toy text -> exactly 2 results
    bwt, eof, cnt = fo.build_bwt(b"mmabcacamabbbca"[::-1])
    t = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, layout=cfg[0], lanes_per_query=cfg[1])
    assert fx.ReTree(".*(a|b)ca").matchSA(t) == [(3, 1, 2), (3, 2, 4)]
    t.close()
    g.close()


# ----------------------------------------------------------------------------------------------- device builder + synthetic texts
def _english_like(words_text, n_words, seed):
    rng = np.random.default_rng(seed)
    vocab = [w for w in words_text.split(b"\r\n") if w]
    perm = rng.permutation(len(vocab))
    ranks = np.minimum(rng.zipf(1.0001, n_words) - 1, len(vocab) - 1)
    out = []
    for i, r in enumerate(ranks):
        out.append(vocab[perm[r]])
        out.append(b"\n" if i % 12 == 11 else b" ")
    return b"".join(out)


def test_device_builder_reproduces_reference_goldens(ref_dir, words_base, words):
    """Any correct suffix sorter yields the unique BWT: the GPU builder must reproduce the bwtdisk goldens
    (T/Indexer.scala:638-820) and the reference's own words.bwt/.aux from the raw text."""
    for name in ["test1024", "test2048", "test2048-2", "test3072", "test", "test-part"]:
        data = open(os.path.join(ref_dir, name + ".txt"), "rb").read()
        o = fo.OracleIndex.load(os.path.join(ref_dir, name + ".cmp"), big_endian=False)
        bwt, eof, cnt = fx.build_bwt(data)
        assert eof == o.eof and np.array_equal(bwt, o.bwt())
        aux = np.frombuffer(open(os.path.join(ref_dir, name + ".cmp.aux"), "rb").read(), dtype="<i8")
        assert np.array_equal(cnt[1:], aux[1:])
    o, text = words
    bwt, eof, cnt = fx.build_bwt(text)
    assert eof == o.eof and np.array_equal(bwt, o.bwt())
    # zero bytes are dropped like FileBWTReader does; tiny / degenerate texts
    for t in [b"", b"a", b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa", b"ab\0ab\0ab", b"abababababababababab" * 50]:
        bwt, eof, cnt = fx.build_bwt(t)
        wb, weof, wcnt = fo.build_bwt(fo.file_to_text_rev(t))
        assert eof == weof and np.array_equal(bwt, wb) and np.array_equal(cnt, wcnt)


def test_index_files_roundtrip(tmp_path, ref_dir):
    data = open(os.path.join(ref_dir, "test3072.txt"), "rb").read()
    base = str(tmp_path / "t3072")
    fx.build_index_files(data, base + ".txt", bigEndian=True, write_fm=True)
    o = fo.OracleIndex.load(base)                          # reads the .fm we wrote (FMLoader rules)
    ref = fo.OracleIndex.load(os.path.join(ref_dir, "test3072.cmp"), big_endian=False)
    assert o.eof == ref.eof and np.array_equal(o.bwt(), ref.bwt()) and np.array_equal(o.fm(), ref.fm())
    g = fx.GpuFMSearcher(base + ".fm", require_fm=True)
    assert g.search(b"ab") == ref.search(b"ab")
    g.close()


@pytest.mark.parametrize("kind", ["random255", "dna", "english"])
@pytest.mark.parametrize("cfg", [(fx.LAYOUT_WM, 4, fx.ACCEL_AUTO), (fx.LAYOUT_PLANES, 4, fx.ACCEL_AUTO), (fx.LAYOUT_WM, 1, fx.ACCEL_NONE), (fx.LAYOUT_WMX, 4, fx.ACCEL_NONE), (fx.LAYOUT_WMX, 2, fx.ACCEL_AUTO),
                                 (fx.LAYOUT_PLANES, 2, fx.ACCEL_NONE), (fx.LAYOUT_PLANES, 4, fx.ACCEL_TEXT), (fx.LAYOUT_WM, 2, fx.ACCEL_KMER)],
                         ids=lambda v: _ids(v) + "-accel%d" % v[2])
def test_synthetic_parity(kind, cfg, words, tmp_path):
    """Scaled-down configs 2/3/5: seeded text -> GPU-built index files -> count / locate / regex vs oracle."""
    rng = np.random.default_rng({"random255": 2, "dna": 7, "english": 4}[kind])
    if kind == "random255":
        text = rng.integers(1, 256, 600_000, dtype=np.uint8).tobytes()
    elif kind == "dna":
        text = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 900_000)].tobytes()
    else:
        text = _english_like(words[1], 90_000, 4)
    base = str(tmp_path / kind)
    fx.build_index_files(text, base, bigEndian=True)
    o = fo.OracleIndex.load(base)
    g = _open(base + ".bwt", cfg, big_endian=True, sa_sample_rate=32, accel=cfg[2])
    info = g.info()
    if kind == "dna":
        assert info["sigma"] == 4 and (info["levels"] == {"wm": 2, "wmx": 1, "planes": 1}[info["layout"]])
    if kind == "random255":
        assert info["sigma"] == 255
    ln = {"random255": 16, "dna": 32, "english": 12}[kind]
    m = 20_000
    offs = rng.integers(0, len(text) - ln, m)
    tarr = np.frombuffer(text, np.uint8)
    pats = np.stack([tarr[s:s + ln][::-1] for s in offs])
    miss = rng.random(m) < 0.1
    pats[miss] = tarr[rng.integers(0, len(text), (int(miss.sum()), ln))] if kind != "random255" else rng.integers(1, 256, (int(miss.sum()), ln), dtype=np.uint8)
    sp, ep = g.count_fixed(pats)
    osp, oep = o.count_batch(pats.reshape(-1), np.arange(0, pats.size + 1, ln, dtype=np.int64), threads=4)
    assert np.array_equal(sp, osp) and np.array_equal(ep, oep)
    assert (ep > sp).sum() >= 0.85 * m
    blocks, steps = g.count_fixed_stats(pats)
    assert steps >= m and blocks > 0
    # locate a subset (short prefixes have many occurrences)
    sub = slice(0, 2000)
    shortp = pats[sub, ln - 4:]
    ssp, sep = g.count_fixed(shortp)
    off, pos = g.locate_batch(ssp, sep)
    sa = o.sa().astype(np.int64)
    for k in range(0, 2000, 7):
        assert np.array_equal(pos[off[k]:off[k + 1]], np.sort(sa[ssp[k]:sep[k]]))
    # regexes from the config-4 templates
    lit = lambda k: bytes(text[int(rng.integers(0, len(text) - k)):][:k])
    esc = lambda b: b"".join((b"\\" + bytes([c])) if c in b"()[]|*+?.\\-" else bytes([c]) for c in b)
    rxs = []
    for _ in range(40):
        a, b2 = sorted(rng.integers(97, 123, 2).tolist())
        rxs.append(esc(lit(3)) + b"[" + bytes([a]) + b"-" + bytes([b2]) + b"]" + esc(lit(2)))
        rxs.append(esc(lit(3)) + b"(" + esc(lit(2)) + b"|" + esc(lit(2)) + b"|" + esc(lit(1)) + b")" + esc(lit(1)))
        rxs.append(esc(lit(2)) + esc(lit(1)) + b"?" + esc(lit(1)) + esc(lit(1)) + b"?" + esc(lit(2)))
        rxs.append(esc(lit(3)) + b"\\d" + esc(lit(2)))
        rxs.append(esc(lit(4)) + b"." + esc(lit(2)))
    trees, keep = [], []
    for rx in rxs:
        try:
            trees.append(fx.ReTree(rx))
            keep.append(rx)
        except fx.FmxError:
            with pytest.raises((retree.ReUnsupported, retree.ReSyntaxError)):
                retree.compile_regex(rx)
    assert len(keep) > 100
    got = g.regex_search_batch(trees)
    for rx, res in zip(keep, got):
        assert res == o.regex_match(rx), rx
    g.close()
    o.close()


@pytest.mark.parametrize("lanes", [1, 2, 4])
def test_regex_ring_overflow_and_rerun(ref_dir, o1024, lanes):
    """6000 '.'-regexes push > 10^6 items through a work ring of 8192 slots: the traversal must notice the overflow, be rerun with
    larger rings and return every result exactly once; later searches of the same set reuse the grown ring."""
    g = _open(os.path.join(ref_dir, "test1024.cmp.bwt"), (fx.LAYOUT_PLANES, lanes))
    rxs = ["%c.%c" % (97 + i % 26, 97 + (i // 26) % 26) for i in range(6000)]
    trees = [fx.ReTree(r) for r in rxs]
    memo = {}
    rset = g.regex_set(trees)
    rset.set_ring(8192)
    for rep in range(3):
        off, ln, sp, ep = rset.search(g, cap_total=64 if rep == 0 else 1 << 20)
        assert g.last_kernel_launches() >= (4 if rep == 0 else 2)             # rep 0: at least one abandoned run (seed + traversal each)
        assert g.last_regex_levels() == 3
        for i, rx in enumerate(rxs):
            if rx not in memo:
                memo[rx] = o1024.regex_match(rx)
            assert list(zip(ln[off[i]:off[i + 1]].tolist(), sp[off[i]:off[i + 1]].tolist(), ep[off[i]:off[i + 1]].tolist())) == memo[rx], rx
        assert off[-1] > 1000
    rset.close()
    g.close()


def test_thompson_max_length(ref_dir, o1024):
    """REParser.matchSA(nfa, sa, maxLength = L) (re2.scala:568, :636-641): follow positions are enqueued only while len < L, matches are
    emitted whatever their length — same multiset as the oracle for L = 1..6 and unlimited, Thompson engine; same meaning on Glushkov sets"""
    g = _open(os.path.join(ref_dir, "test1024.cmp.bwt"), (fx.LAYOUT_PLANES, 2))
    rxs = ["ab*c", "a.*b", "(a|b)c*d", "x.y.z", "qu*", "a(b|c)*d"]
    trees = [fx.ThompsonNFA(r) for r in rxs]
    rset = g.regex_set(trees)
    for L in (1, 2, 3, 4, 6, 0):
        rset.set_limits(L)
        off, ln, sp, ep = rset.search(g, cap_total=1 << 22)
        for i, rx in enumerate(rxs):
            if L == 0 and ".*" in rx:
                continue                                    # unlimited '.*' walks every substring: too slow for the scalar oracle
            got = list(zip(ln[off[i]:off[i + 1]].tolist(), sp[off[i]:off[i + 1]].tolist(), ep[off[i]:off[i + 1]].tolist()))
            assert got == o1024.regex_match_thompson(rx, max_expansions=20_000_000, max_len=L), (rx, L)
    rset.close()
    gl = ["ab*c", "a.*b", "x(a|b)c*d"]
    rset = g.regex_set([fx.ReTree(r) for r in gl])
    from oracle import retree as _rt
    for L in (2, 3, 0):
        rset.set_limits(L)
        off, ln, sp, ep = rset.search(g, cap_total=1 << 22)
        for i, rx in enumerate(gl):
            if L == 0 and ".*" in rx:
                continue
            got = list(zip(ln[off[i]:off[i + 1]].tolist(), sp[off[i]:off[i + 1]].tolist(), ep[off[i]:off[i + 1]].tolist()))
            assert got == o1024.regex_match_tables(_rt.compile_regex(rx).tables(), 20_000_000, True, L)[0], (rx, L)
    rset.close()
    g.close()


def test_regex_device_resident_results(ref_dir, o1024):
    """fmx_regex_set_search_dev: the ordered {regex, len, sp, ep} records and per-regex offsets, left on the device, equal what the host call
    returns — for a handful of results (one CTA's bitonic network) and for many (radix passes)."""
    import torch
    g = _open(os.path.join(ref_dir, "test1024.cmp.bwt"), (fx.LAYOUT_WM, 2))
    for rxs in (["ab.", "q(u|a)x?", "zz", "a", "x\\w\\w"], ["%c.%c?" % (97 + i % 26, 97 + (i // 26) % 26) for i in range(676)]):
        trees = [fx.ReTree(r) for r in rxs]
        rset = g.regex_set(trees)
        off, ln, sp, ep = rset.search(g)
        cap = int(off[-1]) + 5
        d_res = torch.zeros((cap, 4), dtype=torch.int32, device="cuda")
        d_off = torch.zeros(len(rxs) + 1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()                       # stream contract of fmx_regex_set_search_dev (fmgpu.h)
        total = rset.search_dev(g, d_res.data_ptr(), cap, d_off.data_ptr())
        torch.cuda.synchronize()
        assert total == off[-1] and np.array_equal(d_off.cpu().numpy(), off)
        r = d_res.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        assert np.array_equal(r[:total, 1], ln) and np.array_equal(r[:total, 2], sp) and np.array_equal(r[:total, 3], ep)
        assert np.array_equal(r[:total, 0], np.repeat(np.arange(len(rxs)), np.diff(off)))
        with pytest.raises(fx.FmxError) as ei:
            rset.search_dev(g, d_res.data_ptr(), max(total - 1, 0), 0)
        assert ei.value.code == fx.FMX_E_CAPACITY or total == 0
        for i, rx in enumerate(rxs[:40]):
            assert list(zip(ln[off[i]:off[i + 1]].tolist(), sp[off[i]:off[i + 1]].tolist(), ep[off[i]:off[i + 1]].tolist())) == o1024.regex_match(rx), rx
        rset.close()
    g.close()


@pytest.mark.parametrize("rate,accel", [(4, fx.ACCEL_NONE), (32, fx.ACCEL_KMER), (0, fx.ACCEL_AUTO)])
def test_locate_device_resident_and_slabs(ref_dir, rate, accel):
    """fmx_locate_dev (uint32 rows in, int64 offsets + uint32 positions out, all on the device) equals fmx_locate_batch and the oracle's
    sorted sa[sp..ep); with slabs of 1000 occurrences the batch is cut at query boundaries (one query alone exceeds a slab)."""
    import torch
    text = open(os.path.join(ref_dir, "test.txt"), "rb").read()
    o = fo.OracleIndex.load(os.path.join(ref_dir, "test.cmp"), big_endian=False)
    sa = o.sa().astype(np.int64)
    g = fx.GpuFMSearcher(os.path.join(ref_dir, "test.cmp.bwt"), bigEndian=False, sa_sample_rate=rate, accel=accel)
    rng = np.random.default_rng(8)
    pats = [b"", b"a", b"e"] + [text[s:s + int(rng.integers(1, 4))][::-1] for s in rng.integers(0, len(text) - 4, 400)] + [b"zzzz", b"qq"]
    sp, ep = g.count_batch(pats)
    want = [np.sort(sa[a:b]) for a, b in zip(sp, ep)]
    try:
        for slab in (0, 1000, 7):
            fx.lib().fmx_set_locate_slab(slab)
            off, pos = g.locate_batch(sp, ep)
            assert off[-1] == (ep - sp).sum() > 10241
            for k in range(len(pats)):
                assert np.array_equal(pos[off[k]:off[k + 1]], want[k]), (slab, k)
            d_sp = torch.from_numpy(sp.astype(np.uint32).view(np.int32)).cuda()
            d_ep = torch.from_numpy(ep.astype(np.uint32).view(np.int32)).cuda()
            d_off = torch.zeros(len(pats) + 1, dtype=torch.int64, device="cuda")
            d_pos = torch.zeros(int(off[-1]) + 3, dtype=torch.int32, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            with pytest.raises(fx.FmxError) as ei:
                g.locate_dev(d_sp.data_ptr(), d_ep.data_ptr(), len(pats), d_off.data_ptr(), d_pos.data_ptr(), int(off[-1]) - 1, st)
            assert ei.value.code == fx.FMX_E_CAPACITY
            total = g.locate_dev(d_sp.data_ptr(), d_ep.data_ptr(), len(pats), d_off.data_ptr(), d_pos.data_ptr(), d_pos.numel(), st)
            torch.cuda.synchronize()
            assert total == off[-1] and np.array_equal(d_off.cpu().numpy(), off)
            assert np.array_equal(d_pos.cpu().numpy()[:total].astype(np.int64) & 0xFFFFFFFF, pos)
            walk, sort = g.last_locate_ms()
            assert walk > 0 and sort > 0
    finally:
        fx.lib().fmx_set_locate_slab(0)
    g.close()
    o.close()


@pytest.mark.parametrize("sigma", [2, 4])
def test_packed_two_bit_upload(sigma):
    """fmx_count_fixed_packed2: 2-bit codes on the wire, expanded on the device — same (sp, ep) as the byte patterns, uint32 and int64 rows,
    lengths that are not multiples of 4, batches spanning several pipeline chunks"""
    rng = np.random.default_rng(40 + sigma)
    alpha = np.frombuffer(b"ACGT", np.uint8)[:sigma]
    text = alpha[rng.integers(0, sigma, 60_000)].tobytes()
    tp = bytes(fo.file_to_text_rev(text))
    bwt, eof, cnt = fo.build_bwt(fo.file_to_text_rev(text))
    o = fo.OracleIndex.from_bwt(bwt, eof, cnt)
    g = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt)
    g.set_chunk(1000)
    for ln in (1, 7, 16, 32, 45):
        offs = rng.integers(0, len(tp) - ln, 3500)
        arr = np.stack([np.frombuffer(tp[s:s + ln], np.uint8) for s in offs]).copy()
        arr[::6, ln // 2] = alpha[rng.integers(0, sigma, len(arr[::6]))]
        osp, oep = o.count_batch(arr.reshape(-1), np.arange(0, arr.size + 1, ln, dtype=np.int64))
        codes = g.pack2(arr)
        assert codes.shape == (len(arr), (ln + 3) // 4)
        for dt in (np.uint32, np.int64):
            sp, ep = np.full(len(arr), 7, dt), np.full(len(arr), 7, dt)
            g.count_packed2_into(codes, ln, sp, ep)
            assert np.array_equal(sp.astype(np.int64), osp) and np.array_equal(ep.astype(np.int64), oep), (ln, dt)
    g.close()
    o.close()


def test_lcp_creator_and_searcher(ref_dir, tmp_path):
    """fmx_build_lcp == bwtFm2LCP (util.scala:153-212) and fmx_write_lcp_file == LCPCreator.create's file (the equality the reference's
    LCPLoaderTest asserts, T/Indexer.scala:1017-1042: file == bwtFm2LCP sliced to the file's length), on the reference's fixtures and on the
    kind of text that test indexes (testdata/t2: digits); LCPSearcher.getLCP / getStringOn over the files."""
    import shutil
    rng = np.random.default_rng(21)
    digits = bytes(rng.integers(48, 58, 5000, dtype=np.uint8)) + b"0123456789" * 40 + bytes(rng.integers(48, 58, 700, dtype=np.uint8))
    cases = [("test1024.cmp", None), ("test.cmp", None), (None, digits), (None, b"a"), (None, b"ab"), (None, b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa" * 20)]
    for name, text in cases:
        if name:
            o = fo.OracleIndex.load(os.path.join(ref_dir, name), big_endian=False)
            for ext in (".bwt", ".aux"):
                shutil.copy(os.path.join(ref_dir, name + ext), tmp_path / ("c" + ext))
            g = fx.GpuFMSearcher(str(tmp_path / "c.bwt"), bigEndian=False)
        else:
            bwt, eof, cnt = fo.build_bwt(fo.file_to_text_rev(text))
            o = fo.OracleIndex.from_bwt(bwt, eof, cnt)
            g = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, accel=fx.ACCEL_NONE)
        want = o.lcp()
        assert np.array_equal(g.build_lcp(), want), name or text[:20]
        g.write_lcp_file(str(tmp_path / "c"))
        f = np.fromfile(tmp_path / "c.lcp", dtype=">i4")
        assert len(f) == max(o.n - 1, 1) and np.array_equal(f, want[:len(f)])
        g.close()
        o.close()
    o = fo.OracleIndex.load(os.path.join(ref_dir, "test1024.cmp"), big_endian=False)
    for ext in (".bwt", ".aux"):
        shutil.copy(os.path.join(ref_dir, "test1024.cmp" + ext), tmp_path / ("s" + ext))
    ls = fx.LCPSearcher(str(tmp_path / "s.bwt"), bigEndian=False)
    sa, lcp = o.sa(), o.lcp()
    text = open(os.path.join(ref_dir, "test1024.txt"), "rb").read()
    for i in (0, 1, 5, 462, 700, 1023):
        assert ls.getLCP(i) == lcp[i] and ls.getSA(i) == sa[i]
        # the reference reads the forward text file from offset fsize - sa[i] up to a 0 byte (bwtmerger.scala:329-332)
        assert "".join(ls.getStringOn(i, chunk=64)) == text[len(text) - int(sa[i]):].decode("latin-1")
    ls.close()
    o.close()


def test_concurrent_host_threads(ref_dir):
    """An opened index is immutable: batch calls of several host threads run concurrently (each on its own stream set) and every one
    returns exactly what it returns alone — count (all three widths), locate, regex, LF/extraction mixed over 8 threads."""
    import threading
    text = open(os.path.join(ref_dir, "test.txt"), "rb").read()
    o = fo.OracleIndex.load(os.path.join(ref_dir, "test.cmp"), big_endian=False)
    g = fx.GpuFMSearcher(os.path.join(ref_dir, "test.cmp.bwt"), bigEndian=False, sa_sample_rate=8)
    g.set_chunk(500)
    rng = np.random.default_rng(77)
    arr = np.stack([np.frombuffer(text[::-1][s:s + 9], np.uint8) for s in rng.integers(0, len(text) - 9, 4000)]).copy()
    arr[::4, 3] = ord("#")
    osp, oep = o.count_batch(arr.reshape(-1), np.arange(0, arr.size + 1, 9, dtype=np.int64))
    sa = o.sa().astype(np.int64)
    want_loc = np.concatenate([np.sort(sa[a:b]) for a, b in zip(osp[:300], oep[:300])])
    rxs = ["ab.", "q(u|a)x?", "a.*b", "x[a-f]y"]
    want_rx = [o.regex_match(r) for r in rxs]
    rows = rng.integers(0, o.n, 200)
    want_lf = g.get_prev_i_batch(rows)
    errs = []

    def worker(kind):
        try:
            for _ in range(6):
                if kind == 0:
                    sp, ep = g.count_fixed(arr)
                    assert np.array_equal(sp, osp) and np.array_equal(ep, oep)
                elif kind == 1:
                    assert np.array_equal(g.count_only_fixed(arr).astype(np.int64), oep - osp)
                elif kind == 2:
                    off, pos = g.locate_batch(osp[:300], oep[:300])
                    assert np.array_equal(pos, want_loc)
                elif kind == 3:
                    assert g.regex_search_batch([fx.ReTree(r) for r in rxs]) == want_rx
                else:
                    assert np.array_equal(g.get_prev_i_batch(rows), want_lf)
        except Exception as e:                              # noqa: BLE001
            errs.append((kind, repr(e)))
    ths = [threading.Thread(target=worker, args=(k % 5,)) for k in range(8)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    g.close()
    o.close()


def test_scatter_dev_single_gpu():
    """fmx_scatter_dev with this GPU playing three ranks: every rank's slab lands in every gathered buffer at the offset read from the device"""
    import torch
    rng = np.random.default_rng(1)
    sizes = [1001, 0, 4099]
    slabs = [torch.from_numpy(rng.integers(0, 2 ** 31, s).astype(np.int32)).cuda() for s in sizes]
    bufs = [torch.full((sum(sizes) + 8,), -1, dtype=torch.int32, device="cuda") for _ in range(3)]
    offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for r in range(3):
        fx.scatter_dev(slabs[r].data_ptr(), sizes[r], [b.data_ptr() for b in bufs], 3, offs.data_ptr() + 8 * r, 1, st)
    torch.cuda.synchronize()
    want = np.concatenate([np.full(3, -1, np.int32)] + [s.cpu().numpy() for s in slabs] + [np.full(5, -1, np.int32)])
    for b in bufs:
        assert np.array_equal(b.cpu().numpy(), want)


@pytest.mark.parametrize("m,lanes", [(3000, 2), (3003, 2), (3000, 1), (3003, 4)])
def test_fused_count_and_exchange_single_gpu(ref_dir, m, lanes):
    """The fused count+all-gather entry point with this GPU playing every rank: three 'ranks' shard a batch and each stores its
    hit counts into all three gathered buffers; every buffer must end up holding the counts of the whole batch."""
    import torch
    text = open(os.path.join(ref_dir, "test.txt"), "rb").read()
    o = fo.OracleIndex.load(os.path.join(ref_dir, "test.cmp"), big_endian=False)
    g = _open(os.path.join(ref_dir, "test.cmp.bwt"), (fx.LAYOUT_PLANES, lanes))
    rng = np.random.default_rng(3)
    ln, world = 6, 3                                        # m = 3003: shard offsets 1001, 2002 are not 16-byte aligned (scalar stores)
    offs = rng.integers(0, len(text) - ln, m)
    tarr = np.frombuffer(text, np.uint8)
    pats = np.stack([tarr[s:s + ln][::-1] for s in offs])
    pats[::5] = rng.integers(97, 123, (len(pats[::5]), ln), dtype=np.uint8)
    osp, oep = o.count_batch(pats.reshape(-1), np.arange(0, m * ln + 1, ln, dtype=np.int64))
    bufs = [fx.SharedDeviceBuffer(m) for _ in range(world)]
    d_pat = torch.from_numpy(pats).cuda()
    d_sp = torch.zeros(m, dtype=torch.int32, device="cuda")
    d_ep = torch.zeros(m, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    per = m // world
    for r in range(world):
        lo, hi = r * per, (r + 1) * per
        g.count_fixed_dev_gather(d_pat.data_ptr() + lo * ln, ln, hi - lo, d_sp.data_ptr() + lo * 4, d_ep.data_ptr() + lo * 4,
                                 [b.ptr for b in bufs], lo, st)
    torch.cuda.synchronize()
    assert np.array_equal(d_sp.cpu().numpy(), osp) and np.array_equal(d_ep.cpu().numpy(), oep)
    for b in bufs:
        assert np.array_equal(b.to_host().astype(np.int64), oep - osp)
        b.close()
    g.close()


# ----------------------------------------------------------------------------------------------- Thompson engine
@pytest.mark.parametrize("cfg", [(fx.LAYOUT_WM, 4), (fx.LAYOUT_PLANES, 2), (fx.LAYOUT_PLANES, 1)], ids=_ids)
def test_thompson_engine_goldens_and_parity(ref_dir, o1024, cfg):
    g = _open(os.path.join(ref_dir, "test1024.cmp.bwt"), cfg)
    # T/REParser.scala:292-307 on the GPU: Set("ec", "dc", "[2 Results] ac", "bc")
    r = fx.ThompsonNFA("(((b|a)|d)|e)c").matchSA(g)
    shown = {("[%d Results] " % (e - s) if e - s > 1 else "") + g.nextSubstr(s, l).decode() for l, s, e in r}
    assert shown == {"ec", "dc", "[2 Results] ac", "bc"}
    rxs = ["mab", "(b|a)c", "ab*", "ab*c", "a+b", "a(bc)*d", "(ab|cd)+e", "a.b", "a\\db", "x\\w+y", "a?b", "(a|b)*c", "a(b|c)?d", "q(u|a).z?k",
           "a.?b", "((a|b)*aba*)*(a|b)(a|b)", "zz*", "a(b|c)(d|e)f?"]
    got = g.regex_search_batch([fx.ThompsonNFA(rx) for rx in rxs])
    for rx, res in zip(rxs, got):
        assert res == o1024.regex_match_thompson(rx, max_expansions=30_000_000), rx
    # both engines in one batch
    mixed = [fx.ReTree("a(a|b|d|e)c"), fx.ThompsonNFA("a(a|b|d|e)c"), fx.ReTree("ab*"), fx.ThompsonNFA("ab*")]
    res = g.regex_search_batch(mixed)
    assert res[0] == o1024.regex_match("a(a|b|d|e)c") and res[1] == o1024.regex_match_thompson("a(a|b|d|e)c")
    assert res[2] == o1024.regex_match("ab*") and res[3] == o1024.regex_match_thompson("ab*") and len(res[3]) > len(res[2])
    g.close()
    # T/REParser.scala:219-234 (toy SA)
    bwt, eof, cnt = fo.build_bwt(b"mmabcacamabbbca"[::-1])
    t = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, layout=cfg[0], lanes_per_query=cfg[1])
    assert fx.ThompsonNFA("mab").matchSA(t) == [(3, 6, 8)]
    assert fx.ThompsonNFA("(b|a)c").matchSA(t) == [(2, 10, 11), (2, 11, 13)]
    t.close()


def test_thompson_parity_words(words_base, words):
    o, _ = words
    g = _open(words_base + ".bwt", (fx.LAYOUT_PLANES, 2), big_endian=True)
    rxs = ["x(a|b|d|e)c", "th(e|a)(n|t)\\w", "b(oo|ee)+k", "qu.k", "q\\d", "ab(cd)*ef", "ing\r\n", "z(a|e|i|o|u)(a|e|i|o|u)?z"]
    got = g.regex_search_batch([fx.ThompsonNFA(rx) for rx in rxs])
    for rx, res in zip(rxs, got):
        assert res == o.regex_match_thompson(rx, max_expansions=50_000_000), rx
    assert sum(e - s for _, s, e in got[0]) == 90
    g.close()


# ----------------------------------------------------------------------------------------------- DFA engine (dfa.scala)
@pytest.mark.parametrize("cfg", [(fx.LAYOUT_WM, 2), (fx.LAYOUT_PLANES, 2), (fx.LAYOUT_PLANES, 1)], ids=_ids)
def test_dfa_match_sa(cfg, words_base, words):
    """DFA.matchSA on the GPU: the reference's asserted toy case (T/dfa.scala:108-120) verbatim, the bucket quirk, random automata over
    the toy text and over words.*, and DFA handles mixed with ReTree / Thompson handles in one batch."""
    import random
    from dfa_cases import ab_star_c, class_b_star_c
    from findex_b200 import dfa as pd
    from oracle import dfa as od
    bwt, eof, cnt = fo.build_bwt(b"mmabcacadabbbca"[::-1])
    o = fo.OracleIndex.from_bwt(bwt, eof, cnt)
    g = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, layout=cfg[0], lanes_per_query=cfg[1])
    res = pd.DFA.processLinkList(ab_star_c(pd)).matchSA(g)
    assert res == od.DFA(ab_star_c(od)).matchSA(o) and len(res) == 2
    assert sorted(g.nextSubstr(sp, ln)[::-1] for ln, sp, ep in res) == [b"cba", b"cbbba"]          # results(1), results(0) of the reference test
    assert pd.DFA(class_b_star_c(pd, ["c", "d"])).matchSA(g) == []                                  # DFABucket actions are never followed

    def rand(m, rng, alphabet):
        s = m.StartState()
        st = [s] + [m.FinishState() if rng.random() < 0.4 else m.State(str(i)) for i in range(rng.randrange(1, 7))]
        for _ in range(rng.randrange(1, 25)):
            st[rng.randrange(len(st))].link(st[rng.randrange(len(st))], rng.choice(alphabet))
        return s
    for seed in range(40):
        a = pd.DFA(rand(pd, random.Random(seed), list(b"abcdm")))
        b = od.DFA(rand(od, random.Random(seed), list(b"abcdm")))
        assert a.matchSA(g) == b.matchSA(o), seed
    # DFA.fromNFA (dfa.scala:343-389): `a(b|c)*` + `c` as an NFA with epsilon links, determinised in the library
    def nfa(m):
        s, x, y, f = m.NfaStartState(), m.NfaState(), m.NfaState(), m.NfaFinishState()
        s.link(x, "a"); x.epsilon(y); y.link(y, "b"); y.link(y, "c"); y.link(f, "c")
        return s
    assert pd.DFA.fromNFA(nfa(pd)).matchSA(g) == od.DFA(od.from_nfa(nfa(od))).matchSA(o)
    g.close()
    o.close()

    ow, _ = words
    gw = _open(words_base + ".bwt", cfg, big_endian=True)
    autos_p, autos_o = [], []
    for seed in range(12):
        alphabet = list(b"aeitn\r\n")
        autos_p.append(pd.DFA(rand(pd, random.Random(100 + seed), alphabet)))
        autos_o.append(od.DFA(rand(od, random.Random(100 + seed), alphabet)))
    batch = autos_p + [fx.ReTree("th(e|a)(n|t)\\w"), fx.ThompsonNFA("qu.k")]
    got = gw.regex_search_batch(batch)
    nonempty = 0
    for a, b in zip(got[:12], autos_o):
        # a cyclic automaton matches ever longer strings until the text runs out of them: bounded here by the text, not by a cap
        assert a == b.matchSA(ow)
        nonempty += bool(a)
    assert nonempty >= 3
    assert got[12] == ow.regex_match("th(e|a)(n|t)\\w") and got[13] == ow.regex_match_thompson("qu.k", max_expansions=50_000_000)
    gw.close()


# ----------------------------------------------------------------------------------------------- row contexts
@pytest.mark.parametrize("sigma,form", [(2, 32), (4, 32), (20, 32), (200, 32), (2, 8), (3, 8), (4, 8)])
@pytest.mark.parametrize("cfg", [(fx.LAYOUT_WM, 4), (fx.LAYOUT_PLANES, 2), (fx.LAYOUT_PLANES, 1), (fx.LAYOUT_WM, 1), (fx.LAYOUT_WMX, 4)], ids=_ids)
def test_row_context_small_intervals(cfg, sigma, form):
    """FMX_ACCEL_CTX / _CTX8: intervals of 1..8 rows (and larger ones) with every remaining length around the covered window, on texts
    whose chunks repeat 2..12 times; all three symbol packings of the 32-byte form (3, 5 and 8 bits) and the compact 8-byte form"""
    rng = np.random.default_rng(100 + sigma)
    alpha = rng.choice(np.arange(1, 256), sigma, replace=False).astype(np.uint8)
    chunks = [alpha[rng.integers(0, sigma, int(rng.integers(20, 70)))] for _ in range(40)]
    parts = []
    for _ in range(300):
        parts.append(chunks[int(rng.integers(0, len(chunks)))] if rng.random() < 0.6 else alpha[rng.integers(0, sigma, int(rng.integers(1, 50)))])
    text = np.concatenate(parts).tobytes()
    tp = bytes(fo.file_to_text_rev(text))
    bwt, eof, cnt = fo.build_bwt(fo.file_to_text_rev(text))
    o = fo.OracleIndex.from_bwt(bwt, eof, cnt)
    g = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, layout=cfg[0], lanes_per_query=cfg[1],
                         accel=(fx.ACCEL_CTX if form == 32 else fx.ACCEL_CTX8) | fx.ACCEL_KMER, kmer_table_bytes=8 * sigma * sigma)
    info = g.info()
    J = info["ctx_depth"]
    assert info["ctx_entry_bytes"] == form and info["kmer_k"] == 2
    assert J == ({2: 48, 4: 32, 20: 19, 200: 12}[sigma] if form == 32 else 16)
    assert not info["text_shortcut"]                        # row contexts are built without the 16-byte isat entries
    pats = []
    for ln in list(range(1, J + 8)) + [J + 20, 2 * J + 1, 2 * J + 5, 3 * J + 2, 131, 150]:      # every hop plan, and beyond the plan table
        for _ in range(60):
            s = int(rng.integers(0, len(tp) - ln))
            p = bytearray(tp[s:s + ln])
            if rng.random() < 0.3:
                p[int(rng.integers(0, ln))] = int(alpha[rng.integers(0, sigma)])       # near miss
            pats.append(bytes(p))
    pats += [tp[:J], tp[:J + 2], tp[1:J + 1], tp[-J - 1:-1], bytes([int(alpha[0])]) * (J + 2), tp[:J + 1] + b"\0", b"\0" + tp[:J + 1]]
    sp, ep = g.count_batch(pats)
    sizes = set()
    for i, p in enumerate(pats):
        r = o.search(p)
        assert (int(sp[i]), int(ep[i])) == (r if r else (0, 0)), p
        sizes.add(int(ep[i] - sp[i]))
    assert len(sizes & {1, 2, 3, 4, 5, 6, 7, 8}) >= (3 if sigma > 2 else 1) and max(sizes) > 8
    for ln in (J - 4, J, J + 2, J + 3, 2 * J + 4, 2 * J + 5, 132):
        arr = np.frombuffer(b"".join(p[:ln] for p in pats if len(p) >= ln), np.uint8).reshape(-1, ln)
        s2, e2 = g.count_fixed(arr)
        osp, oep = o.count_batch(arr.reshape(-1), np.arange(0, arr.size + 1, ln, dtype=np.int64))
        assert np.array_equal(s2, osp) and np.array_equal(e2, oep)
    off, pos = g.locate_batch(sp[:40], ep[:40])             # the full SA stays resident with either form
    sa = o.sa().astype(np.int64)
    for k in range(40):
        assert np.array_equal(pos[off[k]:off[k + 1]], np.sort(sa[sp[k]:ep[k]]))
    g.close()
    o.close()


@pytest.mark.parametrize("sigma", [4, 26, 255])
def test_every_length_hops(sigma):
    """Row-context hops: patterns of every length 1..100 (hits, near misses, interval sizes 1..8 and wider) cost the same (sp, ep) as
    stepping, for the three symbol packings, with a k-mer table in front (AUTO) and without (CTX alone)."""
    rng = np.random.default_rng(7 + sigma)
    alpha = np.arange(1, 256, dtype=np.uint8) if sigma == 255 else rng.choice(np.arange(1, 256), sigma, replace=False).astype(np.uint8)
    unit = alpha[rng.integers(0, sigma, 3000)]
    text = np.concatenate([unit, alpha[rng.integers(0, sigma, 500)], unit[500:2500], alpha[rng.integers(0, sigma, 300)], unit[1000:1400],
                           unit[1000:1400], unit[1000:1400]]).tobytes()
    tp = bytes(fo.file_to_text_rev(text))
    bwt, eof, cnt = fo.build_bwt(fo.file_to_text_rev(text))
    o = fo.OracleIndex.from_bwt(bwt, eof, cnt)
    for accel, lanes in ((fx.ACCEL_AUTO, 0), (fx.ACCEL_CTX, 1), (fx.ACCEL_CTX, 4)):
        g = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, accel=accel, lanes_per_query=lanes)
        assert g.info()["ctx_entry_bytes"] == 32
        for ln in range(1, 101):
            offs = rng.integers(0, len(tp) - ln, 64)
            arr = np.stack([np.frombuffer(tp[s:s + ln], np.uint8) for s in offs]).copy()
            arr[::5, int(rng.integers(0, ln))] = alpha[rng.integers(0, sigma)]           # near misses
            sp, ep = g.count_fixed(arr)
            osp, oep = o.count_batch(arr.reshape(-1), np.arange(0, arr.size + 1, ln, dtype=np.int64))
            assert np.array_equal(sp, osp) and np.array_equal(ep, oep), ln
        g.close()
    o.close()


# ----------------------------------------------------------------------------------------------- dictionary of wide intervals
def _zipf_text(rng, alpha, n_words, n_vocab):
    """words of 2..9 symbols drawn Zipf(1) over a vocabulary, separated by alpha[0]: k-mers that stay frequent far beyond a shallow table"""
    vocab = [alpha[1 + rng.integers(0, len(alpha) - 1, int(rng.integers(2, 10)))] for _ in range(n_vocab)]
    p = 1.0 / np.arange(1, n_vocab + 1)
    idx = rng.choice(n_vocab, n_words, p=p / p.sum())
    return np.concatenate([np.concatenate([vocab[i], alpha[:1]]) for i in idx]).tobytes()


@pytest.mark.parametrize("sigma,min_rows", [(4, 8), (27, 8), (27, 1), (60, 3), (255, 2)])
@pytest.mark.parametrize("cfg", [(fx.LAYOUT_PLANES, 2), (fx.LAYOUT_PLANES, 1), (fx.LAYOUT_WM, 4), (fx.LAYOUT_WMX, 4)], ids=_ids)
def test_wide_interval_dictionary(cfg, sigma, min_rows):
    """FMX_ACCEL_DICT: (sp, ep) of every wide d-mer beyond the dense table in a hash table, deepest stored prefix by bisection — same
    intervals as stepping for every pattern length, for hits (wide and narrow), near misses, absent symbols and byte 0; with and
    without row contexts behind it; 2-, 5-, 6- and 8-bit keys."""
    rng = np.random.default_rng(900 + sigma + min_rows)
    alpha = np.arange(1, 256, dtype=np.uint8) if sigma == 255 else rng.choice(np.arange(1, 255), sigma, replace=False).astype(np.uint8)
    text = _zipf_text(rng, alpha, 6000, 300)
    tp = bytes(fo.file_to_text_rev(text))
    bwt, eof, cnt = fo.build_bwt(fo.file_to_text_rev(text))
    o = fo.OracleIndex.from_bwt(bwt, eof, cnt)
    bits = max(1, int(np.ceil(np.log2(len(set(text))))))
    absent = bytes([255]) if sigma < 255 else None
    for accel in (fx.ACCEL_KMER | fx.ACCEL_DICT, fx.ACCEL_KMER | fx.ACCEL_DICT | fx.ACCEL_CTX):
        g = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, layout=cfg[0], lanes_per_query=cfg[1], accel=accel,
                             kmer_table_bytes=8 * len(set(text)) ** 2, dict_min_rows=min_rows)
        info = g.info()
        assert info["kmer_k"] == 2 and info["dict_depth"] == min(16, 60 // bits) and info["dict_entries"] > 100, info
        assert info["dict_bytes"] >= 16 * info["dict_entries"] * 2 and info["index_bytes"] > info["dict_bytes"]
        assert info["dict_chain_depth"] >= info["dict_depth"] + (4 if min_rows <= 3 else 0), info       # chain entries behind the deepest keyed level
        pats = []
        for ln in list(range(1, 24)) + [30, 47]:
            for _ in range(80):
                s = int(rng.integers(0, len(tp) - ln))
                q = bytearray(tp[s:s + ln])
                u = rng.random()
                if u < 0.25:
                    q[int(rng.integers(0, ln))] = int(alpha[rng.integers(0, sigma)])       # near miss
                elif u < 0.30 and absent:
                    q[int(rng.integers(0, ln))] = absent[0]
                elif u < 0.33:
                    q[int(rng.integers(0, ln))] = 0
                pats.append(bytes(q))
        sp, ep = g.count_batch(pats)
        wide = 0
        for i, q in enumerate(pats):
            r = o.search(q)
            assert (int(sp[i]), int(ep[i])) == (r if r else (0, 0)), (q, accel)
            wide += int(ep[i] - sp[i]) > min_rows and len(q) > 2
        assert wide > 200
        for ln in (3, 5, 8, 12, 13, 16, 17, 21, 30):
            arr = np.frombuffer(b"".join(q[-ln:] for q in pats if len(q) >= ln), np.uint8).reshape(-1, ln)
            s2, e2 = g.count_fixed(arr)
            osp, oep = o.count_batch(arr.reshape(-1), np.arange(0, arr.size + 1, ln, dtype=np.int64))
            assert np.array_equal(s2, osp) and np.array_equal(e2, oep), ln
        # fewer requests than without the dictionary, and hiding it changes nothing but that
        arr = np.frombuffer(b"".join(q[-12:] for q in pats if len(q) >= 12), np.uint8).reshape(-1, 12)
        with_dict, _ = g.count_fixed_stats(arr)
        g.set_accel_mask(accel & ~fx.ACCEL_DICT)
        assert g.info()["dict_depth"] == 0
        without, _ = g.count_fixed_stats(arr)
        s3, e3 = g.count_fixed(arr)
        osp, oep = o.count_batch(arr.reshape(-1), np.arange(0, arr.size + 1, 12, dtype=np.int64))
        assert np.array_equal(s3, osp) and np.array_equal(e3, oep)
        if min_rows <= 2:                                    # (a high threshold on a text this small stores too little to pay for its probes)
            assert with_dict < without, (with_dict, without)
        g.close()
    o.close()


def test_dictionary_budget_and_uniform_text():
    """dict_bytes bounds the table (whole levels are dropped from the deep end); a uniform text under a saturating table gets no dictionary"""
    rng = np.random.default_rng(77)
    alpha = rng.choice(np.arange(1, 255), 27, replace=False).astype(np.uint8)
    text = _zipf_text(rng, alpha, 6000, 300)
    bwt, eof, cnt = fo.build_bwt(fo.file_to_text_rev(text))
    o = fo.OracleIndex.from_bwt(bwt, eof, cnt)
    tp = bytes(fo.file_to_text_rev(text))
    full = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, accel=fx.ACCEL_KMER | fx.ACCEL_DICT, kmer_table_bytes=8 * 28 ** 2, dict_min_rows=2)
    fi = full.info()
    full.close()
    g = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, accel=fx.ACCEL_KMER | fx.ACCEL_DICT, kmer_table_bytes=8 * 28 ** 2, dict_min_rows=2,
                         dict_bytes=fi["dict_bytes"] // 3)
    gi = g.info()
    assert 2 < gi["dict_depth"] < fi["dict_depth"] and gi["dict_bytes"] <= fi["dict_bytes"] // 3 + 128
    offs = rng.integers(0, len(tp) - 14, 500)
    arr = np.stack([np.frombuffer(tp[s:s + 14], np.uint8) for s in offs])
    sp, ep = g.count_fixed(arr)
    osp, oep = o.count_batch(arr.reshape(-1), np.arange(0, arr.size + 1, 14, dtype=np.int64))
    assert np.array_equal(sp, osp) and np.array_equal(ep, oep)
    g.close()
    o.close()
    utext = alpha[rng.integers(0, 27, 20000)].tobytes()
    bwt, eof, cnt = fo.build_bwt(fo.file_to_text_rev(utext))
    g = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt)          # AUTO: table depth ~ log_27(8n): the few 4-mers that are wide by chance
    assert g.info()["kmer_k"] >= 3 and g.info()["dict_depth"] == 0        # cover next to nothing => no dictionary (no probe per query for nothing)
    g.close()


@pytest.mark.parametrize("sigma", [3, 10, 16, 17, 255])
@pytest.mark.parametrize("lanes", [1, 2, 4])
def test_multiary_wavelet_matrix(sigma, lanes):
    """FMX_LAYOUT_WMX: 4-ary single level (<= 4 symbols), 16-ary single level (<= 16) and two levels (more) — occ on every symbol and a
    spread of rows, LF, count and locate, all against the oracle (the shared parity suites above run the layout through everything else)"""
    rng = np.random.default_rng(500 + sigma)
    alpha = np.arange(1, 256, dtype=np.uint8) if sigma == 255 else rng.choice(np.arange(1, 256), sigma, replace=False).astype(np.uint8)
    text = alpha[rng.integers(0, sigma, 5000)].tobytes()
    tp = bytes(fo.file_to_text_rev(text))
    bwt, eof, cnt = fo.build_bwt(fo.file_to_text_rev(text))
    o = fo.OracleIndex.from_bwt(bwt, eof, cnt)
    g = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, layout=fx.LAYOUT_WMX, lanes_per_query=lanes, accel=fx.ACCEL_NONE, sa_sample_rate=8)
    info = g.info()
    assert info["layout"] == "wmx" and info["levels"] == (1 if sigma <= 16 else 2)
    keys = np.concatenate([np.arange(-1, 140), rng.integers(0, o.n, 300), [o.n - 2, o.n - 1, eof - 1, eof, eof + 1]]).astype(np.int64)
    keys = keys[(keys >= -1) & (keys < o.n)]
    for c in list(alpha[:20]) + [0, int(alpha[-1]), 255 if 255 not in alpha else 254]:
        got = g.occ_batch(np.full(len(keys), c, np.uint8), keys)
        assert got.tolist() == [o.occ(int(c), int(k)) for k in keys], c
    rows = rng.integers(0, o.n, 500)
    assert g.get_prev_i_batch(rows).tolist() == [o.getPrevI(int(r)) for r in rows]
    pats = [tp[s:s + int(rng.integers(1, 9))] for s in rng.integers(0, len(tp) - 9, 600)] + [b"", bytes([int(alpha[0])]) * 3, b"\0"]
    sp, ep = g.count_batch(pats)
    for i, p in enumerate(pats):
        assert (int(sp[i]), int(ep[i])) == (o.search(p) or (0, 0)), p
    off, pos = g.locate_batch(sp[:100], ep[:100])
    sa = o.sa().astype(np.int64)
    for k in range(100):
        assert np.array_equal(pos[off[k]:off[k + 1]], np.sort(sa[sp[k]:ep[k]]))
    g.close()
    o.close()


def test_long_fixed_patterns_leave_shared_memory(ref_dir):
    """fixed-length patterns too long to stage a CTA's worth in shared memory (> ~640 bytes per lane) take the global-memory kernel"""
    text = open(os.path.join(ref_dir, "test.txt"), "rb").read()
    o = fo.OracleIndex.load(os.path.join(ref_dir, "test.cmp"), big_endian=False)
    tp = text[::-1]
    for lanes in (1, 4):
        g = fx.GpuFMSearcher(os.path.join(ref_dir, "test.cmp.bwt"), bigEndian=False, lanes_per_query=lanes)
        for ln in (700, 2600):
            arr = np.stack([np.frombuffer(tp[s:s + ln], np.uint8) for s in range(0, 3000, 100)]).copy()
            arr[::3, ln // 2] = ord("!")
            sp, ep = g.count_fixed(arr)
            osp, oep = o.count_batch(arr.reshape(-1), np.arange(0, arr.size + 1, ln, dtype=np.int64))
            assert np.array_equal(sp, osp) and np.array_equal(ep, oep) and (ep > sp).sum() >= 10
            assert np.array_equal(g.count_only_fixed(arr).astype(np.int64), oep - osp)
        g.close()
    o.close()


def test_mismatched_bwt_and_aux_are_refused(ref_dir):
    """.aux counts that do not describe the .bwt (LF would not be a permutation): FMX_E_FORMAT at open instead of a runaway walk"""
    o = fo.OracleIndex.load(os.path.join(ref_dir, "test1024.cmp"), big_endian=False)
    bwt = np.array(o.bwt(), np.uint8)
    counts = np.bincount(bwt, minlength=256).astype(np.int64)
    counts[0] = 0
    a, b = int(bwt[10]), int(bwt[10]) % 255 + 1
    bad = bwt.copy()
    bad[bad == a] = b                                       # same length, different histogram
    with pytest.raises(fx.FmxError, match="do not belong together") as ei:
        fx.GpuFMSearcher(bwt=bad, eof=o.eof, counts=counts, sa_sample_rate=4)
    assert ei.value.code == fx.FMX_E_FORMAT
    # right histogram, wrong order: LF is a permutation but not one cycle -> refused by the chain walk's own check, no hang
    perm = bwt.copy()
    i, j = [k for k in range(len(perm)) if k != o.eof][:2]
    k2 = next(k for k in range(len(perm) - 1, 0, -1) if k != o.eof and perm[k] != perm[i])
    perm[i], perm[k2] = perm[k2], perm[i]
    try:
        g = fx.GpuFMSearcher(bwt=perm, eof=o.eof, counts=counts, sa_sample_rate=4)
        g.close()                                           # a swap may by chance keep one cycle; then the index is simply another text's
    except fx.FmxError as e:
        assert e.code == fx.FMX_E_FORMAT
    o.close()


def test_total_footprint_cap(ref_dir):
    """fmx_opts.max_total_bytes: everything resident stays under the cap, whatever that leaves out, and results do not change"""
    text = open(os.path.join(ref_dir, "test.txt"), "rb").read()
    o = fo.OracleIndex.load(os.path.join(ref_dir, "test.cmp"), big_endian=False)
    rng = np.random.default_rng(3)
    pats = _patterns(text, rng, 600, 16) + _patterns(text, rng, 300, 30)
    want = [o.search(p) or (0, 0) for p in pats]
    n = o.n
    seen = set()
    for cap in (0, 60 * n, 40 * n, 12 * n, 3 * n, 2 * n):
        g = fx.GpuFMSearcher(os.path.join(ref_dir, "test.cmp.bwt"), bigEndian=False, max_total_bytes=cap)
        info = g.info()
        assert cap == 0 or info["index_bytes"] <= cap, (cap, info)
        seen.add((info["layout"], info["ctx_depth"] > 0, info["kmer_k"]))
        sp, ep = g.count_batch(pats)
        assert [(int(a), int(b)) for a, b in zip(sp, ep)] == want
        g.close()
    assert len(seen) >= 3                                   # the cap really changed what was built
    with pytest.raises(fx.FmxError):
        fx.GpuFMSearcher(os.path.join(ref_dir, "test.cmp.bwt"), bigEndian=False, max_total_bytes=n // 2)
    o.close()


def test_compact_contexts_refuse_larger_alphabets(ref_dir):
    with pytest.raises(fx.FmxError, match="2-bit symbols"):
        fx.GpuFMSearcher(os.path.join(ref_dir, "test1024.cmp.bwt"), bigEndian=False, accel=fx.ACCEL_CTX8)


def test_compact_contexts_are_chosen_when_the_wide_form_does_not_fit(monkeypatch):
    """FMX_ACCEL_AUTO on a 4-symbol text whose 32-byte contexts would not fit the (here artificially small) TLB reach: the compact form
    is built, isat is not, the k-mer table saturates, and count/locate still equal the oracle."""
    rng = np.random.default_rng(5)
    text = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 200_000)].tobytes()
    bwt, eof, cnt = fo.build_bwt(fo.file_to_text_rev(text))
    o = fo.OracleIndex.from_bwt(bwt, eof, cnt)
    monkeypatch.setenv("FMX_TLB_REACH_GB", "%.6f" % ((4e9 + 20 * len(bwt)) / 1e9))      # 8n fits, 32n does not
    g = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt)
    info = g.info()
    assert info["ctx_entry_bytes"] == 8 and info["ctx_depth"] == 16 and not info["text_shortcut"] and 4 ** info["kmer_k"] >= len(bwt) // 2
    tp = bytes(fo.file_to_text_rev(text))
    for ln in (12, 16 + info["kmer_k"], 32, 40, 32 + info["kmer_k"], 75):
        offs = rng.integers(0, len(tp) - ln, 3000)
        arr = np.stack([np.frombuffer(tp[s:s + ln], np.uint8) for s in offs]).copy()
        arr[::7, ln // 2] = ord("A")                        # some near misses
        sp, ep = g.count_fixed(arr)
        osp, oep = o.count_batch(arr.reshape(-1), np.arange(0, arr.size + 1, ln, dtype=np.int64))
        assert np.array_equal(sp, osp) and np.array_equal(ep, oep)
    off, pos = g.locate_batch(sp[:50], ep[:50])
    sa = o.sa().astype(np.int64)
    for k in range(50):
        assert np.array_equal(pos[off[k]:off[k + 1]], np.sort(sa[sp[k]:ep[k]]))
    g.close()
    o.close()


@pytest.mark.parametrize("accel", [fx.ACCEL_AUTO, fx.ACCEL_NONE], ids=["resident_sa", "built_for_the_call"])
def test_write_sa_file_is_sacreator_output(ref_dir, tmp_path, accel):
    """SACreator.create (bwtmerger.scala:535-556): <base>.sa = n x int32 big-endian, no header, == bwtFm2sa (G7, T/Indexer.scala:1043-1068)"""
    o = fo.OracleIndex.load(os.path.join(ref_dir, "test1024.cmp"), big_endian=False)
    g = fx.GpuFMSearcher(os.path.join(ref_dir, "test1024.cmp.bwt"), bigEndian=False, accel=accel)
    g.write_sa_file(str(tmp_path / "t.whatever"))
    raw = (tmp_path / "t.sa").read_bytes()
    assert len(raw) == 4 * o.n
    assert np.array_equal(np.frombuffer(raw, ">i4").astype(np.int64), o.sa().astype(np.int64))
    g.close()
    o.close()


# ----------------------------------------------------------------------------------------------- edge cases
@pytest.mark.parametrize("accel", [fx.ACCEL_AUTO, fx.ACCEL_NONE], ids=["auto", "none"])
@pytest.mark.parametrize("cfg", [(fx.LAYOUT_WM, 2), (fx.LAYOUT_PLANES, 4)], ids=_ids)
@pytest.mark.parametrize("text", [b"", b"a", b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa", b"ab" * 300, bytes(range(1, 256)) * 3,
                                  b"\xff\xfe\xff\xfe\x01\x01\x02", b"abracadabra"], ids=["empty", "one", "run", "period2", "all255", "highbytes", "abra"])
def test_degenerate_texts(text, cfg, accel):
    """Empty text (n = 1), single symbol, periodic text, every byte value, bytes >= 0x80 (Q8: reference-undefined through search's
    signed Byte, defined through getPrevRange; unsigned here): every operator vs the oracle, incl. locate and both regex engines."""
    bwt, eof, cnt = fo.build_bwt(fo.file_to_text_rev(text))
    o = fo.OracleIndex.from_bwt(bwt, eof, cnt)
    g = fx.GpuFMSearcher(bwt=bwt, eof=eof, counts=cnt, layout=cfg[0], lanes_per_query=cfg[1], accel=accel, sa_sample_rate=4)
    n = o.n
    assert g.n == n and g.eof == eof
    alphabet = sorted(set(text)) + [0, 1, 7, 255]
    rng = np.random.default_rng(len(text))
    pats = [b"", b"\0", bytes([alphabet[0]]) * 40] + [bytes(rng.choice(alphabet, int(rng.integers(0, 9))).astype(np.uint8)) for _ in range(300)]
    tp = bytes(fo.file_to_text_rev(text))
    pats += [tp[i:i + k] for i in range(0, max(len(tp) - 1, 1), 7) for k in (1, 3, 14) if tp[i:i + k]]
    sp, ep = g.count_batch(pats)
    for p, a, b in zip(pats, sp, ep):
        r = o.search(p)
        assert (int(a), int(b)) == (r if r else (0, 0)), p
    rows = np.arange(n, dtype=np.int64)
    assert g.get_prev_i_batch(rows).tolist() == [o.getPrevI(int(r)) for r in rows]
    assert g.get_next_i_batch(rows).tolist() == [o.getNextI(int(r)) for r in rows]
    for c in alphabet:
        keys = np.arange(-1, n, dtype=np.int64)
        assert g.occ_batch(np.full(len(keys), c, np.uint8), keys).tolist() == [o.occ(c, int(k)) for k in keys]
    off, pos = g.locate_batch([0], [n])
    assert pos.tolist() == sorted(o.sa().tolist())
    if len(text) >= 2:
        a, b = text[0], text[1]
        rx = bytes([a]) + b"." if a not in b"()[]|*+?.\\-" else b"x."
        assert fx.ReTree(rx).matchSA(g) == o.regex_match(rx)
        assert fx.ThompsonNFA(rx).matchSA(g) == o.regex_match_thompson(rx)
    g.close()
    o.close()


def test_empty_batches_and_argument_errors(ref_dir):
    g = _open(os.path.join(ref_dir, "test1024.cmp.bwt"), (fx.LAYOUT_PLANES, 2), sa_sample_rate=8)
    sp, ep = g.count_batch([])
    assert len(sp) == 0 and len(ep) == 0
    sp, ep = g.count_fixed(np.zeros((0, 16), np.uint8))
    assert len(sp) == 0
    sp, ep = g.count_fixed(np.zeros((5, 0), np.uint8))                  # five empty patterns -> (0, n) each
    assert sp.tolist() == [0] * 5 and ep.tolist() == [g.n] * 5
    off, pos = g.locate_batch([], [])
    assert off.tolist() == [0] and len(pos) == 0
    off, pos = g.locate_batch([5, 9], [5, 9])                           # empty intervals
    assert off.tolist() == [0, 0, 0]
    assert g.regex_search_batch([]) == []
    assert g.getIntervalPrevRange(0, g.n, ord("z"), ord("a")) == []     # cstart > cend: the reference's loop does not run
    for bad in ([-1], [g.n]):
        with pytest.raises(fx.FmxError) as e:
            g.get_prev_i_batch(bad)
        assert e.value.code == fx.FMX_E_ARG
    with pytest.raises(fx.FmxError) as e:
        g.prev_range_batch([0], [g.n + 1], [97])
    assert e.value.code == fx.FMX_E_ARG
    with pytest.raises(fx.FmxError) as e:
        g.locate_batch([10], [5])
    assert e.value.code == fx.FMX_E_ARG
    g2 = _open(os.path.join(ref_dir, "test1024.cmp.bwt"), (fx.LAYOUT_WM, 2), accel=fx.ACCEL_NONE)
    with pytest.raises(fx.FmxError) as e:                               # no samples, no full SA
        g2.locate_batch([0], [3])
    assert e.value.code == fx.FMX_E_ARG
    g2.close()
    # capacity protocol of locate: too small -> FMX_E_CAPACITY and the required total
    import ctypes as C
    sp = np.array([0], np.int64)
    ep = np.array([100], np.int64)
    off = np.zeros(2, np.int64)
    pos = np.zeros(10, np.int64)
    rc = fx.lib().fmx_locate_batch(g.h, sp.ctypes.data_as(C.c_void_p), ep.ctypes.data_as(C.c_void_p), 1, 10, off.ctypes.data_as(C.c_void_p),
                                   pos.ctypes.data_as(C.c_void_p))
    assert rc == fx.FMX_E_CAPACITY and off[1] == 100
    g.close()


def test_regex_set_is_reusable(ref_dir, o1024, words_base, words):
    """compile once, search many times: one device-resident regex set searched repeatedly and against two different indexes.  The set's
    follow lists are pruned by the alphabet of the index it was created for, so it serves any index whose symbols all occur there
    (test1024's a-z inside words' alphabet) and refuses one with other symbols."""
    g = _open(os.path.join(ref_dir, "test1024.cmp.bwt"), (fx.LAYOUT_PLANES, 2))
    w = _open(words_base + ".bwt", (fx.LAYOUT_WM, 2), big_endian=True)
    rxs = ["a(a|b|d|e)c", "ab", "q[a-z]q", "a.b", "x\\wy", "zz?z?", "a+b"]
    trees = [fx.ReTree(r) for r in rxs[:5]] + [fx.ThompsonNFA(r) for r in rxs[5:]]
    rs = w.regex_set(trees)
    for _ in range(3):
        off, ln, sp, ep = rs.search(g, cap_total=4)                          # also exercises the capacity retry
        for i, rx in enumerate(rxs):
            got = list(zip(ln[off[i]:off[i + 1]].tolist(), sp[off[i]:off[i + 1]].tolist(), ep[off[i]:off[i + 1]].tolist()))
            want = o1024.regex_match(rx) if i < 5 else o1024.regex_match_thompson(rx)
            assert got == want, rx
    off, ln, sp, ep = rs.search(w)
    o, _ = words
    for i, rx in enumerate(rxs):
        got = list(zip(ln[off[i]:off[i + 1]].tolist(), sp[off[i]:off[i + 1]].tolist(), ep[off[i]:off[i + 1]].tolist()))
        assert got == (o.regex_match(rx) if i < 5 else o.regex_match_thompson(rx)), rx
    rs.close()
    narrow = g.regex_set(trees)                                              # pruned for a-z: words' '\r', '\n' are not covered
    with pytest.raises(fx.FmxError, match="create the set against this index") as ei:
        narrow.search(w)
    assert ei.value.code == fx.FMX_E_ARG
    narrow.close()
    w.close()
    g.close()


def test_concurrent_callers_and_two_indexes(ref_dir, words_base, words):
    """An opened index is immutable; batch calls from several host threads (ctypes drops the GIL) and on several indexes at once
    must not interfere (thread-safe batch calls)."""
    import threading
    o1 = fo.OracleIndex.load(os.path.join(ref_dir, "test.cmp"), big_endian=False)
    o2, text2 = words
    g1 = _open(os.path.join(ref_dir, "test.cmp.bwt"), (fx.LAYOUT_PLANES, 2), sa_sample_rate=8)
    g2 = _open(words_base + ".bwt", (fx.LAYOUT_WM, 2), big_endian=True)
    text1 = open(os.path.join(ref_dir, "test.txt"), "rb").read()
    rng = np.random.default_rng(21)

    def pats_of(text, m, ln):
        t = np.frombuffer(text, np.uint8)
        offs = rng.integers(0, len(t) - ln, m)
        return np.stack([t[s:s + ln][::-1] for s in offs])
    p1, p2 = pats_of(text1, 20000, 5), pats_of(text2, 20000, 7)
    w1 = o1.count_batch(p1.reshape(-1), np.arange(0, p1.size + 1, 5, dtype=np.int64))
    w2 = o2.count_batch(p2.reshape(-1), np.arange(0, p2.size + 1, 7, dtype=np.int64))
    rx_want = o1.regex_match("a.b")
    errors = []

    def worker(k):
        try:
            for _ in range(10):
                if k % 3 == 0:
                    sp, ep = g1.count_fixed(p1)
                    assert np.array_equal(sp, w1[0]) and np.array_equal(ep, w1[1])
                elif k % 3 == 1:
                    sp, ep = g2.count_fixed(p2)
                    assert np.array_equal(sp, w2[0]) and np.array_equal(ep, w2[1])
                else:
                    assert fx.ReTree("a.b").matchSA(g1) == rx_want
                    off, pos = g1.locate_batch(w1[0][:50], w1[1][:50])
                    assert off[-1] == int((w1[1][:50] - w1[0][:50]).sum())
        except Exception as e:                      # noqa: BLE001
            errors.append(repr(e))
    th = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    g1.close()
    g2.close()

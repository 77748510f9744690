"""
Pins the CPU oracle (oracle/) against the reference's own known-answer tests — vectors G1..G12 of
SURVEY.md §8(c).  T/ = /root/reference/src/test/scala/org/fmindex/tests/.  CPU only.
"""
import os
import re as pyre

import numpy as np
import pytest

from oracle import fm_oracle as fo
from oracle import retree


def _idx(text_rev):
    return fo.OracleIndex.from_text_rev(text_rev)


# ---------------------------------------------------------------- G1  T/Indexer.scala:203-333
def test_g1_abracadabra():
    ix = _idx(b"abracadabra")
    assert bytes(ix.bwt()).replace(b"\0", b"$") == b"ard$rcaaaabb"                 # :203-212
    assert ix.fm().tolist() == [3, 0, 6, 7, 8, 9, 10, 11, 5, 2, 1, 4]              # :239, :272
    assert ix.cf(0) == 0 and ix.cf(ord("a")) == 1 and ix.cf(ord("b")) == 6         # :249-251
    rows = {                                                                       # :286-292
        0: [0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1],
        ord("a"): [1, 1, 1, 1, 1, 1, 2, 3, 4, 5, 5, 5],
        ord("b"): [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2],
        ord("c"): [0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1],
        ord("d"): [0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1],
        ord("r"): [0, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2],
        ord("x"): [0] * 12,
    }
    for c, want in rows.items():
        assert [ix.occ(c, i) for i in range(ix.n)] == want
    assert ix.search(b"bra") == (6, 8)                                             # :296-306
    assert ix.getPrevI(6) == 2 and ix.getNextI(6) == 10 and ix.getNextI(10) == 1   # :308-323
    # :324-333 are asserted on SAISBuilder, whose nextSubstr/prevSubstr (M/sais.scala:110-148) reverse the
    # opposite way round from NaiveFMSearcher's (M/bwtmerger.scala:394-419): same walk, mirrored string.
    assert ix.nextSubstr(6, 4)[::-1] == b"bra\0"
    assert ix.prevSubstr(6, 4)[::-1] == b"cada"


# ---------------------------------------------------------------- G2  T/Indexer.scala:334-351
def test_g2_prev_range():
    ix = _idx(b"mmabcacadabbbca"[::-1])
    assert ix.occ(ord("b"), 6) == 3
    assert ix.getPrevRange(0, 16, ord("a")) == (1, 6)
    assert ix.getPrevRange(1, 6, ord("b")) == (6, 8)
    assert ix.nextSubstr(11, 3)[::-1] == b"cba"
    assert ix.prevSubstr(11, 3)[::-1] == b"aca"


# ---------------------------------------------------------------- G3  T/REParser.scala:236-291
def test_g3_small2(ref_dir):
    data = open(os.path.join(ref_dir, "small2.txt"), "rb").read()
    ix = _idx(fo.file_to_text_rev(data))
    assert bytes(ix.pos2char(k) for k in range(11)) == b"iiiimppssss"
    chain = [0]
    for _ in range(5):
        chain.append(ix.getNextI(chain[-1]))
    assert chain == [0, 5, 4, 10, 9, 3]
    chain = [3]
    for _ in range(6):
        chain.append(ix.getPrevI(chain[-1]))
    assert chain == [3, 9, 10, 4, 5, 0, 1]
    assert chr(ix.bwt()[4]) == "m"


# ---------------------------------------------------------------- G4  T/Indexer.scala:1079-1124
@pytest.fixture(scope="module")
def ix1024(ref_dir):
    # The C-tool goldens are little-endian (BWTLoader(..., false), T/Indexer.scala:643); they hold the same BWT
    # that BWTMerger2.merge produces (asserted by MergerTest), so the on-disk index loads directly from them.
    return fo.OracleIndex.load(os.path.join(ref_dir, "test1024.cmp"), big_endian=False)


def test_g4_test1024(ix1024):
    ix = ix1024
    eof = ix.eof
    assert ix.n == 1025 and eof == 462
    assert bytes(ix.bwt()[:3]) == b"ubx"
    assert ix.bwt()[eof] == 0
    assert ix.getPrevI(eof) == 0
    assert chr(ix.bwt()[ix.getPrevI(eof)]) == "u"          # first char of the file
    assert ix.getNextI(eof) == 517
    assert chr(ix.bwt()[ix.getNextI(eof)]) == "l"          # last char of the file
    assert ix.getPrevI(1) == 48
    assert ix.getPrevI(48) == 649
    assert ix.nextSubstr(1, 3) == b"haa"
    assert chr(ix.bwt()[1000]) == "b"
    assert ix.nextSubstr(ix.getNextI(eof), 100) == \
        b"zajrtzbeqwbxdfpwjflmmsseewuudgfbtzqenjqafwzcnfanycigwsflfvxojxpqhhzekjdkhgsptqveavquuoqujbezdkarayom"
    assert ix.nextSubstr(eof, 100) == \
        b"ajrtzbeqwbxdfpwjflmmsseewuudgfbtzqenjqafwzcnfanycigwsflfvxojxpqhhzekjdkhgsptqveavquuoqujbezdkarayoml"
    assert ix.prevSubstr(1, 5) == b"bqxxa"
    assert ix.prevSubstr(eof, 5) == b"\0uexm"
    assert ix.prevSubstr(ix.getPrevI(eof), 4) == b"uexm"


# ---------------------------------------------------------------- G5  T/REParser.scala:292-307
def test_g5_alternation_counts(ix1024, ref_dir):
    text = open(os.path.join(ref_dir, "test1024.txt"), "rb").read()
    got = {}
    for w in (b"ac", b"bc", b"dc", b"ec"):
        r = ix1024.search(w[::-1])               # search(p) finds reverse(p) in the file
        got[w] = (r[1] - r[0]) if r else 0
        assert got[w] == len(pyre.findall(b"(?=" + w + b")", text))
    assert got == {b"ac": 2, b"bc": 1, b"dc": 1, b"ec": 1}      # Set("ec","dc","[2 Results] ac","bc")
    assert ix1024.search(b"ca") == (83, 85)                     # intervals derived in SURVEY §8c G5
    assert ix1024.search(b"cb") == (85, 86)
    assert ix1024.search(b"cd") == (86, 87)
    assert ix1024.search(b"ce") == (87, 88)
    # the Glushkov builder rejects this shape (Q3): concat (Or, Char) has no case ...
    with pytest.raises(retree.ReUnsupported):
        retree.compile_regex("(a|b|d|e)c")
    # ... but accepts it behind a literal; the result multiset then equals the literal searches
    for lead in b"abcdefghijklmnopqrstuvwxyz":
        res = ix1024.regex_match(bytes([lead]) + b"(a|b|d|e)c")
        want = []
        for mid in b"abde":
            r = ix1024.search(bytes([lead, mid, ord("c")])[::-1])
            if r:
                want.append((3, r[0], r[1]))
        assert res == sorted(want)


# ---------------------------------------------------------------- G6/G7/G8  T/Indexer.scala:638-900, 1043-1068
CMP = [("test1024", 1025, 462), ("test2048", 2049, 1118), ("test2048-2", 2049, 1), ("test3072", 3073, 6),
       ("test", 10241, 2658), ("test-part", 2549, 774)]


@pytest.mark.parametrize("name,n,eof", CMP)
def test_g8_formats_and_g6_fm_and_g7_sa(ref_dir, name, n, eof):
    ix = fo.OracleIndex.load(os.path.join(ref_dir, name + ".cmp"), big_endian=False)
    assert (ix.n, ix.eof) == (n, eof)
    bwt = ix.bwt().copy()
    assert bwt[eof] == 0
    # G6: fm == bwt2occ(bwt with eof->0)  (M/util.scala:121-134) == stable argsort
    assert np.array_equal(ix.fm(), np.argsort(bwt, kind="stable").astype(np.uint32))
    # any correct suffix sorter reproduces the golden BWT / aux from the text (reversed, 0x00 dropped)
    data = open(os.path.join(ref_dir, name + ".txt"), "rb").read()
    b2, eof2, cnt2 = fo.build_bwt(fo.file_to_text_rev(data))
    assert eof2 == eof and np.array_equal(b2, bwt)
    aux = np.frombuffer(open(os.path.join(ref_dir, name + ".cmp.aux"), "rb").read(), dtype="<i8")
    assert np.array_equal(aux[1:], cnt2[1:])
    # G7: sa == bwtFm2sa: row r holds the suffix of T' starting at sa[r]
    sa = ix.sa()
    tprime = bytes(fo.file_to_text_rev(data)) + b"\0"
    assert sa[eof] == 0 and sa[0] == n - 1
    order = sorted(range(n), key=lambda i: tprime[i:])
    assert sa.tolist() == order


def test_g8_bad_sizes(tmp_path, ref_dir):
    raw = open(os.path.join(ref_dir, "test1024.cmp.bwt"), "rb").read()
    (tmp_path / "x.bwt").write_bytes(raw[:-1])
    (tmp_path / "x.aux").write_bytes(open(os.path.join(ref_dir, "test1024.cmp.aux"), "rb").read())
    with pytest.raises(RuntimeError):
        fo.OracleIndex.load(str(tmp_path / "x"), big_endian=False)
    with pytest.raises(RuntimeError):                            # wrong endianness => size check fails
        fo.OracleIndex.load(os.path.join(ref_dir, "test1024.cmp"), big_endian=True)


def test_fm_file_roundtrip(tmp_path, ref_dir):
    ix = fo.OracleIndex.load(os.path.join(ref_dir, "test2048.cmp"), big_endian=False)
    cnt = np.frombuffer(open(os.path.join(ref_dir, "test2048.cmp.aux"), "rb").read(), dtype="<i8")
    fo.write_index_files(str(tmp_path / "t"), ix.bwt(), ix.eof, cnt, big_endian=True, write_fm=True)
    assert os.path.getsize(tmp_path / "t.fm") == 4 * ix.n + 9
    ix2 = fo.OracleIndex.load(str(tmp_path / "t.fm"))
    assert np.array_equal(ix2.fm(), ix.fm()) and ix2.eof == ix.eof


# ---------------------------------------------------------------- G12  T/REParser.scala:10-26
def test_g12_re2post():
    assert retree.re2poststr("abc") == "ab·c·"
    assert retree.re2poststr("a(bb)+a") == "abb·+·a·"
    assert retree.re2poststr("(a|b)") == "ab|"
    assert retree.re2poststr("((a|b)*aba*)*(a|b)(a|b)") == "ab|*a·b·a*·*ab|·ab|·"


@pytest.mark.parametrize("bad", ["*a", "(", "a(b", "|a", "()", "a||b", "[abc", "[a-", "[-a]", "[b-a]"])
def test_re2post_syntax_errors(bad):
    with pytest.raises(retree.ReSyntaxError):
        retree.re2post(bad)


# ---------------------------------------------------------------- G9  T/REParser.scala:319-588
SMOKE = ["abcd", "abcd*", "abc*d", "a*bcd", "a*b*c*d*", "(ab)*", "(ab)*cd", "(ab)*(cd*)*", "(a|b)", "(a|b|d|c)",
         "(a|b*|d|c)", "(a|b*|d|c)*|(abc)", "(a|b|c)|(c|d|e)", "[a-c]", "a[b-d]e", "a[b-d]*e", "a[x.]e", "a\\de",
         "a+", "a****", "a+b", "a+((b|c)+|d)", "a*+", "a+*", "a+*+*++*", "a?", "(abc)?+|a?|bcd", "ab(cd|ef)+gh",
         "(10\\.[0-9]|[1-9][0-9]|[1-2][0-5][0-5]\\.[0-9]|[1-9][0-9]|[1-2][0-5][0-5]\\.[0-9]|[1-9][0-9]|[1-2][0-5][0-5])",
         "ab(cd)*ef", "ab*(cd)*(gh)*ij", "a(cd|ef)*j",
         ".*ab(cd)*(m(k|l)|tm*)(a|abc)(a*|(abc)*)ef(a*b*c*dg*)*gh", "a.*(b|c)d.*f"]


@pytest.mark.parametrize("rx", SMOKE)
def test_g9_reference_smoke_regexes_build(rx):
    retree.compile_regex(rx).tables()            # the reference's REAnalys tests only require "does not throw"


def test_g9_structure():
    t = retree.compile_regex("a")
    assert t.root.parent is None and str(t.root.childs[0].parent) == "F[a]"             # anal1
    t = retree.compile_regex("ab*", remove_nulls=False)
    assert str(t.root.childs[1].childs[0].parent) == "*[b]"                            # anal2
    assert len(retree.compile_regex("a*(b|a)*bB*cd*e*").root.childs) == 3              # anal2.1
    t = retree.compile_regex("a*(b|a)*b?B*c?d*e*")
    assert len(t.root.childs) == 0 and retree.is_null(t.root)                          # anal2.2
    assert retree.compile_regex("abcdef").root.childs[3].num == 4                      # anal3
    assert retree.compile_regex("(a|bX|cYZ)(a|b|c)").root.childs[1].childs[1].num == 4  # anal6
    assert retree.compile_regex("(a|b|c)(a|b|c)").root.childs[1].childs[1].num == 2     # anal7


def test_g9_follows():
    F = retree.compile_regex("abc(cde)*ef").root                                       # anal4.follows
    a, b, c, cdeS, e, f = F.childs
    fol = retree.follows
    assert fol(F) == []
    assert fol(a) == [b] and fol(b) == [c] and fol(cdeS) == [e] and fol(e) == [f] and fol(f) == []
    cdeSF = cdeS.childs[0]
    c2, d2, e2 = cdeSF.childs
    assert fol(cdeSF) == [c2, e] and fol(c2) == [d2] and fol(d2) == [e2] and fol(e2) == [c2, e]
    F = retree.compile_regex("ab?j").root                                              # anal4.follows.or.?
    assert retree.is_null(F.childs[1])
    assert fol(F.childs[0]) == [F.childs[2], F.childs[1].childs[0]]


def test_quirks():
    # Q1: Interval upper bound exclusive
    t = retree.compile_regex("a\\db").tables()
    assert sorted(t["c"][1:-1]) == list(range(ord("0"), ord("9")))
    t = retree.compile_regex("a.b").tables()
    assert sorted(t["c"][1:-1]) == list(range(2, 255))
    t = retree.compile_regex("a.b", line_only=True).tables()
    assert sorted(t["c"][1:-1]) == list(range(0x20, 255))
    t = retree.compile_regex("x[a-c]d").tables()                        # classes are inclusive
    assert sorted(t["c"][1:-1]) == [ord("a"), ord("b"), ord("c")]
    # Q2: follows inside a nested concatenation does not climb when the remaining siblings are nullable
    t = retree.compile_regex("x(ab?|d)c")
    tb = t.tables()
    ia = tb["c"].index(ord("a"))
    assert [tb["c"][j] for j in tb["follows"][ia]] == [ord("b")]
    # Q3: unsupported shapes
    for rx in ["a(bc)d", "(a|b)c", "[a-c]d", "a|b*"]:
        with pytest.raises(retree.ReUnsupported):
            retree.compile_regex(rx)
    for rx in ["x(a|b)c", "ab(cd)*ef", "x[a-c]d"]:
        retree.compile_regex(rx)
    # Q4: a+((b|c)+|d) reduces to F[a]
    assert str(retree.compile_regex("a+((b|c)+|d)").root) == "F[a]"
    # Q5: border trimming
    assert retree.compile_regex(".*foo.*").tables() == retree.compile_regex("foo").tables()
    with pytest.raises(retree.ReUnsupported):
        retree.compile_regex("")                                        # args.pop on an empty stack


# ---------------------------------------------------------------- G10  T/REParser.scala:594-605
def test_g10_glushkov_over_toy_sa():
    ix = _idx(b"mmabcacamabbbca"[::-1])
    res = ix.regex_match(".*(a|b)ca")
    assert len(res) == 2
    assert res == [(3, 1, 2), (3, 2, 4)]                                # triples derived in SURVEY §8c


# ---------------------------------------------------------------- words.* (survey cross-checks + brute force)
@pytest.fixture(scope="module")
def words(words_base):
    ix = fo.OracleIndex.load(words_base)
    # invert the BWT: T'[sa[r]-1] = bwt[r]; file text = reverse(T' without '$')
    sa = ix.sa()
    tp = np.zeros(ix.n, np.uint8)
    bwt = ix.bwt()
    tp[(sa.astype(np.int64) - 1) % ix.n] = bwt
    text = bytes(tp[:-1][::-1])
    return ix, text


def test_words_index_header(words):
    ix, text = words
    assert ix.n == 1916149 and ix.eof == 86533
    assert len(text) == 1916148 and text.count(b"\r\n") == 172820 and len(set(text)) == 28


def test_words_survey_crosschecks(words):
    ix, text = words
    assert ix.search(b"hello"[::-1]) == (1333929, 1333938)
    assert ix.search(b"ing\r\n"[::-1]) == (40318, 52882)
    assert ix.search(b"\nqu"[::-1]) == (1838051, 1838856)
    assert ix.search(b"zzz") is None
    for w in (b"hello", b"ing\r\n", b"\nqu", b"abc", b"q"):
        r = ix.search(w[::-1])
        assert (r[1] - r[0] if r else 0) == len(pyre.findall(b"(?=" + pyre.escape(w) + b")", text))


@pytest.mark.parametrize("rx,total", [("x(a|b|d|e)c", 90), ("ab?c[d-h]", 1668), ("q(u|a)[a-m]z?k", 22),
                                     ("th(e|a)(n|t)\\w", 363), ("b(oo|ee)+k", 158), ("z[aeiou][aeiou]?z", 21)])
def test_words_glushkov_totals(words, rx, total):
    ix, text = words
    res = ix.regex_match(rx)
    assert sum(ep - sp for _, sp, ep in res) == total
    # occurrences = all start positions where python's re matches (shortest-match semantics => count starts
    # per distinct match length is what the triples hold; totals agree for these patterns)
    pyrx = rx.replace("\\w", "[A-y]")
    if rx != "b(oo|ee)+k":
        assert total == len(pyre.findall(("(?=" + pyrx + ")").encode(), text))


def test_locate_definition(words):
    ix, text = words
    n = ix.n
    r = ix.search(b"hello"[::-1])
    pos = ix.locate(*r)
    assert list(pos) == sorted(pos)
    for q in pos:                                     # file offset = (n-1) - q - m   (SURVEY §8a conventions)
        off = (n - 1) - int(q) - 5
        assert text[off:off + 5] == b"hello"


# ---------------------------------------------------------------- G11 / G5: Thompson engine (REParser.matchSA)
def test_g11_thompson_over_toy_sa():
    """T/REParser.scala:219-234 — post2re("ma.b.") == re2post("mab"), post2re("ba|c.") == re2post("(b|a)c").  The strings are
    printed by SAISBuilder.nextSubstr, the mirror image of NaiveFMSearcher's (see test_g1)."""
    ix = _idx(b"mmabcacamabbbca"[::-1])
    r = ix.regex_match_thompson("mab")
    assert [(e - s, ix.nextSubstr(s, l)[::-1]) for l, s, e in r] == [(2, b"bam")]                   # "[2 Results] bam"
    r = ix.regex_match_thompson("(b|a)c")
    assert sorted((e - s, ix.nextSubstr(s, l)[::-1]) for l, s, e in r) == [(1, b"ca"), (2, b"cb")]    # List(ca, [2 Results] cb)


def test_g5_thompson_over_disk_index(ix1024):
    """T/REParser.scala:292-307 — the reference's only asserted regex-over-on-disk-index test:
    post2re("ba|d|e|c.") over test1024 == Set("ec", "dc", "[2 Results] ac", "bc")."""
    assert retree.re2poststr("(((b|a)|d)|e)c") == "ba|d|e|c·"                 # the test's postfix, as a regex
    r = ix1024.regex_match_thompson("(((b|a)|d)|e)c")
    assert r == ix1024.regex_match_thompson("(b|a|d|e)c")
    shown = {("[%d Results] " % (e - s) if e - s > 1 else "") + ix1024.nextSubstr(s, l).decode() for l, s, e in r}
    assert shown == {"ec", "dc", "[2 Results] ac", "bc"}
    assert r == [(2, 83, 85), (2, 85, 86), (2, 86, 87), (2, 87, 88)]


def test_thompson_differs_from_glushkov_where_the_reference_does(ix1024):
    # a position that reaches the MatchState emits and is still expanded: ab* yields every prefix, Glushkov trims b* away
    t = ix1024.regex_match_thompson("ab*")
    g = ix1024.regex_match("ab*")
    assert g == [(1,) + ix1024.search(b"a")] and set(g) <= set(t)
    for rx in ["a*", "a?", "[ab]c", "a|", "", "(a*)*", "a**"]:
        with pytest.raises(retree.ReUnsupported):
            retree.compile_thompson(rx)
    assert retree.compile_thompson("(a|b)c")["firsts"] == [0, 1]            # fine here, MatchError in ReTree (Q3)


# ---------------------------------------------------------------- DFA engine  T/dfa.scala:13-122
from dfa_cases import CDFKLM, DFA1, ab_star_c as _dfa_ab_star_c, class_b_star_c as _dfa_class_b_star_c  # noqa: E402


def test_dfa_match_string_and_buckets():
    from oracle import dfa as od
    d = od.DFA(_dfa_ab_star_c(od))                                                  # T/dfa.scala:62-68
    assert not d.matchString(b"absbc") and d.matchString(b"abbc") and d.matchString(b"abc")
    assert d.n_states == 4                                                          # :89-96
    assert [d.bucket_string(i) for i in range(4)] == ["DFAChar('a'->1)", "DFAChar('b'->2)", "DFAChar('b'->2),DFAChar('c'->3)", ""]
    d2 = od.DFA(_dfa_class_b_star_c(od, CDFKLM))            # :97-103
    assert [d2.bucket_string(i) for i in range(4)] == ["DFABucket('c-d' ->1),DFAChar('f'->1),DFABucket('k-m' ->1)", "DFAChar('b'->2)",
                                                       "DFAChar('b'->2),DFAChar('c'->3)", ""]
    d3 = od.DFA(_dfa_class_b_star_c(od, DFA1))   # :104-105
    assert d3.bucket_string(0) == "DFABucket('c-d' ->1),DFAChar('f'->1),DFABucket('l-m' ->1),DFABucket('\\xfa-\\xff' ->1)"


def test_dfa_match_sa_basics():
    """T/dfa.scala:108-120: `ab*c` over reverse("mmabcacadabbbca") gives exactly "cbbba" and "cba" (SAISBuilder.nextSubstr renders the
    mirrored string, as in G1/G2)."""
    from oracle import dfa as od
    ix = _idx(b"mmabcacadabbbca"[::-1])
    res = od.DFA(_dfa_ab_star_c(od)).matchSA(ix)
    assert len(res) == 2
    assert sorted(ix.nextSubstr(sp, ln)[::-1] for ln, sp, ep in res) == [b"cba", b"cbbba"]
    assert all(ep - sp == 1 for _, sp, ep in res)
    # a bucketed first step is never traversed: `[cd]b*c` finds nothing although "cb..." occurs (StatePoint.expand, M/dfa.scala:247-249)
    assert od.DFA(_dfa_class_b_star_c(od, ["c", "d"])).matchSA(ix) == []

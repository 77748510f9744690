"""
CPU-only checks of the product's host side: the C-ABI library loads and exports every declared symbol, the regex
compiler (C++) agrees table-for-table with the oracle's independent Python restatement, file validation mirrors the
reference's exceptions, and compute calls fail loudly without a GPU (no CPU fallback).
"""
import os
import random
import re

import numpy as np
import pytest

from findex_b200 import build as fbuild
from findex_b200 import fmindex as fx
from oracle import retree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    fbuild.build()


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fmgpu.h")).read()
    names = set(re.findall(r"\b(fmx_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    L = fx.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in L.fmx_version()


REGEXES = ["abcd", "abcd*", "abc*d", "a*bcd", "a*b*c*d*", "(ab)*", "(ab)*cd", "(ab)*(cd*)*", "(a|b)", "(a|b|d|c)",
           "(a|b*|d|c)", "(a|b*|d|c)*|(abc)", "(a|b|c)|(c|d|e)", "[a-c]", "a[b-d]e", "a[b-d]*e", "a[x.]e", "a\\de",
           "a+", "a****", "a+b", "a+((b|c)+|d)", "a*+", "a+*", "a+*+*++*", "a?", "(abc)?+|a?|bcd", "ab(cd|ef)+gh",
           "(10\\.[0-9]|[1-9][0-9]|[1-2][0-5][0-5]\\.[0-9]|[1-9][0-9]|[1-2][0-5][0-5]\\.[0-9]|[1-9][0-9]|[1-2][0-5][0-5])",
           "ab(cd)*ef", "ab*(cd)*(gh)*ij", "a(cd|ef)*j", ".*ab(cd)*(m(k|l)|tm*)(a|abc)(a*|(abc)*)ef(a*b*c*dg*)*gh",
           "a.*(b|c)d.*f", "x(ab?|d)c", "abc(cde)*ef", "ab?j", "a*(b|a)*bB*cd*e*", "a*(b|a)*b?B*c?d*e*", "abcdef",
           "(a|bX|cYZ)(a|b|c)", "(a|b|c)(a|b|c)", ".*(a|b)ca", "x(a|b|d|e)c", "ab?c[d-h]", "q(u|a)[a-m]z?k",
           "th(e|a)(n|t)\\w", "b(oo|ee)+k", "z[aeiou][aeiou]?z", "a\\.b", "a[\\]x]b", "a.b", "\\w+x", "a[aa]b", "a\\"]


@pytest.mark.parametrize("rx", REGEXES)
@pytest.mark.parametrize("line_only", [False, True])
def test_regex_compiler_matches_oracle_tables(rx, line_only):
    want = retree.compile_regex(rx, line_only).tables()
    got = fx.ReTree(rx, lineOnly=line_only).tables()
    assert got == want


BAD = ["*a", "(", "a(b", "|a", "()", "a||b", "[abc", "[a-", "[-a]", "[b-a]", "a(bc)d", "(a|b)c", "[a-c]d", "a|b*", "", "a|", "a)",
       "(a|b)|c", "(ab)(cd)", "a(b)", "(a)(b)", "a|(b|c)d"]


@pytest.mark.parametrize("rx", BAD)
def test_regex_errors_match_oracle(rx):
    try:
        retree.compile_regex(rx)
        want = None
    except retree.ReSyntaxError:
        want = fx.ReSyntaxError
    except retree.ReUnsupported:
        want = fx.ReUnsupported
    if want is None:
        fx.ReTree(rx)
    else:
        with pytest.raises(want):
            fx.ReTree(rx)


def _rand_regex(rng, depth=0):
    r = rng.random()
    if depth > 3 or r < 0.35:
        return rng.choice("abcdxyz")
    if r < 0.45:
        return rng.choice(["[a-c]", "\\d", ".", "[xyz]", "\\w"])
    if r < 0.65:
        return "".join(_rand_regex(rng, depth + 1) for _ in range(rng.randint(2, 4)))
    if r < 0.8:
        return "(" + "|".join(_rand_regex(rng, depth + 1) for _ in range(rng.randint(2, 3))) + ")"
    return _rand_regex(rng, depth + 1) + rng.choice("*+?")


def test_regex_compiler_fuzz_against_oracle():
    rng = random.Random(12345)
    seen = {"ok": 0, "syntax": 0, "unsupported": 0}
    for _ in range(3000):
        rx = _rand_regex(rng)
        try:
            want = retree.compile_regex(rx).tables()
            kind = "ok"
        except retree.ReSyntaxError:
            kind = "syntax"
        except retree.ReUnsupported:
            kind = "unsupported"
        seen[kind] += 1
        if kind == "ok":
            assert fx.ReTree(rx).tables() == want, rx
        else:
            with pytest.raises(fx.ReSyntaxError if kind == "syntax" else fx.ReUnsupported):
                fx.ReTree(rx)
    assert seen["ok"] > 300 and seen["unsupported"] > 300


def test_open_validates_files_like_the_reference(tmp_path, ref_dir):
    raw = open(os.path.join(ref_dir, "test1024.cmp.bwt"), "rb").read()
    aux = open(os.path.join(ref_dir, "test1024.cmp.aux"), "rb").read()
    (tmp_path / "x.bwt").write_bytes(raw[:-1])
    (tmp_path / "x.aux").write_bytes(aux)
    with pytest.raises(fx.FmxError) as e:                                   # "File %s bad size" bwtmerger.scala:153
        fx.GpuFMSearcher(str(tmp_path / "x.bwt"), bigEndian=False)
    assert e.value.code == fx.FMX_E_FORMAT and "bad size" in str(e.value)
    with pytest.raises(fx.FmxError) as e:                                   # wrong endianness flag
        fx.GpuFMSearcher(os.path.join(ref_dir, "test1024.cmp.bwt"), bigEndian=True)
    assert e.value.code == fx.FMX_E_FORMAT
    with pytest.raises(fx.FmxError) as e:
        fx.GpuFMSearcher(str(tmp_path / "missing.fm"))
    assert e.value.code == fx.FMX_E_IO
    (tmp_path / "y.bwt").write_bytes(raw)
    (tmp_path / "y.aux").write_bytes(aux)
    (tmp_path / "y.fm").write_bytes(bytes([8]) + b"\0" * 8)                 # elSize != 4, bwtmerger.scala:261
    with pytest.raises(fx.FmxError) as e:
        fx.GpuFMSearcher(str(tmp_path / "y.aux"), bigEndian=False)
    assert e.value.code == fx.FMX_E_FORMAT and "elSize" in str(e.value)
    (tmp_path / "y.fm").write_bytes(bytes([4]) + (1025).to_bytes(8, "little") + b"\0" * 7)     # size rule :262
    with pytest.raises(fx.FmxError) as e:
        fx.GpuFMSearcher(str(tmp_path / "y.aux"), bigEndian=False)
    assert e.value.code == fx.FMX_E_FORMAT
    os.remove(tmp_path / "y.fm")
    with pytest.raises(fx.FmxError) as e:                                   # the reference needs the .fm; opt-in strictness
        fx.GpuFMSearcher(str(tmp_path / "y.bwt"), bigEndian=False, require_fm=True)
    assert e.value.code == fx.FMX_E_IO


def test_no_gpu_means_loud_failure_not_cpu_fallback(ref_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fx.FmxError) as e:
        fx.GpuFMSearcher(os.path.join(ref_dir, "test1024.cmp.bwt"), bigEndian=False)
    assert e.value.code == fx.FMX_E_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(fx.FmxError):
        fx.build_bwt(b"abracadabra")


# ---------------------------------------------------------------- Thompson engine: product compiler vs oracle tables
T_REGEXES = ["mab", "(b|a)c", "(b|a|d|e)c", "ab*", "ab*c", "a+b", "a(bc)*d", "(ab|cd)+e", "a.b", "a\\db", "x\\w+y", "a?b", "(a|b)*c", "a(b|c)?d",
             "ab(cd|ef)+gh", "q(u|a).z?k", ".*ab", "a.*b", "((a|b)*aba*)*(a|b)(a|b)", "a+((b|c)+|d)", "a\\.b", "(a)(b)", "a(bc)d"]


@pytest.mark.parametrize("rx", T_REGEXES)
def test_thompson_compiler_matches_oracle_tables(rx):
    assert fx.ThompsonNFA(rx).tables() == retree.compile_thompson(rx)


@pytest.mark.parametrize("rx", ["a*", "a?", "[ab]c", "a|", "", "(a*)*", "a**", "(", "*a", "a(b", "x[a-c]", "(a?)*b", "(abc)?+|a?|bcd"])
def test_thompson_errors_match_oracle(rx):
    try:
        retree.compile_thompson(rx)
        want = None
    except retree.ReSyntaxError:
        want = fx.ReSyntaxError
    except retree.ReUnsupported:
        want = fx.ReUnsupported
    if want is None:
        fx.ThompsonNFA(rx)
    else:
        with pytest.raises(want):
            fx.ThompsonNFA(rx)


def test_thompson_compiler_fuzz_against_oracle():
    rng = random.Random(777)
    seen = {"ok": 0, "bad": 0}
    for _ in range(3000):
        rx = _rand_regex(rng)
        try:
            want = retree.compile_thompson(rx)
        except retree.ReSyntaxError:
            want = fx.ReSyntaxError
        except retree.ReUnsupported:
            want = fx.ReUnsupported
        if isinstance(want, dict):
            seen["ok"] += 1
            assert fx.ThompsonNFA(rx).tables() == want, rx
        else:
            seen["bad"] += 1
            with pytest.raises(want):
                fx.ThompsonNFA(rx)
    assert seen["ok"] > 500 and seen["bad"] > 200


def test_regex_compiler_limits_do_not_crash():
    deep = "(" * 5000 + "a" + ")" * 5000
    with pytest.raises(fx.FmxError) as e:
        fx.ReTree(deep)
    assert e.value.code == fx.FMX_E_LIMIT
    with pytest.raises(fx.FmxError) as e:
        fx.ThompsonNFA("a" * 70000)
    assert e.value.code == fx.FMX_E_LIMIT
    long_lit = "ab" * 20000                                   # 40 k positions: follows must stay linear
    t = fx.ReTree(long_lit).tables()
    assert len(t["c"]) == 40000 and t["last"][-1] == 1 and t["follows"][0] == [1]
    opt = "x" + "a?" * 400 + "y"                              # long nullable run: quadratic follows, deep Thompson closure
    assert fx.ReTree(opt).tables() == retree.compile_regex(opt).tables()
    assert fx.ThompsonNFA(opt).tables() == retree.compile_thompson(opt)


# ---------------------------------------------------------------------------------------------- DFA engine (dfa.scala), host side
def _random_automaton(m, rng, n_states, n_links, alphabet):
    s = m.StartState()
    states = [s] + [m.FinishState() if rng.random() < 0.3 else m.State(str(i)) for i in range(n_states - 1)]
    for _ in range(n_links):
        states[rng.randrange(n_states)].link(states[rng.randrange(n_states)], rng.choice(alphabet))
    return s, states


def test_dfa_host_side_matches_oracle():
    """fmx_dfa_create (numbering, moves, finish states, compileBuckets strings) and fmx_dfa_match_string against the oracle's
    restatement: the reference's three asserted automata (T/dfa.scala:13-105) and 300 random ones.  No GPU involved."""
    from findex_b200 import dfa as pd
    from oracle import dfa as od
    from dfa_cases import CDFKLM, DFA1, ab_star_c as _dfa_ab_star_c, class_b_star_c as _dfa_class_b_star_c
    fbuild.build()
    cases = [lambda m: _dfa_ab_star_c(m), lambda m: _dfa_class_b_star_c(m, CDFKLM), lambda m: _dfa_class_b_star_c(m, DFA1)]
    for mk in cases:
        p, o = pd.DFA.processLinkList(mk(pd)), od.DFA(mk(od))
        assert p.n_states == o.n_states and p.moves.tolist() == o.moves and p.finishStates == o.finish
        assert p.buckets == [o.bucket_string(i) for i in range(o.n_states)]
    p = pd.DFA(_dfa_ab_star_c(pd))
    assert p.buckets == ["DFAChar('a'->1)", "DFAChar('b'->2)", "DFAChar('b'->2),DFAChar('c'->3)", ""]      # T/dfa.scala:93-96 verbatim
    assert not p.matchString("absbc") and p.matchString("abbc") and p.matchString("abc")                 # :65-67
    for seed in range(300):
        ns, nl = 1 + seed % 9, seed % 40
        alphabet = [0, 1, 97, 98, 99, 100, 101, 200, 254, 255][: 2 + seed % 9]
        sp_, states_p = _random_automaton(pd, random.Random(seed), ns, nl, alphabet)
        so_, states_o = _random_automaton(od, random.Random(seed), ns, nl, alphabet)
        p, o = pd.DFA(sp_), od.DFA(so_)
        assert p.n_states == o.n_states and p.moves.tolist() == o.moves and p.finishStates == o.finish, seed
        assert [s.dfaIdx for s in states_p] == [s.dfaIdx for s in states_o], seed              # same numbering, unreachable = -1
        assert p.buckets == [o.bucket_string(i) for i in range(o.n_states)], seed
        rng = random.Random(1000 + seed)
        for _ in range(20):
            w = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 7)))
            assert p.matchString(w) == o.matchString(w), (seed, w)
        # the position automaton handed to the frontier kernel: one position per DFAChar action
        tb = p.tables()
        n_char = sum(1 for b in o.buckets for a in b if a[0] == "char")
        assert len(tb["c"]) == n_char and len(tb["firsts"]) == sum(1 for a in o.buckets[0] if a[0] == "char")


def test_dfa_create_rejects_bad_descriptions():
    import ctypes as C
    from findex_b200 import dfa as pd
    L = fx.lib()
    pd._declare(L)
    h = C.c_void_p()
    kind = np.array([1, 1], np.uint8)
    off = np.zeros(3, np.int32)
    z = np.zeros(1, np.int32)
    assert L.fmx_dfa_create(2, fx._ptr(kind), fx._ptr(off), fx._ptr(z), fx._ptr(z), C.byref(h)) == fx.FMX_E_ARG      # no start state
    kind = np.array([0, 0], np.uint8)
    assert L.fmx_dfa_create(2, fx._ptr(kind), fx._ptr(off), fx._ptr(z), fx._ptr(z), C.byref(h)) == fx.FMX_E_ARG      # two start states
    kind = np.array([0, 2], np.uint8)
    off = np.array([0, 1, 1], np.int32)
    to, ch = np.array([5], np.int32), np.array([97], np.int32)
    assert L.fmx_dfa_create(2, fx._ptr(kind), fx._ptr(off), fx._ptr(to), fx._ptr(ch), C.byref(h)) == fx.FMX_E_ARG    # link target out of range


# ---------------------------------------------------------------------------------------------- DFA.fromNFA (dfa.scala:343-389), host side
def _random_nfa(m, rng, n_states, n_links, alphabet):
    st = [m.NfaStartState()] + [m.NfaFinishState() if rng.random() < 0.25 else m.NfaState() for _ in range(n_states - 1)]
    for _ in range(n_links):
        a, b = st[rng.randrange(n_states)], st[rng.randrange(n_states)]
        if rng.random() < 0.3:
            a.epsilon(b)
        else:
            a.link(b, rng.choice(alphabet))
    return st[0]


def _thompson(m, rx):
    """Thompson construction of a tiny regex grammar (letters, concatenation, |, *, parentheses) out of the reference's NFA objects"""
    pos = [0]

    def atom():
        c = rx[pos[0]]
        if c == "(":
            pos[0] += 1
            s, e = alt()
            assert rx[pos[0]] == ")"
            pos[0] += 1
        else:
            s, e = m.NfaState(), m.NfaState()
            s.link(e, c)
            pos[0] += 1
        while pos[0] < len(rx) and rx[pos[0]] == "*":
            pos[0] += 1
            s2, e2 = m.NfaState(), m.NfaState()
            s2.epsilon(s); s2.epsilon(e2); e.epsilon(s); e.epsilon(e2)
            s, e = s2, e2
        return s, e

    def cat():
        s, e = atom()
        while pos[0] < len(rx) and rx[pos[0]] not in "|)":
            s2, e2 = atom()
            e.epsilon(s2)
            e = e2
        return s, e

    def alt():
        s, e = cat()
        while pos[0] < len(rx) and rx[pos[0]] == "|":
            pos[0] += 1
            s2, e2 = cat()
            a, b = m.NfaState(), m.NfaState()
            a.epsilon(s); a.epsilon(s2); e.epsilon(b); e2.epsilon(b)
            s, e = a, b
        return s, e

    s, e = alt()
    start, fin = m.NfaStartState(), m.NfaFinishState()
    start.epsilon(s)
    e.epsilon(fin)
    return start


def test_dfa_from_nfa_matches_oracle_and_the_nfa_language():
    """fmx_dfa_from_nfa against the oracle's restatement (same moves / finish states / bucket strings) on 200 random NFAs, and against
    what the NFA itself accepts: direct simulation under the reference's rule that the initial set never accepts, and Python's re for
    Thompson-built automata (words whose run comes back to the initial set are the documented exception)."""
    from findex_b200 import dfa as pd
    from oracle import dfa as od
    fbuild.build()
    for seed in range(200):
        ns, nl = 1 + seed % 8, seed % 30
        alphabet = [97, 98, 99, 100, 200, 255][: 2 + seed % 5]
        p = pd.DFA.fromNFA(_random_nfa(pd, random.Random(seed), ns, nl, alphabet))
        nfa_o = _random_nfa(od, random.Random(seed), ns, nl, alphabet)
        o = od.DFA(od.from_nfa(nfa_o))
        assert p.n_states == o.n_states and p.moves.tolist() == o.moves and p.finishStates == o.finish, seed
        assert p.buckets == [o.bucket_string(i) for i in range(o.n_states)], seed
        assert 0 not in p.finishStates
        rng = random.Random(5000 + seed)
        for _ in range(30):
            w = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 8)))
            assert p.matchString(w) == od.nfa_accepts(nfa_o, w), (seed, w)
    rng = random.Random(7)
    for rx in ["ab*c", "(a|b)*c", "a(b|c)*d", "(ab|cd)*e", "a*b*c", "((a|b)(c|d))*a", "abc", "a|b|c", "(a*b)*c"]:
        p = pd.DFA.fromNFA(_thompson(pd, rx))
        nfa_o = _thompson(od, rx)
        for _ in range(300):
            w = "".join(rng.choice("abcde") for _ in range(rng.randrange(0, 7)))
            got = p.matchString(w)
            assert got == od.nfa_accepts(nfa_o, w.encode()), (rx, w)
            if got:
                assert re.fullmatch(rx, w), (rx, w)              # never accepts outside the language
    # the initial set is the StartState and never accepts: `a*` as an NFA accepts "" and "a...", its DFA accepts nothing that returns there
    p = pd.DFA.fromNFA(_thompson(pd, "a*"))
    assert not p.matchString("") and 0 not in p.finishStates


def test_dfa_from_nfa_rejects_bad_descriptions():
    import ctypes as C
    from findex_b200 import dfa as pd
    L = fx.lib()
    pd._declare(L)
    L.fmx_dfa_from_nfa.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    h = C.c_void_p()
    fin = np.array([0, 1], np.uint8)
    off = np.array([0, 1, 1], np.int32)
    to, ch = np.array([1], np.int32), np.array([97], np.int32)
    assert L.fmx_dfa_from_nfa(2, fx._ptr(fin), 5, fx._ptr(off), fx._ptr(to), fx._ptr(ch), C.byref(h)) == fx.FMX_E_ARG     # initial out of range
    ch = np.array([300], np.int32)
    assert L.fmx_dfa_from_nfa(2, fx._ptr(fin), 0, fx._ptr(off), fx._ptr(to), fx._ptr(ch), C.byref(h)) == fx.FMX_E_ARG     # character out of range
    ch = np.array([97], np.int32)
    assert L.fmx_dfa_from_nfa(2, fx._ptr(fin), 0, fx._ptr(off), fx._ptr(to), fx._ptr(ch), C.byref(h)) == fx.FMX_OK
    L.fmx_regex_free(h)

"""
findex_b200.dfa — host-side mirror of the reference's DFA engine (src/main/scala/org/fmindex/dfa.scala) over libfmgpu.so.

The reference builds a DFA from state objects and links and searches it over a SuffixWalkingAlgo:

    val s = new StartState(); val a = new State("a"); val f = new FinishState()
    s.link(a, 'a'); a.link(f, 'c')
    val dfa = DFA.processLinkList(s)          // dfa.scala:391-407
    dfa.matchString("ac"); dfa.buckets(0).mkString(","); dfa.matchSA(sa)

Same names here; the numbering, the moves table, compileBuckets and the bucket strings are computed by the library
(fmx_dfa_create, fmx_regex.cpp), the search is DFA.matchSA (:261-289, 500-iteration cap off) on the GPU through the same frontier
kernel as the regex engines.  The object graph is only flattened on this side.
"""
import ctypes as C

import numpy as np

from .fmindex import ReTree, _check, _ptr, lib


class AnyState:
    """trait AnyState (dfa.scala:296-323).  link() prepends, like the reference."""
    KIND = 1

    def __init__(self, name="x"):
        self.name = name
        self.links = []
        self.dfaIdx = -1

    def link(self, to, chr_):
        self.links.insert(0, (to, chr_ if isinstance(chr_, int) else ord(chr_)))


class State(AnyState):
    KIND = 1


class StartState(AnyState):
    KIND = 0

    def __init__(self):
        super().__init__("START")


class FinishState(AnyState):
    KIND = 2

    def __init__(self):
        super().__init__("END")


def _declare(L):
    if getattr(L, "_dfa_declared", False):
        return
    p, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.fmx_dfa_create.argtypes = [i32, p, p, p, p, C.POINTER(p)]
    L.fmx_dfa_info.argtypes = [p, C.POINTER(i32), p, p, p, i32]
    L.fmx_dfa_buckets.argtypes = [p, i32, C.c_char_p, i64, C.POINTER(i64)]
    L.fmx_dfa_match_string.argtypes = [p, p, i64, C.POINTER(i32)]
    L._dfa_declared = True


class DFA(ReTree):
    """DFA.processLinkList(start).  The handle is an fmx_regex, so searcher.regex_search_batch / regex_set accept it, also mixed
    with ReTree / ThompsonNFA handles."""

    def __init__(self, start):                              # noqa: super().__init__ compiles a regex string; a DFA is built from states
        L = lib()
        _declare(L)
        states, seen, todo = [], {}, [start]
        while todo:                                         # any order: the library does the reference's numbering
            s = todo.pop()
            if id(s) in seen:
                continue
            seen[id(s)] = len(states)
            states.append(s)
            todo.extend(to for to, _ in s.links)
        kind = np.array([s.KIND for s in states], np.uint8)
        off = np.zeros(len(states) + 1, np.int32)
        to, ch = [], []
        for i, s in enumerate(states):
            for t, c in s.links:
                to.append(seen[id(t)])
                ch.append(c)
            off[i + 1] = len(to)
        to = np.array(to if to else [0], np.int32)
        ch = np.array(ch if ch else [0], np.int32)
        h = C.c_void_p()
        _check(L.fmx_dfa_create(len(states), _ptr(kind), _ptr(off), _ptr(to), _ptr(ch), C.byref(h)))
        self.h = h
        self.regex = b"<dfa>"
        n = C.c_int32()
        number = np.zeros(len(states), np.int32)
        _check(L.fmx_dfa_info(h, C.byref(n), None, None, _ptr(number), len(states)))
        self.n_states = n.value
        moves = np.zeros((self.n_states, 256), np.int32)
        fin = np.zeros(self.n_states, np.uint8)
        _check(L.fmx_dfa_info(h, None, _ptr(moves), _ptr(fin), None, 0))
        self.moves = moves
        self.finishStates = set(np.flatnonzero(fin).tolist())
        for s, k in zip(states, number.tolist()):
            s.dfaIdx = k

    processLinkList = classmethod(lambda cls, start, debugLevel=0: cls(start))

    @property
    def buckets(self):
        """buckets(i).mkString(",") for every state"""
        out = []
        for i in range(self.n_states):
            need = C.c_int64()
            lib().fmx_dfa_buckets(self.h, i, None, 0, C.byref(need))
            buf = C.create_string_buffer(max(need.value, 1))
            _check(lib().fmx_dfa_buckets(self.h, i, buf, need.value, None))
            out.append(buf.value.decode("latin-1"))
        return out

    def matchString(self, s):
        if isinstance(s, str):
            s = s.encode("latin-1")
        a = np.frombuffer(bytes(s), np.uint8) if s else np.zeros(1, np.uint8)
        m = C.c_int32()
        _check(lib().fmx_dfa_match_string(self.h, _ptr(a), len(s), C.byref(m)))
        return bool(m.value)

    def matchSA(self, sa):
        """DFA.matchSA(sa) with the iteration cap off: sorted list of (len, sp, ep)."""
        return sa.regex_search_batch([self])[0]


# ---- NFA side: NfaBaseState / NfaState / NfaStartState / NfaFinishState with link() and epsilon() (dfa.scala:5-37, 90-95) and
# DFA.fromNFA(initialState) (:343-389); the subset construction runs in the library (fmx_dfa_from_nfa).
class NfaBaseState:
    FINISH = False

    def __init__(self):
        self.links = []                     # (to, chr); chr = -1 for an EpsilonLink

    def link(self, to, chr_):
        self.links.insert(0, (to, chr_ if isinstance(chr_, int) else ord(chr_)))

    def epsilon(self, to):
        self.links.append((to, -1))


class NfaState(NfaBaseState):
    pass


class NfaStartState(NfaBaseState):
    pass


class NfaFinishState(NfaBaseState):
    FINISH = True


def _from_nfa(cls, initial):
    L = lib()
    _declare(L)
    L.fmx_dfa_from_nfa.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    states, seen, todo = [], {}, [initial]
    while todo:
        s = todo.pop()
        if id(s) in seen:
            continue
        seen[id(s)] = len(states)
        states.append(s)
        todo.extend(to for to, _ in s.links)
    fin = np.array([1 if s.FINISH else 0 for s in states], np.uint8)
    off = np.zeros(len(states) + 1, np.int32)
    to, ch = [], []
    for i, s in enumerate(states):
        for t, c in s.links:
            to.append(seen[id(t)])
            ch.append(c)
        off[i + 1] = len(to)
    to = np.array(to if to else [0], np.int32)
    ch = np.array(ch if ch else [0], np.int32)
    h = C.c_void_p()
    _check(L.fmx_dfa_from_nfa(len(states), _ptr(fin), seen[id(initial)], _ptr(off), _ptr(to), _ptr(ch), C.byref(h)))
    self = cls.__new__(cls)
    self.h = h
    self.regex = b"<dfa from nfa>"
    n = C.c_int32()
    _check(L.fmx_dfa_info(h, C.byref(n), None, None, None, 0))
    self.n_states = n.value
    moves = np.zeros((self.n_states, 256), np.int32)
    f = np.zeros(self.n_states, np.uint8)
    _check(L.fmx_dfa_info(h, None, _ptr(moves), _ptr(f), None, 0))
    self.moves = moves
    self.finishStates = set(np.flatnonzero(f).tolist())
    return self


DFA.fromNFA = classmethod(_from_nfa)

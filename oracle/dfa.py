"""
oracle/dfa.py — CPU restatement of the reference's DFA engine.  TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing in the
product); it is the checker, never the thing measured or shipped.

Follows src/main/scala/org/fmindex/dfa.scala ("M/dfa.scala"):
    AnyState / State / StartState / FinishState, Link          M/dfa.scala:291-336
    DFA.processLinkList (reachability, numbering, moves)        M/dfa.scala:391-407, addState :129-143, addLink :153-157
    DFA.compileBuckets, DFAChar / DFABucket and their toString  M/dfa.scala:172-223
    DFA.matchString                                             M/dfa.scala:160-171
    DFA.matchSA, StatePoint.expand                              M/dfa.scala:238-289

Pinned: every assertion of the reference's DFATests (src/test/scala/org/fmindex/tests/dfa.scala:62-122) holds against this file
(tests/test_oracle_golden.py::test_dfa_*): matchString on `ab*c`, the three compileBuckets strings, and the two results "cbbba" / "cba"
of matchSA over reverse("mmabcacadabbbca").

PARITY UNPINNED for `from_nfa` (DFA.fromNFA, M/dfa.scala:343-389, at the end of this file): no reference test touches it; it is checked
against direct NFA simulation instead.

Behaviour kept on purpose:
  * a run of two or more consecutive characters with the same target becomes a DFABucket, and StatePoint.expand only follows
    DFAChar actions (`case _ => None`, :247-249): character ranges are never traversed by matchSA.
  * a state that is a finish state emits its (len, sp, ep) and is expanded all the same (:272-274).
  * numbering: the start state is 0, the others are numbered in the iteration order of the `visited` Set (:400-401).  Scala's Set1..Set4
    keep insertion order (which the reference's tests rely on: 4 states); for larger automata the order is hash-dependent and
    unspecified — insertion order is used here.  Results of matchSA/matchString do not depend on the numbering.
Deviations, as everywhere in this repo: characters are unsigned (the reference indexes moves with a signed Byte, :165, and passes a
signed Byte to getPrevRange, :245 — both fail for bytes >= 0x80), and the 500-iteration cap of matchSA (:269), whose effect
depends on the hash order of a Set, is off: the parity object is the full result multiset.
"""


class AnyState:
    """trait AnyState (M/dfa.scala:296-323): links are prepended."""
    kind = 1

    def __init__(self, name="x"):
        self.name = name
        self.links = []                     # [(to, chr)], head = most recently added
        self.dfaIdx = -1

    def link(self, to, chr_):
        self.links.insert(0, (to, chr_ if isinstance(chr_, int) else ord(chr_)))


class State(AnyState):
    kind = 1


class StartState(AnyState):
    kind = 0

    def __init__(self):
        super().__init__("START")


class FinishState(AnyState):
    kind = 2

    def __init__(self):
        super().__init__("END")


def _reach(s, visited):
    """_processLinkList (M/dfa.scala:392-398): note that the membership test uses the set passed IN, not the growing one."""
    v = list(visited)
    if s not in v:
        v.append(s)
    for to, _ in s.links:
        if to not in visited:
            for x in _reach(to, v):
                if x not in v:
                    v.append(x)
    return v


def _pretty(c):
    return "\\x%x" % c if (c < 0x20 or c > 0x7e) else chr(c)


class DFA:
    def __init__(self, start):
        """DFA.processLinkList(start)"""
        visited = _reach(start, [])
        self.n_states = len(visited)
        self.moves = [[-1] * 256 for _ in range(self.n_states)]
        self.finish = set()
        idx = 1
        for v in visited:                                   # addState
            assert v.dfaIdx < 0, "State already used"
            if v.kind == 0:
                v.dfaIdx = 0
            else:
                v.dfaIdx = idx
                if v.kind == 2:
                    self.finish.add(idx)
                idx += 1
        for v in visited:                                   # addLink
            for to, c in v.links:
                self.moves[v.dfaIdx][c] = to.dfaIdx
        self.states = visited
        self.buckets = self._compile_buckets()

    def _compile_buckets(self):
        """compileBuckets (M/dfa.scala:198-223): per state a list of ('char', target, c) / ('bucket', target, c1, c2)."""
        out = []
        for row in self.moves:
            b, last, start = [], -1, -1
            for j in range(256):
                v = row[j]
                if last != v:
                    if last != -1:
                        b.append(self._action(last, start, j - 1))
                    start, last = j, v
            if last != -1:
                b.append(self._action(last, start, 255))
            out.append(b)
        return out

    @staticmethod
    def _action(state, c1, c2):
        return ("char", state, c1) if c1 == c2 else ("bucket", state, c1, c2)

    def bucket_string(self, i):
        """buckets(i).mkString(",") with the toString of DFAChar / DFABucket (M/dfa.scala:190-196)"""
        parts = []
        for a in self.buckets[i]:
            if a[0] == "char":
                parts.append("DFAChar('%s'->%d)" % (_pretty(a[2]), a[1]))
            else:
                parts.append("DFABucket('%s-%s' ->%d)" % (_pretty(a[2]), _pretty(a[3]), a[1]))
        return ",".join(parts)

    def matchString(self, s):
        cur = 0
        for c in bytes(s):
            cur = self.moves[cur][c]
            if cur == -1:
                return False
        return cur in self.finish

    def matchSA(self, sa):
        """uncapped DFA.matchSA: sorted list of (len, sp, ep)"""
        front = [(0, 0, 0, sa.n)]                            # StatePoint(state, len, sp, ep); `statesFront` is a Set in the reference
        queued = set(front)
        visited = set()
        results = []
        while front:
            st = front.pop()
            queued.discard(st)
            visited.add(st)
            state, ln, sp, ep = st
            new = []
            for a in self.buckets[state]:
                if a[0] != "char":
                    continue                                 # DFABucket: `case _ => None`
                r = sa.getPrevRange(sp, ep, a[2])
                if r is not None:
                    new.append((a[1], ln + 1, r[0], r[1]))
            if state in self.finish:
                results.append((ln, sp, ep))
            for s in new:
                if s not in visited and s not in queued:
                    front.append(s)
                    queued.add(s)
        return sorted(results)


# ---------------------------------------------------------------------------------------------------------------------------
# NFA side and DFA.fromNFA  (M/dfa.scala:5-37 NfaBaseState / NfaLink / EpsilonLink, :97-108 NFA.epsilons / epsilonTransitions,
# :343-389 DFA.fromNFA).  No reference test asserts anything about fromNFA (only the DFAPlay playground calls it): PARITY UNPINNED for
# this part — what is checked instead is that the DFA accepts exactly what the NFA accepts under the reference's own rule that the
# initial state set is the StartState and never a FinishState.
class NfaBaseState:
    finish = False

    def __init__(self):
        self.links = []                     # (to, chr) with chr = None for an EpsilonLink; link() prepends, epsilon() appends (:34-35)

    def link(self, to, chr_):
        self.links.insert(0, (to, chr_ if isinstance(chr_, int) else ord(chr_)))

    def epsilon(self, to):
        self.links.append((to, None))

    def epsilons(self):
        seen, stack = [self], [self]
        while stack:
            s = stack.pop()
            for to, c in s.links:
                if c is None and all(to is not x for x in seen):
                    seen.append(to)
                    stack.append(to)
        return seen


class NfaState(NfaBaseState):
    pass


class NfaStartState(NfaBaseState):
    pass


class NfaFinishState(NfaBaseState):
    finish = True


def from_nfa(initial):
    """DFA.fromNFA(initial): returns the StartState of the equivalent object DFA (feed it to DFA(...)).  Sets are discovered
    breadth-first with characters ascending and links are added in (source set, character) order — the reference leaves both to hash
    iteration order, which only affects the numbering."""
    def closure(states):
        out = []
        for s in states:
            for e in s.epsilons():
                if all(e is not x for x in out):
                    out.append(e)
        return frozenset(id(x) for x in out), out

    key0, set0 = closure([initial])
    order, members, trans = [key0], {key0: set0}, {}
    i = 0
    while i < len(order):
        k = order[i]
        i += 1
        by_chr = {}
        for s in members[k]:
            for to, c in s.links:
                if c is not None:
                    by_chr.setdefault(c, []).append(to)
        trans[k] = []
        for c in sorted(by_chr):
            tk, tset = closure(by_chr[c])
            if tk not in members:
                members[tk] = tset
                order.append(tk)
            trans[k].append((c, tk))
    dstate = {}
    for k in order:
        if k == key0:
            dstate[k] = StartState()
        elif any(s.finish for s in members[k]):
            dstate[k] = FinishState()
        else:
            dstate[k] = State(str(len(dstate)))
    for k in order:
        for c, tk in trans[k]:
            dstate[k].link(dstate[tk], c)
    return dstate[key0]


def nfa_accepts(initial, word):
    """Direct simulation of the NFA under the reference's acceptance rule for its DFAs: accepted iff the state set after the word holds
    an NfaFinishState and is not the initial state's own epsilon closure (that set is the StartState, M/dfa.scala:353-359)."""
    def closure(states):
        out = []
        for s in states:
            for e in s.epsilons():
                if all(e is not x for x in out):
                    out.append(e)
        return out
    init = closure([initial])
    cur = init
    for c in bytes(word):
        nxt = [to for s in cur for to, ch in s.links if ch == c]
        if not nxt:
            return False
        cur = closure(nxt)
    return any(s.finish for s in cur) and {id(x) for x in cur} != {id(x) for x in init}

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

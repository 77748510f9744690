"""
oracle/retree.py — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT.

Pure-Python restatement of the reference's regex front-end and Glushkov ("ReTree") construction:

  * re2post            M/re2/re2.scala:50-185     regex text -> postfix token list
  * ReTree.apply       M/re2/retree.scala:156-370 postfix -> position-automaton tree
  * node semantics     M/re2/retree.scala:10-155  firsts / isNull / follows / isLast
  * postProcess        M/re2/retree.scala:439-482 ('+' rewrite, deep copy)
  * removeBorderNulls  M/re2/retree.scala:371-385
  * setParents/setNums M/re2/retree.scala:386-423

(M/ = /root/reference/src/main/scala/org/fmindex/.)  The quirks Q1-Q5 of SURVEY.md §8(a) are part of
the behaviour and are reproduced, not fixed.  Parity status: PINNED against the reference's own
assertions (T/REParser.scala:10-26, 481-588, 594-605) in tests/test_oracle_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""

MIN_CHAR = 2
MAX_CHAR = 255


class ReSyntaxError(Exception):
    """The reference throws Exception("re2post syntax")."""


class ReUnsupported(Exception):
    """The reference throws MatchError / NoSuchElementException (construction is partial, Q3)."""


# ------------------------------------------------------------------ postfix tokens (re2.scala:24-48)
class Tok:
    __slots__ = ("kind", "c", "start", "end", "alts")

    def __init__(self, kind, c=None, start=None, end=None, alts=None):
        self.kind, self.c, self.start, self.end, self.alts = kind, c, start, end, alts

    def __str__(self):  # the reference's toString, used by the re2poststr goldens
        k = self.kind
        if k == "char":
            return chr(self.c)
        if k == "interval":
            return "." if (self.start == MIN_CHAR and self.end == MAX_CHAR) else "[%c-%c]" % (self.start, self.end)
        if k == "alt":
            return "[" + "".join(chr(x) for x in reversed(self.alts)) + "]"
        return {"cat": "·", "star": "*", "quest": "?", "plus": "+", "or": "|"}[k]


def re2post(s, line_only=False):
    """re2.scala:50-185.  `s` is bytes (or a latin-1 str); returns list[Tok]."""
    if isinstance(s, str):
        s = s.encode("latin-1")
    l = len(s)
    i = 0
    natom = 0
    nalt = 0
    dst = []          # built in forward order (the reference prepends and reverses at the end)
    stack = []
    quoted = False

    def process_char(c, q):
        nonlocal natom
        if natom > 1:
            natom -= 1
            dst.append(Tok("cat"))
        if q:
            if c == ord("w"):
                dst.append(Tok("interval", start=ord("A"), end=ord("z")))
            elif c == ord("d"):
                dst.append(Tok("interval", start=ord("0"), end=ord("9")))
            else:
                dst.append(Tok("char", c=c))
        else:
            if c == ord("."):
                dst.append(Tok("interval", start=0x20 if line_only else MIN_CHAR, end=MAX_CHAR))
            else:
                dst.append(Tok("char", c=c))
        natom += 1

    def process_alt(i0):
        # re2.scala:76-119 ; alts is the reference's list (built by prepending => head = last added)
        nonlocal natom
        i = i0
        alts = []          # alts[0] is the head
        q = False
        end = False
        interval = False

        def pc(c):
            nonlocal interval, alts
            if interval:
                if not alts:
                    raise ReSyntaxError("re2post syntax")
                c_alt = alts[0] + 1
                e_alt = c
                if c_alt > e_alt:
                    raise ReSyntaxError("re2post syntax")
                while c_alt <= e_alt:
                    alts.insert(0, c_alt)
                    c_alt += 1
                interval = False
            else:
                alts.insert(0, c)

        while i < l and not end:
            c = s[i]
            if q:
                pc(c)
                q = False
            elif c == ord("\\"):
                q = True
            elif c == ord("-"):
                interval = True
            elif c == ord("]"):
                end = True
            else:
                pc(c)
            i += 1
        if (not end) or interval:
            raise ReSyntaxError("re2post syntax")
        if natom > 1:
            natom -= 1
            dst.append(Tok("cat"))
        dst.append(Tok("alt", alts=list(alts)))
        natom += 1
        return i

    while i < l:
        c = s[i]
        if not quoted:
            if c == ord("("):
                if natom > 1:
                    natom -= 1
                    dst.append(Tok("cat"))
                stack.append((nalt, natom))
                nalt = 0
                natom = 0
            elif c == ord("|"):
                if natom == 0:
                    raise ReSyntaxError("re2post syntax")
                natom -= 1
                while natom > 0:
                    dst.append(Tok("cat"))
                    natom -= 1
                nalt += 1
            elif c == ord(")"):
                if natom == 0:
                    raise ReSyntaxError("re2post syntax")
                natom -= 1
                while natom > 0:
                    dst.append(Tok("cat"))
                    natom -= 1
                while nalt > 0:
                    dst.append(Tok("or"))
                    nalt -= 1
                if not stack:
                    raise ReUnsupported("NoSuchElementException: unbalanced ')'")   # Stack.pop on empty
                nalt, natom = stack.pop()
                natom += 1
            elif c == ord("["):
                i = process_alt(i + 1) - 1
            elif c == ord("\\"):
                quoted = True
            elif c in (ord("*"), ord("+"), ord("?")):
                if natom == 0:
                    raise ReSyntaxError("re2post syntax")
                dst.append(Tok({ord("*"): "star", ord("+"): "plus", ord("?"): "quest"}[c]))
            else:
                process_char(c, False)
        else:
            process_char(c, True)
            quoted = False
        i += 1
    if stack:
        raise ReSyntaxError("re2post syntax")
    natom -= 1
    while natom > 0:
        dst.append(Tok("cat"))
        natom -= 1
    while nalt > 0:
        dst.append(Tok("or"))
        nalt -= 1
    return dst


def re2poststr(s):
    return "".join(str(t) for t in re2post(s))


# ------------------------------------------------------------------ tree nodes (retree.scala:10-155)
class Node:
    kind = "?"

    def __init__(self):
        self.childs = []       # scala List; index 0 is the head
        self.parent = None     # None stands for RootNode

    def append(self, n):
        n.parent = self
        self.childs.insert(0, n)          # childs ::= n

    @property
    def is_unar(self):
        return self.kind in ("star", "quest", "plus")


class CharNode(Node):
    kind = "char"

    def __init__(self, c):
        super().__init__()
        self.c = c
        self.num = 0

    def __str__(self):
        return chr(self.c) if 0x20 <= self.c < 0x7F else "%02x" % self.c


class StarNode(Node):
    kind = "star"

    def __str__(self):
        return "*[" + ",".join(map(str, self.childs)) + "]"


class QuestNode(Node):
    kind = "quest"

    def __str__(self):
        return "?[" + ",".join(map(str, self.childs)) + "]"


class PlusNode(Node):
    kind = "plus"

    def __str__(self):
        return "+[" + ",".join(map(str, self.childs)) + "]"


class OrNode(Node):
    kind = "or"

    def append(self, n):
        if n.kind == "or":                 # splice: childs :::= on.childs
            for ch in n.childs:
                ch.parent = self
            self.childs = list(n.childs) + self.childs
        else:
            n.parent = self
            self.childs.insert(0, n)

    def __str__(self):
        return "O[" + "|".join(map(str, self.childs)) + "]"


class FollowNode(Node):
    kind = "follow"

    def __str__(self):
        return "F[" + ",".join(map(str, self.childs)) + "]"


def _cls(n):
    """C / O / F / U classification used by the pattern matches in ReTree.apply."""
    if n.kind == "char":
        return "C"
    if n.kind == "or":
        return "O"
    if n.kind == "follow":
        return "F"
    return "U"


# pattern-match tables, retree.scala:181-239 (Or) and :240-295 (Concat); anything else => MatchError (Q3)
_OR_APPEND_TO_A2 = {("C", "O"), ("U", "O"), ("F", "O"), ("O", "O")}
_OR_NEW = {("F", "F"), ("C", "C"), ("U", "F"), ("C", "F"), ("U", "C"), ("F", "C"), ("U", "U")}
_CAT_NEW = {("O", "O"), ("C", "O"), ("C", "C"), ("U", "C"), ("U", "O"), ("C", "U"), ("U", "U")}
_CAT_APPEND_TO_A1 = {("F", "C"), ("F", "O"), ("F", "U")}


# ------------------------------------------------------------------ semantic functions (lazy vals)
def is_null(n):
    k = n.kind
    if k == "char":
        return False
    if k in ("star", "quest"):
        return True
    if k == "plus":
        return all(is_null(c) for c in n.childs)
    if k == "or":
        return any(is_null(c) for c in n.childs)
    return all(is_null(c) for c in n.childs)          # follow


def firsts(n):
    k = n.kind
    if k == "char":
        return [n]
    if k in ("star", "quest", "plus", "or"):
        out = []
        for c in n.childs:
            out.extend(firsts(c))
        return out
    # follow, retree.scala:117-127   (ret :::= x  means  ret = x ::: ret)
    p = list(n.childs)
    ret = []
    while p and is_null(p[0]):
        ret = firsts(p[0]) + ret
        p = p[1:]
    if p:
        ret = firsts(p[0]) + ret
    return ret


def _siblings_after(parent, node):
    idx = next(i for i, ch in enumerate(parent.childs) if ch is node)      # dropWhile(_ != this).tail
    return parent.childs[idx + 1:]


def follows(n):
    """retree.scala:14-38"""
    p = n.parent
    if p is None:
        return []
    if p.kind == "or":
        return follows(p)
    if p.kind == "follow":
        last = _siblings_after(p, n)
        if last:
            ret = firsts(last[0])
            if is_null(last[0]):
                last = last[1:]
                while last and is_null(last[0]):
                    ret = firsts(last[0]) + ret
                    last = last[1:]
                if last:
                    ret = firsts(last[0]) + ret
            return ret
        return follows(p)
    if p.kind == "star":
        return firsts(n) + follows(p)
    if p.kind == "quest":
        return follows(p)
    return []                                            # PlusNode parent


def is_last(n):
    """retree.scala:40-50"""
    p = n.parent
    if p is None:
        return True
    if p.kind == "or" or p.is_unar:
        return is_last(p)
    if p.kind == "follow":
        last = _siblings_after(p, n)
        if (not last) or all(is_null(x) for x in last):
            return is_last(p)
        return False
    return True


# ------------------------------------------------------------------ construction
def post_process(r):
    """retree.scala:439-482: deep copy, child lists re-reversed into forward order, Plus -> x, x*"""
    def process_child(new_l, old_c):
        if old_c.kind == "plus":
            a1 = post_process(old_c.childs[0])
            a2 = StarNode()
            a2.append(post_process(old_c.childs[0]))
            return [a1, a2] + new_l
        return [post_process(old_c)] + new_l

    k = r.kind
    if k == "char":
        return CharNode(r.c)
    if k == "follow":
        nc = FollowNode()
    elif k == "quest":
        nc = QuestNode()
    elif k == "or":
        nc = OrNode()
    elif k == "star":
        nc = StarNode()
    else:
        raise ReUnsupported("MatchError in postProcess: " + k)
    for ch in r.childs:
        nc.childs = process_child(nc.childs, ch)
    return nc


def remove_border_nulls(a1):
    """retree.scala:371-385"""
    n = FollowNode()
    p = list(a1.childs)
    while p and is_null(p[0]):
        p = p[1:]
    p = p[::-1]
    while p and is_null(p[0]):
        p = p[1:]
    while p:
        n.append(p[0])
        p = p[1:]
    return n


def set_parents(r, parent=None):
    r.parent = parent
    for ch in r.childs:
        set_parents(ch, r)


def set_nums(r):
    """retree.scala:393-423"""
    def _set(r, idx0):
        idx = [idx0]

        def __set(r):
            if r.kind == "or":
                nidx = idx[0]
                for ch in r.childs:
                    if ch.kind == "char":
                        ch.num = idx[0]
                        nidx = max(nidx, idx[0] + 1)
                    else:
                        nidx = max(nidx, _set(ch, idx[0]))
                idx[0] = nidx
            else:
                for ch in r.childs:
                    if ch.kind == "char":
                        ch.num = idx[0]
                        idx[0] += 1
                    else:
                        __set(ch)
            return idx[0]

        return __set(r)

    return _set(r, 1)


class ReTree:
    def __init__(self, root):
        self.root = root

    # flat tables handed to the traversal (C oracle) and compared against the product's compiler
    def tables(self):
        states = []

        def walk(n):
            if n.kind == "char":
                n.sid = len(states)
                states.append(n)
            for ch in n.childs:
                walk(ch)

        walk(self.root)
        c = [s.c for s in states]
        last = [1 if is_last(s) else 0 for s in states]
        num = [s.num for s in states]
        fol = [[f.sid for f in follows(s)] for s in states]
        first = [f.sid for f in firsts(self.root)]
        return {"c": c, "last": last, "num": num, "follows": fol, "firsts": first}


def build(post, remove_nulls=True):
    """ReTree.apply, retree.scala:156-370"""
    args = []

    def pop():
        if not args:
            raise ReUnsupported("NoSuchElementException: empty stack")
        return args.pop()

    for t in post:
        k = t.kind
        if k == "interval":
            el = OrNode()
            j = t.start
            while j < t.end:                               # exclusive upper bound (Q1)
                el.append(CharNode(j))
                j += 1
            args.append(el)
        elif k == "alt":
            el = OrNode()
            for c in t.alts:
                el.append(CharNode(c))
            args.append(el)
        elif k == "char":
            args.append(CharNode(t.c))
        elif k == "or":
            a2 = pop()
            a1 = pop()
            key = (_cls(a1), _cls(a2))
            if key in _OR_APPEND_TO_A2:
                a2.append(a1)
                args.append(a2)
            elif key in _OR_NEW:
                el = OrNode()
                el.append(a1)
                el.append(a2)
                args.append(el)
            else:
                raise ReUnsupported("MatchError: OrPoint have no match for a1=%s a2=%s" % (a1, a2))
        elif k == "cat":
            a2 = pop()
            a1 = pop()
            key = (_cls(a1), _cls(a2))
            if key in _CAT_NEW:
                el = FollowNode()
                el.append(a1)
                el.append(a2)
                args.append(el)
            elif key in _CAT_APPEND_TO_A1:
                a1.append(a2)
                args.append(a1)
            else:
                raise ReUnsupported("MatchError: ConcatPoint have no match for a1=%s a2=%s" % (a1, a2))
        elif k in ("plus", "star"):
            a1 = pop()
            if a1.kind == "star":
                args.append(a1)
            elif a1.kind in ("quest", "plus"):
                el = StarNode()
                el.append(a1.childs[0])
                args.append(el)
            else:
                el = PlusNode() if k == "plus" else StarNode()
                el.append(a1)
                args.append(el)
        elif k == "quest":
            a1 = pop()
            if a1.kind == "quest":
                el = QuestNode()
                el.append(a1.childs[0])
                args.append(el)
            elif a1.kind == "star":
                args.append(a1)
            elif a1.kind == "plus":
                el = StarNode()
                el.append(a1.childs[0])
                args.append(el)
            else:
                el = QuestNode()
                el.append(a1)
                args.append(el)
        else:
            raise AssertionError(k)
    a0 = pop()
    if a0.kind == "follow":
        a2 = a0
    else:
        a2 = FollowNode()
        a2.append(a0)
    a1 = post_process(a2)
    a3 = remove_border_nulls(a1) if remove_nulls else a1
    set_parents(a3, None)
    set_nums(a3)
    return ReTree(a3)


def compile_regex(s, line_only=False, remove_nulls=True):
    return build(re2post(s, line_only), remove_nulls)


# ====================================================================================================================
# Thompson engine: REParser.createNFA (M/re2/re2.scala:264-334) + REParser.matchSA (:568-693), restated.
#
# matchSA carries StatePoint(len, state, intervals): every interval of a state point is expanded independently and all
# survivors are handed to every terminal state of the epsilon-closure of `next` (BaseState.outStates, :213-226); a
# MatchState in that closure emits one SAResult(len+1, sp, ep) per interval, the other states are enqueued.  As a multiset
# of (len, sp, ep) this equals an item-wise traversal over "positions", where an IntervalState(start, end) contributes one
# position per char in `start until end` (exclusive end, :472).  Unlike the Glushkov engine a position may both emit and
# go on.  Tables: flags bit0 = emits (MatchState in the closure of next), bit1 = stop after emitting (never set here).
# What throws in the reference: AltPoint has no case in createNFA (MatchError); a nullable regex puts MatchState into the
# start front and StatePoint.expand has no case for it (MatchError); stack underflow (NoSuchElementException).
# ====================================================================================================================
class _TState:
    __slots__ = ("kind", "c", "start", "end", "out", "out1", "out2", "idx")

    def __init__(self, kind, c=0, start=0, end=0):
        self.kind, self.c, self.start, self.end = kind, c, start, end
        self.out = [None]            # LinkState of Const / Interval
        self.out1 = [None]           # LinkStates of Split
        self.out2 = [None]


_MATCH = _TState("match")


def thompson_nfa(post):
    """createNFA: returns the start state."""
    st = []

    def pop():
        if not st:
            raise ReUnsupported("NoSuchElementException: empty stack")
        return st.pop()

    def patch(outs, s):
        for link in outs:
            link[0] = s

    for t in post:
        k = t.kind
        if k == "quest":
            start, outs = pop()
            ns = _TState("split")
            ns.out1[0] = start
            st.append((ns, [ns.out2] + outs))
        elif k == "star":
            start, outs = pop()
            ns = _TState("split")
            ns.out1[0] = start
            patch(outs, ns)
            st.append((ns, [ns.out2]))
        elif k == "plus":
            start, outs = pop()
            ns = _TState("split")
            ns.out1[0] = start
            patch(outs, ns)
            st.append((start, [ns.out2]))
        elif k == "cat":
            s2, o2 = pop()
            s1, o1 = pop()
            patch(o1, s2)
            st.append((s1, o2))
        elif k == "or":
            s2, o2 = pop()
            s1, o1 = pop()
            ns = _TState("split")
            ns.out1[0] = s1
            ns.out2[0] = s2
            st.append((ns, o1 + o2))
        elif k == "char":
            ns = _TState("const", c=t.c)
            st.append((ns, [ns.out]))
        elif k == "interval":
            ns = _TState("interval", start=t.start, end=t.end)
            st.append((ns, [ns.out]))
        else:
            raise ReUnsupported("MatchError: createNFA has no case for " + k)
    start, outs = pop()
    patch(outs, _MATCH)
    return start


def _closure(s):
    """BaseState.outStates / liststates: terminal states reachable through Split states, each once, in visit order."""
    seen, order, on_stack = set(), [], set()

    def add(x):
        if x is None or id(x) in seen:
            return
        if x.kind == "split":
            # the reference adds only terminal states to the set; a split is re-walked whenever it is reached again, and a
            # cycle made of splits only ((a*)* and friends) recurses forever there: StackOverflowError
            if id(x) in on_stack:
                raise ReUnsupported("StackOverflowError: epsilon cycle in the Thompson NFA")
            on_stack.add(id(x))
            add(x.out1[0])
            add(x.out2[0])
            on_stack.discard(id(x))
        else:
            seen.add(id(x))
            order.append(x)
    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))
    add(s)
    return order


def thompson_tables(post):
    start = thompson_nfa(post)
    # enumerate terminal states reachable from the start (BFS over closures), expand intervals into positions
    term, queue = [], []
    idx_of = {}

    def see(s):
        if id(s) not in idx_of and s.kind != "match":
            idx_of[id(s)] = len(term)
            term.append(s)
            queue.append(s)

    first_states = _closure(start)
    if any(s.kind == "match" for s in first_states):
        raise ReUnsupported("MatchError: StatePoint.expand has no case for MatchState (nullable regex)")
    for s in first_states:
        see(s)
    nexts = {}
    while queue:
        s = queue.pop(0)
        cl = _closure(s.out[0])
        nexts[id(s)] = cl
        for x in cl:
            see(x)
    # positions: Const -> 1, Interval -> one per char in [start, end)
    pos_of, c, flags = {}, [], []
    for s in term:
        chars = [s.c] if s.kind == "const" else list(range(s.start, s.end))
        pos_of[id(s)] = list(range(len(c), len(c) + len(chars)))
        emits = 1 if any(x.kind == "match" for x in nexts[id(s)]) else 0
        for ch in chars:
            c.append(ch)
            flags.append(emits)
    follows = []
    for s in term:
        f = [p for x in nexts[id(s)] if x.kind != "match" for p in pos_of[id(x)]]
        for _ in pos_of[id(s)]:
            follows.append(list(f))
    firsts = [p for s in first_states for p in pos_of[id(s)]]
    return {"c": c, "last": flags, "num": [0] * len(c), "follows": follows, "firsts": firsts}


def compile_thompson(s, line_only=False):
    return thompson_tables(re2post(s, line_only))

"""The automata of the reference's DFATests (src/test/scala/org/fmindex/tests/dfa.scala:13-60), built with any module that offers
StartState / State / FinishState (oracle.dfa or findex_b200.dfa)."""


def ab_star_c(m):
    s, a, b, f = m.StartState(), m.State("a"), m.State("b"), m.FinishState()
    s.link(a, "a")
    a.link(b, "b")
    b.link(b, "b")
    b.link(f, "c")
    return s


def class_b_star_c(m, chars):
    s, a, b, f = m.StartState(), m.State("a"), m.State("b"), m.FinishState()
    for c in chars:
        s.link(a, c)
    a.link(b, "b")
    b.link(b, "b")
    b.link(f, "c")
    return s


CDFKLM = ["c", "d", "f", "m", "k", "l"]
DFA1 = ["c", "d", "f", "m", "l", 0xfa, 0xfb, 0xfc, 0xfd, 0xfe, 0xff]

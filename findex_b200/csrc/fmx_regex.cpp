// fmx_regex.cpp — host-side regex front-end of libfmgpu: regex text -> Glushkov position automaton,
// flattened into the tables the device frontier kernel walks.
//
// Behavioural contract = the reference's REParser.re2post (src/main/scala/org/fmindex/re2/re2.scala:50-185)
// followed by ReTree.apply (re2/retree.scala:156-370) with node semantics :10-155, postProcess :439-482,
// removeBorderNulls :371-385, setNums :393-423.  The reference engine is a partial prototype; what it
// rejects (MatchError), mis-handles ('+' inside alternation, Q4; follows that do not climb, Q2)
// or trims (border nullables, Q5; exclusive interval ends, Q1) is reproduced so that match sets are
// identical.  Written from the behavioural description, as an index-based arena instead of the
// reference's linked object graph.
#include "fmx_internal.h"

#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>

namespace fmx {

namespace {

enum TokKind : uint8_t { T_CHAR, T_RANGE, T_SET, T_CAT, T_OR, T_STAR, T_PLUS, T_QUEST };
struct Tok {
    TokKind kind;
    int a = 0, b = 0;               // T_CHAR: a ; T_RANGE: [a,b) as the tree builder reads it
    std::vector<int> set;           // T_SET: members, front = most recently added (reference list order)
};

struct ParseError { int code; std::string msg; };
constexpr size_t kMaxNesting = 1000;       // the tree functions recurse over the nesting depth
constexpr int64_t kMaxRegexBytes = 1 << 16;
[[noreturn]] void syntax() { throw ParseError{FMX_E_SYNTAX, "re2post syntax"}; }
[[noreturn]] void unsupported(const std::string &m) { throw ParseError{FMX_E_UNSUPPORTED, m}; }

// ---- infix -> postfix with explicit concatenation (re2.scala:50-185) ---------------------------------
std::vector<Tok> to_postfix(const uint8_t *s, int64_t l, bool line_only) {
    if (l > kMaxRegexBytes) throw ParseError{FMX_E_LIMIT, "regex longer than 65536 bytes"};
    std::vector<Tok> out;
    struct Frame { int nalt, natom; };
    std::vector<Frame> frames;
    int natom = 0, nalt = 0;
    bool quoted = false;

    auto emit = [&](TokKind k) { Tok t; t.kind = k; out.push_back(std::move(t)); };
    auto before_atom = [&]() { if (natom > 1) { --natom; emit(T_CAT); } };
    auto flush_cats = [&]() { --natom; while (natom > 0) { emit(T_CAT); --natom; } };
    auto flush_ors = [&]() { while (nalt > 0) { emit(T_OR); --nalt; } };

    auto atom_char = [&](int c, bool q) {
        before_atom();
        Tok t;
        if (q && c == 'w')      { t.kind = T_RANGE; t.a = 'A'; t.b = 'z'; }
        else if (q && c == 'd') { t.kind = T_RANGE; t.a = '0'; t.b = '9'; }
        else if (!q && c == '.'){ t.kind = T_RANGE; t.a = line_only ? 0x20 : 2; t.b = 255; }
        else                    { t.kind = T_CHAR;  t.a = c; }
        out.push_back(std::move(t));
        ++natom;
    };

    // "[...]" : returns index just past ']'   (re2.scala:76-119)
    auto atom_set = [&](int64_t i) -> int64_t {
        std::vector<int> members;            // front = head
        bool q = false, closed = false, range = false;
        auto add = [&](int c) {
            if (range) {
                if (members.empty()) syntax();
                int from = members.front() + 1;
                if (from > c) syntax();
                for (int x = from; x <= c; ++x) members.insert(members.begin(), x);
                range = false;
            } else members.insert(members.begin(), c);
        };
        while (i < l && !closed) {
            int c = s[i];
            if (q) { add(c); q = false; }
            else if (c == '\\') q = true;
            else if (c == '-') range = true;
            else if (c == ']') closed = true;
            else add(c);
            ++i;
        }
        if (!closed || range) syntax();
        before_atom();
        Tok t; t.kind = T_SET; t.set = std::move(members);
        out.push_back(std::move(t));
        ++natom;
        return i;
    };

    for (int64_t i = 0; i < l; ++i) {
        int c = s[i];
        if (quoted) { atom_char(c, true); quoted = false; continue; }
        switch (c) {
        case '(':
            before_atom();
            if (frames.size() >= kMaxNesting) throw ParseError{FMX_E_LIMIT, "regex nesting deeper than 1000 (the reference overflows its stack long before)"};
            frames.push_back({nalt, natom});
            nalt = 0; natom = 0;
            break;
        case '|':
            if (natom == 0) syntax();
            flush_cats();
            ++nalt;
            break;
        case ')':
            if (natom == 0) syntax();
            flush_cats();
            flush_ors();
            if (frames.empty()) unsupported("NoSuchElementException: unbalanced ')'");
            nalt = frames.back().nalt; natom = frames.back().natom + 1;
            frames.pop_back();
            break;
        case '[':
            i = atom_set(i + 1) - 1;
            break;
        case '\\':
            quoted = true;
            break;
        case '*': case '+': case '?':
            if (natom == 0) syntax();
            emit(c == '*' ? T_STAR : c == '+' ? T_PLUS : T_QUEST);
            break;
        default:
            atom_char(c, false);
        }
    }
    if (!frames.empty()) syntax();
    flush_cats();
    flush_ors();
    return out;
}

// ---- tree arena --------------------------------------------------------------------------------------
enum NodeKind : uint8_t { N_CHAR, N_OR, N_SEQ, N_STAR, N_QUEST, N_PLUS };
struct Node {
    NodeKind kind;
    int c = 0, num = 0;
    int parent = -1;                 // -1 = RootNode
    int slot = 0;                    // index inside the parent's child list (valid after set_parents)
    std::vector<int> kids;           // front = head of the reference's child list
};

struct Tree {
    std::vector<Node> nodes;
    int make(NodeKind k, int c = 0) { Node n; n.kind = k; n.c = c; nodes.push_back(std::move(n)); return (int)nodes.size() - 1; }
    Node &at(int i) { return nodes[i]; }
    const Node &at(int i) const { return nodes[i]; }

    static bool unary(NodeKind k) { return k == N_STAR || k == N_QUEST || k == N_PLUS; }
    // shape class used by the builder's case tables: 0=Char 1=Or 2=Seq 3=Unary
    int shape(int i) const { NodeKind k = at(i).kind; return k == N_CHAR ? 0 : k == N_OR ? 1 : k == N_SEQ ? 2 : 3; }

    // append == prepend to the child list; an Or absorbing an Or splices its children in front
    void adopt(int p, int ch) {
        if (at(p).kind == N_OR && at(ch).kind == N_OR) {
            std::vector<int> merged = at(ch).kids;
            for (int k : merged) at(k).parent = p;
            merged.insert(merged.end(), at(p).kids.begin(), at(p).kids.end());
            at(p).kids = std::move(merged);
        } else {
            at(ch).parent = p;
            at(p).kids.insert(at(p).kids.begin(), ch);
        }
    }

    bool nullable(int i) const {
        const Node &n = at(i);
        switch (n.kind) {
        case N_CHAR: return false;
        case N_STAR: case N_QUEST: return true;
        case N_OR:   return std::any_of(n.kids.begin(), n.kids.end(), [&](int k) { return nullable(k); });
        default:     return std::all_of(n.kids.begin(), n.kids.end(), [&](int k) { return nullable(k); });  // Seq, Plus
        }
    }

    static void prepend(std::vector<int> &dst, const std::vector<int> &src) { dst.insert(dst.begin(), src.begin(), src.end()); }

    std::vector<int> firsts(int i) const {
        const Node &n = at(i);
        std::vector<int> r;
        if (n.kind == N_CHAR) { r.push_back(i); return r; }
        if (n.kind != N_SEQ) {
            for (int k : n.kids) { auto f = firsts(k); r.insert(r.end(), f.begin(), f.end()); }
            return r;
        }
        size_t p = 0;
        while (p < n.kids.size() && nullable(n.kids[p])) { prepend(r, firsts(n.kids[p])); ++p; }
        if (p < n.kids.size()) prepend(r, firsts(n.kids[p]));
        return r;
    }

    size_t index_in_parent(int i) const { return (size_t)at(i).slot; }

    std::vector<int> follows(int i) const {
        int pi = at(i).parent;
        if (pi < 0) return {};
        const Node &p = at(pi);
        switch (p.kind) {
        case N_OR: case N_QUEST: return follows(pi);
        case N_STAR: { auto r = firsts(i); auto f = follows(pi); r.insert(r.end(), f.begin(), f.end()); return r; }
        case N_SEQ: {
            size_t k = index_in_parent(i) + 1;
            if (k >= p.kids.size()) return follows(pi);
            std::vector<int> r = firsts(p.kids[k]);
            if (nullable(p.kids[k])) {
                ++k;
                while (k < p.kids.size() && nullable(p.kids[k])) { prepend(r, firsts(p.kids[k])); ++k; }
                if (k < p.kids.size()) prepend(r, firsts(p.kids[k]));     // siblings exhausted: no climb (Q2)
            }
            return r;
        }
        default: return {};          // under a Plus
        }
    }

    bool is_last(int i) const {
        int pi = at(i).parent;
        if (pi < 0) return true;
        const Node &p = at(pi);
        if (p.kind == N_OR || unary(p.kind)) return is_last(pi);
        if (p.kind == N_SEQ) {
            for (size_t k = index_in_parent(i) + 1; k < p.kids.size(); ++k)
                if (!nullable(p.kids[k])) return false;
            return is_last(pi);
        }
        return true;
    }
};

// case tables of the builder: [shape(a1)][shape(a2)] ; 0 = MatchError, 1 = fold into existing, 2 = new node
//                      a2:  C  O  S  U
const uint8_t OR_RULE[4][4] = {
    /* a1=C */ {2, 1, 2, 0},
    /* a1=O */ {0, 1, 0, 0},
    /* a1=S */ {2, 1, 2, 0},
    /* a1=U */ {2, 1, 2, 2},
};
const uint8_t CAT_RULE[4][4] = {
    /* a1=C */ {2, 2, 0, 2},
    /* a1=O */ {0, 2, 0, 0},
    /* a1=S */ {1, 1, 0, 1},
    /* a1=U */ {2, 2, 0, 2},
};

struct Builder {
    Tree t;
    std::vector<int> st;
    int pop() { if (st.empty()) unsupported("NoSuchElementException: empty stack"); int x = st.back(); st.pop_back(); return x; }

    int char_class(const std::vector<int> &members) {
        int o = t.make(N_OR);
        for (int c : members) t.adopt(o, t.make(N_CHAR, c));
        return o;
    }

    int wrap(NodeKind k, int child) { int u = t.make(k); t.adopt(u, child); return u; }

    void run(const std::vector<Tok> &post) {
        for (const Tok &tk : post) {
            switch (tk.kind) {
            case T_CHAR: st.push_back(t.make(N_CHAR, tk.a)); break;
            case T_RANGE: { std::vector<int> m; for (int j = tk.a; j < tk.b; ++j) m.push_back(j); st.push_back(char_class(m)); break; }   // end exclusive (Q1)
            case T_SET: st.push_back(char_class(tk.set)); break;
            case T_OR: {
                int a2 = pop(), a1 = pop();
                uint8_t r = OR_RULE[t.shape(a1)][t.shape(a2)];
                if (r == 1) { t.adopt(a2, a1); st.push_back(a2); }
                else if (r == 2) { int o = t.make(N_OR); t.adopt(o, a1); t.adopt(o, a2); st.push_back(o); }
                else unsupported("MatchError: OrPoint have no match");
                break;
            }
            case T_CAT: {
                int a2 = pop(), a1 = pop();
                uint8_t r = CAT_RULE[t.shape(a1)][t.shape(a2)];
                if (r == 1) { t.adopt(a1, a2); st.push_back(a1); }
                else if (r == 2) { int f = t.make(N_SEQ); t.adopt(f, a1); t.adopt(f, a2); st.push_back(f); }
                else unsupported("MatchError: ConcatPoint have no match");
                break;
            }
            case T_STAR: case T_PLUS: {
                int a = pop();
                NodeKind k = t.at(a).kind;
                if (k == N_STAR) st.push_back(a);
                else if (k == N_QUEST || k == N_PLUS) st.push_back(wrap(N_STAR, t.at(a).kids.front()));
                else st.push_back(wrap(tk.kind == T_PLUS ? N_PLUS : N_STAR, a));
                break;
            }
            case T_QUEST: {
                int a = pop();
                NodeKind k = t.at(a).kind;
                if (k == N_STAR) st.push_back(a);
                else if (k == N_QUEST) st.push_back(wrap(N_QUEST, t.at(a).kids.front()));
                else if (k == N_PLUS) st.push_back(wrap(N_STAR, t.at(a).kids.front()));
                else st.push_back(wrap(N_QUEST, a));
                break;
            }
            }
        }
    }
};

// deep copy into `dst` restoring forward child order; every Plus child p becomes  p , Star(p)   (Q4)
int normalise(const Tree &src, int i, Tree &dst) {
    const Node &n = src.at(i);
    if (n.kind == N_CHAR) return dst.make(N_CHAR, n.c);
    if (n.kind == N_PLUS) unsupported("MatchError in postProcess");
    int me = dst.make(n.kind);
    std::vector<int> kids;
    for (int k : n.kids) {                         // iterate the (reversed) list, prepending => forward order
        if (src.at(k).kind == N_PLUS) {
            int inner = src.at(k).kids.front();
            int once = normalise(src, inner, dst);
            int star = dst.make(N_STAR);
            int again = normalise(src, inner, dst);
            dst.at(star).kids.push_back(again);
            kids.insert(kids.begin(), {once, star});
        } else {
            kids.insert(kids.begin(), normalise(src, k, dst));
        }
    }
    dst.at(me).kids = std::move(kids);
    return me;
}

void set_parents(Tree &t, int i, int parent) {
    t.at(i).parent = parent;
    const std::vector<int> &kids = t.at(i).kids;
    for (size_t k = 0; k < kids.size(); ++k) { t.at(kids[k]).slot = (int)k; set_parents(t, kids[k], i); }
}

// priority numbers (only relevant to the reference's capped PQ order; exported for parity checks)
int number_from(Tree &t, int i, int start);
int number_shared(Tree &t, int i, int &idx) {
    Node &n = t.at(i);
    if (n.kind == N_OR) {
        int hi = idx;
        for (int k : n.kids) {
            if (t.at(k).kind == N_CHAR) { t.at(k).num = idx; hi = std::max(hi, idx + 1); }
            else hi = std::max(hi, number_from(t, k, idx));
        }
        idx = hi;
    } else {
        for (int k : n.kids) {
            if (t.at(k).kind == N_CHAR) { t.at(k).num = idx; ++idx; }
            else number_shared(t, k, idx);
        }
    }
    return idx;
}
int number_from(Tree &t, int i, int start) { int idx = start; return number_shared(t, i, idx); }

}  // namespace

// ---- Thompson engine: REParser.createNFA (re2.scala:264-334) + the traversal contract of REParser.matchSA (:568-693) ----
// A state point of the reference carries a list of intervals that are expanded independently, so the result multiset equals an
// item-wise traversal over positions: one per ConstState, one per char of an IntervalState (`start until end`, exclusive).
// follows(p) = the non-match terminal states of the epsilon closure of p's successor; p emits when that closure holds the
// MatchState — and, unlike a Glushkov last position, is still expanded.  Throws where the reference does: AltPoint has no case
// in createNFA, a nullable regex puts MatchState into the start front (no case in StatePoint.expand), an epsilon cycle of split
// states recurses forever in outStates, an operator without operand pops an empty stack.
namespace {
struct TNode {
    int kind;                 // 0 const, 1 interval, 2 split, 3 match
    int c = 0, lo = 0, hi = 0;
    int out = -1, out1 = -1, out2 = -1;
};
struct TLink { int node; int which; };          // which: 0 out, 1 out1, 2 out2
struct TFrag { int start; std::vector<TLink> outs; };

struct TBuilder {
    std::vector<TNode> n;
    int match;
    TBuilder() { n.push_back(TNode{3}); match = 0; }
    int make(int kind) { n.push_back(TNode{kind}); return (int)n.size() - 1; }
    void patch(const std::vector<TLink> &outs, int target) {
        for (const TLink &l : outs) { TNode &x = n[l.node]; (l.which == 0 ? x.out : l.which == 1 ? x.out1 : x.out2) = target; }
    }
    void closure_rec(int s, std::vector<char> &seen, std::vector<char> &onstack, std::vector<int> &order) const {
        if (s < 0 || seen[s]) return;
        if (n[s].kind == 2) {
            if (onstack[s]) unsupported("StackOverflowError: epsilon cycle in the Thompson NFA");
            onstack[s] = 1;
            closure_rec(n[s].out1, seen, onstack, order);
            closure_rec(n[s].out2, seen, onstack, order);
            onstack[s] = 0;
        } else { seen[s] = 1; order.push_back(s); }
    }
    std::vector<int> closure(int s) const {
        std::vector<char> seen(n.size(), 0), onstack(n.size(), 0);
        std::vector<int> order;
        closure_rec(s, seen, onstack, order);
        return order;
    }
};
}  // namespace

int compile_thompson(const uint8_t *re, int64_t len, bool line_only, CompiledRegex &out, std::string &err) {
    try {
        std::vector<Tok> post = to_postfix(re, len, line_only);
        TBuilder b;
        std::vector<TFrag> st;
        auto pop = [&]() { if (st.empty()) unsupported("NoSuchElementException: empty stack"); TFrag f = std::move(st.back()); st.pop_back(); return f; };
        for (const Tok &tk : post) {
            switch (tk.kind) {
            case T_QUEST: { TFrag e = pop(); int ns = b.make(2); b.n[ns].out1 = e.start; std::vector<TLink> o{{ns, 2}}; o.insert(o.end(), e.outs.begin(), e.outs.end()); st.push_back({ns, o}); break; }
            case T_STAR:  { TFrag e = pop(); int ns = b.make(2); b.n[ns].out1 = e.start; b.patch(e.outs, ns); st.push_back({ns, {{ns, 2}}}); break; }
            case T_PLUS:  { TFrag e = pop(); int ns = b.make(2); b.n[ns].out1 = e.start; b.patch(e.outs, ns); st.push_back({e.start, {{ns, 2}}}); break; }
            case T_CAT:   { TFrag e2 = pop(), e1 = pop(); b.patch(e1.outs, e2.start); st.push_back({e1.start, e2.outs}); break; }
            case T_OR:    { TFrag e2 = pop(), e1 = pop(); int ns = b.make(2); b.n[ns].out1 = e1.start; b.n[ns].out2 = e2.start;
                            std::vector<TLink> o = e1.outs; o.insert(o.end(), e2.outs.begin(), e2.outs.end()); st.push_back({ns, o}); break; }
            case T_CHAR:  { int ns = b.make(0); b.n[ns].c = tk.a; st.push_back({ns, {{ns, 0}}}); break; }
            case T_RANGE: { int ns = b.make(1); b.n[ns].lo = tk.a; b.n[ns].hi = tk.b; st.push_back({ns, {{ns, 0}}}); break; }
            default: unsupported("MatchError: createNFA has no case for a character set");
            }
        }
        TFrag e0 = pop();
        b.patch(e0.outs, b.match);

        std::vector<int> first_states = b.closure(e0.start);
        for (int s : first_states) if (s == b.match) unsupported("MatchError: StatePoint.expand has no case for MatchState (nullable regex)");
        std::vector<int> term, idx_of(b.n.size(), -1);
        std::vector<std::vector<int>> nexts;
        auto see = [&](int s) { if (s != b.match && idx_of[s] < 0) { idx_of[s] = (int)term.size(); term.push_back(s); } };
        for (int s : first_states) see(s);
        for (size_t q = 0; q < term.size(); ++q) {
            nexts.push_back(b.closure(b.n[term[q]].out));
            for (int x : nexts.back()) see(x);
        }
        std::vector<std::vector<int>> pos_of(term.size());
        out = CompiledRegex();
        out.stop_on_emit = false;
        for (size_t q = 0; q < term.size(); ++q) {
            const TNode &t = b.n[term[q]];
            const bool emits = std::find(nexts[q].begin(), nexts[q].end(), b.match) != nexts[q].end();
            const int lo = t.kind == 0 ? t.c : t.lo, hi = t.kind == 0 ? t.c + 1 : t.hi;
            for (int ch = lo; ch < hi; ++ch) {
                pos_of[q].push_back((int)out.c.size());
                out.c.push_back((uint8_t)ch);
                out.is_last.push_back(emits ? 1 : 0);
                out.num.push_back(0);
            }
        }
        out.follows_off.push_back(0);
        for (size_t q = 0; q < term.size(); ++q) {
            std::vector<int32_t> f;
            for (int x : nexts[q]) if (x != b.match) for (int p : pos_of[idx_of[x]]) f.push_back(p);
            for (size_t k = 0; k < pos_of[q].size(); ++k) {
                out.follows.insert(out.follows.end(), f.begin(), f.end());
                out.follows_off.push_back((int32_t)out.follows.size());
            }
        }
        for (int s : first_states) for (int p : pos_of[idx_of[s]]) out.firsts.push_back(p);
        return FMX_OK;
    } catch (const ParseError &e) {
        err = e.msg;
        return e.code;
    }
}

int compile_regex(const uint8_t *re, int64_t len, bool line_only, CompiledRegex &out, std::string &err) {
    try {
        std::vector<Tok> post = to_postfix(re, len, line_only);
        Builder b;
        b.run(post);
        int top = b.pop();
        if (b.t.at(top).kind != N_SEQ) top = b.wrap(N_SEQ, top);

        Tree t;
        int root0 = normalise(b.t, top, t);
        // trim nullable children at both borders of the root sequence (Q5)
        std::vector<int> &rk = t.at(root0).kids;
        size_t lo = 0, hi = rk.size();
        while (lo < hi && t.nullable(rk[lo])) ++lo;
        while (hi > lo && t.nullable(rk[hi - 1])) --hi;
        int root = t.make(N_SEQ);
        t.at(root).kids.assign(t.at(root0).kids.begin() + lo, t.at(root0).kids.begin() + hi);
        set_parents(t, root, -1);
        number_from(t, root, 1);

        // flatten: positions in pre-order
        std::vector<int> order, sid(t.nodes.size(), -1);
        std::vector<int> stack{root};
        while (!stack.empty()) {
            int i = stack.back(); stack.pop_back();
            if (t.at(i).kind == N_CHAR) { sid[i] = (int)order.size(); order.push_back(i); }
            const auto &k = t.at(i).kids;
            for (auto it = k.rbegin(); it != k.rend(); ++it) stack.push_back(*it);
        }
        out = CompiledRegex();
        out.follows_off.push_back(0);
        for (int i : order) {
            out.c.push_back((uint8_t)t.at(i).c);
            out.num.push_back(t.at(i).num);
            out.is_last.push_back(t.is_last(i) ? 1 : 0);
            for (int f : t.follows(i)) out.follows.push_back(sid[f]);
            out.follows_off.push_back((int32_t)out.follows.size());
        }
        for (int f : t.firsts(root)) out.firsts.push_back(sid[f]);
        return FMX_OK;
    } catch (const ParseError &e) {
        err = e.msg;
        return e.code;
    }
}

// =====================================================================================================
// DFA engine  (src/main/scala/org/fmindex/dfa.scala)
// =====================================================================================================
// DFA.processLinkList (:391-407): states reachable from the start state, numbered in the order the reference's `visited` set grows
// (start = 0), moves[state][char] from the links, then compileBuckets (:198-223).  For the search the automaton is handed to the
// frontier kernel as a position automaton over its DFAChar EDGES: edge (s --c--> t) consumes c, emits iff t is a finish state and is
// followed by t's DFAChar edges — StatePoint.expand (:238-256) follows DFAChar actions only, a DFABucket (two or more consecutive
// characters with one target) is never traversed, and a finish state emits and is expanded all the same (:272-274).
int compile_dfa(int32_t n_states, const uint8_t *kind, const int32_t *link_off, const int32_t *link_to, const int32_t *link_chr,
                CompiledDfa &dfa, CompiledRegex &out, std::string &err) {
    if (n_states <= 0 || !kind || !link_off) { err = "bad DFA description"; return FMX_E_ARG; }
    int start = -1;
    for (int i = 0; i < n_states; ++i) {
        if (kind[i] > 2) { err = "state kind must be 0 (start), 1 (state) or 2 (finish)"; return FMX_E_ARG; }
        if (kind[i] == 0) { if (start >= 0) { err = "Start state already taken"; return FMX_E_ARG; } start = i; }
    }
    if (start < 0) { err = "no start state"; return FMX_E_ARG; }
    for (int i = 0; i < n_states; ++i)
        for (int k = link_off[i]; k < link_off[i + 1]; ++k)
            if (link_to[k] < 0 || link_to[k] >= n_states || link_chr[k] < 0 || link_chr[k] > 255) { err = "link out of range"; return FMX_E_ARG; }

    // _processLinkList(s, visited): v = visited + s; for (l <- s.links if !visited(l.to)) v = v ++ rec(l.to, v)   — the membership test
    // looks at the set that was passed in.  `order` is an insertion-ordered set.
    struct Walk {
        int32_t n; const int32_t *off, *to;
        std::vector<int> rec(int s, const std::vector<int> &visited) const {
            std::vector<int> v = visited;
            std::vector<char> in_visited((size_t)n, 0), in_v((size_t)n, 0);
            for (int x : visited) { in_visited[(size_t)x] = 1; in_v[(size_t)x] = 1; }
            if (!in_v[(size_t)s]) { v.push_back(s); in_v[(size_t)s] = 1; }
            for (int k = off[s]; k < off[s + 1]; ++k) {
                if (in_visited[(size_t)to[k]]) continue;
                for (int x : rec(to[k], v)) if (!in_v[(size_t)x]) { v.push_back(x); in_v[(size_t)x] = 1; }
            }
            return v;
        }
    } walk{n_states, link_off, link_to};
    const std::vector<int> order = walk.rec(start, {});

    dfa = CompiledDfa();
    dfa.number.assign((size_t)n_states, -1);
    int idx = 1;
    for (int s : order) dfa.number[(size_t)s] = (kind[s] == 0) ? 0 : idx++;
    const int ns = (int)order.size();
    dfa.n_states = ns;
    dfa.moves.assign((size_t)ns * 256, -1);
    dfa.finish.assign((size_t)ns, 0);
    for (int s : order) {
        if (kind[s] == 2) dfa.finish[(size_t)dfa.number[(size_t)s]] = 1;
        for (int k = link_off[s]; k < link_off[s + 1]; ++k)                       // addLink: a later link for the same char overwrites
            dfa.moves[(size_t)dfa.number[(size_t)s] * 256 + (size_t)link_chr[k]] = dfa.number[(size_t)link_to[k]];
    }
    // the reference applies links in `for (v <- visited; l <- v.links)` order, i.e. list order (most recently added first), so the
    // OLDEST link for a character wins; callers pass links in list order and the loop above keeps the last one it sees — the same.

    // compileBuckets
    dfa.bucket_off.assign(1, 0);
    for (int i = 0; i < ns; ++i) {
        int last = -1, start_bucket = -1;
        for (int j = 0; j < 256; ++j) {
            const int v = dfa.moves[(size_t)i * 256 + (size_t)j];
            if (last != v) {
                if (last != -1) dfa.buckets.push_back({last, start_bucket, j - 1});
                start_bucket = j;
                last = v;
            }
        }
        if (last != -1) dfa.buckets.push_back({last, start_bucket, 255});
        dfa.bucket_off.push_back((int32_t)dfa.buckets.size());
    }

    // position automaton over the DFAChar edges
    std::vector<int32_t> first_edge((size_t)ns + 1, 0);
    std::vector<int32_t> edge_target;
    out = CompiledRegex();
    out.stop_on_emit = false;
    for (int i = 0; i < ns; ++i) {
        first_edge[(size_t)i] = (int32_t)out.c.size();
        for (int b = dfa.bucket_off[(size_t)i]; b < dfa.bucket_off[(size_t)i + 1]; ++b) {
            const DfaAction &a = dfa.buckets[(size_t)b];
            if (a.c1 != a.c2) continue;                                        // DFABucket: never traversed
            out.c.push_back((uint8_t)a.c1);
            out.num.push_back(0);
            out.is_last.push_back(dfa.finish[(size_t)a.state]);
            edge_target.push_back(a.state);
        }
    }
    first_edge[(size_t)ns] = (int32_t)out.c.size();
    out.follows_off.push_back(0);
    for (size_t e = 0; e < edge_target.size(); ++e) {
        const int t = edge_target[e];
        for (int32_t f = first_edge[(size_t)t]; f < first_edge[(size_t)t + 1]; ++f) out.follows.push_back(f);
        out.follows_off.push_back((int32_t)out.follows.size());
    }
    for (int32_t f = first_edge[0]; f < first_edge[1]; ++f) out.firsts.push_back(f);
    return FMX_OK;
}

// DFA.fromNFA(initialState) (dfa.scala:343-389): subset construction over an NFA of NfaBaseState objects with NfaLink(to, chr) and
// EpsilonLink(to) links (:5-37), then processLinkList.  A DFA state is a set of NFA states closed under epsilon links; the set of the
// initial state becomes the StartState — also when it contains an NfaFinishState, and also when it is reached again later (:353-359:
// `if (x == initialStateSet) startState`), so it never accepts; any other set with an NfaFinishState is a FinishState.  The reference
// adds the DFA links in the iteration order of a hash map, which decides the numbering of automata with more than four states and
// nothing else; here sets are discovered breadth-first with characters ascending, and links are added in (source, character) order.
int dfa_from_nfa(int32_t n_states, const uint8_t *is_finish, int32_t initial, const int32_t *link_off, const int32_t *link_to,
                 const int32_t *link_chr, std::vector<uint8_t> &kind, std::vector<int32_t> &d_off, std::vector<int32_t> &d_to,
                 std::vector<int32_t> &d_chr, std::string &err) {
    if (n_states <= 0 || !is_finish || !link_off || initial < 0 || initial >= n_states) { err = "bad NFA description"; return FMX_E_ARG; }
    for (int i = 0; i < n_states; ++i)
        for (int k = link_off[i]; k < link_off[i + 1]; ++k)
            if (link_to[k] < 0 || link_to[k] >= n_states || link_chr[k] < -1 || link_chr[k] > 255) { err = "link out of range"; return FMX_E_ARG; }
    typedef std::vector<int32_t> Set;                                   // sorted, unique
    auto closure = [&](const Set &in) {                                 // NFA.epsilons(set): union of s.epsilons
        std::vector<char> seen((size_t)n_states, 0);
        std::vector<int32_t> stack(in.begin(), in.end());
        for (int32_t s : in) seen[(size_t)s] = 1;
        while (!stack.empty()) {
            const int32_t s = stack.back(); stack.pop_back();
            for (int k = link_off[s]; k < link_off[s + 1]; ++k)
                if (link_chr[k] < 0 && !seen[(size_t)link_to[k]]) { seen[(size_t)link_to[k]] = 1; stack.push_back(link_to[k]); }
        }
        Set out;
        for (int32_t s = 0; s < n_states; ++s) if (seen[(size_t)s]) out.push_back(s);
        return out;
    };
    std::vector<Set> sets;                                              // DFA states in discovery order; sets[0] = closure({initial})
    std::vector<std::vector<std::pair<int, int>>> trans;               // per DFA state: (chr, target), chr ascending
    auto find_or_add = [&](const Set &x) {
        for (size_t i = 0; i < sets.size(); ++i) if (sets[i] == x) return (int)i;
        sets.push_back(x); trans.emplace_back();
        return (int)sets.size() - 1;
    };
    find_or_add(closure(Set{initial}));
    for (size_t cur = 0; cur < sets.size(); ++cur) {                    // the queue of psc: every set is expanded once
        if (sets.size() > 100000) { err = "subset construction exceeds 100000 DFA states"; return FMX_E_LIMIT; }
        const Set here = sets[cur];
        for (int c = 0; c < 256; ++c) {                                 // NFA.epsilonTransitions: chr -> union of target closures
            Set raw;
            for (int32_t s : here)
                for (int k = link_off[s]; k < link_off[s + 1]; ++k) if (link_chr[k] == c) raw.push_back(link_to[k]);
            if (raw.empty()) continue;
            std::sort(raw.begin(), raw.end());
            raw.erase(std::unique(raw.begin(), raw.end()), raw.end());
            const int t = find_or_add(closure(raw));
            trans[cur].push_back({c, t});
        }
    }
    const int nd = (int)sets.size();
    kind.assign((size_t)nd, 1);
    kind[0] = 0;
    for (int i = 1; i < nd; ++i)
        for (int32_t s : sets[(size_t)i]) if (is_finish[s]) { kind[(size_t)i] = 2; break; }
    // link() prepends: after adding a state's links with characters ascending, its list reads characters descending
    d_off.assign(1, 0); d_to.clear(); d_chr.clear();
    for (int i = 0; i < nd; ++i) {
        for (auto it = trans[(size_t)i].rbegin(); it != trans[(size_t)i].rend(); ++it) { d_chr.push_back(it->first); d_to.push_back(it->second); }
        d_off.push_back((int32_t)d_to.size());
    }
    return FMX_OK;
}

static std::string pretty_chr(int c) {
    char buf[8];
    if (c < 0x20 || c > 0x7e) std::snprintf(buf, sizeof buf, "\\x%x", c); else std::snprintf(buf, sizeof buf, "%c", c);
    return buf;
}

// buckets(state).mkString(",") with the reference's DFAChar / DFABucket toString (:190-196)
std::string dfa_bucket_string(const CompiledDfa &dfa, int state) {
    std::string s;
    for (int b = dfa.bucket_off[(size_t)state]; b < dfa.bucket_off[(size_t)state + 1]; ++b) {
        const DfaAction &a = dfa.buckets[(size_t)b];
        if (!s.empty()) s += ",";
        if (a.c1 == a.c2) s += "DFAChar('" + pretty_chr(a.c1) + "'->" + std::to_string(a.state) + ")";
        else s += "DFABucket('" + pretty_chr(a.c1) + "-" + pretty_chr(a.c2) + "' ->" + std::to_string(a.state) + ")";
    }
    return s;
}

}  // namespace fmx

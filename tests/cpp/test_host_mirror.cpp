// C++ host mirror test: the reference's own assertions (CombinedIndexingTest, src/test/scala/org/fmindex/tests/Indexer.scala:1079-1124;
// "match SA FMindex", REParser.scala:292-307) rewritten against include/GpuFMSearcher.hpp.  Needs a GPU.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <string>

#include "GpuFMSearcher.hpp"

#define REQUIRE(x) do { if (!(x)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #x); std::exit(1); } } while (0)

int main(int argc, char **argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s <dir with test1024.cmp.bwt/.aux>\n", argv[0]); return 2; }
    const std::string dir = argv[1];
    fmx::GpuFMSearcher sa(dir + "/test1024.cmp.bwt", /*bigEndian=*/false);
    const int64_t eof = sa.eof;
    REQUIRE(sa.n == 1025);
    REQUIRE(eof == 462);
    REQUIRE(sa.getPrevI(eof) == 0);
    REQUIRE(sa.getNextI(eof) == 517);
    REQUIRE(sa.getPrevI(1) == 48);
    REQUIRE(sa.getPrevI(48) == 649);
    REQUIRE(sa.nextSubstr(1, 3) == "haa");
    REQUIRE(sa.nextSubstr(sa.getNextI(eof), 100) == "zajrtzbeqwbxdfpwjflmmsseewuudgfbtzqenjqafwzcnfanycigwsflfvxojxpqhhzekjdkhgsptqveavquuoqujbezdkarayom");
    REQUIRE(sa.nextSubstr(eof, 100) == "ajrtzbeqwbxdfpwjflmmsseewuudgfbtzqenjqafwzcnfanycigwsflfvxojxpqhhzekjdkhgsptqveavquuoqujbezdkarayoml");
    REQUIRE(sa.prevSubstr(1, 5) == "bqxxa");
    REQUIRE(sa.prevSubstr(eof, 5) == std::string("\0uexm", 5));
    REQUIRE(sa.prevSubstr(sa.getPrevI(eof), 4) == "uexm");

    // search / getPrevRange agree: feeding "ac" in file order through getPrevRange == search("ca")
    auto r1 = sa.getPrevRange(0, sa.n, 'a');
    REQUIRE(r1.has_value());
    auto r2 = sa.getPrevRange(r1->first, r1->second, 'c');
    auto s = sa.search("ca");
    REQUIRE(r2.has_value() && s.has_value() && *r2 == *s);
    REQUIRE(s->first == 83 && s->second == 85);                       // "[2 Results] ac"
    REQUIRE(!sa.search("zzzzzzzz").has_value());

    // Glushkov engine: the alternation of REParser.scala:292-307 behind a literal the builder accepts
    std::set<std::string> got;
    for (char lead = 'a'; lead <= 'z'; ++lead) {
        fmx::ReTree t(std::string(1, lead) + "(a|b|d|e)c");
        for (const auto &m : t.matchSA(sa)) {
            REQUIRE(m.len == 3);
            std::string str = m.toString();
            got.insert(str.substr(str.size() - 2));
        }
    }
    REQUIRE(got == (std::set<std::string>{"ac", "bc", "dc", "ec"}));
    // construction is partial exactly where the reference's is (MatchError), syntax errors are exceptions
    bool threw = false;
    try { fmx::ReTree bad("(a|b|d|e)c"); } catch (const fmx::MatchError &) { threw = true; }
    REQUIRE(threw);
    threw = false;
    try { fmx::ReTree bad("a(b"); } catch (const std::runtime_error &e) { threw = std::string(e.what()).find("re2post syntax") != std::string::npos; }
    REQUIRE(threw);
    threw = false;
    try { fmx::GpuFMSearcher bad(dir + "/test1024.cmp.bwt", /*bigEndian=*/true); } catch (const std::runtime_error &e) { threw = std::string(e.what()).find("bad size") != std::string::npos; }
    REQUIRE(threw);
    // "match SA FMindex" (REParser.scala:292-307) verbatim with the Thompson engine: post2re("ba|d|e|c.")
    std::set<std::string> shown;
    for (const auto &m : fmx::ThompsonNFA("(((b|a)|d)|e)c").matchSA(sa)) shown.insert(m.toString());
    REQUIRE(shown == (std::set<std::string>{"ec", "dc", "[2 Results] ac", "bc"}));
    // DFATests (T/dfa.scala:62-105) through the C++ mirror: `ab*c`
    {
        fmx::DFABuilder b;
        const int s = b.addState(fmx::DFABuilder::Start), a = b.addState(fmx::DFABuilder::Plain), bb = b.addState(fmx::DFABuilder::Plain),
                  f = b.addState(fmx::DFABuilder::Finish);
        b.link(s, a, 'a'); b.link(a, bb, 'b'); b.link(bb, bb, 'b'); b.link(bb, f, 'c');
        auto dfa = b.build();
        REQUIRE(!dfa->matchString("absbc") && dfa->matchString("abbc") && dfa->matchString("abc"));
        REQUIRE(dfa->buckets(0) == "DFAChar('a'->1)" && dfa->buckets(1) == "DFAChar('b'->2)");
        REQUIRE(dfa->buckets(2) == "DFAChar('b'->2),DFAChar('c'->3)" && dfa->buckets(3).empty());
        for (const auto &m : dfa->matchSA(sa)) {           // every result renders as a string of the language
            const std::string w = sa.nextSubstr(m.sp, m.len);
            REQUIRE(w.size() >= 3 && w.front() == 'a' && w.back() == 'c' && w.find_first_not_of('b', 1) == w.size() - 1);
        }
        fmx::DFABuilder b2;                                 // `ac`: the index holds it twice ("[2 Results] ac" above)
        const int s2 = b2.addState(fmx::DFABuilder::Start), a2 = b2.addState(fmx::DFABuilder::Plain), f2 = b2.addState(fmx::DFABuilder::Finish);
        b2.link(s2, a2, 'a'); b2.link(a2, f2, 'c');
        const auto r2 = b2.build()->matchSA(sa);
        REQUIRE(r2.size() == 1 && r2[0].len == 2 && r2[0].cnt() == 2 && r2[0].toString() == "[2 Results] ac");
    }
    // fixed-length batch with the reference's Int rows
    {
        const std::string pats = "bqxxa" "zzzzz";          // prevSubstr(1,5) of the index (T/Indexer.scala:1119) reversed is a hit; zzzzz is not
        std::vector<int32_t> sp32, ep32;
        sa.searchFixed(reinterpret_cast<const uint8_t *>(pats.data()), 5, 2, sp32, ep32);
        auto one = sa.search("bqxxa");
        REQUIRE((one ? (sp32[0] == one->first && ep32[0] == one->second) : (sp32[0] == 0 && ep32[0] == 0)));
        REQUIRE(sp32[1] == 0 && ep32[1] == 0);
    }
    std::printf("cpp host mirror ok\n");
    return 0;
}

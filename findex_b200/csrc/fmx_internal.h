// fmx_internal.h — shared declarations inside libfmgpu (not part of the ABI).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/fmgpu.h"

namespace fmx {

// ---- error plumbing --------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int  fail(int code, const char *fmt, ...);

// ---- host-side file layer (fmx_files.cpp) ------------------------------------------------------------
struct IndexFiles {
    int64_t n = 0, eof = 0;
    std::vector<uint8_t> bwt;          // n bytes, byte at eof forced to 0
    int64_t counts[256] = {0};         // raw .aux
};
std::string strip_extension(const std::string &path);       // FilenameUtils.removeExtension semantics
int load_index_files(const std::string &base, bool big_endian, bool require_fm, IndexFiles &out);
int write_index_files(const std::string &base, const uint8_t *bwt, int64_t n, int64_t eof, const int64_t counts[256],
                      bool big_endian, const uint32_t *fm_or_null);

// ---- regex compiler (fmx_regex.cpp) -------------------------------------------------------------------
struct CompiledRegex {
    std::vector<uint8_t> c;            // per position: the byte it consumes
    std::vector<uint8_t> is_last;
    std::vector<int32_t> num;          // reference PQ priority (exported for parity checks only)
    std::vector<int32_t> follows_off;  // CSR, size n_states+1
    std::vector<int32_t> follows;
    std::vector<int32_t> firsts;       // start positions
    bool stop_on_emit = true;          // Glushkov: a last position emits and is not expanded; Thompson: it emits and goes on
};
int compile_regex(const uint8_t *re, int64_t len, bool line_only, CompiledRegex &out, std::string &err);
int compile_thompson(const uint8_t *re, int64_t len, bool line_only, CompiledRegex &out, std::string &err);


// ---- DFA engine (dfa.scala), compiled next to the regex engines in fmx_regex.cpp ----------------------------
struct DfaAction { int state, c1, c2; };       // c1 == c2: DFAChar, else DFABucket
struct CompiledDfa {
    int n_states = 0;                  // reachable states, numbered like DFA.processLinkList (start = 0)
    std::vector<int32_t> number;       // caller's state index -> DFA state (-1 = unreachable)
    std::vector<int32_t> moves;        // n_states x 256, -1 = no move
    std::vector<uint8_t> finish;       // n_states
    std::vector<DfaAction> buckets;    // compileBuckets, concatenated
    std::vector<int32_t> bucket_off;   // n_states + 1
};
// link_off[n_states+1] / link_to / link_chr: every state's links in the reference's list order (most recently added first)
int compile_dfa(int32_t n_states, const uint8_t *kind, const int32_t *link_off, const int32_t *link_to, const int32_t *link_chr,
                CompiledDfa &dfa, CompiledRegex &out, std::string &err);
std::string dfa_bucket_string(const CompiledDfa &dfa, int state);
// DFA.fromNFA: NFA (link_chr = -1 for an EpsilonLink) -> the state kinds and list-ordered links compile_dfa takes
int dfa_from_nfa(int32_t n_states, const uint8_t *is_finish, int32_t initial, const int32_t *link_off, const int32_t *link_to,
                 const int32_t *link_chr, std::vector<uint8_t> &kind, std::vector<int32_t> &d_off, std::vector<int32_t> &d_to,
                 std::vector<int32_t> &d_chr, std::string &err);

}  // namespace fmx

struct fmx_regex {
    fmx::CompiledRegex a;
    fmx::CompiledDfa dfa;              // filled by fmx_dfa_create only
};

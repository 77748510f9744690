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

}  // namespace fmx

struct fmx_regex {
    fmx::CompiledRegex a;
};

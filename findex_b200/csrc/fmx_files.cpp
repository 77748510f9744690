// fmx_files.cpp — the on-disk index layout of the reference, host side.
//
//   <base>.bwt : i64 n | i64 eof | u8 bwt[n]        BWTLoader   src/main/scala/org/fmindex/bwtmerger.scala:144-174
//   <base>.aux : i64 counts[256]                    AUXLoader   bwtmerger.scala:130-142
//   <base>.fm  : u8 elSize(=4) | i64 n | u32be[n]   FMLoader    bwtmerger.scala:252-290 (payload always big-endian)
//
// The header byte order follows the caller's flag (Scala-written files are big-endian, the bwtdisk
// goldens little-endian).  Size validation mirrors the reference's exceptions.
#include "fmx_internal.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <sys/stat.h>

namespace fmx {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
int fail(int code, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
const char *last_error() { return g_err; }

std::string strip_extension(const std::string &path) {
    size_t sep = path.find_last_of("/\\");
    size_t dot = path.find_last_of('.');
    if (dot == std::string::npos) return path;
    if (sep != std::string::npos && dot < sep) return path;
    return path.substr(0, dot);
}

static int64_t get_i64(const uint8_t *p, bool be) {
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) v = (v << 8) | p[be ? i : 7 - i];
    return (int64_t)v;
}
static void put_i64(uint8_t *p, int64_t x, bool be) {
    uint64_t v = (uint64_t)x;
    for (int i = 0; i < 8; ++i) p[be ? 7 - i : i] = (uint8_t)(v >> (8 * i));
}

static int64_t file_size(const std::string &p) {
    struct stat st;
    if (stat(p.c_str(), &st) != 0) return -1;
    return (int64_t)st.st_size;
}

int load_index_files(const std::string &base, bool be, bool require_fm, IndexFiles &out) {
    const std::string pb = base + ".bwt", pa = base + ".aux", pf = base + ".fm";

    int64_t lb = file_size(pb);
    if (lb < 0) return fail(FMX_E_IO, "File %s does not exists", pb.c_str());
    FILE *f = fopen(pb.c_str(), "rb");
    if (!f) return fail(FMX_E_IO, "cannot open %s", pb.c_str());
    uint8_t hdr[16];
    if (lb < 16 || fread(hdr, 1, 16, f) != 16) { fclose(f); return fail(FMX_E_FORMAT, "File %s bad size %lld", pb.c_str(), (long long)lb); }
    out.n = get_i64(hdr, be);
    out.eof = get_i64(hdr + 8, be);
    if (out.n + 16 != lb) {
        fclose(f);
        return fail(FMX_E_FORMAT, "File %s bad size %lld != %lld + 16 ", pb.c_str(), (long long)out.n, (long long)lb);
    }
    if (out.n < 1 || out.eof < 0 || out.eof >= out.n) { fclose(f); return fail(FMX_E_FORMAT, "File %s bad eof %lld", pb.c_str(), (long long)out.eof); }
    out.bwt.resize((size_t)out.n);
    if (fread(out.bwt.data(), 1, (size_t)out.n, f) != (size_t)out.n) { fclose(f); return fail(FMX_E_IO, "short read %s", pb.c_str()); }
    fclose(f);
    out.bwt[(size_t)out.eof] = 0;                       // BWTLoader.read(eof) == 0

    int64_t la = file_size(pa);
    if (la < 0) return fail(FMX_E_IO, "File %s does not exists", pa.c_str());
    if (la < 2048) return fail(FMX_E_FORMAT, "File %s bad size %lld", pa.c_str(), (long long)la);
    f = fopen(pa.c_str(), "rb");
    if (!f) return fail(FMX_E_IO, "cannot open %s", pa.c_str());
    uint8_t ab[2048];
    if (fread(ab, 1, 2048, f) != 2048) { fclose(f); return fail(FMX_E_IO, "short read %s", pa.c_str()); }
    fclose(f);
    int64_t tot = 0;
    for (int i = 0; i < 256; ++i) { out.counts[i] = get_i64(ab + 8 * i, be); if (i) tot += out.counts[i]; }
    if (tot + 1 != out.n) return fail(FMX_E_FORMAT, "File %s counts sum %lld != n-1 = %lld", pa.c_str(), (long long)tot, (long long)(out.n - 1));

    int64_t lf = file_size(pf);
    if (lf < 0) {
        if (require_fm) return fail(FMX_E_IO, "File %s does not exists", pf.c_str());
    } else {
        f = fopen(pf.c_str(), "rb");
        if (!f) return fail(FMX_E_IO, "cannot open %s", pf.c_str());
        uint8_t fh[9];
        size_t got = fread(fh, 1, 9, f);
        fclose(f);
        if (got != 9) return fail(FMX_E_FORMAT, "File %s bad size", pf.c_str());
        if (fh[0] != 4) return fail(FMX_E_FORMAT, "File %s bad elSize %d", pf.c_str(), (int)fh[0]);
        int64_t fn = get_i64(fh + 1, be);
        if (fn * 4 + 9 != lf) return fail(FMX_E_FORMAT, "File %s bad size %lld + 0x9 != %lld(filelen) ", pf.c_str(), (long long)fn, (long long)lf);
        if (fn != out.n) return fail(FMX_E_FORMAT, "File %s size %lld does not match bwt size %lld", pf.c_str(), (long long)fn, (long long)out.n);
    }
    return FMX_OK;
}

int write_index_files(const std::string &base, const uint8_t *bwt, int64_t n, int64_t eof, const int64_t counts[256],
                      bool be, const uint32_t *fm) {
    {
        FILE *f = fopen((base + ".bwt").c_str(), "wb");
        if (!f) return fail(FMX_E_IO, "cannot create %s.bwt", base.c_str());
        uint8_t hdr[16];
        put_i64(hdr, n, be); put_i64(hdr + 8, eof, be);
        bool ok = fwrite(hdr, 1, 16, f) == 16 && fwrite(bwt, 1, (size_t)n, f) == (size_t)n;
        fclose(f);
        if (!ok) return fail(FMX_E_IO, "short write %s.bwt", base.c_str());
    }
    {
        FILE *f = fopen((base + ".aux").c_str(), "wb");
        if (!f) return fail(FMX_E_IO, "cannot create %s.aux", base.c_str());
        uint8_t ab[2048];
        for (int i = 0; i < 256; ++i) put_i64(ab + 8 * i, counts[i], be);
        bool ok = fwrite(ab, 1, 2048, f) == 2048;
        fclose(f);
        if (!ok) return fail(FMX_E_IO, "short write %s.aux", base.c_str());
    }
    if (fm) {
        FILE *f = fopen((base + ".fm").c_str(), "wb");
        if (!f) return fail(FMX_E_IO, "cannot create %s.fm", base.c_str());
        uint8_t fh[9];
        fh[0] = 4; put_i64(fh + 1, n, true);           // RandomAccessFile.writeLong: always big-endian
        bool ok = fwrite(fh, 1, 9, f) == 9;
        std::vector<uint8_t> buf(1 << 20);
        for (int64_t i = 0; ok && i < n;) {
            size_t k = 0;
            for (; k + 4 <= buf.size() && i < n; ++i, k += 4) {
                uint32_t v = fm[i];
                buf[k] = (uint8_t)(v >> 24); buf[k + 1] = (uint8_t)(v >> 16); buf[k + 2] = (uint8_t)(v >> 8); buf[k + 3] = (uint8_t)v;
            }
            ok = fwrite(buf.data(), 1, k, f) == k;
        }
        fclose(f);
        if (!ok) return fail(FMX_E_IO, "short write %s.fm", base.c_str());
    }
    return FMX_OK;
}

}  // namespace fmx

extern "C" const char *fmx_last_error(void) { return fmx::last_error(); }

/*
 * oracle/fm_oracle.c — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT.
 *
 * CPU restatement (plain C) of the reference's FM-index hot path, used only as the checker in
 * tests/, in __graft_entry__.smoke() and in bench.py's cpu_baseline / --impl reference legs.
 * The product (findex_b200/csrc, libfmgpu.so) never links, imports or calls anything in here.
 *
 * Parity status: PINNED. The reference is Scala and no JVM exists in this image, so it cannot be
 * executed; this restatement is checked against every known-answer test the reference's own test
 * suite holds for the path (tests/test_oracle_golden.py, vectors G1..G12 of SURVEY.md §8c).
 *
 * Reference citations are relative to /root/reference/src/main/scala/org/fmindex/ ("M/").
 * Indices are widened to 64 bit (the reference's Int arithmetic fails for n >= 2^29);
 * all formulas are otherwise literal.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define ALPHA 256

typedef struct fmo_index {
    int64_t   n;            /* fm.size = text length + 1            (M/bwtmerger.scala:339) */
    int64_t   eof;          /* row whose BWT char is '$'            (M/bwtmerger.scala:151) */
    uint8_t  *bwt;          /* n bytes, bwt[eof] reads as 0         (M/bwtmerger.scala:155-162) */
    int64_t   cnt[ALPHA];   /* raw .aux counts                      (M/bwtmerger.scala:130-139) */
    int64_t   bs[ALPHA+1];  /* bucketStarts  (cnt[0] forced to 1)   (M/bwtmerger.scala:346-350) ; bs[256]=n */
    int64_t   bs0[ALPHA];   /* bucketStarts0 (raw counts)           (M/bwtmerger.scala:341-345) */
    uint32_t *fm;           /* .fm payload: per-char sorted BWT positions (M/bwtmerger.scala:491-510) */
    uint32_t *sa;           /* optional full suffix array           (M/util.scala:213-224) */
} fmo_index;

static char g_err[512];
const char *fmo_last_error(void) { return g_err; }

/* ---- byte order helpers ------------------------------------------------------------------ */
static int64_t rd_i64(const uint8_t *p, int be) {
    uint64_t v = 0;
    if (be) for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
    else    for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
    return (int64_t)v;
}

static uint8_t *slurp(const char *path, int64_t *len) {
    FILE *f = fopen(path, "rb");
    if (!f) { snprintf(g_err, sizeof g_err, "cannot open %s", path); return NULL; }
    fseek(f, 0, SEEK_END); int64_t L = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t *b = (uint8_t *)malloc(L > 0 ? L : 1);
    if (L > 0 && fread(b, 1, L, f) != (size_t)L) { fclose(f); free(b); snprintf(g_err, sizeof g_err, "short read %s", path); return NULL; }
    fclose(f); *len = L; return b;
}

/* c2bs: M/util.scala:109-119 — exclusive prefix sums of the count table */
static void c2bs(const int64_t *c, int64_t *bs) {
    int64_t tot = 0;
    for (int i = 0; i < ALPHA; i++) { bs[i] = tot; tot += c[i]; }
}

/* FMCreator.create: M/bwtmerger.scala:452-532 (bucketStarts :440-450). Scanning the BWT in row order,
 * c = (i==eof ? 0 : bwt[i]); fm[bkt[c]++] = i, with bkt[0]=0 and bkt[c]=1+sum_{1<=k<c} cnt[k]. */
static void fm_create(fmo_index *ix) {
    int64_t bkt[ALPHA]; int64_t tot = 1;
    bkt[0] = 0;
    for (int i = 1; i < ALPHA; i++) { bkt[i] = tot; tot += ix->cnt[i]; }
    ix->fm = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)ix->n);
    for (int64_t i = 0; i < ix->n; i++) {
        int c = (i == ix->eof) ? 0 : ix->bwt[i];
        ix->fm[bkt[c]++] = (uint32_t)i;
    }
}

static void finish_tables(fmo_index *ix) {
    int64_t c1[ALPHA];
    memcpy(c1, ix->cnt, sizeof c1);
    c2bs(c1, ix->bs0);                 /* bucketStarts0: raw counts      */
    c1[0] = 1;
    c2bs(c1, ix->bs);                  /* bucketStarts : counts[0] := 1  */
    ix->bs[ALPHA] = ix->n;
    ix->bwt[ix->eof] = 0;              /* BWTLoader.read(eof) == 0       */
}

/* Build an index object from in-memory BWT (+eof). cnt may be NULL (then derived from the BWT
 * with the eof row excluded, i.e. what writeAuxFile M/bwtmerger.scala:841-856 stores). */
fmo_index *fmo_from_memory(const uint8_t *bwt, int64_t n, int64_t eof, const int64_t *cnt) {
    fmo_index *ix = (fmo_index *)calloc(1, sizeof *ix);
    ix->n = n; ix->eof = eof;
    ix->bwt = (uint8_t *)malloc((size_t)n);
    memcpy(ix->bwt, bwt, (size_t)n);
    ix->bwt[eof] = 0;
    if (cnt) memcpy(ix->cnt, cnt, sizeof ix->cnt);
    else for (int64_t i = 0; i < n; i++) if (i != eof) ix->cnt[ix->bwt[i]]++;
    finish_tables(ix);
    fm_create(ix);
    return ix;
}

/* NaiveFMSearcher ctor: M/bwtmerger.scala:335-350; loaders :130-174, :252-262.
 * base = path without extension. If <base>.fm exists it is validated and read (FMLoader), otherwise
 * it is materialised in memory with the FMCreator rule (testdata/words.fm is missing upstream). */
fmo_index *fmo_load(const char *base, int big_endian) {
    char p[4096]; int64_t len;
    snprintf(p, sizeof p, "%s.bwt", base);
    uint8_t *b = slurp(p, &len); if (!b) return NULL;
    if (len < 16) { free(b); snprintf(g_err, sizeof g_err, "File %s bad size", p); return NULL; }
    int64_t n = rd_i64(b, big_endian), eof = rd_i64(b + 8, big_endian);
    if (n + 16 != len) { free(b); snprintf(g_err, sizeof g_err, "File %s bad size %lld != %lld + 16", p, (long long)n, (long long)len); return NULL; }
    fmo_index *ix = (fmo_index *)calloc(1, sizeof *ix);
    ix->n = n; ix->eof = eof;
    ix->bwt = (uint8_t *)malloc((size_t)n); memcpy(ix->bwt, b + 16, (size_t)n); free(b);

    snprintf(p, sizeof p, "%s.aux", base);
    b = slurp(p, &len); if (!b || len != 2048) { snprintf(g_err, sizeof g_err, "bad aux %s", p); return NULL; }
    for (int i = 0; i < ALPHA; i++) ix->cnt[i] = rd_i64(b + 8 * i, big_endian);
    free(b);
    finish_tables(ix);

    snprintf(p, sizeof p, "%s.fm", base);
    FILE *f = fopen(p, "rb");
    if (f) {
        fclose(f);
        b = slurp(p, &len); if (!b) return NULL;
        if (len < 9 || b[0] != 4) { snprintf(g_err, sizeof g_err, "File %s bad elSize", p); return NULL; }
        int64_t fn = rd_i64(b + 1, big_endian);
        if (fn * 4 + 9 != len || fn != n) { snprintf(g_err, sizeof g_err, "File %s bad size", p); return NULL; }
        ix->fm = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
        for (int64_t i = 0; i < n; i++) {            /* payload is always big-endian int32: setIntValOn :476-481 */
            const uint8_t *q = b + 9 + 4 * i;
            ix->fm[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
        }
        free(b);
    } else {
        fm_create(ix);
    }
    return ix;
}

void fmo_free(fmo_index *ix) { if (!ix) return; free(ix->bwt); free(ix->fm); free(ix->sa); free(ix); }
int64_t fmo_n(const fmo_index *ix) { return ix->n; }
int64_t fmo_eof(const fmo_index *ix) { return ix->eof; }
const uint8_t  *fmo_bwt(const fmo_index *ix) { return ix->bwt; }
const uint32_t *fmo_fm(const fmo_index *ix) { return ix->fm; }
int64_t fmo_cf(const fmo_index *ix, int c) { return ix->bs[c]; }          /* cf: M/bwtmerger.scala:352 */

/* occ(c,key): M/bwtmerger.scala:354-375 — literal binary search over fm[bucket(c)]:
 * number of entries <= key, i.e. number of c in BWT[0..key]. */
int64_t fmo_occ(const fmo_index *ix, int c, int64_t key) {
    int64_t istart = ix->bs[c];
    int64_t imin = istart;
    int64_t imax = (c == ALPHA - 1) ? ix->n - 1 : ix->bs[c + 1] - 1;
    if (imin <= imax) {
        int found = 0; int64_t imid = 0, ival = 0;
        while (!found && imax >= imin) {
            imid = (imax + imin) / 2;
            ival = ix->fm[imid];
            if (ival < key) imin = imid + 1;
            else if (ival > key) imax = imid - 1;
            else found = 1;
        }
        return (ival <= key) ? (imid - istart + 1) : (imid - istart);
    }
    return 0;
}

/* getPrevRange: M/findex.scala:32-36. Returns 1 iff non-empty. */
int fmo_prev_range(const fmo_index *ix, int64_t sp, int64_t ep, int c, int64_t *sp1, int64_t *ep1) {
    int64_t a = ix->bs[c] + fmo_occ(ix, c, sp - 1);
    int64_t b = ix->bs[c] + fmo_occ(ix, c, ep - 1);
    *sp1 = a; *ep1 = b;
    return a < b;
}

/* search: M/findex.scala:15-31. Returns 1 (Some) / 0 (None); (sp,ep) always written. */
int fmo_search(const fmo_index *ix, const uint8_t *pat, int64_t len, int64_t *sp_out, int64_t *ep_out) {
    int64_t sp = 0, ep = ix->n, i = len - 1;
    while (sp < ep && i >= 0) {
        int c = pat[i]; i--;
        int64_t nsp = ix->bs[c] + fmo_occ(ix, c, sp - 1);
        int64_t nep = ix->bs[c] + fmo_occ(ix, c, ep - 1);
        sp = nsp; ep = nep;
    }
    *sp_out = sp; *ep_out = ep;
    return sp < ep;
}

/* getIntervalPrevRange: M/findex.scala:37-51. Output in the reference's order (descending c,
 * because results are prepended). out_c/out_sp/out_ep must hold cend-cstart+1 entries. */
int64_t fmo_interval_prev_range(const fmo_index *ix, int64_t sp, int64_t ep, int cstart, int cend,
                                int32_t *out_c, int64_t *out_sp, int64_t *out_ep) {
    int64_t k = 0;
    for (int c = cend; c >= cstart; c--) {
        int64_t o1 = fmo_occ(ix, c, sp - 1), o2 = fmo_occ(ix, c, ep - 1);
        if (o1 < o2) { out_c[k] = c; out_sp[k] = ix->bs[c] + o1; out_ep[k] = ix->bs[c] + o2; k++; }
    }
    return k;
}

/* getPrevI / getNextI: M/bwtmerger.scala:386-392 */
int64_t fmo_get_prev_i(const fmo_index *ix, int64_t i) { int c = ix->bwt[i]; return ix->bs[c] + fmo_occ(ix, c, i - 1); }
int64_t fmo_get_next_i(const fmo_index *ix, int64_t i) { return ix->fm[i]; }

/* pos2char: M/bwtmerger.scala:376-385 (uses bucketStarts0) */
int fmo_pos2char(const fmo_index *ix, int64_t key) {
    int i = ALPHA - 1;
    if (ix->bs0[i] > key) { while (ix->bs0[i] > key && i > 0) i--; }
    else { while (ix->bs0[i - 1] == ix->bs0[i] && i > 1) i--; i--; }
    return i;
}

/* nextSubstr: M/bwtmerger.scala:394-405. Writes <= len bytes, returns count. */
int64_t fmo_next_substr(const fmo_index *ix, int64_t sp, int64_t len, uint8_t *out) {
    int64_t cp = ix->fm[sp], k = 0; int eof = 0;
    for (int64_t i = 0; i < len && !eof; i++) {
        uint8_t b = ix->bwt[cp];
        eof = (b == 0);
        out[k++] = b;
        cp = ix->fm[cp];
    }
    for (int64_t a = 0, z = k - 1; a < z; a++, z--) { uint8_t t = out[a]; out[a] = out[z]; out[z] = t; }  /* ret.reverse */
    return k;
}

/* prevSubstr: M/bwtmerger.scala:409-419 (its eof flag is never set, so exactly len chars). */
int64_t fmo_prev_substr(const fmo_index *ix, int64_t sp, int64_t len, uint8_t *out) {
    int64_t cp = sp;
    for (int64_t i = 0; i < len; i++) { out[i] = ix->bwt[cp]; cp = fmo_get_prev_i(ix, cp); }
    return len;
}

/* bwtFm2sa: M/util.scala:213-224 == SACreator.create M/bwtmerger.scala:541-555 */
const uint32_t *fmo_build_sa(fmo_index *ix) {
    if (ix->sa) return ix->sa;
    ix->sa = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)ix->n);
    int64_t i = ix->eof;
    for (int64_t j = 0; j < ix->n; j++) { ix->sa[i] = (uint32_t)j; i = ix->fm[i]; }
    return ix->sa;
}

/* bwtFm2LCP: M/util.scala:153-212 (== LCPCreator.create M/bwtmerger.scala:583-650, which writes the same values to <base>.lcp).
 * Walks the text positions i = 0..n-1 through k = fm(k) starting at the eof row; for row k it compares the suffix of row k with the
 * suffix of row j = k-1 character by character — ibs2c(row) = the F-column character of a row, iterChar advancing both rows by fm —
 * carrying h-1 over to the next text position, and stores h at LCP(k-1); row 0 stores LCP(0) = 0.  out[n]: LCP(n-1) is never written
 * (stays 0); the file holds the first max(n-1, 1) entries. */
static int ibs2c(const fmo_index *ix, int64_t i) {           /* bs.indexWhere(i < _, 0) - 1 ; no such index: -1 - 1 */
    for (int j = 0; j < ALPHA; j++) if (i < ix->bs[j]) return j - 1;
    return -2;
}
static int iter_char(const fmo_index *ix, int64_t j, int64_t h, int64_t *temp) {
    if (h != 0 && *temp == -1) {
        while (h > 0) { j = ix->fm[j]; h--; }
        *temp = j;
    } else if (*temp != -1) {
        j = ix->fm[*temp];
        *temp = j;
    }
    return ibs2c(ix, j);
}
void fmo_build_lcp(const fmo_index *ix, int32_t *out) {
    const int64_t n = ix->n;
    for (int64_t r = 0; r < n; r++) out[r] = 0;
    int64_t i = 0, k = ix->eof, h = 0;
    while (i < n) {
        if (k == 0) out[k] = 0;
        else {
            int64_t temp1 = -1, temp2 = -1, j = k - 1;
            int stop = 0;
            while (i + h < n && !stop) {
                int64_t t1 = temp1, t2 = temp2;
                const int c1 = iter_char(ix, k, h, &t1), c2 = iter_char(ix, j, h, &t2);
                if (c1 == c2) { temp1 = t1; temp2 = t2; h++; }
                else stop = 1;
            }
            out[k - 1] = (int32_t)h;
        }
        if (h > 0) h--;
        k = ix->fm[k];
        i++;
    }
}

static int cmp_i64(const void *a, const void *b) { int64_t x = *(const int64_t *)a, y = *(const int64_t *)b; return (x > y) - (x < y); }

/* locate(sp,ep) := sorted { sa[r] : r in [sp,ep) }   (T' coordinates) */
int64_t fmo_locate(fmo_index *ix, int64_t sp, int64_t ep, int64_t *out) {
    fmo_build_sa(ix);
    int64_t k = 0;
    for (int64_t r = sp; r < ep; r++) out[k++] = ix->sa[r];
    qsort(out, (size_t)k, sizeof(int64_t), cmp_i64);
    return k;
}

/* ---- batched count (for parity at scale and for the CPU baseline) ------------------------- */
typedef struct { const fmo_index *ix; const uint8_t *pat; const int64_t *off; int64_t lo, hi; int64_t *sp, *ep; } job_t;
static void *count_worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (int64_t q = j->lo; q < j->hi; q++) {
        int64_t sp, ep;
        int hit = fmo_search(j->ix, j->pat + j->off[q], j->off[q + 1] - j->off[q], &sp, &ep);
        if (!hit) { sp = 0; ep = 0; }              /* None is reported as (0,0) at the C ABI */
        j->sp[q] = sp; j->ep[q] = ep;
    }
    return NULL;
}
void fmo_count_batch(const fmo_index *ix, const uint8_t *pat, const int64_t *off, int64_t m,
                     int64_t *sp, int64_t *ep, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256]; job_t jb[256];
    for (int t = 0; t < threads; t++) {
        jb[t] = (job_t){ ix, pat, off, m * t / threads, m * (t + 1) / threads, sp, ep };
        pthread_create(&th[t], NULL, count_worker, &jb[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}

/* ---- Glushkov traversal: ReTree._matchSA, M/re2/retree.scala:618-653, uncapped -------------
 * Automaton is passed as flat tables produced by oracle/retree.py (the Python restatement of
 * ReTree.apply). Emits (len,sp,ep) triples; order is unspecified (parity object = sorted multiset).
 * Returns number of triples, or -1 if cap exceeded (then *needed is a lower bound). */
typedef struct { int32_t st; int32_t len; int64_t sp, ep; } item_t;
int64_t fmo_regex_match(const fmo_index *ix, int32_t nstates, const uint8_t *st_c, const uint8_t *st_last,
                        const int32_t *fol_off, const int32_t *fol, const int32_t *firsts, int32_t nfirsts,
                        int64_t cap, int32_t *out_len, int64_t *out_sp, int64_t *out_ep, int64_t max_expansions,
                        int64_t *n_expansions, int stop_on_emit, int64_t max_len) {
    /* max_len = REParser.matchSA's maxLength (re2.scala:636-641): a new non-match state point is enqueued only while its len < maxLength
     * (0 = no limit); matches are emitted whatever their length.  Independent of the order in which items are taken. */
    (void)nstates;
    size_t scap = 1024, top = 0; item_t *stk = (item_t *)malloc(scap * sizeof *stk);
    for (int32_t i = 0; i < nfirsts; i++) {
        if (top == scap) { scap *= 2; stk = (item_t *)realloc(stk, scap * sizeof *stk); }
        stk[top++] = (item_t){ firsts[i], 0, 0, ix->n };
    }
    int64_t nout = 0, nexp = 0;
    while (top) {
        item_t q = stk[--top];
        int64_t sp1, ep1;
        nexp++;
        if (max_expansions && nexp > max_expansions) { free(stk); return -2; }
        if (fmo_prev_range(ix, q.sp, q.ep, st_c[q.st], &sp1, &ep1)) {
            if (st_last[q.st]) {
                if (nout < cap) { out_len[nout] = q.len + 1; out_sp[nout] = sp1; out_ep[nout] = ep1; }
                nout++;
            }
            /* Glushkov (retree.scala:638-645): a last position emits and is not expanded.  Thompson (re2.scala:636-641): the
             * MatchState of the closure emits, the other states of the same closure are still enqueued. */
            if (!(st_last[q.st] && stop_on_emit) && (max_len == 0 || q.len + 1 < max_len)) {
                for (int32_t k = fol_off[q.st]; k < fol_off[q.st + 1]; k++) {
                    if (top == scap) { scap *= 2; stk = (item_t *)realloc(stk, scap * sizeof *stk); }
                    stk[top++] = (item_t){ fol[k], q.len + 1, sp1, ep1 };
                }
            }
        }
    }
    free(stk);
    if (n_expansions) *n_expansions = nexp;
    return nout;
}

/* ---- suffix sorting for TEST index construction (any correct sorter yields the unique BWT) --
 * t = text bytes WITHOUT terminator (the already-reversed text T'), len = n-1. Produces BWT of
 * t+'$' with '$' smaller than every byte, the eof row, and the 256 counts.
 * Comparator sort (memcmp) — O(n log n * LCP); meant for texts up to a few 10 MB. */
static const uint8_t *g_t; static int64_t g_len;
static int cmp_suf(const void *a, const void *b) {
    int64_t i = *(const uint32_t *)a, j = *(const uint32_t *)b;
    if (i == j) return 0;
    int64_t li = g_len - i, lj = g_len - j, l = li < lj ? li : lj;
    int r = memcmp(g_t + i, g_t + j, (size_t)l);
    if (r) return r;
    return li < lj ? -1 : 1;                       /* shorter suffix hits '$' first => smaller */
}
int fmo_build_bwt(const uint8_t *t, int64_t len, uint8_t *bwt_out, int64_t *eof_out, int64_t *cnt_out) {
    int64_t n = len + 1;
    uint32_t *sa = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
    for (int64_t i = 0; i < len; i++) sa[i + 1] = (uint32_t)i;
    sa[0] = (uint32_t)len;                         /* the '$' suffix sorts first */
    g_t = t; g_len = len;
    qsort(sa + 1, (size_t)len, sizeof(uint32_t), cmp_suf);
    memset(cnt_out, 0, sizeof(int64_t) * ALPHA);
    for (int64_t i = 0; i < len; i++) cnt_out[t[i]]++;
    for (int64_t r = 0; r < n; r++) {
        if (sa[r] == 0) { bwt_out[r] = 0; *eof_out = r; }
        else bwt_out[r] = t[sa[r] - 1];
    }
    free(sa);
    return 0;
}

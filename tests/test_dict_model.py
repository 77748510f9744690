"""
The dictionary of wide intervals (FMX_ACCEL_DICT) as an executable model on the CPU: the stored set, the keys and the
lookup order are restated in a few lines of Python over the oracle's suffix array, and the model's answers are compared with the
oracle's plain backward search (findex.scala:15-31).  This pins the ALGORITHM — why a stored entry is exact, why bisection over the
depth finds the deepest stored prefix with per-level thresholds, why {sp of the tier's parent, j, tier, symbols} names a chain entry
uniquely — independently of the CUDA code, which tests/test_gpu_parity.py::test_wide_interval_dictionary checks against the oracle too.
"""
import numpy as np
import pytest

from oracle import fm_oracle as fo

K, KCTX = 2, 8                                             # depth of the dense table in front; rows a context takes (chain threshold)


def _zipf_text(rng, alpha, n_words, n_vocab):
    vocab = [alpha[1 + rng.integers(0, len(alpha) - 1, int(rng.integers(2, 10)))] for _ in range(n_vocab)]
    p = 1.0 / np.arange(1, n_vocab + 1)
    idx = rng.choice(n_vocab, n_words, p=p / p.sum())
    return np.concatenate([np.concatenate([vocab[i], alpha[:1]]) for i in idx]).tobytes()


class Model:
    def __init__(self, o, tp, sigma_bits, min_rows, top_rows, max_depth=None):
        """o: oracle index of T' = tp + '$'.  Stored: every d-mer (consumption order = a prefix of a suffix of T', read forwards) with
        K < d <= D whose interval has more than min_rows rows (more than top_rows at d = D); chain entries past D with more than
        max(min_rows, 8) rows."""
        self.o, self.bits = o, sigma_bits
        self.D = min(16, 60 // sigma_bits)
        self.Jc = max(1, min(8, 22 // sigma_bits))
        sa = o.sa().astype(np.int64)
        n = len(sa)
        syms = sorted(set(tp))
        self.code = {c: i for i, c in enumerate(syms)}
        txt = tp + b"\0"
        self.tab = {}
        self.iv = {}                                       # consumed symbols (bytes, in consumption order) -> interval, every depth
        Dx = (max_depth or (self.D + 7 * self.Jc))
        # groups of rows sharing their first d symbols, straight from the sorted suffixes.  A pattern is consumed last byte first and the
        # interval after d steps holds the suffixes that START with its last d bytes: consumed order = that prefix read backwards.
        for d in range(K + 1, Dx + 1):
            thr = top_rows if d == self.D else (min_rows if d < self.D else max(min_rows, KCTX))
            r = 0
            while r < n:
                s = sa[r]
                pre = txt[s:s + d]
                if len(pre) < d or 0 in pre:
                    r += 1
                    continue
                e = r + 1
                while e < n and txt[sa[e]:sa[e] + d] == pre:
                    e += 1
                cons = pre[::-1]
                self.iv[cons] = (r, e)
                if e - r > thr:                            # (its ancestors hold at least as many rows and meet thresholds that are no higher:
                    self.tab[self._key(cons)] = (r, e)     #  the level-wise construction, which grows from what it kept, reaches it)
                r = e
        self.Dx = max([self.D] + [self._depth_of(k) for k in self.tab])

    def _pack(self, pre):
        v = 0
        for j, c in enumerate(pre):
            v |= self.code[c] << (self.bits * j)
        return v

    def _key(self, cons):
        d = len(cons)
        if d <= self.D:
            return ("lvl", d, self._pack(cons))            # the d codes, first consumed lowest | (d-1) << 60
        t = (d - self.D - 1) // self.Jc + 1
        base = self.D + (t - 1) * self.Jc
        psp = self.iv[cons[:base]][0]                      # sp of the interval the tier starts from
        return ("chain", psp, d - base, t, self._pack(cons[base:]))

    def _depth_of(self, k):
        return k[1] if k[0] == "lvl" else self.D + (k[3] - 1) * self.Jc + k[2]

    def search(self, pat):
        """SuffixAlgo.search with the dictionary in front.  pat is consumed last byte first: the consumed symbols, in order, read a
        prefix of T' suffixes forwards."""
        o = self.o
        cons = bytes(pat[::-1])
        sp, ep, i = 0, o.n, 0
        probes = 0
        if len(cons) > K and all(c in self.code for c in cons[:min(len(cons), self.D)]):
            hi, lo, best = min(len(cons), self.D), K, None
            d = hi
            while lo < hi:
                probes += 1
                k = ("lvl", d, self._pack(cons[:d]))
                if k in self.tab:
                    lo, best = d, self.tab[k]
                else:
                    hi = d - 1
                d = (lo + hi + 1) >> 1
            if best is not None:
                sp, ep = best
                i = lo
                if lo == self.D:                           # chain tiers
                    t = 1
                    while i < len(cons) and i < self.Dx and t <= 7:
                        j = min(len(cons) - i, self.Jc)
                        nxt = cons[i:i + j]
                        if not all(c in self.code for c in nxt):
                            break
                        k = ("chain", sp, j, t, self._pack(nxt))
                        probes += 1
                        if k in self.tab:
                            sp, ep = self.tab[k]
                            i += j
                            if j < self.Jc:
                                break
                            t += 1
                            continue
                        a, b, hit = 0, j - 1, None         # bisection inside the tier
                        while a < b:
                            mid = (a + b + 1) >> 1
                            probes += 1
                            km = ("chain", sp, mid, t, self._pack(nxt[:mid]))
                            if km in self.tab:
                                a, hit = mid, self.tab[km]
                            else:
                                b = mid - 1
                        if hit is not None:
                            sp, ep = hit
                            i += a
                        break
        while i < len(cons) and sp < ep:                   # the ordinary steps
            r = o.getPrevRange(sp, ep, cons[i])
            sp, ep = r if r else (0, 0)
            i += 1
        return ((sp, ep) if sp < ep else None), probes


@pytest.mark.parametrize("sigma,bits,min_rows,top_rows", [(4, 2, 8, 8), (27, 5, 2, 1), (27, 5, 8, 1), (60, 6, 3, 3)])
def test_dictionary_model_equals_backward_search(sigma, bits, min_rows, top_rows):
    rng = np.random.default_rng(40 + sigma + min_rows)
    alpha = rng.choice(np.arange(1, 255), sigma, replace=False).astype(np.uint8)
    text = _zipf_text(rng, alpha, 1500, 60)
    tp = bytes(fo.file_to_text_rev(text))
    bwt, eof, cnt = fo.build_bwt(fo.file_to_text_rev(text))
    o = fo.OracleIndex.from_bwt(bwt, eof, cnt)
    m = Model(o, tp, bits, min_rows, top_rows, max_depth=min(16, 60 // bits) + 2 * max(1, min(8, 22 // bits)))
    assert len(m.tab) > 50 and (m.Dx > m.D or bits != 5)   # keyed levels everywhere, chain entries behind depth 12 on the 5-bit texts
    one_probe = 0
    for ln in list(range(1, 26)) + [31]:
        for _ in range(25):
            s = int(rng.integers(0, len(tp) - ln))
            q = bytearray(tp[s:s + ln])                     # a substring of T' (search() consumes it last byte first)
            u = rng.random()
            if u < 0.25:
                q[int(rng.integers(0, ln))] = int(alpha[rng.integers(0, sigma)])
            elif u < 0.30:
                q[int(rng.integers(0, ln))] = 255           # a byte that does not occur
            got, probes = m.search(bytes(q))
            assert got == o.search(bytes(q)), (bytes(q), got)
            one_probe += probes == 1
    assert one_probe > 50                                   # whole prefixes found with a single probe
    o.close()

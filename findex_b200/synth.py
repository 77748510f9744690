"""
findex_b200.synth — seeded synthetic workloads of BASELINE.json's configs, shared by bench.py, tools/ and the
full-size tests: uniform byte text (cfg 2), English-like Zipf text over a vocabulary (cfg 3/4), DNA (cfg 5), and the cfg-4 regex
templates.  Pure numpy; nothing here touches the GPU or the oracle.
"""
import numpy as np


def uniform_bytes(n, seed=2):
    return np.random.default_rng(seed).integers(1, 256, n, dtype=np.uint8)


def dna(n, seed=7):
    return np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(seed).integers(0, 4, n, dtype=np.uint8)]


def english_like(vocab, n_bytes, seed=4):
    """Words drawn Zipf(s=1) over a seed-permuted vocabulary (list of bytes), space separated, '\\n' every 12 words."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(len(vocab))
    lens = np.array([len(vocab[i]) for i in perm], np.int64) + 1                # + separator
    flat = np.frombuffer(b"".join(vocab[i] + b" " for i in perm), np.uint8)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    p = 1.0 / np.arange(1, len(vocab) + 1)
    cdf = np.cumsum(p / p.sum())
    out = np.empty(n_bytes, np.uint8)
    pos, widx = 0, 0
    while pos < n_bytes:
        k = int(min(4_000_000, max(1000, (n_bytes - pos) // 4)))
        r = np.minimum(np.searchsorted(cdf, rng.random(k)), len(vocab) - 1)
        wl = lens[r]
        ends = np.cumsum(wl)
        total = int(ends[-1])
        chunk = flat[np.repeat(starts[r] - (ends - wl), wl) + np.arange(total)].copy()
        chunk[(ends - 1)[(np.arange(widx, widx + k) % 12) == 11]] = 10
        take = min(total, n_bytes - pos)
        out[pos:pos + take] = chunk[:take]
        pos += take
        widx += k
    return out


def reversed_substrings(text, m, ln, rng):
    """m patterns = substrings of the text at uniform offsets, reversed (what search() consumes: it matches in reverse(file))."""
    offs = rng.integers(0, len(text) - ln, m)
    return text[offs[:, None] + np.arange(ln - 1, -1, -1)[None, :]], offs


def regex_templates(text, rng, m):
    """cfg-4 templates the ReTree grammar accepts: L3[c1-c2]L2, L3(w|w|w)L1, L2 x? y{1,3} L2 (desugared), L3\\dL2, L4.L2."""
    def lit(k):
        s = int(rng.integers(0, len(text) - k))
        return bytes(text[s:s + k])

    def esc(b):
        return b"".join((b"\\" + bytes([c])) if c in b"()[]|*+?.\\-" else bytes([c]) for c in b)
    out = []
    while len(out) < m:
        t = len(out) % 5
        if t == 0:
            a, b = sorted(rng.integers(97, 123, 2).tolist())
            if a == b:
                b = min(a + 1, 122)
                a = b - 1
            out.append(esc(lit(3)) + b"[" + bytes([a]) + b"-" + bytes([b]) + b"]" + esc(lit(2)))
        elif t == 1:
            out.append(esc(lit(3)) + b"(" + esc(lit(2)) + b"|" + esc(lit(2)) + b"|" + esc(lit(3)) + b")" + esc(lit(1)))
        elif t == 2:
            x, y = esc(lit(1)), esc(lit(1))
            out.append(esc(lit(2)) + x + b"?" + y + y + b"?" + y + b"?" + esc(lit(2)))
        elif t == 3:
            out.append(esc(lit(3)) + b"\\d" + esc(lit(2)))
        else:
            out.append(esc(lit(4)) + b"." + esc(lit(2)))
    return out

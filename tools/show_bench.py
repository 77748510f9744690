"""prints the headline and the per-leg summaries of a bench.py JSON line:  python tools/show_bench.py FILE [leg ...]"""
import json
import sys


def main():
    d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    legs = sys.argv[2:]
    top = {k: v for k, v in d.items() if not isinstance(v, (dict, list))}
    print(top)
    for k in ("e2e", "roofline", "clocks"):
        if k in d and (not legs or k in legs):
            print(k, json.dumps(d[k])[:1500])
    e = d.get("english", {})
    if e and (not legs or "english" in legs):
        print("english index", e.get("index"))
        for c in e.get("count", []):
            print("  ", c)
    for k in ("locate", "regex", "cfg5", "sustained", "pcie"):
        v = d.get(k) or e.get(k)
        if v and (not legs or k in legs):
            print(k, json.dumps({a: b for a, b in v.items() if a not in ("clocks", "what", "exchange")})[:2500])
    for c in d.get("sweep", []) if (not legs or "sweep" in legs) else []:
        print("sweep", c)


if __name__ == "__main__":
    main()

"""
tools/ncu_skip.py <launches.csv> <queries> <lanes> — reads an ncu launch list (--metrics gpu__time_duration.sum --csv) of bench.py and
prints how many count_fixed_kernel launches precede the 5th full-batch launch of the timed kernel (the k-mer table fill at open time
uses the same kernel on 2^24-pattern slices), i.e. the value for `ncu -k regex:count_fixed_kernel -s N`.
"""
import csv
import sys

path, m, lanes = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
grid = (m + 256 // lanes - 1) // (256 // lanes)
rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 8]
hdr = rows[0]
ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
seen, full = 0, 0
for r in rows[1:]:
    if "count_fixed_kernel" not in r[ki]:
        continue
    g = int(r[gi].strip("()").split(",")[0])
    if g == grid and "true" not in r[ki] and "(bool)1" not in r[ki]:
        full += 1
        if full == 5:
            print(seen)
            sys.exit(0)
    seen += 1
print(max(seen - 2, 0))

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one steady-state chunk of the
streaming loop (between two consecutive ar_decode launches), per kernel and per stage."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    rows = []
    for row in r:
        try:
            rows.append((row[ki], float(row[vi].replace(",", "")), row[gi]))
        except ValueError:
            pass
    return rows


def main(path, which=-3):
    rows = load(path)
    idx = [i for i, (k, _, _) in enumerate(rows) if "ar_decode" in k]
    a, b = idx[which], idx[which + 1]
    step = rows[a:b]
    tot = sum(v for _, v, _ in step)
    print(f"# {path}: {len(rows)} launches; chunk = launches [{a},{b}) : {len(step)} kernels, {tot / 1e3:.1f} us (ncu, serialised, cold cache)")
    # stage split: ar_decode | V ... (until first magnitude/fill before the DFT gemm = next chunk's E)
    bsq = max(i for i, (k, _, _) in enumerate(step) if "bsq_kernel" in k) if any("bsq_kernel" in k for k, _, _ in step) else None
    first_e = next(i for i, (k, _, _) in enumerate(step) if "magnitude" in k) - 2
    stages = {"A (ar_decode + append)": step[0:2], "V": step[2:first_e], "E": step[first_e:]}
    for name, s in stages.items():
        print(f"  stage {name:24s} {sum(v for _, v, _ in s) / 1e3:9.1f} us  {len(s):4d} launches")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v, g in step:
        k = re.sub(r"\(.*", "", k).replace("svanon::<unnamed>::", "").replace("void ", "")
        agg[k][0] += 1
        agg[k][1] += v
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {v / 1e3:9.1f} us {100 * v / tot:5.1f}%  n={n:4d}  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else -3)

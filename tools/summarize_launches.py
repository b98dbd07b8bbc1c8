"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of one step.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [anchor-kernel-substring]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    rows = []
    for row in r:
        try:
            v = float(row[vi].replace(",", ""))
        except ValueError:
            continue
        u = row[ui]
        v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
        rows.append((row[ki], v))
    return rows


def short(n):
    n = n.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    n = re.sub(r"^void ", "", n)
    m = re.match(r"([A-Za-z0-9_:]+)", n)
    base = m.group(1) if m else n
    if base.startswith("at::native::") and "<" in n:
        inner = re.search(r"at::native::([A-Za-z0-9_]+)", n[len(base):])
        if inner:
            base += "<" + inner.group(1) + ">"
    return base[-80:]


def main():
    rows = load(sys.argv[1])
    anchor = sys.argv[2] if len(sys.argv) > 2 else "corr_fast_kernel"
    names = [n for n, _ in rows]
    idx = [i for i, n in enumerate(names) if anchor in n]
    print("%d launches, %d anchors" % (len(rows), len(idx)))
    if len(idx) < 3:
        return
    period = idx[-1] - idx[-2]
    s = idx[-2] + 1
    step = rows[s:s + period]     # one full period between two anchors (order rotated, content identical)
    # the L2 flush between two steps (a 256 MiB fill, ~70 us) sits in the launch list but outside the timed bracket
    flush = [(n, v) for n, v in step if "vectorized_elementwise_kernel" in n and "FillFunctor" in n and v > 30.0]
    step = [(n, v) for n, v in step if (n, v) not in flush]
    if flush:
        print("(excluded: the L2 flush between steps, %s)" % ", ".join("%.1f us" % v for _, v in flush))
    tot = sum(v for _, v in step)
    print("one step = %d launches, %.1f us of serialised kernel time" % (len(step), tot))
    agg = collections.OrderedDict()
    for n, v in step:
        k = short(n)
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%8.1f us %5.1f%%  x%-3d %s" % (v, 100 * v / tot, c, k))


if __name__ == "__main__":
    main()

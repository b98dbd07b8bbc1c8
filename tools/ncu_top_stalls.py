"""Top stall-sample SASS lines of an `ncu --page source --csv --print-source sass` export: python tools/ncu_top_stalls.py file.csv [n]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]

    def num(x):
        try:
            return int(float(x))
        except ValueError:
            return 0
    seen, uniq = set(), []
    for r in data:
        if r[0] not in seen:
            seen.add(r[0])
            uniq.append(r)
    tot = sum(num(r[i_s]) for r in uniq)
    print("total samples", tot)
    agg = {}
    for r in uniq:
        for i in stall:
            agg[hdr[i]] = agg.get(hdr[i], 0) + num(r[i])
    print("by reason:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for r in sorted(uniq, key=lambda r: -num(r[i_s]))[:n]:
        st = sorted([(num(r[i]), hdr[i][6:]) for i in stall], reverse=True)[:2]
        print("%s %5s %8s  %-70s %s" % (r[0][-5:], r[i_s], r[i_ex], r[i_src][:70], st))


if __name__ == "__main__":
    main()

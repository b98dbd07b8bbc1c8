"""DRAM traffic per launch from an .ncu-rep (dram__bytes_read.sum + dram__bytes_write.sum), summed over the launches whose
kernel name contains the given substring; writes the JSON bench.py reads for `roofline.traffic`.
usage: python tools/ncu_traffic.py rep kernel_substring out.json [launches_per_update]"""
import csv
import io
import json
import subprocess
import sys


def main():
    rep, sub, out = sys.argv[1], sys.argv[2], sys.argv[3]
    per_update = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")

    def to_bytes(v, u):
        v = float(v)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    sel = [r for r in data if sub in r[ki]]
    tot = sum(to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi]) for r in sel)
    n_updates = max(len(sel) // per_update, 1)
    d = dict(kernel=sub, launches_captured=len(sel), launches_per_update=per_update, dram_bytes_per_launch=tot / n_updates,
             source=rep.split("/")[-1], how="ncu --set full --clock-control none (default cache control: flushed before each replay); "
             "dram__bytes_read.sum + dram__bytes_write.sum summed over the kernel's launches of one update")
    json.dump(d, open(out, "w"), indent=1)
    print(d)


if __name__ == "__main__":
    main()

"""Static per-kernel resources of the built library (registers, stack frame, static shared memory) from
`cuobjdump --dump-resource-usage`; no GPU needed.  Usage: python tools/resource_usage.py > profiles/rNN_resource_usage.txt
A non-zero stack frame is per-thread local memory (indexed per-thread arrays or spills) -- the first thing to look at
in a kernel that ncu shows waiting on `lg_throttle` / local loads."""
import os
import re
import shutil
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devo_b200 import _lib  # noqa: E402


def demangle(names):
    filt = shutil.which("c++filt")
    if not filt:
        return names
    out = subprocess.run([filt], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return out[:len(names)]


def short(name):
    """drop the argument list, keep template arguments of the kernel itself"""
    name = re.sub(r"^void ", "", name).replace("(anonymous namespace)::", "")
    depth = 0
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return name[:i]
    return name


def main():
    cu = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    txt = subprocess.run([cu, "--dump-resource-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    rows = []
    fn = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and fn:
            rows.append((fn,) + tuple(int(x) for x in m.groups()))
            fn = None
    names = [short(n) for n in demangle([r[0] for r in rows])]
    seen = {}
    for n, r in zip(names, rows):
        if n.startswith("cub::"):
            n = "cub::" + n.split("<")[0].split("::")[-1] + "<...>"
        seen.setdefault(n, r[1:])
    print("%-110s %5s %6s %7s" % ("kernel (own namespace; cub kernels abbreviated)", "regs", "stack", "s.smem"))
    for n in sorted(seen, key=lambda k: (-seen[k][1], -seen[k][0], k)):
        reg, stack, smem, _ = seen[n]
        print("%-110s %5d %6d %7d" % (n[:110], reg, stack, smem))
    print("\n%d kernels, %d with a stack frame" % (len(seen), sum(1 for v in seen.values() if v[1])), file=sys.stdout)


if __name__ == "__main__":
    main()

"""Opcode histogram of every object in tamtr_b200/csrc/_build (cuobjdump -sass) -> profiles/sass_opcodes.txt.

The evidence that the tcgen05 / TMEM / TMA / vector-reduction paths are in the machine code, not only in the PTX:
UTCHMMA* (tcgen05.mma), UTMALDG / UTMASTG (TMA load / store), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit),
REDG.E.ADD.* (vector reductions), LDGSTS (cp.async).  Run after `make -C tamtr_b200/csrc`."""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = re.compile(r"^(UTC|UTMA|LDTM|STTM|REDG|LDGSTS|UCGABAR|SYNCS|UBLKCP|ATOMG|RED\b|MUFU|HMMA|REDUX)")


def histogram(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    per_fn, fn = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn)
            per_fn[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            per_fn[fn][m.group(1)] += 1
    return per_fn


def main():
    dst = os.path.join(ROOT, "profiles", "sass_opcodes.txt")
    with open(dst, "w") as f:
        f.write("# cuobjdump -sass opcode counts per kernel (sm_100a); only the opcodes that identify a hardware path are\n"
                "# listed individually, followed by the kernel's total instruction count.  tools/sass_opcodes.py\n")
        for obj in sorted(glob.glob(os.path.join(ROOT, "tamtr_b200", "csrc", "_build", "*.o"))):
            f.write(f"\n== {os.path.basename(obj)}\n")
            for fn, c in histogram(obj).items():
                keys = {k: v for k, v in c.items() if KEY.match(k)}
                f.write(f"{fn}\n    total {sum(c.values())}")
                for k in sorted(keys):
                    f.write(f"  {k} {keys[k]}")
                f.write("\n")
    print(dst)


if __name__ == "__main__":
    sys.exit(main())

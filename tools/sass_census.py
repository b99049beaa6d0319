#!/usr/bin/env python
"""Per-kernel census of the Blackwell-specific SASS in libhelmnet_sm100.so (cuobjdump -sass): tcgen05 MMAs (UTCHMMA), TMEM loads /
stores (LDTM / STTM), tcgen05.commit (UTCBAR), bulk copies (UBLKCP), mbarrier waits (SYNCS), programmatic dependent launch
(ACQBULK = griddepcontrol.wait, PREEXIT = griddepcontrol.launch_dependents), packed fp32 FMA (FFMA2), cp.async (LDGSTS).

    python tools/sass_census.py > profiles/r2_sass_census.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "helmnet_b200", "csrc", "libhelmnet_sm100.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "SYNCS", "UTCBAR", "ACQBULK", "PREEXIT", "FFMA2", "LDGSTS", "HMMA."]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
    census, cur, k = [], None, 0
    for line in out.splitlines():
        if "Function :" in line:
            cur = {"name": names[k] if k < len(names) else line.split(":")[1].strip(), "total": 0, **{o: 0 for o in OPS}}
            k += 1
            census.append(cur)
        elif cur is not None and re.search(r"/\*[0-9a-f]{4}\*/", line):
            cur["total"] += 1
            for o in OPS:
                if o in line and not (o == "HMMA." and "UTCHMMA" in line):
                    cur[o] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} -- instruction census per kernel (sm_100a)")
    print(f"# {'kernel':70s} {'instr':>7s} " + " ".join(f"{o.rstrip('.'):>8s}" for o in OPS))
    tot = {o: 0 for o in OPS}
    for c in sorted(census, key=lambda c: -c["UTCHMMA"]):
        nm = re.sub(r"\(.*", "", c["name"])[:70]
        print(f"{nm:72s} {c['total']:7d} " + " ".join(f"{c[o]:8d}" for o in OPS))
        for o in OPS:
            tot[o] += c[o]
    print(f"{'# total':72s} {sum(c['total'] for c in census):7d} " + " ".join(f"{tot[o]:8d}" for o in OPS))
    print("# HMMA (legacy mma.sync) must be 0: the tensor-core path is tcgen05 (UTCHMMA) with TMEM accumulators (LDTM/STTM) and bulk copies (UBLKCP).")


if __name__ == "__main__":
    main()

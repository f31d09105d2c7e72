"""Per-kernel counts of the SASS instructions that identify the Blackwell-specific paths
(cuobjdump -sass on the built libb2p.so).  usage: python tools/sass_evidence.py <round tag> > profiles/<tag>_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "paf_baseband2power_b200", "libb2p.so")
KEYS = ["LDG.E.NA.EFL2.256", "LDG.E.NA.128", "UBLKCP", "SYNCS", "LDS.128", "PRMT", "IMAD", "IADD3", "SHFL", "PREEXIT", "ACQBULK",
        "ATOMG", "BAR.SYNC", "LDG.E.STRONG.GPU", "HMMA", "UTCMMA", "UTCHMMA", "UTMALDG"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "rNN"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    print(f"# SASS evidence (cuobjdump -sass paf_baseband2power_b200/libb2p.so), {tag}")
    print("# per kernel: counts of the instructions that identify the Blackwell-specific paths; absent keys are 0")
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print("# arch:", ", ".join(arch))
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k in ("UBLKCP", "SYNCS", "ATOMG", "SHFL", "PRMT", "IMAD", "IADD3") and op.startswith(k)):
                counts[cur][k] += 1
    for fn, c in counts.items():
        print(fn.replace("_ZN59_GLOBAL__N__", "").split("E", 1)[-1] if False else fn)
        print("    " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("# whole library:", "  ".join(f"{k}={tot[k]}" for k in KEYS))


if __name__ == "__main__":
    main()

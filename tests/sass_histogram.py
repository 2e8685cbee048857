"""SASS opcode histogram per kernel of liblr_b200.so (cuobjdump -sass; no GPU needed): the Blackwell-native evidence
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA loads / stores, UBLKCP = cp.async.bulk,
SYNCS = mbarrier) and the absence of the legacy tensor path (HMMA).   python tests/sass_histogram.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "leftrefill_b200", "liblr_b200.so")
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "MUFU", "LDG", "STG",
       "LDS", "STS", "REDG", "ATOMG", "BAR", "UCGABAR_ARV", "UCGABAR_WAIT", "ELECT", "ACQBULK", "FENCE"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name).replace("lr::", ""), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    lines = [f"SASS opcode counts per kernel of liblr_b200.so (sm_100a; cuobjdump -sass, static instruction counts)",
             f"{'kernel':46s} {'total':>6s} " + " ".join(f"{k[:8]:>8s}" for k in KEY)]
    for name, c in kernels.items():
        lines.append(f"{name[:46]:46s} {sum(c.values()):6d} " + " ".join(f"{c.get(k, 0):8d}" for k in KEY))
    has_hmma = sum(c.get("HMMA", 0) for c in kernels.values())
    lines.append(f"legacy HMMA (mma.sync / wmma) instructions in the library: {has_hmma}")
    libs = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    lines.append("linked libraries: " + ", ".join(sorted({l.split()[0] for l in libs.splitlines() if l.strip()})))
    text = "\n".join(lines) + "\n"
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)


if __name__ == "__main__":
    main()

"""SASS opcode evidence per kernel of libswinfuse.so (cuobjdump -sass): which kernels use the Blackwell tensor path
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP / UTMALDG = bulk / tensor TMA), which the legacy one (HMMA = mma.sync).
    python tools/sass_histogram.py > profiles/<tag>_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "swin-unet-image-fusion_b200", "libswinfuse.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "MUFU", "LDGSTS", "SYNCS", "RED", "ATOMG", "MOVM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["TOTAL"] += 1
            for w in WATCH:
                if op.startswith(w):
                    kernels[cur][w] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"{'kernel':90s} {'instr':>6s} " + " ".join(f"{w:>7s}" for w in WATCH))
    for (k, c), name in sorted(zip(kernels.items(), demangled), key=lambda t: t[1]):
        name = re.sub(r"\(.*", "", name).replace("sf::", "")
        if c["TOTAL"] == 0:
            continue
        print(f"{name[:90]:90s} {c['TOTAL']:6d} " + " ".join(f"{c[w]:7d}" if c[w] else f"{'.':>7s}" for w in WATCH))


if __name__ == "__main__":
    main()

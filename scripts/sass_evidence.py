"""Counts the Blackwell-specific SASS mnemonics (tcgen05 MMA / TMEM loads / bulk and tensor-map copies / mbarrier ops /
LDGSTS gathers) per kernel of the tensor-core objects.  CPU only (cuobjdump).

    python scripts/sass_evidence.py > profiles/rNN_sass_mnemonics.txt
"""
import collections
import re
import subprocess
from pathlib import Path

BUILD = Path(__file__).resolve().parent.parent / "fvdb-core_b200" / "build"
OPS = ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "UTCATOMSWS")

for obj in ("conv_tc.cu.o", "conv_tc_wgrad.cu.o"):
    sass = subprocess.run(["cuobjdump", "-sass", str(BUILD / obj)], capture_output=True, text=True, check=True).stdout
    counts, order, fn = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            order.append(fn)
            continue
        for op in OPS:
            if re.search(rf"\b{op}\b|\b{op}\.", line):
                counts[fn][op] += 1
    names = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
    print(f"== {obj}")
    for fn, name in zip(order, names):
        short = re.sub(r"\(.*", "", name).replace("void fvc::", "").replace("fvc::", "")
        if "conv_tc" not in short:
            continue
        print(f"{short:60s} " + "  ".join(f"{op}={counts[fn][op]}" for op in OPS if counts[fn][op]))

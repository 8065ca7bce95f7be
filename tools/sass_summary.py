"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor loads / stores), UBLKCP (bulk copies), SYNCS (mbarrier),
UTCBAR (tcgen05.commit), from `cuobjdump -sass` of the in-tree library.  Runs without a GPU.
Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal_scene_text_recognition_b200", "libmstr_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", ".2CTA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF",
        "SYNCS", "MUFU.TANH", "HMMA", "FFMA2"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k in KEYS:
            if k in line:
                kernels[cur][k] += 1
    names = list(kernels)
    try:
        dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        pass
    total = collections.Counter()
    print("SASS mnemonic counts per kernel of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(LIB, ROOT))
    print("%-92s %s" % ("kernel", " ".join("%9s" % k for k in KEYS)))
    for n, c in kernels.items():
        total.update(c)
        short = re.sub(r"\(.*", "", demangle.get(n, n)).replace("mstr::", "")
        print("%-92s %s" % (short[:92], " ".join("%9d" % c[k] for k in KEYS)))
    print("%-92s %s" % ("TOTAL", " ".join("%9d" % total[k] for k in KEYS)))


if __name__ == "__main__":
    sys.exit(main())

"""Counts the Blackwell-specific SASS mnemonics per kernel of libpdn_b200.so (B200_PROFILING.md "What proves a Blackwell-native
kernel": tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG, tcgen05.commit -> UTCBAR, elect.sync -> ELECT; the
legacy tensor path would show HMMA). Runs without a GPU: `python tools/sass_evidence.py [> profiles/<tag>_sass_evidence.txt]`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "ELECT", "HMMA", "HGMMA")


def evidence(so=None):
    so = so or os.path.join(ROOT, "pydynet_b200", "libpdn_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur and m.group(1) in WANT:
            cnt[cur][m.group(1)] += 1
    names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
    return {re.sub(r"\(.*", "", n): dict(c) for n, c in zip(names, cnt.values())}


if __name__ == "__main__":
    ev = evidence(sys.argv[1] if len(sys.argv) > 1 else None)
    print(f"{'kernel':44s} " + " ".join(f"{w:>8s}" for w in WANT))
    for k in sorted(ev):
        print(f"{k[:44]:44s} " + " ".join(f"{ev[k].get(w, 0):8d}" for w in WANT))

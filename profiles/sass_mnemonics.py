#!/usr/bin/env python
"""Which kernels of libsivae_b200.so actually use the Blackwell units: counts of the SASS mnemonics that prove tcgen05
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit), TMA (UTMALDG / UTMASTG = cp.async.bulk.tensor
load / store), mbarriers (SYNCS) and LDGSTS (cp.async), per kernel template family.  CPU only (cuobjdump on the built
library):   python profiles/sass_mnemonics.py > profiles/rNN_sass_mnemonics.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "soft-intro-vae-pytorch_b200", "libsivae_b200.so")
KEYS = ["UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = {}
    try:
        mangled = sorted(set(re.findall(r"Function : (\S+)", sass)))
        dem = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.splitlines()
        names = dict(zip(mangled, dem))
    except OSError:
        pass
    cnt = collections.defaultdict(collections.Counter)
    variants = collections.Counter()
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            full = names.get(m.group(1), m.group(1))
            full = full.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
            fn = re.sub(r"^void ", "", re.sub(r"\(.*", "", full)).replace("sivae::", "")
            fn = re.sub(r"<.*", "<...>", fn)
            variants[fn] += 1
            continue
        if fn is None:
            continue
        for k in KEYS:
            if re.search(r"\b%s\b|\b%s\." % (k, k), line):
                cnt[fn][k] += 1
    print("| kernel family (template instances) | " + " | ".join(KEYS) + " |\n|---|" + "---:|" * len(KEYS))
    for fn in sorted(cnt, key=lambda f: -cnt[f]["UTCHMMA"] - cnt[f]["UTMALDG"] - cnt[f]["LDGSTS"]):
        if sum(cnt[fn].values()) == 0:
            continue
        print("| `%s` (%d) | " % (fn, variants[fn]) + " | ".join(str(cnt[fn][k]) for k in KEYS) + " |")
    print("\nCounts are static SASS instructions summed over the template instances; %d kernel families (%d instances) in the "
          "library, %d families use tcgen05.mma, %d use TMA loads." % (len(variants), sum(variants.values()),
                                                                      sum(1 for f in cnt if cnt[f]["UTCHMMA"]), sum(1 for f in cnt if cnt[f]["UTMALDG"])))


if __name__ == "__main__":
    main()

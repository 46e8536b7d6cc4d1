#!/usr/bin/env python
"""Static SASS evidence: how often the Blackwell-specific instructions occur in each kernel of the sm_100a objects built by
witw_b200/csrc/Makefile (cuobjdump -sass; runs without a GPU).  UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05 alloc / dealloc / relinquish, UTMALDG = TMA tile load, SYNCS = mbarrier
operations, UCGABAR = cluster barrier, LDGSTS = cp.async, FFMA2 / FADD2 / FMUL2 = packed f32x2 arithmetic, FMNMX3 = 3-input maximum.
usage: python tools/sass_evidence.py > profiles/sass_evidence_<round>.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "SYNCS", "UCGABAR", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "FMNMX3")
LEGACY = ("HGMMA", "HMMA", "IMMA", "DMMA")


def main():
    print(__doc__.split("usage:")[0].strip().replace("\n", "\n# ").join(["# ", ""]))
    legacy = {}
    for obj in sorted(glob.glob(os.path.join(ROOT, "witw_b200", "csrc", "build", "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        legacy[os.path.basename(obj)] = sum(len(re.findall(r"\b%s\b" % k, sass)) for k in LEGACY)
        name = None
        counts = collections.OrderedDict()
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("witw::", "")
                counts[name] = collections.Counter()
                continue
            if name is None:
                continue
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
            if m:
                op = m.group(1)
                if op in KEYS:
                    counts[name][op + (".2CTA" if ".2CTA" in m.group(2) else "")] += 1
        seen = collections.Counter()
        for fn, c in counts.items():
            if any(k.split(".")[0] in ("UTCHMMA", "UTMALDG", "LDTM", "LDGSTS", "FFMA2", "FMNMX3") for k in c):
                key = (fn.split("<")[0], tuple(sorted(c.items())))
                seen[key] += 1
                if seen[key] == 1:          # template instantiations with identical counts are listed once
                    print("%-58s %s" % (fn[:58], "   ".join("%s %d" % kv for kv in sorted(c.items()))))
    print("# legacy tensor-core instructions (%s) per object:" % " / ".join(LEGACY))
    for k, v in legacy.items():
        print("%s: %d" % (k, v))


if __name__ == "__main__":
    main()

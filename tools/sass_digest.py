"""SASS digest of liblr_b200.so: per kernel, the instruction count and the Blackwell-native mnemonics it contains
(UTCHMMA = tcgen05.mma, UTCHMMA.2CTA = cta_group::2, LDTM = tcgen05.ld, UTMALDG = TMA tensor load (+ .MULTICAST),
UTCBAR = tcgen05.commit, SYNCS = mbarrier, ATOMS = shared-memory atomics, HMMA = legacy mma.sync — must be 0).
Usage: python tools/sass_digest.py > profiles/r2/sass_digest.txt   (runs on the build box: cuobjdump needs no GPU)"""
import collections
import hashlib
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lightretriever_b200", "liblr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
MN = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG.2D.MULTICAST", "UTMALDG", "UTCBAR", "SYNCS", "ATOMS", "LDGSTS", "HMMA", "UBLKCP"]
per = collections.OrderedDict()
name = None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        per[name] = collections.Counter()
        continue
    if name and re.match(r"\s+/\*[0-9a-f]+\*/\s", ln):
        per[name]["instructions"] += 1
        for k in MN:
            if re.search(r"\b" + re.escape(k) + r"\b", ln):
                per[name][k] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.basename(LIB)} sha1 {hashlib.sha1(open(LIB, 'rb').read()).hexdigest()}  arch {re.search(r'arch = (\S+)', sass).group(1)}")
print("# UTCHMMA counts include the .2CTA forms; UTMALDG counts include the multicast forms")
for (mangled, c), nice in zip(per.items(), demangle):
    nice = re.sub(r"\(.*", "", nice)
    tags = " ".join(f"{k}={c[k]}" for k in MN if c[k])
    print(f"{c['instructions']:6d}  {nice[:110]:110s} {tags}")

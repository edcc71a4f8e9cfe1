"""cuobjdump -sass of the built library -> tcgen05 / TMA / TMEM instruction counts per kernel (profiles/rNN_sass_tensor_ops.txt)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "makeupdiffuse_b200/libmkd_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b((?:UTCHMMA|UTMALDG|UTMASTG|UTCBAR|LDTM|STTM|UBLKPF|UTMAPF|HMMA)(?:\.\w+)*)")
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern.replace("(anonymous namespace)::", "").replace("void ", ""))
        counts[kern] = collections.Counter()
        continue
    if kern:
        for op in pat.findall(line):
            counts[kern][op] += 1
print(f"# cuobjdump -sass {so}: tcgen05 / TMA / TMEM instruction counts per kernel")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA load, UTMASTG = TMA store, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UBLKPF = bulk L2 prefetch, HMMA = legacy mma.sync")
tot = collections.Counter()
for k, c in counts.items():
    if c:
        print(k)
        print("    " + ", ".join(f"{op} x{n}" for op, n in sorted(c.items())))
        tot.update(c)
print("# library totals: " + ", ".join(f"{op} x{n}" for op, n in sorted(tot.items())))

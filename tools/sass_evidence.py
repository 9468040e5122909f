"""SASS mnemonic counts per kernel of libmadrigal_b200.so -> profiles/r02_sass_evidence.csv (runs without a GPU)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "madrigal_b200", "lib", "libmadrigal_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "STSM", "LDSM", "LDGSTS", "HMMA", "ACQBULK"]
counts, cur, i = collections.OrderedDict(), None, 0
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names[i]; i += 1
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for c in cols:
            if op.startswith(c):
                counts[cur][c] += 1
out = ["# SASS evidence (cuobjdump -sass madrigal_b200/lib/libmadrigal_b200.so; tools/sass_evidence.py): instructions per kernel",
       "# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UTCBAR = tcgen05.commit, "
       "SYNCS = mbarrier, STSM/LDSM = stmatrix/ldmatrix, LDGSTS = cp.async, HMMA = warp-level mma.sync (per-(drug, head) attention only)",
       "kernel," + ",".join(cols)]
for k, c in counts.items():
    if not k.startswith("mdg::") and "mdg::" not in k:
        continue
    if sum(c.values()) == 0:
        continue
    short = re.sub(r"\(.*", "", k.replace("void ", ""))
    out.append('"%s",' % short + ",".join(str(c[x]) for x in cols))
open(os.path.join(ROOT, "profiles", "r02_sass_evidence.csv"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))

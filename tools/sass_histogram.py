"""SASS opcode histogram of the built library (cuobjdump -sass): python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt
Proves which Blackwell paths are in the binary: UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store
(.MULTICAST = cluster multicast), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, HMMA = mma.sync."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "cpp-paddle-ocr_b200", "libb200ocr.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
total = collections.Counter()
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        total[base] += 1
        per[cur][base] += 1
        if ".MULTICAST" in op:
            total[base + ".MULTICAST"] += 1
            per[cur][base + ".MULTICAST"] += 1
print(f"# SASS opcode histogram of cpp-paddle-ocr_b200/libb200ocr.so (cuobjdump -sass, sm_100a); tools/sass_histogram.py")
print(f"# total instructions: {sum(v for k, v in total.items() if '.' not in k)} in {len(per)} kernels\n")
print("## Blackwell-specific / notable opcodes (whole library)")
for op in ["UTCHMMA", "UTMALDG", "UTMALDG.MULTICAST", "UTMASTG", "LDTM", "UTCBAR", "UTCBAR.MULTICAST", "UCGABAR_ARV", "UTCATOMSWS", "SYNCS",
           "HMMA", "FFMA", "HFMA2", "FMUL", "MUFU", "LDG", "STG", "LDS", "STS", "LDGSTS", "ATOMG", "BAR", "SHFL", "ACQBULK", "CCTL"]:
    print(f"{total.get(op, 0):8d}  {op}")
print("\n## kernels that issue tcgen05 MMA (UTCHMMA), TMA loads (UTMALDG, of which multicast), TMEM loads (LDTM), mma.sync (HMMA)")
for k, c in per.items():
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"] or c["HMMA"]:
        name = re.sub(r"^_ZN\d+b200ocr\d+_GLOBAL__N__[0-9a-f_]+", "", k)[:110]
        print(f"UTCHMMA {c['UTCHMMA']:3d}  UTMALDG {c['UTMALDG']:3d} (mc {c['UTMALDG.MULTICAST']:2d})  UTMASTG {c['UTMASTG']:2d}  LDTM {c['LDTM']:3d}  "
              f"HMMA {c['HMMA']:4d}  total {sum(v for kk, v in c.items() if '.' not in kk):6d}  {name}")

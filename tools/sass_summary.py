"""cuobjdump -sass of the built library -> profiles/sass_conv3_tc_rNN.txt: per conv3_tc_kernel variant a histogram of the
Blackwell-specific SASS mnemonics (tcgen05 MMA / TMEM / TMA / mbarrier), then the full listing (encodings stripped) of the
variant that carries most of the FLOPs.  No GPU needed.  usage: python tools/sass_summary.py [out] [variant-substring]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "deepwmh_b200", "lib", "libdeepwmh_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_conv3_tc_r02.txt")
pick = sys.argv[2] if len(sys.argv) > 2 else "conv3_tc_kernel<__half, 2, true, false, true, false, false>"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UBLKCP", "SYNCS", "ELECT", "FENCE", "LDS", "STS", "LDG", "STG", "SHFL", "BAR"]

txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = re.search(r"arch = (\S+)", txt)
funcs = re.split(r"\n\s*Function : ", txt)[1:]
rows, listing = [], None
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    if "conv3_tc_kernel" not in name:
        continue
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = dem.replace("void dwmh::", "").split("(")[0]
    ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)
    hist = collections.Counter(i.split(".")[0] for i in ins)
    rows.append((dem, len(ins), hist))
    if pick in dem:
        lines = []
        for ln in f.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*.*\*/\s*$", ln)
            if m:
                lines.append("%s  %s" % (m.group(1), m.group(2).rstrip()))
        listing = (dem, lines)
with open(out, "w") as fo:
    fo.write("# cuobjdump -sass deepwmh_b200/lib/libdeepwmh_b200.so (%s), tools/sass_summary.py\n" % (arch.group(1) if arch else "?"))
    fo.write("# SASS mnemonics that prove the Blackwell path: UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st (TMEM),\n")
    fo.write("# UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.\n")
    fo.write("%-64s %7s " % ("variant <T, KSTEPS, SMALL_CB, TCONV, DUAL, XFORM, FIRST>", "instrs") + " ".join("%7s" % k[:7] for k in KEYS) + "\n")
    for dem, n, hist in rows:
        fo.write("%-64s %7d " % (dem[:64], n) + " ".join("%7d" % hist.get(k, 0) for k in KEYS) + "\n")
    if listing:
        fo.write("\n# full listing of %s (%d instructions; encodings stripped)\n" % (listing[0], len(listing[1])))
        fo.write("\n".join(listing[1]) + "\n")
print(out, len(rows), "variants")

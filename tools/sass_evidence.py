"""Counts the Blackwell-native SASS mnemonics per kernel of the built library (cuobjdump -sass) -> profiles/r02_sass_evidence.txt.
UTC*MMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG / UBLKCP = TMA tensor / bulk copies,
SYNCS = mbarrier, LDGSTS = cp.async, HMMA = mma.sync (f16 / bf16 / tf32 forms listed separately), HFMA2.BF16 / MUFU.TANH.BF16 = the
packed activation arithmetic."""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "se3conv3d_b200", "lib", "libse3conv3d_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCQMMA", r"\bUTCQMMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
        ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"),
        ("HMMA.16816.F32.BF16", r"HMMA\.16816\.F32\.BF16"), ("HMMA.16816.F32(f16)", r"HMMA\.16816\.F32 "),
        ("HMMA.1688.TF32", r"HMMA\.1688\.F32\.TF32"), ("HFMA2.BF16", r"HFMA2\.BF16"), ("HMUL2.BF16", r"HMUL2\.BF16"),
        ("MUFU.TANH.BF16", r"MUFU\.TANH\.BF16"), ("MUFU.TANH", r"MUFU\.TANH "), ("LDSM", r"\bLDSM")]
kern, cur = OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = Counter()
        continue
    if cur is None:
        continue
    for name, pat in KEYS:
        if re.search(pat, line):
            kern[cur][name] += 1
names = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
want = re.compile(sys.argv[1] if len(sys.argv) > 1 else r"k_gemm|k_agg|k_edge|k_conv_fused|k_dx_segsum")
lines = ["# SASS evidence (tools/sass_evidence.py: cuobjdump -sass se3conv3d_b200/lib/libse3conv3d_b200.so, build of this commit)",
         "# per kernel: static count of the Blackwell-native mnemonics (UTC*MMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld,",
         "# UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDGSTS = cp.async, HMMA = mma.sync)", ""]
for (mangled, cnt), nm in sorted(zip(kern.items(), names), key=lambda kv: kv[1]):
    if not want.search(nm) or not cnt:
        continue
    lines.append("%-118s %s" % (nm[:118], " ".join("%s=%d" % (k, cnt[k]) for k, _ in KEYS if cnt[k])))
open(os.path.join(ROOT, "profiles", "r02_sass_evidence.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:4] + [l for l in lines if "k_gemm_tma" in l or "k_agg_tc<32, 2, false, 2>" in l or "k_edge_row_tc<32, 2, 2, false>" in l]))

"""Per-source-line warp-instruction counts and stall-sample shares from an ncu report captured with
--import-source on.  Usage: python tools/ncu_lines.py report.ncu-rep file.cuh [divisor] [min_per_unit]"""
import collections, csv, io, os, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 3.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
cur = hdr = None
agg, st = collections.Counter(), collections.Counter()
def num(s):
    try: return int(s)
    except ValueError: return 0
for r in csv.reader(io.StringIO(raw)):
    if len(r) == 2 and r[0] == "File Path":
        cur, hdr = r[1].split('/')[-1], None
        continue
    if len(r) == 2: continue
    if hdr is None:
        if "Instructions Executed" in r: hdr = {k: i for i, k in enumerate(r)}
        continue
    if r[hdr["Line No"]]:
        key = (cur, num(r[hdr["Line No"]]))
        agg[key] += num(r[hdr["Instructions Executed"]])
        st[key] += num(r[hdr["# Samples"]])
byfile = collections.Counter()
for (f, l), n in agg.items(): byfile[f] += n
print("per unit:", {k: round(v / div, 1) for k, v in byfile.most_common()}, "total", round(sum(byfile.values()) / div, 1))
tot = max(sum(st.values()), 1)
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-attention-visual-odometry_b200", "csrc")
src = open(os.path.join(root, fname)).read().split('\n')
for (f, l), n in sorted(agg.items()):
    if f == fname and (n / div >= thr or st[(f, l)] / tot > 0.004):
        print("%4d %8.1f %5.1f%%  %s" % (l, n / div, 100 * st[(f, l)] / tot, src[l - 1][:105] if l - 1 < len(src) else ""))

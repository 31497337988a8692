"""Top SASS lines by stall samples from an .ncu-rep source page: python scripts/ncu_hot.py rep kernel-regex [n]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several launches may match: take the first block
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hdr_i[0]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
hdr = rows[start]
body = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
si = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
total = sum(int(r[si] or 0) for r in body)
print("kernel", rows[start - 1][1][:80], "total samples", total, "instructions", len(body))
agg = {}
for r in body:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("stall mix:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(body, key=lambda r: -int(r[si] or 0))[:n]:
    top = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print("%6d %5.1f%%  %-70s %s" % (int(r[si]), 100.0 * int(r[si]) / max(total, 1), r[1].strip()[:70],
                                     " ".join("%s=%d" % (h[6:], v) for v, h in top if v)))

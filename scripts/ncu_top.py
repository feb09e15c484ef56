"""Top stall lines of an ncu report (source page, SASS view):  ncu_top.py <report.ncu-rep> [N]"""
import csv, subprocess, sys
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?", " total samples:", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for idx, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][col["# Samples"]] or 0))[:n]:
    s = int(r[col["# Samples"]] or 0)
    top = sorted(((int(r[col[x]] or 0), x) for x in stalls), reverse=True)[:2]
    print("%5d  %5.1f%%  #%4d  %-70s %s" % (s, 100.0 * s / max(tot, 1), idx, r[col["Source"]].strip()[:70], top))

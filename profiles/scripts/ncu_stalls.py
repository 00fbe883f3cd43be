"""Measurement tooling: the instructions of an ncu --set full --import-source capture where warps spend their samples.
    python profiles/scripts/ncu_stalls.py x.ncu-rep [top]
Prints the SASS instructions with the most warp-stall samples and the dominant stall reason of each."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
total = sum(int(r[col["# Samples"]] or 0) for r in body)
print(f"total samples {total}")
agg = {}
for h in reasons:
    agg[h] = sum(int(r[col[h]] or 0) for r in body)
print("by reason:", ", ".join(f"{h[6:]} {100 * v / max(total, 1):.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
body.sort(key=lambda r: -int(r[col["# Samples"]] or 0))
for r in body[:top]:
    n = int(r[col["# Samples"]] or 0)
    why = max(reasons, key=lambda h: int(r[col[h]] or 0))
    print(f"{100 * n / max(total, 1):5.1f}%  {why[6:]:14s} {r[col['Source']][:110]}")

"""Measurement tooling: per-kernel DRAM traffic per launch from an ncu launch list (one training step).
    python profiles/scripts/dram_traffic_from_launches.py <launches.csv> <workload> <batch> <note> > profiles/r2_dram_traffic.json
The CSV is what profiles/scripts/r2_call8.sh / r2_final.sh produce (columns kernel, time_ns, dram_read_bytes,
dram_write_bytes); keys of the output are the kernel names of bench.py's roofline table."""
import csv
import json
import sys

NAMES = [("attn_bwd_kernel", "attn_bwd_kernel"), ("attn_fwd2_kernel", "attn_fwd2_kernel"), ("attn_delta_kernel", "attn_delta_kernel"),
         ("ln_bwd", "ln_backward"), ("ln_fwd", "ln_forward"), ("head_fwd", "head_forward"), ("head_bwd", "head_backward"),
         ("approx_ndcg_kernel", "arb_approx_ndcg"), ("pack_rows_kernel", "pack_rows")]

path, workload, batch, note = sys.argv[1:5]
acc = {}
for row in csv.DictReader(l for l in open(path) if not l.startswith("#")):
    for needle, key in NAMES:
        if needle in row["kernel"]:
            t = acc.setdefault(key, [0.0, 0])
            t[0] += float(row["dram_read_bytes"]) + float(row["dram_write_bytes"])
            t[1] += 1
            break
out = {"note": note, workload: {str(int(batch)): {k: int(v[0] / v[1]) for k, v in acc.items()}}}
print(json.dumps(out, indent=1))

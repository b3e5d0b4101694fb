"""Aggregate the warp-stall samples of `ncu --page source --csv` over windows of 100 SASS instructions (the roles of
a warp-specialised kernel occupy disjoint address ranges).   usage: ncu_source_regions.py src.csv [window]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
W = int(sys.argv[2]) if len(sys.argv) > 2 else 100
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
for s in range(0, len(body), W):
    chunk = body[s:s + W]
    n = sum(int(r[ix["# Samples"]] or 0) for r in chunk)
    ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in chunk)
    agg = {h: sum(int(r[ix[h]] or 0) for r in chunk) for h in stalls}
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:4]
    first = chunk[0][ix["Source"]][:40]
    print(f"[{s:5d}] samples {n:7d} ({100*n/tot:5.1f}%) inst_exec {ex:9d}  {' '.join(f'{k[6:]}={v}' for k, v in top if v)}   | {first}")

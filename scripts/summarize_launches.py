"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time and share of the
LAST step (launch lists are cold-cache and serialised: compare shares, not absolutes)."""
import collections, csv, re, sys
path, steps_total = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
n = len(rows)
sub = rows[-(n // steps_total):]
# bench.py flushes L2 (one uint8 fill kernel) at the start of every TIMED step: when the marker is there, the step = the
# launches from the last flush up to the start of the next forward (its stem kernel) — robust against the extra forwards
# bench.py runs outside the timed region
marks = [i for i, r in enumerate(rows) if "FillFunctor<unsigned char>" in r["Kernel Name"]]
if marks:
    stems = [i for i in range(marks[-1], n) if "stem_im2col" in rows[i]["Kernel Name"]]
    sub = rows[marks[-1]:(stems[1] if len(stems) > 1 else n)]
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in sub:
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    v *= {"us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(u, 1.0)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)[:78]
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print(f"# {path}: {len(sub)} launches in the last timed step, {T:.3f} ms serialised")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:30]:
    print(f"{v:9.3f} ms {100 * v / T:5.1f}%  x{cnt[k]:4d}  {k}")

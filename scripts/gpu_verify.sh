# Full verification of the committed state without profilers.  usage: gpu_verify.sh <tag>
T=gpurun_out/$1
mkdir -p $T
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $T/pytest_gpu.txt
timeout 600 python __graft_entry__.py smoke > $T/smoke.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > $T/bench.json 2> $T/bench.err
cat $T/pytest_gpu.txt; tail -1 $T/smoke.txt; python - $T/bench.json <<'P'
import json, sys
j = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], "roof", j["roofline"]["frac"], "tensor", j["roofline_tensor"].get("frac"), "clocks", j["clocks"])
P
tail -2 $T/bench.err

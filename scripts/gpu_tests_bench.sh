# selected GPU tests + the full bench line.  usage: gpu_tests_bench.sh <tag> "<pytest -k expr>"
T=gpurun_out/$1
mkdir -p $T
timeout 1500 python -m pytest tests -m gpu -q -x -k "$2" 2>&1 | tail -40 > $T/pytest_sel.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $T/bench.json 2> $T/bench.err
cat $T/pytest_sel.txt; cat $T/bench.json | cut -c1-300; python - $T/bench.json <<'P'
import json, sys
j = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value", j["value"], "e2e", j["e2e"]["value"], "launches", j["gpu_launches"], "full_model", j.get("full_model"))
P
tail -3 $T/bench.err

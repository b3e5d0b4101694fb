# usage: gpu_r2s.sh <tag>  : all GPU tests, bench, PlaneTRHead launch list
T=gpurun_out/$1
mkdir -p $T
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $T/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $T/bench.json 2> $T/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $T/planetr_launches.csv python scripts/profile_planetr.py 64 > $T/planetr.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $T/launches.csv python bench.py --steps 1 --warmup 3 --only-value > $T/ncu_bench.log 2>&1
cat $T/pytest_gpu.txt | tail -8; python - $T/bench.json <<'P'
import json, sys
j = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], "launches", j["gpu_launches"], "roof", j["roofline"]["frac"], "full_model", j.get("full_model"))
P
tail -3 $T/bench.err; tail -2 $T/planetr.log

# quick iteration: selected GPU tests + bench value only.  usage: gpu_quick.sh <tag> "<pytest -k expr>"
set -x
T=gpurun_out/$1
mkdir -p $T
timeout 900 python -m pytest tests -m gpu -q -x -k "$2" 2>&1 | tail -15 > $T/pytest_sel.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $T/launches.csv python bench.py --steps 1 --warmup 3 --only-value > $T/ncu_bench.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --only-value > $T/bench_value.json 2> $T/bench.err
cat $T/pytest_sel.txt; cat $T/bench_value.json; tail -3 $T/bench.err

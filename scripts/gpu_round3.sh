# tests + bench + planes launch list.  usage: gpu_round3.sh <tag>
set -x
T=gpurun_out/$1
mkdir -p $T
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $T/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $T/bench.json 2> $T/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $T/planes_launches.csv python scripts/planes_bench.py > $T/planes_bench.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $T/launches.csv python bench.py --steps 1 --warmup 3 --only-value > $T/ncu_bench.log 2>&1
tail -5 $T/pytest_gpu.txt; cat $T/bench.json | cut -c1-600; tail -3 $T/bench.err

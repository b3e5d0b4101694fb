# One gpurun call for a full evidence round.  usage: gpu_round.sh <tag>   (outputs under gpurun_out/<tag>/)
set -x
T=gpurun_out/$1
mkdir -p $T
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $T/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $T/bench.json 2> $T/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $T/bench_reference.json 2>> $T/bench.err
timeout 300 python scripts/score_bench.py > $T/score_bench.jsonl 2> $T/score_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $T/launches.csv python bench.py --steps 1 --warmup 3 --only-value > $T/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 1 -c 1 -f -o $T/score_tc python scripts/profile_score.py 512 256 2 > $T/ncu_score.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $T/score_launches.csv python scripts/profile_score.py 512 256 3 > /dev/null 2>&1
timeout 120 python scripts/score_trace.py 512 256 0 > $T/score_timeline_cta0.txt 2>&1
tail -3 $T/pytest_gpu.txt; cat $T/bench.json; cat $T/bench_reference.json; head -1 $T/score_bench.jsonl; tail -3 $T/bench.err

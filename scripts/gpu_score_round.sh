# one gpurun call for a scoring-kernel iteration: parity tests, timing, launch list, ncu --set full.  usage: gpu_score_round.sh <tag>
set -x
T=gpurun_out/$1
mkdir -p $T
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $T/pytest_gpu.txt
timeout 300 python scripts/score_bench.py > $T/score_bench.jsonl 2> $T/score_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $T/score_launches.csv python scripts/profile_score.py 512 256 3 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 1 -c 1 -f -o $T/score_tc python scripts/profile_score.py 512 256 2 > $T/ncu_score.log 2>&1
tail -4 $T/pytest_gpu.txt; cat $T/score_bench.jsonl; tail -3 $T/score_bench.err

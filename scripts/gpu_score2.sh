# Scoring-kernel iteration: parity subset + sweep + timeline (+ optional extras).  usage: gpu_score2.sh <tag>
set -x
T=gpurun_out/$1
mkdir -p $T
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config2.py tests/test_gpu_runtime_graph.py -m gpu -q -x 2>&1 | tail -8 > $T/pytest_score.txt
timeout 300 python scripts/score_bench.py > $T/score_bench.jsonl 2> $T/score_bench.err
timeout 120 python scripts/score_trace.py 512 256 0 > $T/score_timeline_cta0.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $T/score_launches.csv python scripts/profile_score.py 512 256 3 > /dev/null 2>&1
cat $T/pytest_score.txt; cat $T/score_bench.jsonl; tail -3 $T/score_bench.err

# quick scoring iteration: parity tests of the scoring path + graph-timed roofline config + timeline.  usage: gpu_score_quick.sh <tag>
T=gpurun_out/$1
mkdir -p $T
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > $T/pytest_gpu.txt
timeout 300 python scripts/score_bench.py roofline > $T/score_bench.jsonl 2> $T/score_bench.err
timeout 120 python scripts/score_trace.py 512 256 0 > $T/trace_0.txt 2>&1
cat $T/pytest_gpu.txt; cat $T/score_bench.jsonl; tail -2 $T/score_bench.err

set -x
mkdir -p gpurun_out/r1e
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1e/pytest_gpu.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r1e/bench.json 2> gpurun_out/r1e/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1e/bench_reference.json 2>> gpurun_out/r1e/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1e/launches.csv python bench.py --steps 1 --warmup 3 --only-value > gpurun_out/r1e/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 1 -c 1 -o gpurun_out/r1e/score_tc python scripts/profile_score.py 512 256 2 > gpurun_out/r1e/ncu_score.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1e/score_launches.csv python scripts/profile_score.py 512 256 3 > /dev/null 2>&1
tail -3 gpurun_out/r1e/pytest_gpu.txt; cat gpurun_out/r1e/bench.json

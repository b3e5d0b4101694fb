# One gpurun call for a full evidence round (round 2).  usage: gpu_round2.sh <tag> [quick|full]   (outputs under gpurun_out/<tag>/)
set -x
T=gpurun_out/$1
mkdir -p $T
nvidia-smi -L > $T/gpus.txt
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $T/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $T/bench.json 2> $T/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $T/launches.csv python bench.py --steps 1 --warmup 3 --only-value > $T/ncu_bench.log 2>&1
if [ "$2" == "full" ]; then
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $T/bench_reference.json 2>> $T/bench.err
for s in k7 conv256 gnn; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 3 -c 1 -f -o $T/gemm_$s python scripts/profile_gemm.py $s 1 > $T/ncu_gemm_$s.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 1 -c 1 -f -o $T/score_tc python scripts/profile_score.py 512 256 2 > $T/ncu_score.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $T/score_launches.csv python scripts/profile_score.py 512 256 3 > /dev/null 2>&1
timeout 120 python scripts/score_trace.py 512 256 0 > $T/score_timeline_cta0.txt 2>&1
fi
tail -5 $T/pytest_gpu.txt; cat $T/bench.json; tail -3 $T/bench.err

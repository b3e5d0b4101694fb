# Final evidence run of a round.  usage: gpu_final.sh <tag>   (outputs under gpurun_out/<tag>/)
T=gpurun_out/$1
mkdir -p $T
nvidia-smi -L > $T/gpus.txt
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $T/pytest_gpu.txt
timeout 600 python __graft_entry__.py smoke > $T/smoke.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > $T/bench.json 2> $T/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $T/bench_reference.json 2>> $T/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $T/launches.csv python bench.py --steps 1 --warmup 3 --only-value > $T/ncu_bench.log 2>&1
timeout 300 python scripts/profile_gemm.py all 10 > $T/gemm_shapes.jsonl 2> $T/gemm_shapes.err
for s in k7 conv256 gnn res; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 3 -c 1 -f -o $T/gemm_$s python scripts/profile_gemm.py $s 1 > $T/ncu_gemm_$s.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 1 -c 1 -f -o $T/score_tc python scripts/profile_score.py 512 256 2 > $T/ncu_score.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $T/score_launches.csv python scripts/profile_score.py 512 256 3 > /dev/null 2>&1
timeout 120 python scripts/score_trace.py 512 256 0 > $T/score_timeline_cta0.txt 2>&1
timeout 300 python scripts/score_bench.py > $T/score_bench.jsonl 2> $T/score_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $T/planetr_launches.csv python scripts/profile_planetr.py 64 > $T/planetr.log 2>&1
tail -5 $T/pytest_gpu.txt; tail -2 $T/smoke.txt; cat $T/bench.json | cut -c1-400; cat $T/bench_reference.json | cut -c1-300; cat $T/gemm_shapes.jsonl; tail -3 $T/bench.err

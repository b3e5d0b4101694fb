# GEMM-engine evidence (shapes timed with CUDA events + ncu --set full captures) and the PlaneTRHead launch list.  usage: <tag>
T=gpurun_out/$1
mkdir -p $T
timeout 900 python -m pytest tests -m gpu -q -x -k "planetr or stage_entry" 2>&1 | tail -5 > $T/pytest_planetr.txt
timeout 300 python scripts/profile_gemm.py all 10 > $T/gemm_shapes.jsonl 2> $T/gemm_shapes.err
for s in k7 conv256 gnn res; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 3 -c 1 -f -o $T/gemm_$s python scripts/profile_gemm.py $s 1 > $T/ncu_gemm_$s.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $T/planetr_launches.csv python scripts/profile_planetr.py 64 > $T/planetr.log 2>&1
timeout 300 python - > $T/planetr_time.txt 2>&1 <<'P'
import sys, torch
sys.argv = ["x", "64"]
exec(open("scripts/profile_planetr.py").read())
P
cat $T/pytest_planetr.txt; cat $T/gemm_shapes.jsonl; tail -2 $T/gemm_shapes.err; tail -1 $T/planetr_time.txt

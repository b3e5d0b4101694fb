# plane-list tests + timing + launch metrics after a change to planes.cu.  usage: gpu_planes_verify.sh <tag>
T=gpurun_out/$1
mkdir -p $T
timeout 600 python -m pytest tests -m gpu -q -x -k "planes or planetr or smoke" 2>&1 | tail -4 > $T/pytest_planes.txt
timeout 120 python scripts/planes_bench.py 64 128 > $T/planes_bench.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:plane_argmax -s 2 -c 1 python scripts/planes_bench.py 128 > $T/ncu_planes.log 2>&1
cat $T/pytest_planes.txt; tail -2 $T/planes_bench.txt; grep -E "gpu__time|inst_executed" $T/ncu_planes.log

# ncu --set full capture of plane_argmax_kernel (row f1 post-processing) on 128 images.  usage: gpu_planes_ncu.sh <tag>
T=gpurun_out/$1
mkdir -p $T
timeout 600 ncu --set full --clock-control none --import-source on -k regex:plane_argmax -s 2 -c 1 -f -o $T/plane_argmax python scripts/planes_bench.py 128 > $T/ncu_planes.log 2>&1
timeout 120 python scripts/planes_bench.py 64 128 > $T/planes_bench.txt 2>&1
tail -3 $T/planes_bench.txt; tail -2 $T/ncu_planes.log

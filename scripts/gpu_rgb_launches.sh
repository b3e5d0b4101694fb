# Launch list of the RGB -> backbone -> head step (side measurement of bench.py).  usage: gpu_rgb_launches.sh <tag>
T=gpurun_out/$1
mkdir -p $T
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $T/rgb_launches.csv python bench.py --side from_rgb --pairs 64 > $T/rgb_side.log 2>&1
tail -2 $T/rgb_side.log

# multi-GPU check.  usage: gpu_multi.sh <tag> <N>
T=gpurun_out/$1
N=$2
mkdir -p $T
nvidia-smi -L > $T/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x 2>&1 | tail -15 > $T/pytest_multigpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $T/bench_n$N.json 2> $T/bench_n$N.err
cat $T/pytest_multigpu.txt; cat $T/bench_n$N.json | cut -c1-1500; tail -5 $T/bench_n$N.err

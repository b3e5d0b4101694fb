# selected GPU tests only.  usage: gpu_tests.sh <tag> "<pytest -k expr>"
T=gpurun_out/$1
mkdir -p $T
timeout 1500 python -m pytest tests -m gpu -q -x -k "$2" 2>&1 | tail -40 > $T/pytest_sel.txt
cat $T/pytest_sel.txt

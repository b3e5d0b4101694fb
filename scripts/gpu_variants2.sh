# A/B of scoring-kernel library variants (interleaved, graph replays).  usage: gpu_variants2.sh <tag> <variant names...>
T=gpurun_out/$1; shift
mkdir -p $T
for rep in 1 2; do
python scripts/score_quick.py >> $T/variants.jsonl 2>> $T/variants.err
for v in "$@"; do
  NSAC_B200_LIB=build/variants/$v.so timeout 120 python scripts/score_quick.py >> $T/variants.jsonl 2>> $T/variants.err
done
done
NSAC_B200_LIB=build/variants/$1.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config2.py -m gpu -q -x -k "score or scoring" 2>&1 | tail -3 >> $T/variants.err
cat $T/variants.jsonl; tail -4 $T/variants.err

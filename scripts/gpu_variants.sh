T=gpurun_out/$1
mkdir -p $T
python scripts/score_quick.py >> $T/variants.jsonl 2>> $T/variants.err
for v in abl1 abl2 abl3 abl4 abl5; do
  NSAC_B200_LIB=build/variants/$v.so timeout 120 python scripts/score_quick.py >> $T/variants.jsonl 2>> $T/variants.err
done
cat $T/variants.jsonl; tail -2 $T/variants.err

#!/bin/bash
# A/B of two builds of the library on the bench shapes (1 GPU): default vs MMR_B200_LIB=<variant>.
#   gpurun --timeout 900 -- 'bash scripts/ab_epi.sh libmmr_b200_nosplit.so tag'
alts=$1; tag=${2:-ab}; out=gpurun_out; mkdir -p $out   # alts: comma-separated variant libraries
pkg=multi_modal_retrieval_predict_project_b200
common="--steps 20 --warmup 5 --no-extra --no-cpu-baseline --latency-queries 0"
for shape in ${SHAPES:-1250000:4096 5000000:4096 10000000:4096}; do
  rows=${shape%%:*}; batch=${shape##*:}
  for lib in default ${alts//,/ }; do
    if [ "$lib" = default ]; then unset MMR_B200_LIB; else export MMR_B200_LIB=$PWD/$pkg/$lib; fi
    python bench.py $common --rows $rows --batch $batch > $out/${tag}_${rows}_${lib%.so}.json 2> $out/${tag}_${rows}_${lib%.so}.err
    echo "rows=$rows batch=$batch lib=$lib exit=$?"
    python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_${rows}_${lib%.so}.json"))
    r = d["roofline"]
    print("   q/s %.0f  ms/step %.3f  kernel_ms %.3f  TFLOP/s %.0f  clocks %s  parity %s" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["achieved"], d["clocks"], d.get("parity_check", {}).get("ok")))
except Exception as e:
    print("   no line:", e)
PY
  done
done

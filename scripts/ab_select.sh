#!/bin/bash
# Per-launch duration of the selection kernel for several builds of the library (ncu launch list; run the plain loop first).
#   gpurun -- 'bash scripts/ab_select.sh plain "default selcta"; bash scripts/ab_select.sh ncu "default selcta"'
mode=$1; libs=${2:-default}; P=$PWD/multi_modal_retrieval_predict_project_b200
loop() {
  for rows in ${ROWS_LIST:-1250000 10000000}; do for lib in $libs; do
    if [ $lib = default ]; then unset MMR_B200_LIB; else export MMR_B200_LIB=$P/libmmr_b200_$lib.so; fi
    echo "== rows=$rows lib=$lib"; ROWS=$rows python scripts/diag_tail.py 2>&1 | tail -1
  done; done
}
if [ "$mode" = plain ]; then loop; else
  ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -k regex:select --csv \
      --log-file gpurun_out/ab_select.csv bash scripts/ab_select.sh plain "$libs" > gpurun_out/ab_select.log 2>&1
  echo "ncu exit=$?"
  python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/ab_select.csv")) if len(r) > 10]
h = rows[0]; ix = {n: i for i, n in enumerate(h)}
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        key = (r[ix["Process ID"]], r[ix["Kernel Name"]][:60]); v = float(r[ix["Metric Value"]].replace(",", ""))
    except Exception:
        continue
    agg.setdefault(key, []).append(v)
for (pid, name), v in agg.items():
    v = sorted(v); print(pid, name, "n=%d min %.1f p50 %.1f max %.1f %s" % (len(v), v[0], v[len(v) // 2], v[-1], rows[1][ix["Metric Unit"]]))
PY
fi

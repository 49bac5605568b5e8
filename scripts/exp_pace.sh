#!/bin/bash
# Pacing window x kernel variant at the headline shape: kernel time (CUDA events, 10 launches) and DRAM bytes /
# L2 hit rate of one launch (ncu metrics only).  Needs the -DMMR_DIAG library.
#   gpurun --timeout 1200 -- 'bash scripts/exp_pace.sh'
pkg=$PWD/multi_modal_retrieval_predict_project_b200
out=gpurun_out; mkdir -p $out
export MMR_B200_LIB=$pkg/libmmr_b200_diag.so NO_TRACE=1 ROWS=${ROWS:-10000000}
for variant in ${VARIANTS:-long short}; do
  for pace in ${PACES:-0 16 32 64}; do
    export VARIANT=$variant MMR_B200_GEMM_PACE=$pace
    echo "=== variant=$variant pace=$pace"
    python scripts/diag_gemm.py 2>&1 | tail -1 && \
    ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
        --clock-control none -k regex:gemm_topk_kernel --launch-skip 3 -c 1 --csv --log-file $out/pace_${variant}_${pace}.csv \
        python scripts/diag_gemm.py > /dev/null 2>&1
    grep -E "dram__bytes_read|hit_rate|time_duration|tensor" $out/pace_${variant}_${pace}.csv | awk -F'","' '{print "   " $(NF-2) " = " $NF " " $(NF-1)}'
  done
done

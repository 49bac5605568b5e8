#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/diag_gemm.sh'   (diag libraries built here first)
pkg=$PWD/multi_modal_retrieval_predict_project_b200
out=gpurun_out; mkdir -p $out
run() {  # lib debug rows trace?
  if [ -n "$4" ]; then
    MMR_B200_LIB=$pkg/$1 MMR_B200_GEMM_DEBUG=$2 ROWS=$3 MMR_B200_GEMM_TRACE=/tmp/gt_$$.bin python scripts/diag_gemm.py
  else
    MMR_B200_LIB=$pkg/$1 MMR_B200_GEMM_DEBUG=$2 ROWS=$3 python scripts/diag_gemm.py
  fi
}
for lib in ${LIBS:-libmmr_b200_diag.so libmmr_b200_diag_nosplit.so}; do
  echo "=== $lib"
  for d in ${DEBUGS:-0}; do run $lib $d ${ROWS:-1250000} ${TRACE:-}; done
done

#!/bin/bash
# Round-2 `ncu --set full` captures (run on the GPU box through gpurun AFTER the same programs have exited 0 without ncu).
# Raw pages land in gpurun_out/ncu_raw_<kernel>_r2.csv; copy the ones to keep into profiles/.
set -u
mkdir -p gpurun_out
cap() {   # name, kernel regex, skip, count, program...
  local name=$1 k=$2 s=$3 c=$4; shift 4
  timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:$k -s $s -c $c -o gpurun_out/_r2_$name -f "$@" > gpurun_out/_r2_$name.log 2>&1
  ncu -i gpurun_out/_r2_$name.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${name}_r2.csv 2>/dev/null
  rm -f gpurun_out/_r2_$name.ncu-rep
}
python tools/one_call.py s3batch 140 8 > /dev/null 2>&1 || { echo "plain s3batch run failed"; exit 1; }
python tools/ncu_t3_rows.py 16 > /dev/null 2>&1 || { echo "plain t3 run failed"; exit 1; }
cap cfm_tail cfm_tail_kernel 300 2 python tools/one_call.py s3batch 140 8
cap attn_fa attn_fa_kernel 300 2 python tools/one_call.py s3batch 140 8
cap gemm_tc gemm_tc_kernel 200 3 python tools/one_call.py s3batch 140 8
cap gemv_tc gemv_tc_kernel 4 4 python tools/ncu_t3_rows.py 16
cap gemv_2rows gemv_kernel 4 4 python tools/ncu_t3_rows.py 1
cap decode_attn decode_attn_kernel 2 2 python tools/ncu_t3_rows.py 16
ls -la gpurun_out/ncu_raw_*_r2.csv

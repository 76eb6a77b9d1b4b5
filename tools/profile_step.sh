#!/bin/bash
# ncu evidence for one benchmark step (run under gpurun, ONE GPU): launch list + `--set full` captures of the step's kernels.
#   tools/profile_step.sh <tag>     -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_{gru_if1,gru_if4,gru_fused,gru_seq,layout,viterbi,gemm,conv}.ncu-rep
# A number printed by a run under ncu is never a bench value; the bench line itself comes from a separate plain run.
set -u
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
# every launch of a single-stream step with its device time (cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches.csv $B --in-flight 1 > $out/${tag}_under_ncu.log 2>&1
cap() {  # name, kernel regex, skip, count, in-flight
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o $out/${tag}_$1 $B --in-flight $5 >> $out/${tag}_under_ncu.log 2>&1
}
cap gru_if1 gru_tc 10 2 1
cap gru_if4 gru_tc 8 1 4
cap gru_fused gru_fused 16 2 4
cap gru_seq gru_seq 50 2 8
cap layout block_layout 30 2 8
cap viterbi viterbi_k1024 3 1 1
cap gemm gemm_tf32x3 12 6 1
cap conv conv1d 3 1 1
ls -la $out/${tag}_*

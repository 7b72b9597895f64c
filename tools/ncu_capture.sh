#!/bin/bash
# ncu --set full captures of the kernels quoted in profiles/ (run on ONE B200 through gpurun; ~40 replays per kernel).
#   bash tools/ncu_capture.sh <tag>      -> gpurun_out/<tag>_*.ncu-rep
set -u
TAG=${1:-r1}
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
mkdir -p $OUT
# 1. tcgen05 GEMM at 8192^3 (tensor-bound)
timeout 600 $NCU -k regex:'k_gemm_tc<256>' -s 3 -c 1 -f -o $OUT/${TAG}_gemm8k python bench_all.py --only gemm --steps 1 > $OUT/${TAG}_gemm8k.log 2>&1
# 2. lm_head GEMM + argmax epilogue inside the decode step (batch 512), and the decode attention kernel
PDN_BENCH_TOTAL_LEN=12 timeout 600 $NCU -k regex:'k_gemm_tc<256>' -s 6 -c 1 -f -o $OUT/${TAG}_lmhead python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 512 > $OUT/${TAG}_lmhead.log 2>&1
PDN_BENCH_TOTAL_LEN=12 timeout 600 $NCU -k regex:'k_attention_fwd' -s 60 -c 1 -f -o $OUT/${TAG}_attn_decode python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 512 > $OUT/${TAG}_attn.log 2>&1
# 3. HBM-bound row kernels at encoder sizes
timeout 600 $NCU -k regex:'k_softmax_fwd|k_adam|k_feat_apply|k_feat_bwd_dx' -s 4 -c 4 -f -o $OUT/${TAG}_rows python bench_all.py --only rows --steps 1 > $OUT/${TAG}_rows.log 2>&1
ls -la $OUT/*.ncu-rep

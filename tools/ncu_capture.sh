#!/bin/bash
# ncu --set full captures of the kernels quoted in profiles/ (run on ONE B200 through gpurun; ~40 replays per kernel).
#   bash tools/ncu_capture.sh <tag>      -> gpurun_out/<tag>_*.ncu-rep
set -u
TAG=${1:-r1f}
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
mkdir -p $OUT
# 1. KV-cache decode attention (the top kernel of the bench step): batch 1024, context 130, standalone micro-benchmark
timeout 300 $NCU -k regex:k_attention_rows -s 5 -c 1 -f -o $OUT/${TAG}_attn_rows python bench_all.py --only decode --steps 1 > $OUT/${TAG}_attn_rows.log 2>&1
# 2. tcgen05 GEMM at 8192^3 (tensor-bound)
timeout 300 $NCU -k regex:k_gemm_tc -s 3 -c 1 -f -o $OUT/${TAG}_gemm8k python bench_all.py --only gemm --steps 1 > $OUT/${TAG}_gemm8k.log 2>&1
# 3. TMA-tiled convolution (forward, backward-data, backward-weight) at 64->128 56x56 batch 128, and the flash-attention passes
timeout 400 $NCU -k regex:'k_conv_tma|k_attn_tc' -s 9 -c 8 -f -o $OUT/${TAG}_conv_attn python bench_all.py --only micro --steps 2 > $OUT/${TAG}_conv_attn.log 2>&1
# 4. lm_head GEMM with the argmax epilogue inside a short decode (batch 1024)
PDN_BENCH_TOTAL_LEN=8 timeout 300 $NCU -k regex:'k_gemm_tc<256' -s 3 -c 1 -f -o $OUT/${TAG}_lmhead python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_lmhead.log 2>&1
ls -la $OUT/*.ncu-rep

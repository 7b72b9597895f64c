#!/bin/bash
# ncu --set full captures of the kernels quoted in profiles/ (run on ONE B200 through gpurun; ~40 replays per kernel).
#   bash tools/ncu_capture.sh <tag>      -> gpurun_out/<tag>_*.ncu-rep
set -u
TAG=${1:-r1}
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
mkdir -p $OUT
# 1. tcgen05 GEMM at 8192^3 (tensor-bound): 4th launch = first timed step
timeout 600 $NCU -k regex:k_gemm_tc -s 3 -c 1 -f -o $OUT/${TAG}_gemm8k python bench_all.py --only gemm --steps 1 > $OUT/${TAG}_gemm8k.log 2>&1
# 2. decode step at batch 512: the lm_head GEMM + argmax epilogue is every 25th k_gemm_tc launch (6 layers x 4 + 1)
PDN_BENCH_TOTAL_LEN=12 timeout 600 $NCU -k regex:k_gemm_tc -s 124 -c 1 -f -o $OUT/${TAG}_lmhead python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 512 > $OUT/${TAG}_lmhead.log 2>&1
#    decode attention at a realistic context (~130 cached keys): skip 3 warm-up passes (3 x 6 x 252 launches) + half a pass
timeout 900 $NCU -k regex:k_attention_fwd -s 5300 -c 1 -f -o $OUT/${TAG}_attn_decode python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 512 > $OUT/${TAG}_attn.log 2>&1
# 3. HBM-bound row kernels at encoder sizes (one of each)
timeout 600 $NCU -k regex:'k_softmax_fwd|k_adam|k_feat_apply|k_feat_bwd_dx|k_feat_reduce' -s 6 -c 12 -f -o $OUT/${TAG}_rows python bench_all.py --only rows --steps 1 > $OUT/${TAG}_rows.log 2>&1
# 4. tensor-core flash attention (forward pass kernel and dK kernel) at B128 H8 S512 hd64
timeout 600 $NCU -k regex:k_attn_tc -s 10 -c 5 -f -o $OUT/${TAG}_attn_tc python bench_all.py --only micro_att --steps 1 > $OUT/${TAG}_attn_tc.log 2>&1
ls -la $OUT/*.ncu-rep

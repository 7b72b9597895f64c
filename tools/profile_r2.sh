#!/bin/bash
# Round-2 profiling pass (one GPU): launch lists (gpu__time_duration, the B200_PROFILING.md recipe) of the bench command and of the
# new kernels' workloads, plus one `--set full` capture each of k_decode_mega, k_gru_persist_fwd/bwd and k_attention_rows.
# Outputs land in gpurun_out/ (copy the summaries into profiles/).
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
echo "== launch list: bench.py (short passes)"
PDN_BENCH_TOTAL_LEN=16 timeout 600 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file gpurun_out/r2_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-dp-train --no-token-check > gpurun_out/r2_launches_bench.log 2>&1
echo "== launch list: LeNet step (eager + recorded) and GRU step"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r2_launches_c2.csv \
  python bench_all.py --only c2 --no-cpu --steps 2 > gpurun_out/r2_launches_c2.log 2>&1
T=64 timeout 600 $NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/r2_launches_gru.csv \
  python tools/dbg/gru_train_step.py > gpurun_out/r2_launches_gru.log 2>&1
echo "== full captures"
TOTAL=48 timeout 900 $NCU --set full --import-source on -k regex:k_decode_mega -s 60 -c 1 -f -o gpurun_out/r2_mega python tools/dbg/b1_short.py > gpurun_out/r2_mega.log 2>&1
T=128 timeout 900 $NCU --set full --import-source on -k regex:k_gru_persist -c 2 -f -o gpurun_out/r2_gru python tools/dbg/gru_train_step.py > gpurun_out/r2_gru.log 2>&1
for f in r2_mega r2_gru; do
  [ -f gpurun_out/$f.ncu-rep ] && ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
done
ls -la gpurun_out/r2_* | head -30

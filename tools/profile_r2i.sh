#!/bin/bash
# Round-2 (final state) ncu pass on one GPU: --set full captures of the kernels added late in the round (persistent LSTM / RNN,
# thin first-layer convolution, attention passes after the split / staging changes) + the LeNet launch list. Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
T=64 timeout 600 $NCU --set full --import-source on -k regex:k_lstm_persist -s 2 -c 2 -f -o gpurun_out/r2i_lstm python tools/dbg/lstm_time.py > gpurun_out/r2i_lstm.log 2>&1
T=64 timeout 600 $NCU --set full --import-source on -k regex:k_rnn_persist -s 2 -c 2 -f -o gpurun_out/r2i_rnn python tools/dbg/rnn_time.py > gpurun_out/r2i_rnn.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_conv_thin -c 2 -f -o gpurun_out/r2i_thin python tools/dbg/thin_conv_check.py > gpurun_out/r2i_thin.log 2>&1
B=32 timeout 600 $NCU --set full --import-source on -k regex:k_attn_tc -s 10 -c 5 -f -o gpurun_out/r2i_attn python tools/dbg/attn_trace.py > gpurun_out/r2i_attn.log 2>&1
for f in r2i_lstm r2i_rnn r2i_thin r2i_attn; do
  [ -f gpurun_out/$f.ncu-rep ] && python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/${f}_summary.txt 2>&1
done
cat gpurun_out/r2i_*_summary.txt | cut -c1-400

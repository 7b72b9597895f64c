#!/bin/bash
# Round-end measurement pass on one B200: GPU tests, smoke, bench (both arms), bench_all. Outputs under gpurun_out/<tag>_*.
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/${tag}_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; tail -c 400 gpurun_out/${tag}_bench_ref.json
timeout 1200 python bench_all.py > gpurun_out/${tag}_bench_all.jsonl 2> gpurun_out/${tag}_bench_all.err; cut -c1-260 gpurun_out/${tag}_bench_all.jsonl

set -x
timeout 600 python -m pytest tests/test_llama_fast_gpu.py -m gpu -x -q 2>&1 | tail -5
for nb in 1 2 4; do
  PDN_DECODE_BRANCHES=$nb timeout 300 python bench.py --steps 10 --warmup 3 --no-b1 --no-dp-train --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('NB', $nb, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'], 3), 'att_us', round(d['roofline']['launch_us'], 1), 'lm_us', round(d['roofline']['gemm_view']['launch_us'], 1), 'tokcheck', d.get('token_check', {}).get('ok'), 'same', d['resident_vs_e2e'])
"
done
PDN_DECODE_BRANCHES=4 PDN_DECODE_ORDER=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-b1 --no-dp-train --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('NB 4 unordered value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'], 3), 'att_us', round(d['roofline']['launch_us'], 1), 'tokcheck', d.get('token_check', {}).get('ok'))
"

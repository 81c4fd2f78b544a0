cd /root/repo
for rep in 1 2; do
timeout 600 python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunked flush:', round(j['ms_per_step'],4), j['roofline']['stages_ms'], j['fast_path_misses'], j['front_half'])"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or cfg or fast_path or graph" 2>&1 | tail -4

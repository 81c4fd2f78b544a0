cd /root/repo
for m in 0 1 2; do
  touch abc-dls_b200/csrc/fused.cuh
  (cd abc-dls_b200/csrc && EXTRA="-DABCDLS_QMODE=$m" bash build.sh > /dev/null 2>&1)
  for rep in 1 2; do
  timeout 600 python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mode $m', round(j['ms_per_step'],4), j['roofline']['stages_ms']['k_fused_pass'], j['fast_path_misses'])"
  done
done

cd /root/repo
for cfg in "160 16" "128 24" "96 32"; do
  set -- $cfg
  touch abc-dls_b200/csrc/fused.cuh
  (cd abc-dls_b200/csrc && EXTRA="-DABCDLS_FRING_KB=$1 -DABCDLS_FQCAP=$2" bash build.sh > /dev/null 2>&1)
  for rep in 1 2; do
  timeout 600 python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ring $1 KB qcap $2:', round(j['ms_per_step'],4), j['roofline']['stages_ms']['k_fused_pass'], j['fast_path_misses'])"
  done
done

cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['host_enqueue_ms'], j['roofline']['stages_ms'], j['fast_path_misses'])"; done
bash scratch/launches.sh v8

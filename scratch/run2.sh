cd /root/repo
for iv in 0.01 0.05 10; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --clock-interval $iv 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('$iv', j['ms_per_step'], j['host_enqueue_ms'], j['clocks'], j['roofline']['stages_ms'])"
done

cd /root/repo
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_g.json 2>gpurun_out/bench_g.err
python -c "
import json
j=json.load(open('gpurun_out/bench_g.json')); print(j['ms_per_step'], j['host_enqueue_ms'], j['cuda_graph'], j['front_half'], j['fast_path_misses'], j['gpu_launches'])"
ABCDLS_NO_GRAPH=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('nograph', j['ms_per_step'], j['host_enqueue_ms'], j['cuda_graph'], j['front_half'])"
bash scratch/launches.sh v7

# the 1-GPU measurement set behind profiles/r02_* and BASELINE.md (run under gpurun from the repo root)
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_cfg5_1gpu.json 2> gpurun_out/f1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/f1.err
python bench.py --config cfg3 --steps 50 --warmup 5 > gpurun_out/r02_bench_cfg3_1gpu.json 2>> gpurun_out/f1.err
python bench.py --workload postpr --steps 50 --warmup 5 > gpurun_out/r02_bench_cfg2_postpr_1gpu.json 2>> gpurun_out/f1.err
python bench.py --workload smc --steps 3 --warmup 1 > gpurun_out/r02_bench_cfg4_smc_1gpu.json 2>> gpurun_out/f1.err
python bench.py --workload sweep --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_bench_sweep_1gpu.json 2>> gpurun_out/f1.err
python bench.py --workload cv4abc > gpurun_out/r02_bench_cfg1_cv4abc.json 2>> gpurun_out/f1.err
python bench.py --rows 50000000 --stats 24 --params 24 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_bench_wide24_1gpu.json 2>> gpurun_out/f1.err
ABCDLS_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fused_pass|k_sample_fine|k_fine_brackets" --launch-skip 6 -c 3 -o gpurun_out/prof_r02_pass_fine -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/f1.err

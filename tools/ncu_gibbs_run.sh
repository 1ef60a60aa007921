set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_golden_reference.py -m gpu -q -x 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:gibbs_level_kernel -s 27 -c 9 -o gpurun_out/r2_gibbs_levels -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-predict-leg > gpurun_out/r2_ncu_gibbs.log 2>&1
ncu --set full --clock-control none --import-source on -k build_level_kernel -s 4 -c 1 -o gpurun_out/r2_build_leaf -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-predict-leg > gpurun_out/r2_ncu_full_b.log 2>&1
tail -2 gpurun_out/r2_ncu_full_b.log
python tools/perf_probe.py C4 6 2>&1 | grep -E "it [25]|best"

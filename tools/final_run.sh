# Round-end measurement pass on one B200 (run through gpurun): GPU tests, bench (both arms), ncu launch list + DRAM traffic of
# one BUILD, ncu full captures of the two dominant BUILD launches, benches of the other configs.  Outputs under gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 40 --warmup 5 > gpurun_out/r2_bench_C4_n1.json 2> gpurun_out/r2_bench_C4_n1.err; tail -c 300 gpurun_out/r2_bench_C4_n1.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_C4_ref.json 2> gpurun_out/r2_bench_C4_ref.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 260 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-predict-leg > gpurun_out/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:build_level_kernel_ref -s 23 -c 1 -o gpurun_out/r2_build_level8 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-predict-leg > gpurun_out/r2_ncu_full_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"build_level_kernel<" -s 4 -c 1 -o gpurun_out/r2_build_leaf -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-predict-leg > gpurun_out/r2_ncu_full_b.log 2>&1
for w in C1 C2 C3; do python bench.py --workload $w --steps 200 --warmup 5 > gpurun_out/r2_bench_${w}_n1.json 2> gpurun_out/r2_bench_${w}_n1.err; cut -c1-160 gpurun_out/r2_bench_${w}_n1.json; done
python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline --no-predict-leg > gpurun_out/r2_bench_C5_n1.json 2> gpurun_out/r2_bench_C5_n1.err; cut -c1-200 gpurun_out/r2_bench_C5_n1.json
ST_PROFILE_BUILD=1 python tools/perf_probe.py C4 2 > gpurun_out/r2_probe_build.log 2>&1
ST_PROFILE_GIBBS=1 python tools/perf_probe.py C4 5 > gpurun_out/r2_probe_gibbs.log 2>&1

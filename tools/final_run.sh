set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 40 --warmup 4 > gpurun_out/bench_r1f_n1.json 2> gpurun_out/bench_r1f_n1.err; tail -c 600 gpurun_out/bench_r1f_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1f_ref.json 2> gpurun_out/bench_r1f_ref.err
ST_PROFILE_BUILD=1 python tools/perf_probe.py C4 2 > gpurun_out/probe_f_prof.log 2>&1
ST_PROFILE_MCMC=1 python tools/perf_probe.py C4 8 > gpurun_out/probe_f.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 140 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:build_level_kernel --launch-skip 7 --launch-count 2 -o gpurun_out/prof_build_r1f -f python tools/perf_probe.py C4 1 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gibbs_level_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/prof_gibbs_r1f -f python tools/perf_probe.py C4 1 > gpurun_out/ncu_f3.log 2>&1
ls -la gpurun_out/*.ncu-rep

# Round-end measurement pass on one B200 (run through gpurun): GPU tests, bench (both arms), phase profiles, ncu launch list.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 40 --warmup 4 > gpurun_out/bench_r1g_n1.json 2> gpurun_out/bench_r1g_n1.err; tail -c 300 gpurun_out/bench_r1g_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1g_ref.json 2> gpurun_out/bench_r1g_ref.err
ST_PROFILE_BUILD=1 python tools/perf_probe.py C4 2 > gpurun_out/probe_g_build.log 2>&1
ST_PROFILE_GIBBS=1 python tools/perf_probe.py C4 5 > gpurun_out/probe_g_gibbs.log 2>&1
python tools/perf_probe.py C4 9 > gpurun_out/probe_g.log 2>&1; tail -4 gpurun_out/probe_g.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_g1.log 2>&1
for w in C1 C2 C3; do python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1g_$w.json 2> gpurun_out/bench_r1g_$w.err; cut -c1-120 gpurun_out/bench_r1g_$w.json; done

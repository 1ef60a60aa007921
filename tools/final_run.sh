# Round-end measurement pass on one B200 (run through gpurun): GPU tests, smoke, bench (both arms), ncu launch list + DRAM traffic
# of one BUILD, benches of the other configs.  Outputs under gpurun_out/ (the summaries worth keeping are copied to profiles/).
# The ncu full captures of the dominant launches are made by tools/ncu_gibbs_run.sh / the commands quoted in profiles/r2_ncu_*.txt.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 40 --warmup 5 > gpurun_out/r2_bench_C4_n1.json 2> gpurun_out/r2_bench_C4_n1.err; tail -c 300 gpurun_out/r2_bench_C4_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_C4_ref.json 2> gpurun_out/r2_bench_C4_ref.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 260 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-predict-leg > gpurun_out/r2_ncu_launches.log 2>&1
for w in C1 C2 C3; do python bench.py --workload $w --steps 200 --warmup 5 > gpurun_out/r2_bench_${w}_n1.json 2> gpurun_out/r2_bench_${w}_n1.err; cut -c1-160 gpurun_out/r2_bench_${w}_n1.json; done

# round-2 record run on one B200: GPU tests, shape timings, bench lines, ncu launch list + one full capture
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.log 2>&1; tail -3 gpurun_out/r02_gputests.log
SHAPES="c5c:12 c5b:12 c5a:12 c5:16 map50:60" bash scripts/ab_run.sh > gpurun_out/r02_shapes.log 2>&1; cat gpurun_out/r02_shapes.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; tail -c 300 gpurun_out/r02_bench_c5.json
timeout 900 python bench.py --workload map50 --steps 20 --warmup 5 > gpurun_out/r02_bench_map50.json 2> gpurun_out/r02_bench_map50.err; tail -c 300 gpurun_out/r02_bench_map50.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --workload c5 --instances 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_launchlist.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dsqp_refine -c 1 -o gpurun_out/r02_refine_full python bench.py --workload c5 --instances 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -12

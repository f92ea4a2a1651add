# final record of round 2 (one B200): bench lines after the horizon buckets, launch list
set -x
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; tail -c 200 gpurun_out/r02_bench_c5.json
timeout 600 python bench.py --workload real --steps 10 --warmup 3 > gpurun_out/r02_bench_real_n1.json 2>/dev/null; tail -c 200 gpurun_out/r02_bench_real_n1.json
timeout 600 python bench.py --workload map50 --steps 20 --warmup 5 > gpurun_out/r02_bench_map50.json 2>/dev/null; tail -c 200 gpurun_out/r02_bench_map50.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --workload c5 --instances 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_launchlist.log 2>&1

set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.log 2>&1; tail -3 gpurun_out/r02_gputests.log
for w in "c5c 12" "map50 60" "c5 16" "c5a 12"; do set -- $w; timeout 300 python scripts/dev_shape.py $1 $2 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('perRC', d['workload'], '%.0f QP/s'%d['qp_per_s'], d['launch']['smem_bytes'], d['launch']['ctas_per_sm'])"; done > gpurun_out/r02_ab.log 2>&1; cat gpurun_out/r02_ab.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; tail -c 600 gpurun_out/r02_bench_c5.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --workload c5 --instances 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_launchlist.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dsqp_refine -c 1 -o gpurun_out/r02_refine_full python bench.py --workload c5 --instances 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -8

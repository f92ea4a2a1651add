# configs[4] at its full size (4096 instances, 409 600 agents) on one B200: one bench line for the record
mkdir -p gpurun_out
free -g | head -2
mem=$(free -g | awk '/Mem:/ {print $7}')
if [ "$mem" -lt 120 ]; then echo "not enough host memory for the pinned end-to-end buffers ($mem GB free)"; exit 0; fi
timeout 1500 python bench.py --instances 4096 --steps 3 --warmup 3 > gpurun_out/r02_bench_c5_4096_n1.json 2> gpurun_out/r02_bench_c5_4096_n1.err
tail -c 400 gpurun_out/r02_bench_c5_4096_n1.json; tail -3 gpurun_out/r02_bench_c5_4096_n1.err

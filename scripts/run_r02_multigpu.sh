# round-2 record run on N GPUs of one box: instance-sharded c5 (strong scaling) and agent-partitioned map100_a100
set -x
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_c5_n$N.json 2> gpurun_out/r02_bench_c5_n$N.err; tail -c 400 gpurun_out/r02_bench_c5_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload map100_a100 --partition agents --steps 5 --warmup 3 > gpurun_out/r02_bench_map100_agents_n$N.json 2> gpurun_out/r02_bench_map100_agents_n$N.err; tail -c 600 gpurun_out/r02_bench_map100_agents_n$N.json

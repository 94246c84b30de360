#!/bin/bash
# multi-GPU check + bench at N GPUs:  scripts/multi_bench.sh N [steps]
N=${1:-2}; K=${2:-200}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 scripts/multi_gpu_check.py > gpurun_out/multi_check_n$N.log 2>&1; tail -6 gpurun_out/multi_check_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps $K --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 1800 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
OQ_MATVEC=stream python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus $N --steps $K --warmup 5 --no-parity > gpurun_out/bench_stream_n$N.json 2> gpurun_out/bench_stream_n$N.err; head -c 330 gpurun_out/bench_stream_n$N.json

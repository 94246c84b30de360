mkdir -p gpurun_out
N=8
for v in 6 1; do
OQ_PANEL_P=$v OQ_TIMELINE=gpurun_out/tl_p$v python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$v bench.py --gpus $N --steps 100 --warmup 10 --no-parity > gpurun_out/tl_bench_p$v.json 2> gpurun_out/tl_bench_p$v.err
python scripts/timeline.py gpurun_out/tl_p$v.rank0.call1 gpurun_out/tl_p$v.rank5.call1 gpurun_out/tl_p$v.rank0.call2
done

#!/bin/bash
# One multi-GPU call (N GPUs of one box): multi-GPU parity tests, the driver's scaling measurement in miniature,
# the configs[4] assembly sweep and the coupled configs[3]-scale problem.   scripts/multi_round.sh N [sweeponly]
N=${1:-8}
mkdir -p gpurun_out
tr() { port=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
if [[ "$2" != "sweeponly" ]]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_multi_n$N.log; cat gpurun_out/pytest_multi_n$N.log
  bash scripts/scale_quick.sh $N
fi
tr 29811 scripts/assembly_sweep.py --out gpurun_out/assembly_sweep_n$N.json > gpurun_out/sweep_n$N.log 2>&1; tail -2 gpurun_out/sweep_n$N.log | cut -c1-400
if [[ "$2" != "sweeponly" ]]; then
  tr 29812 scripts/coupled_scaling.py --out gpurun_out/coupled_20k_19k_n$N.json > gpurun_out/coupled_a_n$N.log 2>&1; tail -1 gpurun_out/coupled_a_n$N.log | cut -c1-700
  # the same problem with the mantle operands in class form, and configs[3] as written (20k + 80k cells) in class form
  tr 29814 scripts/coupled_scaling.py --form classes --gf11 fft --out gpurun_out/coupled_20k_19k_classes_n$N.json > gpurun_out/coupled_c_n$N.log 2>&1; tail -1 gpurun_out/coupled_c_n$N.log | cut -c1-300
  tr 29815 scripts/coupled_scaling.py --form classes --gf11 fft --mantle 50 41 39 --steps 5 --out gpurun_out/coupled_20k_80k_classes_n$N.json > gpurun_out/coupled_d_n$N.log 2>&1; tail -1 gpurun_out/coupled_d_n$N.log | cut -c1-300
  if [[ $N -ge 8 ]]; then
    tr 29813 scripts/coupled_scaling.py --mantle 50 33 34 --out gpurun_out/coupled_20k_56k_n$N.json > gpurun_out/coupled_b_n$N.log 2>&1; tail -1 gpurun_out/coupled_b_n$N.log | cut -c1-700
  fi
fi

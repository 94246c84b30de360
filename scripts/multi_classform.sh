#!/bin/bash
# One multi-GPU call for the class-form path: the driver's scaling measurement in miniature (dense configs[2]) and the
# coupled problems with the mantle operands in class form, incl. configs[3] as written.   scripts/multi_classform.sh N
N=${1:-8}
mkdir -p gpurun_out
tr() { port=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
bash scripts/scale_quick.sh $N
tr 29821 scripts/coupled_scaling.py --form classes --gf11 fft --mantle 50 41 39 --steps 10 --out gpurun_out/coupled_20k_80k_classes_n$N.json > gpurun_out/coupled_d_n$N.log 2>&1; tail -1 gpurun_out/coupled_d_n$N.log | cut -c1-200
tr 29822 scripts/coupled_scaling.py --form classes --gf11 fft --out gpurun_out/coupled_20k_19k_classes_n$N.json > gpurun_out/coupled_c_n$N.log 2>&1; tail -1 gpurun_out/coupled_c_n$N.log | cut -c1-200

#!/bin/bash
# A/B at 8 GPUs on ONE box: peer-wait styles of the streaming pair (driver settings K=20, W=5 and a long run)
N=8; mkdir -p gpurun_out; port=29800
run() { name=$1; K=$2; shift 2; port=$((port+1)); env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps $K --warmup 5 --no-parity > gpurun_out/ab8_$name.json 2> gpurun_out/ab8_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab8_$name.json").read().strip().splitlines()[-1])
    print("$name K=$K", "us/step %.2f"%(1e3*d["ms_per_step"]), "evals/s %.0f"%d["value"], "kernel_us %.2f"%(1e3*d["roofline"]["kernel_ms"]), "e2e %.0f"%d["e2e"]["value"])
except Exception as e:
    print("$name FAILED", e)
PY
}
run serial_a 20 OQ_WAIT=serial
run warp_a 20 OQ_WAIT=warp
run serial_b 400 OQ_WAIT=serial
run warp_b 400 OQ_WAIT=warp
run serial_c 20 OQ_WAIT=serial
run warp_c 20 OQ_WAIT=warp

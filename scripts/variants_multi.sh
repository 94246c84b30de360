#!/bin/bash
# N-GPU bench of kernel variants (resident value + e2e):  scripts/variants_multi.sh N
N=${1:-8}
mkdir -p gpurun_out
port=29600
run() { name=$1; shift; port=$((port+1)); env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 400 --warmup 10 --no-parity > gpurun_out/varN${N}_$name.json 2> gpurun_out/varN${N}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/varN${N}_$name.json").read().strip().splitlines()[-1])
    print("N=$N $name", "us/step %.2f"%(1e3*d["ms_per_step"]), "evals/s %.0f"%d["value"], "kernel_us %.2f"%(1e3*d["roofline"]["kernel_ms"]), "frac %.3f"%d["roofline"]["frac"], "e2e %.0f"%d["e2e"]["value"])
except Exception as e:
    print("N=$N $name FAILED", e)
PY
}
run panel6 OQ_PANEL_P=6
run panel2 OQ_PANEL_P=2
run panel1 OQ_PANEL_P=1
run panel4 OQ_PANEL_P=4
run panel6_split OQ_PANEL_P=6 OQ_FORCING=split
run stream OQ_MATVEC=stream

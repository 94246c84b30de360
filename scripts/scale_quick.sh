#!/bin/bash
# the driver's scaling measurement in miniature: bench.py at N with the driver's K/W, three repetitions
N=${1:-8}
mkdir -p gpurun_out
for rep in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+rep)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scaleq_n${N}_$rep.json 2> gpurun_out/scaleq_n${N}_$rep.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scaleq_n${N}_$rep.json").read().strip().splitlines()[-1])
    print("N=$N rep $rep: us/step %.2f evals/s %.0f kernel %s frac %.3f e2e %.0f parity %s %.2e"%(1e3*d["ms_per_step"], d["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["e2e"]["value"], d["parity"]["pass"], d["parity"]["max_rel_err"]))
except Exception as e:
    print("rep $rep FAILED", e); print(open("gpurun_out/scaleq_n${N}_$rep.err").read()[-1500:])
PY
done

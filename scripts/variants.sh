#!/bin/bash
# N=1 bench of kernel variants (resident value only)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 200 --warmup 5 --no-extra --no-cpu --no-parity > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err; python - <<PY
import json
d=json.loads(open("gpurun_out/var_$name.json").read().strip().splitlines()[-1])
print("$name", "ms/step %.5f"%d["ms_per_step"], "evals/s %.1f"%d["value"], "kernel_ms %.5f"%d["roofline"]["kernel_ms"], "e2e %.1f"%d["e2e"]["value"])
PY
}
run panel6 OQ_PANEL_P=6
run panel4 OQ_PANEL_P=4
run panel2 OQ_PANEL_P=2
run panel1 OQ_PANEL_P=1
run panel6_split OQ_PANEL_P=6 OQ_FORCING=split
run panel1_split OQ_PANEL_P=1 OQ_FORCING=split
run stream OQ_MATVEC=stream
run panel6_nokeep OQ_PANEL_P=6 OQ_MATVEC_KEEP_MB=0 OQ_MATVEC_PINGPONG=0
run stream_nokeep OQ_MATVEC=stream OQ_MATVEC_KEEP_MB=0 OQ_MATVEC_PINGPONG=0

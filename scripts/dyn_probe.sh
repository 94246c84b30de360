#!/bin/bash
# compare the static-span and dynamic-unit matvec kernels on one GPU (tests first, then timings)
set -u
OQ_MATVEC=dyn timeout 400 python -m pytest tests/test_gpu_rhs.py tests/test_gpu_greens.py tests/test_gpu_solve.py tests/test_gpu_hex8.py -m gpu -q -x -k "not twins" 2>&1 | tail -8
pick='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["roofline"]["achieved"],1), round(d["roofline"]["kernel_ms"]*1e3,2))'
echo "static:"; timeout 200 python bench.py --steps 300 --warmup 5 --no-extra 2>/dev/null | python -c "$pick"
for pl in 2 4 8 16; do
  echo "dyn piece=$pl:"; OQ_MATVEC=dyn OQ_MATVEC_PIECE=$pl timeout 200 python bench.py --steps 300 --warmup 5 --no-extra 2>/dev/null | python -c "$pick"
done
echo "static shards:"; timeout 200 python scripts/shard_probe.py 2>&1 | tail -4
echo "dyn shards:"; OQ_MATVEC=dyn timeout 200 python scripts/shard_probe.py 2>&1 | tail -4
echo "dyn shards piece=2:"; OQ_MATVEC=dyn OQ_MATVEC_PIECE=2 timeout 200 python scripts/shard_probe.py 2>&1 | tail -4

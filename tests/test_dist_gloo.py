"""World-size-2 tests of the N>1 host logic on CPU (gloo): shard arithmetic, handle exchange order, local state
slicing, and the reference arm's rank-0-only contract.  The peer-memory data path itself needs GPUs
(scripts/multi_gpu_check.py under torchrun on the GPU box)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
import oetqf_b200 as oq
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
nf, ne = 16384, 37
rows = oq.dist.shard_range(nf, world, rank, align=4)
elems = oq.dist.shard_range(ne, world, rank)
mine = bytes([rank]) * 128
got = oq.dist.exchange_handles(mine)
assert [g[0] for g in got] == list(range(world)) and all(len(g) == 128 for g in got)
v = np.arange(nf, dtype=float).reshape(256, 64, order="F")
eps = np.arange(ne * 6, dtype=float).reshape(ne, 6, order="F")
loc = oq.dist.local_state((v, v + 1, eps, eps * 2, v + 2), rows, elems, kind="viscoelastic")
out = [None] * world
dist.all_gather_object(out, dict(rows=rows, elems=elems, v0=float(loc[0][0]), n=[int(x.size) for x in loc],
                                 e00=float(loc[2][0])))
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
"""


def _torchrun(args, timeout=240):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29533"] + args,
                          capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_shard_range_properties(oq):
    for n, world, align in [(16384, 8, 4), (32, 3, 4), (37, 2, 1), (5, 8, 1), (100, 1, 4)]:
        sh = oq.dist.all_shards(n, world, align)
        assert sh[0][0] == 0 and sh[-1][1] == n
        for (a0, a1), (b0, b1) in zip(sh[:-1], sh[1:]):
            assert a1 == b0 and a0 <= a1
        for a0, a1 in sh[:-1]:
            assert (a1 - a0) % align == 0 or a1 == n


def test_two_rank_exchange_and_slicing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    res = _torchrun([str(script)])
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("[{")][-1]
    out = json.loads(line)
    assert out[0]["rows"] == [0, 8192] and out[1]["rows"] == [8192, 16384]
    assert out[0]["elems"] == [0, 19] and out[1]["elems"] == [19, 37]
    assert out[1]["v0"] == 8192.0 and out[1]["e00"] == 19.0
    assert out[0]["n"] == [8192, 8192, 6 * 19, 6 * 19, 8192] and out[1]["n"][2] == 6 * 18


def test_reference_arm_prints_once_under_torchrun():
    res = _torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3",
                     "--warmup", "1"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0

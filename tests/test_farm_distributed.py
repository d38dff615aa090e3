"""The N>1 path: streams sharded round-robin over ranks, no collective on the data path, only the
barrier + max-over-ranks reduction of the timing.  Two gloo ranks on CPU (host simulator backend) against a
single-process run of the same streams."""
import importlib
import os
import subprocess
import sys

import pytest

from harness import farm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NUM_STREAMS, FRAMES, TRIS, RECTS = 6, 2, 60, 3

WORKER = r'''
import importlib, json, os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
from harness import farm, workloads as wl
from sidebyside import build_hostsim
rdp = importlib.import_module("parallel-rdp_b200")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = build_hostsim()
mine = farm.streams_for_rank({n}, rank, world)
digests = {{k: farm.render_stream(lambda rdram, flags: rdp.CommandProcessor(rdram, wl.HIDDEN_SIZE, flags=flags, library=lib), k, {frames}, {tris}, {rects}) for k in mine}}
# timing plumbing of bench.py: barrier on both sides, max over ranks
dist.barrier()
t = torch.tensor([10.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
gathered = [None] * world
dist.all_gather_object(gathered, digests)
if rank == 0:
    merged = {{}}
    for g in gathered:
        assert not (set(g) & set(merged)), "a stream was rendered by two ranks"
        merged.update(g)
    print("RESULT " + json.dumps({{"digests": {{str(k): v for k, v in merged.items()}}, "max_ms": float(t.item()), "world": world}}))
dist.destroy_process_group()
'''


def test_round_robin_assignment():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            mine = farm.streams_for_rank(1024, r, world)
            assert all(k % world == r for k in mine)
            seen += mine
        assert sorted(seen) == list(range(1024))
    assert [farm.stream_kind(k) for k in range(4)] == ["c2", "c3", "c4", "c2"]  # config 2 + (k mod 3)


@pytest.mark.timeout(600)
def test_two_gloo_ranks_render_the_same_streams_as_one_process(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, n=NUM_STREAMS, frames=FRAMES, tris=TRIS, rects=RECTS))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=580)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    import json
    res = json.loads(line[len("RESULT "):])
    assert res["world"] == 2 and res["max_ms"] == 11.0
    assert sorted(int(k) for k in res["digests"]) == list(range(NUM_STREAMS))

    # the same streams in this process, one after the other
    from harness import workloads as wl
    from sidebyside import build_hostsim
    rdp = importlib.import_module("parallel-rdp_b200")
    lib = build_hostsim()
    for k in range(NUM_STREAMS):
        serial = farm.render_stream(lambda rdram, flags: rdp.CommandProcessor(rdram, wl.HIDDEN_SIZE, flags=flags, library=lib), k, FRAMES, TRIS, RECTS)
        assert serial == res["digests"][str(k)], f"stream {k} differs between the sharded and the serial run"

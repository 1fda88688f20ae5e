"""N > 1 host logic on CPU: world_size-2 gloo process group, iteration sharding + the accumulator sum-reduce.
The per-rank 'renderer' here is the oracle (tests may use it); on GPUs the same gpugi_b200.parallel functions
drive RendererSystem.SetIterationShard + an NCCL all-reduce of the device accumulator."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
os.environ["OMP_NUM_THREADS"] = "2"
import numpy as np, torch
import torch.distributed as dist
from gpugi_b200 import parallel, scenes
from oracle import pyoracle as orc

rank, world, _ = parallel.env_rank_world()
d = parallel.init_process_group("gloo")
assert d is not None and d.get_world_size() == world == 2
TOTAL = 6
mine = parallel.shard_iterations(TOTAL, rank, world)
assert mine == [parallel.global_iteration(j, rank, world) for j in range(len(mine))]
sc = scenes.make_box1(tess=2)
W = H = 32
C = scenes.BOX1_CAMERA
cam = orc.camera(C["pos"], C["lookat"], C["fov"], W, H)

def samples_for(k):
    ls = orc.LightSampler(sc.positions(), sc.triangles(), sc.triangle_materials(), sc.emissivity(), sc.point_lights())
    ls.seed(1000 + k)   # a pure function of the GLOBAL iteration index
    return ls.generate(128)

acc = np.zeros((H, W, 4), np.float32)
for k in mine:
    acc, _ = orc.render_pt(sc.desc, cam, W, H, samples_for(k), orc.wanghash(k), 16, accum=acc)
local = acc.copy()                     # this rank's own image (the all-reduce below works in place)
t = torch.from_numpy(acc)
total = parallel.reduce_accumulator(t, len(mine), d)
assert total == TOTAL
if rank == 0:
    ref = np.zeros((H, W, 4), np.float32)
    for k in range(TOTAL):
        ref, _ = orc.render_pt(sc.desc, cam, W, H, samples_for(k), orc.wanghash(k), 16, accum=ref)
    # identical up to fp32 summation order (each rank sums its own iterations first)
    assert np.allclose(t.numpy(), ref, rtol=1e-5, atol=1e-6), float(np.abs(t.numpy() - ref).max())
    print("RANK0 OK")
# the transport of the multi-GPU read-out's rendezvous (gi_comm_connect leaves it to the caller): every rank gets every handle, in rank order
import gpugi_b200 as g
mine_h = bytes([rank + 1]) * 64
hs = parallel.exchange_handles(mine_h, d)
assert hs == [bytes([1]) * 64, bytes([2]) * 64]
# what the reduce-scatter kernel computes, restated with numpy on this rank's rows: rank-order sum of the per-rank images == the all-reduced image
first, rows = g.comm_slice(H, rank, world)
gathered = [torch.zeros_like(torch.from_numpy(local)) for _ in range(world)]
d.all_gather(gathered, torch.from_numpy(local))
mine_rows = gathered[0].numpy()[first:first + rows].copy()
for r in range(1, world):
    mine_rows += gathered[r].numpy()[first:first + rows]
assert np.array_equal(mine_rows, t.numpy()[first:first + rows]) or np.allclose(mine_rows, t.numpy()[first:first + rows], rtol=1e-6, atol=1e-7)
if rank == 0:
    print("COMM OK")
d.barrier()
d.destroy_process_group()
'''


def test_shard_iterations_partition():
    from gpugi_b200 import parallel
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            it = parallel.shard_iterations(21, r, world)
            assert all(k % world == r for k in it)
            seen += it
        assert sorted(seen) == list(range(21))


def test_two_rank_gloo_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "RANK0 OK" in out.stdout and "COMM OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_comm_slice_partitions_the_rows(g):
    """gi_comm_slice (pure, no GPU): the ranks' row ranges tile the image exactly, in rank order, for every world size the ABI allows."""
    for height in (1, 7, 60, 1080, 2160):
        for world in range(1, 17):
            nxt = 0
            for rank in range(world):
                first, rows = g.comm_slice(height, rank, world)
                assert first == nxt
                nxt = first + rows
            assert nxt == height
    import pytest
    with pytest.raises(g.GpugiError):
        g.comm_slice(10, 3, 3)


def test_rel_mse_definition():
    from gpugi_b200 import parallel
    r = np.ones((4, 4, 3), np.float32)
    x = r * 1.1
    assert abs(parallel.rel_mse(x, r) - 3 * 0.01 / 1.01) < 1e-6
    assert abs(parallel.mse(x, r) - 0.03) < 1e-6

"""Multi-GPU read-out behind the C ABI (gi_comm_create / gi_comm_connect / gi_read_accumulator_reduced[_async], csrc/gi_comm.cuh): one process per
GPU under torchrun, exchange blocks shared over CUDA IPC, snapshot + reduce-scatter over NVLink peer loads.  Needs >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _num_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch
import gpugi_b200 as g
from gpugi_b200 import parallel, scenes

rank, world, local = parallel.env_rank_world()
torch.cuda.set_device(local)
dist = parallel.init_process_group("nccl")
W, H = 200, 120
sc = scenes.make_box1(tess=3)
C = scenes.BOX1_CAMERA

def system(streaming):
    rs = g.RendererSystem(local)
    rs.SetScene(sc); rs.SetScreenSize(W, H); rs.SetCamera(C["pos"], C["lookat"], C["fov"]); rs.SetRenderer("PT")
    rs.SetIterationShard(rank, world)
    rs.SetStreaming(streaming)
    parallel.connect_ranks(rs.device, dist)
    return rs

first, rows = g.comm_slice(H, rank, world)
ITER = 7

# ---- 1. synchronous reduced read: every rank's rows == the rank-order sum of the per-rank images, bit for bit
rs = system(False)
for _ in range(ITER):
    rs.Draw()
own = rs.ReadBackbuffer().copy()
mine = np.zeros((rows, W, 4), np.float32)
rs.ReadBackbufferReduced(mine)
every = [torch.zeros((H, W, 4), dtype=torch.float32, device="cuda") for _ in range(world)]
dist.all_gather(every, torch.from_numpy(own).cuda())
want = every[0].cpu().numpy().copy()
for r in range(1, world):
    want = want + every[r].cpu().numpy()      # the same order of additions as k_comm_reduce
assert own[..., :3].sum() > 0
assert np.array_equal(mine.view(np.uint32), want[first:first + rows].view(np.uint32)), "rank %%d: reduced rows differ from the sequential sum" %% rank
# the ranks rendered DIFFERENT iterations (k = rank mod world)
if world > 1:
    assert not np.array_equal(every[0].cpu().numpy(), every[1].cpu().numpy())
# repeated reads (both snapshot buffers, flag protocol over several epochs)
for rep in range(5):
    rs.Draw()
    rs.ReadBackbufferReduced(mine)
own = rs.ReadBackbuffer().copy()
dist.all_gather(every, torch.from_numpy(own).cuda())
want = every[0].cpu().numpy().copy()
for r in range(1, world):
    want = want + every[r].cpu().numpy()
rs.ReadBackbufferReduced(mine)
assert np.array_equal(mine.view(np.uint32), want[first:first + rows].view(np.uint32))
rs.close()

# ---- 2. asynchronous reduced reads inside a stream: whole iterations only, and after the drain the same image as above
rs = system(True)
buf = [torch.empty((rows, W, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)]
for k in range(ITER + 5):
    rs.Draw()
    in_flight = rs.ReadBackbufferReducedAsync(buf[k & 1])
    assert in_flight == min(k + 1, 17)
rs.WaitReads()
final = np.zeros((rows, W, 4), np.float32)
rs.ReadBackbufferReduced(final)                # drains
assert np.array_equal(final.view(np.uint32), want[first:first + rows].view(np.uint32)), "rank %%d: streamed + reduced image differs" %% rank
# (with 12 iterations and 17 in flight every asynchronous snapshot was still empty)
assert not buf[0].any() and not buf[1].any()
rs.close()

# ---- 3. a communicator of the wrong size is refused after a resize
rs = system(False)
rs.SetScreenSize(64, 64)
try:
    rs.ReadBackbufferReduced(np.zeros((64, 64, 4), np.float32))
    raise SystemExit("reduced read after a resize was not refused")
except g.GpugiError as e:
    assert "screen size changed" in str(e)
rs.close()
dist.barrier()
if rank == 0:
    print("MULTI OK world=%%d" %% world)
dist.destroy_process_group()
'''


@pytest.mark.skipif(_num_gpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_reduced_readout_equals_sequential_sum(tmp_path, world):
    if _num_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
                          "--master-port", str(29650 + world), str(script)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and ("MULTI OK world=%d" % world) in out.stdout, out.stdout[-3000:] + out.stderr[-6000:]


def test_single_rank_communicator(g, scenes, box1):
    """world == 1: the reduced read is the plain read (no peer, no IPC)."""
    import numpy as np
    from gpugi_b200 import parallel
    W, H = 96, 64
    rs = g.RendererSystem(0)
    rs.SetScene(box1); rs.SetScreenSize(W, H); rs.SetCamera(scenes.BOX1_CAMERA["pos"], scenes.BOX1_CAMERA["lookat"], scenes.BOX1_CAMERA["fov"]); rs.SetRenderer("PT")
    with pytest.raises(g.GpugiError):
        rs.ReadBackbufferReduced(np.zeros((H, W, 4), np.float32))   # no communicator yet
    assert parallel.connect_ranks(rs.device, None) == (0, 1)
    for _ in range(3):
        rs.Draw()
    a = rs.ReadBackbuffer()
    b = rs.ReadBackbufferReduced(np.zeros((H, W, 4), np.float32))
    assert a[..., :3].sum() > 0 and np.array_equal(a, b)
    rs.SetStreaming(True)
    for k in range(4):
        rs.Draw()
        rs.ReadBackbufferReducedAsync(b)
    rs.WaitReads()
    assert np.array_equal(rs.ReadBackbufferReduced(b), rs.ReadBackbuffer())
    rs.close()

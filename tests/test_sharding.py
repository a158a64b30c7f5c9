"""Multi-GPU sharding logic on CPU: world_size=2 `gloo` processes.  Each rank traces its row
band / ray range (with the test-only host build of the traversal core, since there is no GPU
here), writes it into its slot of the gathered buffer and all_gathers -- the same plumbing
bench.py uses with NCCL on the B200 box.  The gathered frame must equal the oracle's."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, W, H, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    from common import HostEmul
    from oracle.oracle_py import Oracle
    pkg = g.load_pkg()
    emul, orc = HostEmul(), Oracle()
    orc.srand(1)
    s = orc.scene_random(200)           # every rank builds the same replicated tree
    h = emul.build(s, 200, 0, 200, 0)
    cam = orc.default_camera()
    # frame: row bands aligned to the CTA tile height
    y0, y1 = pkg.shard_rows(H, rank, world)
    bands = [pkg.shard_rows(H, r, world) for r in range(world)]
    assert bands[0][0] == 0 and bands[-1][1] == H and all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
    rows_max = max(b[1] - b[0] for b in bands)
    full = torch.zeros((world, rows_max * W, 8), dtype=torch.float32)
    mine, _ = emul.trace(h, emul.primary_rays(cam, W, H, y0, y1), threads=2)
    full[rank, :(y1 - y0) * W] = torch.from_numpy(mine.view(np.float32).reshape(-1, 8))
    dist.all_gather_into_tensor(full.view(-1, 8), full[rank].clone())
    frame = np.concatenate([full[r, :(bands[r][1] - bands[r][0]) * W].numpy() for r in range(world)])
    # ray buffer: contiguous ranges
    orc.srand(3)
    rays = orc.rays_benchmark(1001)
    lo, hi = pkg.shard_range(len(rays), rank, world)
    part, _ = emul.trace(h, rays[lo:hi], threads=1)
    sizes = [pkg.shard_range(len(rays), r, world) for r in range(world)]
    pad = max(b - a for a, b in sizes)
    buf = torch.zeros((world, pad, 8), dtype=torch.float32)
    buf[rank, :hi - lo] = torch.from_numpy(part.view(np.float32).reshape(-1, 8))
    dist.all_gather_into_tensor(buf.view(-1, 8), buf[rank].clone())
    gathered = np.concatenate([buf[r, :b - a].numpy() for r, (a, b) in enumerate(sizes)])
    if rank == 0:
        q.put((frame.tobytes(), gathered.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("W,H", [(160, 120), (96, 50)])
def test_two_rank_band_gather_equals_full_frame(W, H):
    from oracle.oracle_py import Oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, W, H, q)) for r in range(2)]
    for p in procs:
        p.start()
    frame, gathered = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    orc = Oracle()
    orc.srand(1)
    s = orc.scene_random(200)
    nodes = orc.build(s, 0, 200, 0)
    cam = orc.default_camera()
    want, _ = orc.trace_bvh(nodes, s, 200, orc.primary_rays(cam, W, H))
    assert frame == want.tobytes()
    orc.srand(3)
    rays = orc.rays_benchmark(1001)
    want, _ = orc.trace_bvh(nodes, s, 200, rays)
    assert gathered == want.tobytes()


def test_shard_arithmetic():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_pkg()
    for n in (0, 1, 7, 1000, 16777216):
        for world in (1, 2, 3, 4, 8):
            r = [pkg.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    # weighted bands (the root also consumes the frame): they tile the frame on 16-row boundaries, the root's band is
    # root_share of an equal share, the others differ by at most one block; share 1 (or world 1) is the equal split
    for H in (600, 2160, 4320, 6144, 50):
        for world in (1, 2, 4, 8):
            for share in (1.0, 0.9, 0.8, 0.5):
                b = [pkg.shard_rows_weighted(H, k, world, share) for k in range(world)]
                assert b[0][0] == 0 and b[-1][1] == H
                assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
                assert all(y0 % 16 == 0 for y0, _ in b)
                if share == 1.0 or world == 1:
                    assert b == [pkg.shard_rows(H, k, world) for k in range(world)]
                elif H >= 16 * world * 4:
                    rows = [y1 - y0 for y0, y1 in b]
                    assert abs(rows[0] - H / world * share) <= 16
                    assert max(rows[1:]) - min(rows[1:]) <= 16 and min(rows[1:]) >= rows[0]
    for H in (600, 2160, 4320, 50, 7):
        for world in (1, 2, 4, 8):
            b = [pkg.shard_rows(H, k, world) for k in range(world)]
            assert b[0][0] == 0 and b[-1][1] == H
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert all(y0 % 8 == 0 for y0, _ in b)

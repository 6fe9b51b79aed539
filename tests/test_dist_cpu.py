"""N>1 host logic on CPU: world_size-2 gloo process group (127.0.0.1), batch sharding + max-over-ranks timing +
index gather, i.e. everything bench.py --gpus N does besides launching kernels."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nuwa_pytorch_b200.parallel import gather_indices, max_over_ranks, rank_slice


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = rank_slice(7, rank, world)
    local = torch.arange(s, e, dtype=torch.int64)[:3].clone()  # 3 ids per rank (equal shapes for all_gather)
    t = max_over_ranks(0.5 + rank, dist)
    g = gather_indices(local, dist)
    dist.barrier()
    q.put((rank, (s, e), t, g.tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == (0, 4) and res[1][1] == (4, 7)          # contiguous, disjoint, covering shards
    assert res[0][2] == res[1][2] == 1.5                         # every rank sees the slowest rank's time
    assert res[0][3] == res[1][3] == [0, 1, 2, 4, 5, 6]          # rank-ordered gather


def test_rank_slice_covers_everything():
    for n in (1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            parts = [rank_slice(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))


def _reducer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nuwa_pytorch_b200.parallel import GradAllReduce
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)      # rank-dependent "gradients"
    red = GradAllReduce(dist, max_bucket_elems=128, min_bucket_elems=300)   # small buckets: chunking and merging
    red.begin(flat)
    red.ready(700, 900)   # ranges become final in backward order, with holes that finish() must cover;
    n0 = len(red.handles)  # 200 < min bucket: held back
    red.ready(300, 700)   # adjacent -> merged with the pending range into one 600-element bucket (5 chunks of <= 128)
    n1 = len(red.handles)
    red.ready(0, 0)
    red.ready(100, 150)   # not adjacent to anything pending, below the minimum: goes out in finish()
    red.finish()
    assert n0 == 0 and n1 == 5 and red.pending is None
    q.put((rank, flat.tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_mean():
    """The one collective of the training path: every element of the flat gradient buffer ends up as the mean over
    ranks exactly once, whatever order the ranges were announced in."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (torch.arange(1000, dtype=torch.float32) * 1.5).tolist()   # mean of x*1 and x*2
    assert res[0][1] == want and res[1][1] == want


def _convert_worker(rank, world, port, path, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nuwa_pytorch_b200.data import convert_video_tensor_dataset_to_indices
    from tests.helpers_data import StubVAE, StubVideos
    vae = StubVAE(image_size=64, num_layers=4, codebook=97)
    videos = StubVideos(n=5, frames=3, channels=3, size=64, seed=7)
    shape = convert_video_tensor_dataset_to_indices(vae=vae, raw_video_dataset=videos, num_frames=3, path=path, batch_videos=2,
                                                    rank=rank, world_size=world, barrier=dist.barrier)
    q.put((rank, tuple(shape), sum(vae.calls)))
    dist.destroy_process_group()


def test_two_rank_dataset_conversion_writes_the_reference_bytes(tmp_path):
    """data.convert_video_tensor_dataset_to_indices sharded over 2 ranks (videos are independent units, no data-path
    collective): the shared memmap equals the file the UNMODIFIED reference writer produced (tests/golden)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    path = str(tmp_path / "idx.bin")
    procs = [ctx.Process(target=_convert_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == (5, 48)
    assert (res[0][2], res[1][2]) == (3, 2)  # rank 0 encoded videos 0-2, rank 1 videos 3-4
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "video_indices_small.bin")
    assert open(path, "rb").read() == open(golden, "rb").read()

"""CPU, world_size 2 over gloo: the host-side multi-rank logic of the sharded paths (no collective on
the data path: ranks own disjoint clips; the only exchanges are bench.py's barrier + max-over-ranks
timing reduction, exercised here)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from melspec_gpt_vqvae_b200.feature_extraction.extract_codes import shard
    import bench
    paths = ["clip%03d" % i for i in range(11)]
    mine = shard(paths, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = bench.max_over_ranks(10.0 + rank)          # max-over-ranks timing reduction used by bench.py
    seeds = bench.rank_seed(783435, rank)
    if rank == 0:
        out.put((gathered, t, seeds))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_timing_reduction_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, t, seed0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert gathered[0] + gathered[1] == ["clip%03d" % i for i in range(11)]
    assert set(gathered[0]).isdisjoint(gathered[1])
    assert t == 11.0 and seed0 == 783435


def test_bench_roofline_arithmetic_matches_survey_figures():
    """bench.py's algorithmic bytes of the decode loop against SURVEY section 8(d) / BASELINE.md section 4: 604.9 MB of bf16
    weights per position, 98 304 B of KV per sequence and context position, ~382 GB per 64 clips of 265 tokens."""
    import bench
    from melspec_gpt_vqvae_b200 import synthetic
    cfg = synthetic.GPT_VAS
    one = bench.decode_algorithmic_bytes(1, 1, cfg)
    C, L, V = cfg["n_embd"], cfg["n_layer"], cfg["vocab_size"]
    weights = (L * 12 * C * C + V * C) * 2
    assert abs(weights / 1e6 - 604.9) < 1.0          # 604.2 MB of Linear weights (SURVEY rounds in the embeddings)
    assert one == weights + 98304
    total = bench.decode_algorithmic_bytes(64, 265, cfg)
    assert total == 265 * weights + 64 * 98304 * (265 * 266 // 2)
    assert abs(total / 1e9 - 381.9) < 0.2

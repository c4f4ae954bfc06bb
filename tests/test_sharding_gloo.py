"""CPU, world_size=2 (gloo): the utterance partition used for multi-GPU inference is a disjoint cover that equals
the reference's np.array_split, and bench.py's reference arm only speaks on rank 0."""
import json
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avvad.sharding import shard_bounds, shard_list, batches

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_match_array_split():
    for n in (0, 1, 7, 9, 10, 1250, 10000):
        for w in (1, 2, 3, 4, 8):
            parts = np.array_split(np.arange(n), w)
            for r in range(w):
                a, b = shard_bounds(n, w, r)
                assert list(range(a, b)) == parts[r].tolist()


def _worker(rank, world, port, n_items, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_list(list(range(n_items)), world, rank)
    # every rank reports how many frames it would process; all-gather only to CHECK the partition (the data path has
    # no collective)
    t = torch.zeros(n_items, dtype=torch.int32)
    t[mine] = 1
    dist.all_reduce(t)
    cover_ok = bool((t == 1).all())
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(mine)]))
    if rank == 0:
        json.dump({"cover_ok": cover_ok, "sizes": [int(s) for s in sizes]}, open(os.path.join(out_dir, "r.json"), "w"))
    dist.destroy_process_group()


def test_two_rank_partition_is_a_disjoint_cover(tmp_path):
    port = 29400 + os.getpid() % 500
    mp.spawn(_worker, args=(2, port, 1251, str(tmp_path)), nprocs=2, join=True)
    r = json.load(open(tmp_path / "r.json"))
    assert r["cover_ok"] and r["sizes"] == [626, 625]


def test_batches_grouping():
    assert list(batches(list(range(5)), 2)) == [[0, 1], [2, 3], [4]]


def test_reference_arm_prints_only_on_rank0():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""

"""Host logic of the sharded batch evaluation (BASELINE config 4): rank blocks + call plan at the full 10,000-utterance
size.  No GPU: only the partition / grouping arithmetic of avvad/evaluate.py and avvad/sharding.py."""
import numpy as np
import pytest

from avvad.evaluate import padding_overhead, plan_calls
from avvad.sharding import shard_bounds


def _frame_counts(n, seed=0):
    # BASELINE config 4: N ~ U{64,000..102,400} samples -> T = 247..397 frames (hop 256, pad-at-end framing of the front end)
    ns = np.random.default_rng(seed).integers(64000, 102401, size=n)
    return [int(1 + (m - 1) // 256) for m in ns]  # any monotone map of the length is enough for the plan properties


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_config4_partition_and_plan(world):
    T = _frame_counts(10000)
    seen = []
    for rank in range(world):
        a, b = shard_bounds(len(T), world, rank)
        shard = T[a:b]
        assert abs(len(shard) - 10000 / world) < 1
        calls = plan_calls(shard, 256)
        flat = [i for c in calls for i in c]
        assert sorted(flat) == list(range(len(shard)))               # every utterance of the shard exactly once
        assert all(len(c) == 256 for c in calls[:-1]) and 0 < len(calls[-1]) <= 256
        for c in calls:                                               # descending lengths inside and across calls
            assert all(shard[c[k]] >= shard[c[k + 1]] for k in range(len(c) - 1))
        for c0, c1 in zip(calls, calls[1:]):
            assert shard[c0[-1]] >= shard[c1[0]]
        unsorted = plan_calls(shard, 256, sort_by_length=False)
        assert [i for c in unsorted for i in c] == list(range(len(shard)))
        # sorting is what removes the collate padding: ~20 % of the trunk work in list order, a few % sorted
        assert padding_overhead(shard, unsorted) > 0.15
        assert padding_overhead(shard, calls) < 0.06
        seen.extend(range(a, b))
    assert seen == list(range(10000))


def test_plan_edge_cases():
    assert plan_calls([], 4) == []
    assert plan_calls([5], 4) == [[0]]
    assert plan_calls([3, 3, 3], 2) == [[0, 1], [2]]                  # stable for equal lengths
    assert plan_calls([1, 9, 4], 8) == [[1, 2, 0]]
    assert padding_overhead([], []) == 0.0
    with pytest.raises(ValueError):
        plan_calls([1, 2], 0)


# ---- the evaluation loop itself, two gloo ranks on the CPU, with a stub in place of the device pipeline ----------------
class _StubPipeline:
    """Stands in for AVVADPipeline.infer_host: posterior of frame t of an utterance = a fixed function of that
    utterance's own samples and frames only (as the per-utterance mode guarantees on the device), so any mistake in the
    host loop -- wrong row order after sorting, lengths attached to the wrong utterance, stale staging contents,
    a shard boundary off by one -- changes the result."""

    def __init__(self):
        self.calls = []

    def infer_host(self, wave, n_samples, video, n_src, lengths=None, per_utterance=None, t_max=None):
        import torch

        assert per_utterance is True
        B, T = wave.shape[0], max(lengths)
        self.calls.append((B, T))
        post = torch.full((B, T, 1), 0.5)
        for b in range(B):
            w = wave[b, : n_samples[b]].double()
            v = video[b, : n_src[b]].double()
            key = float(w.sum()) * 1e-3 + float(v.mean()) * 1e-2
            t = torch.arange(lengths[b], dtype=torch.float64)
            post[b, : lengths[b], 0] = torch.sigmoid(torch.sin(key + 0.1 * t)).float()
        return post, (post > 0.5).int()


def _utterance(i):
    rng = np.random.default_rng(1000 + i)
    n = int(rng.integers(3000, 9000))
    f = int(rng.integers(6, 18))
    return rng.standard_normal(n).astype(np.float32) * 0.1, rng.integers(0, 256, size=(f, 67, 67), dtype=np.uint8)


def _eval_worker(rank, world, port, n_items, out_dir):
    import json
    import os

    import torch
    import torch.distributed as dist

    from avvad.evaluate import evaluate_sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    utts = [_utterance(i) for i in range(n_items)]
    pipe = _StubPipeline()
    a, b, res = evaluate_sharded(pipe, utts, world, rank, batch_size=4)
    sums = torch.zeros(n_items, dtype=torch.float64)
    lens = torch.zeros(n_items, dtype=torch.int64)
    for i, (soft, hard) in zip(range(a, b), res):
        assert torch.equal(hard, (soft > 0.5).to(hard.dtype))
        sums[i] = float(soft.double().sum())
        lens[i] = soft.numel()
    dist.all_reduce(sums)   # only to CHECK the result: the data path has no collective
    dist.all_reduce(lens)
    if rank == 0:
        json.dump({"sums": sums.tolist(), "lens": lens.tolist(), "calls": pipe.calls, "block": [a, b]},
                  open(os.path.join(out_dir, "eval.json"), "w"))
    dist.destroy_process_group()


def test_sharded_evaluation_two_gloo_ranks_equals_one_by_one(tmp_path):
    import json
    import os

    import torch
    import torch.multiprocessing as mp

    from avvad.evaluate import evaluate_shard
    from avvad.pipeline import AVVADPipeline

    n_items = 11
    port = 29900 + os.getpid() % 90
    mp.spawn(_eval_worker, args=(2, port, n_items, str(tmp_path)), nprocs=2, join=True)
    got = json.load(open(tmp_path / "eval.json"))
    assert got["block"] == [0, 6]                       # np.array_split: 6 + 5
    assert [c[0] for c in got["calls"]] == [4, 2]       # rank 0: calls of 4 and 2 utterances
    # one utterance per call, list order, no sorting: the reference's loop
    utts = [_utterance(i) for i in range(n_items)]
    one = evaluate_shard(_StubPipeline(), utts, batch_size=1, sort_by_length=False)
    T = AVVADPipeline.frame_counts([len(w) for w, _ in utts], [v.shape[0] for _, v in utts])
    assert got["lens"] == T
    for i, (soft, _) in enumerate(one):
        assert soft.numel() == T[i]
        assert abs(float(soft.double().sum()) - got["sums"][i]) < 1e-9, i

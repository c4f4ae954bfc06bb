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

"""GPU (B200): the lock-free block hash on its own (vh_map_*), the replacement of vhashing::HashTable."""
import numpy as np
import pytest

from util import key_set

pytestmark = pytest.mark.gpu


def test_insert_dedupes_under_contention(vh):
    rng = np.random.RandomState(0)
    uniq = rng.randint(-5000, 5000, (40000, 3)).astype(np.int32)
    uniq = np.unique(uniq, axis=0)
    keys = np.concatenate([uniq] * 8)                  # every key 8 times, shuffled: many warps race on the same CAS
    rng.shuffle(keys)
    m = vh.BlockHashMap(1 << 16, 4, 1 << 17)
    slots = m.insert(keys)
    assert slots.min() >= 0
    assert len(m) == len(uniq)                         # no duplicate insert (reference race)
    # equal keys share a slot, different keys never do
    by_key = {}
    for k, s in zip(map(tuple, keys.tolist()), slots.tolist()):
        assert by_key.setdefault(k, s) == s
    assert len(set(by_key.values())) == len(uniq)
    assert key_set(m.keys()) == key_set(uniq)          # key_heap == set of inserted keys
    assert len(m.keys()) == len(uniq)
    assert np.array_equal(m.find(keys), slots)
    absent = uniq + np.array([20000, 0, 0], np.int32)
    assert np.all(m.find(absent) == -1)
    m.close()


def test_erase_and_reinsert(vh):
    m = vh.BlockHashMap(1 << 12, 4, 1 << 12)
    keys = np.stack(np.meshgrid(np.arange(-8, 8), np.arange(-8, 8), np.arange(0, 8), indexing="ij"), -1).reshape(-1, 3).astype(np.int32)
    s0 = m.insert(keys)
    assert len(m) == len(keys) == 2048
    half = keys[::2]
    er = m.erase(np.concatenate([half, half]))          # duplicates in the batch erase once
    assert er.sum() == len(half)
    assert len(m) == len(keys) - len(half)
    assert np.all(m.find(half) == -1) and np.all(m.find(keys[1::2]) == s0[1::2])
    assert key_set(m.keys()) == key_set(keys[1::2])
    s1 = m.insert(half)                                 # slots are recycled from the free list
    assert len(m) == len(keys) and s1.min() >= 0
    assert len(set(s1.tolist()) | set(s0[1::2].tolist())) == len(keys)
    m.close()


def test_capacity_errors_are_reported(vh):
    m = vh.BlockHashMap(256, 4, 100)                    # 1024 entries, 100 value slots
    keys = np.arange(300 * 3, dtype=np.int32).reshape(-1, 3)
    with pytest.raises(vh.VhError) as ei:
        m.insert(keys)
    assert ei.value.code == 5                           # VH_ERR_POOL_FULL, not a hang
    m.close()
    with pytest.raises(vh.VhError):
        vh.BlockHashMap(16, 4, 16).insert(np.array([[1 << 21, 0, 0]], np.int32))   # out of the 21-bit range

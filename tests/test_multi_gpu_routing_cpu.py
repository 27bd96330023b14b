"""Host-side routing of the shard-aware flat index (vectorlite_b200/multi_gpu.py) on CPU: the shards
are stand-ins that only store rows, so this covers where rows go, what a delete / re-split does to the global
storage order and the per-shard position bases — not the search (that needs the device: tests/test_multi_gpu.py)."""
import numpy as np
import pytest

import vectorlite_b200 as vl
from vectorlite_b200.multi_gpu import MultiGpuFlatIndex


class StoreOnlyShard:
    """The mutation half of FlatIndex (flat.rs:82-96) over numpy arrays."""

    def __init__(self, dim, device):
        self.dim, self.device = dim, device
        self.ids, self.rows, self.pos_base, self.closed = [], [], None, False

    def add(self, v):
        self.ids.append(int(v.id))
        self.rows.append(np.asarray(v.values, dtype=np.float32))

    def add_batch(self, ids, rows, texts=None, metadata=None):
        for i, r in zip(ids, rows):
            self.ids.append(int(i))
            self.rows.append(np.asarray(r, dtype=np.float32))

    def delete(self, id_):
        j = self.ids.index(int(id_))
        del self.ids[j], self.rows[j]

    def len(self):
        return len(self.ids)

    def export(self):
        return np.array(self.ids, dtype=np.uint64), np.array(self.rows, dtype=np.float32).reshape(len(self.ids), self.dim)

    def get_vector(self, id_):
        return vl.Vector(id=int(id_), values=self.rows[self.ids.index(int(id_))])

    def set_pos_base(self, base):
        self.pos_base = int(base)

    def close(self):
        self.closed = True


def _mk(dim, devices, **kw):
    return MultiGpuFlatIndex(dim, devices, shard_factory=StoreOnlyShard, **kw)


def test_tail_routing_deletes_and_resplit():
    rng = np.random.default_rng(1)
    rows = rng.standard_normal((400, 4)).astype(np.float32)
    idx = _mk(4, [0, 1, 2], shard_rows=10)
    assert idx.is_empty() and idx.max_id() is None and idx.len() == 0 and idx.num_shards() == 3
    for i in range(25):
        idx.add(vl.Vector(100 + i, rows[i], f"t{i}", None))
    assert idx.shard_sizes() == [10, 10, 5]
    assert [s.pos_base for s in idx._shards] == [0, 10, 20]          # global position of each shard's first row
    with pytest.raises(ValueError, match="already exists"):          # flat.rs:87, over the whole store
        idx.add(vl.Vector(103, rows[0]))
    with pytest.raises(ValueError, match="dimension"):               # flat.rs:84
        idx.add(vl.Vector(999, [1.0]))
    idx.delete(777)                                                  # flat.rs:93-96: missing id is Ok
    idx.delete(104)
    assert idx.shard_sizes() == [9, 10, 5] and [s.pos_base for s in idx._shards] == [0, 9, 19]
    assert idx.shard_of(104) is None and idx.get_vector(104) is None and idx.max_id() == 124
    idx.add(vl.Vector(200, rows[200]))                               # the tail does not move back into freed room
    assert idx.shard_sizes() == [9, 10, 6]
    for i in range(4):
        idx.add(vl.Vector(201 + i, rows[201 + i]))
    assert idx.shard_sizes() == [9, 10, 10]
    old = list(idx._shards)
    idx.add(vl.Vector(300, rows[300]))                               # last shard full → even re-split, headroom
    assert all(s.closed for s in old)
    sizes = idx.shard_sizes()
    assert sum(sizes) == 30 and max(sizes) - min(sizes) <= 1 and idx._shard_rows >= 13
    want = [100 + i for i in range(25) if i != 4] + [200, 201, 202, 203, 204, 300]
    ids, erows = idx.export()                                        # global storage order survives everything
    assert list(map(int, ids)) == want
    assert np.array_equal(erows[0], rows[0]) and np.array_equal(erows[-1], rows[300])
    assert idx.get_vector(103).text == "t3"                          # text / metadata kept across the re-split
    bases = [s.pos_base for s in idx._shards]
    assert bases == [0, sizes[0], sizes[0] + sizes[1]]


def test_bulk_load_split_and_append():
    rng = np.random.default_rng(2)
    rows = rng.standard_normal((30000, 4)).astype(np.float32)
    ids = np.arange(30000, dtype=np.uint64)
    idx = _mk(4, [0, 1, 2, 3])
    idx.add_batch(ids[:20000], rows[:20000])                         # empty index: split evenly
    assert idx.shard_sizes() == [5000, 5000, 5000, 5000]
    small = _mk(4, [0, 1, 2, 3])
    small.add_batch(ids[:100], rows[:100])                           # tiny loads stay together
    assert small.shard_sizes() == [100, 0, 0, 0]
    with pytest.raises(ValueError, match="already exists"):
        idx.add_batch(ids[19990:20010], rows[19990:20010])
    with pytest.raises(ValueError, match="dimension"):
        idx.add_batch(ids[:2], rows[:2, :3])
    idx.add_batch(ids[20000:20010], rows[20000:20010])               # later rows go to the tail shard
    assert idx.shard_sizes() == [5000, 5000, 5000, 5010] and idx.shard_of(20005) == 3
    e_ids, e_rows = idx.export()
    assert np.array_equal(e_ids, ids[:20010]) and np.array_equal(e_rows, rows[:20010])

"""Host logic of the multi-GPU product path without a GPU: contiguous block shards, the (device,
lane) scheduler of DeviceGroup and the ordered gather, driven with stub device states that record what they were given."""
import numpy as np
import pandas as pd
import pytest

from glow_b200.engine import DeviceGroup, GenotypeBlock, contiguous_blocks


class _StubState:
    """Stands in for DeviceState: a block's "result" is its own first row, lanes hold one block each."""

    def __init__(self, device, n_lanes=2, log=None):
        self.device, self.n_lanes, self.log = device, n_lanes, log if log is not None else []
        self.busy = {}
        self.max_inflight = 0

    def submit(self, lane, block, contig=0, **kw):
        assert lane not in self.busy, 'lane reused before wait'
        self.busy[lane] = (block, contig)
        self.max_inflight = max(self.max_inflight, len(self.busy))
        self.log.append(('submit', self.device, lane, int(block.data[0, 0])))

    def wait(self, lane):
        block, contig = self.busy.pop(lane)
        self.log.append(('wait', self.device, lane, int(block.data[0, 0])))
        return {'id': int(block.data[0, 0]), 'contig': contig, 'device': self.device}

    def close(self):
        pass


def _group(n_dev, n_lanes=2):
    g = DeviceGroup.__new__(DeviceGroup)
    log = []
    g.states = [_StubState(d, n_lanes, log) for d in range(n_dev)]
    g.devices = list(range(n_dev))
    g._comm = None
    return g, log


def _blocks(n):
    return [(GenotypeBlock(np.full((1, 4), k, dtype=np.float64), 'f64'), k % 3) for k in range(n)]


@pytest.mark.parametrize('n_blocks,world', [(0, 1), (1, 4), (7, 2), (8, 8), (23, 4), (1000, 8)])
def test_contiguous_blocks_partition(n_blocks, world):
    ranges = [contiguous_blocks(n_blocks, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n_blocks
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize('n_dev,n_blocks', [(1, 5), (2, 9), (4, 3), (3, 17)])
def test_round_robin_pipeline_yields_in_submission_order(n_dev, n_blocks):
    g, log = _group(n_dev)
    got = list(g.run_pipelined(iter(_blocks(n_blocks))))
    assert [r['id'] for r in got] == list(range(n_blocks))
    assert [r['device'] for r in got] == [k % n_dev for k in range(n_blocks)]
    assert [r['contig'] for r in got] == [k % 3 for k in range(n_blocks)]
    if n_blocks >= 2 * n_dev:  # every device has both lanes busy at some point: copies overlap compute
        assert all(s.max_inflight == 2 for s in g.states)
    assert not any(s.busy for s in g.states)


@pytest.mark.parametrize('n_dev,n_blocks', [(2, 10), (4, 4), (4, 30), (8, 13)])
def test_contiguous_shards_feed_all_devices_and_keep_order(n_dev, n_blocks):
    g, log = _group(n_dev)
    blocks = _blocks(n_blocks)
    ranges = [contiguous_blocks(n_blocks, d, n_dev) for d in range(n_dev)]
    out = {}
    for d, seq, res in g.run_shards([iter(blocks[a:b]) for a, b in ranges]):
        assert res['device'] == d
        out[(d, seq)] = res['id']
    flat = [out[k] for k in sorted(out)]
    assert flat == list(range(n_blocks))  # shard-major concatenation = the single-GPU order
    # all devices are started before any device receives its second block (they run concurrently)
    first_submits = [e for e in log if e[0] == 'submit'][:min(n_dev, n_blocks)]
    assert len({e[1] for e in first_submits}) == len(first_submits)
    assert not any(s.busy for s in g.states)


def test_sources_expose_block_ranges(tmp_path):
    import glow_b200
    rg = np.random.default_rng(1)
    X = rg.standard_normal((23, 9))
    src = glow_b200.ArraySource(X, block_variants=4)
    assert src.n_blocks() == 6
    got = [b.data for _, b in src.blocks(2, 5)]
    assert np.array_equal(np.concatenate(got), X[8:20])
    N, G = 101, 50
    rb = (N + 3) // 4
    body = rg.integers(0, 256, (G, rb), dtype=np.uint8)
    path = tmp_path / 'x.bed'
    path.write_bytes(bytes([0x6c, 0x1b, 0x01]) + body.tobytes())
    bed = glow_b200.PlinkBedSource(str(path), N, block_variants=8, read_threads=2)
    assert bed.n_blocks() == 7
    # two shards of one file streamed at the same time own separate staging rings
    it_a, it_b = bed.blocks(0, 3), bed.blocks(3, 7)
    rows_a, rows_b = [], []
    for (ma, a), (mb, b) in zip(it_a, it_b):
        rows_a.append(a.data.copy())
        rows_b.append(b.data.copy())
        assert ma.index[0] != mb.index[0]
    rows_b += [b.data.copy() for _, b in it_b]
    assert np.array_equal(np.concatenate(rows_a), body[:24]) and np.array_equal(np.concatenate(rows_b), body[24:])

"""CPU tier: the host logic of bench.py's north-star batch -- tiling, layout, strong-scaling partition, the tiled
comparison -- and both arms' identical `config` object."""
import numpy as np
import torch

import bench


def test_ns_layout_and_partition():
    rng = np.random.default_rng(1)
    d = 64
    comp = rng.integers(30_000, 500_000, size=d)
    plain = rng.integers(65_536, 1 << 20, size=d)
    for total, world in ((1024, 1), (1024, 2), (1000, 4), (1024, 8)):
        in_off, out_off, ranges = bench.ns_layout(comp, plain, total, world)
        assert len(in_off) == total + 1 and in_off[0] == 0 and out_off[0] == 0
        assert (np.diff(in_off.astype(np.int64)) == np.tile(comp, -(-total // d))[:total]).all()
        assert (out_off % 16 == 0).all()
        assert (np.diff(out_off.astype(np.int64)) >= np.tile(plain, -(-total // d))[:total]).all()
        # contiguous, complete, equal shares of the compressed bytes to within one stream
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
        share = [int(in_off[hi] - in_off[lo]) for lo, hi in ranges]
        assert max(share) - min(share) <= 2 * int(comp.max())
    # whole repetitions of the distinct set: the cuts fall on repetition boundaries and every shard starts 16-aligned
    in_off, out_off, ranges = bench.ns_layout(comp, plain, 16 * d, 8)
    assert [lo for lo, _ in ranges] == [2 * d * r for r in range(8)]


def test_tile_slice_and_equal():
    base = torch.arange(1000, dtype=torch.int64).to(torch.uint8)
    virt = base.repeat(5)
    for a, b in ((0, 1000), (0, 5000), (17, 4321), (999, 1001), (2000, 2000), (2500, 2501)):
        t = bench.tile_slice(torch, base, a, b, pad=32)
        assert t.numel() == b - a + 32 and torch.equal(t[:b - a], virt[a:b]) and int(t[b - a:].sum()) == 0
        assert bench.tile_equal(torch, t, base, a, b)
        if b > a:
            t[(b - a) // 2] ^= 1
            assert not bench.tile_equal(torch, t, base, a, b)


def test_both_arms_state_the_same_config():
    c = bench.ns_config(65536, 4096)
    assert set(c) == {"workload", "streams_total", "distinct_streams"}
    assert "65536" in c["workload"] and "4096 distinct" in c["workload"] and "strong-scaled" in c["workload"]
    # sizes are a function of the seed only: every rank (and the reference arm) derives the same batch
    assert [bench.ns_plain_len(j) for j in range(5)] == [bench.ns_plain_len(j) for j in range(5)]
    assert all(65536 <= bench.ns_plain_len(j) <= 1 << 20 for j in range(200))

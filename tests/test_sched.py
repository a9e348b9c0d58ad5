"""CPU tier: the stream-placement planner (lzma_rs_b200/csrc/lzb_sched.h, host-only C++) -- every stream is scheduled
exactly once, batches of similar streams keep the plain queue, a north-star size mix gets less crowded SMs for its
longest streams, and the C++ launch model agrees with its Python twin (tools/sched_sim.py)."""
import os
import sys

import numpy as np
import pytest

import emul_py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sched_sim  # noqa: E402

PARK = 0xFFFFFFFF


def _ns_work(n=8192):
    return np.sort(sched_sim.ns_sizes(n, min(n, 4096)) / 65536.0)[::-1]


def test_uniform_batches_keep_the_plain_queue():
    for n in (1, 100, 4096, 8192, 20000):
        order, info = emul_py.sched_plan(np.full(n, 3.0))
        assert not info["throttled"] and info["n_static"] == 0 and info["parked"] == 0
        assert order.tolist() == list(range(n))
        assert info["grid"] == max(1, min(148, (n + 27) // 28))


def test_size_mix_parks_warps_next_to_the_longest_streams():
    work = _ns_work()
    order, info = emul_py.sched_plan(work)
    assert info["throttled"] and info["n_static"] == info["grid"] * 28
    real = order[order != PARK]
    assert sorted(real.tolist()) == list(range(len(work)))          # every stream exactly once
    assert int((order == PARK).sum()) == info["parked"] > 0
    assert (order[info["n_static"]:] != PARK).all()                  # sentinels only in the pre-assigned part
    assert info["parked"] <= 0.15 * 148 * 28
    assert info["predicted"] < 0.97 * info["plain"]
    # CTA 0 holds the longest streams and the fewest warps
    first = order[:28]
    assert first[0] == 0 and (first == PARK).sum() >= (order[28 * 100: 28 * 101] == PARK).sum()


def test_cpp_model_matches_python_twin():
    work = _ns_work()
    t_cpp = emul_py.sched_simulate(work)
    t_py = sched_sim.simulate(work) / sched_sim.t_unit(28)
    assert t_cpp == pytest.approx(t_py, rel=1e-9)
    order, info = emul_py.sched_plan(work)
    counts = [(order[c * 28:(c + 1) * 28] != PARK).sum() for c in range(info["grid"])]
    queue = work[order[order != PARK]]
    assert emul_py.sched_simulate(queue, counts) == pytest.approx(info["predicted"], rel=1e-9)


def test_small_or_single_round_batches_are_never_throttled():
    rng = np.random.default_rng(5)
    for n in (10, 1000, 4144):
        w = np.sort(rng.random(n) * 15 + 1)[::-1]
        _, info = emul_py.sched_plan(w)
        assert not info["throttled"]

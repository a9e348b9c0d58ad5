"""CPU tier: the product's host planning code (lzb_plan.cpp) + K1's decode core compiled as plain C++ (1-lane warp,
tests/host_emulation) against the oracle, on the whole parity corpus.  The same corpus runs on the real GPU path in
tests/test_gpu_parity.py; this tier exists so the decode logic and the container walk are checked on every CPU run."""
import pytest

import cases
import emul_py
import parity
from lzma_rs_b200 import _native


def _decode(fmt, streams, opts):
    opt = _native.make_options(opts.get("unpacked_mode", 0), opts.get("provided"), opts.get("memlimit"))
    res = emul_py.decode_batch(fmt, streams, opt)
    # end-marker .lzma has no size bound: retry on LZB_E_CAPACITY like lzb_decompress_alloc does
    for i, r in enumerate(res):
        cap = None
        while int(r.status["code"]) == _native.E_CAPACITY:
            cap = max(int(r.status["a0"]) * 2, 1 << 16) if cap is None else cap * 2
            r = emul_py.decode_batch(fmt, [streams[i]], opt, capacities=[cap])[0]
        res[i] = r
    return res


@pytest.mark.parametrize("family", ["valid_lzma2_cases", "valid_lzma_cases", "hand_encoded_cases",
                                    "truncation_and_corruption_cases", "xz_cases", "xz_chain_cases"])
def test_emulated_kernel_matches_oracle(family):
    bad = []
    n = 0
    for (fmt, okey), named in parity.group_cases(getattr(cases, family)()).items():
        bad += parity.check_group(_decode, fmt, dict(okey), named)
        n += len(named)
    assert not bad, f"{len(bad)}/{n} mismatches:\n" + "\n".join(bad[:40])


@pytest.mark.parametrize("family", ["valid_lzma2_cases", "valid_lzma_cases", "hand_encoded_cases",
                                    "truncation_and_corruption_cases", "xz_cases", "xz_chain_cases"])
def test_emulated_latency_kernel_matches_oracle(family, monkeypatch):
    """The latency form of K1 (LAT: look-ahead tree walks, probabilities fetched ahead of their decisions, whole literal
    table in "shared memory") compiled as plain C++: same corpus, same oracle."""
    monkeypatch.setenv("LZB_EMUL_LAT", "1")
    test_emulated_kernel_matches_oracle(family)


def test_structured_fuzz_latency_kernel(monkeypatch):
    monkeypatch.setenv("LZB_EMUL_LAT", "1")
    bad = structured_fuzz(_decode, 20261018, 300)
    assert not bad, f"{len(bad)} mismatches:\n" + "\n".join(bad[:20])


def fuzz_regressions():
    """Inputs on which tools/fuzz_soak.py once found a mismatch (tests/golden/fuzz_regressions/<seed>-r<round>-f<fmt>-<i>.bin)."""
    import glob
    import os
    import re
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuzz_regressions")
    out = {}
    for path in sorted(glob.glob(os.path.join(d, "*.bin"))):
        fmt = int(re.search(r"-f(\d)-", os.path.basename(path)).group(1))
        out.setdefault(fmt, []).append((os.path.basename(path), open(path, "rb").read()))
    return out


def test_fuzz_regressions():
    regs = fuzz_regressions()
    assert regs
    for fmt, named in regs.items():
        bad = parity.check_group(_decode, fmt, {}, named)
        assert not bad, "\n".join(bad)


def structured_fuzz(decode, seed, n):
    """A fixed-seed slice of tools/fuzz_soak.py's structure-aware generators: .xz files assembled field by field with
    valid CRCs around odd values, and .lzma / LZMA2 streams built symbol by symbol with a real range encoder."""
    import os
    import random
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import corpus
    import fuzz_soak
    rnd = random.Random(seed)
    bad = parity.check_group(decode, 2, {}, [(f"xz-{i}", fuzz_soak.xz_structured(rnd, corpus)) for i in range(n)])
    gen = [fuzz_soak.lzma_structured(rnd, corpus) for _ in range(2 * n)]
    for fmt, opts in ((0, {}), (0, {"unpacked_mode": 1, "provided": 7}), (0, {"memlimit": 100}), (1, {})):
        bad += parity.check_group(decode, fmt, opts, [(f"l{fmt}-{i}", s) for i, (f, s) in enumerate(gen) if f == fmt])
    return bad


def test_structured_fuzz():
    bad = structured_fuzz(_decode, 20261017, 500)
    assert not bad, f"{len(bad)} mismatches:\n" + "\n".join(bad[:20])

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

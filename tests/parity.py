"""Shared parity checker: a decode implementation (GPU library or CPU-tier emulation) against the oracle."""
import collections

import oracle_py as oracle

ORACLE = {0: lambda s, o: oracle.lzma_decompress(s, o.get("unpacked_mode", 0), o.get("provided"), o.get("memlimit")),
          1: lambda s, o: oracle.lzma2_decompress(s),
          2: lambda s, o: oracle.xz_decompress(s)}


def group_cases(cases):
    """Groups (name, fmt, stream, opts) by (fmt, opts) so each group is ONE batch call."""
    groups = collections.OrderedDict()
    for name, fmt, stream, opts in cases:
        key = (fmt, tuple(sorted(opts.items())))
        groups.setdefault(key, []).append((name, stream))
    return groups


def check_group(decode_batch, fmt, opts, named_streams):
    """decode_batch(fmt, streams, opts_dict) -> list of results with .data .consumed .display .status.
    Returns list of mismatch descriptions (empty = parity)."""
    streams = [s for _, s in named_streams]
    results = decode_batch(fmt, streams, opts)
    bad = []
    for (name, s), r in zip(named_streams, results):
        ref = ORACLE[fmt](s, opts)
        code = int(r.status["code"])
        if code < 0:  # capacity / unsupported: never acceptable in these tests
            bad.append(f"{name}: internal status {code}: {r.display}")
            continue
        if r.display != ref.display:
            bad.append(f"{name}: display {r.display!r} != oracle {ref.display!r}")
        elif r.data != ref.out:
            bad.append(f"{name}: output differs (len {len(r.data)} vs oracle {len(ref.out)})")
        elif ref.ok and r.consumed != ref.consumed:
            bad.append(f"{name}: consumed {r.consumed} != oracle {ref.consumed}")
    return bad

#!/usr/bin/env python3
"""tools/check_sass.py -- guards K1 against a ptxas heuristic flip.

K1 is bound by instruction issue on the vector ALU/FMA pipes.  Its control flow is warp-uniform by construction, and
for some source shapes ptxas moves the whole bit loop onto the *uniform datapath* (UIADD3/USEL/UISETP/UIMAD ...), whose
single pipe per SM sub-partition then throttles the kernel (measured on B200: C3 238 ms -> 339 ms, issue-active 85 % ->
65 %, math_pipe_throttle 12 % -> 22 %; profiles/r01_uniform_flip.md).  The flip shows in the static opcode mix, so the
build checks it:   python tools/check_sass.py lzma_rs_b200/liblzma_b200.so [max_share]
prints size / registers-independent uniform-op share per decode kernel and exits 1 if any exceeds max_share (0.12).

Second guard (round 2): reconvergence barriers.  K1's control flow is warp-uniform; when ptxas's divergence analysis
loses that (a called function's result or stream-index dependent code around decode_item in the plain-queue loop), it
wraps every branch of the bit loop in BSSY / BSYNC pairs: 10 -> 78 barriers and +20 % instructions in the kernel
(round-1's non-sched mirror kernel had 87).  More than MAX_BSSY (32) in a decode kernel fails the build; the lc+lp > 4
kernel (rare path, literal table in global memory), the single-stream kernel of the raw decoders and the `fill`
instantiations (run-length batches: there the dist-1 shortcut that triggers the barriers is measured 13 % faster) are exempt.
"""
import re
import subprocess
import sys


MAX_BSSY = 32


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, res = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            res[cur].append(m.group(1).split(".")[0])
    return res


def main():
    lib = sys.argv[1]
    limit = float(sys.argv[2]) if len(sys.argv) > 2 else 0.12
    bad = 0
    for name, ops in kernels(lib).items():
        if not name.startswith("lzb_decode"):
            continue
        ops = [o for o in ops if o != "NOP"]
        u = sum(1 for o in ops if o.startswith("U") or o in ("R2UR", "S2UR", "LDCU", "VOTEU"))
        share = u / max(1, len(ops))
        bssy = sum(1 for o in ops if o == "BSSY")
        flag = "" if share <= limit else "   <-- uniform-datapath flip"
        if bssy > MAX_BSSY and not any(x in name for x in ("biglit", "carry", "fill")):
            flag += "   <-- reconvergence barriers in the bit loop"
            bad += 1
        bad += share > limit
        print(f"{name:34s} {len(ops):5d} instructions ({len(ops) * 16 / 1024:5.1f} KB)  uniform-datapath ops {100 * share:5.1f} %"
              f"  BSSY {bssy:3d}{flag}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

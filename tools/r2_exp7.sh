#!/bin/bash
O=gpurun_out
mkdir -p $O
V=build/variants
timeout 900 python tools/kbench.py --config c2 --steps 7 $V/r2c.so $V/r2c_pin.so $V/r2c.so $V/r2c_pin.so > $O/r2_exp7_c2.txt 2>&1
timeout 900 python tools/kbench.py --config ns --streams 8192 --steps 3 $V/r2c.so $V/r2c_pin.so > $O/r2_exp7_ns.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_abi2.py tests/test_gpu_parity.py -x -q -k "gated or drain or peer or pinned or multi or stored_only or batch_shape or config4" 2>&1 | tail -8 > $O/r2_exp7_pytest.txt
for mode in "" "LZB_GATE_BYTES=1" "LZB_DRAIN_ROUNDS=2" "LZB_DRAIN_ROUNDS=2 LZB_GATE_BYTES=1"; do
  ( env $mode LZB_TRACE=1 timeout 900 python bench.py --streams 8192 --distinct 512 --steps 3 --warmup 3 --no-verify ) > $O/r2_exp7_small_$(echo $mode | tr ' =' '__').json 2> $O/r2_exp7_small_$(echo $mode | tr ' =' '__').err
done
( LZB_TRACE=1 timeout 1500 python bench.py --steps 3 --warmup 3 ) > $O/r2_exp7_ns_full.json 2> $O/r2_exp7_ns_full.err
( LZB_GATE_BYTES=1 LZB_TRACE=1 timeout 1500 python bench.py --steps 3 --warmup 3 ) > $O/r2_exp7_ns_full_bytes.json 2> $O/r2_exp7_ns_full_bytes.err
cat $O/r2_exp7_c2.txt $O/r2_exp7_ns.txt $O/r2_exp7_pytest.txt
for f in $O/r2_exp7_small_*.err $O/r2_exp7_ns_full*.err; do echo $f; grep lzb_trace $f | tail -2; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_exp7_*.json")):
    try:
        t=open(f).read(); j=json.loads(t[t.index('{"metric"'):].splitlines()[0])
        print(f, "value %.3f (%.1f ms)  e2e %.3f (%.1f ms)"%(j["value"],j["ms_per_step"],j["e2e"]["value"],j["e2e"]["ms_per_step"]))
    except Exception as e: print(f,"ERR",e)
PY

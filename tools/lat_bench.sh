#!/bin/bash
# tools/lat_bench.sh -- latency kernels (lzb_decode_lat_*) against the throughput kernels (LZB_NO_LAT=1) at small stream
# counts, and one --set full capture of each family with one warp per SM.  Results: profiles/r02_k1_lat_vs_default.txt,
# profiles/r02_k1_{lat,default}_148streams.txt (python profiles/summarize_ncu.py gpurun_out/lat_*.ncu-rep).
O=gpurun_out
mkdir -p $O
L=lzma_rs_b200/liblzma_b200.so
for n in 1 148 592 1024 1184; do
  echo "== streams $n"
  timeout 300 python tools/kbench.py --config c2 --streams $n --steps 7 $L $L@LZB_NO_LAT=1 2>&1 | tail -n 2
done > $O/lat_kbench.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lzb_decode_lat_kernel -s 3 -c 1 -f -o $O/lat_lat148 \
  python tools/kbench.py --config c2 --streams 148 --steps 2 $L > $O/lat_lat148.log 2>&1
LZB_NO_LAT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lzb_decode_kernel -s 3 -c 1 -f -o $O/lat_def148 \
  python tools/kbench.py --config c2 --streams 148 --steps 2 $L > $O/lat_def148.log 2>&1
cat $O/lat_kbench.txt

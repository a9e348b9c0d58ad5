#!/bin/bash
O=gpurun_out
mkdir -p $O
L=lzma_rs_b200/liblzma_b200.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lzb_decode_lat_kernel -s 3 -c 1 -f -o $O/lat2_lat148 \
  python tools/kbench.py --config c2 --streams 148 --steps 2 $L > $O/lat2_lat148.log 2>&1
LZB_NO_LAT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lzb_decode_kernel -s 3 -c 1 -f -o $O/lat2_def148 \
  python tools/kbench.py --config c2 --streams 148 --steps 2 $L > $O/lat2_def148.log 2>&1
tail -3 $O/lat2_lat148.log $O/lat2_def148.log

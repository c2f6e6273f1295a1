#!/bin/bash
mkdir -p gpurun_out
echo "== step trace"; DVD_LIB=dvd_b200/libdvd_b200_trace.so DVD_NO_GRAPH=1 timeout 300 python tools/step_trace.py 2048 1536 4608 384 2>&1 | tail -56

#!/bin/bash
# build the library with the trace hooks in a scratch copy of csrc, run the timeline, restore the normal build
cd flowgnn_b200/csrc && cp ../libflowgnn_b200.so /tmp/lib_keep.so && rm -f build/gin_fused.o build/gin_tc2.o build/api.o && make -j8 EXTRA=-DFG_TC2_TRACE > /dev/null 2>&1; cd ../..
timeout 300 python tools/trace_fused.py 2>&1 | tee gpurun_out/${1:-r2f}_trace.txt
cp /tmp/lib_keep.so flowgnn_b200/libflowgnn_b200.so

#!/bin/bash
mkdir -p gpurun_out
cp dfd-clip_b200/libdfdclip_b200.so /tmp/lib_keep.so
cp dfd-clip_b200/libdfdclip_b200_trace.so dfd-clip_b200/libdfdclip_b200.so
: > gpurun_out/mha_trace.log
for v in ${VARIANTS:-1}; do
  echo "== variant $v" >> gpurun_out/mha_trace.log
  DFD_MHA_VARIANT=$v python tools/mha_trace.py >> gpurun_out/mha_trace.log 2>&1
done
cp /tmp/lib_keep.so dfd-clip_b200/libdfdclip_b200.so
cat gpurun_out/mha_trace.log

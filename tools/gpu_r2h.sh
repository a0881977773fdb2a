#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2h_dec_attn.log
cp dfd-clip_b200/libdfdclip_b200.so /tmp/lib_new.so
for rep in 1 2; do
  for which in old new; do
    if [ $which = old ]; then cp dfd-clip_b200/libdfdclip_b200_old.so dfd-clip_b200/libdfdclip_b200.so; else cp /tmp/lib_new.so dfd-clip_b200/libdfdclip_b200.so; fi
    echo "$which: $(timeout 120 python tools/prof_kernel.py dec_attn 50 2>&1 | tail -1)" >> gpurun_out/r2h_dec_attn.log
  done
done
cp /tmp/lib_new.so dfd-clip_b200/libdfdclip_b200.so
ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:dec_attn_stream -c 2 --csv --log-file gpurun_out/r2h_ncu.csv python tools/prof_kernel.py dec_attn 2 > /dev/null 2>&1
grep -E "dec_attn_stream" gpurun_out/r2h_ncu.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tail -5 >> gpurun_out/r2h_dec_attn.log
cat gpurun_out/r2h_dec_attn.log

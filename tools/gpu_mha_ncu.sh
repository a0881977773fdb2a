#!/bin/bash
mkdir -p gpurun_out
for v in ${VARIANTS:-1 4}; do
  DFD_MHA_VARIANT=$v ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second \
    --clock-control none -k regex:mha_fwd -c 3 --csv --log-file gpurun_out/mha_ncu_v$v.csv python tools/prof_kernel.py mha 3 > gpurun_out/mha_ncu_v$v.log 2>&1
  python - "$v" <<'PY'
import csv,sys
v=sys.argv[1]
rows=[r for r in csv.reader(l for l in open('gpurun_out/mha_ncu_v%s.csv'%v) if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
last=max(r[ii] for r in rows[1:])
for r in rows[1:]:
    if r[ii]==last: print(v, r[ki][:40], r[mi], r[vi])
PY
done

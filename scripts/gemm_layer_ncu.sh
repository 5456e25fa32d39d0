#!/bin/bash
# Profiling aid: duration, tensor-pipe activity and L1 hit rate of the four projection GEMMs of encoder layer 1
# (one ncu pass over `bench.py --only encode`; run under gpurun after the plain command has succeeded).
python bench.py --only encode --no-cpu --steps 2 --warmup 1 > gpurun_out/enc_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct,sm__cycles_active.avg \
  --clock-control none -k regex:"gemm_tc2|attention_tc" -s 5 -c 5 --csv --log-file gpurun_out/gemm_layer.csv \
  python bench.py --only encode --no-cpu --steps 2 --warmup 1 > gpurun_out/gemm_layer_ncu.log 2>&1
python - <<'PY'
import csv, json
rows = [r for r in csv.reader(open("gpurun_out/gemm_layer.csv")) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
out = {}
for r in rows[1:]:
    key = (r[ix["ID"]], r[ix["Kernel Name"]][:60])
    out.setdefault(key, {})[r[ix["Metric Name"]]] = r[ix["Metric Value"]]
for (i, name), m in out.items():
    print(i, name, " ".join(f"{k.split('.')[0].split('__')[-1]}={v}" for k, v in m.items()))
print(json.loads(open("gpurun_out/enc_plain.log").read().strip().splitlines()[-1])["extra"]["encode"]["chunks_per_s"]
      if False else open("gpurun_out/enc_plain.log").read()[-4000:].split('"chunks_per_s": ')[1].split(",")[0])
PY

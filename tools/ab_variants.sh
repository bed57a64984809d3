#!/bin/bash
# A/B of tuning builds on one GPU: tools/ab_variants.sh base p1 p2 ...  (libraries from `python __graft_entry__.py --variant NAME DEFS`)
# For every variant: the configs[1] bench line (resident + roofline), configs[2] training step, configs[3] shard.
mkdir -p gpurun_out
for v in "$@"; do
  export B2LDDMM_LIB=$PWD/build/variants/$v/libb2lddmm.so
  [ "$v" = main ] && unset B2LDDMM_LIB
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ab_${v}_bench.json 2> gpurun_out/ab_${v}_bench.err
  timeout 300 python tools/bench_configs.py c3f > gpurun_out/ab_${v}_c3f.json 2>&1
  timeout 300 python tools/bench_configs.py c4 > gpurun_out/ab_${v}_c4.json 2>&1
  python - <<PY
import json
def last(p):
    try:
        return json.loads([l for l in open(p) if l.startswith("{")][-1])
    except Exception as e:
        return {"err": str(e)}
b = last("gpurun_out/ab_${v}_bench.json"); c3 = last("gpurun_out/ab_${v}_c3f.json"); c4 = last("gpurun_out/ab_${v}_c4.json")
print("AB ${v}: fwd kernel_ms", b.get("roofline", {}).get("kernel_ms"), "frac", b.get("roofline", {}).get("frac"), "step_ms", b.get("ms_per_step"),
      "| c3f ms", c3.get("ms_per_step"), "| c4 ms", c4.get("ms_per_step"))
PY
done

#!/bin/bash
# data-parallel knob sweep: tools/dp_sweep.sh N "ENV1=.. ENV2=.. -- bench flags" ...
# each spec: environment assignments, then "--", then extra bench.py flags
N=$1; shift
mkdir -p gpurun_out/r02
i=0
for spec in "$@"; do
  envs="${spec%%--*}"; flags="${spec#*--}"
  port=$((29700 + i)); i=$((i + 1))
  out=gpurun_out/r02/sweep_n${N}_$i.json
  env $envs python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port $port bench.py --gpus $N --steps 40 --warmup 5 $flags > $out 2> ${out%.json}.err
  python - "$out" "$spec" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%-60s ms=%.3f value=%.1f e2e=%.1f" % (sys.argv[2], d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print("%-60s FAILED %s" % (sys.argv[2], e))
PY
done

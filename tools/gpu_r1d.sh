#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out

timeout 900 python bench.py --no-cpu > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], json.dumps(d["roofline"]["kernels"], indent=0), d.get("e2e"))
PY

#!/bin/bash
# dev helper run on the GPU box: full GPU test suite + a bench line (tag = $1)
tag=${1:-x}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1
tail -3 gpurun_out/${tag}_tests.log | cut -c1-300
python bench.py --no-cpu > gpurun_out/bench_${tag}.log 2>&1
python - <<PY
import json
d = json.loads([x for x in open("gpurun_out/bench_${tag}.log") if x.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "share", d["roofline"]["kernel_share_of_step"],
      "grint", d["secondary"]["value"], "n512", d["secondary_n512"]["value"])
PY

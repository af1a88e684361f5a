#!/bin/bash
# One short GPU-box call: the GPU parity suite and one bench line (about 45 s of box time).
#   gpurun --timeout 400 -- 'bash tools/gpu_check.sh'
mkdir -p gpurun_out/check
python -m pytest tests -m gpu -x -q > gpurun_out/check/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/check/pytest.log
tail -3 gpurun_out/check/pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/check/bench.json 2> gpurun_out/check/bench.err
python - gpurun_out/check/bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["stage_ms"].items()}, "e2e", round(d["e2e"]["build_s"] * 1e3, 1),
      "launches", d["gpu_launches"], "frac", round(d["roofline"]["frac"], 3))
PY

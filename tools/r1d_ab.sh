mkdir -p gpurun_out/r1d
python -m pytest tests -m gpu -x -q > gpurun_out/r1d/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1d/pytest.log
tail -3 gpurun_out/r1d/pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1d/bench_new.json 2> gpurun_out/r1d/bench_new.err
python - gpurun_out/r1d/bench_new.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()}, "e2e", round(d["e2e"]["build_s"]*1e3,1), "launches", d["gpu_launches"])
print(d["roofline"])
PY

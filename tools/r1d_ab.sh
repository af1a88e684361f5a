mkdir -p gpurun_out/r1d
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 > gpurun_out/r1d/bench_final_n1.json 2> gpurun_out/r1d/bench_final_n1.err; echo "bench rc $?"
timeout 200 python bench.py --workload nio-afm-standin-k444 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1d/bench_nio_n1.json 2> gpurun_out/r1d/bench_nio_n1.err; echo "nio rc $?"
timeout 150 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r1d/bench_reference_arm.json 2> gpurun_out/r1d/bench_reference_arm.err; echo "ref rc $?"
for f in bench_final_n1 bench_nio_n1 bench_reference_arm; do python - gpurun_out/r1d/$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d.get("ms_per_step",0),2), {k:round(v,2) for k,v in d.get("stage_ms",{}).items()}, "value", round(d.get("value",0),1), "e2e", d.get("e2e"), "cpu", d.get("cpu_baseline"))
except Exception as e:
    print("ERR", sys.argv[1], e)
PY
done

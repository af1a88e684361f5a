mkdir -p gpurun_out/r1d
python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1d/bench_n2.json 2> gpurun_out/r1d/bench_n2.err
tail -c 1500 gpurun_out/r1d/bench_n2.json

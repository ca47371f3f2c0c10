timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python scripts_tmp/defer_probe.py 2>&1 | tail -7
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | cut -c1-330
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/b.log 2>&1

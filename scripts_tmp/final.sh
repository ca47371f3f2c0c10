mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tc_chunk_kernel -s 3 -c 1 -o gpurun_out/prof_tc -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_run.log 2>&1
echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b.log 2>&1
echo "ncu2 rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 300 gpurun_out/bench_ref.json
python tests/shape_bench.py > gpurun_out/shapes.jsonl 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv > gpurun_out/clocks.txt

python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -c . gpurun_out/launches_bench_r1.csv

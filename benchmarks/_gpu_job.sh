python benchmarks/sanitize_small.py > gpurun_out/san_plain.log 2>&1 && timeout 800 compute-sanitizer --tool memcheck --print-limit 20 python benchmarks/sanitize_small.py > gpurun_out/san_memcheck.log 2>&1
tail -n 3 gpurun_out/san_plain.log; tail -n 15 gpurun_out/san_memcheck.log

python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --windows 32768 --steps 2 --warmup 1 --no-config2 > gpurun_out/r02_launches_final.log 2>&1; echo ncu rc=$?
compute-sanitizer --tool memcheck python tools/sanitizer_workload.py 0.5 2>&1 | tail -8 > gpurun_out/r02_sanitizer_memcheck.txt; echo memcheck rc=$?
compute-sanitizer --tool racecheck python tools/sanitizer_workload.py 0.5 2>&1 | tail -8 > gpurun_out/r02_sanitizer_racecheck.txt; echo racecheck rc=$?
tail -n 2 gpurun_out/r02_sanitizer_memcheck.txt gpurun_out/r02_sanitizer_racecheck.txt

# ncu --set full captures of the hot kernels (one launch each, after a warm-up run) + racecheck; outputs under gpurun_out/
N="ncu --set full --clock-control none --import-source on -f"
$N --kernel-name-base demangled -k "regex:Tier<unsigned char" --launch-skip 3 -c 2 -o gpurun_out/r02f_poa python tools/profile_workload.py 8192 150 2 > gpurun_out/r02f_poa.log 2>&1; echo poa rc=$?
$N -k regex:k_index --launch-skip 1 -c 1 -o gpurun_out/r02f_index python tools/profile_workload.py 8192 150 2 > gpurun_out/r02f_index.log 2>&1; echo index rc=$?
$N --kernel-name-base demangled -k "regex:Tier<unsigned short, 2, 1024" --launch-skip 2 -c 1 -o gpurun_out/r02f_w1 python tools/profile_workload.py 4000 20 2 > gpurun_out/r02f_w1.log 2>&1; echo w1 rc=$?
$N -k "regex:^k_reanchor$" --launch-skip 1 -c 1 -o gpurun_out/r02f_reanchor python tools/reanchor_bench.py 1500 > gpurun_out/r02f_reanchor.log 2>&1; echo reanchor rc=$?
compute-sanitizer --tool racecheck python tools/sanitizer_workload.py 0.5 2>&1 | tail -8 > gpurun_out/r02_sanitizer_racecheck.txt; echo racecheck rc=$?

set -x
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke2.py > gpurun_out/san2_mem.txt 2>&1; tail -3 gpurun_out/san2_mem.txt
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_smoke2.py > gpurun_out/san2_race.txt 2>&1; tail -3 gpurun_out/san2_race.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nr2d1_kernel|strain_kernel" -c 3 -o gpurun_out/r1_nr_strain python tools/bench_rows.py > gpurun_out/ncu_rows.log 2>&1; tail -2 gpurun_out/ncu_rows.log
ncu -i gpurun_out/r1_nr_strain.ncu-rep --page details > gpurun_out/r1_ncu_details_nr2d1_strain_configB.txt 2>&1; ls -la gpurun_out | tail -5

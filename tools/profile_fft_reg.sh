ncu --set full --clock-control none --import-source on -k regex:"fftcc2d_reg" -s 2 -c 1 -o gpurun_out/r1_fftcc2d_reg40 python bench.py --config C --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/r1_fftcc2d_reg40.ncu-rep --page details > gpurun_out/r1_ncu_details_fftcc2d_reg40_configC.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fftcc3d_reg" -s 1 -c 1 -o gpurun_out/r1_fftcc3d_reg60 python bench.py --config F --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/r1_fftcc3d_reg60.ncu-rep --page details > gpurun_out/r1_ncu_details_fftcc3d_reg60_configF.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"icgn|fftcc|gradient3d|prefilter3d" -c 100 --csv --log-file gpurun_out/r1_launches_configC.csv python bench.py --config C --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"icgn|fftcc|gradient3d|prefilter3d" -c 100 --csv --log-file gpurun_out/r1_launches_configF.csv python bench.py --config F --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep -E "Duration|Issue Slots Busy|Mem Pipes Busy|Registers Per|Achieved Active Warps" gpurun_out/r1_ncu_details_fftcc2d_reg40_configC.txt gpurun_out/r1_ncu_details_fftcc3d_reg60_configF.txt | head -12

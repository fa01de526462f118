# Regenerates the profiles/ evidence on a GPU box (run through gpurun; outputs land in gpurun_out/)
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
for c in B C D E; do python bench.py --config $c 2>/dev/null | tail -1 > gpurun_out/r1_bench_config$c.json; done
python bench.py --impl reference 2>/dev/null | tail -1 > gpurun_out/r1_bench_configB_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"icgn|fftcc|gradient3d|prefilter3d" -c 400 --csv --log-file gpurun_out/r1_launches_configB.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"icgn|fftcc|gradient3d|prefilter3d" -c 400 --csv --log-file gpurun_out/r1_launches_configD.csv python bench.py --config D --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"icgn2d_kernel" -s 3 -c 1 -o gpurun_out/r1_icgn2d1_final python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/r1_icgn2d1_final.ncu-rep --page details > gpurun_out/r1_ncu_details_icgn2d1_configB_final.txt 2>&1
ncu -i gpurun_out/r1_icgn2d1_final.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys,csv,json
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
g=lambda k: float(r[h.index(k)].replace(',',''))
print(json.dumps({'kernel':'icgn2d_kernel<6,16,false,1> config B','dram_bytes_read':g('dram__bytes_read.sum'),'dram_bytes_write':g('dram__bytes_write.sum'),'unit':r[h.index('dram__bytes_read.sum')] and rows[1][h.index('dram__bytes_read.sum')]}))
" > gpurun_out/traffic_raw.json
cat gpurun_out/pytest_gpu.txt
# the reference's own 2D example (compiled unchanged against the shim) on its shipped image pair: its timing file next to the shipped one
rm -rf /tmp/ex && mkdir -p "/tmp/ex/d:/dic_tests/2d_dic" && cp tests/golden/oht_cfrp_0.bmp tests/golden/oht_cfrp_4.bmp "/tmp/ex/d:/dic_tests/2d_dic/"
(cd /tmp/ex && for i in 1 2; do $OLDPWD/examples/bin/test_2d_dic_fftcc_icgn1 < /dev/null > /tmp/ex/out.txt 2>&1; done; cat /tmp/ex/out.txt | head -5)
cp "/tmp/ex/d:/dic_tests/2d_dic/oht_cfrp_4_fftcc_icgn1_r16_time.csv" gpurun_out/r1_example_2d_dic_fftcc_icgn1_time.csv

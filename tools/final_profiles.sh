# Regenerates the round-2 evidence under profiles/ on a GPU box (run through gpurun; outputs land in gpurun_out/, copied to profiles/ here)
set -x
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/r2_pytest_gpu.txt
python bench.py > $O/r2_bench_configB.json 2> $O/r2_bench_configB.err
for c in C D E F A; do python bench.py --config $c --no-extras 2>/dev/null | tail -1 > $O/r2_bench_config$c.json; done
python bench.py --impl reference 2>/dev/null | tail -1 > $O/r2_bench_configB_reference_arm.json
OCB_ICGN2D_TMEM=0 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > $O/r2_bench_configB_smem_variant.json
OCB_NO_TMA=1 python bench.py --config D --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > $O/r2_bench_configD_no_tma.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"icgn|fftcc|gradient3d|prefilter3d" -c 400 --csv --log-file $O/r2_launches_configB.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"icgn|fftcc|gradient3d|prefilter3d" -c 400 --csv --log-file $O/r2_launches_configD.csv python bench.py --config D --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:icgn2d_kernel -s 3 -c 1 -o $O/r2_icgn2d1_tmem python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:icgn3d1_kernel -s 1 -c 1 -o $O/r2_icgn3d1_final python bench.py --config D --steps 1 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fftcc2d_w32 -s 3 -c 1 -o $O/r2_fftcc2d_w32_tma python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fftcc3d_w32 -s 1 -c 1 -o $O/r2_fftcc3d_w32_tma python bench.py --config D --steps 1 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
# the reference's own example programs (compiled unchanged against the shim) on their shipped data: the timing files they write
R=$PWD
rm -rf /tmp/ex && mkdir -p "/tmp/ex/d:/dic_tests/2d_dic" && cp tests/golden/oht_cfrp_0.bmp tests/golden/oht_cfrp_4.bmp "/tmp/ex/d:/dic_tests/2d_dic/"
(cd /tmp/ex && for i in 1 2 3; do python3 -c "
import subprocess,time,sys
t=time.time(); o=subprocess.run(['$R/examples/bin/test_2d_dic_fftcc_icgn1'],stdin=subprocess.DEVNULL,capture_output=True,text=True); print(o.stdout.strip().replace(chr(10),' | '), '| process wall %.3f s' % (time.time()-t))"; done) > $O/r2_example_2d_runs.txt 2>&1
cp "/tmp/ex/d:/dic_tests/2d_dic/oht_cfrp_4_fftcc_icgn1_r16_time.csv" $O/r2_example_2d_dic_fftcc_icgn1_time.csv
compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > $O/r2_sanitizer_memcheck.txt 2>&1
tail -3 $O/r2_sanitizer_memcheck.txt
cat $O/r2_pytest_gpu.txt $O/r2_example_2d_runs.txt

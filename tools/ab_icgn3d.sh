# A/B harness for icgn3d.cu build variants (opencorr_b200/lib/variants/*.so, selected with OCB_LIB_PATH), config D
run() { python bench.py --no-cpu-baseline --steps 5 --config D 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'step_ms', round(d['ms_per_step'],3), 'icgn_ms', round(d['roofline']['kernel_ms'],3), d['results']['iteration_histogram'][:8])"; }
run default
for v in opencorr_b200/lib/variants/*.so; do OCB_LIB_PATH=$PWD/$v run $(basename $v .so); done

run() { python bench.py --no-cpu-baseline --steps 20 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'step_ms', round(d['ms_per_step'],4), 'icgn_ms', round(d['roofline']['kernel_ms'],4), 'e2e', d['e2e']['ms_per_step_spread_rank0'])"; }
run base
for u in 1 2 6 11; do OCB_LIB_PATH=$PWD/opencorr_b200/lib/variants/u$u.so run unroll$u; done
for w in 4 6 8 10; do OCB_ICGN2D_MAX_WARPS=$w run warps$w; done

# A/B harness for icgn2d.cu: warps per POI (OCB_ICGN2D_WPP) and build variants (opencorr_b200/lib/variants/*.so via OCB_LIB_PATH)
run() { python bench.py --no-cpu-baseline --steps 20 --config ${2:-B} 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', '${2:-B}', 'step_ms', round(d['ms_per_step'],4), 'icgn_ms', round(d['roofline']['kernel_ms'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), d['results']['iteration_histogram'][:8])"; }
for c in B C E A; do
OCB_ICGN2D_WPP=1 run wpp1 $c
OCB_ICGN2D_WPP=2 run wpp2 $c
for v in opencorr_b200/lib/variants/*.so; do [ -f $v ] && OCB_LIB_PATH=$PWD/$v run $(basename $v .so) $c; done
done

mkdir -p "/tmp/ex/d:/dic_tests/2d_dic" && cp tests/golden/oht_cfrp_*.bmp "/tmp/ex/d:/dic_tests/2d_dic/"
R=$PWD
cd /tmp/ex
for i in 1 2 3; do s=$(date +%s.%N); $R/examples/bin/test_2d_dic_fftcc_icgn1 < /dev/null | grep takes; e=$(date +%s.%N); echo "wall $(echo "$e - $s" | bc) s"; done
cat "/tmp/ex/d:/dic_tests/2d_dic/oht_cfrp_4_fftcc_icgn1_r16_time.csv"

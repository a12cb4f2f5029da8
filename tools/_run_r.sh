OUT=gpurun_out; TAG=r02r
timeout 600 bash tools/ab_variants.sh "nowarm" 10 2 20 20 0 300 1 heat > $OUT/${TAG}_warm_n10_nx20_cp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "nowarm" 10 2 20 20 0 100 2 heat > $OUT/${TAG}_warm_n10_nx20_sp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "nowarm" 7 2 5 5 0 300 1 heat > $OUT/${TAG}_warm_n7_nx5_cp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "nowarm" 13 2 50 20 0 100 > $OUT/${TAG}_warm_n13_nx50_cp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "nowarm" 20 2 50 20 0 30 > $OUT/${TAG}_warm_n20_nx50_cp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "nowarm" 11 4 10 10 0 30 1 heat > $OUT/${TAG}_warm_config5_slice.txt 2>&1

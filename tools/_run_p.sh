OUT=gpurun_out; TAG=r02p
timeout 600 bash tools/ab_variants.sh "f512" 10 2 20 20 0 300 1 heat > $OUT/${TAG}_f512_small_cp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "f512" 18 2 20 20 0 30 1 heat > $OUT/${TAG}_f512_large_cp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "fs2 fs4" 7 2 5 5 0 300 1 heat > $OUT/${TAG}_fs_5_small_cp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "fs2 fs4" 10 2 10 10 0 300 1 heat > $OUT/${TAG}_fs_10_small_cp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "fs2 fs4" 19 2 10 10 0 30 1 heat > $OUT/${TAG}_fs_10_large_cp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "fs2 fs4" 20 2 5 5 0 30 1 heat > $OUT/${TAG}_fs_5_large_cp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "h512" 13 2 50 20 0 100 > $OUT/${TAG}_h512_n13_cp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "h512" 20 2 50 20 0 30 > $OUT/${TAG}_h512_n20_cp.txt 2>&1

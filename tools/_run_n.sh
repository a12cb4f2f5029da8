OUT=gpurun_out; TAG=r02o
timeout 600 bash tools/ab_variants.sh "t256 t256_128 t128_256" 10 2 20 20 0 300 1 heat > $OUT/${TAG}_threads_small_cp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "t256 t256_128 t128_256" 10 2 20 20 0 100 2 heat > $OUT/${TAG}_threads_small_sp.txt 2>&1
timeout 600 bash tools/ab_variants.sh "t256 t256_128 t128_256" 18 2 20 20 0 30 1 heat > $OUT/${TAG}_threads_large_cp.txt 2>&1

OUT=gpurun_out; TAG=r02l
V=$PWD/spock_b200/csrc/_variants
# the prefetching builds must pass the parity suite before their timings mean anything
SPOCK_B200_LIB=$V/pf444.so timeout 600 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_gpu_tests_pf444.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_gpu_tests_pf444.log
SPOCK_B200_LIB=$V/mmpf1.so timeout 600 python -m pytest tests -m gpu -q -x -k "stage or mm or factor" > $OUT/${TAG}_gpu_tests_mmpf1.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_gpu_tests_mmpf1.log
timeout 900 bash tools/ab_variants.sh "pf222 pf444 pf888" 20 2 50 20 0 30 > $OUT/${TAG}_dual3_pf_config3.txt 2>&1
timeout 900 bash tools/ab_variants.sh "pf444 mmpf1 mmpf2" 20 2 50 20 1 30 > $OUT/${TAG}_pf_per_stage.txt 2>&1
timeout 600 bash tools/ab_variants.sh "pf222 pf444 pf888" 11 4 10 10 0 30 1 heat > $OUT/${TAG}_dual3_pf_config5.txt 2>&1

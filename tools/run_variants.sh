set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s6_pytest.log
python tools/quick_time.py 18 2 50 20 1 20 > gpurun_out/s6_qt_minb4.json 2> gpurun_out/s6_err.log
python tools/quick_time.py 11 4 10 10 1 20 1 heat > gpurun_out/s6_qt_d4_minb4.json 2>> gpurun_out/s6_err.log
cp spock_b200/csrc/libspock_b200.so /tmp/lib_main.so
for v in minb3 minb5 minb6; do
  cp spock_b200/csrc/_variants/$v.so spock_b200/csrc/libspock_b200.so
  python tools/quick_time.py 18 2 50 20 1 20 > gpurun_out/s6_qt_$v.json 2>> gpurun_out/s6_err.log
  python tools/quick_time.py 11 4 10 10 1 20 1 heat > gpurun_out/s6_qt_d4_$v.json 2>> gpurun_out/s6_err.log
done
cp /tmp/lib_main.so spock_b200/csrc/libspock_b200.so
cat gpurun_out/s6_pytest.log
for f in gpurun_out/s6_qt_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['ms_per_pass'], d['profile_ms_per_pass'])"; done

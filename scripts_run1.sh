for f in tests/test_gpu_map.py tests/test_gpu_edt.py; do
  echo "=== $f"; timeout -s KILL 500 python -m pytest $f -q -m gpu 2>&1 | tail -25
done

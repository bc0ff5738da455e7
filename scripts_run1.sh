for f in tests/test_gpu_golden.py tests/test_gpu_search.py; do
  echo "=== $f"; timeout -s KILL 600 python -m pytest $f -q -m gpu -x 2>&1 | tail -25
done

timeout 900 python -m pytest tests/test_knots_gpu.py tests/test_i8_gpu.py tests/test_golden_r_gpu.py tests/test_fit_gpu.py -q -x 2>&1 | tail -5
echo "== vi knots"; timeout 120 python tools/run_vi.py 1000000 1024 8 3 vi 1 | tail -2
echo "== vi"; timeout 120 python tools/run_vi.py 1000000 1024 8 3 | tail -1

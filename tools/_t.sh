timeout 900 python -m pytest tests/test_vi_gpu.py tests/test_i8_gpu.py tests/test_knots_gpu.py tests/test_edges_gpu.py tests/test_golden_r_gpu.py tests/test_fit_gpu.py -q 2>&1 | tail -3
for k in 0 1; do echo "== K2=$k"; SRGP_K2=$k timeout 120 python tools/run_vi.py 1000000 1024 8 5 | tail -2; done
echo "== n=125k (8-GPU shard)"; for k in 0 1; do SRGP_K2=$k timeout 120 python tools/run_vi.py 125000 1024 8 5 | tail -1; done

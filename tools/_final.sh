python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/r02_gputests_final.log; cat gpurun_out/r02_gputests_final.log
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; python tools/show_bench.py gpurun_out/r02_bench_final.json | head -3
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; head -c 600 gpurun_out/r02_bench_reference.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-check --no-secondary --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_bench_summary.csv
gzip -f gpurun_out/r02_launches_bench.csv
head -14 gpurun_out/r02_launches_bench_summary.csv

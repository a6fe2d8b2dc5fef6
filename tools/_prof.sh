cap() { # name kernel-regex
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 4 -c 1 -o /tmp/$1 -f python tools/run_vi.py 1000000 1024 8 1 > gpurun_out/r02_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/r02_ncu_$1_details.txt 2>/dev/null
}
cap km2_final i8_km2
cap gram2_final i8_gram2
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-check --no-secondary --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_bench_summary.csv
gzip -f gpurun_out/r02_launches_bench.csv
cat gpurun_out/r02_launches_bench_summary.csv | head -20

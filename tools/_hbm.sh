python tools/bench_k1.py > gpurun_out/r02_k1.json 2>&1
python tools/bench_k5.py > gpurun_out/r02_k5.json 2>&1
cat gpurun_out/r02_k1.json gpurun_out/r02_k5.json
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 9 -c 1 -o /tmp/k1 -f python tools/bench_k1.py 200000 1024 8 > /dev/null 2>&1
ncu -i /tmp/k1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_k1_raw.csv
ncu -i /tmp/k1.ncu-rep --page details > gpurun_out/r02_ncu_k1_details.txt
ncu --set full --clock-control none -k regex:assemble_kernel -s 22 -c 1 -o /tmp/k2 -f python tools/bench_k1.py 200000 1024 8 > /dev/null 2>&1
ncu -i /tmp/k2.ncu-rep --page raw --csv > gpurun_out/r02_ncu_k2lc_raw.csv
ncu --set full --clock-control none -k regex:assemble_kernel -s 35 -c 1 -o /tmp/k3 -f python tools/bench_k1.py 200000 1024 8 > /dev/null 2>&1
ncu -i /tmp/k3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_k2tau_raw.csv
ncu --set full --clock-control none --import-source on -k regex:omega_dk -s 5 -c 1 -o /tmp/k5 -f python tools/bench_k5.py 200000 1024 8 > /dev/null 2>&1
ncu -i /tmp/k5.ncu-rep --page raw --csv > gpurun_out/r02_ncu_k5_raw.csv
ncu -i /tmp/k5.ncu-rep --page details > gpurun_out/r02_ncu_k5_details.txt
ls -la gpurun_out

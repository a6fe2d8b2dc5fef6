cap() { # name kernel-regex skip
  SRGP_NO_OVERLAP=1 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/$1 -f python tools/run_vi.py 200000 1024 8 1 > gpurun_out/r02_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/r02_ncu_$1_source.csv 2>/dev/null
}
cap gen_data gen_slices_datarows 3
cap gen_knot gen_slices_knotrows 3
cap potrf potrf_diag 3

for n in 1 2 4 8; do
  if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_scale_$n.json 2> gpurun_out/r02_scale_$n.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale_$n.json 2> gpurun_out/r02_scale_$n.err; fi
  python tools/show_bench.py gpurun_out/r02_scale_$n.json | head -2
done

for k in node_fwd node_bwd; do
ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -o gpurun_out/r2f_$k python bench.py --steps 1 --warmup 1 --profile-steps 0 --no-cpu-baseline --no-cuda-graph > gpurun_out/r2f_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"lbwd_kernel" -c 2 -o gpurun_out/r2f_lbwd python scripts/bench_rmlp.py > gpurun_out/r2f_lbwd.log 2>&1
ls -la gpurun_out/*.ncu-rep

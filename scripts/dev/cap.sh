set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err || exit 1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_traffic_cfg2.csv python bench.py --steps 1 --warmup 3 --profile-steps 0 --no-cpu-baseline > gpurun_out/r2e_traffic.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lbwd_kernel|rmlp_fwd" -c 3 -o gpurun_out/r2e_lbwd python scripts/bench_rmlp.py > gpurun_out/r2e_lbwd.log 2>&1
ls -la gpurun_out | tail -5

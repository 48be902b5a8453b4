T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
TAG=${TAG:-r2j}
$T bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_8gpu.json 2> gpurun_out/${TAG}_8gpu.err
$T bench.py --gpus 8 --steps 10 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_8gpu_strong.json 2>> gpurun_out/${TAG}_8gpu.err
tail -c 300 gpurun_out/${TAG}_8gpu.err
for f in gpurun_out/${TAG}_bench_8gpu*.json; do python -c "
import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['scaling'], d['config'].get('graphs_per_gpu'))"; done

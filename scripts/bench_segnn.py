"""SEGNN (BASELINE.json configs[2] / SURVEY cfg-3) forward+backward throughput on synthetic 5000-node PBC clouds.
   python scripts/bench_segnn.py [graphs_per_gpu] [lmax_attributes] [dense|blocked] [tensor cores 1|0]  -> one JSON line (edges/s), CUDA events, 3 warm-ups."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from eqnn_jax_b200 import graph_utils as GU, ops
from eqnn_jax_b200.equivariant_graph_utils import get_equivariant_graph
from eqnn_jax_b200.irreps import Irreps
from eqnn_jax_b200.nequip import IrrepsArray
from eqnn_jax_b200.segnn import SEGNN

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
LMAX_ATTR = int(sys.argv[2]) if len(sys.argv) > 2 else 1
LOWERING = sys.argv[3] if len(sys.argv) > 3 else "dense"
TC = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
N, K = bench.N_NODES, bench.K_NN
halos, _ = bench.synth_batch(0, B)
x = torch.from_numpy(halos[..., :3].copy()).cuda()
pbc = GU.get_apply_pbc(std=(bench.STD[:3] / bench.BOX).astype(np.float32))
# train_velocities.py:364-378,454-458: d_hidden 128, 3 x 3, sum, node task -> 1x1o, l_max_hidden 2, |r| messages, PBC
model = SEGNN(d_hidden=128, n_layers=3, message_passing_steps=3, message_passing_agg="sum", task="node", output_irreps="1x1o",
              l_max_hidden=2, residual=True, lowering=LOWERING, tensor_cores=TC)
target = torch.randn(B, N, 3, device="cuda")


def step(params=None):
    s, r, _ = ops.knn_pbc(x, K, pbc.cell, pbc.cell_inv, want_dr=False)
    g = get_equivariant_graph(IrrepsArray(Irreps("1o"), x), x, None, s, r, None, None, None, None, LMAX_ATTR, apply_pbc=pbc)
    if params is None:
        return g
    model.forward_backward(params, g, lambda o: (o - target) * (2.0 / o.numel()))    # MSE gradient (device-side elementwise)
    return g


params = model.init(0, step())
for _ in range(3):
    step(params)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 3
e0.record()
for _ in range(steps):
    step(params)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
ops.TIMER = ops.KernelTimer(); step(params); prof = ops.TIMER.summary(); ops.TIMER = None
if os.environ.get("KERNELS"):          # per-kernel device time of one step (CUPTI through torch.profiler)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as pr:
        step(params)
        torch.cuda.synchronize()
    rows = sorted(((e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count, e.key)
                   for e in pr.key_averages()), reverse=True)
    for t, n, k in rows[:24]:
        print(f"{t / 1e3:9.3f} ms  x{n:4d}  {k[:110]}", file=sys.stderr)
E = B * N * K
print(json.dumps({"metric": "SEGNN fwd+bwd edges/sec", "value": E / (ms * 1e-3), "unit": "edges/s", "ms_per_step": ms,
                  "config": {"workload": "SEGNN l_max_hidden 2, 3 steps x 3 layers, 5000-node kNN-20 PBC clouds (cfg-3)",
                             "graphs_per_gpu": B, "lmax_attributes": LMAX_ATTR, "gemm": f"{'tcgen05 3xTF32' if TC else 'fp32 CUDA-core'}, {LOWERING} TPLG lowering"},
                  "breakdown_ms": {k: t for k, (n, t) in sorted(prof.items())}, "params": params.num_params()}))

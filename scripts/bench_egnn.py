"""EGNN (BASELINE.json configs[3]) forward+backward throughput on synthetic 5000-node PBC clouds.
   python scripts/bench_egnn.py [graphs_per_gpu]   -> one JSON line (edges/s), CUDA-event timing, >=3 warm-ups."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from eqnn_jax_b200 import graph_utils as GU, ops
from eqnn_jax_b200.egnn import EGNN

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N, K = bench.N_NODES, bench.K_NN
halos, y = bench.synth_batch(0, B)
x = torch.from_numpy(halos[..., :3].copy()).cuda()
yt = torch.from_numpy(y).cuda()
pbc = GU.get_apply_pbc(std=(bench.STD[:3] / bench.BOX).astype(np.float32))
# train_cosmology.py:67-81 EGNN_PARAMS + n_radial_basis 32, r_max 0.6, apply_pbc (:403-408)
model = EGNN(message_passing_steps=4, d_hidden=128, n_layers=3, activation="gelu", soft_edges=False, positions_only=True,
             tanh_out=False, decouple_pos_vel_updates=True, message_passing_agg="mean", readout_agg="mean",
             mlp_readout_widths=(4, 2, 2), task="graph", n_outputs=2, apply_pbc=pbc, n_radial_basis=32, r_max=0.6)
g0 = GU.GraphsTuple(nodes=x, edges=None, receivers=None, senders=None, globals=None, n_node=None, n_edge=None)
params = model.init(0, g0)
loss = torch.empty(1, device="cuda"); gp = torch.empty(B, 2, device="cuda")

def step():
    s, r, _ = ops.knn_pbc(x, K, pbc.cell, pbc.cell_inv, want_dr=False)
    g = GU.GraphsTuple(nodes=x, edges=None, receivers=r, senders=s, globals=None, n_node=None, n_edge=None)
    def grad(out):
        ops.mse_loss(out, yt, 1.0, loss, gp)
        return gp
    model.forward_backward(params, g, grad)

for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 5
e0.record()
for _ in range(steps): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
ops.TIMER = ops.KernelTimer(); step(); prof = ops.TIMER.summary(); ops.TIMER = None
E = B * N * K
flops = 4 * E * (2 * (32 * 128 + 4 * 128 * 128 + 128) * 3 - 2 * 32 * 128)
print(json.dumps({"metric": "EGNN fwd+bwd edges/sec", "value": E / (ms * 1e-3), "unit": "edges/s", "ms_per_step": ms,
                  "config": {"workload": "EGNN positions-only, 4 steps, 5000-node kNN-20 PBC clouds (cfg-4)", "graphs_per_gpu": B},
                  "mlp_tflops_algorithmic": flops / (ms * 1e-3) / 1e12,
                  "breakdown_ms": {k: t for k, (n, t) in sorted(prof.items())}, "params": params.num_params()}))

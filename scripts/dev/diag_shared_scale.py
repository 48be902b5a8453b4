"""Diagnostic: gradients of the benchmarked model with one radial-MLP evaluation per unique edge length (share_radial=True;
the zero class sums 5000 x B self-loop cotangents into ONE row, which also sets the fp16 operand scale of the backward
kernels) against one evaluation per edge, at cloud sizes the CPU oracle cannot reach.  Prints, per Flax module, the largest
difference relative to the module's gradient scale, and the ratio of the zero row's cotangent to the largest other row."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from eqnn_jax_b200 import graph_utils as GU, ops
from eqnn_jax_b200.nequip import NequIP
from eqnn_jax_b200.train import wrap_graph

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
h, y = bench.synth_batch(0, B)
hd = torch.from_numpy(h).cuda()
pbc = GU.get_apply_pbc(std=(bench.STD[:3] / bench.BOX).astype(np.float32))
s_, r_, _ = ops.knn_pbc(hd, bench.K_NN, pbc.cell, pbc.cell_inv, want_dr=False)
res = {}
for share in (False, True):
    model = NequIP(**bench.MODEL_KW, share_radial=share)
    g = wrap_graph(model, hd, s_, r_, bench.K_NN, pbc)
    params = model.init(0, g)
    yt = torch.from_numpy(y).cuda()
    out = model.forward_backward(params, g, lambda o: 2.0 * (o - yt) / o.numel())
    torch.cuda.synchronize()
    res[share] = (out.clone(), params.grad.clone(), params)
print("forward bit-identical:", torch.equal(res[False][0], res[True][0]))
ga, gb, params = res[False][1], res[True][1], res[True][2]
mods = {}
for path, shape, off in params.layout:
    n = int(np.prod(shape))
    a, b = ga[off:off + n], gb[off:off + n]
    m = mods.setdefault(path[0], [0.0, 0.0])
    m[0] = max(m[0], float(a.abs().max()))
    m[1] = max(m[1], float((a - b).abs().max()))
worst = 0.0
for k, (scale, err) in mods.items():
    rel = err / max(scale, 1e-30)
    worst = max(worst, rel)
    print(f"{k:32s} scale {scale:10.3e}  max diff {err:10.3e}  rel {rel:8.2e}")
print("worst module-relative difference:", f"{worst:.2e}")

#!/usr/bin/env python
"""Benchmark: NequIP forward+backward edges/sec on synthetic Quijote-shaped halo clouds.

Workload (BASELINE.json configs[1]): 5000-node clouds, kNN k=20 with PBC in a 1000 Mpc/h box,
NequIP l_max=2 (benchmarks/galaxies/train_cosmology.py NEQUIP_PARAMS + n_radial_basis 64,
r_cutoff 0.6), graph-level regression of 2 targets, MSE loss.  One "step" = one full pass of
the hot path over a batch of `--graphs-per-gpu` clouds per GPU: kNN graph build, CSR + edge
geometry, forward, loss, backward, gradient all-reduce (N>1) and the AdamW update.

    python bench.py --gpus N --steps K --warmup W            (torchrun for N>1)
    python bench.py --impl reference ...                     (CPU reference arm, host cores)
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# benchmarks/galaxies/dataset.py:16-17 (standardisation constants of the halo catalogue)
MEAN = np.array([499.91877075908684, 500.0947802559321, 499.964508664328,
                 -0.0512854365234889, -0.01263126442198149, -0.06458034372345466])
STD = np.array([288.71092533309235, 288.7525818573022, 288.70234893905575,
                344.0231468131901, 343.9333673335964, 344.071876710777])
BOX = 1000.0
N_NODES, K_NN = 5000, 20
MODEL_KW = dict(d_hidden=128, l_max=2, sphharm_norm="integral", message_passing_steps=4, n_layers=3,
                message_passing_agg="mean", readout_agg="mean", mlp_readout_widths=(4, 2, 2), task="graph",
                n_outputs=2, n_radial_basis=64, r_cutoff=0.6)
IRREPS_IN = "1o+1o+1x0e"


def synth_batch(seed0: int, n_graphs: int):
    """Quijote-shaped synthetic halos: uniform positions in the box, Gaussian velocities, standardised."""
    halos = np.empty((n_graphs, N_NODES, 7), dtype=np.float32)
    for i in range(n_graphs):
        rng = np.random.default_rng(seed0 + i)
        pos = rng.uniform(0.0, BOX, size=(N_NODES, 3))
        vel = rng.normal(0.0, 344.0, size=(N_NODES, 3))
        halos[i, :, :6] = ((np.concatenate([pos, vel], -1) - MEAN) / STD).astype(np.float32)
        halos[i, :, 6] = 1.0
    y = np.random.default_rng(10_000 + seed0).normal(size=(n_graphs, 2)).astype(np.float32)
    return halos, y


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    which="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, which="fallback")


def ncu_traffic():
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum of `ncu --set full`) per edge of the kernels of one
    message-passing step, from the committed capture profiles/ncu_traffic.json (written by profiles/traffic.py from an
    .ncu-rep of THIS bench command).  The file records the tree it was taken on; bench.py reports it next to the
    number and returns None when the capture is missing -- nothing is hard-coded."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p))


def traffic_of(tr, workload, patterns, edges):
    """Sum over the kernels whose name contains one of `patterns`: bytes per edge (all launches of one message-passing
    step) and the number of launches; None when the capture does not cover the workload."""
    if tr is None or workload not in tr.get("workloads", {}):
        return None
    w = tr["workloads"][workload]
    tot, n = 0.0, 0
    for k, v in w["kernels"].items():
        if any(pt in k for pt in patterns):
            tot += v["dram_bytes_per_edge_step"]
            n += v["launches_per_step"]
    return {"bytes_per_edge_step": tot, "launches_per_step": n, "per_launch": tot * edges / max(n, 1),
            "source": f"profiles/ncu_traffic.json ({w['capture']}, tree {tr.get('tree', '?')})"} if n else None


def algorithmic(n_live, d_live, d_in=7, k=K_NN, n_rb=64, H=128, L=3):
    """per edge and message-passing step."""
    b_fwd = 8 + 4 * n_live + (12 + 4 * d_in + 4 * d_live) / k
    b_bwd = 8 + 8 * n_live + (12 + 8 * d_in + 4 * d_live) / k
    f_fwd = 2 * (n_rb * H + (L - 1) * H * H + H * n_live)
    f_bwd = 2 * f_fwd - 2 * n_rb * H
    return b_fwd, b_bwd, f_fwd, f_bwd


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# CPU reference arm: the oracle restatement on the host cores (JAX is not installable here)
# --------------------------------------------------------------------------------------
def cpu_reference_step(halos, y, tree, threads):
    """One fwd+bwd of the reference algorithm (dense PBC kNN + NequIP + MSE grads) on one cloud batch."""
    import torch
    from oracle import graph_ref as G
    from oracle import nequip_ref as NQ
    torch.set_num_threads(threads)
    cfg = NQ.NequIPConfig(**MODEL_KW)
    t0 = time.perf_counter()
    g = G.build_graph(halos, None, K_NN, apply_pbc=G.get_apply_pbc((STD[:3] / BOX).astype(np.float32)))
    p = NQ.to_torch(tree, torch.float32, requires_grad=True)
    out = NQ.apply_batched(p, cfg, g._replace(nodes=torch.tensor(halos), edges=None), IRREPS_IN)
    loss = ((out - torch.tensor(y)) ** 2).mean()
    loss.backward()
    return time.perf_counter() - t0, float(loss.detach())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import nequip_ref as NQ
    threads = os.cpu_count() or 1
    tree = NQ.init_params(NQ.NequIPConfig(**MODEL_KW), IRREPS_IN, seed=0)
    halos, y = synth_batch(0, 1)
    for _ in range(args.warmup):
        cpu_reference_step(halos, y, tree, threads)
    ts = [cpu_reference_step(halos, y, tree, threads)[0] for _ in range(args.steps)]
    edges = N_NODES * K_NN
    val = edges * len(ts) / sum(ts)
    line = {"impl": "reference", "metric": "NequIP fwd+bwd edges/sec", "value": val, "unit": "edges/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(ts) / len(ts),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "NequIP lmax=2, 5000-node kNN-20 PBC clouds, graph regression (cfg-2)",
                       "graphs_per_step": 1, "edges_per_graph": edges},
            "cpu_baseline": {"value": val, "unit": "edges/s", "cores": threads, "kind": "port",
                             "sample": "1 cloud (100k edges) per step: dense PBC kNN + fwd + bwd, torch-CPU fp32 "
                                       "restatement (JAX/e3nn-jax not installable in this image)"},
            "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# the other model families on the path: SEGNN (cfg-3) and EGNN (cfg-4)
# --------------------------------------------------------------------------------------
SEGNN_KW = dict(d_hidden=128, n_layers=3, message_passing_steps=3, message_passing_agg="sum", task="node",
                output_irreps="1x1o", l_max_hidden=2, residual=True)        # train_velocities.py:364-378,454-458
SEGNN_LMAX_ATTR = 1
EGNN_KW = dict(message_passing_steps=4, d_hidden=128, n_layers=3, activation="gelu", soft_edges=False,
               positions_only=True, tanh_out=False, decouple_pos_vel_updates=True, message_passing_agg="mean",
               readout_agg="mean", mlp_readout_widths=(4, 2, 2), task="graph", n_outputs=2, n_radial_basis=32,
               r_max=0.6)                                                   # train_cosmology.py:67-81,403-408
FAMILIES = {
    # cpu_nodes: the CPU arm's cloud; the SEGNN restatement needs minutes per step at 5000 nodes, so it runs 1000-node
    # clouds (edges/s is per-edge normalised) and the line says so (same_config false)
    "segnn": dict(metric="SEGNN fwd+bwd edges/sec", graphs=8, cpu_nodes=1000,
                  name="SEGNN l_max_hidden 2, attributes l<=1, 3 steps x 3 layers, node task -> 1x1o, 5000-node kNN-20 "
                       "PBC clouds (cfg-3)"),
    "egnn": dict(metric="EGNN fwd+bwd edges/sec", graphs=32, cpu_nodes=5000,
                 name="EGNN positions-only, 4 steps, graph regression, 5000-node kNN-20 PBC clouds (cfg-4)"),
}


def family_targets(kind, halos, y):
    """Synthetic targets: SEGNN regresses per-node velocities (train_velocities.py), EGNN the 2 graph labels."""
    return halos[..., 3:6].copy() if kind == "segnn" else y


def cpu_family_step(kind, pos, target, tree, threads):
    """One fwd+bwd of the oracle restatement (dense PBC kNN + steerable attributes + model + MSE grads) on the host."""
    import torch
    from oracle import graph_ref as G
    torch.set_num_threads(threads)
    pbc = G.get_apply_pbc((STD[:3] / BOX).astype(np.float32))
    t0 = time.perf_counter()
    g = G.build_graph(pos, None, K_NN, apply_pbc=pbc)
    if kind == "segnn":
        from oracle import nequip_ref as NQ, segnn_ref as S
        from oracle.e3nn_ref import IrrepsArray as OIA
        st = G.get_equivariant_graph(OIA("1o", torch.tensor(pos, dtype=torch.float32)), pos, None, g.senders, g.receivers,
                                     g.n_node, g.n_edge, None, None, SEGNN_LMAX_ATTR, apply_pbc=pbc)
        p = NQ.to_torch(tree, torch.float32, requires_grad=True)
        out = S.apply_batched(p, S.SEGNNConfig(**SEGNN_KW), st, dtype=torch.float32).array
    else:
        from oracle import egnn_ref as EG, nequip_ref as NQ
        p = NQ.to_torch(tree, torch.float32, requires_grad=True)
        cfg = EG.EGNNConfig(**{**EGNN_KW, "apply_pbc": pbc})
        out = EG.apply_batched(p, cfg, g._replace(nodes=torch.tensor(pos), edges=None))
    loss = ((out - torch.tensor(target)) ** 2).mean()
    loss.backward()
    return time.perf_counter() - t0, float(loss.detach())


def cpu_family_baseline(kind, n_nodes, reps, warm=1):
    threads = os.cpu_count() or 1
    halos, y = synth_batch(0, 1)
    halos, y = halos[:, :n_nodes], y
    pos, target = halos[..., :3].copy(), family_targets(kind, halos, y)
    if kind == "segnn":
        from oracle import segnn_ref as S
        from oracle.so3 import Irreps as OI
        tree = S.init_params(S.SEGNNConfig(**SEGNN_KW), "1o", OI.spherical_harmonics(SEGNN_LMAX_ATTR), "1x0e", seed=0)
    else:
        from oracle import egnn_ref as EG
        tree = EG.init_params(EG.EGNNConfig(**EGNN_KW), 3, seed=0)
    for _ in range(warm):
        cpu_family_step(kind, pos, target, tree, threads)
    ts = [cpu_family_step(kind, pos, target, tree, threads)[0] for _ in range(reps)]
    edges = n_nodes * K_NN
    return edges * len(ts) / sum(ts), 1e3 * sum(ts) / len(ts), threads, \
        f"1 cloud of {n_nodes} nodes ({edges} edges) per step: dense PBC kNN + attributes + fwd + bwd, torch-CPU fp32 " \
        f"restatement (JAX/e3nn-jax not installable in this image)"


def run_family(args):
    kind = args.workload
    fam = FAMILIES[kind]
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        val, ms, threads, sample = cpu_family_baseline(kind, fam["cpu_nodes"], max(args.steps, 1), max(args.warmup, 1))
        print(json.dumps({"impl": "reference", "metric": fam["metric"], "value": val, "unit": "edges/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": fam["name"], "graphs_per_step": 1},
                          "cpu_baseline": {"value": val, "unit": "edges/s", "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
              flush=True)
        return
    import torch
    import torch.distributed as dist
    from eqnn_jax_b200 import _lib, graph_utils as GU, ops
    from eqnn_jax_b200.train import TrainState, allreduce_mean_, cosine_decay_lr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a GPU (no CPU fallback exists for this path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    B = args.graphs_per_gpu if "--graphs-per-gpu" in sys.argv else fam["graphs"]
    E_step = B * N_NODES * K_NN
    pbc = GU.get_apply_pbc(std=(STD[:3] / BOX).astype(np.float32))
    host = []
    for j in range(2):
        h, y = synth_batch(1000 * rank + 100 * j, B)
        host.append((torch.from_numpy(h[..., :3].copy()).pin_memory(),
                     torch.from_numpy(np.ascontiguousarray(family_targets(kind, h, y))).pin_memory()))
    resident = [(x.to(dev), t.to(dev)) for x, t in host]

    if kind == "segnn":
        from eqnn_jax_b200.equivariant_graph_utils import get_equivariant_graph
        from eqnn_jax_b200.irreps import Irreps
        from eqnn_jax_b200.nequip import IrrepsArray
        from eqnn_jax_b200.segnn import SEGNN
        model = SEGNN(**SEGNN_KW)

        def make_graph(x, s, r):
            return get_equivariant_graph(IrrepsArray(Irreps("1o"), x), x, None, s, r, None, None, None, None,
                                         SEGNN_LMAX_ATTR, apply_pbc=pbc)
    else:
        from eqnn_jax_b200.egnn import EGNN
        model = EGNN(apply_pbc=pbc, **EGNN_KW)

        def make_graph(x, s, r):
            return GU.GraphsTuple(nodes=x, edges=None, receivers=r, senders=s, globals=None, n_node=None, n_edge=None)

    def graph_of(x):
        s, r, _ = ops.knn_pbc(x, K_NN, pbc.cell, pbc.cell_inv, want_dr=False)
        return make_graph(x, s, r)

    params = model.init(0, graph_of(resident[0][0]))      # same seed on every rank = replicated parameters
    state = TrainState(model, params)
    n = params.flat.numel()
    loss_slot = state.bucket[n:]

    def train(x, t):
        """kNN graph + attributes, forward, MSE, backward, gradient all-reduce, AdamW (the step of eqnn_jax_b200.train)."""
        g = graph_of(x)

        def grad_fn(out):
            o2, t2 = out.reshape(-1, out.shape[-1]), t.reshape(-1, t.shape[-1])
            if state._gpred is None or state._gpred.shape != o2.shape:
                state._gpred = torch.empty_like(o2)
            ops.mse_loss(o2, t2, 1.0, loss_slot, state._gpred)
            return state._gpred.view(out.shape)

        model.forward_backward(params, g, grad_fn)
        allreduce_mean_(state.bucket, world)
        state.step += 1
        lr = cosine_decay_lr(state.step - 1, state.learning_rate, state.decay_steps)
        ops.adamw(params.flat, params.grad, state.mu, state.nu, lr, 0.9, 0.999, 1e-8, state.weight_decay, state.step)
        return loss_slot

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    xbuf, tbuf = torch.empty_like(resident[0][0]), torch.empty_like(resident[0][1])
    loss_host = torch.empty((1,), dtype=torch.float32).pin_memory()

    def step_resident(i):
        train(*resident[i % 2])

    def step_e2e(i):
        x, t = host[i % 2]
        xbuf.copy_(x, non_blocking=True)
        tbuf.copy_(t, non_blocking=True)
        loss_host.copy_(train(xbuf, tbuf), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = L.eqnn_launch_count()
    ms, t0, t1 = timed(step_resident, args.steps)
    launches = L.eqnn_launch_count() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    for i in range(2):
        step_e2e(i)
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    ops.TIMER = ops.KernelTimer()
    step_resident(0)
    prof = ops.TIMER.summary()
    ops.TIMER = None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    if kind == "segnn":
        plan = model.plan(Irreps("1o"), SEGNN_LMAX_ATTR, 1)
        Nt, Et = B * N_NODES, E_step
        edge_layers = {id(t) for st in plan.steps for t in st["edge"]}
        # the first edge layer of a step runs its two wide products at the nodes (h @ M_s, h @ M_r: segnn._edge0_fwd)
        at_nodes = {id(st["edge"][0]) for st in plan.steps if model._edge0_ok(plan, st["edge"][0], 1)}
        # ... and so does the last one (no gate, only aggregated: segnn._edgeL_fwd)
        last_at_nodes = {id(st["edge"][-1]) for st in plan.steps if model._edgeL_ok(plan, st)}
        flops = 0.0
        for t in plan.tplgs:
            if not t.live:
                continue
            rows = Et if id(t) in edge_layers else Nt
            passes = 2 if t is plan.embed else 3                 # forward, weight gradient (+ data gradient)
            if id(t) in at_nodes or id(t) in last_at_nodes:
                # GEMM work as executed: node rows (first layer: the h_s and h_r blocks; last layer: Dy GEMMs on S_j)
                flops += 3 * 2.0 * Nt * (t.Dx - (1 if id(t) in at_nodes else 0)) * t.Dy * t.Do
            else:
                flops += passes * 2.0 * rows * t.Dx * t.Dy * t.Do
        n_g, ms_g = prof.get("segnn_gemm", (0, 0.0))
        tf = flops / (ms_g * 1e-3) / 1e12 if ms_g > 0 else float("nan")
        roof = {"kernel": "TensorProductLinearGate GEMMs fwd + dgrad + wgrad (tc_gemm_kernel: tcgen05 kind::tf32, 3 MMAs "
                          "per product, any strides)", "bound": "tensor", "achieved": tf, "peak": pk["tf_sust"],
                "unit": "TFLOP/s", "frac": tf / pk["tf_sust"], "traffic": None,
                "peak_source": f"bf16_tflops_sustained of {pk['which']}", "launches_per_step": n_g, "ms_per_step": ms_g,
                "flops_per_step": flops,
                "note": "flops = 2 Dx Dy Do per row and pass of the dense-lowered layer as executed (the first and the "
                        "last edge layer of every step at the node rows), 1 per product; the kernel issues 3 tf32 MMAs per product at half "
                        "the bf16 rate, so its own ceiling is peak / 6"}
    else:
        per_edge_step = (128 + 512) + 4 * 1024 + 516 + (1028 + 4 * 1536 + 640) + (640 + 4 * 1024 + 516)
        n_f, ms_f = prof.get("egnn_mlp_fwd", (0, 0.0))
        n_b, ms_b = prof.get("egnn_mlp_bwd", (0, 0.0))
        gbs = E_step * EGNN_KW["message_passing_steps"] * per_edge_step / ((ms_f + ms_b) * 1e-3) / 1e9 \
            if ms_f + ms_b > 0 else float("nan")
        roof = {"kernel": "edge / coordinate MLP GEMMs fwd + dgrad + wgrad (tc_rowgemm_kernel / tc_wgrad_kernel)",
                "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                "traffic": None, "peak_source": f"hbm_gbs of {pk['which']}", "bytes_per_edge_step": per_edge_step,
                "launches_per_step": n_f + n_b, "ms_per_step": ms_f + ms_b,
                "note": "algorithmic traffic of the 32 -> 128 -> (4x) 128 -> 1 per-edge chain: fwd 640 | 4x1024 | 516; "
                        "dgrad 1028 | 4x1536 | 640; wgrad 640 | 4x1024 | 516 bytes per edge and message-passing step"}
    line = {"metric": fam["metric"], "value": world * E_step * args.steps / (ms * 1e-3), "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": fam["name"], "graphs_per_gpu": B, "edges_per_gpu_step": E_step,
                       "step": "kNN graph build + attributes + forward + MSE + backward + grad all-reduce + AdamW",
                       "l2": "inputs larger than L2 (per-step working set of several GB)", "parallelism": f"dp{world}"},
            "e2e": {"value": world * E_step * args.steps / (ms_e2e * 1e-3), "unit": "edges/s",
                    "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(xbuf.numel() * 4 + tbuf.numel() * 4), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "breakdown_ms_per_step": {k: t for k, (c, t) in sorted(prof.items())}}
    if world == 1 and not args.no_cpu_baseline:
        val, _, threads, sample = cpu_family_baseline(kind, fam["cpu_nodes"], 2)
        line["cpu_baseline"] = {"value": val, "unit": "edges/s", "cores": threads, "kind": "port", "sample": sample,
                                "same_config": fam["cpu_nodes"] == N_NODES}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--graphs-per-gpu", type=int, default=32)   # train_cosmology.py:648 batch_size
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=2)     # 0: skip the instrumented pass (ncu captures)
    # cfg2 = the workload BASELINE.json quotes the metric on (default); cfg5 = its large-graph scaling case
    # (50 000-node clouds, k = 32, l_max = 3), 2 clouds per GPU unless --graphs-per-gpu is given
    # segnn / egnn: the other model families of the path (BASELINE.json configs[2], configs[3]) through the same step
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg5", "segnn", "egnn"])
    # weak: --graphs-per-gpu clouds on every GPU (the driver's scaling runs); strong: the reference's regime, a GLOBAL
    # batch of --graphs-per-gpu clouds split over the GPUs (train_cosmology.py:283-298: batch 32 over the devices)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    # the rank-local part of the step (graph build, forward, loss, backward) is captured once in a CUDA graph and
    # replayed -- what jax.jit gives the reference; --no-cuda-graph launches the ~130 kernels one by one
    ap.add_argument("--no-cuda-graph", action="store_true")
    # default: the radial MLP is evaluated once per UNIQUE edge length (the two directions of a reciprocated edge have
    # bit-identical lengths: NequIP.share_radial, csrc/pairs.cu); --no-share-radial evaluates it once per edge
    ap.add_argument("--no-share-radial", action="store_true")
    args = ap.parse_args()
    if args.workload in FAMILIES:
        return run_family(args)
    global N_NODES, K_NN
    workload_name = "NequIP lmax=2, 5000-node kNN-20 PBC clouds, graph regression (cfg-2)"
    if args.workload == "cfg5":
        N_NODES, K_NN = 50_000, 32
        MODEL_KW["l_max"] = 3
        workload_name = "NequIP lmax=3, 50000-node kNN-32 PBC clouds, graph regression (cfg-5)"
        if "--graphs-per-gpu" not in sys.argv:
            args.graphs_per_gpu = 2
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from eqnn_jax_b200 import _lib, graph_utils as GU, ops
    from eqnn_jax_b200.nequip import NequIP
    from eqnn_jax_b200.train import GraphedTrainStep, TrainState, train_step, wrap_nodes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a GPU (no CPU fallback exists for this path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    B = args.graphs_per_gpu
    if args.scaling == "strong":
        if B % world:
            raise SystemExit(f"--scaling strong: a global batch of {B} clouds does not split over {world} GPUs")
        B //= world
    E_step = B * N_NODES * K_NN

    # inputs: two different host batches per rank (pinned), cycled
    host = []
    for j in range(2):
        h, y = synth_batch(1000 * rank + 100 * j, B)
        host.append((torch.from_numpy(h).pin_memory(), torch.from_numpy(y).pin_memory()))
    resident = [(h.to(dev), y.to(dev)) for h, y in host]
    apply_pbc = GU.get_apply_pbc(std=(STD[:3] / BOX).astype(np.float32))   # train_cosmology.py:401

    model = NequIP(**MODEL_KW, share_radial=not args.no_share_radial)
    g0 = GU.GraphsTuple(nodes=wrap_nodes(resident[0][0]), edges=None, receivers=None, senders=None, globals=None,
                        n_node=None, n_edge=None)
    params = model.init(0, g0)                       # same seed on every rank = replicated parameters
    state = TrainState(model, params)
    plan = model.plan(IRREPS_IN)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    graphed, graph_note = None, "off (--no-cuda-graph)"
    if not args.no_cuda_graph:
        try:
            graphed = GraphedTrainStep(state, resident[0][0].shape, resident[0][1].shape, K_NN, apply_pbc, world)
            graphed.capture(*resident[0])
            graph_note = "rank-local part of the step replayed from one CUDA graph"
        except Exception as e:      # keep the measurement: fall back to launching kernel by kernel
            graphed, graph_note = None, f"capture failed ({type(e).__name__}: {e}); eager launches"
            torch.cuda.synchronize()

    def run_step(h, y):
        return graphed(h, y) if graphed is not None else train_step(state, h, y, K_NN, apply_pbc, world)

    def step_resident(i):
        h, y = resident[i % 2]
        run_step(h, y)

    hbuf = torch.empty_like(resident[0][0])
    ybuf = torch.empty_like(resident[0][1])
    loss_host = torch.empty((1,), dtype=torch.float32).pin_memory()

    def step_e2e(i):
        h, y = host[i % 2]
        if graphed is not None:                         # pinned host memory straight into the graph's input buffers
            loss = graphed(h, y)
        else:
            hbuf.copy_(h, non_blocking=True)
            ybuf.copy_(y, non_blocking=True)
            loss = train_step(state, hbuf, ybuf, K_NN, apply_pbc, world)
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the reference reads the loss every step (:468)

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = L.eqnn_launch_count()
    ms, t0, t1 = timed(step_resident, args.steps)
    launches = L.eqnn_launch_count() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    for i in range(2):
        step_e2e(i)
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    if graphed is not None:
        # a replayed graph does not pass through the library's launch counter: count one eager step of the same code
        # (the kernels the graph holds) and add the per-step AdamW launch that stays outside the graph
        l1 = L.eqnn_launch_count()
        train_step(state, *resident[0], K_NN, apply_pbc, world)
        launches = (L.eqnn_launch_count() - l1) * args.steps

    # instrumented pass: per-kernel CUDA-event durations on the launching stream (eager launches)
    ops.TIMER = ops.KernelTimer()
    for i in range(args.profile_steps):
        train_step(state, *resident[i % 2], K_NN, apply_pbc, world)
    prof = ops.TIMER.summary()
    ops.TIMER = None
    if args.profile_steps == 0:          # capture runs: the plain line only
        if rank == 0:
            print(json.dumps({"metric": "NequIP fwd+bwd edges/sec", "value": world * E_step * args.steps / (ms * 1e-3),
                              "unit": "edges/s", "ms_per_step": ms / args.steps, "gpu_launches": int(launches),
                              "n_gpus": world, "scaling": args.scaling, "graphs_per_gpu": B, "cuda_graph": graph_note,
                              "edges_per_gpu_step": E_step, "note": "capture run (--profile-steps 0)"}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    n_live, d_live = plan.tp.n_live, plan.tp.d_live
    b_fwd, b_bwd, f_fwd, f_bwd = algorithmic(n_live, d_live, k=K_NN)
    steps_mp = MODEL_KW["message_passing_steps"]

    def per_launch(tag):
        n, t = prof.get(tag, (0, 0.0))
        return n, (t / n if n else float("nan"))

    n_f, ms_mf = per_launch("radial_mlp_fwd")
    n_b, ms_mb = per_launch("radial_mlp_bwd")
    n_c, ms_mc = per_launch("pair_combine")           # shared radial evaluations: summing the pairs' cotangents
    ms_mc = 0.0 if ms_mc != ms_mc else ms_mc
    mlp_ms_step = (n_f * ms_mf + n_b * ms_mb + n_c * ms_mc) / args.profile_steps
    # radial-MLP evaluations per edge: 1, or (unique edge lengths) / edges with share_radial -- measured on this batch
    from eqnn_jax_b200.train import wrap_graph
    s_, r_, _ = ops.knn_pbc(resident[0][0], K_NN, apply_pbc.cell, apply_pbc.cell_inv, want_dr=False)
    pg = model.prepare(wrap_graph(model, resident[0][0], s_, r_, K_NN, apply_pbc))
    evals_per_edge = float(pg.pairs.n_unique.item()) / E_step if pg.pairs is not None else 1.0
    del pg
    mlp_flops_step = E_step * evals_per_edge * steps_mp * (f_fwd + f_bwd)     # flops of the evaluations that are executed
    mlp_tf = mlp_flops_step / (mlp_ms_step * 1e-3) / 1e12 if mlp_ms_step > 0 else float("nan")
    fused = model.fuse_radial_mlp and ops.radial_mlp_supported(MODEL_KW["n_radial_basis"], MODEL_KW["d_hidden"],
                                                               MODEL_KW["n_layers"], n_live)
    stash = fused and model.radial_mlp_backward == "stash"
    H = MODEL_KW["n_layers"]
    if fused:
        # bytes the present design moves per edge and message-passing step:
        #   forward  d (4) + w_t (4 n) [+ z stash 512 H]
        #   backward dw (4 n) + z stash in / g out (2 x 512 H) | recompute: activations + gradients out (256 + 1024 H)
        #   weight gradients: basis 256 + z (or activations) 512 H + g 512 H + dw 4 n
        mlp_bytes = (4 + 4 * n_live + (512 * H if stash else 0)) + (4 * n_live + (1024 * H if stash else 256 + 1024 * H)) + \
                    (256 + 1024 * H + 4 * n_live)
        if stash and H >= 2 and not os.environ.get("EQNN_RMLP_NO_LBWD"):
            # fused per-layer backward (rmlp_lbwd_kernel): max|dw| pass 4 n | layer H: z_H + z_{H-1} + dw in, g_{H-1} out |
            # layers H-1..2: z_{l-1} + g_l in, g_{l-1} out | first layer's weight gradient: basis + g_1 | output layer's: z_H + dw
            # (the forward stash holds z_1 .. z_{H-1}, gelu(z_H), gelu'(z_H): H + 1 matrices)
            mlp_bytes = (4 + 4 * n_live + 512 * (H + 1)) + 4 * n_live + (1024 + 4 * n_live + 512) + (H - 2) * 1536 + \
                        (256 + 512) + (512 + 4 * n_live)
    else:
        mlp_bytes = (256 + 512) + 2 * 1024 + (512 + 4 * n_live) + (4 * n_live + 1024) + 2 * 1536 + \
                    (512 + 4 * n_live) + 2 * 1024 + (256 + 512)
    # per EDGE: the MLP's bytes per evaluation x evaluations per edge, plus (shared evaluations) the pair sum of the
    # cotangents: per-slot records in, per-evaluation columns + the two slot indices out / in
    mlp_bytes_edge = mlp_bytes * evals_per_edge + ((4 * n_live) + (4 * n_live + 8) * evals_per_edge if evals_per_edge < 1.0 else 0)
    mlp_gbs = E_step * steps_mp * mlp_bytes_edge / (mlp_ms_step * 1e-3) / 1e9 if mlp_ms_step > 0 else float("nan")
    tr = ncu_traffic()
    tr_mlp = traffic_of(tr, args.workload, ("rmlp_", "tc_wgrad", "tc_rowgemm", "absmax", "wgrad_reduce", "lbwd_reduce", "pair_combine", "zero_sum"), E_step)
    tr_tp = traffic_of(tr, args.workload, ("tpconv_fwd", "tpconv_bwd", "sender_reduce", "segment_gather"), E_step)
    _, ms_tf = per_launch("tpconv_fwd")
    _, ms_tb = per_launch("tpconv_bwd")
    _, ms_ts = per_launch("sender_reduce")            # the sender-side sum of dL/dxt belongs to the backward pass
    ms_ts = 0.0 if ms_ts != ms_ts else ms_ts
    tp_gbs = E_step * (b_fwd + b_bwd) / ((ms_tf + ms_tb + ms_ts) * 1e-3) / 1e9
    _, ms_knn = per_launch("knn_pbc")
    total_prof = sum(t for _, t in prof.values()) / args.profile_steps

    value = world * E_step * args.steps / (ms * 1e-3)
    e2e = world * E_step * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": "NequIP fwd+bwd edges/sec", "value": value, "unit": "edges/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name,
                   "graphs_per_gpu": B, "edges_per_gpu_step": E_step, "message_passing_steps": steps_mp,
                   "cuda_graph": graph_note,
                   "radial_mlp": (f"one evaluation per unique edge length (the two directions of a reciprocated edge have "
                                  f"bit-identical lengths and share it; cotangents summed per pair): {evals_per_edge:.3f} "
                                  f"evaluations per edge on this batch") if evals_per_edge < 1.0 else "one evaluation per edge",
                   "step": "kNN graph build + CSR/geometry" + (" + edge pairing" if evals_per_edge < 1.0 else "") +
                           " + forward + MSE + backward + grad all-reduce + AdamW",
                   "l2": "inputs larger than L2 (per-step working set of several GB)",
                   "parallelism": f"dp{world}"},
        "e2e": {"value": e2e, "unit": "edges/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(hbuf.numel() * 4 + ybuf.numel() * 4), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        # dominant kernels = the radial MLP (tensor-core bound dense contraction).  `achieved` counts
        # the ALGORITHMIC flops (one per product: 2 (n_rb H + (L-1) H^2 + H n_live) forward, twice that minus the first
        # layer's data gradient backward); the kernels issue three fp16 MMAs per product (hi*hi + lo*hi + hi*lo, fp32-class
        # accuracy), so their own ceiling is peak / 3.  The HBM view of the same launches is kept beside it.
        "roofline": {"kernel": ("fused radial MLP: rmlp_fwd_kernel (Bessel + all layers, activations in tensor memory) + "
                                "rmlp_lbwd_kernel per hidden layer (data gradient + weight gradient in one pass) + tc_wgrad_kernel x2; "
                                "tcgen05 kind::f16 / kind::tf32, 3 MMAs per product") if fused else
                               "radial MLP GEMMs fwd+dgrad+wgrad (tc_rowgemm_kernel / tc_wgrad_kernel, tcgen05 kind::tf32 x3)",
                     "bound": "tensor", "achieved": mlp_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                     "frac": mlp_tf / pk["tf_sust"],
                     "traffic": tr_mlp["per_launch"] if tr_mlp else None,
                     "traffic_source": tr_mlp["source"] if tr_mlp else "no ncu capture of this workload committed",
                     "traffic_bytes_per_edge_step": tr_mlp["bytes_per_edge_step"] if tr_mlp else None,
                     "peak_source": f"bf16_tflops_sustained of {pk['which']}",
                     "flops_per_edge_step": f_fwd + f_bwd, "executed_mma_per_product": 3,
                     "flops_per_evaluation": f_fwd + f_bwd, "evaluations_per_edge": evals_per_edge,
                     "launches_per_step": (n_f + n_b) / args.profile_steps, "ms_per_step": mlp_ms_step,
                     "ms_fwd_per_launch": ms_mf, "ms_bwd_per_call": ms_mb,
                     "note": "flops per SURVEY.md 8(d).2 with the live output columns, counted for the radial-MLP evaluations "
                             "that are EXECUTED (evaluations_per_edge x edges); time = CUDA events around every "
                             "eqnn_radial_mlp_fwd / _bwd (+ eqnn_pair_combine) call of the instrumented pass"},
        "roofline_mlp_hbm": {"kernel": "same launches, HBM view (bytes the design moves: forward d + w_t + z stash; backward per hidden "
                                       "layer z_{l-1} + g_l in, g_{l-1} out; first / output layer weight gradients basis + g_1, z_H + dw)",
                             "bound": "hbm", "achieved": mlp_gbs, "peak": pk["hbm"], "unit": "GB/s",
                             "frac": mlp_gbs / pk["hbm"], "bytes_per_edge_step": mlp_bytes_edge,
                             "bytes_per_evaluation": mlp_bytes,
                             "traffic": tr_mlp["per_launch"] if tr_mlp else None,
                             "peak_source": f"hbm_gbs of {pk['which']}"},
        "roofline_tpconv": {"kernel": "tpconv_fwd + tpconv_bwd + sender-side reduction", "bound": "hbm", "achieved": tp_gbs,
                            "peak": pk["hbm"], "unit": "GB/s", "frac": tp_gbs / pk["hbm"],
                            "edge_steps_per_s": E_step / ((ms_tf + ms_tb + ms_ts) * 1e-3), "ms_sender_reduce": ms_ts,
                            "traffic": tr_tp["per_launch"] if tr_tp else None,
                            "traffic_source": tr_tp["source"] if tr_tp else "no ncu capture of this workload committed",
                            "traffic_bytes_per_edge_step": tr_tp["bytes_per_edge_step"] if tr_tp else None,
                            "bytes_per_edge_step": b_fwd + b_bwd, "ms_fwd": ms_tf, "ms_bwd": ms_tb,
                            "n_live": n_live, "d_live": d_live, "peak_source": f"hbm_gbs of {pk['which']}"},
        "breakdown_ms_per_step": {k: t / args.profile_steps for k, (n, t) in sorted(prof.items())},
        "knn": {"ms_per_batch": ms_knn, "pairs_per_s": B * N_NODES * N_NODES / (ms_knn * 1e-3),
                "note": "cell-list search (bit-identical to the dense sweep); pairs_per_s is the dense-equivalent rate N^2/t"},
        "profiled_kernel_ms_per_step": total_prof,
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import nequip_ref as NQ
        threads = os.cpu_count() or 1
        tree = NQ.init_params(NQ.NequIPConfig(**MODEL_KW), IRREPS_IN, seed=0)
        h1, y1 = synth_batch(0, 1)
        cpu_reference_step(h1, y1, tree, threads)               # warm-up
        ts = [cpu_reference_step(h1, y1, tree, threads)[0] for _ in range(2)]
        line["cpu_baseline"] = {"value": N_NODES * K_NN * len(ts) / sum(ts), "unit": "edges/s", "cores": threads,
                                "kind": "port", "sample": "1 cloud (100k edges), dense PBC kNN + fwd + bwd, "
                                "torch-CPU fp32 restatement of the reference (JAX not installable here)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

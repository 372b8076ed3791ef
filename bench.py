#!/usr/bin/env python
"""bench.py -- headline benchmark of the MM-PDE hot path on B200 (contract in the task statement).

Workload (BASELINE.json configs[1]): Burgers 2-D MM-PDE, base_resolution 31x48x48, batch 16 per GPU
(N = 36 864 nodes, E = 1 290 240 edges per graph), moved mesh + learned interpolation both ways + two
6-layer MP-PDE processors, k = 35.  One *step* = one body of training_loop_branch
(/root/reference/train_helper_2d.py:95-131): create_data, moved-mesh graph, uniform graph, both solvers,
interpolation, MSE, backward, AdamW.  metric = edge-updates/s (fwd+bwd): 2 solvers x E x 6 layers per step.

  value : inputs (the trajectory batch) resident in HBM, CUDA events, max over ranks.
  e2e   : the same step through the public API with HOST (pinned) buffers; H2D of the step's inputs and a
          D2H read of the loss inside the timed region.
  roofline : the dominant kernel (mmpde_edge_bwd) timed with CUDA events on its launch stream, every launch
          of the timed region; algorithmic FLOPs (reference formulation) / duration.
  kernels  : the same for every kernel kind of the step (launches per step, average duration, bound, fraction).
  cpu_baseline : the oracle port of the reference path on the host cores, the full batch-16 step, rank 0, N=1.
  parity_probe (N > 1): sharded step vs the global batch recomputed in one process (loss, gradients).
  cylinder / c4 : BASELINE.json configs[2] / configs[3] as sub-records of the same line (--no-extras skips them).
  --impl reference : only the CPU oracle (the reference needs PyG, not installable), same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

RES = [31, 48, 48]
CY_RES = [30, 2521]
BATCH = 16
K_NEIGH = 35
LAYERS = 6
FLOP_PER_EDGE_FWD = 99328          # 2*260*128 + 2*128*128 (reference formulation)
# The mesh mover of the step: a seeded, default-initialised DMM with the reference's constructor arguments
# (/root/reference/mmpde.py:199 for Burgers, /root/reference/README.md:31 for the cylinder); trained checkpoints live on
# Google Drive.  Default initialisation moves the nodes by about one cell.
MOVER_SEED = 4321
DMM_ARRAY = dict(branch_layer=7, trunk_layer=[2, 32, 512], out_layer=[1024, 512, 1])
DMM_GRAPH = dict(branch_layer=[4, 3], trunk_layer=[2, 16, 512], out_layer=[1024, 512, 1])


RESULT_OUT = sys.stdout


def _ncu_capture(kernel="edge_bwd"):
    """The committed `ncu --set full` capture of a kernel of THIS build (profiles/r02_ncu_<kernel>.json, written by
    profiles/ncu_summary.py --json on the CPU box from the .ncu-rep a gpurun call brought back): DRAM bytes per launch,
    the duration ncu saw and the source hash of the kernel file it was taken on.  None when absent."""
    path = os.path.join(ROOT, "profiles", f"r02_ncu_{kernel}.json")
    if os.path.exists(path):
        return json.load(open(path))
    return None


def _ncu_traffic():
    d = _ncu_capture()
    return None if d is None else d.get("dram_bytes_read", 0) + d.get("dram_bytes_write", 0)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm": p["hbm_gbs"], "bf16": p["bf16_tflops_sustained"], "source": "MEASURED_PEAKS.json (sustained)"}
    return {"hbm": 6650.0, "bf16": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock and the throttle reasons seen between t_begin and t_end (perf_counter stamps of the timed
        region; nvidia-smi is started well before it because its first sample can take a second on an 8-GPU box)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t_begin is None or (t_begin - 0.05 <= t <= t_end + 0.15)]
        window = "timed region"
        if not rows:                                   # region shorter than the sampling period: nearest samples under load
            rows, window = [r for _, r in self.rows[-5:]], "last samples before the end of the timed region"
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------
def _oracle_setup(batch, seed=0):
    from oracle import creator, itp, loops, pdes, processor
    from mmpde_b200 import synthetic
    torch.manual_seed(seed)
    pde = pdes.burgers()
    pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = RES
    gc = creator.GraphCreator_FS_2D(pde, K_NEIGH, "knn", 1, RES[0], knn_backend="sklearn")
    model, model_b = processor.MP_PDE_Solver_2D(pde), processor.MP_PDE_Solver_2D(pde)
    net = itp.ItpNet(RES[1], RES[2], [128, 64], [128, 64], [1, 4, 16, 4, 1])
    opt = torch.optim.AdamW([{"params": model.parameters()}, {"params": model_b.parameters()},
                             {"params": net.parameters()}], lr=2e-3)
    fields = synthetic.burgers_fields(batch, RES[0], RES[1], RES[2], seed=seed)
    from oracle import dmm as odmm
    torch.manual_seed(MOVER_SEED)                 # the same seeded default-initialised DMM as the CUDA arm (mmpde.py:199)
    mover = odmm.DMM(s=RES[1], mode="array", **DMM_ARRAY).eval()
    return gc, model, model_b, net, opt, fields, mover, loops


def cpu_reference_step_time(sample_batch, steps, warmup, budget_s=1200.0):
    """Times the oracle's training_loop_branch body on the host cores for a `sample_batch`-trajectory batch.
    Returns (edge-updates/s, mean seconds per timed step, threads, steps actually timed): the loop stops early when the
    timed steps have used `budget_s` (the driver's limit for the whole arm is 1800 s)."""
    torch.set_num_threads(os.cpu_count() or 1)
    gc, model, model_b, net, opt, fields, mover, loops = _oracle_setup(sample_batch)
    model.train(); model_b.train(); net.train()
    loader = [(fields, fields)]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        loops.training_loop_branch(model, model_b, net, mover, [0], sample_batch, opt, None, loader, gc, loops.criterion)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
            if sum(times) > budget_s:
                break
    n = RES[1] * RES[2]
    edge_updates = 2 * sample_batch * n * K_NEIGH * LAYERS
    mean = sum(times) / len(times)
    return edge_updates / mean, mean, torch.get_num_threads(), len(times)


WORKLOAD_BURGERS = ("Burgers 2D MM-PDE training step (DMM-moved mesh [default-initialised DMM, array mode] + interpolation "
                    "+ 2x 6-layer processor), 31x48x48, k=35")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # The SAME configuration as our arm: the full batch-16 step (N = 36 864, E = 1 290 240, both solvers, interpolation,
    # backward, AdamW), ~11 s per step on the box's 16 cores.  One warm-up step at most (allocator, thread pools).
    # (MMPDE_BENCH_REF_BATCH shrinks the sample for the CPU-only contract test; the driver never sets it.)
    sample = int(os.environ.get("MMPDE_BENCH_REF_BATCH", BATCH))
    value, sec, cores, timed = cpu_reference_step_time(sample, max(1, args.steps), min(max(args.warmup, 0), 1))
    line = {
        "impl": "reference", "metric": "edge-updates/sec (fwd+bwd)", "value": value, "unit": "edge-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_BURGERS, "per_gpu_batch": BATCH, "timed_sample_batch": sample,
                   "steps_timed": timed, "warmup_run": min(max(args.warmup, 0), 1)},
        "cpu_baseline": {"value": value, "unit": "edge-updates/s", "cores": cores, "kind": "port",
                         "sample": (f"the full batch-{BATCH} step" if sample == BATCH else f"batch {sample} of the batch-{BATCH} step")
                                   + f", {timed} timed steps (mean {sec:.1f} s); "
                                   "plain-torch oracle port of the reference path incl. sklearn kd-tree kNN; "
                                   "the reference itself needs torch_geometric/torch_cluster, not installable offline"},
        "e2e": {"value": value, "unit": "edge-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
# per-kernel accounting: ALGORITHMIC work of one C-ABI call from its arguments (include/mmpde_b200.h), SURVEY.md 8(d)
H = 128


def _kernel_work(name, a, ctx):
    """-> (bound, work) with work in FLOP (tensor) or bytes (hbm); None when the kernel has no roofline entry."""
    if name == "mmpde_edge_fwd":
        return "tensor", FLOP_PER_EDGE_FWD * a[4]
    if name == "mmpde_edge_bwd":
        return "tensor", 2 * FLOP_PER_EDGE_FWD * a[4]
    if name == "mmpde_node_gemm":                       # 2*M*128*K, K = 128 per segment (+4 extension columns)
        k = H * (2 if a[2] else 1) + (4 if a[10] else 0)
        return "tensor", 2 * a[20] * H * k
    if name == "mmpde_node_wgrad_grouped":
        return "tensor", ctx.get("wgrad_flops", 0)
    if name in ("mmpde_bn_stats", "mmpde_bn_stats_fused"):   # one pass over [M,128] (+ the residual operand)
        return "hbm", a[4] * H * 4 * (2 if a[2] else 1)
    if name == "mmpde_bn_apply":
        return "hbm", a[4] * H * 4 * ((2 if a[2] else 1) + 1)
    if name in ("mmpde_bn_bwd_reduce", "mmpde_bn_bwd_reduce_fused"):
        return "hbm", a[9] * H * 4 * (2 + (1 if a[7] else 0))
    if name == "mmpde_bn_bwd_apply":
        return "hbm", a[9] * H * 4 * (2 + (1 if a[7] else 0) + 1 + (1 if a[17] else 0))
    if name == "mmpde_decoder_fwd":
        return "hbm", a[2] * (H * 4 + 4)
    if name == "mmpde_decoder_bwd":
        return "hbm", a[2] * (2 * H * 4 + 4)
    if name == "mmpde_itp_fwd":                         # ~144 B / query
        return "hbm", a[4] * 144
    if name == "mmpde_itp_bwd":
        return "hbm", a[4] * (144 + 30 * 4)
    if name == "mmpde_itp_fwd_tc":                      # tcgen05: 2*(62*128 + 128*64 + 64*30) = 36 096 FLOP / query (x3 executed)
        return "tensor", a[4] * 36096
    if name == "mmpde_itp_bwd_tc":                      # forward again + the two data-gradient contractions (weight gradients
        return "tensor", a[4] * (36096 + 2 * (30 * 64 + 64 * 128))      # go out as a grouped node_wgrad launch)
    return None, 0


def profile_kernels(train_step_eager, steps, peaks, barrier):
    """Every C-ABI launch of `steps` eager steps bracketed by CUDA events on its launch stream -> per kernel kind:
    launches per step, mean duration, bound, achieved / peak.  (Nodes of a replayed CUDA graph cannot carry events.)"""
    import ctypes
    from mmpde_b200 import _cabi
    import mmpde_b200.ops as ops_mod
    real_call = _cabi.call
    rec = []

    def profiled_call(name, *a):
        ctx = {}
        if name == "mmpde_node_wgrad_grouped":
            tasks = ctypes.cast(a[0], ctypes.POINTER(_cabi.WgradTask))
            ctx["wgrad_flops"] = sum(2 * tasks[i].M * H * ((H if tasks[i].B else 0) + (5 if (tasks[i].Bext or tasks[i].dbias) else 0))
                                     for i in range(a[1]))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = real_call(name, *a)
        e.record()
        rec.append((name, s, e) + _kernel_work(name, a, ctx))
        return rc

    ops_mod._cabi.call = profiled_call
    os.environ["MMPDE_OVERLAP_SOLVERS"] = "0"          # one stream: every kernel timed alone, not beside the other solver
    try:
        barrier()
        for _ in range(steps):
            train_step_eager()
        barrier()
    finally:
        os.environ.pop("MMPDE_OVERLAP_SOLVERS")
        ops_mod._cabi.call = real_call
    by = {}
    for name, s, e, bound, work in rec:
        d = by.setdefault(name, {"n": 0, "ms": 0.0, "work": 0.0, "bound": bound, "all": []})
        t = s.elapsed_time(e)
        d["n"] += 1
        d["ms"] += t
        d["all"].append(t)
        d["work"] += work
    table = []
    for name, d in sorted(by.items(), key=lambda kv: -kv[1]["ms"]):
        srt = sorted(d["all"])
        row = {"name": name, "launches_per_step": d["n"] / steps, "avg_us": 1e3 * d["ms"] / d["n"],
               "min_us": 1e3 * srt[0], "median_us": 1e3 * srt[len(srt) // 2], "max_us": 1e3 * srt[-1],
               "us_per_step": 1e3 * d["ms"] / steps, "bound": d["bound"]}
        if d["bound"] == "tensor" and d["ms"] > 0:
            row["achieved_tflops"] = d["work"] / (d["ms"] * 1e-3) / 1e12
            row["frac"] = row["achieved_tflops"] / peaks["bf16"]
        elif d["bound"] == "hbm" and d["ms"] > 0:
            row["achieved_gbs"] = d["work"] / (d["ms"] * 1e-3) / 1e9
            row["frac"] = row["achieved_gbs"] / peaks["hbm"]
        table.append(row)
    return table, by


def kineto_table(step, steps, path, barrier):
    """Kernel durations (CUPTI, through torch.profiler) summed per kernel name over `steps` replayed steps -> text table."""
    import collections
    from torch.profiler import ProfilerActivity, profile
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    barrier()
    if path is None:
        return
    wall = e0.elapsed_time(e1) / steps
    tot, cnt = collections.Counter(), collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name
            for cut in ("<", "("):
                name = name.split(cut)[0]
            tot[name] += ev.device_time_total
            cnt[name] += 1
    with open(path, "w") as f:
        f.write(f"step {wall:.3f} ms under the profiler; kernels+copies busy {sum(tot.values()) / steps / 1e3:.3f} ms/step\n")
        for name, v in tot.most_common(60):
            f.write(f"  {v / steps / 1e3:8.3f} ms  x{cnt[name] / steps:6.1f}  avg {v / cnt[name]:8.1f} us  {name[:90]}\n")


def parity_probe(world, rank, dev, build_models, fields_of_rank, fwd_bwd, bucket_cls):
    """N > 1: one batch-sharded step (sync-BN across ranks, averaged gradient bucket) against the SAME global batch
    recomputed in one process on this rank's GPU (single-rank COMM) -- loss and every parameter gradient.  Driver-run
    evidence that the scaled numbers are numbers for the right result (tests/multi/sharded_step_parity.py is the test)."""
    import random
    import torch.distributed as dist
    import mmpde_b200.ops as ops_mod
    models = build_models()
    params = [p for m in models for p in m.parameters()]
    state0 = [{k: v.clone() for k, v in m.state_dict().items()} for m in models]
    rnd = random.Random(4321)
    steps_all = [[rnd.randrange(1, RES[0] - 1) for _ in range(BATCH)] for _ in range(world)]
    bucket = bucket_cls(params)
    loss_s = fwd_bwd(models, fields_of_rank(rank).to(dev), steps_all[rank])
    bucket.allreduce()
    dist.all_reduce(loss_s)
    loss_s = float(loss_s) / world
    g_s = [p.grad.detach().clone() if p.grad is not None else None for p in params]
    comm, ops_mod.COMM = ops_mod.COMM, ops_mod._Comm()
    try:
        for m, s0 in zip(models, state0):
            m.load_state_dict(s0)
            m.zero_grad(set_to_none=True)
        fields = torch.cat([fields_of_rank(r) for r in range(world)]).to(dev)
        loss_g = float(fwd_bwd(models, fields, [s for ss in steps_all for s in ss]))
    finally:
        ops_mod.COMM = comm
    g_g = [p.grad.detach() if p.grad is not None else None for p in params]
    worst, num, den = 0.0, 0.0, 0.0
    scale = max(float(b.norm()) for b in g_g if b is not None)
    for a, b in zip(g_s, g_g):
        if a is None or b is None:           # parameters outside the step's graph (ItpNet.layers3): no gradient either way
            continue
        nb = float(b.double().norm())
        d = float((a.double() - b.double()).norm())
        num, den = num + d * d, den + nb * nb
        if nb > 1e-4 * scale:                # (a bias in front of a BatchNorm has an analytically zero gradient: noise / noise)
            worst = max(worst, d / nb)
    out = torch.tensor([abs(loss_s - loss_g) / abs(loss_g), worst, (num / max(den, 1e-300)) ** 0.5], device=dev,
                       dtype=torch.float64)
    dist.all_reduce(out, op=dist.ReduceOp.MAX)
    del models, bucket
    return {"loss_rel": float(out[0]), "grad_rel_max_per_tensor": float(out[1]), "grad_rel_all": float(out[2]),
            "global_batch": world * BATCH,
            "what": "one sharded MM-mode step (sync-BN across ranks, averaged gradient bucket) vs the same global batch "
                    "recomputed on one GPU with the single-rank path; max over ranks (the frozen DMM mover runs in cuDNN / cuBLAS, "
                    "whose kernels and summation order depend on the batch size: the two runs see mesh coordinates that differ in "
                    "the last bits; with the same mesh the whole-gradient difference is 2.5e-7, tests/multi/sharded_step_parity.py)"}


def c4_record(rank, world, dev, steps, nodes=1000000):
    """BASELINE.json configs[3]: synthetic 1 M-node mesh, 6 layers, k = 35, fwd + bwd.  N = 1: the whole graph on the GPU.
    N > 1: graph-partitioned, one part per rank, one halo exchange per layer and direction; every rank ALSO times the
    unpartitioned graph on its own GPU, so the record carries the efficiency against N = 1 measured in the same run."""
    import numpy as np
    import torch.distributed as dist
    from mmpde_b200 import dist as mdist, ops, partition as pt
    from mmpde_b200.PDEs import burgers
    from mmpde_b200.gnn_2d import MP_PDE_Solver_2D
    side = int(round(nodes ** 0.5))
    n = side * side
    rng = np.random.default_rng(0)
    g = np.stack(np.meshgrid(np.linspace(0, 1, side), np.linspace(0, 1, side), indexing="ij"), -1).reshape(-1, 2)
    xy = torch.from_numpy((g + rng.uniform(-0.3, 0.3, g.shape) / (side - 1)).astype(np.float32)).to(dev)
    xy = xy[pt.morton_order(xy)].contiguous()
    edges = ops.EdgeList.from_knn(ops.knn_indices_grid(xy, xy, K_NEIGH, 0, True), has_pad=False)
    torch.manual_seed(0)
    u = torch.randn(n, 1, device=dev)
    pos = torch.cat((torch.full((n, 1), 7.0, device=dev), xy), 1)
    r = torch.randn(n, 1, device=dev)
    model = MP_PDE_Solver_2D(burgers()).to(dev).train()
    params = list(model.parameters())

    class Whole:
        pass
    whole = Whole()
    whole.x, whole.pos, whole.edge_index, whole.batch, whole._edges = u, pos, None, None, edges

    def timed(fn):
        for _ in range(2):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def whole_step():
        model.zero_grad(set_to_none=True)
        ((model(whole) * r).sum() / n).backward()

    comm = ops.COMM
    ops.COMM = ops._Comm()
    try:
        ms1 = timed(whole_step)
    finally:
        ops.COMM = comm
    rec = {"workload": f"synthetic {n}-node jittered lattice (Morton order), k=35, 6 layers, hidden 128, fwd+bwd",
           "nodes": n, "edges": int(edges.n_edges), "n1_ms_per_step": ms1,
           "n1_edge_updates_per_s": edges.n_edges * LAYERS / (ms1 * 1e-3)}
    if world > 1:
        (part,), (plan,) = pt.split_graph(u, pos, edges.src, edges.dst, world, ranks=[rank])
        exch = mdist.HaloExchange(plan)
        bucket = mdist.GradBucket(params)
        n_edges_total = int(edges.n_edges)
        del edges, whole
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()

        def part_step():
            model.zero_grad(set_to_none=True)
            (out,) = model.forward_partitioned([part], exch)
            ((out * r[plan.owned]).sum() / n).backward()
            bucket.allreduce(average=False)

        ms = timed(part_step)
        halo = torch.tensor([plan.n_halo], device=dev)
        dist.all_reduce(halo, op=dist.ReduceOp.MAX)
        rec.update(ms_per_step=ms, edge_updates_per_s=n_edges_total * LAYERS / (ms * 1e-3), parts=world,
                   halo_rows_max=int(halo), halo_bytes_per_layer_per_direction=int(halo) * H * 4,
                   efficiency_vs_n1=ms1 / (world * ms), exchange=type(exch).__name__ + getattr(exch, "kind", ""))
    else:
        rec.update(ms_per_step=ms1, edge_updates_per_s=rec["n1_edge_updates_per_s"], parts=1)
    return rec


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    from mmpde_b200 import _cabi, dist as mdist, synthetic
    from mmpde_b200.PDEs import burgers
    from mmpde_b200.data_creator_2d import GraphCreator_FS_2D
    from mmpde_b200.gnn_2d import MP_PDE_Solver_2D
    from mmpde_b200.interpolate import ItpNet
    from mmpde_b200.mesh.dmm_model import DMM
    from mmpde_b200.mmpde import criterion
    from mmpde_b200.train_helper_2d import StepGraph, _forward_gnn, _overlap_solvers, test_timestep_losses, training_loop_branch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    rank, world, dev = mdist.init_from_env()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    _cabi.lib()
    peaks = _peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    def workload(kind):
        """Models, graph creator and this rank's trajectories of one BASELINE.json training workload."""
        torch.manual_seed(0)
        if kind == "cylinder":                   # configs[2]: flow around a cylinder, base_resolution 30,2521
            from mmpde_b200.PDEs import cy
            cloud = synthetic.cylinder_cloud(CY_RES[1], seed=0)
            pde = cy(ori_grid=cloud, device=dev)
            pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = CY_RES
            w = {"gc": GraphCreator_FS_2D(pde, K_NEIGH, "knn", 1, CY_RES[0]), "nodes_per_sample": CY_RES[1],
                 "net": ItpNet(CY_RES[1], None, [128, 64], [128, 64], [1, 4, 16, 4, 1]).to(dev),
                 "name": "Flow around a cylinder MM-PDE training step (DMM-moved nodes [default-initialised DMM, graph mode] + "
                         "interpolation + 2x 6-layer processor), 30x2521 unstructured nodes, k=35",
                 "mover": lambda: DMM(mode="graph", grid=cloud.to(dev), **DMM_GRAPH),
                 "fields": lambda r: synthetic.cylinder_fields(BATCH, cloud, CY_RES[0], seed=100 + r)}
        else:
            pde = burgers()
            pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = RES
            w = {"gc": GraphCreator_FS_2D(pde, K_NEIGH, "knn", 1, RES[0]), "nodes_per_sample": RES[1] * RES[2],
                 "net": ItpNet(RES[1], RES[2], [128, 64], [128, 64], [1, 4, 16, 4, 1]).to(dev), "name": WORKLOAD_BURGERS,
                 "mover": lambda: DMM(s=RES[1], mode="array", **DMM_ARRAY),
                 "fields": lambda r: synthetic.burgers_fields(BATCH, RES[0], RES[1], RES[2], seed=100 + r)}
        w["pde"] = pde
        w["model"], w["model_b"] = MP_PDE_Solver_2D(pde).to(dev), MP_PDE_Solver_2D(pde).to(dev)
        torch.manual_seed(MOVER_SEED)
        w["mover"] = w["mover"]().to(dev).eval()
        w["params"] = [p for m in (w["model"], w["model_b"], w["net"]) for p in m.parameters()]
        w["opt"] = torch.optim.AdamW([{"params": w["model"].parameters()}, {"params": w["model_b"].parameters()},
                                      {"params": w["net"].parameters()}], lr=2e-3, capturable=not args.no_graph,
                                     fused=True if not args.no_graph else None)
        w["step_graph"] = None if args.no_graph else StepGraph()
        w["bucket"] = mdist.GradBucket(w["params"]) if world > 1 else None
        for m in (w["model"], w["model_b"], w["net"]):
            m.train()
        return w

    def stepper(w):
        after = w["bucket"].allreduce if w["bucket"] is not None else None

        def train_step(fields, graph="default"):
            return training_loop_branch(w["model"], w["model_b"], w["net"], w["mover"], [0], BATCH, w["opt"], None,
                                        [(fields, fields)], w["gc"], criterion, dev, after_backward=after,
                                        step_graph=w["step_graph"] if graph == "default" else graph)
        return train_step

    cyl = args.workload == "cylinder"
    # ---- N > 1: parity probe of the sharded step, before anything is recorded -------------------------------
    probe = None
    if world > 1 and not cyl and not args.no_extras:
        def build_models():
            w0 = workload("burgers")
            probe_ctx["w"] = w0
            return (w0["model"], w0["model_b"], w0["net"])

        probe_ctx = {}

        def fwd_bwd(models, fields, steps):
            w0 = probe_ctx["w"]
            for m in models:
                m.zero_grad(set_to_none=True)
            data, labels = w0["gc"].create_data(fields, steps)
            pred = _forward_gnn(models[0], models[1], models[2], w0["mover"], GraphCreator_FS_2D(w0["pde"], K_NEIGH, "knn", 1, RES[0]),
                                data, labels, steps, dev)
            loss = criterion(pred, labels.to(dev).reshape(-1, 1))
            loss.backward()
            return loss.detach()

        probe = parity_probe(world, rank, dev, build_models,
                             lambda r: synthetic.burgers_fields(BATCH, RES[0], RES[1], RES[2], seed=100 + r), fwd_bwd,
                             mdist.GradBucket)
        probe_ctx.clear()
        torch.cuda.empty_cache()

    w = workload(args.workload)
    step_graph = w["step_graph"]
    train_step = stepper(w)
    nodes_per_sample = w["nodes_per_sample"]
    # weak scaling: every rank owns its own batch of 16 trajectories (global batch 16*G), seeded per rank
    fields_host = w["fields"](rank).pin_memory()
    fields_dev = fields_host.to(dev)
    n_nodes = BATCH * nodes_per_sample
    n_edges = n_nodes * K_NEIGH
    edge_updates_per_step = 2 * n_edges * LAYERS

    import random
    random.seed(1234 + rank)
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        train_step(fields_dev)

    # ---- device-resident timing (value) ---------------------------------------------------------------------
    launches0 = _cabi.launches
    torch.cuda.profiler.start()          # ncu --profile-from-start off captures exactly the timed region
    t_begin = time.perf_counter()
    ms_step = timed(lambda: train_step(fields_dev), args.steps)
    t_end = time.perf_counter()
    torch.cuda.profiler.stop()
    launches = _cabi.launches - launches0
    if os.environ.get("MMPDE_KINETO"):   # CUPTI durations of every kernel inside the replayed step (any N; rank 0 writes the table)
        kineto_table(lambda: train_step(fields_dev), 3, os.environ["MMPDE_KINETO"] if rank == 0 else None, barrier)
    # ---- every kernel of the step, timed with events on its launch stream in an eager pass of the SAME step ---
    # (with the step replayed from a CUDA graph the kernels inside it cannot carry events; --no-graph: same eager path)
    kern_steps = min(args.steps, 3)
    kernels, by = profile_kernels(lambda: train_step(fields_dev, graph=None), kern_steps, peaks, barrier)
    dom = by.get("mmpde_edge_bwd", {"n": 0, "ms": 0.0})
    kern_avg_ms = dom["ms"] / max(dom["n"], 1)

    # ---- end to end through the public API with host buffers -----------------------------------------------
    # Every step: H2D of that step's input slices from the pinned trajectory batch (inside create_graph) and a D2H
    # read of that step's loss.  The read is the usual asynchronous logging pattern: the loss is copied to pinned
    # host memory on the stream and looked at one step later (the last one before the clock stops), so the host
    # keeps queueing work instead of draining the GPU after every step.
    host_loss = torch.zeros(2, dtype=torch.float32).pin_memory()     # two slots: step i is in flight while i-1 is read
    seen = []

    def e2e_run(steps):
        pending = None
        for i in range(steps):
            losses = train_step(fields_host)       # create_graph moves the step's slices host -> device
            host_loss[i & 1:(i & 1) + 1].copy_(losses[-1].reshape(1), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:                # step i is queued: now look at the loss of step i-1
                pending[0].synchronize()
                seen.append(float(host_loss[pending[1]]))
            pending = (ev, i & 1)
        pending[0].synchronize()
        seen.append(float(host_loss[pending[1]]))  # the last step's loss is read inside the timed region too

    e2e_run(1)
    n_seen = len(seen)
    ms_e2e = timed(lambda: e2e_run(args.steps), 1) / args.steps
    assert len(seen) - n_seen == args.steps and all(v == v for v in seen), "every timed step must deliver its loss"
    clocks = sampler.stop(t_begin, time.perf_counter())
    h2d = 2 * BATCH * nodes_per_sample * 4 + (BATCH * 8 if step_graph is not None else 0)   # data + labels slices (+ step indices)
    d2h = 4

    # ---- rollout (teacher-forced per-time-step test sweep, no_grad) ----------------------------------------
    w["model"].eval(); w["model_b"].eval(); w["net"].eval()

    def rollout_step():
        test_timestep_losses(w["model"], w["model_b"], w["net"], w["mover"], [7], BATCH, [(fields_dev, fields_dev)], w["gc"],
                             criterion, dev, step_graph=step_graph)

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(3):
            rollout_step()
        ms_roll = timed(rollout_step, max(args.steps, 3))

    value = world * edge_updates_per_step / (ms_step * 1e-3)
    e2e_value = world * edge_updates_per_step / (ms_e2e * 1e-3)
    alg_flops = 2 * FLOP_PER_EDGE_FWD * n_edges    # backward = 2x forward FLOPs (dgrad + wgrad)
    exe_flops = 2 * 3 * 2 * H * H * n_edges        # executed on the tensor pipe: 2 contractions x 3 split-bf16 products
    achieved = alg_flops / (kern_avg_ms * 1e-3) / 1e12 if kern_avg_ms > 0 else 0.0
    cap = _ncu_capture()

    # ---- the other BASELINE.json configurations as sub-records of the same line ------------------------------
    extras = {}
    if not args.no_extras and not cyl:
        if step_graph is not None:
            step_graph.release()
        del train_step
        torch.cuda.empty_cache()
        wc = workload("cylinder")
        cstep = stepper(wc)
        cf = wc["fields"](rank).to(dev)
        for _ in range(4):
            cstep(cf)
        ms_c = timed(lambda: cstep(cf), max(args.steps // 2, 3))
        e_c = BATCH * CY_RES[1] * K_NEIGH
        extras["cylinder"] = {"workload": wc["name"], "per_gpu_batch": BATCH, "ms_per_step": ms_c,
                              "edge_updates_per_s": world * 2 * e_c * LAYERS / (ms_c * 1e-3),
                              "parallelism": f"batch-sharded dp{world}"}
        if wc["step_graph"] is not None:
            wc["step_graph"].release()
        del wc, cstep, cf
        torch.cuda.empty_cache()
        extras["c4"] = c4_record(rank, world, dev, steps=2)
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not cyl:      # the CPU arm times the headline workload
        v, sec, cores, timed_n = cpu_reference_step_time(BATCH, 1, 1)
        cpu = {"value": v, "unit": "edge-updates/s", "cores": cores, "kind": "port",
               "sample": f"the full batch-{BATCH} step, 1 warm-up + {timed_n} timed step ({sec:.1f} s); oracle port "
                         "(reference needs PyG, not installable offline)"}
    if rank == 0:
        line = {
            "metric": "edge-updates/sec (fwd+bwd)", "value": value, "unit": "edge-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "per_gpu_batch": BATCH, "nodes_per_gpu": n_nodes,
                       "edges_per_graph": n_edges,
                       "parallelism": f"batch-sharded dp{world}, sync-BN ("
                                      + ("sums exchanged over NVLink peer memory in one kernel" if getattr(mdist.ops.COMM, "peer", None) is not None
                                         else "NCCL all-reduce of the sums" if world > 1 else "single rank") + "), flat grad all-reduce",
                       "launch": "eager" if step_graph is None else "CUDA graph replay of the whole step (StepGraph)"
                                 + (", the two solvers as parallel graph branches whose persistent kernels run side by side at "
                                    "half width (74 CTAs each, train_helper_2d._branch_ctas)" if _overlap_solvers(dev) else ""),
                       "l2": "per-step working set ~1.7 GB > 126 MB L2, no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "edge-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": "mmpde_edge_bwd", "bound": "tensor", "achieved": achieved, "peak": peaks["bf16"],
                         "unit": "TFLOP/s", "frac": achieved / peaks["bf16"], "traffic": _ncu_traffic(),
                         "traffic_source": None if cap is None else {k: cap.get(k) for k in ("file", "gpu_time_us", "source_sha16", "build")},
                         "avg_launch_ms": kern_avg_ms, "launches_timed": dom["n"],
                         "algorithmic_flops_per_launch": alg_flops, "executed_flops_per_launch": exe_flops,
                         "frac_executed": exe_flops / (kern_avg_ms * 1e-3) / 1e12 / peaks["bf16"] if kern_avg_ms > 0 else None,
                         "peak_source": peaks["source"],
                         "timed_in": "eager single-stream pass of the same step right after the timed region, every kernel alone on the "
                                     "whole chip (nodes of a replayed CUDA graph cannot carry events; in the replayed step the same kernel "
                                     "runs on half the SMs beside the other solver's, for twice as long)"
                                     if step_graph is not None else "eager pass after the timed region",
                         "share_of_step": kern_avg_ms * dom["n"] / kern_steps / ms_step if ms_step > 0 else None},
            "kernels": kernels,
            "rollout": {"steps_per_s": world * 1e3 / ms_roll, "ms_per_step": ms_roll,
                        "definition": "one pass of train_helper_2d.py:173-185 for one batch of 16, eval, no_grad"},
        }
        if probe is not None:
            line["parity_probe"] = probe
        line.update(extras)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    if step_graph is not None:
        step_graph.release()             # recorded graphs hold captured NCCL work: drop them before the communicator
    if world > 1:
        # The line is out.  Leave without the interpreter / NCCL teardown: with collectives captured in CUDA graphs
        # destroy_process_group() was seen to block after the run had finished (profiles/r01_bench_2gpu_v12.json).
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", dest="no_graph", action="store_true",
                    help="queue every launch from Python instead of replaying the recorded step")
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-extras", dest="no_extras", action="store_true",
                    help="skip the parity probe (N > 1) and the cylinder / c4 sub-records")
    ap.add_argument("--workload", default="burgers", choices=["burgers", "cylinder"],
                    help="burgers = BASELINE.json configs[1] (the headline, default); cylinder = configs[2]")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: whatever libraries write to file descriptor 1 (the NCCL version
    # banner, cuDNN notices) is sent to stderr, the result goes to a private copy of the original stdout.
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

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
  cpu_baseline : the oracle port of the reference path on the host cores, bounded sample, rank 0, N=1.
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


RESULT_OUT = sys.stdout


def _ncu_traffic():
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel, from the committed ncu --set full capture
    (profiles/r01_edge_bwd_ncu.json, written by profiles/ncu_summary.py --json on this CPU box)."""
    path = os.path.join(ROOT, "profiles", "r01_edge_bwd_ncu.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("dram_bytes_read", 0) + d.get("dram_bytes_write", 0)
    return None


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
    mover = synthetic.AnalyticMover()
    return gc, model, model_b, net, opt, fields, mover, loops


def cpu_reference_step_time(sample_batch, steps, warmup):
    """Times the oracle's training_loop_branch body on the host cores for a `sample_batch`-trajectory batch."""
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
    n = RES[1] * RES[2]
    edge_updates = 2 * sample_batch * n * K_NEIGH * LAYERS
    best = min(times)
    return edge_updates / best, best, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 2
    # exactly --steps timed steps of the bounded sample (1.5 s each on the GPU box's 16 cores), at most 2 warm-ups
    value, sec, cores = cpu_reference_step_time(sample, max(1, args.steps), min(max(args.warmup, 0), 2))
    line = {
        "impl": "reference", "metric": "edge-updates/sec (fwd+bwd)", "value": value, "unit": "edge-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Burgers 2D MM-PDE training step (moved mesh + interpolation + 2x 6-layer processor), "
                               "31x48x48, k=35", "per_gpu_batch": BATCH, "timed_sample_batch": sample},
        "cpu_baseline": {"value": value, "unit": "edge-updates/s", "cores": cores, "kind": "port",
                         "sample": f"batch {sample} of the batch-{BATCH} step (per-edge cost is size-independent); "
                                   "plain-torch oracle port of the reference path incl. sklearn kd-tree kNN; "
                                   "the reference itself needs torch_geometric/torch_cluster, not installable offline"},
        "e2e": {"value": value, "unit": "edge-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    from mmpde_b200 import _cabi, dist as mdist, synthetic
    from mmpde_b200.PDEs import burgers
    from mmpde_b200.data_creator_2d import GraphCreator_FS_2D
    from mmpde_b200.gnn_2d import MP_PDE_Solver_2D
    from mmpde_b200.interpolate import ItpNet
    from mmpde_b200.mmpde import criterion
    from mmpde_b200.train_helper_2d import StepGraph, test_timestep_losses, training_loop_branch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    rank, world, dev = mdist.init_from_env()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    _cabi.lib()
    torch.manual_seed(0)
    cyl = args.workload == "cylinder"
    if cyl:                                  # BASELINE.json configs[2]: flow around a cylinder, base_resolution 30,2521
        from mmpde_b200.PDEs import cy
        cloud = synthetic.cylinder_cloud(CY_RES[1], seed=0)
        pde = cy(ori_grid=cloud, device=dev)
        pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = CY_RES
        gc = GraphCreator_FS_2D(pde, K_NEIGH, "knn", 1, CY_RES[0])
        net = ItpNet(CY_RES[1], None, [128, 64], [128, 64], [1, 4, 16, 4, 1]).to(dev)
        nodes_per_sample, workload = CY_RES[1], ("Flow around a cylinder MM-PDE training step (moved nodes + interpolation + 2x 6-layer "
                                                 "processor), 30x2521 unstructured nodes, k=35")
    else:
        pde = burgers()
        pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = RES
        gc = GraphCreator_FS_2D(pde, K_NEIGH, "knn", 1, RES[0])
        net = ItpNet(RES[1], RES[2], [128, 64], [128, 64], [1, 4, 16, 4, 1]).to(dev)
        nodes_per_sample, workload = RES[1] * RES[2], ("Burgers 2D MM-PDE training step (moved mesh + interpolation + 2x 6-layer "
                                                       "processor), 31x48x48, k=35")
    model, model_b = MP_PDE_Solver_2D(pde).to(dev), MP_PDE_Solver_2D(pde).to(dev)
    mover = synthetic.AnalyticMover().to(dev)
    params = [p for m in (model, model_b, net) for p in m.parameters()]
    opt = torch.optim.AdamW([{"params": model.parameters()}, {"params": model_b.parameters()},
                             {"params": net.parameters()}], lr=2e-3, capturable=not args.no_graph,
                            fused=True if not args.no_graph else None)
    step_graph = None if args.no_graph else StepGraph()
    bucket = mdist.GradBucket(params) if world > 1 else None
    after = bucket.allreduce if bucket is not None else None
    # weak scaling: every rank owns its own batch of 16 trajectories (global batch 16*G), seeded per rank
    if cyl:
        fields_host = synthetic.cylinder_fields(BATCH, cloud, CY_RES[0], seed=100 + rank).pin_memory()
    else:
        fields_host = synthetic.burgers_fields(BATCH, RES[0], RES[1], RES[2], seed=100 + rank).pin_memory()
    fields_dev = fields_host.to(dev)
    n_nodes = BATCH * nodes_per_sample
    n_edges = n_nodes * K_NEIGH
    edge_updates_per_step = 2 * n_edges * LAYERS

    model.train(); model_b.train(); net.train()

    def train_step(fields, graph=step_graph):
        return training_loop_branch(model, model_b, net, mover, [0], BATCH, opt, None, [(fields, fields)], gc,
                                    criterion, dev, after_backward=after, step_graph=graph)

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

    import random
    random.seed(1234 + rank)
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        train_step(fields_dev)

    # ---- device-resident timing (value) with per-launch events on the dominant kernel -----------------------
    _cabi_profile = []
    real_call = _cabi.call

    def profiled_call(name, *a):
        if name == "mmpde_edge_bwd":
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = real_call(name, *a)
            e.record()
            _cabi_profile.append((s, e))
            return rc
        return real_call(name, *a)

    import mmpde_b200.ops as ops_mod
    # With the step replayed from a CUDA graph the kernels inside it cannot carry events, so the dominant kernel is
    # timed in an eager pass of the SAME step (same inputs, same process) right after the timed region; without the
    # graph (--no-graph) the events sit inside the timed region itself.
    if step_graph is None:
        ops_mod._cabi.call = profiled_call
    launches0 = _cabi.launches
    torch.cuda.profiler.start()          # ncu --profile-from-start off captures exactly the timed region
    t_begin = time.perf_counter()
    ms_step = timed(lambda: train_step(fields_dev), args.steps)
    t_end = time.perf_counter()
    torch.cuda.profiler.stop()
    launches = _cabi.launches - launches0
    if step_graph is not None:
        ops_mod._cabi.call = profiled_call
        os.environ["MMPDE_OVERLAP_SOLVERS"] = "0"      # time the kernel alone: on one stream, not beside the other solver's kernels
        barrier()
        for _ in range(args.steps):
            train_step(fields_dev, graph=None)
        barrier()
        os.environ.pop("MMPDE_OVERLAP_SOLVERS")
    ops_mod._cabi.call = real_call
    kern_ms = [s.elapsed_time(e) for s, e in _cabi_profile]
    kern_avg_ms = sum(kern_ms) / max(len(kern_ms), 1)

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
    model.eval(); model_b.eval(); net.eval()

    def rollout_step():
        test_timestep_losses(model, model_b, net, mover, [7], BATCH, [(fields_dev, fields_dev)], gc, criterion, dev,
                             step_graph=step_graph)

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(3):
            rollout_step()
        ms_roll = timed(rollout_step, max(args.steps, 3))

    value = world * edge_updates_per_step / (ms_step * 1e-3)
    e2e_value = world * edge_updates_per_step / (ms_e2e * 1e-3)
    peaks = _peaks()
    alg_flops = 2 * FLOP_PER_EDGE_FWD * n_edges    # backward = 2x forward FLOPs (dgrad + wgrad)
    achieved = alg_flops / (kern_avg_ms * 1e-3) / 1e12 if kern_avg_ms > 0 else 0.0

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not cyl:      # the CPU arm times the headline workload
        v, sec, cores = cpu_reference_step_time(4, 3, 1)
        cpu = {"value": v, "unit": "edge-updates/s", "cores": cores, "kind": "port",
               "sample": f"batch 4 of the batch-{BATCH} step, 1 warm-up + 3 timed steps (best {sec:.1f} s); oracle port "
                         "(reference needs PyG, not installable offline)"}
    if rank == 0:
        line = {
            "metric": "edge-updates/sec (fwd+bwd)", "value": value, "unit": "edge-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "per_gpu_batch": BATCH, "nodes_per_gpu": n_nodes,
                       "edges_per_graph": n_edges,
                       "parallelism": f"batch-sharded dp{world}, sync-BN ("
                                      + ("sums exchanged over NVLink peer memory in one kernel" if getattr(ops_mod.COMM, "peer", None) is not None
                                         else "NCCL all-reduce of the sums" if world > 1 else "single rank") + "), flat grad all-reduce",
                       "launch": "eager" if step_graph is None else "CUDA graph replay of the whole step (StepGraph)"
                                 + (", the two solvers as parallel graph branches" if world == 1 else ""),
                       "l2": "per-step working set ~1.7 GB > 126 MB L2, no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "edge-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": "mmpde_edge_bwd", "bound": "tensor", "achieved": achieved, "peak": peaks["bf16"],
                         "unit": "TFLOP/s", "frac": achieved / peaks["bf16"], "traffic": _ncu_traffic(),
                         "avg_launch_ms": kern_avg_ms, "launches_timed": len(kern_ms),
                         "algorithmic_flops_per_launch": alg_flops, "peak_source": peaks["source"],
                         "timed_in": "timed region" if step_graph is None else "eager single-stream pass of the same step after the "
                                     "timed region (graph nodes cannot carry events)",
                         "share_of_step": kern_avg_ms * len(kern_ms) / args.steps / ms_step if ms_step > 0 else None},
            "rollout": {"steps_per_s": world * 1e3 / ms_roll, "ms_per_step": ms_roll,
                        "definition": "one pass of train_helper_2d.py:173-185 for one batch of 16, eval, no_grad"},
        }
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

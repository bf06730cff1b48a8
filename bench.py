#!/usr/bin/env python
"""bench.py -- points/sec of the PNEConvLayerRotEquiv hot path (hierarchy + frames + neighbourhoods +
21 fused convolutions, forward and backward) on the dfaust_I_rot_pca_2F shapes (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU port of the reference path

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every key.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CLOUDS, N_POINTS = 32, 6890
METRIC = "points/sec PNEConvLayerRotEquiv fwd+bwd (dfaust_I_rot_pca_2F hot path, F=2)"
WORKLOAD = ("dfaust_I_rot_pca_2F hot path: grid hierarchy (0.04;0.05,0.1,0.2,0.4) + kNN16/PCA frames + ball-query CSRs + 21 "
            "PNEConvLayerRotEquiv fwd+bwd, 32 clouds x 6890 points per GPU, F=2")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", type=int, default=int(os.environ.get("SE3_PRECISION", "1")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="stack", choices=["stack", "fpn", "sweep"],
                    help="stack: hierarchy + the 21 convolutions fwd+bwd (BASELINE metric); fpn: the full FPN training step")
    ap.add_argument("--strong", action="store_true", help="32 clouds in total (32/G per GPU) instead of 32 per GPU")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra measurements (other configs / precisions)")
    ap.add_argument("--thread", action="store_true",
                    help="pipelined mode: build the next hierarchy from a worker thread (measured slower: 5.37 vs 5.17 ms)")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="stack workload: build every hierarchy on the convolutions' stream (no overlap of the next batch's "
                         "hierarchy with the current batch's convolutions)")
    return ap.parse_args()


class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes(spec_sizes, s=4):
    """SURVEY 8(d) gather model: bytes one conv forward / backward must move; s = bytes per feature element (4 = the
    API's fp32 features, the SURVEY's definition; 2 = the bf16 rows the precision-1 kernels actually gather)."""
    fwd = bwd = 0.0
    for (m, e, f, cin, cout) in spec_sizes:
        kbar = e / max(m, 1)
        w = (cin * 32 * cout + 9 * 32 + 32) * 4
        fwd += m * (kbar * (4 + 12 + 36 * f + f * cin * s) + (12 + 36 * f + 4) + f * cout * s) + w
        bwd += m * (kbar * (2 * (4 + 12 + 36 * f) + f * cin * s + f * cout * s) + f * cout * s + f * cin * s) + 2 * w
    return fwd, bwd


def kernel_algorithmic_bytes(size, s=4):
    """The same model split over the three gather kernels: the forward figure belongs to the
    forward aggregation; the backward figure is its two gather passes -- the edge-gradient kernel (gathers x
    rows, reads the output gradient) and the transposed aggregation (gathers dy rows, produces dx)."""
    (m, e, f, cin, cout) = size
    kbar = e / max(m, 1)
    geo = 4 + 12 + 36 * f
    w = (cin * 32 * cout + 9 * 32 + 32) * 4
    return [m * (kbar * (geo + f * cin * s) + (12 + 36 * f + 4) + f * cout * s) + w,      # forward aggregation
            m * (kbar * (geo + f * cout * s) + f * cin * s) + w,                          # transposed aggregation
            m * (kbar * (geo + f * cin * s) + f * cout * s) + w]                          # edge gradient


def run_reference(args, rank, world):
    """CPU port of the reference path (oracle/), all host threads, bounded sample: ONE cloud per step."""
    if rank != 0:
        return
    import torch
    from oracle import hierarchy_oracle as ho
    from se3conv3d_b200 import workloads as wl
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    pts, batch = wl.synthetic_bodies(1, N_POINTS, seed=0)
    specs = wl.dfaust_conv_specs()
    torch.manual_seed(0)
    params = []
    for (_, _, _, _, cin, cout) in specs:
        params.append((torch.empty(9, 32).uniform_(-1 / 3, 1 / 3), torch.zeros(32),
                       torch.empty(cin, 32, cout).uniform_(-1, 1) / (cin * 32) ** 0.5))
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        ho.dfaust_step_cpu(pts.numpy(), batch.numpy(), specs, params, wl.DFAUST_CFG)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = N_POINTS * len(times) / total
    sample = "1 of %d clouds (%d points) per step through the full hot path, oracle port (torch CPU + C)" % (
        N_CLOUDS, N_POINTS)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "points_per_gpu": N_CLOUDS * N_POINTS, "precision": 0,
                   "sample": sample + "; per-cloud cost is independent of the other clouds (no edge crosses two clouds), so "
                             "points/s of one cloud = points/s of the 32-cloud batch on the same cores"},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cuda_timed(torch, fn, steps, warmup, flush=None, after=None):
    """Total device milliseconds of `steps` calls of fn (CUDA events on the current stream, L2 flushed in between)."""
    for _ in range(warmup):
        fn()
        if after:
            after()
    tot = 0.0
    for _ in range(steps):
        if flush is not None:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
        if after:
            after()
    return tot


def load_seg_models():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import stage_reference_models as srm
        return srm.import_models("seg_models")
    except Exception:
        return None


def measure_fpn(torch, dist, shard, wl, dev, rank, world, precision, steps, warmup, flush, strong=False):
    """Full FPN training step (unmodified reference model code): device-resident and end to end."""
    seg = load_seg_models()
    if seg is None:
        return {"unavailable": "reference model sources are not staged (tools/stage_reference_models.py)"}
    n_clouds = max(N_CLOUDS // world, 1) if strong else N_CLOUDS
    pts_h, batch_h = wl.synthetic_bodies(n_clouds, N_POINTS, seed=1000 + rank)
    labels_h = torch.randint(0, 20, (pts_h.shape[0],), generator=torch.Generator().manual_seed(7 + rank))
    pts_h, batch_h, labels_h = pts_h.pin_memory(), batch_h.pin_memory(), labels_h.pin_memory()
    if precision == 1:
        # bf16 mode: the blocks' Linear layers (torch matmuls) run on tf32 tensor cores instead of SIMT fp32
        torch.backends.cuda.matmul.allow_tf32 = True
    step = wl.FpnStep(dev, seg, precision=precision)
    pts_d, batch_d, labels_d = pts_h.to(dev), batch_h.to(dev), labels_h.to(dev)
    from se3conv3d_b200 import _lib
    l0 = _lib.launch_count()
    ms_dev = cuda_timed(torch, lambda: step.step(pts_d, batch_d, labels_d, n_clouds), steps, warmup, flush)
    launches = (_lib.launch_count() - l0) // (steps + warmup)

    def e2e():
        p = pts_h.to(dev, non_blocking=True)
        b = batch_h.to(dev, non_blocking=True)
        lab = labels_h.to(dev, non_blocking=True)
        return float(step.step(p, b, lab, n_clouds).item())
    ms_e2e = cuda_timed(torch, e2e, steps, max(warmup, 3), flush)
    shard.barrier(dev)
    ms_dev = shard.max_over_ranks(ms_dev, dev)
    ms_e2e = shard.max_over_ranks(ms_e2e, dev)
    n_pts = n_clouds * N_POINTS * world
    g = world
    return {"metric": "points/sec full FPN training step (FPNSegUNetMLPGeluRotEqFAUST fwd + CE loss + bwd + grad all-reduce "
                      "+ clip + AdamW, hierarchy rebuilt every step)",
            "value": n_pts * steps / (ms_dev * 1e-3), "unit": "points/s", "ms_per_step": ms_dev / steps,
            "e2e": {"value": n_pts * steps / (ms_e2e * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": int(pts_h.numel() * 4 + batch_h.numel() * 4 + labels_h.numel() * 8),
                    "d2h_bytes_per_step": 4},
            "scaling": "strong" if strong else "weak", "clouds_per_gpu": n_clouds, "precision": precision,
            "linear_layers": "tf32 (torch.backends.cuda.matmul.allow_tf32)" if precision == 1 else "fp32",
            "model_parameters": int(sum(p.numel() for p in step.model.parameters())),
            "allreduce": {"collective": ("NCCL all-reduce (AVG) of one flat fp32 gradient buffer, %d buckets launched from "
                                         "post-accumulate hooks during backward" % len(step.reducer.buckets)) if g > 1 else
                                        "none (one replica; gradients are not flattened)",
                          "bytes_per_step": int(step.reducer.bytes),
                          "nvlink_bytes_per_gpu_per_step": int(2 * (g - 1) / g * step.reducer.bytes) if g > 1 else 0},
            "own_kernel_launches_per_step": int(launches)}


def measure_sweep(torch, shard, dev, world, precision, flush):
    """BASELINE configs[4]: N in {64 k, 256 k, 1 M, 4 M} x F in {1, 2, 4} x (Cin, Cout) in {32, 64, 128, 256}, restricted to
    the cases whose [N F, Cin K] tile fits comfortably (4 M points only at 32 channels, 1 M up to 64, F = 4 up to 64 channels);
    every rank runs every case on its own replica, times are the max over ranks, points/s is the whole job's."""
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    from se3conv3d_b200.pc import PointcloudRotEquiv, BQNeighborhood
    rows = []
    for n in (65536, 262144, 1048576, 4194304):
        side = (n / 8192) ** (1.0 / 3.0)
        pts = torch.rand(n, 3, generator=torch.Generator().manual_seed(0)) * side
        for f in (1, 2, 4):
            chans = [c for c in (32, 64, 128, 256)
                     if n * f * c * 32 * 2 * 4 <= 40e9 and not (f == 4 and c > 64) and not (n > 1048576 and f > 2)]
            if not chans:
                continue
            cfg = {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False, "n_frames": f}
            pc = PointcloudRotEquiv(pts.to(dev), torch.zeros(n, dtype=torch.int32, device=dev), cfg)
            nb = BQNeighborhood(pc, pc, 0.1)
            e = int(nb.neighbors_.shape[0])
            for c in chans:
                torch.manual_seed(2)
                layer = PNEConvLayerRotEquiv(9, c, c, 32, "mlp_gelu").to(dev)
                layer.precision = precision
                layer.norm_neigh_dist_.fill_(10.0)
                layer.norm_num_neighs_.fill_(n / max(e, 1))
                x = torch.randn(n * f, c, device=dev, requires_grad=True)
                dy = torch.randn(n * f, c, device=dev) / c ** 0.5

                def one():
                    y = layer(pc, pc, x, nb)
                    y.backward(dy)
                shard.barrier(dev)
                ms = shard.max_over_ranks(cuda_timed(torch, one, 3, 2, flush) / 3, dev)
                fb, bb = algorithmic_bytes([(n, e, f, c, c)])
                rows.append({"n": n, "edges": e, "f": f, "c_in": c, "c_out": c, "ms_fwd_bwd": ms,
                             "points_per_s": n * world / (ms * 1e-3), "alg_gbs_s4_per_gpu": (fb + bb) / (ms * 1e-3) / 1e9})
                layer = x = dy = one = None
                torch.cuda.empty_cache()
            pc = nb = None
            torch.cuda.empty_cache()
    return rows


def measure_extras(torch, wl, dev, precision, flush):
    """Other BASELINE configs / precisions on the same kernels (a few iterations each; not the headline)."""
    from se3conv3d_b200.layers import PNEConvLayerRotEquiv
    from se3conv3d_b200.pc import PointcloudRotEquiv, BQNeighborhood
    out = {}

    def layer_case(n, side, r, f, cin, cout, prec, reps=5):
        pts = torch.rand(n, 3, generator=torch.Generator().manual_seed(0)) * side
        cfg = {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False, "n_frames": f}
        pc = PointcloudRotEquiv(pts.to(dev), torch.zeros(n, dtype=torch.int32, device=dev), cfg)
        nb = BQNeighborhood(pc, pc, r)
        e = int(nb.neighbors_.shape[0])
        torch.manual_seed(2)
        layer = PNEConvLayerRotEquiv(9, cin, cout, 32, "mlp_gelu").to(dev)
        layer.precision = prec
        layer.norm_neigh_dist_.fill_(1.0 / r)
        layer.norm_num_neighs_.fill_(n / max(e, 1))
        x = torch.randn(n * f, cin, device=dev, requires_grad=True)
        dy = torch.randn(n * f, cout, device=dev) / cout ** 0.5

        def one():
            y = layer(pc, pc, x, nb)
            y.backward(dy)
        ms = cuda_timed(torch, one, reps, 2, flush) / reps
        fb, bb = algorithmic_bytes([(n, e, f, cin, cout)])
        return {"n": n, "edges": e, "f": f, "c_in": cin, "c_out": cout, "precision": prec, "ms_fwd_bwd": ms,
                "points_per_s": n / (ms * 1e-3), "alg_gbs_s4": (fb + bb) / (ms * 1e-3) / 1e9}
    # BASELINE configs[0]: one 8192-point cloud, F=2, 32 -> 64, r = 0.1
    out["config1_bf16"] = layer_case(8192, 1.0, 0.1, 2, 32, 64, 1, reps=10)
    out["config1_fp32"] = layer_case(8192, 1.0, 0.1, 2, 32, 64, 0, reps=5)
    # BASELINE configs[4] subset: 256 k points at constant density (k ~ 34), F = 2, the four channel pairs
    side = (262144 / 8192) ** (1.0 / 3.0)
    out["config5_256k_f2"] = [layer_case(262144, side, 0.1, 2, c, c, 1, reps=3) for c in (32, 64, 128, 256)]
    out["config5_256k_f1_32"] = layer_case(262144, side, 0.1, 1, 32, 32, 1, reps=3)
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    from se3conv3d_b200 import _lib, shard, workloads as wl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    clocks = ClockSampler(local_rank)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    if args.workload == "sweep":
        # BASELINE configs[4]: single same-level layer, constant density (k ~ 34), every rank its own replica of the cloud
        # (weak scaling, no collective); one JSON line with the whole table, `value` = the 1 M-point F = 2 32 -> 32 case
        if rank == 0:
            clocks.start()
        table = measure_sweep(torch, shard, dev, world, args.precision, flush)
        clk = clocks.stop() if rank == 0 else None
        if rank == 0:
            head = [t for t in table if t["n"] == 1048576 and t["f"] == 2 and t["c_in"] == 32][0]
            print(json.dumps({"metric": "points/sec PNEConvLayerRotEquiv fwd+bwd, single layer sweep (BASELINE configs[4])",
                              "value": head["points_per_s"], "unit": "points/s", "n_gpus": world, "steps": 3, "warmup": 2,
                              "ms_per_step": head["ms_fwd_bwd"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                              "dtype": "f32" if args.precision == 0 else "bf16", "data": "synthetic",
                              "config": {"workload": "layer sweep: N x F x C at constant density (r = 0.1, cube side ~ N^(1/3)), one replica "
                                                     "of every cloud per GPU; value = N 1 M, F 2, 32 -> 32; L2 flushed between iterations",
                                         "precision": args.precision},
                              "clocks": clk, "extra": {"sweep": table}}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    if args.workload == "fpn":
        if rank == 0:
            clocks.start()
        res = measure_fpn(torch, dist, shard, wl, dev, rank, world, args.precision, args.steps, args.warmup, flush, args.strong)
        clk = clocks.stop() if rank == 0 else None
        if rank == 0:
            if "unavailable" in res:
                print(json.dumps({"metric": "points/sec full FPN training step", "unavailable": res["unavailable"]}), flush=True)
            else:
                line = {"metric": res["metric"], "value": res["value"], "unit": "points/s", "n_gpus": world, "steps": args.steps,
                        "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                        "scaling": res["scaling"], "vs_baseline": None, "dtype": "f32" if args.precision == 0 else "bf16",
                        "data": "synthetic",
                        "config": {"workload": "dfaust_I_rot_pca_2F full training step, %d clouds x %d points per GPU, F=2, "
                                               "unmodified reference models/ over this package" % (res["clouds_per_gpu"], N_POINTS),
                                   "precision": args.precision, "l2": "256 MB flush write between timed iterations"},
                        "clocks": clk, "e2e": res["e2e"], "gpu_launches": res["own_kernel_launches_per_step"] * args.steps,
                        "extra": {k: res[k] for k in ("allreduce", "model_parameters", "clouds_per_gpu")}}
                print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- workload: 32 clouds x 6890 points per GPU (weak scaling: clouds are independent, no collective)
    n_clouds = max(N_CLOUDS // world, 1) if args.strong else N_CLOUDS   # strong scaling: the 32 clouds are split over the ranks
    pts_h, batch_h = wl.synthetic_bodies(n_clouds, N_POINTS, seed=shard.shard_seed(rank))
    pts_h, batch_h = pts_h.pin_memory(), batch_h.pin_memory()
    step = wl.DfaustStep(dev, precision=args.precision, seed=0)
    pts_d, batch_d = pts_h.to(dev), batch_h.to(dev)
    pcs, neighs = step.build_hierarchy(pts_d, batch_d, n_batches=n_clouds)
    step.calibrate(pcs, neighs)
    step.make_inputs(pcs)

    def hot_step(p, b):
        pcs_, neighs_ = step.build_hierarchy(p, b, n_batches=n_clouds)
        return step.conv_fwd_bwd(pcs_, neighs_, return_output=True)

    def barrier():
        shard.barrier(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
            step.zero_grad()
        barrier()
        tot = cuda_timed(torch, fn, steps, 0, flush, step.zero_grad)
        barrier()
        return shard.max_over_ranks(tot, dev)

    # rank 0 samples its GPU's clocks / throttle reasons during the timed region (one nvidia-smi poller, not one per
    # rank: eight pollers on one box contend with the launch-bound hierarchy builders for the driver)
    if rank == 0:
        clocks.start()
    launches0 = _lib.launch_count()
    # (1) device-resident: hierarchy + neighbourhoods + conv stack fwd+bwd, every step on its own batch object.
    #     sequential: each step builds its hierarchy, then runs its convolutions, on one stream (per-step events).
    #     pipelined (default): the hierarchy of step i + 1 is built on a second stream while the convolutions of step i
    #     run (K builds and K conv stacks inside ONE timed region; the L2 flush runs inside it, before every conv stack).
    ms_seq = timed(lambda: hot_step(pts_d, batch_d), args.steps, args.warmup)
    launches = (_lib.launch_count() - launches0) // max(args.steps + args.warmup, 1) * args.steps
    side = torch.cuda.Stream(dev, priority=-1)   # the hierarchy builder's tiny kernels go ahead of queued conv kernels

    def timed_pipeline(items, steps, warmup, after=None, drain=None):
        step.run_pipelined(items[:max(warmup, 1)], n_clouds, side, threaded=args.thread)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step.run_pipelined(items[:steps], n_clouds, side, before_conv=lambda: flush.fill_(1), after_conv=after,
                           threaded=args.thread)
        if drain is not None:      # every result copy has landed on the host before the clock stops
            drain()
            torch.cuda.current_stream().wait_stream(copy_stream)
        e1.record()
        e1.synchronize()
        side.synchronize()
        barrier()
        return shard.max_over_ranks(e0.elapsed_time(e1), dev)
    if args.no_pipeline:
        ms_total = ms_seq
    else:
        ms_total = timed_pipeline([(pts_d, batch_d)] * max(args.steps, args.warmup), args.steps, args.warmup)

    # (2) end to end: pinned host buffers -> H2D -> hot path -> D2H of the step's result, the seg-head output
    #     [42 k x 2, 32] fp32, into a pinned host buffer
    y_probe = hot_step(pts_d, batch_d)
    step.zero_grad()
    y_host = torch.empty((int(y_probe.shape[0] * 1.05) + 64, y_probe.shape[1]), dtype=torch.float32).pin_memory()

    def e2e_step():
        p = pts_h.to(dev, non_blocking=True)
        b = batch_h.to(dev, non_blocking=True)
        y = hot_step(p, b)
        y_host[:y.shape[0]].copy_(y, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return y.shape[0]

    if args.no_pipeline:
        ms_e2e = timed(e2e_step, args.steps, max(args.warmup, 3))
    else:
        def h2d():
            return pts_h.to(dev, non_blocking=True), batch_h.to(dev, non_blocking=True)

        y_hosts = [y_host, torch.empty_like(y_host).pin_memory()]
        d2h_ev = [None, None]
        copy_stream = torch.cuda.Stream(dev)

        def d2h(y, i):
            # the step's result goes to pinned host memory (double buffered); the host waits for the copy of step i - 1
            # before it reuses that buffer, i.e. every result is on the host inside the timed region
            # the copy runs on its own stream (copy engine) behind an event, so the next step's kernels do not queue
            # behind 10.8 MB of PCIe traffic on the conv stream
            if d2h_ev[i & 1] is not None:
                d2h_ev[i & 1].synchronize()
            copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy_stream):
                y_hosts[i & 1][:y.shape[0]].copy_(y, non_blocking=True)
                y.record_stream(copy_stream)
                d2h_ev[i & 1] = torch.cuda.Event()
                d2h_ev[i & 1].record(copy_stream)
        ms_e2e = timed_pipeline([h2d] * max(args.steps, args.warmup, 3), args.steps, max(args.warmup, 3), after=d2h,
                                drain=lambda: [e.synchronize() for e in d2h_ev if e is not None])
    d2h_bytes = int(y_probe.shape[0] * y_probe.shape[1] * 4)
    # (3) convolutions only (hierarchy cached), and the dominant layer alone for the roofline
    ms_conv = timed(lambda: step.conv_fwd_bwd(pcs, neighs), args.steps, args.warmup)
    clk = clocks.stop() if rank == 0 else None

    sizes = [(pcs[lo].pts_.shape[0], nb.conv_geometry(pcs[li], pcs[lo]).n_edges, 2, cin, cout)
             for (_, li, lo, _, cin, cout), nb in zip(step.specs, neighs)]
    fwd_b, bwd_b = algorithmic_bytes(sizes)
    fwd_b2, bwd_b2 = algorithmic_bytes(sizes, s=2)
    # per-layer fwd+bwd time (one C-ABI forward + one backward call), L2 flushed between iterations
    per_layer = []
    xs, dys = step.inputs
    for layer, nb, (name, li, lo, _, cin, cout), x, dy, sz in zip(step.layers, neighs, step.specs, xs, dys, sizes):
        def one():
            y = layer(pcs[li], pcs[lo], x, nb)
            y.backward(dy)
        t = timed(one, max(args.steps // 2, 3), 2)
        fb, bb = algorithmic_bytes([sz])
        per_layer.append({"name": name, "ms": t / max(args.steps // 2, 3), "alg_bytes": fb + bb,
                          "m": sz[0], "e": sz[1], "c_in": cin, "c_out": cout})

    # dominant kernel: device time of the three gather kernels over one conv stack (CUDA events on the launch
    # stream inside the library, se3_profile_*), L2 flushed before every stack
    def flushed_stack():
        flush.fill_(1)
        step.conv_fwd_bwd(pcs, neighs)
        step.zero_grad()
    stack_prof = _lib.profile_kernels(flushed_stack, reps=3)
    stack_alg = [0.0, 0.0, 0.0]
    for sz in sizes:
        for i, v in enumerate(kernel_algorithmic_bytes(sz)):
            stack_alg[i] += v
    kern_rows = []
    for i, (kname, ms, cnt) in enumerate(stack_prof):
        per_stack_ms = ms * cnt / 3.0
        kern_rows.append({"kernel": kname, "launches_per_step": cnt // 3, "ms_per_step": per_stack_ms,
                          "alg_gb_per_step": stack_alg[i] / 1e9,
                          "achieved_gbs": stack_alg[i] / max(per_stack_ms * 1e-3, 1e-12) / 1e9})
    dom_k = max(range(3), key=lambda i: kern_rows[i]["ms_per_step"])
    # ... and its largest launch (the layer with the most edges) timed alone, which is also the launch the
    # committed ncu capture under profiles/ measures (dram traffic)
    big = max(range(len(sizes)), key=lambda i: sizes[i][1])
    big_name = step.specs[big][0]

    def flushed_big():
        flush.fill_(1)
        y = step.layers[big](pcs[step.specs[big][1]], pcs[step.specs[big][2]], xs[big], neighs[big])
        y.backward(dys[big])
        step.zero_grad()
    big_prof = _lib.profile_kernels(flushed_big, reps=max(args.steps // 2, 5))
    dom_ms = big_prof[dom_k][1]
    dom_bytes = kernel_algorithmic_bytes(sizes[big])[dom_k]
    dom_bytes2 = kernel_algorithmic_bytes(sizes[big], s=2)[dom_k]
    achieved = dom_bytes / max(dom_ms * 1e-3, 1e-12) / 1e9
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tr.get(["k_agg_tc_fwd", "k_agg_tc_tr", "k_edge_tc"][dom_k], {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    dom = max(per_layer, key=lambda d: d["ms"])

    n_pts_global = n_clouds * N_POINTS * world
    value = n_pts_global * args.steps / (ms_total * 1e-3)
    e2e_value = n_pts_global * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "f32" if args.precision == 0 else "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD if not args.strong else WORKLOAD.replace("32 clouds x 6890 points per GPU",
                                                                               "%d clouds x 6890 points per GPU (32 in total)" % n_clouds),
                   "points_per_gpu": n_clouds * N_POINTS, "precision": args.precision,
                   "l2": "256 MB flush write between timed iterations" if args.no_pipeline else
                         "256 MB flush write before every conv stack, inside the timed region",
                   "pipelined": (not args.no_pipeline),
                   "pipeline": None if args.no_pipeline else
                   "hierarchy of step i+1 built on a second CUDA stream while the convolutions of step i run; K hierarchy "
                   "builds + K conv stacks inside one timed region (CUDA events on the conv stream, both streams drained)",
                   "level_points": [int(p.pts_.shape[0]) for p in pcs], "edges_total": int(sum(s[1] for s in sizes))},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(pts_h.numel() * 4 + batch_h.numel() * 4), "d2h_bytes_per_step": d2h_bytes,
                "d2h": "the seg-head output y [rows, 32] fp32 (the result of the step) into pinned host memory"},
        "gpu_launches": int(launches),
        "breakdown_ms": {"hierarchy_frames_neighbourhoods": (ms_seq - ms_conv) / args.steps,
                         "conv_fwd_bwd_x21": ms_conv / args.steps, "sequential_step": ms_seq / args.steps},
        "conv_only_points_per_s": n_pts_global * args.steps / (ms_conv * 1e-3),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic,
                     "traffic_source": "STATIC: dram__bytes_read+write of this launch from the committed ncu --set full capture "
                                       "(profiles/ncu_traffic.json), not measured in this run",
                     "frac_bf16_rows": dom_bytes2 / max(dom_ms * 1e-3, 1e-12) / 1e9 / peak,
                     "byte_model": "SURVEY 8(d) gather model; `achieved`/`frac` with s = 4 B per feature element (the API's "
                                   "fp32 features, the SURVEY's definition), `frac_bf16_rows` with s = 2 B (the bf16 rows the "
                                   "kernel gathers)",
                     "kernel": "%s, launch of layer %s (M=%d E=%d %d->%d, F=2)" % (
                         kern_rows[dom_k]["kernel"], big_name, sizes[big][0], sizes[big][1], sizes[big][3], sizes[big][4]),
                     "kernel_ms": dom_ms, "alg_bytes_per_launch": dom_bytes, "alg_bytes_per_launch_bf16_rows": dom_bytes2,
                     "peak_source": "MEASURED_PEAKS.json (measured copy bandwidth)" if peaks else "fallback 6650 GB/s",
                     "timing": "CUDA events on the launching stream around the kernel, L2 flushed between iterations",
                     "gather_kernels_per_step": kern_rows,
                     "largest_launch_all_kernels": [
                         {"kernel": k, "ms": ms,
                          "achieved_gbs": kernel_algorithmic_bytes(sizes[big])[i] / max(ms * 1e-3, 1e-12) / 1e9,
                          "achieved_gbs_bf16_rows": kernel_algorithmic_bytes(sizes[big], s=2)[i] / max(ms * 1e-3, 1e-12) / 1e9}
                         for i, (k, ms, _) in enumerate(big_prof)],
                     "layer_fwd_bwd": {"layer": dom["name"], "ms": dom["ms"],
                                       "achieved_gbs": dom["alg_bytes"] / (dom["ms"] * 1e-3) / 1e9},
                     "stack_alg_gb_per_step": (fwd_b + bwd_b) / 1e9,
                     "stack_achieved_gbs": (fwd_b + bwd_b) / (ms_conv / args.steps * 1e-3) / 1e9,
                     "stack_achieved_gbs_bf16_rows": (fwd_b2 + bwd_b2) / (ms_conv / args.steps * 1e-3) / 1e9},
        "per_layer_ms": {d["name"]: round(d["ms"], 4) for d in per_layer},
    }
    extra = {}
    if not args.no_extra:
        # the full FPN training step (with the gradient all-reduce when N > 1) rides along on every run
        try:
            extra["fpn_step"] = measure_fpn(torch, dist, shard, wl, dev, rank, world, args.precision, max(args.steps // 2, 5), 3,
                                            flush)
        except Exception as e:  # the headline line must survive a failure of an extra
            extra["fpn_step"] = {"error": repr(e)[:300]}
        if world == 1:
            try:
                # the fp32 exactness mode on the same conv stack
                step0 = wl.DfaustStep(dev, precision=0, seed=0)
                step0.calibrate(pcs, neighs)
                step0.make_inputs(pcs)
                ms0 = cuda_timed(torch, lambda: step0.conv_fwd_bwd(pcs, neighs), 3, 1, flush, step0.zero_grad) / 3
                extra["precision0_conv_stack"] = {"ms": ms0, "points_per_s": n_clouds * N_POINTS / (ms0 * 1e-3)}
                del step0
                extra.update(measure_extras(torch, wl, dev, args.precision, flush))
            except Exception as e:
                extra["error"] = repr(e)[:300]
    line["extra"] = extra
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # CPU port of the same hot path on a bounded sample (one cloud), all host threads
        from oracle import hierarchy_oracle as ho
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        p1, b1 = wl.synthetic_bodies(1, N_POINTS, seed=0)
        params = [(l.proj_axes_.detach().cpu(), l.proj_biases_.detach().cpu(), l.conv_weights_.detach().cpu())
                  for l in step.layers]
        ho.dfaust_step_cpu(p1.numpy(), b1.numpy(), step.specs, params, wl.DFAUST_CFG)
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 20):
            ho.dfaust_step_cpu(p1.numpy(), b1.numpy(), step.specs, params, wl.DFAUST_CFG)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        line["cpu_baseline"] = {"value": N_POINTS / dt, "unit": "points/s", "cores": cores, "kind": "port",
                                "sample": "1 of 32 clouds (6890 points) per step, oracle port (torch CPU + C), %d reps" % reps}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

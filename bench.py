#!/usr/bin/env python
"""Benchmark of the GASFM graph-attention hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one forward+backward of the full GASFM model over one synthetic scene.  At N=1 the
workload is BASELINE.json configs[1]: 300 views x 50k points, ~500k observations, n_feat_proj=256,
4 heads, 12 layers (other widths as shipped).  Prints ONE JSON line (see the task contract):
  value  GAT-layer edges/s fwd+bwd with the scene resident in HBM  (E x 26 edge-level GATs / step)
  e2e    same metric through the public API with the scene in pinned HOST memory: H2D copy of the
         observations, index build, forward+backward, D2H of the predictions -- every step
  roofline      the dominant edge-attention kernel timed alone with CUDA events (achieved GB/s of
                algorithmic bytes over the measured HBM peak)
  cpu_baseline  the CPU oracle (port of the reference's PyTorch path) on the host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = dict(name="cfg2", m=300, n=50_000, n_obs=500_000, n_feat_proj=256, num_layers=12, seed=0)
# BASELINE.json configs[2] (1,000 views x 300k points, ~5M observations) at the shipped width and at d=256
CFG3 = dict(name="cfg3", m=1000, n=300_000, n_obs=5_000_000, n_feat_proj=32, num_layers=12, seed=0)
CFG3_WIDE = dict(CFG3, name="cfg3_d256", n_feat_proj=256)
WORKLOADS = {"cfg2": CFG2, "cfg3": CFG3, "cfg3_d256": CFG3_WIDE}
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": FALLBACK_HBM_GBS}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(cfg):
    from gasfm_b200.config import gasfm_conf
    from gasfm_b200.models.graph_attn_sfm import GraphAttnSfMNet
    from gasfm_b200.scene import Scene
    from oracle import gasfm_cpu  # synthetic scene generator only (shared with the tests)

    idx, vals = gasfm_cpu.synthetic_observations(cfg["m"], cfg["n"], cfg["n_obs"], cfg["seed"])
    scene = Scene.from_observations(idx, vals, cfg["m"], cfg["n"])
    conf = gasfm_conf(n_feat_proj=cfg["n_feat_proj"], num_layers=cfg["num_layers"])
    torch.manual_seed(cfg["seed"])
    model = GraphAttnSfMNet(conf)
    return conf, model, scene


def surrogate_loss(out):
    """Scalar touching both predictions (the reprojection loss is outside the hot path)."""
    return out["Ps_norm"].square().mean() + out["pts3D"].square().mean()


def step_device(model, scene):
    model.zero_grad(set_to_none=True)
    out = model(scene)
    loss = surrogate_loss(out)
    loss.backward()
    return loss


def timed(fn, steps, warmup, sync_dist=False):
    for _ in range(warmup):
        fn()
    if sync_dist:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        fn()
    end.record()
    if sync_dist:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / steps
    if sync_dist:
        t = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def timed_batches(fn, batches=5, per_batch=5, warmup=5):
    """Average launch duration over batches x per_batch launches (CUDA events).  A batch that is more
    than 10x the median batch (a one-off stall of a fresh process, seen once on this pool) is dropped."""
    ms = [timed(fn, per_batch, warmup if b == 0 else 0) for b in range(batches)]
    med = float(np.median(ms))
    kept = [t for t in ms if t <= 10 * med]
    return float(np.mean(kept))


def kernel_roofline(cfg, peaks, peak_kind):
    """The hot kernels alone at the workload's shapes (E observations, width d = n_feat_proj), CUDA events.

    Edge-attention kernels: HBM-bound; achieved = algorithmic bytes / time over the
    measured copy bandwidth.  Projection GEMMs (fwd + dX): the scaled 2 x FP16 split (three kind::f16 MMAs per
    product at the bf16 rate) leaves them HBM-bound: achieved = (read A + write C) / time; the tensor share is
    listed as ``tensor_frac``; the fp16 weight gradient likewise.  Concatenated input gradient (and GASFM_GEMM=tf32x3):
    3xTF32, tensor-bound: achieved =
    3 * 2MNK / time over the tf32 rate, taken as half of the measured dense bf16 rate; HBM fraction listed too.
    Every E-sized operand (E x 256 fp32 = 507 MB) exceeds the 126 MB L2, so launches are cold.
    ``calls_per_step`` x time picks the dominant kernel of the step."""
    from gasfm_b200 import ops
    from gasfm_b200.index import ObservationIndex
    from oracle import gasfm_cpu

    peak_gbs = float(peaks["hbm_gbs"])
    peak_tf32 = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
    dev = torch.device("cuda")
    H, HC, L = 4, cfg["n_feat_proj"], cfg["num_layers"]
    idx, _ = gasfm_cpu.synthetic_observations(cfg["m"], cfg["n"], cfg["n_obs"], cfg["seed"])
    E = idx.shape[1]
    oi = ObservationIndex(torch.from_numpy(idx).to(dev), cfg["m"], cfg["n"])
    torch.manual_seed(0)
    XL = torch.randn(E, HC, device=dev)
    att = torch.randn(1, H, HC // H, device=dev) * 0.2
    res = {}
    for name, plan, T in (("tracks", oi.by_track, cfg["n"]), ("views", oi.by_view, cfg["m"])):
        XR = torch.randn(T, HC, device=dev)
        out, mx, sm = ops.gat_edge_partial(XL, XR, att, plan, H)
        out = out / sm.repeat_interleave(HC // H, dim=1).clamp_min(1e-30)
        dO = torch.randn(T, HC, device=dev)
        fwd_ms = timed_batches(lambda: ops.gat_edge_partial(XL, XR, att, plan, H))
        bwd_ms = timed_batches(lambda: ops.gat_edge_backward_raw(XL, XR, att, out, mx, sm, dO, plan, H))
        fwd_bytes = E * (HC * 4 + 4) + T * (2 * HC * 4 + 8 * H)
        bwd_bytes = E * (2 * HC * 4 + 4) + T * (4 * HC * 4 + 8 * H)
        res[f"gat_fwd_{name}"] = dict(bound="hbm", ms=fwd_ms, work=fwd_bytes, calls=L + 1)
        res[f"gat_bwd_{name}"] = dict(bound="hbm", ms=bwd_ms, work=bwd_bytes, calls=L + 1)
        del XR, out, dO
    if ops.gemm_tf32x3_supported(E, HC, HC, HC, HC):
        W = torch.randn(HC, HC, device=dev) / HC ** 0.5
        n_blocks3 = L - 1                             # stateful blocks: lin_l x2 + lin_proj of the same x
        n_gemm = 3 * n_blocks3 + 2                    # + lin_l x2 in the final update
        flops = 3 * 2.0 * E * HC * HC                 # three split products
        io_bytes = 2 * E * HC * 4                     # read A once, write C once (weights stay in L2)
        f16 = ops.GEMM_KIND == "f16x2" and ops.gemm_f16x2_supported(E, HC, HC, HC, HC)
        gemm_ms = timed_batches(lambda: ops.gemm_tc(XL, W))
        if f16:
            # scaled 2 x FP16 split: the tensor time is half the 3xTF32 kernel's, the kernel is HBM-bound.
            # Forward: the three projections of a block in one launch (A read once, three outputs written).
            grp_ms = timed_batches(lambda: ops.gemm_f16x2_groups(XL, [W, W, W], [None, None, None]))
            res["gemm_f16x2 x3 groups (forward projections)"] = dict(
                bound="hbm", ms=grp_ms, work=4 * E * HC * 4, calls=n_blocks3, tensor_frac=3 * flops / grp_ms / 1e9 / (2.0 * peak_tf32))
            res["gemm_f16x2 (single projection)"] = dict(bound="hbm", ms=gemm_ms, work=io_bytes, calls=2,
                                                         tensor_frac=flops / gemm_ms / 1e9 / (2.0 * peak_tf32))
        else:
            res["gemm_tf32x3 (forward projections)"] = dict(bound="tensor", ms=gemm_ms, work=flops, calls=n_gemm, hbm_bytes=io_bytes)
        # input gradient of a block's three projections: one 3xTF32 GEMM over [dY0 | dY1 | dY2] (K = 3 HC)
        Wcat = torch.cat([W, W, W], dim=1)
        dx_ms = timed_batches(lambda: ops.gemm_tf32x3_cat([XL, XL, XL], Wcat))
        res["gemm_tf32x3_cat (dX over 3 dY)"] = dict(bound="tensor", ms=dx_ms, work=3 * flops, calls=n_blocks3, hbm_bytes=4 * E * HC * 4)
        if f16 and ops.WGRAD_KIND == "f16x2" and ops.wgrad_f16x2_supported(E, HC, HC, HC, HC):
            # fp16 weight gradient (operand maxima come from the GEMMs above): HBM-bound like the forward projections
            amax = XL.abs().max().reshape(1)
            wg_ms = timed_batches(lambda: ops.wgrad_f16x2(XL, XL, amax, amax))
            res["wgrad_f16x2 (dW)"] = dict(bound="hbm", ms=wg_ms, work=io_bytes, calls=n_gemm,
                                           tensor_frac=flops / wg_ms / 1e9 / (2.0 * peak_tf32))
        else:
            wg_ms = timed_batches(lambda: ops.wgrad_tf32x3(XL, XL))
            res["wgrad_tf32x3 (dW)"] = dict(bound="tensor", ms=wg_ms, work=flops, calls=n_gemm, hbm_bytes=io_bytes)
    for v in res.values():
        if v["bound"] == "hbm":
            v["achieved"], v["peak"], v["unit"] = v["work"] / v["ms"] / 1e6, peak_gbs, "GB/s"
        else:
            v["achieved"], v["peak"], v["unit"] = v["work"] / v["ms"] / 1e9, peak_tf32, "TFLOP/s"
            v["hbm_frac"] = round(v["hbm_bytes"] / v["ms"] / 1e6 / peak_gbs, 4)
        v["frac"] = v["achieved"] / v["peak"]
    dom = max(res, key=lambda k: res[k]["ms"] * res[k]["calls"])
    d = res[dom]
    roof = {"bound": d["bound"], "kernel": dom, "achieved": round(d["achieved"], 1), "peak": round(d["peak"], 1),
            "peak_source": peak_kind + (" (hbm_gbs)" if d["bound"] == "hbm" else " (bf16_tflops / 2 = tf32 MMA rate)"),
            "unit": d["unit"], "frac": round(d["frac"], 4), "traffic": None,
            "algorithmic_work_per_launch": d["work"], "ms_per_launch": round(d["ms"], 4),
            "all_kernels": {k: {"bound": v["bound"], "ms": round(v["ms"], 4), "calls_per_step": v["calls"],
                                "achieved": round(v["achieved"], 1), "unit": v["unit"], "frac": round(v["frac"], 4),
                                **({"hbm_frac": v["hbm_frac"]} if "hbm_frac" in v else {}),
                                **({"tensor_frac": round(v["tensor_frac"], 4)} if "tensor_frac" in v else {})}
                            for k, v in res.items()}}
    return roof


def cpu_baseline(cfg, layers=1, repeats=1):
    """The oracle (= port of the reference's PyTorch path, materialising what the reference
    materialises) forward+backward on the host cores.  Bounded sample: the same scene and widths
    but ``layers`` GASFM blocks instead of 12 (the reference's memory use grows by ~10 GB per
    block at this size); edges/s counts the 2*(layers+1) edge-level GATs run."""
    from gasfm_b200.config import gasfm_conf
    from gasfm_b200.models.graph_attn_sfm import GraphAttnSfMNet
    from oracle import gasfm_cpu

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    idx, vals = gasfm_cpu.synthetic_observations(cfg["m"], cfg["n"], cfg["n_obs"], cfg["seed"])
    scene = gasfm_cpu.scene_from_sparse(torch.from_numpy(idx), torch.from_numpy(vals), cfg["m"], cfg["n"])
    torch.manual_seed(cfg["seed"])
    model = GraphAttnSfMNet(gasfm_conf(n_feat_proj=cfg["n_feat_proj"], num_layers=layers))
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    E = idx.shape[1]
    best = float("inf")
    for _ in range(repeats):
        for p in params.values():
            p.grad = None
        t0 = time.perf_counter()
        out = gasfm_cpu.gasfm_forward(params, scene)
        surrogate_loss(out).backward()
        best = min(best, time.perf_counter() - t0)
    n_gat = 2 * (layers + 1)
    return {"value": E * n_gat / best, "unit": "edges/s", "cores": cores, "kind": "port",
            "sample": f"same scene (E={E}), same widths, {layers} of {cfg['num_layers']} blocks "
                      f"({n_gat} edge-level GATs), fwd+bwd, best of {repeats}, {best:.2f} s"}, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = dict(CFG2)
    times, base = [], None
    for i in range(args.warmup + args.steps):
        base, t = cpu_baseline(cfg, layers=1, repeats=1)
        if i >= args.warmup:
            times.append(t)
    E = int(base["sample"].split("E=")[1].split(")")[0])
    ms = 1e3 * float(np.mean(times))
    value = E * 4 / (ms / 1e3)
    base["value"] = value
    line = {"impl": "reference", "metric": "gat_layer_edges_per_sec_fwd_bwd", "value": value, "unit": "edges/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, E, 1), "cpu_baseline": base,
            "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(cfg, E, n_gpus):
    return {"workload": f"{cfg['name']}: GASFM fwd+bwd, {cfg['m']} views x {cfg['n']} points, E={E} observations, "
                        f"n_feat_proj={cfg['n_feat_proj']}, 4 heads, {cfg['num_layers']} layers, shipped other widths",
            "edge_level_gats_per_step": 2 * (cfg["num_layers"] + 1),
            "cache": f"inputs larger than L2 (one [E,{cfg['n_feat_proj']}] fp32 tensor = {E * cfg['n_feat_proj'] * 4 / 1e6:.0f} MB vs 126 MB L2)",
            "parallelism": "single GPU" if n_gpus == 1 else f"tracks sharded over {n_gpus} GPUs"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="default cfg2 = the configuration the headline metric is quoted on")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = one scene of N x the workload's tracks; strong = the workload's scene itself")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the device-resident step eagerly instead of as a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    from gasfm_b200 import _lib
    _lib.load()   # fail loudly if the CUDA library is missing
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        from gasfm_b200 import dist as gdist
        return gdist.bench_main(args, WORKLOADS[args.workload], workload_config, ClockSampler, measured_peaks, timed,
                                surrogate_loss)
    dev = torch.device("cuda", local_rank)
    cfg = dict(WORKLOADS[args.workload])
    peaks, peak_kind = measured_peaks()

    conf, model, scene_host = build_workload(cfg)
    model = model.to(dev)
    scene_host.pin_memory()
    E = scene_host.x.indices.shape[1]
    n_gat = 2 * (cfg["num_layers"] + 1)

    # ---- device-resident throughput ------------------------------------------------------------
    scene_dev = scene_host.to(dev)
    step_device(model, scene_dev)           # builds and caches the CSR/CSC index
    torch.cuda.synchronize()
    launches0 = _lib.launch_count
    eager_ms = timed(lambda: step_device(model, scene_dev), max(2, args.steps // 2), args.warmup)
    launches = (_lib.launch_count - launches0) // (max(2, args.steps // 2) + args.warmup)
    # the device-resident step is replayed as ONE CUDA graph (same kernels, no per-launch host overhead)
    step_fn, graphed = (lambda: step_device(model, scene_dev)), False
    if not args.no_graph:
        try:
            from gasfm_b200.graphs import GraphedStep
            gstep = GraphedStep(model, scene_dev, surrogate_loss)
            step_fn, graphed = gstep, True
        except Exception as exc:  # capture is an optimisation; report and fall back to eager timing
            print(f"[bench] CUDA graph capture failed, timing eagerly: {exc}", file=sys.stderr)
            torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(step_fn, args.steps, args.warmup)
    clocks = sampler.stop()
    value = E * n_gat / (ms / 1e3)

    # ---- forward only (ms per scene) -----------------------------------------------------------
    def fwd_only():
        with torch.no_grad():
            model(scene_dev)
    fwd_ms = timed(fwd_only, max(3, args.steps), 2)

    # ---- end to end: host scene -> device -> fwd+bwd -> predictions back on the host -----------
    h2d = (scene_host.x.values.numel() * 4 + scene_host.x.indices.numel() * 8 +
           scene_host.x.cam_per_pts.numel() * 8 + scene_host.x.pts_per_cam.numel() * 8 +
           sum(w.valid_indices.numel() * 8 for k, w in scene_host.graph_wrappers.items() if k.endswith("2global")))
    d2h_holder = {}

    def step_e2e():
        s = scene_host.to(dev, non_blocking=True)
        model.zero_grad(set_to_none=True)
        out = model(s)
        loss = surrogate_loss(out)
        loss.backward()
        d2h_holder["Ps"] = out["Ps_norm"].detach().cpu()
        d2h_holder["pts"] = out["pts3D"].detach().cpu()
        d2h_holder["loss"] = float(loss)
    e2e_ms = timed(step_e2e, args.steps, args.warmup)
    d2h = d2h_holder["Ps"].numel() * 4 + d2h_holder["pts"].numel() * 4 + 4
    e2e_value = E * n_gat / (e2e_ms / 1e3)

    roofline = None if args.no_roofline else kernel_roofline(cfg, peaks, peak_kind)
    line = {"metric": "gat_layer_edges_per_sec_fwd_bwd", "value": value, "unit": "edges/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, E, 1), "forward_ms_per_scene": fwd_ms,
            "cuda_graph": graphed, "eager_ms_per_step": eager_ms,
            "e2e": {"value": e2e_value, "unit": "edges/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
            "clocks": clocks, "roofline": roofline}
    if not args.no_cpu_baseline and rank == 0:
        del scene_dev
        torch.cuda.empty_cache()
        line["cpu_baseline"], _ = cpu_baseline(cfg, layers=1, repeats=1)
    print(json.dumps(line))


if __name__ == "__main__":
    main()

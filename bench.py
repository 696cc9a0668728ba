#!/usr/bin/env python
"""Benchmark of the GASFM graph-attention hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3_d256|cfg3|cfg2]
    torchrun --nproc-per-node N bench.py --gpus N ...        (N > 1: one rank per GPU)

A "step" is one forward+backward of the full GASFM model over one synthetic scene.  The default workload is the
configuration BASELINE.json's target sentence names -- configs[2]: 1,000 views x 300k points, ~5M observations -- at
n_feat_proj=256, 4 heads, 12 layers (other widths as shipped).  With N > 1 the SAME scene is track-sharded over the
GPUs (strong scaling; --scaling weak grows the scene with N instead).  Prints ONE JSON line (see the task contract):
  value   GAT-layer edges/s fwd+bwd with the scene resident in HBM (E x 26 edge-level GATs / step), CUDA-graph replay
  e2e     same metric through the public API with the scene in pinned HOST memory: H2D copy of the observations,
          index build, forward+backward, D2H of the predictions -- every step
  roofline       the dominant hot kernel timed alone with CUDA events (algorithmic bytes / time over the measured HBM
                 peak), the other hot kernels under ``all_kernels``, and the layer-level floor under ``layer``
  cpu_baseline   the CPU oracle (port of the reference's PyTorch path) on the host cores, bounded sample   (N = 1)
  parity         N = 1: GPU vs that CPU run on the same sample scene and weights;  N > 1: the track-sharded model vs
                 the single-GPU model on a small scene (outputs and every parameter gradient)
  cfg2           N = 1: the same measurements on configs[1] (300 x 50k, ~500k observations), the round-1 headline
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
# BASELINE.json configs[0] as a TRAINING sample: the shipped model on a 20-view x 2k-point scene (~30 % density) -- the size
# the reference's multi-scene training actually draws (code/datasets/ScenesDataSet.py:30-39); launch-bound without graph replay
CFG1 = dict(name="cfg1", m=20, n=2000, n_obs=12_000, n_feat_proj=32, num_layers=12, seed=0)
WORKLOADS = {"cfg1": CFG1, "cfg2": CFG2, "cfg3": CFG3, "cfg3_d256": CFG3_WIDE}
DEFAULT_WORKLOAD = "cfg3_d256"
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback
METRIC = "gat_layer_edges_per_sec_fwd_bwd"


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


_SCENES = {}


def observations(m, n, n_obs, seed):
    from gasfm_b200.synthetic import synthetic_observations

    key = (m, n, n_obs, seed)
    if key not in _SCENES:
        _SCENES[key] = synthetic_observations(m, n, n_obs, seed)
    return _SCENES[key]


def make_model(cfg, layers=None):
    from gasfm_b200.config import gasfm_conf
    from gasfm_b200.models.graph_attn_sfm import GraphAttnSfMNet

    torch.manual_seed(cfg["seed"])
    return GraphAttnSfMNet(gasfm_conf(n_feat_proj=cfg["n_feat_proj"], num_layers=layers or cfg["num_layers"]))


def build_workload(cfg):
    """(conf, model, host scene) of a single-GPU workload (also used by tools/profile_step.py, tools/one_step.py)."""
    from gasfm_b200.scene import Scene

    idx, vals = observations(cfg["m"], cfg["n"], cfg["n_obs"], cfg["seed"])
    model = make_model(cfg)
    return None, model, Scene.from_observations(idx, vals, cfg["m"], cfg["n"])


def step_device(model, scene):
    model.zero_grad(set_to_none=True)
    out = model(scene)
    loss = out["Ps_norm"].square().mean() + out["pts3D"].square().mean()
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


# ---------------------------------------------------------------------------------------------
# roofline of the hot kernels, each timed alone
# ---------------------------------------------------------------------------------------------
def kernel_roofline(cfg, peaks, peak_kind, forward_ms=None, recompute=False):
    """The hot kernels alone at the workload's shapes (E observations, width d = n_feat_proj), CUDA events.

    Edge-attention kernels, the scaled 2 x FP16 projections and the fp16 weight gradient are HBM-bound: achieved =
    algorithmic bytes (GEMMs: read A once + write C once) / time over the measured copy bandwidth.  The
    concatenated input gradient (3xTF32) is tensor-bound: achieved = ALGORITHMIC 2 M N K / time over the tf32 MMA rate
    (half the measured dense bf16 rate); the three split products it actually issues are listed as ``pipe_util``.
    Every E-sized operand exceeds the 126 MB L2, so launches are cold.  ``calls_per_step`` x time picks the dominant one."""
    from gasfm_b200 import ops
    from gasfm_b200.index import ObservationIndex

    peak_gbs = float(peaks["hbm_gbs"])
    peak_tf32 = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
    dev = torch.device("cuda")
    H, HC, L = 4, cfg["n_feat_proj"], cfg["num_layers"]
    idx, _ = observations(cfg["m"], cfg["n"], cfg["n_obs"], cfg["seed"])
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
        flops = 2.0 * E * HC * HC                     # algorithmic flops of one projection
        io_bytes = 2 * E * HC * 4                     # read A once, write C once (weights stay in L2)
        f16 = ops.GEMM_KIND == "f16x2" and ops.gemm_f16x2_supported(E, HC, HC, HC, HC)
        gemm_ms = timed_batches(lambda: ops.gemm_tc(XL, W))
        if f16:
            ln_fused = ops.LN_FUSED and ops.gemm_f16x2_ln_y_supported(E, HC, HC, HC, 3 * HC)
            if ln_fused:
                # as the model runs it: LayerNorm + ReLU inside the operand producer (x_raw in, three projections out)
                gamma, beta = torch.rand(HC, device=dev) + 0.5, torch.randn(HC, device=dev) * 0.1
                grp_ms = timed_batches(lambda: ops.gemm_f16x2_groups_ln(XL, gamma, beta, 1e-5, [W, W, W], [None, None, None]))
            else:
                grp_ms = timed_batches(lambda: ops.gemm_f16x2_groups(XL, [W, W, W], [None, None, None]))
            # write_frac: the bytes this kernel WRITES against the write half of the measured copy bandwidth (a copy moves
            # hbm_gbs / 2 in each direction)
            res["gemm_f16x2 x3 groups (LayerNorm + ReLU + forward projections)" if ln_fused else "gemm_f16x2 x3 groups (forward projections)"] = dict(
                bound="hbm", ms=grp_ms, work=4 * E * HC * 4, calls=n_blocks3, pipe_util=9 * flops / grp_ms / 1e9 / (2.0 * peak_tf32),
                write_frac=3 * E * HC * 4 / grp_ms / 1e6 / (peak_gbs / 2))
            if recompute and ln_fused:
                # activation recompute: backward rebuilds the two attention sources and relu(LN(x)) in one more pass over x_raw
                rc_ms = timed_batches(lambda: ops.gemm_f16x2_groups_ln(XL, gamma, beta, 1e-5, [W, W], [None, None], want_y=True))
                res["gemm_f16x2 x2 groups + relu(LN(x)) (recompute in backward)"] = dict(
                    bound="hbm", ms=rc_ms, work=4 * E * HC * 4, calls=n_blocks3, pipe_util=6 * flops / rc_ms / 1e9 / (2.0 * peak_tf32),
                    write_frac=3 * E * HC * 4 / rc_ms / 1e6 / (peak_gbs / 2))
            res["gemm_f16x2 (single projection)"] = dict(bound="hbm", ms=gemm_ms, work=io_bytes, calls=2,
                                                         pipe_util=3 * flops / gemm_ms / 1e9 / (2.0 * peak_tf32),
                                                         write_frac=E * HC * 4 / gemm_ms / 1e6 / (peak_gbs / 2))
        else:
            res["gemm_tf32x3 (forward projections)"] = dict(bound="tensor", ms=gemm_ms, work=flops, calls=n_gemm, hbm_bytes=io_bytes,
                                                            pipe_util=3 * flops / gemm_ms / 1e9 / peak_tf32)
        # input gradient of a block's three projections: one 3xTF32 GEMM over [dY0 | dY1 | dY2] (K = 3 HC)
        Wcat = torch.cat([W, W, W], dim=1)
        if ops.DX_KIND == "f16x2" and ops.gemm_f16x2_cat_supported(E, HC, 3, HC):
            # fp16 path: half the MMA time of 3xTF32, which leaves the kernel between the two roofs; credited as HBM-bound
            # (read the three dY once + write dX once)
            # (one-pass form: the row maxima come from the kernels that produced the dY_i, as in the model's backward)
            rowmax = XL.abs().amax(dim=1)
            dx_ms = timed_batches(lambda: ops.gemm_f16x2_cat([XL, XL, XL], Wcat, rowmax=[rowmax, rowmax, rowmax]))
            res["gemm_f16x2_cat (dX over 3 dY, row maxima from upstream)"] = dict(
                bound="hbm", ms=dx_ms, work=4 * E * HC * 4, calls=n_blocks3, pipe_util=9 * flops / dx_ms / 1e9 / (2.0 * peak_tf32))
        else:
            dx_ms = timed_batches(lambda: ops.gemm_tf32x3_cat([XL, XL, XL], Wcat))
            res["gemm_tf32x3_cat (dX over 3 dY)"] = dict(bound="tensor", ms=dx_ms, work=3 * flops, calls=n_blocks3, hbm_bytes=4 * E * HC * 4,
                                                         pipe_util=9 * flops / dx_ms / 1e9 / peak_tf32)
        if f16 and ops.WGRAD_KIND == "f16x2" and ops.wgrad_f16x2_supported(E, HC, HC, HC, HC):
            amax = XL.abs().max().reshape(1)
            wg_ms = timed_batches(lambda: ops.wgrad_f16x2(XL, XL, amax, amax))
            res["wgrad_f16x2 (dW, single)"] = dict(bound="hbm", ms=wg_ms, work=io_bytes, calls=2,
                                                   pipe_util=3 * flops / wg_ms / 1e9 / (2.0 * peak_tf32))
            if ops.WGRAD_MULTI:
                # the three weight gradients of a block in one launch: algorithmic bytes = three dY + x ONCE
                amax3 = amax.repeat(3)
                wg3_ms = timed_batches(lambda: ops.wgrad_f16x2_multi([XL, XL, XL], XL, amax3, amax))
                res["wgrad_f16x2_multi (3 dW of a block)"] = dict(bound="hbm", ms=wg3_ms, work=4 * E * HC * 4, calls=n_blocks3,
                                                                  pipe_util=9 * flops / wg3_ms / 1e9 / (2.0 * peak_tf32))
        else:
            wg_ms = timed_batches(lambda: ops.wgrad_tf32x3(XL, XL))
            res["wgrad_tf32x3 (dW)"] = dict(bound="tensor", ms=wg_ms, work=flops, calls=n_gemm, hbm_bytes=io_bytes,
                                            pipe_util=3 * flops / wg_ms / 1e9 / peak_tf32)
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
                                **({"pipe_util": round(v["pipe_util"], 4)} if "pipe_util" in v else {}),
                                **({"write_frac": round(v["write_frac"], 4)} if "write_frac" in v else {})}
                            for k, v in res.items()}}
    try:   # DRAM traffic of the dominant kernel from the committed ncu capture of this shape, when there is one
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f).get(cfg["name"], {}).get(dom)
        if t:
            roof["traffic"], roof["traffic_source"] = t["bytes"], t["source"]
    except Exception:
        pass
    if forward_ms:
        # fused-layer floor of SURVEY.md 8d: x read once, x_out written once, the per-edge index once -- per block
        floor_bytes = E * ((HC + HC) * 4 + 8) * L
        floor_ms = floor_bytes / peak_gbs / 1e6
        roof["layer"] = {"floor_bytes_forward": floor_bytes, "floor_ms_forward": round(floor_ms, 3),
                         "forward_ms_per_scene": round(forward_ms, 3), "frac": round(floor_ms / forward_ms, 4)}
    return roof


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's PyTorch path) on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------
CPU_SAMPLE_LAYERS = 2          # blocks 0 (width 2 -> d) and 1 (d -> d) + the final update: 6 edge-level GATs, 4 at full width
CPU_SAMPLE_OBS = 500_000       # the reference's materialise-everything style needs ~10 GB per full-width block at this size


def cpu_sample_cfg(cfg):
    """Bounded sample of a workload for the CPU arm: the same views, density, widths and weights-init seed, but
    ``CPU_SAMPLE_OBS`` observations (a proportional subset of the tracks) and ``CPU_SAMPLE_LAYERS`` of the blocks."""
    frac = min(1.0, CPU_SAMPLE_OBS / cfg["n_obs"])
    return dict(cfg, n=max(1000, int(round(cfg["n"] * frac))), n_obs=int(round(cfg["n_obs"] * frac)), num_layers=CPU_SAMPLE_LAYERS)


def sample_loss(out):
    return out["Ps_norm"].square().mean() + out["pts3D"].square().mean()


def cpu_baseline(cfg, repeats=1):
    """-> (cpu_baseline dict, seconds, reference outputs / gradients for the parity check)."""
    from oracle import gasfm_cpu

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    s = cpu_sample_cfg(cfg)
    idx, vals = observations(s["m"], s["n"], s["n_obs"], s["seed"])
    scene = gasfm_cpu.scene_from_sparse(torch.from_numpy(idx), torch.from_numpy(vals), s["m"], s["n"])
    model = make_model(s)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    E = idx.shape[1]
    best, out = float("inf"), None
    for _ in range(repeats):
        for p in params.values():
            p.grad = None
        t0 = time.perf_counter()
        out = gasfm_cpu.gasfm_forward(params, scene)
        sample_loss(out).backward()
        best = min(best, time.perf_counter() - t0)
    n_gat = 2 * (s["num_layers"] + 1)
    ref = {"out": {k: out[k].detach() for k in ("Ps_norm", "pts3D")},
           "grads": {k: p.grad.detach() for k, p in params.items() if p.grad is not None}}
    info = {"value": E * n_gat / best, "unit": "edges/s", "cores": cores, "kind": "port",
            "sample": f"{s['m']} views x {s['n']} points (E={E}: the workload's density on a subset of its tracks), same widths, "
                      f"{s['num_layers']} of {cfg['num_layers']} blocks ({n_gat} edge-level GATs, {n_gat - 2} at full width), "
                      f"fwd+bwd, best of {repeats}, {best:.2f} s"}
    return info, best, ref


def _grad_error(got, want):
    """Worst per-parameter gradient error, each scaled by max(|g_param|, 1e-3 * largest gradient) (tests/conftest.py)."""
    gscale = 1e-3 * max(float(v.abs().max()) for v in want.values())
    worst, key = 0.0, None
    for k, w in want.items():
        err = float((got[k] - w).abs().max()) / max(gscale, float(w.abs().max()))
        if err > worst:
            worst, key = err, k
    return worst, key


def parity_vs_cpu(cfg, ref, dev):
    """The GPU model on the CPU arm's sample scene with the same weights: outputs and all gradients vs the oracle."""
    from gasfm_b200.scene import Scene

    s = cpu_sample_cfg(cfg)
    idx, vals = observations(s["m"], s["n"], s["n_obs"], s["seed"])
    model = make_model(s).to(dev)
    out = model(Scene.from_observations(idx, vals, s["m"], s["n"]).to(dev))
    sample_loss(out).backward()
    errs = {k: float((out[k].detach().cpu() - ref["out"][k]).abs().max() / max(1.0, float(ref["out"][k].abs().max())))
            for k in ("Ps_norm", "pts3D")}
    got = {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}
    worst, key = _grad_error(got, ref["grads"])
    return {"against": "CPU oracle (fp32) on the cpu_baseline sample scene, same weights", "E": int(idx.shape[1]),
            "max_err_outputs": errs, "tolerance_outputs": 1e-4, "worst_grad_err": worst, "worst_grad_param": key,
            "tolerance_grads": 2e-3, "ok": bool(max(errs.values()) < 1e-4 and worst < 2e-3)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = dict(WORKLOADS[args.workload])
    times, base = [], None
    for i in range(args.warmup + args.steps):
        base, t, _ = cpu_baseline(cfg, repeats=1)
        if i >= args.warmup:
            times.append(t)
    s = cpu_sample_cfg(cfg)
    E = observations(s["m"], s["n"], s["n_obs"], s["seed"])[0].shape[1]
    ms = 1e3 * float(np.mean(times))
    value = E * 2 * (s["num_layers"] + 1) / (ms / 1e3)
    base["value"] = value
    # the SAME config object our arm prints for this --gpus / --workload / --scaling (the CPU arm itself is one process)
    weak = args.gpus > 1 and args.scaling == "weak"
    n_total, obs_total = (cfg["n"] * args.gpus, cfg["n_obs"] * args.gpus) if weak else (cfg["n"], cfg["n_obs"])
    E_full = observations(cfg["m"], n_total, obs_total, cfg["seed"])[0].shape[1]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "edges/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": args.scaling if args.gpus > 1 else "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(dict(cfg, n=n_total), E_full, args.gpus, args.scaling if args.gpus > 1 else "strong"),
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(cfg, E, n_gpus, scaling):
    return {"workload": f"{cfg['name']}: GASFM fwd+bwd, {cfg['m']} views x {cfg['n']} points, E={E} observations, "
                        f"n_feat_proj={cfg['n_feat_proj']}, 4 heads, {cfg['num_layers']} layers, shipped other widths",
            "edge_level_gats_per_step": 2 * (cfg["num_layers"] + 1),
            "cache": f"inputs larger than L2 (one [E,{cfg['n_feat_proj']}] fp32 tensor = {E * cfg['n_feat_proj'] * 4 / 1e6 / n_gpus:.0f} MB per GPU vs 126 MB L2)",
            "parallelism": "single GPU" if n_gpus == 1 else
                           f"tracks sharded over {n_gpus} GPUs ({scaling} scaling), per-view softmax partials merged by peer-memory kernels over NVLink"}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def measure_workload(cfg, args, dev, rank, world, exchange, scaling):
    """Device-resident step (graph replay), forward only, and the end-to-end step of one workload."""
    from gasfm_b200 import _lib
    from gasfm_b200 import dist as gdist
    from gasfm_b200.graphs import GraphedStep
    from gasfm_b200.scene import Scene

    sync = world > 1
    weak = sync and scaling == "weak"
    m = cfg["m"]
    n_total, obs_total = (cfg["n"] * world, cfg["n_obs"] * world) if weak else (cfg["n"], cfg["n_obs"])
    idx, vals = observations(m, n_total, obs_total, cfg["seed"])
    E_total = idx.shape[1]
    if sync:
        scene_host = gdist.shard_scene(idx, vals, m, n_total, rank, world, exchange).pin_memory()
    else:
        scene_host = Scene.from_observations(idx, vals, m, n_total).pin_memory()
    model = make_model(cfg).to(dev)
    n_gat = 2 * (cfg["num_layers"] + 1)
    scene_dev = scene_host.to(dev).prepare()
    bucket = gdist.LocalGradBucket(model, exchange) if sync else None

    def loss_fn(out):
        # replicated term (every rank evaluates it) + rank-local term (its own tracks); N = 1: the plain mean over all points
        return out["Ps_norm"].square().mean() + out["pts3D"].square().sum() / (4 * n_total)

    def step(scene):
        if bucket is not None:
            bucket.prepare()
        else:
            model.zero_grad(set_to_none=True)
        out = model(scene)
        loss = loss_fn(out)
        loss.backward()
        if bucket is not None:
            bucket.allreduce()
        return out, loss

    step(scene_dev)
    torch.cuda.synchronize()
    l0 = _lib.launch_count
    n_eager = 3
    eager_ms = timed(lambda: step(scene_dev), n_eager, 1, sync_dist=sync)
    launches = (_lib.launch_count - l0) // (n_eager + 1)
    _ops = __import__("gasfm_b200.ops", fromlist=["ops"])
    recompute = bool(_ops.activation_recompute_used())
    recompute_blocks = list(_ops.last_recompute_plan)          # [blocks that recompute their activations, blocks]
    peak_mem_gb = torch.cuda.max_memory_allocated() / 2 ** 30

    # the device-resident step is replayed as ONE CUDA graph (same kernels, no per-launch host overhead); sharded steps too:
    # their exchanges are peer-memory kernels, not NCCL calls
    step_fn, graphed, gstep = (lambda: step(scene_dev)), False, None
    if not args.no_graph and (not sync or isinstance(exchange, gdist.PeerExchange)):
        ok = 1
        try:
            torch.cuda.empty_cache()
            gstep = GraphedStep(model, scene_dev, loss_fn, warmup=2,
                                before_forward=bucket.prepare if bucket else None,
                                after_backward=bucket.allreduce if bucket else None)
        except Exception as exc:  # capture is an optimisation; report and fall back to eager timing
            print(f"[bench] rank {rank}: CUDA graph capture failed, timing eagerly: {exc}", file=sys.stderr)
            ok = 0
        if sync:
            flag = torch.tensor([ok], device=dev)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            ok = int(flag.item())
        if ok:
            step_fn, graphed = gstep, True
        else:
            gstep = None
            torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    ms = timed(step_fn, args.steps, args.warmup, sync_dist=sync)
    clocks = sampler.stop()
    del gstep, step_fn
    torch.cuda.empty_cache()

    def fwd_only():
        with torch.no_grad():
            model(scene_dev)
    fwd_ms = timed(fwd_only, max(3, args.steps // 2), 2, sync_dist=sync)

    # ---- end to end: host scene -> device -> index build -> fwd+bwd -> predictions back on the host ----
    holder = {}

    def step_e2e():
        out, loss = step(scene_host.to(dev, non_blocking=True))
        holder["Ps"] = out["Ps_norm"].detach().cpu()
        holder["pts"] = out["pts3D"].detach().cpu()
        holder["loss"] = float(loss.detach())
    del scene_dev
    torch.cuda.empty_cache()
    e2e_eager_ms = timed(step_e2e, max(3, args.steps // 2), 2, sync_dist=sync)
    e2e_ms, e2e_graphed = e2e_eager_ms, False
    if not args.no_graph and (not sync or isinstance(exchange, gdist.PeerExchange)):
        # the same host-to-host step with the index build inside ONE captured graph (gasfm_b200.graphs.StreamedStep): per step
        # the pinned host scene is copied into the captured buffers, the graph replayed, predictions + loss + index status read back
        from gasfm_b200.graphs import StreamedStep
        ok, streamed = 1, None
        torch.cuda.empty_cache()
        try:
            streamed = StreamedStep(model, scene_host, loss_fn, outputs=("Ps_norm", "pts3D"), device=dev,
                                    before_forward=bucket.prepare if bucket else None,
                                    after_backward=bucket.allreduce if bucket else None)
        except Exception as exc:
            print(f"[bench] rank {rank}: streamed-step capture failed, e2e stays eager: {exc}", file=sys.stderr)
            ok = 0
        if sync:
            flag = torch.tensor([ok], device=dev)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            ok = int(flag.item())
        if ok:
            def step_streamed():
                holder.update(streamed(scene_host))
            e2e_ms, e2e_graphed = timed(step_streamed, max(3, args.steps // 2), 2, sync_dist=sync), True
            holder["Ps"], holder["pts"] = holder["Ps_norm"], holder["pts3D"]
        del streamed
        torch.cuda.empty_cache()
    h2d = (scene_host.x.values.numel() * 4 + scene_host.x.indices.numel() * 8 +
           scene_host.x.cam_per_pts.numel() * 8 + scene_host.x.pts_per_cam.numel() * 8 +
           sum(w.valid_indices.numel() * 8 for k, w in scene_host.graph_wrappers.items() if k.endswith("2global")))
    d2h = holder["Ps"].numel() * 4 + holder["pts"].numel() * 4 + 4
    if sync:
        t = torch.tensor([h2d, d2h], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t)
        h2d, d2h = int(t[0].item()), int(t[1].item())
    if exchange is not None:
        exchange.check()
    return {"E": E_total, "n_total": n_total, "ms": ms, "value": E_total * n_gat / (ms / 1e3), "eager_ms": eager_ms,
            "forward_ms": fwd_ms, "e2e_ms": e2e_ms, "e2e_value": E_total * n_gat / (e2e_ms / 1e3), "h2d": int(h2d), "d2h": int(d2h),
            "launches": int(launches), "graphed": graphed, "clocks": clocks, "recompute": recompute,
            "recompute_blocks": recompute_blocks, "peak_mem_gb": round(peak_mem_gb, 1),
            "e2e_graphed": e2e_graphed, "e2e_eager_ms": e2e_eager_ms}


def parity_sharded_vs_single(dev, rank, world, exchange):
    """A small scene through the track-sharded model on all ranks and through the single-GPU model on rank 0: outputs
    and every parameter gradient must agree (the driver's pytest box has one GPU; this is the NCCL-free multi-GPU path's
    driver-side parity record)."""
    from gasfm_b200 import dist as gdist
    from gasfm_b200.scene import Scene

    cfg = dict(CFG2, n=20_000, n_obs=200_000, num_layers=2, seed=1)
    idx, vals = observations(cfg["m"], cfg["n"], cfg["n_obs"], cfg["seed"])
    model = make_model(cfg).to(dev)
    g = torch.Generator().manual_seed(1)
    wP = torch.rand(cfg["m"], 3, 4, generator=g).to(dev)
    wX = torch.rand(4, cfg["n"], generator=g).to(dev)
    sh = gdist.shard_scene(idx, vals, cfg["m"], cfg["n"], rank, world, exchange).to(dev)
    lo, hi = sh.shard.col_begin, sh.shard.col_end
    bucket = gdist.LocalGradBucket(model, exchange)
    bucket.prepare()
    o = model(sh)
    ((o["Ps_norm"] * wP).sum() + (o["pts3D"] * wX[:, lo:hi]).sum()).backward()
    bucket.allreduce()
    got = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    got_out = {"Ps_norm": o["Ps_norm"].detach().clone(), "pts3D": o["pts3D"].detach().clone()}
    exchange.check()
    res = None
    if rank == 0:
        model.zero_grad(set_to_none=True)
        o1 = model(Scene.from_observations(idx, vals, cfg["m"], cfg["n"]).to(dev))
        ((o1["Ps_norm"] * wP).sum() + (o1["pts3D"] * wX).sum()).backward()
        o1 = {k: v.detach() for k, v in o1.items()}
        want = {k: p.grad.detach() for k, p in model.named_parameters() if p.grad is not None}
        worst, key = _grad_error(got, want)
        e_ps = float((got_out["Ps_norm"] - o1["Ps_norm"]).abs().max() / max(1.0, float(o1["Ps_norm"].abs().max())))
        e_pts = float((got_out["pts3D"] - o1["pts3D"][:, lo:hi]).abs().max() / max(1.0, float(o1["pts3D"].abs().max())))
        res = {"against": f"single-GPU model on rank 0, {cfg['m']} x {cfg['n']} scene (E={idx.shape[1]}), d=256, 2 blocks",
               "max_err_outputs": {"Ps_norm": e_ps, "pts3D(rank 0 tracks)": e_pts}, "tolerance_outputs": 1e-4,
               "worst_grad_err": worst, "worst_grad_param": key, "tolerance_grads": 2e-3,
               "ok": bool(max(e_ps, e_pts) < 1e-4 and worst < 2e-3)}
    torch.distributed.barrier()
    return res


def exchange_timings(dev, world, exchange, m, hc=256, heads=4):
    """The cross-GPU exchange kernels alone (CUDA events, max over ranks): one log-sum-exp merge of [m, hc] partials, one sum of an
    [m, hc] gradient and one of a 4.5 M-float bucket -- with the NVLink bound next to them: every rank pushes its payload to
    world - 1 peers, bytes / 770 GB/s (measured peer-copy bandwidth per direction, /opt/skills/guides/B200_PROFILING.md)."""
    torch.manual_seed(0)
    acc = torch.randn(m, hc, device=dev)
    mx, sm = torch.randn(m, heads, device=dev), torch.rand(m, heads, device=dev) + 1.0
    grad = torch.randn(m, hc, device=dev)
    bucket = torch.randn(4_500_000, device=dev)
    res = {}
    for name, fn, floats in (("lse_merge", lambda: exchange.lse_merge(acc, mx, sm, heads), m * (hc + 2 * heads)),
                             ("allreduce_sum [m, d]", lambda: exchange.allreduce_sum(grad), m * hc),
                             ("allreduce_sum local-gradient bucket", lambda: exchange.allreduce_sum(bucket), bucket.numel())):
        us = 1e3 * timed(fn, 20, 5, sync_dist=True)
        pushed = floats * 4 * (world - 1)
        res[name] = {"us": round(us, 1), "bytes_pushed_per_rank": pushed, "nvlink_bound_us": round(pushed / 770e3, 1)}
    return res


def run_ours(args):
    from gasfm_b200 import _lib
    from gasfm_b200 import dist as gdist

    _lib.load()   # fail loudly if the CUDA library is missing
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg = dict(WORKLOADS[args.workload])
    peaks, peak_kind = measured_peaks()
    exchange, exchange_kind = None, None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
        if args.exchange == "peer":
            try:
                exchange, exchange_kind = gdist.PeerExchange(dev), "peer-memory kernels (cudaIpc over NVLink), CUDA-graph replay"
            except Exception as exc:
                print(f"[bench] rank {rank}: peer-memory exchange unavailable ({exc}); using NCCL all_gather", file=sys.stderr)
        flag = torch.tensor([0 if exchange is None else 1], device=dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        if int(flag.item()) == 0:
            exchange, exchange_kind = gdist.CollectiveExchange(dev), "NCCL all_gather + merge kernels, eager"

    r = measure_workload(cfg, args, dev, rank, world, exchange, args.scaling)
    parity = parity_sharded_vs_single(dev, rank, world, exchange) if world > 1 else None
    exch = exchange_timings(dev, world, exchange, cfg["m"]) if world > 1 else None
    if rank != 0:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
        return
    scaling = args.scaling if world > 1 else "strong"
    wc = workload_config(dict(cfg, n=r["n_total"]), r["E"], world, scaling)
    line = {"metric": METRIC, "value": r["value"], "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms"], "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wc,
            "forward_ms_per_scene": r["forward_ms"], "cuda_graph": r["graphed"], "eager_ms_per_step": r["eager_ms"],
            "activation_recompute": r["recompute"], "recompute_blocks": r["recompute_blocks"], "peak_mem_gb_eager_step": r["peak_mem_gb"],
            "exchange": exchange_kind, "exchange_kernels": exch,
            "e2e": {"value": r["e2e_value"], "unit": "edges/s", "ms_per_step": r["e2e_ms"],
                    "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "cuda_graph": r["e2e_graphed"], "eager_ms_per_step": r["e2e_eager_ms"]},
            "gpu_launches": r["launches"] * args.steps, "gpu_launches_per_step": r["launches"],
            "clocks": r["clocks"], "roofline": None, "cpu_baseline": None, "parity": parity}
    if world == 1:
        if not args.no_roofline:
            line["roofline"] = kernel_roofline(cfg, peaks, peak_kind, r["forward_ms"], recompute=r["recompute"])
            torch.cuda.empty_cache()
        if not args.no_cpu_baseline:
            line["cpu_baseline"], _, ref = cpu_baseline(cfg, repeats=1)
            line["parity"] = parity_vs_cpu(cfg, ref, dev)
            del ref
        if args.workload != "cfg2" and not args.no_cfg2:
            torch.cuda.empty_cache()
            r2 = measure_workload(dict(CFG2), args, dev, rank, 1, None, "strong")
            line["cfg2"] = {"workload": workload_config(CFG2, r2["E"], 1, "strong")["workload"], "value": r2["value"], "unit": "edges/s",
                            "ms_per_step": r2["ms"], "eager_ms_per_step": r2["eager_ms"], "forward_ms_per_scene": r2["forward_ms"],
                            "e2e": {"value": r2["e2e_value"], "ms_per_step": r2["e2e_ms"], "h2d_bytes_per_step": r2["h2d"],
                                    "d2h_bytes_per_step": r2["d2h"], "cuda_graph": r2["e2e_graphed"],
                                    "eager_ms_per_step": r2["e2e_eager_ms"]},
                            "gpu_launches_per_step": r2["launches"], "cuda_graph": r2["graphed"], "activation_recompute": r2["recompute"]}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# cfg4: multi-scene training step, scenes sharded over the GPUs (BASELINE.json configs[3])
# ---------------------------------------------------------------------------------------------
CFG4_SCENES = 8


def cfg4_scene_shapes():
    """8 synthetic scenes, 100 ... 500 views, 60 tracks per view, 3 % density."""
    shapes = []
    for k in range(CFG4_SCENES):
        m = 100 + (400 * k) // (CFG4_SCENES - 1)
        shapes.append((m, 60 * m, int(0.03 * m * 60 * m), 100 + k))
    return shapes


def run_cfg4(args):
    """One optimisation step of the reference's multi-scene learning (code/train.py:61-137): forward + ESFM loss + backward
    over every scene of the batch, gradients SUMMED over the scenes, Adam step.  Scenes are dealt to the ranks in snake
    order (largest with smallest).  Timed form: every local scene's forward+loss+backward replayed as a CUDA graph into one flat
    gradient bucket, one NCCL all-reduce, Adam.  The eager forms (bucketed all-reduce overlapped with backward / one flat
    all-reduce after backward) are timed next to it: at the shipped d = 32 they are launch-bound."""
    from gasfm_b200 import _lib
    from gasfm_b200 import dist as gdist
    from gasfm_b200.config import ConfigTree, gasfm_conf
    from gasfm_b200.loss_functions import ESFMLoss
    from gasfm_b200.models.graph_attn_sfm import GraphAttnSfMNet
    from gasfm_b200.scene import Scene

    _lib.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sync = world > 1
    if sync:
        torch.distributed.init_process_group("nccl", device_id=dev)
    shapes = cfg4_scene_shapes()
    order = sorted(range(CFG4_SCENES), key=lambda k: -shapes[k][2])
    owner = {k: (i % world if (i // world) % 2 == 0 else world - 1 - i % world) for i, k in enumerate(order)}
    mine = [k for k in range(CFG4_SCENES) if owner[k] == rank]
    hosts, E_total = [], 0
    for k, (m, n, n_obs, seed) in enumerate(shapes):
        if k in mine:
            idx, vals = observations(m, n, n_obs, seed)
            hosts.append(Scene.from_observations(idx, vals, m, n).pin_memory())
    E_mine = sum(h.x.indices.shape[1] for h in hosts)
    conf = gasfm_conf()
    conf["loss"] = ConfigTree.from_dict(dict(infinity_pts_margin=1e-4, hinge_loss=True, hinge_loss_weight=1,
                                             pts_grad_equalization_pre_perspective_divide=True,
                                             normalize_grad_wrt_valid_projections_only=True))
    torch.manual_seed(0)
    model = GraphAttnSfMNet(conf).to(dev)
    loss_fn = ESFMLoss(conf)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    reducer = gdist.BucketedGradReducer(model) if sync else None
    scenes_dev = [h.to(dev).prepare() for h in hosts]
    n_params = sum(p.numel() for p in model.parameters())

    def step(scenes, overlap=True):
        if reducer is not None and overlap:
            reducer.prepare()
        else:
            opt.zero_grad(set_to_none=True)
        total = None
        for i, sc in enumerate(scenes):
            loss = loss_fn(model(sc), sc)
            if reducer is not None and overlap and i < len(scenes) - 1:
                with reducer.accumulate_only():
                    loss.backward()
            else:
                loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        if reducer is not None and overlap:
            reducer.finish()
        elif sync:
            gdist.allreduce_gradients(model.parameters())        # A/B: one flat all-reduce after backward
        opt.step()
        return total

    flat_ms = timed(lambda: step(scenes_dev, overlap=False), max(2, args.steps // 2), 2, sync_dist=sync) if sync else None
    opt.zero_grad(set_to_none=True)
    launches_probe = _lib.launch_count
    step(scenes_dev)
    launches = _lib.launch_count - launches_probe
    eager_ms = timed(lambda: step(scenes_dev), max(2, args.steps // 2), 2, sync_dist=sync)

    # graph replay: forward + loss + backward of every local scene captured once (no collective inside a capture); the
    # gradients of all scenes accumulate in ONE flat bucket, which a single NCCL all-reduce sums over the ranks afterwards
    graphed, replay_step = False, None
    if not args.no_graph:
        ok = 1
        try:
            from gasfm_b200.graphs import GraphedStep
            if reducer is not None:
                reducer.enabled = False                       # the bucket hooks must not fire inside a capture
            flat = gdist.LocalGradBucket(model, select=lambda name: True)
            graphs = [GraphedStep(model, sc, (lambda out, sc=sc: loss_fn(out, sc)), warmup=2,
                                  before_forward=flat.prepare if i == 0 else flat.attach) for i, sc in enumerate(scenes_dev)]

            def replay_step():
                total = None
                for g in graphs:
                    loss = g()
                    total = loss.detach() if total is None else total + loss.detach()
                if sync:
                    flat.allreduce_nccl()
                opt.step()
                return total
        except Exception as exc:
            print(f"[bench] rank {rank}: CUDA graph capture failed, timing eagerly: {exc}", file=sys.stderr)
            ok = 0
        if sync:
            flag = torch.tensor([ok], device=dev)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            ok = int(flag.item())
        graphed = bool(ok)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(replay_step if graphed else (lambda: step(scenes_dev)), args.steps, args.warmup, sync_dist=sync)
    clocks = sampler.stop()
    if graphed:
        del graphs
        flat.attach()                                         # leave .grad pointing at live memory for the eager e2e steps
        if reducer is not None:
            reducer.enabled = True
    torch.cuda.empty_cache()
    holder = {}

    def step_e2e():
        holder["loss"] = float(step([h.to(dev, non_blocking=True) for h in hosts]))
    e2e_ms = timed(step_e2e, max(3, args.steps // 2), 2, sync_dist=sync)
    h2d = sum(h.x.values.numel() * 4 + h.x.indices.numel() * 8 + h.x.cam_per_pts.numel() * 8 + h.x.pts_per_cam.numel() * 8 for h in hosts)
    t = torch.tensor([float(E_mine), float(h2d), float(launches)], device=dev, dtype=torch.float64)
    if sync:
        torch.distributed.all_reduce(t)
    E_total, h2d, launches_all = int(t[0].item()), int(t[1].item()), int(t[2].item())
    if rank == 0:
        n_gat = 2 * (12 + 1)
        line = {"metric": METRIC, "value": E_total * n_gat / (ms / 1e3), "unit": "edges/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"cfg4: multi-scene training step over {CFG4_SCENES} synthetic scenes (100..500 views, 60 tracks per view, "
                                       f"3% density, E={E_total} observations in total), shipped GASFM model ({n_params} parameters) replicated, "
                                       "sparse ESFM loss, gradients SUMMED over scenes and ranks (one NCCL all-reduce of a flat bucket after the graph replays), Adam step",
                           "edge_level_gats_per_step": n_gat * CFG4_SCENES, "parallelism": f"{CFG4_SCENES} scenes dealt to {world} GPU(s)",
                           "cache": "per-scene forward+backward replayed as CUDA graphs, one NCCL all-reduce of the flat gradient bucket, Adam"},
                "scenes_per_s": CFG4_SCENES / (ms / 1e3),
                "eager_ms_per_step": eager_ms,
                "grad_allreduce": None if not sync else {"bytes": n_params * 4, "eager_ms_per_step_flat_after_backward": flat_ms,
                                                         "eager_ms_per_step_bucketed_overlapped": eager_ms},
                "e2e": {"value": E_total * n_gat / (e2e_ms / 1e3), "unit": "edges/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * world},
                "gpu_launches": launches_all * args.steps, "gpu_launches_per_step": launches_all, "clocks": clocks,
                "cuda_graph": graphed, "roofline": None, "cpu_baseline": None, "parity": None}
        print(json.dumps(line))
    if sync:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS) + ["cfg4"],
                    help="default cfg3_d256 = BASELINE.json's target scene (1,000 x 300k, ~5M observations) at d = 256")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N>1: strong = the workload's scene itself, sharded; weak = one scene of N x the workload's tracks")
    ap.add_argument("--exchange", default="peer", choices=["peer", "collective"],
                    help="N>1: peer = hand-written peer-memory kernels (graph-captured step); collective = NCCL all_gather arm (eager)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-cfg2", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the device-resident step eagerly instead of as a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "cfg4":
            args.workload = "cfg2"      # the CPU arm's bounded sample is scene-shaped; cfg4's scenes are cfg2-sized
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    if args.workload == "cfg4":
        return run_cfg4(args)
    return run_ours(args)


if __name__ == "__main__":
    main()

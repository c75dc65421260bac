#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json configs[2]).

A step = one forward+backward pass of the rasterizer over one view:
1,048,576 pixel-aligned Gaussians (2 context panoramas x 512 x 1024, SH degree 4 = 340 B/Gaussian),
one native-equirectangular 512x1024 target view, MSE seed gradient, all five gradient outputs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); every rank renders its own scene/view (weak scaling,
views are independent) and the only collective is the NCCL all-reduce of the scalar loss.

Prints ONE JSON line (see the task contract): value = Gaussians/s with inputs resident in HBM,
e2e = the same through the decoder-level API with pinned HOST buffers (H2D + layout prep + fwd + bwd +
D2H loss inside the timed region), roofline for the dominant kernel (CUDA events recorded by the
library around each stage, on the launching stream), cpu_baseline = the C oracle on the host cores.

``--impl reference`` times the CPU oracle port (the reference has no CPU implementation and its CUDA
extension is an absent pip dependency) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 512, 1024
SH_DEGREE = 4
CONFIG_ID = 3  # seed = 1234 + config id
WORKLOAD = ("configs[2]: 1,048,576 pixel-aligned Gaussians (2 context ERP 512x1024, SH deg 4, 340 B/Gaussian), "
            "one 512x1024 native-ERP target view, forward+backward (MSE seed gradient; dL/d means, cov, opacity, SH, means2D)")
REF_SAMPLE_P = 1 << 30   # the reference arm runs the FULL workload (every Gaussian of the view), like the GPU arm


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Polls SM clock + throttle reasons of one GPU through NVML while the benchmark runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=1.0)
        s = self.samples
        return {
            "sm_mhz": statistics.median(s) if s else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(s),
            "how": "NVML polled every 20 ms from the first warm-up step to the end of the timed region",
        }


def build_scene(device, seed):
    import torch
    from splatter360_b200 import synthetic
    sc = synthetic.pixel_aligned_scene(H, W, sh_degree=SH_DEGREE, seed=seed, device=device)
    return sc


def algorithmic_bytes(P, P_vis, N):
    """Per-view algorithmic HBM bytes per stage (SURVEY.md sec. 8d adapted to the 48-B geometry record,
    the 36-B SH Jacobian and matrix binning).  Stage names are the library's timer slots: in the matrix-binning path
    "emit" = count matrix (mb_count), "scan" = column scan (mb_colscan), "tile_sort" = ranked scatter (mb_scatter)."""
    npix = H * W
    tiles = (H // 16) * (W // 16)
    chunks = (P + 2047) // 2048
    return {
        "preprocess": 340 * P + (48 + 36) * P_vis + (4 + 8 + 1 + 8) * P,     # inputs; geom record + SH Jacobian; radii+rect+clamp+key/id
        "depth_sort": 4 * (4 + 8 + 8) * P + 4 * P,                          # hist read + 4 passes x scatter r/w 16 B
        "scan": 2 * 4 * chunks * tiles,                                     # count matrix read + prefix write
        "emit": (4 + 8) * P + 4 * chunks * tiles,                           # id + rect gather; count matrix row out
        "tile_sort": (4 + 8) * P + 4 * chunks * tiles + 8 * tiles + 4 * N,  # id + rect, prefix row, ranges; 4 B per instance out
        "tile_ranges": 4 * tiles + 8 * tiles + 8 * tiles,
        "render_fwd": (4 + 48) * N + 20 * npix,                             # id + record per instance; colour + T + n_contrib out
        "render_bwd": 20 * npix + (4 + 48) * N + 72 * N,                    # pixel grads, T, n_contrib; instance refetch; 36-B grad record r-m-w
        "preprocess_bwd": 2 * 48 * P + (12 + 24 + 4 + 36 + 1 + 4) * P_vis + 352 * P,   # accumulator zero-fill + read; inputs + Jacobian; all gradient writes
    }


def run_ours(args):
    import torch
    import torch.distributed as dist
    from splatter360_b200 import _lib, camera, synthetic
    from splatter360_b200.decoder import render_erp
    from splatter360_b200.loss import mse_loss
    from splatter360_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    K, Wm = args.steps, max(args.warmup, 3)
    seed = 1234 + CONFIG_ID + 1000 * rank
    sc = build_scene(dev, seed)
    P = sc.means.shape[0]
    means = sc.means.contiguous().requires_grad_()
    cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous().requires_grad_()
    opac = sc.opacities[:, None].contiguous().requires_grad_()
    shs = sc.harmonics.permute(0, 2, 1).contiguous().requires_grad_()
    poses = synthetic.trajectory(K + Wm, seed=rank).to(dev)
    cams = camera.erp_camera(poses)
    bg = torch.zeros(3, device=dev)
    target = torch.rand(3, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(seed))
    from splatter360_b200.parallel import AsyncLossReducer
    from splatter360_b200.rasterizer import CapacityTracker
    reducer = AsyncLossReducer(dev)   # NCCL all-reduce of the scalar loss, issued async: step i+1 does not wait for it
    # sync-free steady state: the instance buffers of step i are sized from the counts of earlier steps (+25 %); the
    # device-side overflow flag of every step is collected asynchronously and checked after the timed region
    tracker = None if args.exact_counts else CapacityTracker()

    def step(i):
        s = GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=bg, scale_modifier=1.0,
            viewmatrix=cams.view_matrix[i], projmatrix=cams.full_projection[i], sh_degree=SH_DEGREE,
            campos=cams.campos[i], prefiltered=False, debug=False, projection="erp", capacity_tracker=tracker)
        m2d = torch.zeros_like(means, requires_grad=True)
        for t in (means, cov6, opac, shs):
            t.grad = None
        color, _ = GaussianRasterizer(s)(means3D=means, means2D=m2d, shs=shs, colors_precomp=None,
                                         opacities=opac, cov3D_precomp=cov6)
        loss = mse_loss(color, target)          # fused loss + seed gradient (reference: loss_mse.py:30-31)
        loss.backward()
        reducer.submit(loss)

    graphed = None
    if args.graph:
        # steady-state API for a fixed problem shape: forward + fused MSE loss + backward as ONE CUDA-graph launch, Gaussian
        # tensors used in place, only the camera block changes between steps (splatter360_b200.graph.GraphedStep)
        from splatter360_b200.graph import GraphedStep
        s0 = GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=bg, scale_modifier=1.0, viewmatrix=cams.view_matrix[0],
            projmatrix=cams.full_projection[0], sh_degree=SH_DEGREE, campos=cams.campos[0], prefiltered=False, debug=False,
            projection="erp")
        graphed = GraphedStep(s0, means.detach(), cov6.detach(), opac.detach().reshape(-1), shs.detach(), target, margin=1.3)
        launches_per_replay = None

        def step(i):  # noqa: F811
            graphed.set_camera(cams.view_matrix[i], cams.full_projection[i], cams.campos[i])
            loss, _ = graphed.replay()
            reducer.submit(loss)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for i in range(Wm):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    _lib.profile_read(reset=True)
    _lib.profile_enable(not args.graph)   # stage events cannot be recorded inside a replayed graph
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("timed")
    e0.record()
    for i in range(K):
        step(Wm + i)
    reducer.flush()           # every all-reduce of the timed steps completes inside the timed region
    e1.record()
    torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    overflow_steps = []
    if graphed is not None and graphed.overflowed():
        raise SystemExit("bench invalid: the graphed step's instance capacity overflowed")
    if tracker is not None and graphed is None:
        tracker.flush()
        overflow_steps = [int(x) for x in tracker.overflowed]
        if overflow_steps:
            raise SystemExit(f"bench invalid: instance capacity overflowed at steps {overflow_steps} (CapacityTracker margin too small)")
    clocks = sampler.stop() if sampler else None
    _lib.profile_enable(False)
    launches = _lib.launch_count() - l0
    stages = _lib.profile_read(reset=True)
    if graphed is not None:
        # per-stage times and the launch count of the SAME steps run eagerly right after the timed region (stage events
        # cannot be recorded inside a replayed graph; a replay launches exactly the kernels its capture recorded)
        from splatter360_b200 import rasterizer as R_
        tr = CapacityTracker()
        mse_grad = lambda c_: 2.0 * (c_ - target) / c_.numel()

        def eager(i):
            s_ = GaussianRasterizationSettings(
                image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=bg, scale_modifier=1.0, viewmatrix=cams.view_matrix[i],
                projmatrix=cams.full_projection[i], sh_degree=SH_DEGREE, campos=cams.campos[i], prefiltered=False, debug=False,
                projection="erp", capacity_tracker=tr)
            c_, st_ = R_.forward_raw(s_, means.detach(), cov6.detach(), opac.detach().reshape(-1), shs.detach(), None)
            R_.backward_raw(s_, means.detach(), cov6.detach(), opac.detach().reshape(-1), shs.detach(), None, st_, mse_grad(c_))

        for i in range(3):
            eager(Wm + i)
        torch.cuda.synchronize()
        _lib.profile_read(reset=True); _lib.profile_enable(True)
        l1 = _lib.launch_count()
        n_e = min(K, 20)
        for i in range(n_e):
            eager(Wm + i)
        torch.cuda.synchronize()
        _lib.profile_enable(False)
        launches = (_lib.launch_count() - l1 + n_e) * K // n_e   # + the fused loss kernel of the graphed step
        stages = _lib.profile_read(reset=True)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    final_loss = float(reducer.latest().item())

    # instance statistics of the last step (for the algorithmic-byte model)
    with torch.no_grad():
        from splatter360_b200 import rasterizer as R
        s_last = GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=bg, scale_modifier=1.0,
            viewmatrix=cams.view_matrix[Wm + K - 1], projmatrix=cams.full_projection[Wm + K - 1],
            sh_degree=SH_DEGREE, campos=cams.campos[Wm + K - 1], prefiltered=False, debug=False, projection="erp")
        _, st = R.forward_raw(s_last, means.detach(), cov6.detach(), opac.detach().reshape(-1), shs.detach(), None)
        N, P_vis = st.num_rendered, st.num_visible
        del st

    # ---- e2e: decoder-level API with pinned host buffers --------------------------------------
    e2e = None
    if not args.no_e2e:
        host = dict(
            means=sc.means.detach().cpu().pin_memory(), cov=sc.covariances.detach().cpu().pin_memory(),
            sh=sc.harmonics.detach().cpu().pin_memory(), op=sc.opacities.detach().cpu().pin_memory(),
            target=target.cpu().pin_memory(), poses=poses.cpu().pin_memory())
        h2d = sum(host[k].numel() * 4 for k in ("means", "cov", "sh", "op", "target")) + 64
        near = torch.ones(1, device=dev)
        far = torch.full((1,), 100.0, device=dev)

        from splatter360_b200.io import HostSceneFeeder
        feeder = HostSceneFeeder(dev)

        def host_inputs(i):
            return dict(means=host["means"], cov=host["cov"], sh=host["sh"], op=host["op"], target=host["target"],
                        pose=host["poses"][i:i + 1])

        e2e_tracker = None if args.exact_counts else CapacityTracker()

        def e2e_compute(d):
            m, c, sh, o = (d[k].requires_grad_() for k in ("means", "cov", "sh", "op"))
            img = render_erp(d["pose"], near, far, (H, W), bg[None], m[None], c[None], sh[None], o[None],
                             scale_invariant=False, capacity_tracker=e2e_tracker)
            loss = mse_loss(img[0], d["target"])
            loss.backward()
            if world > 1:
                l = loss.detach().reshape(1).clone()
                dist.all_reduce(l)
                return float(l.item())
            return float(loss.item())  # D2H read of the step result

        def e2e_run(n, first):
            """n steps; the upload of step i+1 (copy stream) overlaps the rasterization of step i."""
            t = feeder.submit(host_inputs(first))
            for i in range(n):
                d = feeder.get(t)
                if i + 1 < n:
                    t = feeder.submit(host_inputs(first + (i + 1) % K))
                e2e_compute(d)

        Ke = max(3, min(K, 20))
        e2e_run(3, 0)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e2e_run(Ke, Wm)
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item()) / Ke
        e2e = {"value": P * world / (e2e_ms * 1e-3), "unit": "Gaussians/s", "ms_per_step": e2e_ms, "steps": Ke,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4, "numa_bind": _NUMA["bind"],
               "api": "splatter360_b200.io.HostSceneFeeder (pinned host -> device, upload of step i+1 overlaps step i) + "
                      "splatter360_b200.decoder.render_erp fwd + bwd + loss D2H; every step's H2D copy is inside the timed region"}

    # ---- the reference's own way of producing this panorama: six 90-degree pinhole faces of edge H/2
    # (model_wrapper_erp.py:202-205, 336-345) + Cube2Equirec -- once as six separate rasterizer calls (the reference's
    # call pattern through this library's kernels), once as ONE batched pass + the stitch kernel.  Side measurement on
    # rank 0 at N=1; it does not enter `value`.
    cube6 = None
    if world == 1 and not args.no_cube6:
        try:
            cube6 = cube6_run(dev, means.detach(), cov6.detach(), opac.detach().reshape(-1), shs.detach(), poses[Wm + K - 1])
        except Exception as ex:   # a side measurement must never cost the headline line
            cube6 = {"error": f"{type(ex).__name__}: {ex}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / K
    value = P * world * K / (total_ms * 1e-3)
    peak, peak_src = load_peaks()
    alg = algorithmic_bytes(P, P_vis, N)
    stage_ms = {k: (v[0] / max(v[1], 1)) for k, v in stages.items()}
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    achieved = alg[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get(dom)
        except Exception:
            traffic = None
    whole = sum(alg.values())
    # what actually bounds the dominant kernel, from the committed ncu capture of this same command (profiles/)
    ncu_note = None
    try:
        import re
        kname = {"render_bwd": "render_backward_kernel", "render_fwd": "render_forward_kernel",
                 "preprocess": "preprocess_kernel", "preprocess_bwd": "preprocess_backward_kernel"}.get(dom)
        txt = open(os.path.join(ROOT, "profiles", "r02_ncu_full_summary.txt")).read()
        blk = next(b for b in txt.split("== ")[1:] if kname and kname in b.split("\n")[0])
        pick = lambda key: float(re.search(key + r"\s+([\d.]+)", blk).group(1))
        ncu_note = {"source": "profiles/r02_ncu_full_summary.txt (ncu --set full of this command)",
                    "issue_slot_utilisation_pct": pick("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "fma_pipe_utilisation_pct": pick("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                    "dram_utilisation_pct": pick("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
    except Exception:
        ncu_note = None
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg[dom], "kernel_ms": stage_ms[dom],
        "note": "render kernels are FP32-ALU/latency bound, not HBM bound (DESIGN.md sec. 5); the fraction is reported against the HBM roof as the contract asks",
        "whole_step": {"algorithmic_bytes": whole, "achieved": whole / (ms_per_step * 1e-3) / 1e9,
                       "frac": whole / (ms_per_step * 1e-3) / 1e9 / peak},
        "stages_ms": stage_ms,
        "stage_timing": ("CUDA events recorded by the library around every stage on the launching stream; the timed region itself is "
                         "ONE graph launch per step (events cannot be recorded inside a replayed graph), so the per-stage events come from "
                         "an eager pass over the same steps right after it (bench.py --no-graph has them inside the timed region: same numbers)"
                         if graphed is not None else "CUDA events recorded by the library around every stage on the launching stream, inside the timed region"),
        "stage_kernels": {"emit": "mb_count_kernel", "scan": "mb_colscan_kernel", "tile_ranges": "tile_scan_kernel", "tile_sort": "mb_scatter_kernel",
                          "depth_sort": "rs_global_hist_kernel + 4 x rs_onesweep_kernel"},
        "stages_GBps": {k: (alg[k] / (stage_ms[k] * 1e-3) / 1e9 if stage_ms[k] > 0 else None) for k in alg},
        "instances": {"P": P, "P_visible": P_vis, "N": N},
        "ncu": ncu_note,
    }

    cpu_baseline, parity = None, None
    if world == 1 and not args.no_cpu:
        import numpy as np
        # image-like seed gradient (low-pass random field, what an MSE against a real image produces) for the parity verdict;
        # white noise -- the worst case for float32: the per-Gaussian moment sums cancel almost completely, two builds of the
        # ORACLE alone (with / without FMA contraction) then differ by 5e-5 ... 1e-4 -- is reported next to it
        lo = torch.randn(1, 3, H // 16, W // 16, generator=torch.Generator().manual_seed(0))
        dL = (torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)[0] / (3 * H * W)).numpy()
        pose_last = poses[Wm + K - 1].cpu()
        cpu_baseline, oref = cpu_oracle_run(sc, pose_last, P, dL=dL, keep=True)
        _, oref64 = cpu_oracle_run(sc, pose_last, P, dL=dL, keep=True, f64=True)
        gpu_args = (s_last._replace(capacity_tracker=None), means.detach(), cov6.detach(), opac.detach().reshape(-1), shs.detach())
        parity = parity_block(oref, oref64, *gpu_args, dL)
        parity["seed_gradient"] = "image-like (bilinear upsampling of 1/16-resolution Gaussian noise)"
        dLw = np.random.default_rng(0).standard_normal((3, H, W)).astype(np.float32) / (3 * H * W)
        _, oref_w = cpu_oracle_run(sc, pose_last, P, dL=dLw, keep=True)
        _, oref_w64 = cpu_oracle_run(sc, pose_last, P, dL=dLw, keep=True, f64=True)
        pw = parity_block(oref_w, oref_w64, *gpu_args, dLw)
        parity["white_noise_seed"] = {k: v for k, v in pw.items() if k in ("color", "d_means", "d_cov", "d_opac", "d_shs", "ok")}

    out = {
        "metric": "gaussians_per_s_fwd_bwd", "value": value, "unit": "Gaussians/s", "n_gpus": world,
        "steps": K, "warmup": Wm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "views_per_s": world * K / (total_ms * 1e-3),
        "config": {"workload": WORKLOAD, "projection": "erp", "P": P, "image": [H, W], "sh_degree": SH_DEGREE,
                   "views_per_step_per_gpu": 1,
                   "l2_policy": "inputs (356 MB/view) and gradient outputs (369 MB/view) are larger than the 126 MB L2",
                   "sharding": "one independent view per GPU per step; NCCL all-reduce of the scalar loss only",
                   "api": ("splatter360_b200.graph.GraphedStep.replay(): forward + fused MSE loss + backward captured as one CUDA graph, "
                           "Gaussians resident in HBM and used in place, camera block updated every step" if graphed is not None else
                           "diff_gaussian_rasterization-compatible GaussianRasterizer autograd call, inputs resident in HBM"),
                   "instance_buffers": ("graph: capacity = 1.3 x the count of the warm-up view; device overflow flag checked after the timed region, not set"
                                        if graphed is not None else "exact: count read back every view" if tracker is None else
                                        "sync-free: CapacityTracker (capacity = 1.25 x largest count seen; device overflow flag of every timed step checked, none set)")},
        "e2e": e2e, "gpu_launches": int(launches) * world, "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu_baseline, "parity": parity, "final_loss": final_loss, "cube6_reference_style": cube6,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()



def _dist_setup(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not getattr(args, "no_numa_bind", False):
        # pinned host buffers of the e2e legs land on the NUMA node this rank's GPU hangs off
        from splatter360_b200.io import bind_to_gpu_numa_node
        _NUMA["bind"] = bind_to_gpu_numa_node(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, local_rank, dev


_NUMA = {"bind": None}


def _timed(fn, K, Wm, world, dev):
    """Wm warm-up calls, then K timed calls bracketed by barrier + synchronize; device time, max over ranks (ms)."""
    import torch
    import torch.distributed as dist
    for i in range(Wm):
        fn(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(K):
        fn(Wm + i)
    b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_config5(args):
    """BASELINE.json configs[4]: 3M Gaussians, 1024x2048 ERP video path, forward only, 4 target frames per GPU
    (32 frames on 8 GPUs), the SAME scene on every GPU (/root/reference/src/model/model_wrapper_erp.py:412-432 renders
    the interpolated trajectory of one scene).  value: frames rendered from the HBM-resident scene.  e2e: the scene
    starts in pinned host memory on rank 0 ONLY, is uploaded once and broadcast over NVLink (NCCL), every rank renders
    its 4 frames and copies them to pinned host memory (where the reference's video writer takes them)."""
    import torch
    import torch.distributed as dist
    from splatter360_b200 import _lib, camera, parallel, synthetic
    from splatter360_b200 import rasterizer as R
    rank, world, local_rank, dev = _dist_setup(args)
    _lib.load()
    Hv, Wv, Pv, F = 1024, 2048, 3_000_000, 4
    K, Wm = args.steps, max(args.warmup, 3)
    sc = synthetic.random_cloud_scene(Pv, seed=1234 + 5, ref_width=2048, device=dev)
    means = sc.means.contiguous(); cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous()
    opac = sc.opacities.contiguous(); shs = sc.harmonics.permute(0, 2, 1).contiguous()
    n_frames = F * world
    poses = synthetic.trajectory(n_frames * (K + Wm), seed=0).to(dev)
    cams = camera.erp_camera(poses)
    bg = torch.zeros(3, device=dev)
    tracker = None if args.exact_counts else R.CapacityTracker()

    def settings(j):
        return R.GaussianRasterizationSettings(
            image_height=Hv, image_width=Wv, tanfovx=1.0, tanfovy=1.0, bg=bg, scale_modifier=1.0, viewmatrix=cams.view_matrix[j],
            projmatrix=cams.full_projection[j], sh_degree=SH_DEGREE, campos=cams.campos[j], prefiltered=False, debug=False,
            projection="erp", capacity_tracker=tracker)

    def my_frames(i):   # frames of step i rendered by this rank (round-robin over the ranks)
        return [i * n_frames + f for f in parallel.shard_views(n_frames, rank, world)]

    def step(i, scene=None):
        m, c, o, sh = scene if scene is not None else (means, cov6, opac, shs)
        return [R.forward_raw(settings(j), m, c, o, sh, None)[0] for j in my_frames(i)]

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    _lib.profile_read(reset=True)
    l0 = None

    def timed_step(i):
        nonlocal l0
        if i == Wm:
            _lib.profile_read(reset=True); _lib.profile_enable(True); l0 = _lib.launch_count()
        step(i)

    total_ms = _timed(timed_step, K, Wm, world, dev)
    _lib.profile_enable(False)
    launches = _lib.launch_count() - l0
    stages = _lib.profile_read(reset=True)
    if tracker is not None:
        tracker.flush()
        if tracker.overflowed:
            raise SystemExit(f"bench invalid: instance capacity overflowed at calls {tracker.overflowed}")
    clocks = sampler.stop() if sampler else None
    with torch.no_grad():
        _, st = R.forward_raw(settings(0)._replace(capacity_tracker=None), means, cov6, opac, shs, None)
        N, P_vis = st.num_rendered, st.num_visible
        del st

    # ---- e2e: rank 0 holds the scene in pinned host memory; upload once, NCCL broadcast, render, frames to pinned host
    e2e = None
    if not args.no_e2e:
        host = None
        if rank == 0:
            host = [t.detach().cpu().pin_memory() for t in (means, cov6, opac, shs)]
        bufs = [torch.empty_like(t) for t in (means, cov6, opac, shs)]
        out_host = [torch.empty((3, Hv, Wv), dtype=torch.float32).pin_memory() for _ in range(F)]
        e2e_tracker = None if args.exact_counts else R.CapacityTracker()

        copy_stream = torch.cuda.Stream(device=dev)
        d2h_stream = torch.cuda.Stream(device=dev)

        def e2e_step(i):
            # one upload on rank 0, chunk k+1 over PCIe while chunk k is broadcast over NVLink
            parallel.upload_and_broadcast_scene(host, bufs, src=0, copy_stream=copy_stream)
            cur = torch.cuda.current_stream()
            for h_, j in zip(out_host, my_frames(i)):
                frame = R.forward_raw(settings(j)._replace(capacity_tracker=e2e_tracker), bufs[0], bufs[1], bufs[2], bufs[3], None)[0]
                ev = torch.cuda.Event(); ev.record(cur)
                d2h_stream.wait_event(ev)                      # frame k goes to the host while frame k+1 renders
                with torch.cuda.stream(d2h_stream):
                    h_.copy_(frame, non_blocking=True)
                frame.record_stream(d2h_stream)
            d2h_stream.synchronize()                           # the frames are on the host when the step ends
            cur.wait_stream(d2h_stream)

        Ke = max(3, min(K, 10))
        e2e_ms = _timed(e2e_step, Ke, 2, world, dev) / Ke
        scene_bytes = sum(t.numel() * 4 for t in bufs)
        e2e = {"value": Pv * n_frames / (e2e_ms * 1e-3), "unit": "Gaussians/s", "ms_per_step": e2e_ms, "steps": Ke,
               "frames_per_s": n_frames / (e2e_ms * 1e-3),
               "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(n_frames * 3 * Hv * Wv * 4),
               "broadcast_bytes_per_rank": int(scene_bytes) if world > 1 else 0, "numa_bind": _NUMA["bind"],
               "api": "pinned host scene on rank 0 -> parallel.upload_and_broadcast_scene (one H2D upload in 64-MB chunks, each chunk "
                      "broadcast by NCCL over NVLink while the next one uploads) -> rasterizer.forward_raw per frame -> every frame to "
                      "pinned host on a copy stream while the next one renders; all inside the timed region"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_per_step = total_ms / K
    value = Pv * n_frames * K / (total_ms * 1e-3)
    peak, peak_src = load_peaks()
    stage_ms = {k: (v[0] / max(v[1], 1)) for k, v in stages.items() if v[1]}
    npix = Hv * Wv
    alg = {"preprocess": 340 * Pv + 48 * P_vis + 21 * Pv, "render_fwd": 52 * N + 20 * npix}
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    ach = alg.get(dom, 0) / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms.get(dom) else 0.0
    out = {"metric": "gaussians_per_s_fwd", "value": value, "unit": "Gaussians/s", "n_gpus": world, "steps": K, "warmup": Wm,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "frames_per_s": n_frames * K / (total_ms * 1e-3),
           "config": {"workload": "configs[4]: 3,000,000 random-cloud Gaussians (SH deg 4), 1024x2048 native-ERP video path, "
                                  "forward only, 4 target frames per GPU per step, scene replicated on every GPU",
                      "P": Pv, "image": [Hv, Wv], "frames_per_gpu": F, "sharding": "frames round-robin over the ranks; the scene is "
                      "broadcast once per step in the e2e leg; no collective in the render path",
                      "l2_policy": "the scene (1.02 GB) is larger than the 126 MB L2"},
           "e2e": e2e, "gpu_launches": int(launches) * world, "clocks": clocks,
           "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "traffic": None, "peak_source": peak_src, "stages_ms": stage_ms, "instances": {"P": Pv, "P_visible": P_vis, "N": N}},
           "cpu_baseline": None}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_config4(args):
    """BASELINE.json configs[3]: the HM3D evaluation shape -- per GPU one scene of 1,048,576 pixel-aligned Gaussians and one
    512x1024 target panorama rendered the REFERENCE's way: six 256x256 pinhole faces (DecoderSplattingCUDA, here in one
    batched pass) + Cube2Equirec, MSE loss on the panorama, backward to the Gaussians.  Weak scaling: N scenes on N GPUs."""
    import torch
    import torch.distributed as dist
    from splatter360_b200 import _lib, cubemap, synthetic
    from splatter360_b200.decoder import DecoderSplattingCUDA, Gaussians
    from splatter360_b200.loss import mse_loss
    from splatter360_b200.parallel import AsyncLossReducer
    rank, world, local_rank, dev = _dist_setup(args)
    _lib.load()
    K, Wm = args.steps, max(args.warmup, 3)
    Fw = H // 2
    sc = build_scene(dev, 1234 + 4 + 1000 * rank)
    P = sc.means.shape[0]
    g = Gaussians(sc.means[None].contiguous().requires_grad_(), sc.covariances[None].contiguous().requires_grad_(),
                  sc.harmonics[None].contiguous().requires_grad_(), sc.opacities[None].contiguous().requires_grad_())
    poses = synthetic.trajectory(K + Wm, seed=rank).to(dev)
    faces = cubemap.cube_face_extrinsics(poses)                       # [steps, 6, 4, 4]
    Kf = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev).expand(1, 6, 3, 3)
    near = torch.ones(1, 6, device=dev); far = torch.full((1, 6), 100.0, device=dev)
    dec = DecoderSplattingCUDA(sync_free=not args.exact_counts).to(dev)
    c2e = cubemap.Cube2Equirec(Fw, H, W).to(dev)
    target = torch.rand(1, 3, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank))
    reducer = AsyncLossReducer(dev)

    params = (g.means, g.covariances, g.harmonics, g.opacities)
    ext = faces[0][None].clone()                                       # static pose block of the captured step

    def loss_fn():
        out = dec(g, ext, Kf, near, far, (Fw, Fw))
        pano = c2e.from_faces(out.color)                               # [1, 3, H, W]
        return mse_loss(pano, target)

    def eager_step(i):
        for t in params:
            t.grad = None
        ext.copy_(faces[i][None], non_blocking=True)
        loss = loss_fn()
        loss.backward()
        reducer.submit(loss)

    graphed = None
    used_graph = bool(args.graph and not args.exact_counts)
    if used_graph:
        from splatter360_b200.graph import GraphedAutogradStep
        graphed = GraphedAutogradStep(loss_fn, params, trackers=dec.capacity_trackers)

        def step(i):
            ext.copy_(faces[i][None], non_blocking=True)
            reducer.submit(graphed.replay())
    else:
        step = eager_step

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = None

    def timed_step(i):
        nonlocal l0
        if i == Wm and graphed is None:
            _lib.profile_read(reset=True); _lib.profile_enable(True); l0 = _lib.launch_count()
        step(i)

    total_ms = _timed(timed_step, K, Wm, world, dev)
    reducer.flush()
    ovf = []
    if graphed is not None:
        # stage timers and the launch count cannot be taken inside a replayed graph: an eager pass right after it
        if graphed.overflowed():
            raise SystemExit("bench invalid: the graphed step's pair / instance capacity overflowed")
        graphed.release()
        graphed = None           # drops the captured autograd graph (its leaf accumulators belong to the capture stream)
        for t in params:
            t.grad = None
        n_e = min(K, 10)
        _lib.profile_read(reset=True); _lib.profile_enable(True); l0 = _lib.launch_count()
        for i in range(Wm, Wm + n_e):
            eager_step(i)
        torch.cuda.synchronize()
        reducer.flush()
        launches = (_lib.launch_count() - l0) * K // n_e
    else:
        launches = _lib.launch_count() - l0
    _lib.profile_enable(False)
    stages = _lib.profile_read(reset=True)
    if dec.capacity_trackers:
        for t in dec.capacity_trackers.values():
            t.flush(); ovf += t.overflowed
    if ovf:
        raise SystemExit(f"bench invalid: pair / instance capacity overflowed at calls {ovf}")
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_per_step = total_ms / K
    peak, peak_src = load_peaks()
    stage_ms = {k: (v[0] / max(v[1], 1)) for k, v in stages.items() if v[1]}
    out = {"metric": "gaussians_per_s_fwd_bwd", "value": P * world * K / (total_ms * 1e-3), "unit": "Gaussians/s", "n_gpus": world,
           "steps": K, "warmup": Wm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "views_per_s": world * K / (total_ms * 1e-3),
           "config": {"workload": "configs[3]: HM3D evaluation shape, per GPU one scene of 1,048,576 pixel-aligned Gaussians -> one "
                                  "512x1024 target panorama as six 256x256 pinhole faces (one batched pass) + Cube2Equirec, "
                                  "MSE loss, forward + backward", "P": P, "image": [H, W], "faces": [6, Fw, Fw],
                      "sharding": "one scene / target per GPU; NCCL all-reduce of the scalar loss only",
                      "api": "DecoderSplattingCUDA.forward (reference decoder contract) + cubemap.Cube2Equirec + loss.mse_loss"
                             + (", captured once by graph.GraphedAutogradStep and replayed (poses written in place)" if used_graph else ""),
                      "graph": used_graph,
                      "l2_policy": "inputs (369 MB/scene) and gradients are larger than the 126 MB L2"},
           "e2e": None, "gpu_launches": int(launches) * world, "clocks": clocks,
           "roofline": {"bound": "hbm", "kernel": max(stage_ms, key=lambda k: stage_ms[k]) if stage_ms else None, "achieved": None,
                        "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src, "stages_ms": stage_ms},
           "cpu_baseline": None, "final_loss": float(reducer.latest().item())}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def parity_block(oref, oref64, settings, means, cov6, opac, shs, dL):
    """GPU path vs the oracle on the bench workload itself (the last timed pose, all Gaussians; the float32 oracle run is
    the one that also gives `cpu_baseline`).  Per output three relative-L2 numbers: GPU vs float32 oracle, GPU vs the
    float64 build of the same oracle source, float32 oracle vs float64 oracle -- the last one is what float32 arithmetic
    costs on this workload, and no two float32 implementations can be asked to agree better than that.  ok: every output
    is within 1e-4 of the float32 oracle, or within 1e-4 of the float64 oracle, or no further from the float64 oracle than
    1.5x the float32 oracle is."""
    import numpy as np
    import torch
    from splatter360_b200 import rasterizer as R
    color, st = R.forward_raw(settings, means, cov6, opac, shs, None)
    g = R.backward_raw(settings, means, cov6, opac, shs, None, st, torch.from_numpy(dL).to(means.device))
    torch.cuda.synchronize()

    def rel(a, b):
        a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

    outs = {"color": (color, "color"), "d_means": (g["means3D"], "d_means"), "d_cov": (g["cov3D"], "d_cov6"),
            "d_opac": (g["opacities"].reshape(-1), "d_opac"), "d_shs": (g["shs"], "d_shs")}
    rd = np.abs(st.radii.cpu().numpy().astype(np.int64) - oref["radii"].astype(np.int64))
    out = {"against": "oracle/raster_oracle.c (float32) and its float64 build, on the bench workload (last timed pose, all Gaussians)",
           "metric": "relative L2: [GPU vs f32 oracle, GPU vs f64 oracle, f32 oracle vs f64 oracle]", "tolerance": 1e-4,
           "radii_mismatch_fraction": float((rd != 0).mean()), "radii_max_abs_diff": int(rd.max(initial=0))}
    ok = out["radii_mismatch_fraction"] <= 1e-5 and out["radii_max_abs_diff"] <= 1
    for k, (t, ok_) in outs.items():
        a = t.cpu().numpy()
        e32, e64, fl = rel(a, oref[ok_]), rel(a, oref64[ok_]), rel(oref[ok_], oref64[ok_])
        out[k] = [e32, e64, fl]
        ok = ok and (e32 < 1e-4 or e64 < 1e-4 or e64 <= 1.5 * fl)
    out["ok"] = bool(ok)
    return out


def cube6_run(dev, means, cov6, opac, shs, pose, iters=20):
    """Six cube faces of edge H/2 for the same scene and pose, forward+backward with an MSE seed on the stitched
    panorama: (a) six separate rasterizer calls, (b) one batched pass (s360_multi_*), both followed by the stitch
    kernel.  CUDA events, median of `iters`."""
    import torch
    from splatter360_b200 import camera, cubemap, rasterizer as R
    F = H // 2
    faces = cubemap.cube_face_extrinsics(pose)
    Kf = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev)[None].repeat(6, 1, 1)
    cam = camera.pinhole_camera(faces, Kf, torch.ones(6, device=dev), torch.full((6,), 100.0, device=dev))
    mk = lambda vm, pm, cp: R.GaussianRasterizationSettings(
        image_height=F, image_width=F, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev), scale_modifier=1.0,
        viewmatrix=vm.contiguous(), projmatrix=pm.contiguous(), sh_degree=SH_DEGREE, campos=cp.contiguous(),
        prefiltered=False, debug=False, projection="pinhole")
    s6 = mk(cam.view_matrix, cam.full_projection, cam.campos)
    s1 = [mk(cam.view_matrix[k], cam.full_projection[k], cam.campos[k]) for k in range(6)]
    c2e = cubemap.Cube2Equirec(F, H, W).to(dev)
    target = torch.rand(3, H, W, device=dev)
    info = {}

    def stitch_grad(faces_img):
        f = faces_img[None].detach().requires_grad_()
        pano = c2e.from_faces(f)
        (g,) = torch.autograd.grad(pano, f, 2 * (pano - target[None]) / pano.numel())
        return g[0]

    def batched():
        color, st = R.forward_views_raw(s6, means, cov6, opac, shs, None)
        info.update(pairs=st.num_pairs, instances=st.num_rendered)
        R.backward_views_raw(s6, means, cov6, opac, shs, None, st, stitch_grad(color))

    def separate():
        outs = [R.forward_raw(s1[k], means, cov6, opac, shs, None) for k in range(6)]
        g = stitch_grad(torch.stack([o[0] for o in outs]))
        for k in range(6):
            R.backward_raw(s1[k], means, cov6, opac, shs, None, outs[k][1], g[k])

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    t_b, t_s = timeit(batched), timeit(separate)
    P = means.shape[0]
    return {"what": "same scene/pose as the reference renders it: six 256x256 pinhole faces + Cube2Equirec, fwd+bwd (MSE on the panorama)",
            "batched_one_pass_ms": t_b, "six_separate_calls_ms": t_s, "speedup": t_s / t_b,
            "batched_gaussians_per_s": P / (t_b * 1e-3), "batched_views_per_s": 1e3 / t_b, **info}


def cpu_oracle_run(sc, pose, P_sample, repeats=1, dL=None, keep=False, f64=False):
    """Time the C oracle (OpenMP, all host threads) on P_sample Gaussians of the scene: fwd+bwd, one view.
    keep=True also returns the oracle's image and gradients (the parity check of the bench line)."""
    import numpy as np
    import torch
    import oracle
    from splatter360_b200 import camera, synthetic
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would handicap the CPU arm)
    oracle.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    means = sc.means.detach().cpu()
    n = means.shape[0]
    idx = torch.arange(n) if P_sample >= n else torch.randperm(n, generator=torch.Generator().manual_seed(0))[:P_sample]
    cam = camera.erp_camera(pose[None])
    m = means[idx].numpy()
    c6 = synthetic.cov3x3_to_cov6(sc.covariances.detach().cpu()[idx]).numpy()
    op = sc.opacities.detach().cpu()[idx].numpy()
    sh = sc.harmonics.detach().cpu()[idx].permute(0, 2, 1).contiguous().numpy()
    if dL is None:
        dL = np.random.default_rng(0).standard_normal((3, H, W)).astype(np.float32) / (3 * H * W)
    kw = dict(H=H, W=W, view=cam.view_matrix[0].numpy(), proj=cam.full_projection[0].numpy(),
              campos=cam.campos[0].numpy(), sh_degree=SH_DEGREE, mode="erp", dL_dpix=dL, stages=False)
    best, out = None, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = oracle.render(m, c6, op, shs=sh, f64=f64, **kw)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    res = {"value": len(idx) / best, "unit": "Gaussians/s", "cores": oracle.num_threads(), "kind": "port",
           "seconds": best,
           "sample": f"1 view fwd+bwd, {len(idx)} of the workload's {n} Gaussians at 512x1024 ERP, C oracle with OpenMP"}
    return (res, out) if keep else res


def run_reference(args):
    """CPU arm: the oracle port on the host cores (rank 0 only), on the SAME workload as the GPU arm -- every Gaussian of
    the view, same scene seed, same trajectory, same warm-up rule."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import oracle
    from splatter360_b200 import synthetic
    oracle.build()
    K, Wm = args.steps, max(args.warmup, 3)   # same warm-up rule as the GPU arm
    K = min(K, 40)                            # bounded: a step is ~1 s of CPU work on 16 threads
    Wm = min(Wm, 10)
    sc = build_scene("cpu", 1234 + CONFIG_ID)
    P = sc.means.shape[0]
    poses = synthetic.trajectory(K + Wm, seed=0)
    for i in range(Wm):
        cpu_oracle_run(sc, poses[i], P)
    t0 = time.perf_counter()
    last = None
    for i in range(K):
        last = cpu_oracle_run(sc, poses[Wm + i], P)
    dt = time.perf_counter() - t0
    value = P * K / dt
    sample = (f"each step = 1 view fwd+bwd over ALL {P} Gaussians of the workload at 512x1024 ERP (C oracle, OpenMP, "
              f"{last['cores']} threads); steps capped at 40, warm-up at 10")
    out = {
        "impl": "reference", "metric": "gaussians_per_s_fwd_bwd", "value": value, "unit": "Gaussians/s",
        "n_gpus": args.gpus, "steps": K, "warmup": Wm, "ms_per_step": dt / K * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "views_per_s": K / dt,
        "config": {"workload": WORKLOAD, "projection": "erp", "P": P, "image": [H, W], "sh_degree": SH_DEGREE,
                   "views_per_step_per_gpu": 1},
        "cpu_baseline": {"value": value, "unit": "Gaussians/s", "cores": last["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gaussians/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference has no CPU path and its CUDA rasterizer is an absent, un-pinned pip dependency; this arm is the oracle port (DESIGN.md)",
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="leave the process's CPU affinity alone (default: the CPUs of the GPU's NUMA node)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cube6", action="store_true")
    ap.add_argument("--exact-counts", action="store_true", help="read the instance count back every view (no CapacityTracker)")
    ap.add_argument("--graph", dest="graph", action="store_true", default=True,
                    help="config 3: time the step through graph.GraphedStep (one CUDA-graph launch per step; default)")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="configs 3 / 4: time the eager autograd calls instead of the CUDA-graph step")
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5],
                    help="BASELINE.json config: 3 (default, the headline line), 4 (reference-style six faces + stitch, one "
                         "scene per GPU), 5 (3M Gaussians, 1024x2048 video path, 4 frames per GPU, replicated scene)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 4:
        run_config4(args)
    elif args.config == 5:
        run_config5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

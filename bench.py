#!/usr/bin/env python
"""bench.py -- headline benchmark of the AD-GS hot path (BASELINE.json: fwd+bwd Mpix/s and
Gaussians/s, % of HBM roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one forward+backward of gaussian_renderer.render() per rank (trajectory at t and at
flow_time -> projection -> binning -> blend -> full backward to every parameter gradient), with
fixed seeded cotangents on the five output images so that no loss kernels are timed
(SURVEY.md section 8d). N > 1: launched by torchrun, one rank per GPU; each rank renders its own
view of an N-view batch and the parameter gradients are summed by one NCCL all-reduce over a flat
buffer inside the timed step (weak scaling).

The JSON line carries `value` (inputs resident in HBM), `e2e` (same metric through the public
API with the per-step host inputs copied from pinned memory and a device->host read of a
metric), `roofline` (dominant kernel, CUDA-event timed through the library's stage hooks),
`cpu_baseline` (reference trajectory+SH path on the host cores, rank 0, N=1 only), `clocks`,
`gpu_launches`.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: KITTI-MOT-shaped frame, 1M Gaussians (25 % objects, 32 control points)
    "kitti-375x1242-1M": dict(W=1242, H=375, n=1_000_000, obj_frac=0.25, median_radius_px=3.0),
    # BASELINE.json configs[2]: Waymo-shaped 3-camera rig (front-left / front / front-right, yaw -45 / 0 / +45 degrees,
    # scene/dataset_readers.py:261-357), 3M Gaussians filling the three frusta; one step = all three views
    "waymo-3cam-1066x1600-3M": dict(W=1600, H=1066, n=3_000_000, obj_frac=0.25, median_radius_px=3.0,
                                    rig_yaw=(-45.0, 0.0, 45.0)),
    # one camera of that rig
    "waymo-1066x1600-3M": dict(W=1600, H=1066, n=3_000_000, obj_frac=0.25, median_radius_px=3.0),
    # BASELINE.json configs[4]: stress -- 10M Gaussians, median radius 12 px, half of them inside 5 % of the screen
    "stress-1920x1280-10M": dict(W=1920, H=1280, n=10_000_000, obj_frac=0.25, median_radius_px=12.0, cluster=(0.5, 0.05)),
    "stress-1920x1280-2M": dict(W=1920, H=1280, n=2_000_000, obj_frac=0.25, median_radius_px=12.0, cluster=(0.5, 0.05)),
    # smaller variants for quick checks
    "kitti-375x1242-200k": dict(W=1242, H=375, n=200_000, obj_frac=0.25, median_radius_px=3.0),
    "tiny": dict(W=320, H=192, n=20_000, obj_frac=0.25, median_radius_px=3.0),
}
T_CAMERA, T_FLOW = 0.37, 0.41
METRIC = "fwd+bwd Mpix/s"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe). Samples NVML from a
    background thread every ~5 ms (nvidia-smi -lms cannot sample a 50 ms region); falls back to an
    nvidia-smi loop if NVML is not importable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = False

    def _nvml_loop(self, nv, handle):
        R = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
             "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
             "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
             "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(handle))
                for name, bit in R.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES if it is a plain index list
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                parts = [p.strip() for p in vis.split(",") if p.strip()]
                if self.gpu < len(parts) and parts[self.gpu].isdigit():
                    idx = int(parts[self.gpu])
            handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self._stop = False
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            sm = sorted(self.samples)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": sorted(self.reasons), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def algorithmic_bytes(n_scene, n_obj, R, px):
    """B_alg = N_s*1220 + N_o*2790 + R*404 + Px*84 per view, and its split by stage."""
    n = n_scene + n_obj
    stages = {
        "per_gaussian_forward": n_scene * 496 + n_obj * 1020,
        "binning": n * 24 + R * 172,                       # scan + emit + sort + ranges
        "blend_forward": R * 60 + px * 40,
        "blend_backward": R * 172 + px * 44,
        "per_gaussian_backward": n_scene * 700 + n_obj * 1748,
    }
    total = n_scene * 1220 + n_obj * 2790 + R * 404 + px * 84
    return total, stages


def scene_tensors(wl, device, seed=0):
    """The seeded synthetic scene of a workload in the REFERENCE's tensor layout (both arms build exactly this)."""
    import numpy as np
    from adgs_b200 import scenes
    n, n_obj = wl["n"], int(wl["n"] * wl["obj_frac"])
    n_scene = n - n_obj
    if "rig_yaw" in wl:
        # a third of the Gaussians in each camera's frustum, shuffled so that scene / object rows mix all three
        parts, k = [], len(wl["rig_yaw"])
        for i, yaw in enumerate(wl["rig_yaw"]):
            cam_i = scenes.make_camera(wl["W"], wl["H"], 90.0, yaw_deg=yaw, time=T_CAMERA)
            cnt = n // k + (1 if i < n % k else 0)
            parts.append(scenes.random_cloud(cnt, cam_i, seed=seed + 17 * i, median_radius_px=wl["median_radius_px"]))
        perm = np.random.default_rng(seed).permutation(n)
        cloud = {key: np.concatenate([p[key] for p in parts])[perm] for key in parts[0]}
    else:
        cam0 = scenes.make_camera(wl["W"], wl["H"], 90.0, time=T_CAMERA)
        cloud = scenes.random_cloud(n, cam0, seed=seed, median_radius_px=wl["median_radius_px"], cluster=wl.get("cluster"))
    tensors = scenes.random_model_tensors(n_scene, n_obj, scenes.BENCH_ORDER_ARGS, cloud, seed=seed + 1, device=device)
    return tensors, n_scene, n_obj


def build_ours(wl, device, seed=0):
    import torch
    from adgs_b200 import scenes
    from adgs_b200.gaussian_model import GaussianModel
    tensors, n_scene, n_obj = scene_tensors(wl, device, seed)
    model = GaussianModel.from_reference(tensors, scenes.BENCH_ORDER_ARGS, device=device)
    del tensors
    torch.cuda.empty_cache()
    return model, n_scene, n_obj


def views_of_step(wl, rank, device):
    """The (camera, t, flow_t) list one rank renders per step: its view of the multi-timestep batch, or -- for a
    camera rig -- every camera of the rig at this rank's timestep."""
    from adgs_b200 import scenes
    if "rig_yaw" in wl:
        t = T_CAMERA + 0.05 * rank
        return [(scenes.make_camera(wl["W"], wl["H"], 90.0, yaw_deg=yaw, time=t, device=device), t, t + (T_FLOW - T_CAMERA))
                for yaw in wl["rig_yaw"]]
    return [view_for_rank(wl, rank, device)]


def make_config(workload, wl, world, views_per_rank):
    """Identical in both arms (the driver compares it); arm-specific notes travel outside `config`."""
    return {"workload": workload, "views_per_step": world * views_per_rank, "image": [wl["H"], wl["W"]],
            "gaussians": wl["n"], "object_fraction": wl["obj_frac"], "control_points": 32, "bspline_order": 5,
            "fourier_terms": 6, "sh_degree": 3, "flow": True, "objmask": True, "inv_depth": True,
            "l2": "inputs (>1.5 GB of parameters per step) exceed the 126 MB L2",
            "parallelism": f"dp{world}: {views_per_rank} view(s) per rank and step, weak scaling"}


def view_for_rank(wl, rank, device):
    """8-view batch = 4 timesteps x 2 cameras; rank r renders view r."""
    from adgs_b200 import scenes
    yaw = 0.0 if rank % 2 == 0 else 8.0
    t = T_CAMERA + 0.05 * (rank // 2)
    cam = scenes.make_camera(wl["W"], wl["H"], 90.0, yaw_deg=yaw, time=t, device=device)
    return cam, t, t + (T_FLOW - T_CAMERA)


def make_cotangents(wl, device, seed=7, pinned=False):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    H, W = wl["H"], wl["W"]
    host = {k: torch.randn(c, H, W, generator=g) for k, c in (("color", 3), ("depth", 1), ("opacity", 1), ("flow", 3),
                                                             ("semantic", 1))}
    if pinned:
        return {k: v.pin_memory() for k, v in host.items()}
    return {k: v.to(device) for k, v in host.items()}


class InputStager:
    """Device-side staging of the per-step inputs of the e2e loop: a ring of preallocated device buffers, refilled every
    step from pinned host memory on a copy stream.

    A fresh device tensor per step -- `host.to(device)` on the copy stream -- works too (`ring=False`), but its block
    comes from the caching allocator's copy-stream pool, which grows by one cudaMalloc whenever the host gets one more
    step ahead than it had been before; a cudaMalloc costs ~14 ms on these boxes, i.e. +0.14 ms/step over a 100-step
    run for every one that lands inside it (profiles/r2_ab_e2e_ablation.txt).

    A slot is refilled only after the kernels that read it have finished: the copy stream waits for the event its last
    reader recorded (`release`), the host never does. `cuda` is torch.cuda (the CPU test passes a stand-in)."""

    def __init__(self, host_tensors, device, copy_stream, ring=True, slots=4, cuda=None):
        import torch
        self.cuda = cuda if cuda is not None else torch.cuda
        self.host = list(host_tensors)
        self.device, self.copy_stream, self.ring, self.slots = device, copy_stream, ring, slots
        self.buffers = [[torch.empty(h.shape, dtype=h.dtype, device=device) for h in self.host]
                        for _ in range(slots)] if ring else []
        self.slot_free = [None] * slots
        self.staged = 0

    def stage(self, skip=()):
        """Queue this step's copies, in the order of `host_tensors` (small camera block first: a tiny copy issued behind
        the 16.7 MB one would wait for it on the host-to-device copy engine and stall the forward).
        -> (slot, [device tensor or None for indices in `skip`], [event recorded behind each copy])."""
        k = self.staged % self.slots
        self.staged += 1
        out, events = [], []
        with self.cuda.stream(self.copy_stream):
            if self.ring and self.slot_free[k] is not None:
                self.copy_stream.wait_event(self.slot_free[k])
            for i, h in enumerate(self.host):
                if i in skip:
                    d = None
                elif self.ring:
                    d = self.buffers[k][i]
                    d.copy_(h, non_blocking=True)
                else:
                    d = h.to(self.device, non_blocking=True)
                ev = self.cuda.Event()
                ev.record(self.copy_stream)
                out.append(d)
                events.append(ev)
        return k, out, events

    def reads_on_current_stream(self, t):
        """Per-step tensors (ring=False): tell the allocator that the compute stream reads them."""
        if not self.ring and t is not None:
            t.record_stream(self.cuda.current_stream(self.device))

    def release(self, k):
        """Everything queued on the compute stream so far may read slot k."""
        if self.ring:
            ev = self.cuda.Event()
            ev.record(self.cuda.current_stream(self.device))
            self.slot_free[k] = ev


def run_ours(args):
    import torch
    import torch.distributed as dist
    from adgs_b200 import _lib as L
    from adgs_b200.gaussian_renderer import render
    from adgs_b200.parallel import MultiViewStep

    rank, world, local = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = L.load()
    wl = WORKLOADS[args.workload]
    model, n_scene, n_obj = build_ours(wl, device)
    if args.sync_free_outstanding is not None:
        model.sync_free_outstanding = args.sync_free_outstanding
    cam, t, flow_t = view_for_rank(wl, rank, device)
    cot = make_cotangents(wl, device)
    pipe = SimpleNamespace(inv_depth=True, debug=False, sync_free=True)
    flow_pkg = [flow_t, None, None, None, None, None]
    params = model.hot_parameters()
    px = wl["W"] * wl["H"]

    def outputs_and_cotangents(res, c):
        return ((res["render"], res["depth"], res["img_opacity"], res["img_flow"], res["img_semantic"]),
                (c["color"], c["depth"][0], c["opacity"][0], c["flow"], c["semantic"]))

    rig = "rig_yaw" in wl
    vpr = len(wl["rig_yaw"]) if rig else 1          # views per rank and step
    px = vpr * px
    if rig and world * vpr > 8:
        # one exchange round blends world x views-per-rank views (ADGS_MAX_VIEWS = 8 per round)
        raise SystemExit(f"workload {args.workload}: {vpr} views per rank x {world} ranks exceeds the 8 views of one "
                         f"exchange round; run it on at most {8 // vpr} GPU(s)")
    exchange = (world > 1 and args.parallel == "exchange") or rig
    mv = MultiViewStep(model) if (world > 1 and not exchange) else None
    ex = None
    if exchange:
        # Gaussian-sharded front/back end + exchange of splats (adgs_b200/parallel.py:SplatExchangeStep); a camera
        # rig on one GPU is the same machinery at world size 1 (all views of the step in the multi-view kernels)
        from adgs_b200.parallel import SplatExchangeStep
        shard = model.shard(rank, world) if world > 1 else model
        ex = SplatExchangeStep(shard)
        all_views = []
        for r in range(world):
            for c_r, t_r, f_r in views_of_step(wl, r, device):
                all_views.append((c_r, f_r))
        cot_dict = {"color": cot["color"], "depth": cot["depth"], "opacity": cot["opacity"], "flow": cot["flow"],
                    "semantic": cot["semantic"]}

    def step():
        if ex is not None:
            ex.run(all_views, lambda v, r: cot_dict, pipe, views_per_rank=vpr)
            return None
        if mv is None:
            for p in params:
                p.grad = None
            res = render(cam, model, None, pipe, flow_pkg=flow_pkg, render_objmask=True)
            outs, cots = outputs_and_cotangents(res, cot)
            torch.autograd.backward(outs, cots)
            return res
        views = [None] * world   # this rank's shard is exactly its own view
        views[rank] = (cam, flow_pkg)
        mv.run(views, lambda v: render(v[0], model, None, pipe, flow_pkg=v[1], render_objmask=True),
               lambda v, r: outputs_and_cotangents(r, cot), reduce_stats=False)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ---------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.adgs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (lib.adgs_launch_count() - launches0) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms], device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    mpix = world * px / (ms * 1e-3) / 1e6
    gauss = world * wl["n"] / (ms * 1e-3)

    # ---- roofline: per-stage CUDA events through the library hooks (rank 0) ------------------------
    roof, stage_ms, step_roof = None, None, None
    R = int(getattr(model, "_last_num_rendered", 0))
    if True:
        import ctypes as C
        ns = lib.adgs_profile_num_stages()
        ms_buf = (C.c_float * ns)()
        cnt_buf = (C.c_int32 * ns)()
        torch.cuda.synchronize()
        lib.adgs_profile_begin()
        prof_steps = max(3, min(args.steps, 10))
        for _ in range(prof_steps):
            step()      # every rank takes part (the step holds collectives when N > 1)
        torch.cuda.synchronize()
        lib.adgs_profile_end(ms_buf, cnt_buf)
        stage_ms = {lib.adgs_profile_stage_name(i).decode(): ms_buf[i] / prof_steps for i in range(ns)}
        R = int(getattr(model, "_last_num_rendered", R))
        if ex is not None:
            R = max(0, int((ex._capacity - 65536) / 1.3)) * vpr      # the arena bound tracks the largest single view
        total_b, per_stage = algorithmic_bytes(vpr * n_scene, vpr * n_obj, R, px)
        peak, peak_src = measured_hbm_peak()
        kernel_stages = {k: stage_ms[k] for k in ("per_gaussian_forward", "blend_forward", "blend_backward",
                                                   "per_gaussian_backward")}
        top = max(kernel_stages, key=kernel_stages.get)
        ach = per_stage[top] / (kernel_stages[top] * 1e-3) / 1e9
        # per-launch DRAM bytes and warp instructions of each kernel from the committed `ncu --set full` capture of this
        # same workload (profiles/traffic.json); they scale with the views of a step
        captured = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                captured = json.load(open(tpath))
            except Exception:
                captured = {}
        same_wl = captured.get("_workload") == args.workload
        sm_mhz = 1965.0
        try:
            sm_mhz = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"])
        except Exception:
            pass
        issue_peak = 148 * 4 * sm_mhz * 1e6          # warp instructions / s: 4 schedulers per SM, one issue per clock

        def kernel_roofline(k):
            c = captured.get(k) if same_wl else None
            t = kernel_stages[k] * 1e-3
            out = {"ms": round(kernel_stages[k], 4), "algorithmic_bytes": per_stage[k],
                   "hbm_frac_algorithmic": round(per_stage[k] / t / 1e9 / peak, 4)}
            if isinstance(c, dict):
                if c.get("dram_bytes"):
                    out["dram_bytes"] = c["dram_bytes"]
                    out["hbm_frac_measured_traffic"] = round(c["dram_bytes"] / t / 1e9 / peak, 4)
                if c.get("warp_instructions"):
                    out["warp_instructions"] = c["warp_instructions"]
                    out["issue_min_ms"] = round(c["warp_instructions"] / issue_peak * 1e3, 4)
                    out["issue_frac"] = round(c["warp_instructions"] / issue_peak / t, 4)
            return out

        per_kernel = {k: kernel_roofline(k) for k in kernel_stages}
        traffic = per_kernel[top].get("dram_bytes")
        roof = {"bound": "hbm", "kernel": top, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": per_stage[top], "kernel_ms": round(kernel_stages[top], 4),
                "issue": {"peak_warp_instructions_per_s": issue_peak, "sm_mhz": sm_mhz,
                          "warp_instructions": per_kernel[top].get("warp_instructions"),
                          "min_ms_at_one_issue_per_clock": per_kernel[top].get("issue_min_ms"),
                          "frac": per_kernel[top].get("issue_frac"),
                          "note": "the blend kernels are bound by issue slots, not by HBM (DESIGN.md section 4): "
                                  "frac = warp instructions (ncu capture in profiles/) / (148 SMs x 4 schedulers x clock) "
                                  "/ measured kernel time"},
                "per_kernel": per_kernel,
                "note": "reported against HBM per BASELINE; `issue` is the bound that actually binds the blend kernels"}
        step_ach = total_b / (ms * 1e-3) / 1e9
        step_roof = {"algorithmic_bytes_per_view": total_b, "achieved_GBps": round(step_ach, 1),
                     "frac_of_hbm_peak": round(step_ach / peak, 4), "num_rendered": R}

    # ---- e2e: host buffers in, host metric out, every step ---------------------------------------------
    # host-side inputs of one step, each group packed in ONE pinned buffer (one H2D copy per group):
    # the five cotangent planes (standing in for the ground-truth images the losses consume) and
    # the camera (view matrix, full projection, centre).
    H, W = wl["H"], wl["W"]
    hc = make_cotangents(wl, device, pinned=False)
    host_cot = torch.cat([hc[k].cpu().reshape(-1) for k in ("color", "depth", "opacity", "flow", "semantic")]).pin_memory()
    cam_e2e = all_views[rank * vpr][0] if ex is not None else cam
    host_cam = torch.cat([cam_e2e.world_view_transform.cpu().reshape(-1), cam_e2e.full_proj_transform.cpu().reshape(-1),
                          cam_e2e.camera_center.cpu().reshape(-1)]).pin_memory()
    h2d = host_cot.numel() * 4 + host_cam.numel() * 4
    d2h = 4

    def split_cot(flat):
        px_ = H * W
        o, out = 0, {}
        for k, ch in (("color", 3), ("depth", 1), ("opacity", 1), ("flow", 3), ("semantic", 1)):
            out[k] = flat[o:o + ch * px_].view(ch, H, W)
            o += ch * px_
        return out

    copy_stream = torch.cuda.Stream(device=device)
    readback_stream = torch.cuda.Stream(device=device)
    static_flat = host_cot.to(device) if args.e2e_ablate == "cot" else None
    # the host-to-device rate of this box for exactly this transfer, alone (into one preallocated buffer, so that no
    # allocation sits between the copies): h2d_bytes / rate is a floor of the e2e step whatever the kernels do -- the
    # copy of step i+1 runs under step i, but one copy per step has to fit into a step
    h2d_ms = None
    try:
        probe_dst = torch.empty(host_cot.shape, dtype=host_cot.dtype, device=device)
        with torch.cuda.stream(copy_stream):
            for _ in range(2):
                probe_dst.copy_(host_cot, non_blocking=True)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(copy_stream)
            for _ in range(8):
                probe_dst.copy_(host_cot, non_blocking=True)
            c1.record(copy_stream)
        c1.synchronize()
        h2d_ms = c0.elapsed_time(c1) / 8
        del probe_dst
    except Exception as exc:   # the probe must never take the bench line down with it
        sys.stderr.write(f"h2d probe failed: {exc}\n")

    # per-step inputs: camera block first, cotangent planes behind it, through the copy stream (InputStager above);
    # --e2e-staging alloc keeps the fresh-tensor-per-step behaviour for comparison
    stager = InputStager([host_cam, host_cot], device, copy_stream, ring=(args.e2e_staging == "ring"))

    def stage_inputs():
        """-> (slot, camera on the device, its event, cotangents on the device, their event)"""
        k, (dcam, flat), (cam_ready, ready) = stager.stage(skip=(1,) if args.e2e_ablate == "cot" else ())
        return k, dcam, cam_ready, (static_flat if flat is None else flat), ready

    use_on_compute_stream = stager.reads_on_current_stream
    release_inputs = stager.release

    # device -> host read of the step's metric: an asynchronous copy into pinned memory that the host
    # consumes two steps later (after the next two steps have been queued), the way a training loop logs its
    # loss without draining the GPU. Every step's value is read inside the timed region; the last one
    # is drained before the closing event. --e2e-blocking restores the read-then-launch order.
    metric_ring = [torch.empty((1,), dtype=torch.float32).pin_memory() for _ in range(4)]
    metric_pending = []
    metric_log = []

    host_wait = [0.0]   # seconds the host spent waiting for the GPU (everything else in the loop is enqueue work)

    def consume_metric():
        buf, ev = metric_pending.pop(0)
        w0 = time.perf_counter()
        ev.synchronize()
        host_wait[0] += time.perf_counter() - w0
        metric_log.append(float(buf[0].item()))

    def read_metric(value_on_device):
        if args.e2e_blocking:
            metric_log.append(float(value_on_device.detach().item()))
            return
        buf = metric_ring[(len(metric_log) + len(metric_pending)) % 4]   # at most three reads in flight
        # the 4-byte device-to-host copy runs on a side stream: in the compute stream it would sit between this
        # step's last kernel and the next step's first one
        val = value_on_device.detach().reshape(1)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(device))
        with torch.cuda.stream(readback_stream):
            readback_stream.wait_event(done)
            buf.copy_(val, non_blocking=True)
            val.record_stream(readback_stream)
            ev = torch.cuda.Event()
            ev.record(readback_stream)
        metric_pending.append((buf, ev))
        if len(metric_pending) > 2:
            consume_metric()

    def drain_metrics():
        while metric_pending:
            consume_metric()

    def e2e_step_exchange():
        slot, dcam, cam_ready, flat, ready = stage_inputs()
        torch.cuda.current_stream(device).wait_event(cam_ready)
        use_on_compute_stream(dcam)
        views_ = list(all_views)
        views_[rank * vpr] = (all_views[rank * vpr][0]._replace(world_view_transform=dcam[0:16].view(4, 4),
                                                                full_proj_transform=dcam[16:32].view(4, 4),
                                                                camera_center=dcam[32:35]), all_views[rank * vpr][1])
        box = {}

        def cots(v, r):
            torch.cuda.current_stream(device).wait_event(ready)
            use_on_compute_stream(flat)
            box["res"] = r
            return split_cot(flat)

        ex.run(views_, cots, pipe, views_per_rank=vpr)
        release_inputs(slot)
        read_metric(box["res"]["img_opacity"].mean())

    def e2e_step():
        if ex is not None:
            return e2e_step_exchange()
        # host -> device: camera matrices first (needed by the forward), cotangent planes behind them on the copy
        # stream so that the PCIe transfer overlaps the forward; the backward waits for them (stage_inputs).
        ablate = args.e2e_ablate      # diagnosis only (the reported e2e runs with "none")
        slot, dcam, cam_ready, flat, ready = stage_inputs()
        torch.cuda.current_stream(device).wait_event(cam_ready)
        use_on_compute_stream(dcam)
        vc = cam if ablate == "cam" else cam._replace(world_view_transform=dcam[0:16].view(4, 4),
                                                      full_proj_transform=dcam[16:32].view(4, 4),
                                                      camera_center=dcam[32:35])
        c = split_cot(flat)
        for p in params:
            p.grad = None
        model._grad_sink = mv.bucket.views if mv is not None else None
        try:
            res = render(vc, model, None, pipe, flow_pkg=flow_pkg, render_objmask=True)
            torch.cuda.current_stream(device).wait_event(ready)
            if flat is not static_flat:
                use_on_compute_stream(flat)
            outs, cots = outputs_and_cotangents(res, c)
            torch.autograd.backward(outs, cots)
        finally:
            model._grad_sink = None
        release_inputs(slot)
        if mv is not None:
            mv.bucket.all_reduce()
        if ablate == "metric":
            metric_log.append(0.0)
        else:
            read_metric(res["img_opacity"].mean())   # device -> host read of a metric

    e2e_warm = max(3, args.warmup)
    for _ in range(e2e_warm):
        e2e_step()
    drain_metrics()
    barrier()
    host_wait[0] = -model.__dict__.get("_host_wait_s", 0.0)   # + the waits inside render() (binning counters)
    h0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    drain_metrics()
    e1.record()
    host_wait[0] += model.__dict__.get("_host_wait_s", 0.0)
    host_ms = (time.perf_counter() - h0 - host_wait[0]) * 1e3 / args.steps
    barrier()
    assert len(metric_log) == args.steps + e2e_warm and all(math.isfinite(v) for v in metric_log)
    ms_e2e = e0.elapsed_time(e1) / args.steps
    if world > 1:
        tt = torch.tensor([ms_e2e], device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = float(tt.item())
    e2e = {"value": round(world * px / (ms_e2e * 1e-3) / 1e6, 2), "unit": "Mpix/s", "ms_per_step": round(ms_e2e, 4),
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           # python + ctypes + launch time per step with the waits on the GPU taken out: when it approaches
           # ms_per_step the loop is bound by the host, not by the device
           "host_enqueue_ms_per_step": round(host_ms, 4),
           "staging": ("ring of 4 preallocated device buffers refilled from pinned memory every step" if stager.ring
                       else "fresh device tensor per step"),
           "h2d_alone": None if not h2d_ms else {
               "ms_per_step": round(h2d_ms, 4), "GBps": round(host_cot.numel() * 4 / (h2d_ms * 1e-3) / 1e9, 2),
               "note": "this step's pinned-memory transfer timed on its own (idle GPU): the PCIe floor of the e2e "
                       "step on this box"},
           **({"INVALID_ablated": args.e2e_ablate} if args.e2e_ablate != "none" else {}),
           "readback": "blocking .item() every step" if args.e2e_blocking else
                       "async copy to pinned memory every step, consumed two steps later; drained inside the timed region"}

    # ---- the step right after the path: fused Adam over all 18 parameter groups, timed on
    #      its own (NOT part of `value`): 28 B per parameter element (p, g, m, v read; p, m, v written) -------
    adam = None
    if ex is None and mv is None:
        from adgs_b200.optimizer import FusedAdam, GROUP_NAMES
        res = step()
        opt = FusedAdam(model, {n: 1e-4 for n in GROUP_NAMES}, eps=1e-15)
        n_el = sum(p.numel() for p in params)
        for _ in range(3):
            opt.step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            opt.step()
        e1.record()
        torch.cuda.synchronize()
        ms_adam = e0.elapsed_time(e1) / 10
        peak, _src = measured_hbm_peak()
        adam = {"ms_per_step": round(ms_adam, 4), "elements": n_el, "algorithmic_bytes": n_el * 28,
                "achieved_GBps": round(n_el * 28 / (ms_adam * 1e-3) / 1e9, 1),
                "frac_of_hbm_peak": round(n_el * 28 / (ms_adam * 1e-3) / 1e9 / peak, 4),
                "note": "adgs_adam_step, one launch for the reference's 18 groups; not included in value/e2e"}
        del opt, res

    # ---- the step right before the backward: fused L1 + SSIM image loss, forward + backward,
    #      timed on its own against the torch composition the reference runs (utils/loss_utils.py:20-58) ----------
    loss_fe = None
    if ex is None and mv is None:
        from adgs_b200 import losses as LS
        gt_img = torch.rand(3, wl["H"], wl["W"], device=device)
        pred_img = (gt_img + 0.1 * torch.randn_like(gt_img)).requires_grad_(True)

        C3, Hh, Ww = 3, wl["H"], wl["W"]
        planes = torch.empty((3, C3, Hh, Ww), device=device)
        partial = torch.empty((lib.adgs_image_loss_partial_floats(C3, Hh, Ww),), device=device)
        out3 = torch.empty((3,), device=device)
        gone = torch.ones((1,), device=device)
        d_img = torch.empty_like(gt_img)
        stream_h = torch.cuda.current_stream(device).cuda_stream
        pimg = pred_img.detach()

        def fused_loss():
            # the two C-ABI calls adgs_b200.losses.image_loss makes (forward incl. the derivative planes, backward)
            lib.adgs_image_loss_forward(C3, Hh, Ww, pimg.data_ptr(), gt_img.data_ptr(), planes[0].data_ptr(),
                                        planes[1].data_ptr(), planes[2].data_ptr(), partial.data_ptr(), 0.8, 0.2,
                                        out3.data_ptr(), stream_h)
            lib.adgs_image_loss_backward(C3, Hh, Ww, pimg.data_ptr(), gt_img.data_ptr(), planes[0].data_ptr(),
                                         planes[1].data_ptr(), planes[2].data_ptr(), gone.data_ptr(), 0.8,
                                         gone.data_ptr(), -0.2, d_img.data_ptr(), stream_h)

        def time_it(fn, n=50):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        ms_fused = time_it(fused_loss)
        peak, _src = measured_hbm_peak()
        nbytes = 3 * px * 4 * (2 + 3 + 5 + 1)   # fwd: img, gt read, 3 planes written; bwd: 5 planes read, 1 written
        loss_fe = {"ms_fwd_bwd": round(ms_fused, 4), "algorithmic_bytes": nbytes,
                   "achieved_GBps": round(nbytes / (ms_fused * 1e-3) / 1e9, 1),
                   "note": "adgs_image_loss_forward/backward: (1-l)*L1 + l*(1-SSIM) on (3,H,W); not included in value/e2e"}

    # ---- a whole training iteration on the native pieces (render -> fused losses -> backward -> fused Adam,
    #      adgs_b200/train_step.py = the inner part of train.py:74-167); runs last: it moves the parameters -----
    train_it = None
    if ex is None and mv is None:
        from adgs_b200.train_step import training_iteration
        targs = SimpleNamespace(percent_dense=0.01, object_extent=10.0, min_camera_extent=10.0, feature_lr=0.0025,
                                opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001, rotation_deform_lr=0.001,
                                shs_deform_lr=0.0025, gs_time_sigma_lr=1e-2, position_lr_init=0.00016,
                                position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=60_000,
                                position_deform_lr_scale=0.2, obj_position_lr_scale=0.8, scene_position_lr_scale=1.0)
        topt = SimpleNamespace(lambda_dssim=0.2, lambda_l1=1.0, lambda_depth=0.1, lambda_flow=0.0, lambda_obj=0.1,
                               lambda_sky=0.05, lambda_sigma=0.01)
        Hh, Ww = wl["H"], wl["W"]
        view = SimpleNamespace(image_height=Hh, image_width=Ww, FoVx=cam.FoVx, FoVy=cam.FoVy,
                               world_view_transform=cam.world_view_transform, full_proj_transform=cam.full_proj_transform,
                               camera_center=cam.camera_center, time=t, original_image=torch.rand(3, Hh, Ww, device=device),
                               depth=torch.rand(Hh, Ww, device=device) + 0.1,
                               semantic=(torch.rand(Hh, Ww, device=device) > 0.7).float(),
                               sky=(torch.rand(Hh, Ww, device=device) > 0.8).float())
        for p in params:
            p.grad = None
        per_mode = {}
        for mode in ("dense", "window_aware"):
            model.scene_extent = 20.0
            model.training_setup(targs, window_aware=(mode == "window_aware"))
            model.reset_active_columns()
            for it in range(1, 11):
                training_iteration(model, view, topt, pipe, it, frame_gap=1.0 / 96)
            best = None
            for rep in range(3):   # best of three 20-iteration windows: the first window still sees allocator growth
                torch.cuda.synchronize()
                e0.record()
                for it in range(11 + 20 * rep, 31 + 20 * rep):
                    training_iteration(model, view, topt, pipe, it, frame_gap=1.0 / 96)
                e1.record()
                torch.cuda.synchronize()
                ms_w = e0.elapsed_time(e1) / 20
                best = ms_w if best is None else min(best, ms_w)
            per_mode[mode] = round(best, 4)
            model.optimizer = None
        train_it = {"ms_per_iteration": per_mode, "note": "render + L1/SSIM/depth/obj/sky losses + backward + fused Adam "
                    "(18 groups), one view per iteration like train.py; not part of value/e2e"}

    # ---- densification: densify_and_prune of the whole model incl. Adam moments, and the
    #      near-index K-NN, each timed on its own with synthetic statistics; runs last: it changes the model --------
    densify = None

    def densify_section():
        from adgs_b200 import densify as DN
        from adgs_b200.gaussian_model import PARAM_NAMES as _PN
        gen = torch.Generator(device=device).manual_seed(7)
        ms_runs, rows = [], None
        model.optimizer = None
        # every run starts from the same 1 M-Gaussian parameters (a copy of the bench model with fresh Adam state):
        # the first run grows the caching allocator, the later ones are the steady-state cost inside a training run
        base = {k: getattr(model, k).detach().clone() for k in _PN}
        base_time, base_ns, base_no = model.gs_time.clone(), model.n_scene, model.n_obj
        for rep in range(3):
            for k in _PN:
                setattr(model, k, torch.nn.Parameter(base[k].clone()))
            model.gs_time, model.n_scene, model.n_obj = base_time.clone(), base_ns, base_no
            model.scene_extent, model.object_extent, model.percent_dense = 20.0, 5.0, 0.01
            model.training_setup(targs, window_aware=False)
            model.object_extent = 5.0
            model.denom = torch.randint(0, 4, (model.get_pts_num, 1), generator=gen, device=device).float()
            model.xyz_gradient_accum = model.denom * 0.0002 * torch.exp(torch.randn((model.get_pts_num, 1), generator=gen,
                                                                                    device=device))
            n_before = model.get_pts_num
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            model.densify_and_prune(0.0002, 0.0002, 0.005, False)
            torch.cuda.synchronize()
            ms_runs.append((time.perf_counter() - t0) * 1e3)
            rows = (n_before, model.get_pts_num)
        del base
        per_g = 380 + 1052 * wl["obj_frac"]
        pts4 = torch.cat([model.xyz.detach()[model.n_scene:], model.gs_time.reshape(-1, 1) * 20.0], dim=-1).contiguous()
        anchors = pts4[torch.randperm(pts4.shape[0], device=device)[:pts4.shape[0] // 8]].contiguous()
        DN.knn_points(anchors, pts4, 8)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            DN.knn_points(anchors, pts4, 8)
        e1.record()
        torch.cuda.synchronize()
        ms_knn = e0.elapsed_time(e1) / 3
        return {"densify_and_prune_ms": round(min(ms_runs[1:]), 3), "first_call_ms": round(ms_runs[0], 3),
                   "rows_before_after": rows,
                   "algorithmic_bytes": int(3 * per_g * (rows[0] + rows[1])),
                   "knn_points_ms": round(ms_knn, 3), "knn_anchors_points_K": [anchors.shape[0], pts4.shape[0], 8],
                   "knn_pairs_per_s": round(anchors.shape[0] * pts4.shape[0] / (ms_knn * 1e-3), 1),
                   "note": "host wall time incl. allocation of the new arrays and the one read-back of the row counts; "
                           "parameters + both Adam moments gathered in one launch; not part of value/e2e"}

    if ex is None and mv is None:
        try:
            densify = densify_section()
        except Exception as exc:   # an auxiliary measurement must never take the headline line down with it
            densify = {"error": f"{type(exc).__name__}: {exc}"}
        model.optimizer = None

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline()

    # ---- the other BASELINE configs that fit one GPU, measured by the same harness (short runs; the headline stays
    #      `value` above): configs[2] as a real 3-camera step, configs[4] the 10 M-Gaussian stress frame --------------
    others = None
    if world == 1 and args.workload == "kitti-375x1242-1M" and not args.no_other_workloads:
        del model, params
        torch.cuda.empty_cache()
        others = {}
        for name in ("waymo-3cam-1066x1600-3M", "stress-1920x1280-10M"):
            try:
                others[name] = measure_workload(name, device, steps=10, warmup=3)
            except Exception as exc:
                others[name] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()

    if ex is not None and ex.timing_report():
        print(f"rank {rank} exchange phases (ms):", ex.timing_report(), file=sys.stderr)
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(mpix, 2), "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args.workload, wl, world, vpr),
            "parallelism_note": ((f"{world} rank(s); Gaussians sharded, every view blended on one rank; splats and 64-byte "
                                  f"gradient records exchanged through {ex.exchange if world > 1 else 'local'} memory "
                                  "(no gradient all-reduce)") if exchange else
                                 f"views sharded over {world} rank(s), flat-buffer NCCL all-reduce of parameter gradients"),
            "gaussians_per_s": round(gauss, 1),
            "e2e": e2e, "gpu_launches": round(launches, 1), "clocks": clocks,
            "roofline": roof, "step_roofline": step_roof, "stage_ms": {k: round(v, 4) for k, v in (stage_ms or {}).items()},
            "cpu_baseline": cpu_base, "optimizer_step": adam, "loss_front_end": loss_fe, "train_iteration": train_it,
            "densification": densify, "other_workloads": others,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_workload(name, device, steps=10, warmup=3):
    """One workload on one GPU through the public API: value (inputs resident in HBM), per-stage CUDA-event times,
    whole-step algorithmic-byte roofline. A camera rig runs all its cameras per step (multi-view kernels)."""
    import ctypes as C
    import torch
    from adgs_b200 import _lib as L
    from adgs_b200.gaussian_renderer import render
    lib = L.load()
    wl = WORKLOADS[name]
    model, n_scene, n_obj = build_ours(wl, device)
    views = views_of_step(wl, 0, device)
    vpr = len(views)
    cot = make_cotangents(wl, device)
    pipe = SimpleNamespace(inv_depth=True, debug=False, sync_free=True)
    params = model.hot_parameters()
    px = vpr * wl["W"] * wl["H"]
    ex = None
    if vpr > 1:
        from adgs_b200.parallel import SplatExchangeStep
        ex = SplatExchangeStep(model)
        ex_views = [(c, f) for c, t, f in views]

    def step():
        if ex is not None:
            ex.run(ex_views, lambda v, r: cot, pipe, views_per_rank=vpr)
            return
        for p in params:
            p.grad = None
        cam, t, flow_t = views[0]
        res = render(cam, model, None, pipe, flow_pkg=[flow_t, None, None, None, None, None], render_objmask=True)
        torch.autograd.backward((res["render"], res["depth"], res["img_opacity"], res["img_flow"], res["img_semantic"]),
                                (cot["color"], cot["depth"][0], cot["opacity"][0], cot["flow"], cot["semantic"]))

    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ns = lib.adgs_profile_num_stages()
    ms_buf, cnt_buf = (C.c_float * ns)(), (C.c_int32 * ns)()
    lib.adgs_profile_begin()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    lib.adgs_profile_end(ms_buf, cnt_buf)
    stage_ms = {lib.adgs_profile_stage_name(i).decode(): round(ms_buf[i] / 3, 4) for i in range(ns)}
    if ex is not None:
        R = max(0, int((ex._capacity - 65536) / 1.3)) * vpr
    else:
        R = int(getattr(model, "_last_num_rendered", 0))
    total_b, _ = algorithmic_bytes(vpr * n_scene, vpr * n_obj, R, px)
    peak, _src = measured_hbm_peak()
    ach = total_b / (ms * 1e-3) / 1e9
    return {"metric": METRIC, "value": round(px / (ms * 1e-3) / 1e6, 2), "unit": "Mpix/s", "ms_per_step": round(ms, 4),
            "steps": steps, "config": make_config(name, wl, 1, vpr),
            "gaussians_per_s": round(vpr * wl["n"] / (ms * 1e-3), 1), "num_rendered": R, "stage_ms": stage_ms,
            "step_roofline": {"algorithmic_bytes_per_step": total_b, "achieved_GBps": round(ach, 1),
                              "frac_of_hbm_peak": round(ach / peak, 4)}}


def cpu_baseline(budget_s=15.0):
    """BASELINE config 1: the reference's B-spline trajectory + SH evaluation (no rasterizer) on CPU
    torch, 100k Gaussians (all objects), 32 control points, forward+backward, all host threads.
    This is the oracle port (oracle/trajectory_oracle.py) being TIMED AS A BASELINE, nothing more."""
    import torch
    from oracle import trajectory_oracle as TO
    from adgs_b200 import scenes
    n = 100_000
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = TO.random_reference_model(0, n, scenes.BENCH_ORDER_ARGS, seed=0, device="cpu", requires_grad=True)
    campos = torch.zeros(3)
    params = [getattr(ref, f) for f in ref.trainable()]

    def it():
        for p in params:
            p.grad = None
        flow = ref.get_deformed_xyz(T_FLOW)
        pkg = ref.get_deformed_pkg(T_CAMERA)
        rgb = TO.sh_colors(pkg["shs"], pkg["xyz"], campos, 3)
        (rgb.sum() + flow.sum() + pkg["opacity"].sum() + pkg["rotation"].sum() + ref.get_scaling().sum()).backward()

    it()
    t0 = time.perf_counter()
    k = 0
    while True:
        it()
        k += 1
        if time.perf_counter() - t0 > budget_s or k >= 50:
            break
    dt = (time.perf_counter() - t0) / k
    return {"value": round(n / dt, 1), "unit": "Gaussians/s (trajectory + SH colour, fwd+bwd, no rasterizer)",
            "cores": cores, "kind": "port", "ms_per_iter": round(dt * 1e3, 2),
            "sample": f"BASELINE configs[0]: {n} object Gaussians, 32 control points, k=5, F=6, SH degree 3; {k} iterations"}


def run_reference(args):
    """Reference arm: the reference's own implementation of the path -- its UNMODIFIED CUDA rasterizer (oracle/_ref,
    compiled from /root/reference by oracle/Makefile) driven by the torch trajectory exactly as
    gaussian_renderer.render() drives it (oracle/ref_pipeline.py) -- on the same workload, metric and timing harness.
    The reference has no CPU rasterizer, so its implementation of the path IS this GPU one; it is the number to beat.
    N > 1 (B-REF-N of BASELINE.md): the reference has no distributed code, so N ranks each hold a full replica, render
    their own view of the batch and sum the parameter gradients with a plain torch.distributed.all_reduce -- the
    'naive data parallel' curve. If neither a GPU nor oracle/_ref is available, the numpy oracle port is timed on a
    bounded sample instead."""
    rank, world, local = dist_env()
    import torch
    from oracle import ref_module as REF
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    if not (torch.cuda.is_available() and REF.available()):
        if rank == 0:
            print(json.dumps(reference_port_line(args, cores)))
        return
    import torch.distributed as dist
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    res = reference_measure(args.workload, rank, world, device, args.steps, max(args.warmup, 3), stage_split=True)
    others = None
    if world == 1 and args.workload == "kitti-375x1242-1M" and not args.no_other_workloads:
        others = {}
        for name in ("waymo-3cam-1066x1600-3M", "stress-1920x1280-10M"):
            try:
                r = reference_measure(name, 0, 1, device, 3, 3, stage_split=False)
                others[name] = {"metric": METRIC, "value": r["value"], "unit": "Mpix/s", "ms_per_step": r["ms"],
                                "steps": 3, "config": make_config(name, WORKLOADS[name], 1, r["vpr"])}
            except Exception as exc:
                others[name] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()
    blocking = None
    if rank == 0 and world == 1 and not args.launch_blocking_child and args.workload == "kitti-375x1242-1M":
        # train.py:277 sets CUDA_LAUNCH_BLOCKING=1 for training: how the reference actually runs. Must be set before
        # CUDA initialises, hence a child process (5 steps).
        try:
            env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "5",
                                  "--warmup", "3", "--workload", args.workload, "--launch-blocking-child",
                                  "--no-other-workloads"], env=env, capture_output=True, text=True, timeout=300)
            blocking = json.loads(out.stdout.strip().splitlines()[-1])["ms_per_step"]
        except Exception as exc:
            blocking = f"{type(exc).__name__}: {exc}"
    if rank == 0:
        value = res["value"]
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": res["ms"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": make_config(args.workload, wl, world, res["vpr"]),
                "parallelism_note": ("full replica per rank, one view each, torch.distributed.all_reduce of every parameter "
                                     "gradient (the reference itself is single-GPU: train.py:55-61)" if world > 1 else
                                     "single GPU, one view per iteration like train.py"),
                "gaussians_per_s": round(world * res["vpr"] * wl["n"] / (res["ms"] * 1e-3), 1),
                "stage_ms": res["stages"], "ms_per_step_with_CUDA_LAUNCH_BLOCKING": blocking,
                "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": "reference",
                                 "sample": "unmodified reference CUDA rasterizer (oracle/_ref) + torch trajectory on the "
                                           "GPU, full workload; runs on the GPU because the reference has no CPU rasterizer"},
                "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "other_workloads": others}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def reference_measure(workload, rank, world, device, steps, warmup, stage_split):
    """ms/step of the reference pipeline for `workload` on this rank's views (max over ranks), and -- on request -- a
    per-stage split measured with CUDA events in a separate, synchronised pass."""
    import torch
    import torch.distributed as dist
    from oracle import ref_module as REF
    from oracle import trajectory_oracle as TO
    from oracle.ref_pipeline import reference_render
    from adgs_b200 import scenes
    wl = WORKLOADS[workload]
    tensors, n_scene, n_obj = scene_tensors(wl, device)
    for k, v in tensors.items():
        if k != "gs_time":
            v.requires_grad_(True)
    ref = TO.ReferenceModel(scenes.BENCH_ORDER_ARGS, True, **tensors)
    views = views_of_step(wl, rank, device)
    vpr = len(views)
    px = vpr * wl["W"] * wl["H"]
    cot = make_cotangents(wl, device)
    params = [getattr(ref, f) for f in ref.trainable()]
    cots = (cot["color"], cot["depth"], cot["opacity"], cot["flow"], cot["semantic"])

    def case(cam):
        return dict(cam=cam, W=wl["W"], H=wl["H"], n=wl["n"], background=torch.zeros(3, device=device),
                    tan_fovx=math.tan(cam.FoVx * 0.5), tan_fovy=math.tan(cam.FoVy * 0.5), degree=3, inv_depth=True)

    cases = [case(cam) for cam, _, _ in views]

    def step():
        for p in params:
            p.grad = None
        for c, (cam, t, flow_t) in zip(cases, views):       # one view per iteration, gradients accumulate
            (color, radii, depth, opac, flow, sem), _ = reference_render(ref, c, t, flow_t, REF)
            torch.autograd.backward((color, depth, opac, flow, sem), cots)
        if world > 1:
            for p in params:
                if p.grad is not None:
                    dist.all_reduce(p.grad)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        tt = torch.tensor([ms], device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    stages = None
    if stage_split and rank == 0:
        # where the reference's step goes: its ATen trajectory, its rasterizer kernels, autograd through the trajectory
        # (a separate pass with a device synchronisation between the stages; not part of `value`)
        acc = {"trajectory_forward": 0.0, "rasterizer_forward": 0.0, "rasterizer_backward": 0.0,
               "trajectory_backward_autograd": 0.0}
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        reps = 3
        from oracle.ref_pipeline import RefRasterize
        for _ in range(reps):
            for p in params:
                p.grad = None
            c, (cam, t, flow_t) = cases[0], views[0]
            torch.cuda.synchronize()
            ev[0].record()
            flow = ref.get_deformed_xyz(flow_t)
            pkg = ref.get_deformed_pkg(t)
            sem = ref.get_obj_mask().float()[..., None]
            scaling = ref.get_scaling()
            leaves = [x.detach().requires_grad_(True) for x in (pkg["xyz"], pkg["opacity"], scaling, pkg["rotation"],
                                                                 pkg["shs"], flow)]
            ev[1].record()
            out = RefRasterize.apply(REF, c, leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[5], sem)
            ev[2].record()
            torch.autograd.backward((out[0], out[2], out[3], out[4], out[5]), cots)
            ev[3].record()
            torch.autograd.backward((pkg["xyz"], pkg["opacity"], scaling, pkg["rotation"], pkg["shs"], flow),
                                    tuple(x.grad for x in leaves))
            ev[4].record()
            torch.cuda.synchronize()
            for i, k in enumerate(acc):
                acc[k] += ev[i].elapsed_time(ev[i + 1]) / reps
        stages = {k: round(v, 4) for k, v in acc.items()}
    del ref, tensors, params
    torch.cuda.empty_cache()
    return {"ms": round(ms, 4), "value": round(world * px / (ms * 1e-3) / 1e6, 3), "vpr": vpr, "stages": stages}


def reference_port_line(args, cores):
    """No GPU / no oracle/_ref: the numpy oracle port of the rasterizer on a bounded sample."""
    from oracle import raster_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as Hh
    c = Hh.make_case(n=2000, W=96, H=64, seed=1, device="cpu")
    s = Hh.oracle_settings(c)
    n_ = Hh.to_np
    cot = Hh.cotangents(c, device="cpu")
    t0 = time.perf_counter()
    k = 0
    while k < args.steps and time.perf_counter() - t0 < 60:
        out, st = O.rasterize_forward(s, n_(c["means3D"]), n_(c["opacity"]), n_(c["scales"]), n_(c["rotations"]),
                                      None, n_(c["sh"]), None, n_(c["flow_points"]), n_(c["semantic"]))
        O.rasterize_backward(s, st, out, n_(c["means3D"]), n_(cot["color"]), n_(cot["depth"]), n_(cot["flow"]),
                             n_(cot["semantic"]), n_(cot["opacity"]), n_(c["scales"]), n_(c["rotations"]), None,
                             n_(c["sh"]), n_(c["flow_points"]), n_(c["semantic"]))
        k += 1
    ms = (time.perf_counter() - t0) / max(k, 1) * 1e3
    value = round(96 * 64 / (ms * 1e-3) / 1e6, 3)
    wl = WORKLOADS[args.workload]
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args.workload, wl, 1, 1),
            "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": "port",
                             "sample": "numpy oracle port on the host, bounded sample: 2000 Gaussians at 96x64 (rasterizer only)"},
            "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti-375x1242-1M", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-staging", default="ring", choices=["ring", "alloc"],
                    help="device staging of the per-step inputs: ring of preallocated buffers, or a fresh tensor per step")
    ap.add_argument("--e2e-ablate", default="none", choices=["none", "cot", "cam", "metric"],
                    help="diagnosis: leave one per-step transfer out of the e2e loop (the line is then not a valid e2e)")
    ap.add_argument("--sync-free-outstanding", type=int, default=None,
                    help="sync-free forwards the host may run ahead of the device (GaussianModel.sync_free_outstanding)")
    ap.add_argument("--no-other-workloads", action="store_true",
                    help="skip the short runs of the other BASELINE configs (waymo 3-camera, stress 10M) at N=1")
    ap.add_argument("--launch-blocking-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--e2e-blocking", action="store_true", help="e2e: block on the metric read before queuing the next step")
    ap.add_argument("--parallel", default="exchange", choices=["exchange", "allreduce"],
                    help="multi-GPU data path (N > 1): splat exchange (default) or replicated model + gradient all-reduce")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

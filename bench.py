#!/usr/bin/env python
"""bench.py -- lift-splat frames/s + achieved HBM GB/s vs the B200 roofline, beside a CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--shape NAME]

One "step" = one pass of the hot path (height softmax -> per-frame geometry + voxel-run plan ->
fused BEV reduction) over one batch of B synthetic frames per GPU, every frame with its own
calibration (so the geometry is recomputed every step, as in training / multi-site inference).
For N > 1 launch through ``python -m torch.distributed.run`` (one rank per GPU); frames are
sharded across ranks, there is no collective on the data path ("replicas only").

Prints ONE JSON line (rank 0).  ``--impl reference`` times the CPU restatement of the reference path
(oracle/lift_splat_oracle.py: the reference's geometry calls + torch-CPU index_add_) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from sgv3d_b200.shapes import get_shape  # noqa: E402
from sgv3d_b200.synthetic import make_activations, make_mats  # noqa: E402

METRIC = "lift_splat_frames_per_sec"
UNIT = "frames/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic(kernels, shape_name, frames):
    """DRAM bytes (read + write) of one step from the committed ``ncu --set full`` capture
    (profiles/traffic_r01.json, written by tools/ncu_traffic.py; per launch at the capture's batch), summed over the kernels of this step."""
    p = _traffic_file()
    if p is None:
        return None
    t = json.load(open(p))
    if t.get("shape") != shape_name or t.get("frames_per_launch") != frames:
        return None
    if kernels is None:          # every kernel of the captured training step (plan + forward + backward)
        return sum(t["dram_bytes_per_launch"].values())
    tot = 0
    for base in sorted({k.split("(")[0] for k in kernels}):   # the two prep roles were captured as one launch
        if base not in t["dram_bytes_per_launch"]:
            return None
        tot += t["dram_bytes_per_launch"][base]
    return tot


def _traffic_file():
    for name in ("traffic_r02.json", "traffic_r01.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return p
    return None


def _pct(xs, q):
    xs = sorted(xs)
    if not xs:
        return None
    k = (len(xs) - 1) * q
    lo, hi = int(k), min(int(k) + 1, len(xs) - 1)
    return xs[lo] + (xs[hi] - xs[lo]) * (k - lo)


def _windows(fn, steps, repeats, sync):
    """`repeats` back-to-back windows of `steps` calls, each timed with CUDA events on the current stream;
    returns ms per step of every window."""
    out = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(repeats):
        sync()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        sync()
        out.append(e0.elapsed_time(e1) / steps)
    return out


def _stats(xs):
    return {"n_windows": len(xs), "ms_per_step_median": statistics.median(xs), "ms_per_step_p10": _pct(xs, 0.1),
            "ms_per_step_p90": _pct(xs, 0.9), "ms_per_step_min": min(xs), "ms_per_step_max": max(xs)}


def _mats_dict(mats, device):
    return {"sensor2ego_mats": mats["sensor2ego"].unsqueeze(1).to(device),
            "sensor2virtual_mats": mats["sensor2virtual"].unsqueeze(1).to(device),
            "intrin_mats": mats["intrin"].unsqueeze(1).to(device),
            "ida_mats": mats["ida"].unsqueeze(1).to(device),
            "reference_heights": mats["reference_heights"].unsqueeze(1).to(device),
            "bda_mat": mats["bda"].to(device) if mats.get("bda") is not None else None}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING a timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.idx), "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def count(self) -> int:
        """samples written so far"""
        try:
            return sum(1 for _ in open(self.f.name))
        except OSError:
            return 0

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].strip().lower() == "active":
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _prep_path(dev, B):
    """which implementation of the per-camera 4x4 prep (lss_fpn.py:361,367,392) served this run"""
    from sgv3d_b200 import view_transform as VT
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    ok = VT._CAMERA_PREP_OK.get((idx, (B, 1, 4, 4)))
    if ok:
        return "sgv3d_camera_prep (one kernel, verified bit-identical to torch at first use)"
    if VT._INVERSE_KERNEL_OK.get(idx):
        return "sgv3d_inverse4x4 + torch matmul (camera_prep differed from torch on this installation)"
    return "torch inverse / matmul (library prep kernels differed from torch on this installation)"


def _pipeline_name(plan):
    sel = int(plan.desc.reserved[0])
    from sgv3d_b200 import _native as N
    blk = bool(N.lib().sgv3d_lift_splat_uses_block_pipeline(plan.desc))
    from sgv3d_b200 import view_transform as VT
    auto = sel == 0 or VT._DEFAULT_PIPELINE == VT.PIPELINE_AUTO
    return ("pixel-block" if blk else "voxel-tile") + (" (auto policy)" if auto else " (forced)")


def _dist():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(shape, frames: int, warmup: int, budget_s: float = 40.0):
    """Times the CPU restatement of the reference forward path, one frame per call (the reference's
    get_geometry materialises ~40 MB of intermediates per frame; batching frames does not help it).
    Returns (frames_per_s, frames_timed, cores, seconds)."""
    from oracle import lift_splat_oracle as O
    from oracle import ref_cpu as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fr = O.create_frustum(shape.final_dim, shape.downsample, shape.d_bound)
    vs, vc, vn = O.grid_buffers(shape.x_bound, shape.y_bound, shape.z_bound)
    # the reference's own get_geometry / create_frustum (bytecode of the unmodified lss_fpn.py, oracle/_ref) when it
    # was staged by `make -C oracle ref`; else the line-by-line port
    ref_mod = R.make_module(shape) if R.available() else None
    # a few distinct inputs, rotated (generating them is not part of the timed path)
    nin = 4
    inputs = []
    for i in range(nin):
        mats = make_mats(shape, 1, 1, seed=1000 + i, bda="identity")
        logits, ctx = make_activations(shape, 1, 1, seed=i)
        inputs.append((mats, logits, ctx))
    total, n = 0.0, 0
    t_begin = time.perf_counter()
    for i in range(warmup + frames):
        mats, logits, ctx = inputs[i % nin]
        t0 = time.perf_counter()
        if ref_mod is not None:
            R.forward(ref_mod, logits, ctx, mats)
        else:
            O.lift_splat_forward(logits, ctx, fr, mats, vc, vs, vn)
        dt = time.perf_counter() - t0
        if i >= warmup:
            total += dt
            n += 1
        if time.perf_counter() - t_begin > budget_s and n >= 2:
            break
    return n / total, n, cores, total, ("reference" if ref_mod is not None else "port")


_KIND_TEXT = {"reference": "the reference's own LSSFPN.get_geometry / create_frustum (unmodified lss_fpn.py, byte-compiled into "
                           "oracle/_ref) + the lss_fpn.py:462-495 glue + torch-CPU index_add_ for the CUDA-only voxel_pooling op",
              "port": "torch-CPU port of the reference path (oracle/lift_splat_oracle.py)"}


def run_reference(args):
    rank, world, _ = _dist()
    if rank != 0:
        return
    shape = get_shape(args.shape)
    B, K, W = args.batch, args.steps, max(args.warmup, 1)
    # one step = the same B frames per step as our arm; the run is capped at ~2.5 minutes of CPU work
    fps, n, cores, total, kind = cpu_reference_run(shape, B * K, B * W if B * W < 64 else 64, budget_s=150.0)
    steps_done = n / B
    sample = (f"{n} frame(s) of {shape.name} ({steps_done:.2f} step(s) of {B} frames), forward only, {_KIND_TEXT[kind]}, "
              f"{total:.1f} s on {cores} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * B / fps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{shape.name} lift-splat forward (softmax + get_geometry + lift + index_add_ pooling) on "
                               f"the host CPU, {B} frames per step", "frames_per_step_per_gpu": B, "shape": shape.name,
                   "D": shape.D, "fH": shape.fH, "fW": shape.fW, "C": shape.channels, "grid": list(shape.grid)},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    rank, world, local = _dist()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = None
    if world > 1:
        from sgv3d_b200.sharding import bind_to_gpu_numa_node
        numa_node = bind_to_gpu_numa_node(local)   # pinned staging buffers local to the GPU's PCIe root
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout at the first collective; stdout carries exactly one JSON line,
        # so the banner is sent to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier(device_ids=[local])
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    from sgv3d_b200 import LiftSplat, _native as N
    shape = get_shape(args.shape)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    mod = LiftSplat(shape.x_bound, shape.y_bound, shape.z_bound, shape.d_bound, shape.final_dim,
                    shape.downsample, shape.channels).to(dev)
    # two rotating input sets, each far larger than L2 together with the output
    nsets = 2
    sets = []
    for i in range(nsets):
        seed = 10_000 * rank + i
        mats = make_mats(shape, B, 1, seed=seed, bda="identity")
        logits, ctx = make_activations(shape, B, 1, seed=seed, device=dev, generator_device=dev)
        hf = torch.cat((logits, ctx), 1).contiguous()
        sets.append((hf, _mats_dict(mats, dev), mats))
    in_bytes = sets[0][0].numel() * 4
    out_bytes = B * shape.channels * shape.grid[0] * shape.grid[1] * 4

    def eager_step(i):
        hf, md, _ = sets[i % nsets]
        with torch.no_grad():
            return mod.forward_single_sweep(hf, md)

    # the public serving entry point: one CUDA graph per resident input set (geometry + plan + forward are
    # all replayed every step; only the ~25 host-side launches are folded into one cudaGraphLaunch)
    from sgv3d_b200 import LiftSplatGraph
    for i in range(nsets):
        eager_step(i)
    N.launch_count(reset=True)
    graphs = [LiftSplatGraph(mod, hf, md, warmup=1) for hf, md, _ in sets]
    launches_per_step = N.launch_count(reset=True) // (2 * nsets)   # 1 warm-up + 1 captured call per set

    def step(i):
        return graphs[i % nsets]()

    if args.eager:
        step = eager_step

    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    N.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = N.launch_count(reset=True) if args.eager else launches_per_step * K
    # the timed region lasts a few milliseconds, shorter than nvidia-smi's sampling period: keep the very same
    # step loop running (untimed) for 0.4 s so that the clock / throttle samples are taken under this load
    t_ext = time.perf_counter()
    i = K
    # (nvidia-smi needs up to a second to start when eight ranks launch one each: keep the load up until it has
    # delivered at least ten samples, three seconds at most)
    while True:
        el = time.perf_counter() - t_ext
        if (el >= 0.4 and (sampler.p is None or sampler.count() >= 10)) or el >= 3.0:
            break
        for _ in range(16):
            step(i)
            i += 1
        torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["span"] = "timed region + >= 0.4 s of the same step loop (nvidia-smi -lms 20)"
    # ---- >= 100 iterations, median + p10 / p90: more windows of the same K steps (the contract window
    # above stays the one `value` is computed from)
    R = max(1, args.repeats)
    fwd_windows = [ms / K] + _windows(step, K, R - 1, barrier)

    # ---- training step: plan + fused forward + fused backward (north star: fwd+bwd roofline), raw plan API on the
    # head's logits, distinct calibration per frame, rotating input sets, graph-free (3 library calls per step)
    from sgv3d_b200.view_transform import LiftSplatPlan  # noqa: F401
    D_, C_ = shape.D, shape.channels
    train_sets = []
    ctx_dt = torch.bfloat16 if args.ctx == "bf16" else torch.float32
    for hf, md, _ in sets:
        plan = mod.make_plan(md, 0, C_, ctx_dt)
        gb = torch.randn(B, C_, shape.grid[1], shape.grid[0], device=dev)
        gout = torch.empty_like(hf)
        # BASELINE config 4: bf16 context (a separate tensor: the height logits stay fp32), fp32 accumulation and gradients
        cx = hf[:, D_:D_ + C_].to(ctx_dt).contiguous() if ctx_dt != torch.float32 else hf[:, D_:D_ + C_]
        train_sets.append((plan, hf, gb, gout, cx))

    def train_step(i):
        plan, hf, gb, gout, cx = train_sets[i % nsets]
        plan.rebuild()
        plan.forward(hf[:, :D_], cx, logits=True)
        plan.backward(gb, hf[:, :D_], cx, logits=True, out_height=gout[:, :D_], out_context=gout[:, D_:D_ + C_])

    for i in range(W):
        train_step(i)
    train_windows = _windows(train_step, K, R, barrier)

    ms_eager = None
    if not args.eager:
        for i in range(3):
            eager_step(i)
        barrier()
        e0.record()
        for i in range(K):
            eager_step(i)
        e1.record()
        barrier()
        ms_eager = e0.elapsed_time(e1)

    # ---- per-kernel durations (CUDA events inside the library, same loop) -------------------------
    N.profile_enable(True)
    N.profile_report()
    for i in range(K):
        eager_step(i)
    torch.cuda.synchronize()
    prof = N.profile_report()
    N.profile_enable(False)
    kern = {k: {"launches": n, "avg_us": 1e3 * t / n, "share": 0.0} for k, (n, t) in prof.items()}
    lib_ms = sum(t for _, t in prof.values())
    for k, (n, t) in prof.items():
        kern[k]["share"] = t / lib_ms if lib_ms else 0.0
    dominant = max(prof.items(), key=lambda kv: kv[1][1])[0] if prof else None

    # ---- end to end: pinned host inputs -> H2D -> forward -> D2H of the BEV map ---------------------
    # through the public serving loop (sgv3d_b200.pipeline.LiftSplatPipeline: upload / compute / download
    # streams, rotating slots); every step uploads its inputs and downloads its BEV map.
    from sgv3d_b200.pipeline import LiftSplatPipeline
    hf_host = [s[0].cpu().pin_memory() for s in sets]
    md_host = [{k: (v.cpu().pin_memory() if v is not None else None) for k, v in s[1].items()} for s in sets]
    pipe = LiftSplatPipeline(mod, sets[0][0], sets[0][1], depth=3, device=dev)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    checksum = 0.0

    def e2e_run(n):
        nonlocal checksum
        pending = []
        for i in range(n):
            pending.append(pipe.submit(hf_host[i % nsets], md_host[i % nsets]))
            if len(pending) == pipe.depth:
                checksum += float(pipe.result(pending.pop(0))[0, 0, 0, 0])   # consume the oldest result on the host
        while pending:
            checksum += float(pipe.result(pending.pop(0))[0, 0, 0, 0])

    e2e_run(4)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    e2e_run(K)
    torch.cuda.synchronize()
    ms_e2e = 1e3 * (time.perf_counter() - t0)   # host clock: the last D2H has landed in pinned memory
    barrier()

    # ---- max over ranks ---------------------------------------------------------------------------
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, ms_e2e] + fwd_windows + train_windows, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        fwd_windows = [float(x) for x in t[2:2 + len(fwd_windows)]]
        train_windows = [float(x) for x in t[2 + len(fwd_windows):]]
    frames_total = B * K * world
    value = frames_total / (ms * 1e-3)
    e2e_value = frames_total / (ms_e2e * 1e-3)

    extra = {}
    if not args.quick:
        shapes = shape_lines(args, dev, world, barrier)     # every rank takes part (sharded like the headline)
        if rank == 0:
            extra = extra_measurements(args, shape, mod, sets, dev)
            extra["shapes"] = shapes

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    fwd_bytes = shape.fused_forward_bytes() * B           # 8.77 MB/frame at DAIR-R50
    cb = 2 if args.ctx == "bf16" else 4
    # + 12.29 MB/frame = 21.06 MB/frame with fp32 context (19.40 MB with bf16 context, fp32 grad_ctx)
    train_bytes = shape.fused_forward_bytes(cb) * B + shape.fused_backward_bytes(cb) * B
    # forward-only step: ONE CUDA-graph launch (4x4 prep + plan + forward); its duration is the device-timed contract
    # window above.  Training step: plan + forward + backward, timed the same way (CUDA events on the launch stream).
    step_ms = ms / K
    fwd_achieved = fwd_bytes / (step_ms * 1e-3) / 1e9 if step_ms else 0.0
    train_ms = statistics.median(train_windows)
    train_achieved = train_bytes / (train_ms * 1e-3) / 1e9
    tf = _traffic_file()
    # headline roofline: the north star's target quantity -- fused forward + backward of one step against the HBM
    # roofline; the forward-only (inference) step is reported beside it
    roofline = {
        "bound": "hbm", "achieved": train_achieved, "peak": peak, "unit": "GB/s", "frac": train_achieved / peak,
        "traffic": _ncu_traffic(None, shape.name, B) if args.ctx == "f32" else None,
        "traffic_source": ("committed ncu --set full capture %s (same shape and batch); not re-measured in this run"
                           % os.path.relpath(tf, ROOT)) if tf else None,
        "kernel": "fused lift-splat training step = plan + forward + backward kernels of one step (every kernel the "
                  "library launches for %d frames), %s context" % (B, args.ctx),
        "algorithmic_bytes_per_launch": train_bytes, "launch_ms": train_ms, "windows": _stats(train_windows),
        "peak_source": peak_src, "frac_of_8TBs_nominal": train_achieved / 8000.0,
        "forward_only": {
            "kernel": "4x4 prep + plan + forward of one step, one CUDA-graph launch (the step `value` is computed from)",
            "algorithmic_bytes_per_launch": fwd_bytes, "launch_ms": step_ms, "achieved": fwd_achieved,
            "frac": fwd_achieved / peak, "windows": _stats(fwd_windows),
            "traffic": _ncu_traffic(list(kern.keys()), shape.name, B),
            "kernel_sum_ms": lib_ms / K if K else 0.0},
        "dominant_kernel": dominant, "dominant_share": kern[dominant]["share"] if dominant else None,
        "kernels": kern,
    }
    dom_fwd = max(((k, v) for k, v in kern.items()), key=lambda kv: kv[1]["avg_us"])[0] if kern else None
    if dom_fwd is not None:
        roofline["forward_only"]["dominant_kernel"] = dom_fwd
    cpu = None
    if world == 1:
        fps, n, cores, total, kind = cpu_reference_run(shape, 2000, 2, budget_s=12.0)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n} frame(s) of {shape.name}, forward only, {_KIND_TEXT[kind]}, {total:.1f} s"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{shape.name} fused lift-splat forward (softmax + per-frame geometry/plan + BEV "
                               f"reduction), {B} frames per step per GPU, distinct calibration per frame",
                   "shape": shape.name, "frames_per_step_per_gpu": B, "D": shape.D, "fH": shape.fH, "fW": shape.fW,
                   "C": shape.channels, "grid": list(shape.grid), "arith": "PAIR (torch-CUDA bmm order)",
                   "l2": f"working set per step ({(in_bytes + out_bytes) / 1e6:.0f} MB in+out, 2 rotating input "
                         f"sets) exceeds the 126 MB L2",
                   "sharding": "frames across ranks, no collective (replicas only)",
                   "numa": "rank 0 bound to node %s" % numa_node if numa_node is not None else "not bound",
                   "camera_prep_path": _prep_path(dev, B),
                   "pipeline": _pipeline_name(train_sets[0][0]),
                   "timing": "W >= 3 warm-up steps; %d windows of %d steps each (CUDA events, barrier + synchronize on "
                             "both sides); `value` from the first window, median / p10 / p90 in roofline.*.windows" % (R, K)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K,
                "how": "LiftSplatPipeline: pinned host -> H2D -> CUDA-graph step -> D2H to pinned host, 3 slots on "
                       "upload/compute/download streams; host wall clock until the last BEV map is in host memory"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if ms_eager is not None:
        extra["eager_launch_frames_per_s_per_gpu"] = B * K / (ms_eager * 1e-3)
        extra["eager_launch_ms_per_step"] = ms_eager / K
    if extra:
        line["extra"] = extra
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def _time_loop(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def shape_lines(args, dev, world, barrier):
    """The other BASELINE.json configs, short, on EVERY rank (frames sharded like the headline): forward step (plan rebuilt
    every step, CUDA-graph replay) and training step (plan + forward + backward) on the second mandatory shape SGV3D-BSM-R50,
    the Rope3D-shaped ones (config 3) and the bf16-context training config (config 4: 8 frames per GPU).  Times are the max
    over ranks; frames/s are whole-job aggregates."""
    from sgv3d_b200 import LiftSplat, LiftSplatGraph
    peak, _ = _peaks()
    out = {}

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / iters
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    rank = _dist()[0]
    for nm, nb, dt in (("sgv3d_bsm_r50", 16, torch.float32), ("rope3d_r50", 32, torch.float32),
                       ("rope3d_native", 32, torch.float32), ("dair_r50", 8, torch.bfloat16),
                       ("sgv3d_bsm_r50", 8, torch.bfloat16)):
        sh = get_shape(nm)
        m2 = LiftSplat(sh.x_bound, sh.y_bound, sh.z_bound, sh.d_bound, sh.final_dim, sh.downsample, sh.channels).to(dev)
        mats2 = make_mats(sh, nb, 1, seed=77 + 1000 * rank, bda="identity")
        md2 = _mats_dict(mats2, dev)
        lg2, cx2 = make_activations(sh, nb, 1, seed=77 + 1000 * rank, device=dev, generator_device=dev)
        key = f"{nm}_b{nb}_{'bf16ctx' if dt == torch.bfloat16 else 'f32'}"
        rec = {"frames_per_gpu": nb, "n_gpus": world, "ctx_dtype": str(dt).split(".")[-1]}
        ctx_b = 2 if dt == torch.bfloat16 else 4
        fb2, bb2 = sh.fused_forward_bytes(ctx_b) * nb, sh.fused_backward_bytes(ctx_b) * nb
        if dt == torch.float32:
            hf2 = torch.cat((lg2, cx2), 1).contiguous()
            g2 = LiftSplatGraph(m2, hf2, md2, warmup=1)
            ms_f = timed(g2)
            rec["forward_step"] = {"ms": ms_f, "frames_per_s": world * nb / (ms_f * 1e-3),
                                   "frac_of_measured_peak": fb2 / ms_f / 1e6 / peak}
            del g2
        plan2 = m2.make_plan(md2, 0, sh.channels, dt)
        cxd = cx2.to(dt)
        gb2 = torch.randn(nb, sh.channels, sh.grid[1], sh.grid[0], device=dev)

        def tstep():
            plan2.rebuild()
            plan2.forward(lg2, cxd, logits=True)
            plan2.backward(gb2, lg2, cxd, logits=True)
        ms_t = timed(tstep)
        rec["train_step"] = {"ms": ms_t, "frames_per_s": world * nb / (ms_t * 1e-3), "algorithmic_bytes_per_gpu": fb2 + bb2,
                             "achieved_GBs_per_gpu": (fb2 + bb2) / ms_t / 1e6,
                             "frac_of_measured_peak": (fb2 + bb2) / ms_t / 1e6 / peak}
        rec["pipeline"] = _pipeline_name(plan2)
        out[key] = rec
        del plan2, m2
    # consumer layout: the headline shape with the BEV map written / its gradient read in
    # torch.channels_last order (same values; for a BEV trunk in that memory format).  Not the headline: the reference
    # returns a contiguous (B, C, Y, X) map.
    sh, nb = get_shape("dair_r50"), 64
    m3 = LiftSplat(sh.x_bound, sh.y_bound, sh.z_bound, sh.d_bound, sh.final_dim, sh.downsample, sh.channels,
                   bev_channels_last=True).to(dev)
    md3 = _mats_dict(make_mats(sh, nb, 1, seed=78 + 1000 * rank, bda="identity"), dev)
    lg3, cx3 = make_activations(sh, nb, 1, seed=78 + 1000 * rank, device=dev, generator_device=dev)
    hf3 = torch.cat((lg3, cx3), 1).contiguous()
    fb3, bb3 = sh.fused_forward_bytes(4) * nb, sh.fused_backward_bytes(4) * nb
    g3 = LiftSplatGraph(m3, hf3, md3, warmup=1)
    ms_f = timed(g3)
    del g3
    plan3 = m3.make_plan(md3, 0, sh.channels)
    gb3 = torch.randn(nb, sh.channels, sh.grid[1], sh.grid[0], device=dev).contiguous(memory_format=torch.channels_last)

    def tstep3():
        plan3.rebuild()
        plan3.forward(hf3[:, :sh.D], hf3[:, sh.D:], logits=True)
        plan3.backward(gb3, hf3[:, :sh.D], hf3[:, sh.D:], logits=True)
    ms_t = timed(tstep3)
    out["dair_r50_b64_f32_bev_channels_last"] = {
        "frames_per_gpu": nb, "n_gpus": world, "ctx_dtype": "float32", "bev_memory_format": "torch.channels_last",
        "forward_step": {"ms": ms_f, "frames_per_s": world * nb / (ms_f * 1e-3), "frac_of_measured_peak": fb3 / ms_f / 1e6 / peak},
        "train_step": {"ms": ms_t, "frames_per_s": world * nb / (ms_t * 1e-3), "algorithmic_bytes_per_gpu": fb3 + bb3,
                       "achieved_GBs_per_gpu": (fb3 + bb3) / ms_t / 1e6, "frac_of_measured_peak": (fb3 + bb3) / ms_t / 1e6 / peak},
        "pipeline": _pipeline_name(plan3)}
    del plan3, m3
    return out


def extra_measurements(args, shape, mod, sets, dev):
    """Secondary numbers (not the headline): training step, cached plan, batch-1 latency, the op-level
    drop-in."""
    from sgv3d_b200 import _native as N, voxel_pooling, lift_splat
    out = {}
    B = args.batch
    hf, md, mats = sets[0]
    D, C = shape.D, shape.channels
    peak, _ = _peaks()
    height = hf[:, :D].softmax(1).contiguous()
    ctx = hf[:, D:].contiguous()
    plan = mod.make_plan(md, 0, C)
    gb = torch.randn(B, C, shape.grid[1], shape.grid[0], device=dev)
    it = max(5, args.steps // 2)
    ms_plan = _time_loop(plan.rebuild, it)
    ms_fwd = _time_loop(lambda: plan.forward(height, ctx), it)
    ms_bwd = _time_loop(lambda: plan.backward(gb, height, ctx), it)
    fb = shape.fused_forward_bytes() * B
    bb = shape.fused_backward_bytes() * B
    out["phases_ms"] = {"plan": ms_plan, "forward_cached_plan": ms_fwd, "backward": ms_bwd}
    out["cached_plan_forward"] = {"frames_per_s": B / (ms_fwd * 1e-3), "achieved_GBs": fb / ms_fwd / 1e6,
                                  "frac_of_measured_peak": fb / ms_fwd / 1e6 / peak}
    out["train_step_fwd_bwd"] = {"frames_per_s": B / ((ms_plan + ms_fwd + ms_bwd) * 1e-3),
                                 "achieved_GBs": (fb + bb) / (ms_plan + ms_fwd + ms_bwd) / 1e6,
                                 "frac_of_measured_peak": (fb + bb) / (ms_plan + ms_fwd + ms_bwd) / 1e6 / peak,
                                 "algorithmic_bytes": fb + bb}
    # batch-1 latency through the module call site
    md1 = {k: (v[:1].contiguous() if v is not None else None) for k, v in md.items()}
    hf1 = hf[:1].contiguous()
    with torch.no_grad():
        out["batch1_latency_us"] = 1e3 * _time_loop(lambda: mod.forward_single_sweep(hf1, md1), 20)
    from sgv3d_b200 import LiftSplatGraph
    g1 = LiftSplatGraph(mod, hf1, md1)
    out["batch1_graph_latency_us"] = 1e3 * _time_loop(g1, 50)
    # BASELINE config 2 (batch-1 inference on a static roadside camera): plan built once, replay = forward kernels
    g1s = LiftSplatGraph(mod, hf1, md1, static_calibration=True)
    out["batch1_static_camera_graph_latency_us"] = 1e3 * _time_loop(g1s, 50)
    del g1, g1s
    # frames/s of the headline step (plan rebuilt every step, CUDA-graph replay) over the batch sweep of
    # BASELINE config 3
    sweep = {}
    for nb in (8, 16, 32, 64, 128):
        if nb > 2 * B:
            continue
        reps = (nb + B - 1) // B
        hfb = hf.repeat(reps, 1, 1, 1)[:nb].contiguous()
        mdb = {k: (v.repeat(reps, *([1] * (v.dim() - 1)))[:nb].contiguous() if v is not None else None)
               for k, v in md.items()}
        gb_ = LiftSplatGraph(mod, hfb, mdb, warmup=1)
        sweep[str(nb)] = nb / (_time_loop(gb_, 10) * 1e-3)
        del gb_, hfb, mdb
    out["frames_per_s_by_batch"] = sweep
    # op-level drop-in (materialised frustum features are an API input there); the reference's own kernel,
    # recompiled for sm_100a, is timed on the same inputs by tests/bench_reference_kernel.py (test infrastructure)
    nb = min(B, 4)
    idx = mod.get_geometry_indices(md["sensor2ego_mats"][:nb, 0], md["sensor2virtual_mats"][:nb, 0],
                                   md["intrin_mats"][:nb, 0], md["ida_mats"][:nb, 0],
                                   md["reference_heights"][:nb, 0], md["bda_mat"][:nb])
    feat = (height[:nb].unsqueeze(1) * ctx[:nb].unsqueeze(2)).reshape(nb, 1, C, D, shape.fH, shape.fW)
    feat = feat.permute(0, 1, 3, 4, 5, 2).contiguous()
    with torch.no_grad():
        ms_op = _time_loop(lambda: voxel_pooling(idx, feat, list(shape.grid)), 5)
    ob = shape.op_forward_bytes() * nb
    out["op_level_forward"] = {"frames": nb, "ms": ms_op, "frames_per_s": nb / (ms_op * 1e-3),
                               "achieved_GBs": ob / ms_op / 1e6, "frac_of_measured_peak": ob / ms_op / 1e6 / peak}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per step per GPU")
    ap.add_argument("--shape", default="dair_r50")
    ap.add_argument("--quick", action="store_true", help="skip the secondary measurements")
    ap.add_argument("--repeats", type=int, default=5, help="timed windows of --steps steps each (median / p10 / p90)")
    ap.add_argument("--ctx", default="f32", choices=["f32", "bf16"], help="context dtype of the training step (config 4: bf16)")
    ap.add_argument("--eager", action="store_true", help="time per-kernel launches instead of CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

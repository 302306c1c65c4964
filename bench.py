#!/usr/bin/env python
"""bench.py — scenes/s of one Voxel-DETR training step on synthetic Waymo-shaped scenes.

    python bench.py --gpus N --steps K --warmup W            # the B200 path (efg_b200, CUDA kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): Voxel-DETR 1-frame,
300 queries, 2 scenes of 150k points per GPU, Waymo grid 1504x1504x40; one step = GPU voxelization
(+fused mean-VFE) -> sparse ResNet18 -> FPN(p3) -> 3 box-attention encoder layers -> top-k proposals
-> 3 decoder layers -> heads -> Hungarian matching -> losses -> backward -> gradient all-reduce ->
AdamW update.  fp32 throughout (the reference runs fp32, SURVEY.md §5 AMP row).

JSON line keys (see the task contract): value = scenes/s with the point clouds already in HBM;
e2e = the same step through the public model API from pinned HOST buffers (H2D of the points and
D2H of the loss inside the timed region); roofline = the dominant efg_b200 kernel family measured
with CUDA events in-situ; cpu_baseline = the oracle-backed model on the host cores on a bounded
sample (one scene).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

POINTS_PER_SCENE = 150000
SCENES_PER_GPU = 2
NUM_QUERIES = 300
WORKLOAD = "voxel_detr_waymo_1f_q300_bs2_150kpts"
# --workload -> (config.workload name, scenes per GPU, points per scene, BASELINE.json config index)
WORKLOADS = {
    "voxel_detr": (WORKLOAD, 2, 150000, 2),
    "conquer": ("conquer_waymo_1f_q300_dn3_bs2_150kpts", 2, 150000, 3),
    "centerpoint_waymo": ("centerpoint_waymo_1f_bs4_150kpts", 4, 150000, 1),
    "centerpoint_nusc": ("centerpoint_nuscenes_11sweeps_bs1_400kpts", 1, 400000, 4),
    "config1": ("voxelize_subm16_20kpts_bs1", 1, 20000, 0),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="efgb200", choices=["efgb200", "reference", "torch_gpu"],
                    help="efgb200: this repo's CUDA path; reference: the reference algorithm on the host CPU; torch_gpu: the "
                         "same module graph over plain-torch gather-mm-index_add ops on the GPU (labelled stand-in for the "
                         "reference's spconv GPU path, which is not installable here)")
    ap.add_argument("--workload", default="voxel_detr", choices=sorted(WORKLOADS),
                    help="voxel_detr = BASELINE.json configs[2] (the metric's configuration, default); the others are the "
                         "remaining BASELINE.json configs")
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--scenes", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prefetch", action="store_true",
                    help="voxelize / build rulebooks in line with the step instead of one batch ahead on a side stream")
    ap.add_argument("--no-graph", action="store_true",
                    help="keep the static section (FPN top-down, transformer, heads) eager instead of CUDA-graphed")
    ap.add_argument("--cpu-points", type=int, default=POINTS_PER_SCENE)
    ap.add_argument("--profile-step", action="store_true",
                    help="for `ncu --profile-from-start off`: after the warm-up run ONE step between "
                         "cudaProfilerStart/Stop and exit (no timing, no JSON line)")
    a = ap.parse_args()
    name, scenes, points, _ = WORKLOADS[a.workload]
    a.workload_name = name
    a.scenes = scenes if a.scenes is None else a.scenes
    a.points = points if a.points is None else a.points
    return a


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region by a background thread through NVML (in-process;
    the NVML calls release the GIL, so the enqueueing thread is not held up), a few samples per second.
    Why not the recipe's `nvidia-smi -lms 100` subprocess: on this pool one NVML clock/reason query takes 4-24 ms of
    host time, and a 10 Hz poll (or a blocking query after every step) slowed the partly launch-bound step by 15-40 %
    (measured: 57-68 ms per step with, 47 ms without).  nvidia-smi is only the fallback when NVML is unavailable."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index, period_s=0.2):
        self.index = index
        self.period = period_s
        self.sm, self.reasons, self.cost_ms = [], set(), []
        self.max_mhz = None
        self.nvml = None
        self.handle = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].strip().isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None
        # The first clock / reason query of a process takes ~60 ms on a fresh box and holds a driver lock the launching
        # thread needs: take it (and throw it away) here, outside the timed region.
        self.sample()
        self.sm, self.reasons, self.cost_ms = [], set(), []

    def sample(self):
        t0 = time.perf_counter()
        if self.nvml is not None:
            n = self.nvml
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = int(get(self.handle))
                for bit, name in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
        else:
            self._sample_smi()
        self.cost_ms.append((time.perf_counter() - t0) * 1e3)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip().splitlines()[0]
            parts = [p.strip() for p in out.split(",")]
            self.sm.append(float(parts[0]))
            self.max_mhz = float(parts[1])
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if val.lower().startswith("active"):
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.wait(self.period):
            self.sample()

    def start(self):
        """Call right before the timed region (the first sample is taken one period in)."""
        self._stop.clear()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        """Call after the last step has been ENQUEUED and before the closing synchronize: takes one more sample while
        the device is still executing, then ends the thread."""
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=10)
        if not self.sm:
            self.sample()

    def result(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "host_ms_per_sample": round(float(np.mean(self.cost_ms)), 3),
                "how": ("NVML" if self.nvml else "nvidia-smi") + " from a background thread every %.0f ms during the timed region" % (self.period * 1e3)}


def kernel_of(family):
    """Launch family (efg_b200.ops.PROFILER tag) -> the __global__ function it launches."""
    if family.startswith("spconv_tc_wgrad"):
        return "spconv_wgrad_tc_kernel[sparse conv]"
    if family == "dense_tc_wgrad":
        return "spconv_wgrad_tc_kernel[dense linear]"
    if family.startswith("spconv_tc_c"):
        return "spconv_tc_kernel[sparse conv]"
    if family == "dense_tc_gemm":
        return "spconv_tc_kernel[dense linear]"
    table = {"box_attn_fwd": "box_attn_fwd_tile_kernel", "box_attn_bwd": "box_attn_bwd_tile_kernel",
             "box_grid_softmax_fwd": "box_grid_softmax_kernel", "box_grid_softmax_bwd": "box_grid_softmax_kernel",
             "spconv_pack_weights": "pack_weights_kernel", "colsum": "colsum_partial_kernel"}
    if family in table:
        return table[family]
    if family.startswith("spconv_gemm"):
        return "spconv_fwd_kernel"
    if family.startswith("spconv_wgrad"):
        return "spconv_wgrad_kernel"
    return family


def ncu_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` from the newest committed
    ncu capture of one bench step (profiles/*traffic*.json, written by scripts/ncu_summarize.py) — a profiler
    number, so it is read from the committed capture, never measured inside a timed run."""
    import glob

    best = (None, None)
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*traffic*.json"))):
        try:
            with open(path) as f:
                js = json.load(f)
        except Exception:
            continue
        for name, d in js.get("kernels", {}).items():
            if name.split("::")[-1] == kernel.split("[")[0] and d.get("dram_bytes_per_launch"):
                best = (int(d["dram_bytes_per_launch"]), os.path.relpath(path, ROOT))
    return best


def make_scenes(n_scenes, n_points, seed, spec=None):
    from efg_b200.data import WAYMO, make_scene

    return [make_scene(n_points, spec or WAYMO, seed=seed * 100 + i) for i in range(n_scenes)]


def build_workload(args, device, backend=None):
    """-> (model, scene spec, config) of --workload on `device` (backend None = the CUDA kernels)."""
    from efg_b200.config import centerpoint_config, conquer_config, voxel_detr_config
    from efg_b200.data import NUSCENES, WAYMO

    kw = {} if backend is None else {"backend": backend}
    if args.workload == "voxel_detr":
        from efg_b200.detectors.voxel_detr import VoxelDETR

        cfg = voxel_detr_config(model={"device": device, "transformer": {"num_queries": NUM_QUERIES}})
        return VoxelDETR(cfg, **kw), WAYMO, cfg
    if args.workload == "conquer":
        from efg_b200.detectors.conquer import ConQueR

        cfg = conquer_config(model={"device": device, "transformer": {"num_queries": NUM_QUERIES}})
        return ConQueR(cfg, **kw), WAYMO, cfg
    from efg_b200.detectors.centerpoint import VoxelNet

    if args.workload == "centerpoint_waymo":
        cfg = centerpoint_config(model={"device": device}, dataset={"max_voxel_num": 120000})
        return VoxelNet(cfg, **kw), WAYMO, cfg
    # nuScenes multi-sweep CenterPoint (CPN/config.yaml:14-15, 95-121): 6 task heads with velocity
    tasks = [{"num_classes": 1, "class_names": ["car"]}, {"num_classes": 2, "class_names": ["truck", "construction_vehicle"]},
             {"num_classes": 2, "class_names": ["bus", "trailer"]}, {"num_classes": 1, "class_names": ["barrier"]},
             {"num_classes": 2, "class_names": ["motorcycle", "bicycle"]}, {"num_classes": 2, "class_names": ["pedestrian", "traffic_cone"]}]
    cfg = centerpoint_config(
        dataset={"classes": NUSCENES.classes, "pc_range": NUSCENES.pc_range, "voxel_size": NUSCENES.voxel_size,
                 "max_points_in_voxel": 10, "max_voxel_num": 160000},
        model={"device": device, "head": {"tasks": tasks, "misc": {"dataset": "nuscenes", "weight": 0.25, "code_weights": [1.0] * 8 + [0.2, 0.2],
                                                               "common_heads": {"reg": [2, 2], "height": [1, 2], "dim": [3, 2],
                                                                                "rot": [2, 2], "vel": [2, 2]}}}})
    return VoxelNet(cfg, **kw), NUSCENES, cfg



# -------------------------------------------------------------------------------------------------
# B200 arm
# -------------------------------------------------------------------------------------------------
def loss_total(losses):
    """Sum of the loss terms a trainer optimises (Voxel-DETR / ConQueR: every `loss*` key; CenterPoint: `<task>_loss`)."""
    keys = [k for k in losses if k.startswith("loss")] or [k for k in losses if k.endswith("_loss") and k.count("_") == 1]
    return sum(losses[k] for k in keys)


def rooflines(summary, prof_steps, prof_ms, peaks):
    """Per kernel group: algorithmic bytes / flops over the CUDA-event time of its launches (ops.PROFILER tags), against
    the measured peaks.  Sparse-conv launches and dense-linear launches of the same __global__ function are separate
    groups, so the dense GEMMs cannot flatter the sparse-conv figure."""
    f16_peak = peaks["bf16_tflops"]           # kind::f16 (the default bf16x3 mode); kind::tf32 (wgrad) runs at half
    groups = {}
    for fam, d in summary.items():
        k = groups.setdefault(kernel_of(fam), {"launches": 0, "ms": 0.0, "bytes": 0, "flops": 0, "families": []})
        for key in ("launches", "ms", "bytes", "flops"):
            k[key] += d[key]
        k["families"].append(fam)
    out = {}
    for name, dd in groups.items():
        sec = max(dd["ms"], 1e-9) * 1e-3
        gbs, tfs = dd["bytes"] / sec / 1e9, dd["flops"] / sec / 1e12
        tpeak = f16_peak / 2.0 if "wgrad" in name else f16_peak
        ridge = tpeak * 1e12 / (peaks["hbm_gbs"] * 1e9)
        intensity = dd["flops"] / max(dd["bytes"], 1)
        traffic, traffic_src = ncu_traffic(name)
        e = {"kernel": name, "families": sorted(dd["families"]), "launches_per_step": dd["launches"] / prof_steps,
             "avg_launch_ms": round(dd["ms"] / dd["launches"], 4), "ms_per_step": round(dd["ms"] / prof_steps, 4),
             "share_of_step": round(dd["ms"] / prof_steps / prof_ms, 4),
             "algorithmic_bytes_per_launch": int(dd["bytes"] / dd["launches"]),
             "algorithmic_flops_per_launch": int(dd["flops"] / dd["launches"]),
             "flop_per_byte": round(intensity, 1), "ridge_flop_per_byte": round(ridge, 1),
             "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks["source"],
             "hbm": {"achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(gbs / peaks["hbm_gbs"], 4)},
             "tensor": {"achieved": round(tfs, 2), "peak": round(tpeak, 1), "unit": "TFLOP/s", "frac": round(tfs / tpeak, 4),
                        "note": "algorithmic flops (the bf16x3 / 3xTF32 splits issue 3x as many); peak = measured sustained "
                                "bf16 rate (kind::f16), half of it for the tf32 wgrad kernel"}}
        side = "tensor" if (dd["flops"] and intensity > ridge) else "hbm"
        e.update(bound=side, achieved=e[side]["achieved"], peak=e[side]["peak"], unit=e[side]["unit"], frac=e[side]["frac"])
        out[name] = e
    return out


def run_efgb200(args, backend=None):
    import torch.distributed as dist

    from efg_b200 import _lib, ops
    from efg_b200.parallel import GradAverager, init_distributed

    rank, local_rank, world = init_distributed()
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # The whole job runs on one non-default stream.  CUDA-graph capture of the static section needs it: autograd ties a
    # parameter's gradient accumulator to the stream the parameter was first used on, and a capture may fork / join
    # any stream except the legacy default one ("operation would make the legacy stream depend on a capturing blocking
    # stream").  Every event below is recorded on this stream.
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    torch.manual_seed(0)
    # the dense 2-D convolutions of the FPN / RPN are cuDNN library calls on fixed shapes: let cuDNN pick its algorithm
    torch.backends.cudnn.benchmark = bool(int(os.environ.get("EFGB_CUDNN_BENCHMARK", "1")))
    if args.workload == "config1":
        return run_config1(args, dev)

    model, spec, cfg = build_workload(args, "cuda:%d" % local_rank, backend)
    model.train()
    averager = GradAverager(model)
    averager.broadcast_parameters()
    # the reference's optimizer (AdamW, VD/config.yaml solver); fused = torch's single multi-tensor CUDA implementation
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.01,
                            betas=(0.9, 0.99), eps=1e-9, fused=True)

    # a few distinct batches so consecutive steps do not see identical data
    n_batches = 2
    host_batches = [make_scenes(args.scenes, args.points, seed=1 + rank * 10 + b, spec=spec) for b in range(n_batches)]
    pinned = [[(torch.from_numpy(p).pin_memory(), a) for p, a in hb] for hb in host_batches]
    resident = [[(t.to(dev), a) for t, a in pb] for pb in pinned]
    h2d_bytes = sum(t.numel() * 4 for t, _ in pinned[0])
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def inputs_of(batch):
        return [({"points": p}, {"annotations": a}) for p, a in batch]

    def step(batch, prepared=None):
        averager.zero_grad()          # one memset per gradient bucket; p.grad are views into the buckets
        losses = model(inputs_of(batch)) if prepared is None else model(inputs_of(batch), prepared=prepared)
        total = loss_total(losses)
        total.backward()              # bucket all-reduces are launched from inside backward (post-accumulate hooks)
        averager.finish()
        averager.hide_unused()        # parameters of pruned branches keep grad = None, as under DDP
        opt.step()
        return total

    # Input pipeline: the index part of batch i + 1 (H2D copy of the points, voxelizer, the strided rulebooks — the
    # only places that read a count back from the device) runs on its own high-priority stream while step i executes,
    # so the training stream never synchronises and the host stays ahead of the GPU (model.prepare).
    prefetch = backend is None and not args.no_prefetch and hasattr(model, "prepare")
    prep_stream = torch.cuda.Stream(device=dev, priority=-1) if prefetch else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = []

    def timed(batches, steps, from_host, sampler=None):
        """EXACTLY `steps` steps between one pair of CUDA events, bracketed by barrier + synchronize on both sides.
        L2 is flushed before every step (a 256 MiB memset inside the timed span, ~0.04 ms).  With host inputs every
        step copies its pinned point clouds to the device and reads its loss back: the read of step i is an
        asynchronous copy into pinned memory that the host waits for while step i + 1 is in flight (every read
        completes inside the timed span)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        last = None
        back = [torch.zeros(2, dtype=torch.float32).pin_memory() for _ in range(2)]   # [loss, matcher status]
        back_ev = [torch.cuda.Event(), torch.cuda.Event()]

        def read_back(i):
            back_ev[i & 1].synchronize()
            if back[i & 1][1] != 0:
                raise ValueError("matrix contains invalid numeric entries (device linear_sum_assignment)")
            return float(back[i & 1][0])

        gc.collect()   # start every timed region from the same collector state (the collector itself stays on: the autograd
        barrier()      # objects of a step hold device memory through reference cycles, and without it the allocator grows)
        if sampler is not None:
            sampler.start()
        ev0.record()
        host_t0 = time.perf_counter()
        prepared = None
        for i in range(steps):
            l2_flush.zero_()
            b = batches[i % len(batches)]
            if prefetch:
                if prepared is None:
                    prepared = model.prepare(inputs_of(b), prep_stream)   # first step of the span
                total = step(b, prepared)
                prepared = model.prepare(inputs_of(batches[(i + 1) % len(batches)]), prep_stream) if i + 1 < steps else None
            else:
                if from_host:
                    b = [(t.to(dev, non_blocking=True), a) for t, a in b]
                total = step(b)
            if not from_host:
                # bound the host's run-ahead to one step (what the asynchronous loss read does in the e2e arm): a host
                # that free-runs until the driver's launch queue is full enqueues SLOWER (measured at N = 2 on one box:
                # 57 ms of enqueue per step free-running, 38 ms throttled)
                back_ev[i & 1].record()
                if i >= 1:
                    back_ev[(i - 1) & 1].synchronize()
            if from_host:
                # D2H read of the step's result, one step behind the launches
                back[i & 1][0:1].copy_(total.detach().reshape(1), non_blocking=True)
                status = ops.lsa_status_tensor(dev) if backend is None else None
                if status is not None:
                    back[i & 1][1:2].copy_(status.reshape(1).float(), non_blocking=True)
                back_ev[i & 1].record()
                if i >= 1:
                    last = read_back(i - 1)
        if from_host and steps > 0:
            last = read_back(steps - 1)
        ev1.record()
        host_ms.append((time.perf_counter() - host_t0) * 1e3 / max(steps, 1))   # host enqueue time per step
        if sampler is not None:
            sampler.stop()  # the device is still executing the tail of the last step
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    for i in range(max(args.warmup, 3)):
        step(resident[i % n_batches])
    barrier()

    # CUDA graphs for the static-shape section of the step (FPN top-down, encoder, decoder, heads; forward + backward)
    graph_state = "none"
    if backend is None and not args.no_graph and not args.profile_step and hasattr(model, "enable_static_graph"):
        if True:
            ok = model.enable_static_graph([({"points": p}, {"annotations": a}) for p, a in resident[0]])
            what = {"voxel_detr": "static section (FPN top-down + transformer + heads), forward and backward",
                    "conquer": "encoder section (FPN top-down + encoder + proposal head), forward and backward; the decoders' "
                               "length depends on the ground truth (denoising groups) and stays eager"}.get(
                args.workload, "static section (RPN neck + centre heads + losses), forward and backward")
            graph_state = what if ok else "none (capture failed: %s)" % model.static_graph_error
            if not ok:
                sys.stderr.write("bench.py: CUDA graph capture failed, running eagerly: %s\n" % model.static_graph_error)
            for i in range(3):
                step(resident[i % n_batches])
            barrier()

    if args.profile_step:
        ops.enable_nvtx()
        torch.cuda.profiler.start()
        step(resident[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    # untimed warm-up of the two measured loops themselves (side-stream pipeline, pinned buffers, the allocator pools of the
    # prefetch stream): the steps above ran the plain in-line path
    timed(resident, 3, from_host=False)
    timed(pinned, 2, from_host=True)
    host_ms.clear()
    # everything alive now (model, graphs, caches) goes to the permanent generation: a full collection inside a timed
    # region then only walks the objects of the last steps instead of pausing for tens of ms (the collector stays on —
    # the autograd objects of a step hold device memory through reference cycles)
    gc.collect()
    gc.freeze()
    launches0 = _lib.lib().efgb_launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get("EFGB_BENCH_NO_CLOCKS") else None
    ms_dev, _ = timed(resident, args.steps, from_host=False, sampler=sampler)
    clocks = sampler.result() if sampler is not None else None
    launches_eager = _lib.lib().efgb_launch_count() - launches0
    # kernels inside the CUDA graphs are launched by the replay, not by a host call: counted at capture time, per replay
    launches_graph = int(getattr(model, "static_graph_launches", 0) or 0) * args.steps if graph_state.startswith(("static", "encoder")) else 0
    launches = launches_eager + launches_graph
    ms_e2e, last_loss = timed(pinned, args.steps, from_host=True)

    scenes_total = args.scenes * world * args.steps
    value = scenes_total / (ms_dev / 1e3)
    e2e_value = scenes_total / (ms_e2e / 1e3)

    # in-situ kernel attribution (extra steps, CUDA events around each library launch); eager: a graph replay runs no
    # Python, so the per-launch events could not be recorded inside it
    graphed_call = getattr(model, "_static_call", None)
    if graphed_call is not None:
        model._static_call = None
    ops.PROFILER = ops.KernelProfiler()
    prof_steps = 2
    t_prof0 = torch.cuda.Event(enable_timing=True)
    t_prof1 = torch.cuda.Event(enable_timing=True)
    t_prof0.record()
    for i in range(prof_steps):
        step(resident[i % n_batches])
    t_prof1.record()
    summary = ops.PROFILER.summary()
    ops.PROFILER = None
    prof_ms = t_prof0.elapsed_time(t_prof1) / prof_steps

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    kernels = {}
    for fam, d in summary.items():
        ms_per_launch = d["ms"] / d["launches"]
        kernels[fam] = {"launches_per_step": d["launches"] / prof_steps, "ms_per_step": round(d["ms"] / prof_steps, 4),
                        "gbs": round(d["bytes"] / d["launches"] / (ms_per_launch * 1e-3) / 1e9, 1),
                        "tflops": round(d["flops"] / d["launches"] / (ms_per_launch * 1e-3) / 1e12, 2)}
    impl = "efgb200" if backend is None else "torch_gpu"
    out = {
        "metric": "scenes/sec %s fwd+bwd" % {"voxel_detr": "Voxel-DETR", "conquer": "ConQueR"}.get(args.workload, "CenterPoint"),
        "value": round(value, 3), "unit": "scenes/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_dev / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload_name, "scenes_per_gpu": args.scenes, "points_per_scene": args.points,
                   "num_queries": NUM_QUERIES if args.workload in ("voxel_detr", "conquer") else None,
                   "grid": "x".join(str(int(g)) for g in spec.grid_size), "step": "voxelize+fwd+bwd+allreduce+adamw",
                   "conv_precision": ops.CONV_PRECISION if backend is None else "fp32 torch ops",
                   "cuda_graph": graph_state, "parallelism": "dp%d" % world,
                   "input_pipeline": ("index part of the next batch (H2D, voxelizer, strided rulebooks) on a side stream during "
                                      "the current step; host run-ahead bounded to one step by an event wait (e2e: the asynchronous loss read)") if prefetch else
                                     "in line (no prefetch)",
                   "l2": "flushed before every timed step (256 MiB memset, inside the timed span)"},
        "e2e": {"value": round(e2e_value, 3), "unit": "scenes/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3), "last_loss": last_loss},
        "gpu_launches": int(launches),
        "gpu_launches_detail": {"host_launched": int(launches_eager), "replayed_from_cuda_graphs": int(launches_graph)},
        "host_enqueue_ms_per_step": round(host_ms[0], 3),   # Python + launch time of a step (no synchronisation inside)
        "clocks": clocks,
    }
    if backend is None:
        rl = rooflines(summary, prof_steps, prof_ms, peaks)
        # `roofline`: the kernel the metric names — the sparse-conv forward / dgrad kernel on SPARSE launches only;
        # `rooflines`: every kernel group (dense linears, wgrad, box attention, voxelizer, rulebooks) incl. the largest by time
        headline = "spconv_tc_kernel[sparse conv]"
        out["roofline"] = rl.get(headline) or max(rl.values(), key=lambda e: e["ms_per_step"])
        out["rooflines"] = {k: {kk: v[kk] for kk in ("bound", "achieved", "peak", "unit", "frac", "ms_per_step", "launches_per_step",
                                                     "share_of_step", "hbm", "tensor")}
                            for k, v in sorted(rl.items(), key=lambda kv: -kv[1]["ms_per_step"])}
        out["dominant_kernel_by_time"] = max(rl.values(), key=lambda e: e["ms_per_step"])["kernel"]
        out["kernels"] = kernels
        out["profiled_step_ms"] = round(prof_ms, 3)
    else:
        out["impl"] = impl
        out["config"]["note"] = ("GPU comparator: the same module graph over oracle/ (plain torch gather-mm-index_add sparse conv with "
                                 "host rulebooks, torch box attention, host scipy matching) on CUDA tensors — a labelled STAND-IN "
                                 "for the reference's spconv GPU path, which cannot be installed here; NOT spconv")
        out["gpu_launches"] = 0
    if not args.no_cpu_baseline and world == 1 and backend is None:  # rank 0 at N = 1 only
        out["cpu_baseline"] = cpu_baseline(args, steps=1)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_config1(args, dev):
    """BASELINE.json configs[0]: one 20k-point cloud -> voxelize -> SubMConv3d(16 -> 16, 3^3), batch 1, forward only;
    the GPU row, and the CPU row (reference numba-equivalent C voxelizer + oracle sparse conv) at 1 thread and all cores."""
    from efg_b200 import _lib, ops
    from efg_b200.data import WAYMO, make_scene
    from efg_b200.spconv import SparseConvTensor, SubMConv3d

    pts, _ = make_scene(args.points, WAYMO, seed=0)
    gpts = torch.from_numpy(pts).to(dev)
    offs = torch.tensor([0, pts.shape[0]], dtype=torch.int32, device=dev)
    conv = SubMConv3d(16, 16, 3, padding=1, bias=False, indice_key="c1").to(dev)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=torch.Generator().manual_seed(1)) * 0.05)
    proj = torch.randn(5, 16, generator=torch.Generator().manual_seed(0)).to(dev)   # 5 point features -> 16 channels
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        r = ops.hard_voxelize_batched(gpts, offs, WAYMO.voxel_size, WAYMO.pc_range, 5, 150000, coors_dim=4, want_voxels=False)
        m = int(r["counts"][-1].item())
        x = SparseConvTensor(r["mean"][:m] @ proj, r["coors"][:m].contiguous(), [41, 1504, 1504], 1)
        with torch.no_grad():
            return conv(x).features, m

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    launches0 = _lib.lib().efgb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        l2_flush.zero_()
        y, m = step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = _lib.lib().efgb_launch_count() - launches0
    out = {"metric": "scenes/sec voxelize+SubMConv3d(16->16) fwd", "value": round(1e3 / ms, 2), "unit": "scenes/s", "n_gpus": 1,
           "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": args.workload_name, "points": args.points, "voxels": m, "conv_precision": ops.CONV_PRECISION,
                      "l2": "flushed before every timed step"},
           "gpu_launches": int(launches)}
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = config1_cpu_rows(pts)
    print(json.dumps(out), flush=True)


def config1_cpu_rows(pts):
    """The CPU rows of BASELINE.md section 4a: C voxelizer (== the reference's numba loop, goldens) + oracle sparse conv,
    at 1 thread (the reference sets OMP_NUM_THREADS=1, cli/main.py:142) and at all cores."""
    from efg_b200.data import WAYMO
    from oracle import sparse_conv as sc
    from oracle import voxelize as ovox

    w = torch.randn(16, 3, 3, 3, 16, generator=torch.Generator().manual_seed(1)) * 0.05
    proj = torch.randn(5, 16, generator=torch.Generator().manual_seed(0))
    rows = {}
    for threads in (1, os.cpu_count() or 1):
        torch.set_num_threads(threads)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            v, c, n = ovox.hard_voxelize(pts, WAYMO.voxel_size, WAYMO.pc_range, 5, 150000)
            feats = torch.from_numpy(ovox.mean_vfe(v, n)) @ proj
            nbr = sc.subm_rulebook(np.pad(c, ((0, 0), (1, 0))), 1, [41, 1504, 1504], 3)
            sc.conv(feats, w, None, nbr)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        rows["threads_%d" % threads] = {"value": round(1.0 / best, 3), "unit": "scenes/s", "ms": round(best * 1e3, 2)}
    return {"value": rows["threads_%d" % (os.cpu_count() or 1)]["value"], "unit": "scenes/s", "cores": os.cpu_count() or 1,
            "kind": "port", "rows": rows, "sample": "best of 3 passes of the whole config (voxelize + rulebook + conv), oracle CPU path"}


# -------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle restatement; spconv itself is not installable here)
# -------------------------------------------------------------------------------------------------
def cpu_step_fn(n_points, seed=1):
    from efg_b200.config import voxel_detr_config
    from efg_b200.detectors.voxel_detr import VoxelDETR
    from oracle.backend_cpu import cpu_backend, voxelized_sample

    torch.manual_seed(0)
    cfg = voxel_detr_config(model={"device": "cpu", "transformer": {"num_queries": NUM_QUERIES}})
    model = VoxelDETR(cfg, backend=cpu_backend(), prune_unused=False).train()  # the reference evaluates every FPN level,
    model.stacked_losses = False      # ... every decoder layer's loss separately,
    model.reuse_proposal_head = False  # ... and the proposal head twice (VD/transformer.py:56, VD/voxel_detr.py:145)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01, betas=(0.9, 0.99), eps=1e-9)
    scenes = make_scenes(1, n_points, seed=seed)

    def step():
        opt.zero_grad(set_to_none=True)
        # the reference voxelizes on the host (numba loop == oracle C loop), then collates
        batch = [(voxelized_sample(p, cfg.dataset), {"annotations": a}) for p, a in scenes]
        losses = model(batch)
        total = sum(v for k, v in losses.items() if k.startswith("loss"))
        total.backward()
        opt.step()
        return float(total)

    return step


def cpu_baseline(args, steps=1):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_fn(args.cpu_points)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": round(steps / dt, 5), "unit": "scenes/s", "cores": cores, "kind": "port",
            "sample": "%d step(s) of ONE scene (%d pts, full 1504x1504x40 grid, 300 queries), oracle CPU backend, "
                      "torch threads=%d, %.1f s" % (steps, args.cpu_points, cores, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_fn(args.cpu_points)
    warm = min(args.warmup, 1)
    for _ in range(warm):
        step()
    steps = max(1, min(args.steps, 3))  # bounded: each step is tens of seconds of host work
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = steps / dt
    out = {
        "impl": "reference", "metric": "scenes/sec Voxel-DETR fwd+bwd", "value": round(value, 5), "unit": "scenes/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": round(dt / steps * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "scenes_per_step": 1, "points_per_scene": args.cpu_points,
                   "num_queries": NUM_QUERIES, "grid": "1504x1504x40", "step": "voxelize+fwd+bwd+adamw",
                   "note": "reference algorithm on host CPU: oracle restatement (spconv is not installable here); "
                           "each step is a bounded sample of one scene"},
        "cpu_baseline": {"value": round(value, 5), "unit": "scenes/s", "cores": cores, "kind": "port",
                         "sample": "%d step(s) x 1 scene, %.1f s" % (steps, dt)},
        "e2e": {"value": round(value, 5), "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "torch_gpu":
        from oracle.backend_cpu import cpu_backend  # the GPU stand-in comparator runs the oracle ops on CUDA tensors

        a.no_cpu_baseline = True
        run_efgb200(a, backend=cpu_backend())
    else:
        run_efgb200(a)

#!/usr/bin/env python
"""bench.py — MSDeformAttn fwd+bwd throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

Workloads (all M=8, D=32, L=4, P=4, synthetic tensors, DINO 4-scale R50 pyramid):
  encoder6   (default; BASELINE configs[1]) six deformable-encoder layers' MSDeformAttn, bs=2 per GPU,
             Lq=S=22,223, fp32, encoder-realistic locations "E"; one step = 6 forwards then 6 backwards
             on six distinct input sets (960 MB > L2, so no layer finds its inputs cached).
  decoder6   (configs[2]) six decoder cross-attention layers, bs=2, Lq=1100, bf16 value, locations "Dn".
  encoder1_hr1333 / encoder1_hr2000   (configs[4]) one encoder layer at 1333x1333 / 1600x2000.
  encoder_layer_ddp   (configs[3]) the whole encoder-layer train step under DDP.
Add --deterministic for the bitwise-reproducible grad_value mode, --graph to replay the step from a CUDA graph.

The DEFAULT invocation (no --workload) times encoder6 for the headline line and then appends, in the same JSON
line, "workloads": every other BASELINE config (decoder6 eager and graph-replayed, the two high-resolution
shapes, atomic and deterministic, and the DDP encoder-layer step — on every N, 1 included), each with its own
value / ms_per_step / roofline / clocks, and "legacy_cuda": the reference's own CUDA kernels (recompiled for
sm_100a, oracle/_ref) timed on the same B200 and the same inputs — a comparator leg outside the product's timed
region — and "aux_passes": the HBM-bound passes either side of the op (residual + LayerNorm, value preparation,
proposals, query selection) at the headline shape, each against its own roofline (rank 0 only).
--no-extra skips all three.

One process per GPU (torchrun for N>1); the op never communicates (images are independent), so ranks
only meet at the barriers around the timed region; value = queries processed by all ranks / max-over-ranks
device time.  Prints ONE JSON line on rank 0.

--impl reference times the reference's CPU path for this op — ms_deform_attn_core_pytorch (grid_sample)
forward + autograd backward, restated in oracle/msda_oracle.py because /root/reference does not travel to
the GPU box — on the host cores, rank 0 only.  It never imports the richsem_b200 package (whose import loads
the CUDA library): the synthetic-input module is loaded by file path.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path
from types import SimpleNamespace

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "MSDeformAttn fwd+bwd queries/s & % memory roofline, DINO 4-scale R50 800x1333"
UNIT = "queries/s"

WORKLOADS = {
    # name: (image hw, layers, batch/GPU, kind, Lq (None = S), value dtype)
    "encoder6": ((800, 1333), 6, 2, "E", None, "f32"),
    "encoder1": ((800, 1333), 1, 2, "E", None, "f32"),
    "encoder6_bf16": ((800, 1333), 6, 2, "E", None, "bf16"),
    "decoder6": ((800, 1333), 6, 2, "Dn", 1100, "bf16"),
    "decoder6_f32": ((800, 1333), 6, 2, "Dn", 1100, "f32"),
    "encoder1_hr1333": ((1333, 1333), 1, 2, "E", None, "f32"),
    "encoder1_hr2000": ((1600, 2000), 1, 2, "E", None, "f32"),
    # config 4: whole encoder-layer train step (MSDeformAttn + Linear projections + FFN), DDP all-reduce
    "encoder_layer_ddp": ((800, 1333), 1, 2, "layer", None, "f32"),
    # the whole 6-layer deformable encoder (MSDeformAttn + Linears + LayerNorm + FFN), forward + backward,
    # captured in one CUDA graph; --eager replays it launch by launch instead
    "encoder_stack6": ((800, 1333), 6, 2, "stack", None, "f32"),
}
# what the default invocation appends to the headline line: key -> (workload, options)
EXTRA = [
    ("encoder6_graph", "encoder6", {"graph": True}),  # the headline step replayed from one CUDA graph (no events inside)
    ("decoder6", "decoder6", {}),
    ("decoder6_graph", "decoder6", {"graph": True}),
    ("encoder1_hr1333", "encoder1_hr1333", {}),
    ("encoder1_hr2000", "encoder1_hr2000", {}),
    ("encoder1_hr2000_deterministic", "encoder1_hr2000", {"deterministic": True}),
]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="encoder6", choices=sorted(WORKLOADS))
    ap.add_argument("--deterministic", action="store_true")
    ap.add_argument("--lib-flags", type=lambda x: int(x, 0), default=0,
                    help="extra MSDA_FLAG_* bits for forward and backward (experiments)")
    ap.add_argument("--kernel", type=int, default=0, help="msda_opts.kernel_hint (MSDA_KERNEL_*; experiments)")
    ap.add_argument("--graph", action="store_true", help="op workloads: capture one step (all forwards + backwards) in a "
                    "CUDA graph and replay it; matters for decoder-sized calls, whose kernels (14-36 us) are shorter "
                    "than the Python launch path.  Per-launch times are then not available: the roofline is the step's")
    ap.add_argument("--eager", action="store_true", help="encoder_stack6: no CUDA graph")
    ap.add_argument("--fuse-prologue", action="store_true", help="encoder_stack6: fused softmax + "
                    "sampling-location prologue (SURVEY 8f-1)")
    ap.add_argument("--no-fuse-epilogue", action="store_true", help="encoder_stack6: residual + LayerNorm as the two "
                    "PyTorch kernels instead of the library's one-pass kernel")
    ap.add_argument("--tf32", action="store_true", help="encoder_stack6: let the PyTorch Linear layers use TF32 tensor "
                    "cores (torch.backends.cuda.matmul.allow_tf32); the default is the reference's strict fp32")
    ap.add_argument("--padding", action="store_true", help="encoder_stack6: image 1 of each pair is padded (mask path)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the additional workloads / comparator leg of the default line")
    ap.add_argument("--e2e-steps", type=int, default=12)
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_synthetic():
    """richsem_b200/synthetic.py loaded by PATH: importing the package would dlopen libmsda_b200.so, which the CPU
    reference arm must not do (the module itself needs torch only)."""
    spec = importlib.util.spec_from_file_location("_msda_synthetic", ROOT / "richsem_b200" / "synthetic.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload_config(name, syn):
    """The workload-defining part of `config`, identical in both arms (product and --impl reference)."""
    hw, layers, bs, kind, lq, vdt = WORKLOADS[name]
    shapes = syn.level_shapes(*hw)
    S = sum(h * w for h, w in shapes)
    return {"workload": name, "image": f"{hw[0]}x{hw[1]}", "levels": [list(s) for s in shapes], "S": S,
            "Lq": S if lq is None else lq, "batch_per_gpu": bs, "layers_per_step": layers, "heads": 8, "head_dim": 32,
            "points": 4, "locations": {"layer": "E", "stack": "E"}.get(kind, kind), "value_dtype": vdt,
            "queries_per_step_per_gpu": layers * bs * (S if lq is None else lq),
            "unit_of_value": "queries/s = queries processed / time; a product step processes queries_per_step_per_gpu "
                             "queries per GPU, a step of the CPU reference arm a bounded sample of them (one layer of "
                             "one image) — the per-query rate is what both arms report",
            "l2_policy": f"{layers} distinct input set(s) per step, visited round-robin; a step's inputs and outputs "
                         "exceed the 126 MB L2, so no layer finds its operands cached; no explicit flush"}


# --------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background during the timed region)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, interval_ms=50):
        self.index, self.proc, self.lines, self.interval = index, None, [], interval_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", str(self.interval)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=2.0):
        """Blocks until nvidia-smi has delivered its first sample, so that a short timed region is not over before
        the sampler runs."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, since=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.06)  # let a sample taken at the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if since is not None and ts < since:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU path (grid_sample), bounded sample
# --------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, hw=(800, 1333), n_sets=1):
    """One step = one encoder layer of ONE image (bs=1), fp32, fwd + autograd bwd of the reference's CPU path
    (BASELINE configs[0]) on all host cores; steps visit `n_sets` distinct input sets round-robin."""
    import torch

    from oracle.msda_oracle import core_pytorch_fwd_bwd

    syn = load_synthetic()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    shapes = syn.level_shapes(*hw)
    sets = [syn.make_inputs("E", 1, shapes, "cpu", seed=1234 + k) for k in range(max(1, n_sets))]
    times = []
    for it in range(warmup + steps):
        i = sets[it % len(sets)]
        t0 = time.perf_counter()
        core_pytorch_fwd_bwd(i["value"], shapes, i["loc"], i["attw"], i["grad_out"])
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    qps = steps * sets[0]["Lq"] / total
    return dict(value=qps, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"{steps} steps x (one encoder layer of one image: bs=1, Lq=S={sets[0]['Lq']}, fp32, fwd + autograd "
                       f"bwd of the grid_sample formulation = ms_deform_attn_core_pytorch restated in oracle/, "
                       f"bit-identical), {len(sets)} input set(s) round-robin, {warmup} warm-up, {total:.1f} s timed",
                ms_per_step=1e3 * total / steps)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    syn = load_synthetic()
    # the driver's K and W are honoured; a step is a bounded sample (one layer of one image, ~0.3 s on 16 cores), so
    # that K = 20 ends within seconds; a very large K is capped to keep the run within minutes
    steps, warmup = max(1, min(args.steps, 400)), max(0, min(args.warmup, 20))
    hw, layers, bs, kind, lq, vdt = WORKLOADS[args.workload]
    hw = hw if kind == "E" else (800, 1333)
    r = cpu_reference_run(steps, warmup, hw, n_sets=layers)
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, syn),
        "arm": {"parallelism": "rank 0 only, all host cores", "grad_value_mode": "autograd of grid_sample",
                "step": "a bounded sample of the workload's step: ONE encoder layer of ONE image per step (the CPU "
                        "path needs ~0.3 s for it); value is the per-query rate, comparable across arms"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# multi-rank helpers (batch sharding: no collective on the data path, ranks meet only here)
# --------------------------------------------------------------------------------------------
def max_over_ranks(ms, world, device):
    """Device time of the slowest rank (the contract's max-over-ranks)."""
    import torch
    import torch.distributed as dist

    if world <= 1:
        return float(ms)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def rank_seed(base, rank, layer):
    """Every rank / layer gets its own synthetic shard (weak scaling: per-GPU work is fixed)."""
    return base + 100 * rank + layer


def aggregate_qps(queries_per_rank_step, world, ms_per_step):
    return world * queries_per_rank_step / (ms_per_step * 1e-3)


def sync_all(c):
    if c.world > 1:
        c.dist.barrier()
    c.torch.cuda.synchronize()


# --------------------------------------------------------------------------------------------
# B200 arm: one op workload
# --------------------------------------------------------------------------------------------
def time_op_workload(c, name, steps, warmup, graph=False, deterministic=False, lib_flags=0, kernel=0, want_e2e=False,
                     e2e_steps=12, min_seconds=0.0, keep_sets=None):
    """Times `steps` steps of an op workload on every rank.  min_seconds > 0 (the additional workloads of the default
    line): the step count is raised so that the timed region lasts at least that long — their steps are a few hundred
    microseconds, and a region that short would be over before nvidia-smi delivers one clocks sample."""
    torch, ext, _capi, syn = c.torch, c.ext, c._capi, c.syn
    hw, layers, bs, kind, lq, vdt = WORKLOADS[name]
    shapes = syn.level_shapes(*hw)
    tdt = torch.bfloat16 if vdt == "bf16" else torch.float32
    sets = [syn.make_inputs(kind, bs, shapes, c.dev, seed=rank_seed(1234, c.rank, i), lq=lq, dtype=tdt) for i in range(layers)]
    S, Lq = sets[0]["S"], sets[0]["Lq"]
    shp, st = sets[0]["shapes"], sets[0]["starts"]
    queries_per_step = layers * bs * Lq
    flags = (_capi.FLAG_DETERMINISTIC if deterministic else 0) | lib_flags
    vb = 2 if vdt == "bf16" else 4
    fwd_bytes, bwd_bytes = syn.algorithmic_bytes(bs, S, Lq, value_bytes=vb, out_bytes=vb)
    in_bytes = sum(s[k].numel() * s[k].element_size() for s in sets for k in ("value", "loc", "attw", "grad_out"))

    def eager_step(timing=None):
        for i, s in enumerate(sets):
            if timing is not None:
                timing["f0"][i].record()
            s["out"] = ext.ms_deform_attn_forward(s["value"], shp, st, s["loc"], s["attw"], 64, _flags=lib_flags, _kernel=kernel)
            if timing is not None:
                timing["f1"][i].record()
        for i in reversed(range(layers)):
            s = sets[i]
            if timing is not None:
                timing["b0"][i].record()
            s["grads"] = ext.ms_deform_attn_backward(s["value"], shp, st, s["loc"], s["attw"], s["grad_out"], 64,
                                                     _flags=flags, _kernel=kernel)
            if timing is not None:
                timing["b1"][i].record()

    step = eager_step
    warmup = max(warmup, 3)
    for _ in range(warmup):
        step()
    sync_all(c)
    launches_per_step = None
    if graph:
        l0 = _capi.launch_count()
        step()
        launches_per_step = _capi.launch_count() - l0
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        step = lambda timing=None: g.replay()
        for _ in range(3):
            step()
        sync_all(c)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    if min_seconds > 0:
        a, b = ev(), ev()
        a.record()
        for _ in range(3):
            step()
        b.record()
        torch.cuda.synchronize()
        est = max_over_ranks(a.elapsed_time(b), c.world, c.dev) / 3 * 1e-3
        steps = max(steps, int(math.ceil(min_seconds / max(est, 1e-6))))
    # Per-layer CUDA events inside the timed region give the fwd / bwd split of every launch — unless the calls are so
    # short (decoder layers: 14-36 us kernels) that recording four events per layer would itself slow the step down;
    # those workloads time the plain loop and take the split from a second, instrumented pass.
    est_probe = (ev(), ev())
    est_probe[0].record()
    step()
    est_probe[1].record()
    torch.cuda.synchronize()
    inner_events = not graph and est_probe[0].elapsed_time(est_probe[1]) / (2 * layers) > 0.1
    mk = lambda n: [{k: [ev() for _ in range(layers)] for k in ("f0", "f1", "b0", "b1")} for _ in range(n)]
    timings = mk(steps) if inner_events else None
    start, stop = ev(), ev()
    sampler = None
    if c.rank == 0:
        sampler = ClockSampler(c.dev.index).start()
        sampler.wait_first()
    t_region = time.perf_counter()
    launches0 = _capi.launch_count()
    sync_all(c)
    start.record()
    for k in range(steps):
        step(timings[k] if inner_events else None)
    stop.record()
    sync_all(c)
    launches = _capi.launch_count() - launches0 if not graph else launches_per_step * steps
    clocks = sampler.stop(since=t_region) if sampler is not None else None
    ms_per_step = max_over_ranks(start.elapsed_time(stop), c.world, c.dev) / steps
    value = aggregate_qps(queries_per_step, c.world, ms_per_step)

    split_note = "CUDA events around every launch inside the timed region"
    if graph:
        # no events inside a replayed graph: split the step in the ratio of the algorithmic bytes (reported as such)
        per_layer = ms_per_step / layers
        fwd_ms = per_layer * fwd_bytes / (fwd_bytes + bwd_bytes)
        bwd_ms = per_layer - fwd_ms
        split_note = "graph replay: step time split by algorithmic bytes"
    else:
        if not inner_events:
            timings = mk(min(steps, 20))
            for tm in timings:
                eager_step(tm)
            torch.cuda.synchronize()
            split_note = "fwd / bwd ratio from a second, instrumented pass, scaled to the timed region's step time"
        fwd_ms = statistics.mean(tm["f0"][i].elapsed_time(tm["f1"][i]) for tm in timings for i in range(layers))
        bwd_ms = statistics.mean(tm["b0"][i].elapsed_time(tm["b1"][i]) for tm in timings for i in range(layers))
        if not inner_events:
            scale = (ms_per_step / layers) / (fwd_ms + bwd_ms)
            fwd_ms, bwd_ms = fwd_ms * scale, bwd_ms * scale
    peak, peak_src = peaks()

    e2e = None
    if want_e2e and not graph:  # e2e is measured on the eager path (the public call with host buffers)
        e2e = run_e2e(c, sets, shp, st, flags, queries_per_step, e2e_steps)

    dom = "backward" if bwd_ms >= fwd_ms else "forward"
    dom_bytes, dom_ms = (bwd_bytes, bwd_ms) if dom == "backward" else (fwd_bytes, fwd_ms)
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(name, {}).get(dom)
        except Exception:
            traffic = None
    arm = {"grad_value_mode": "deterministic (canonical order in a block, 64-bit fixed-point accumulation across blocks)"
           if deterministic else "atomic (merged on chip per window cell, then fp32 L2 reductions)",
           "parallelism": f"batch-sharded x{c.world}, no collective in the op",
           "launch": "one CUDA graph per step, replayed (per-launch times not measured: fwd / bwd split by "
                     "algorithmic bytes)" if graph else "eager, one library call per layer and direction",
           "input_bytes_per_step": in_bytes, "fwd_bwd_split": split_note}
    res = {
        "value": value, "unit": UNIT, "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup, "dtype": vdt,
        "config": workload_config(name, syn), "arm": arm,
        "roofline": {"bound": "hbm", "kernel": f"msda {dom} ({'zero-fill of grad_value + ' if dom == 'backward' else ''}kernel), avg of "
                     f"{steps * layers} launches", "achieved": dom_bytes / (dom_ms * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": dom_bytes / (dom_ms * 1e-3) / 1e9 / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms},
        "roofline_fwd_bwd": {"achieved": (fwd_bytes + bwd_bytes) * layers / (ms_per_step * 1e-3) / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": (fwd_bytes + bwd_bytes) * layers / (ms_per_step * 1e-3) / 1e9 / peak,
                             "fwd_ms_per_layer": fwd_ms, "bwd_ms_per_layer": bwd_ms,
                             "fwd_GBps": fwd_bytes / (fwd_ms * 1e-3) / 1e9, "bwd_GBps": bwd_bytes / (bwd_ms * 1e-3) / 1e9,
                             "bytes_per_query": (fwd_bytes + bwd_bytes) / (bs * Lq)},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if keep_sets is not None:
        keep_sets.extend(sets)
    return res


def compact(res):
    """What an additional workload contributes to the default line."""
    r, rf = res["roofline"], res["roofline_fwd_bwd"]
    return {"value": res["value"], "unit": res["unit"], "ms_per_step": res["ms_per_step"], "steps": res["steps"],
            "dtype": res["dtype"], "config": {k: res["config"][k] for k in ("workload", "image", "S", "Lq", "batch_per_gpu",
                                                                          "layers_per_step", "locations", "value_dtype")},
            "arm": {k: res["arm"][k] for k in ("grad_value_mode", "launch")},
            "roofline": {"bound": "hbm", "frac": rf["frac"], "achieved": rf["achieved"], "peak": rf["peak"], "unit": "GB/s",
                         "what": "fwd+bwd algorithmic bytes of the step / step time", "dominant_kernel_frac": r["frac"],
                         "fwd_ms_per_layer": rf["fwd_ms_per_layer"], "bwd_ms_per_layer": rf["bwd_ms_per_layer"]},
            "gpu_launches": res["gpu_launches"], "clocks": res["clocks"]}


# --------------------------------------------------------------------------------------------
# comparator leg: the reference's own CUDA kernels on the same B200 (oracle/_ref, never product code)
# --------------------------------------------------------------------------------------------
def run_legacy_cuda(c, sets, ours_fwd_ms, ours_bwd_ms):
    """Times the reference's ms_deform_im2col_cuda.cuh kernels — compiled unmodified for sm_100a by
    oracle/build_ref.py — on the headline workload's own input sets (bs=2 encoder layers, fp32), including the
    zero-fills its host wrapper performs (ms_deform_attn_cuda.cu:54,121-123).  Rank 0 only; runs after the product's
    timed region and shares nothing with it."""
    torch = c.torch
    try:
        from oracle.msda_oracle import LegacyCuda
    except Exception as e:  # pragma: no cover
        return {"unavailable": f"oracle import failed: {e}"}
    if not LegacyCuda.available():
        return {"unavailable": "oracle/_ref/libmsda_legacy.so is not built (the reference tree exists in the build container only)"}
    legacy = LegacyCuda()
    k = [0]

    def nxt():
        k[0] = (k[0] + 1) % len(sets)
        s = sets[k[0]]
        return (s["value"], s["shapes"], s["starts"], s["loc"], s["attw"]), s["grad_out"]

    def timeit(fn, iters=20, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    f_ms = timeit(lambda: legacy.forward(*nxt()[0]))

    def bwd():
        a, g = nxt()
        legacy.backward(*a, g)

    b_ms = timeit(bwd)
    n, lq = sets[0]["value"].shape[0], sets[0]["Lq"]
    return {"what": "reference CUDA kernels (ms_deform_im2col_cuda.cuh:237-299, 301-403) recompiled for sm_100a, same "
                    "B200, same inputs as the headline workload; comparator only",
            "fwd_ms_per_layer": f_ms, "bwd_ms_per_layer": b_ms, "value": n * lq / ((f_ms + b_ms) * 1e-3), "unit": UNIT,
            "ours_fwd_ms_per_layer": ours_fwd_ms, "ours_bwd_ms_per_layer": ours_bwd_ms,
            "speedup_fwd": f_ms / ours_fwd_ms, "speedup_bwd": b_ms / ours_bwd_ms,
            "speedup_fwd_bwd": (f_ms + b_ms) / (ours_fwd_ms + ours_bwd_ms), "input_sets": len(sets)}


def run_aux_passes(c, iters=48, n_sets=4):
    """The HBM-bound passes either side of the sampling core at the headline shape (bs = 2,
    S = 22,223, C = 256), each against its own roofline: algorithmic bytes / time / measured HBM peak.  Inputs rotate over
    n_sets distinct buffers (4 x 45.5 MB > the 126 MB L2); one round over all sets is captured in a CUDA graph and the
    graph is replayed, so the figures are device time (the passes take 10-40 us, less than a Python launch)."""
    torch, _capi, syn = c.torch, c._capi, c.syn
    from richsem_b200.MultiScaleDeformableAttention import _stream
    from richsem_b200.ops.functions import gen_encoder_output_proposals, topk_proposals
    from richsem_b200.ops.functions.aux_functions import cast_value_bf16, class_scores

    dev = c.dev
    shapes = syn.level_shapes(800, 1333)
    shp, _, S = syn.level_tensors(shapes, dev)
    n, ch = 2, 256
    rows = n * S
    full = rows * ch * 4
    pk = peaks()[0]
    gen = torch.Generator(device=dev).manual_seed(99)
    xs = [torch.randn(n, S, ch, generator=gen, device=dev) for _ in range(n_sets)]
    rs = [torch.randn(n, S, ch, generator=gen, device=dev) for _ in range(n_sets)]
    parts = []
    for i in range(n):  # image 1: right 30 % / bottom 20 % of every level is padding
        fh, fw = (1.0, 1.0) if i % 2 == 0 else (0.8, 0.7)
        lv = []
        for h, w in shapes:
            m = torch.ones(h, w, dtype=torch.bool)
            m[: max(1, round(h * fh)), : max(1, round(w * fw))] = False
            lv.append(m.reshape(-1))
        parts.append(torch.cat(lv))
    mask = torch.stack(parts).to(dev)
    masked = int(mask.sum())

    def timeit(fns):
        for f in fns:
            f()
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                keep = [f() for f in fns]
        torch.cuda.current_stream(dev).wait_stream(side)
        rounds = max(1, iters // len(fns))
        g.replay()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(rounds):
            g.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        del keep, g
        return e0.elapsed_time(e1) / (rounds * len(fns))

    out = {}

    def put(name, ms, alg, ref):
        gbs = alg / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "algorithmic_bytes": alg, "GBps": gbs, "frac": gbs / pk, "reference": ref}

    l0 = _capi.launch_count()
    # residual add + LayerNorm (deformable_transformer.py:871-872), forward and backward through the C ABI
    gamma, beta = torch.rand(ch, device=dev) + 0.5, torch.randn(ch, device=dev) * 0.1
    outs = [torch.empty_like(x) for x in xs]
    stats = torch.empty(2, rows, device=dev)

    def ln_fwd(k):
        _capi.check(_capi.lib.msda_add_layernorm_f32(_stream(dev), xs[k].data_ptr(), rs[k].data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                                     rows, ch, 1e-5, outs[k].data_ptr(), stats[0].data_ptr(), stats[1].data_ptr()),
                    "msda_add_layernorm_f32")

    put("add_layernorm_fwd", timeit([lambda k=k: ln_fwd(k) for k in range(n_sets)]), 3 * full + rows * 8,
        "src = norm(src + src2), deformable_transformer.py:871-872")
    ws_bytes = _capi.lib.msda_add_layernorm_workspace_bytes(rows, ch)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    gw, gb = torch.empty(ch, device=dev), torch.empty(ch, device=dev)

    def ln_bwd(k):  # grad_out = another set's residual; the statistics of the last forward serve every set (timing only)
        _capi.check(_capi.lib.msda_add_layernorm_backward_f32(
            _stream(dev), rs[(k + 1) % n_sets].data_ptr(), xs[k].data_ptr(), rs[k].data_ptr(), gamma.data_ptr(), stats[0].data_ptr(),
            stats[1].data_ptr(), rows, ch, outs[k].data_ptr(), gw.data_ptr(), gb.data_ptr(), ws.data_ptr(), ws_bytes),
            "msda_add_layernorm_backward_f32")

    put("add_layernorm_bwd", timeit([lambda k=k: ln_bwd(k) for k in range(n_sets)]), 4 * full + 2 * rows * 8,
        "its backward incl. weight / bias gradients (two kernels)")
    del outs, rs
    # value preparation (ms_deform_attn.py:94-97): mask zeroing + bf16 cast
    put("value_prepare_bf16", timeit([lambda x=x: cast_value_bf16(x, mask) for x in xs]),
        (rows - masked) * ch * 4 + rows + rows * ch * 2, "value.masked_fill(mask, 0).to(bf16), ms_deform_attn.py:94-97")
    # two-stage proposals (utils.py:10-65)
    keep_rows = int(torch.isfinite(gen_encoder_output_proposals(xs[0], mask, shp)[1][..., 0]).sum())
    put("encoder_proposals", timeit([lambda x=x: gen_encoder_output_proposals(x, mask, shp) for x in xs]),
        keep_rows * ch * 4 + rows + rows * ch * 4 + rows * 16, "gen_encoder_output_proposals, utils.py:10-65")
    del xs
    # query selection (deformable_transformer.py:367-369), 91 classes
    ls = [torch.randn(n, S, 91, generator=gen, device=dev) for _ in range(n_sets)]
    put("class_scores_K91", timeit([lambda x=x: class_scores(x) for x in ls]), rows * 91 * 4 + rows * 4,
        "enc_outputs_class.max(-1)[0], deformable_transformer.py:369")
    put("topk_proposals_K91_k900", timeit([lambda x=x: topk_proposals(x, 900) for x in ls]),
        rows * 91 * 4 + 2 * rows * 4 + n * 900 * 8, "torch.topk(scores, 900, dim=1)[1], :367-369 (latency-bound: one cluster per image)")
    out["_note"] = ("bs=2, S=22223, C=256; CUDA-graph replay over %d rotating buffer sets (larger than L2); frac = algorithmic "
                    "bytes / time / measured HBM peak" % n_sets)
    out["_launches_per_round"] = int(_capi.launch_count() - l0)
    return out


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import richsem_b200  # noqa: F401
    from richsem_b200 import MultiScaleDeformableAttention as ext
    from richsem_b200 import _capi, synthetic as syn

    c = SimpleNamespace(torch=torch, dist=dist, rank=rank, world=world, dev=dev, ext=ext, _capi=_capi, syn=syn)
    kind = WORKLOADS[args.workload][3]
    if kind == "layer":
        r = run_encoder_layer_ddp(c, args.steps, args.warmup)
        if rank == 0:
            print(json.dumps(ddp_line(r, args)))
    elif kind == "stack":
        run_encoder_stack(args, c)
    else:
        default_line = args.workload == "encoder6" and not (args.no_extra or args.graph or args.deterministic
                                                             or args.lib_flags or args.kernel)
        sets = [] if default_line else None
        res = time_op_workload(c, args.workload, args.steps, args.warmup, graph=args.graph, deterministic=args.deterministic,
                               lib_flags=args.lib_flags, kernel=args.kernel, want_e2e=not args.no_e2e,
                               e2e_steps=args.e2e_steps, keep_sets=sets)
        extras, legacy, aux = None, None, None
        if default_line:
            if rank == 0:
                legacy = run_legacy_cuda(c, sets[:4], res["roofline_fwd_bwd"]["fwd_ms_per_layer"],
                                         res["roofline_fwd_bwd"]["bwd_ms_per_layer"])
            del sets
            sync_all(c)
            extras = {}
            for key, name, opt in EXTRA:
                torch.cuda.empty_cache()
                extras[key] = compact(time_op_workload(c, name, min(args.steps, 50), args.warmup, min_seconds=0.25, **opt))
            torch.cuda.empty_cache()
            extras["encoder_layer_ddp"] = run_encoder_layer_ddp(c, min(args.steps, 50), args.warmup)
            torch.cuda.empty_cache()
            aux = run_aux_passes(c) if rank == 0 else None
            sync_all(c)
        if rank == 0:
            cpu_baseline = None
            if world == 1 and not args.no_cpu_baseline:
                r = cpu_reference_run(steps=12, warmup=2)
                cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line = {
                "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": res["steps"],
                "warmup": res["warmup"], "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": res["dtype"], "data": "synthetic", "config": res["config"], "arm": res["arm"],
                "roofline": res["roofline"], "roofline_fwd_bwd": res["roofline_fwd_bwd"], "cpu_baseline": cpu_baseline,
                "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": res["clocks"], "lib": _capi.build_info(),
            }
            if extras is not None:
                line["workloads"] = extras
                line["legacy_cuda"] = legacy
                line["aux_passes"] = aux
            print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_encoder_layer_ddp(c, steps, warmup):
    """BASELINE config 4: per rank one reference-equivalent deformable encoder layer (d_ffn 2048, relu,
    dropout 0; deformable_transformer.py:825-881) on its own bs=2 shard; loss = out.square().mean(); forward +
    backward with DDP's bucketed NCCL all-reduce of the 1,282,176 fp32 parameters (main.py:204-206); no optimizer
    step.  N = 1 runs the same step without the wrapper.  The same model is also timed WITHOUT the
    DDP wrapper on every rank, so that the line names what the all-reduce costs."""
    torch, _capi, syn = c.torch, c._capi, c.syn
    from torch.nn.parallel import DistributedDataParallel as DDP

    from richsem_b200.encoder_layer import DeformableEncoderLayer, encoder_reference_points

    hw, _, bs, _, _, _ = WORKLOADS["encoder_layer_ddp"]
    shapes = syn.level_shapes(*hw)
    shp, st, S = syn.level_tensors(shapes, c.dev)
    torch.manual_seed(1234)
    layer = DeformableEncoderLayer().to(c.dev)
    with torch.no_grad():  # leave the degenerate init so that every gradient path does real work
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
        layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    n_params = sum(p.numel() for p in layer.parameters())
    gen = torch.Generator(device=c.dev).manual_seed(4321 + c.rank)
    src = torch.randn(bs, S, 256, generator=gen, device=c.dev)
    pos = torch.randn(bs, S, 256, generator=gen, device=c.dev)
    ref = encoder_reference_points(shapes, bs, c.dev)

    def make_step(model):
        def step():
            model.zero_grad(set_to_none=True)
            out = model(src, pos, ref, shp, st, None)
            out.square().mean().backward()
        return step

    def timed(step, n, sample_clocks):
        for _ in range(max(warmup, 3)):
            step()
        sync_all(c)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = None
        if sample_clocks and c.rank == 0:
            sampler = ClockSampler(c.dev.index).start()
            sampler.wait_first()
        t_region = time.perf_counter()
        l0 = _capi.launch_count()
        sync_all(c)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        sync_all(c)
        launches = _capi.launch_count() - l0
        clocks = sampler.stop(since=t_region) if sampler is not None else None
        return max_over_ranks(e0.elapsed_time(e1), c.world, c.dev) / n, launches, clocks

    steps = max(steps, 10)
    local_ms, _, _ = timed(make_step(layer), steps, False)          # no wrapper: what one GPU does alone
    model = DDP(layer, device_ids=[c.dev.index]) if c.world > 1 else layer
    ms, launches, clocks = timed(make_step(model), steps, True)
    fwd_bytes, bwd_bytes = syn.algorithmic_bytes(bs, S, S)
    peak, _ = peaks()
    return {"value": c.world * bs * S / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "dtype": "f32",
            "config": {"workload": "encoder_layer_ddp", "image": f"{hw[0]}x{hw[1]}", "S": S, "batch_per_gpu": bs,
                       "params": n_params,
                       "step": "encoder layer fwd + bwd (MSDeformAttn and residual + LayerNorm on libmsda_b200, "
                               "Linears / FFN in PyTorch fp32), no optimizer"},
            "arm": {"parallelism": f"DDP x{c.world}: bucketed NCCL all-reduce of {n_params * 4 / 1e6:.2f} MB of fp32 "
                                   "gradients per step" if c.world > 1 else "1 GPU: the same step, no wrapper, no collective"},
            "ms_per_step_without_ddp_wrapper": local_ms, "allreduce_exposed_ms": ms - local_ms,
            "limited_by": "the step is ~95 % strict-fp32 cuBLAS GEMMs (FFN 256<->2048 and the four projections); the "
                          "all-reduce overlaps the backward except for its last bucket (allreduce_exposed_ms)",
            "roofline": {"bound": "hbm", "frac": (fwd_bytes + bwd_bytes) / (ms * 1e-3) / 1e9 / peak, "unit": "GB/s",
                         "achieved": (fwd_bytes + bwd_bytes) / (ms * 1e-3) / 1e9, "peak": peak,
                         "what": "MSDeformAttn fwd+bwd algorithmic bytes / WHOLE step time (the op is a few % of this step)"},
            "gpu_launches": int(launches), "clocks": clocks}


def ddp_line(r, args):
    return {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": r["steps"], "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            **{k: r[k] for k in ("config", "arm", "ms_per_step_without_ddp_wrapper", "allreduce_exposed_ms", "limited_by",
                                 "roofline", "gpu_launches", "clocks")}}


def run_encoder_stack(args, c):
    """per rank the 6-layer deformable encoder (deformable_transformer.py:470-618, 825-881; d_ffn 2048,
    relu, dropout 0) on its own bs=2 shard, forward + backward of loss = out.square().mean(), captured once in a CUDA
    graph and replayed (or launched eagerly with --eager).  No optimizer, no collective (single-rank glue measurement;
    under torchrun every rank runs its own replica and the slowest one counts)."""
    torch, _capi, syn = c.torch, c._capi, c.syn
    rank, world, dev = c.rank, c.world, c.dev
    from richsem_b200.encoder_layer import DeformableEncoder, GraphedTrainStep

    hw, layers, bs, _, _, _ = WORKLOADS[args.workload]
    if args.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    shapes = syn.level_shapes(*hw)
    shp, st, S = syn.level_tensors(shapes, dev)
    torch.manual_seed(1234)
    model = DeformableEncoder(layers, fuse_prologue=args.fuse_prologue, fuse_epilogue=not args.no_fuse_epilogue).to(dev)
    with torch.no_grad():  # leave the degenerate init so that every gradient path does real work
        for layer in model.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
            layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    src = torch.randn(bs, S, 256, generator=gen, device=dev)
    pos = torch.randn(bs, S, 256, generator=gen, device=dev)
    valid = torch.ones(bs, len(shapes), 2, device=dev)
    mask = None
    if args.padding:
        rows = []
        for i in range(bs):
            fh, fw = (1.0, 1.0) if i % 2 == 0 else (0.8, 0.7)
            parts = []
            for h, w in shapes:
                m = torch.ones(h, w, dtype=torch.bool)
                m[: max(1, round(h * fh)), : max(1, round(w * fw))] = False
                parts.append(m.reshape(-1))
            rows.append(torch.cat(parts))
            valid[i, :, 0], valid[i, :, 1] = fw, fh
        mask = torch.stack(rows).to(dev)
    loss_fn = lambda out: out.square().mean()
    call_args = (src, pos, shp, st, valid, mask)

    def eager_step():
        for p in model.parameters():
            p.grad = None
        loss_fn(model(*call_args)).backward()

    l0 = _capi.launch_count()
    eager_step()
    launches_per_step = _capi.launch_count() - l0
    if args.eager:
        step = eager_step
    else:
        graphed = GraphedTrainStep(model, loss_fn, call_args)
        step = lambda: graphed.graph.replay()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all(c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = None
    if rank == 0:
        sampler = ClockSampler(dev.index).start()
        sampler.wait_first()
    t_region = time.perf_counter()
    l0 = _capi.launch_count()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all(c)
    launches = _capi.launch_count() - l0
    clocks = sampler.stop(since=t_region) if sampler is not None else None
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev) / args.steps
    if rank == 0:
        n_params = sum(p.numel() for p in model.parameters())
        print(json.dumps({
            "metric": METRIC, "value": world * layers * bs * S / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "image": f"{hw[0]}x{hw[1]}", "S": S, "batch_per_gpu": bs,
                       "layers_per_step": layers, "params": n_params, "mode": "eager" if args.eager else "cuda graph",
                       "fuse_prologue": bool(args.fuse_prologue), "fuse_epilogue": not args.no_fuse_epilogue, "padding_mask": bool(args.padding),
                       "linear_precision": "tf32" if args.tf32 else "fp32 (reference default)",
                       "parallelism": f"replicas x{world}, no collective",
                       "step": "6-layer deformable encoder fwd + bwd (MSDeformAttn + elementwise neighbours on "
                               "libmsda_b200, residual + LayerNorm on libmsda_b200 unless fuse_epilogue is false, Linears / FFN in PyTorch), no optimizer"},
            # a replayed graph launches the library's kernels without passing through its host entry points:
            # the count below is what the capture recorded times the replays
            "gpu_launches": int(launches) if args.eager else int(launches_per_step * args.steps),
            "clocks": clocks, "lib": _capi.build_info()}))


def run_e2e(c, sets, shp, st, flags, queries_per_step, e2e_steps):
    """Host buffers in, host buffers out: per step every layer's value / locations / weights / grad_out
    are copied from pinned host memory, and the output and the three gradients are copied back.
    Copies run on their own streams so that transfers of neighbouring layers overlap the kernels."""
    torch, ext = c.torch, c.ext
    layers = len(sets)
    names_in = ("value", "loc", "attw", "grad_out")
    host_in = [{k: s[k].cpu().pin_memory() for k in names_in} for s in sets]
    # two generations of device input buffers: the upload of step k+1 runs while the kernels of step k
    # still read theirs, so the H2D and D2H links both stay busy across step boundaries
    dev_gen = [[{k: torch.empty_like(s[k]) for k in names_in} for s in sets] for _ in range(2)]
    gen_free = [None, None]  # event: the last kernels that read this generation have finished
    host_out = None
    h2d, d2h, comp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
    h2d_bytes = sum(t.numel() * t.element_size() for hi in host_in for t in hi.values())
    counter = [0]
    d2h_done = []

    def e2e_step():
        nonlocal host_out
        ev_in = []
        g = counter[0] & 1
        counter[0] += 1
        dev_in = dev_gen[g]
        if gen_free[g] is not None:
            h2d.wait_event(gen_free[g])
        # bounded pipeline depth: the HOST does not enqueue step k before the downloads of step k-2 have finished, so
        # the caching allocator cycles through a fixed set of output blocks (the op allocates its results) instead
        # of cudaMalloc-ing a new 956 MB set for every step the host runs ahead
        if len(d2h_done) >= 2:
            d2h_done[-2].synchronize()
        with torch.cuda.stream(h2d):
            for i in range(layers):
                for k in names_in:
                    dev_in[i][k].copy_(host_in[i][k], non_blocking=True)
                e = torch.cuda.Event(); e.record(h2d); ev_in.append(e)
        outs, grads = [None] * layers, [None] * layers
        ev_f, ev_b = [], [None] * layers
        for i in range(layers):
            comp.wait_event(ev_in[i])
            outs[i] = ext.ms_deform_attn_forward(dev_in[i]["value"], shp, st, dev_in[i]["loc"], dev_in[i]["attw"], 64)
            e = torch.cuda.Event(); e.record(comp); ev_f.append(e)
        for i in reversed(range(layers)):
            grads[i] = ext.ms_deform_attn_backward(dev_in[i]["value"], shp, st, dev_in[i]["loc"], dev_in[i]["attw"],
                                                   dev_in[i]["grad_out"], 64, _flags=flags)
            e = torch.cuda.Event(); e.record(comp); ev_b[i] = e
        gen_free[g] = ev_b[0]
        if host_out is None:
            host_out = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (outs[i], *grads[i])]
                        for i in range(layers)]
        with torch.cuda.stream(d2h):
            for i in range(layers):
                d2h.wait_event(ev_f[i])
                host_out[i][0].copy_(outs[i], non_blocking=True)
            for i in reversed(range(layers)):
                d2h.wait_event(ev_b[i])
                for j in range(3):
                    host_out[i][1 + j].copy_(grads[i][j], non_blocking=True)
            e = torch.cuda.Event(); e.record(d2h); d2h_done.append(e)
            del d2h_done[:-2]
        for i in range(layers):  # keep device results alive until the copies are ordered after them
            for t in (outs[i], *grads[i]):
                t.record_stream(d2h)

    for _ in range(4):
        e2e_step()
    sync_all(c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    comp.wait_stream(d2h)  # the timed region ends when the last result is in host memory
    comp.wait_stream(h2d)
    e1.record()
    sync_all(c)
    ms = max_over_ranks(e0.elapsed_time(e1), c.world, c.dev) / e2e_steps
    d2h_bytes = sum(t.numel() * t.element_size() for ho in host_out for t in ho)
    return {"value": c.world * queries_per_step / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
            "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": ms, "steps": e2e_steps,
            "note": "pinned host buffers -> H2D -> fwd/bwd through the drop-in API -> D2H of out + 3 grads, copies on side streams, device input buffers double-buffered across steps"}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

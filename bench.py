"""bench.py -- patch-pairs/s of the BiDateNet hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|infer|scene] [--impl reference]

One "step" = one pass of the hot path over one batch of 64 synthetic 13x256x256 patch pairs per GPU.

  train (default; BASELINE.json configs[2], and configs[3] under torchrun): forward + Tversky(0.1, 0.9) loss + backward +
        SGD step (reference train.py:88-95); for N > 1 the step contains the ONE NCCL all-reduce of gradients + BatchNorm
        running statistics (fabric_b200.distributed.DataParallelStep).  The JSON line also carries the eval forward
        (configs[1]) as the `infer` sub-record, the stock torch / cuDNN "library bar" on the same GPU
        (`library_baseline`), and the reference's CPU path (`cpu_baseline`).
  infer: the eval forward alone as the headline (reference train.py:193).
  scene: BASELINE configs[4], full-scene tiled inference sharded by row bands.

Prints ONE JSON line (DESIGN.md section "Measurement").  For N > 1 launch with torchrun; ranks shard pairs data-parallel
(weak scaling, 64 pairs per rank).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_GFLOP_PER_PAIR = 92.577      # SURVEY.md 8d, true Cin = 13
TRAIN_GFLOP_PER_PAIR = 275.8
PAIRS = 64
SIZE = 256
LR = 1e-3                        # metadata.json:41


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d["bf16_tflops"], tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(tflops=1590.0, tflops_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median of the samples under load (upper half)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference
def _cpu_reference_step(workload, batch):
    """One hot-path pass of the reference's CPU implementation on `batch` pairs of 13x256x256.  Uses the UNMODIFIED
    reference modules (oracle/_ref, or /root/reference in the build container; kind "reference") when they are present,
    else the restatement oracle/bidatenet_oracle.py (kind "port").  Returns (callable, kind, description)."""
    import torch
    from oracle import bidatenet_oracle as O
    from oracle import ref_loader
    sd = O.make_state_dict(seed=0)
    x1, x2, labels = O.make_inputs(batch, SIZE, seed=1)
    if ref_loader.available():
        RefNet, ref_metrics, _, root = ref_loader.load()
        model = RefNet(13, 2)
        model.load_state_dict(sd)
        crit = ref_metrics.TverskyLoss(alpha=0.1, beta=0.9)            # metadata.json:42-44, helpers.py:311-312
        opt = torch.optim.SGD(model.parameters(), lr=LR)               # train.py:55
        if workload == "train":
            model.train()

            def one():                                                  # train.py:88-95
                opt.zero_grad()
                loss = crit(model(x1, x2), labels)
                loss.backward()
                opt.step()
                return float(loss.detach())
        else:
            model.eval()

            def one():                                                  # train.py:193
                with torch.no_grad():
                    return model(x1, x2)
        where = "oracle/_ref" if root.endswith("_ref") else root
        return one, "reference", f"unmodified reference modules ({where}: models/bidate_model.py, utils/metrics.py), fp32 torch-CPU"
    crit = lambda l, t: O.tversky_loss(l, t, 0.1, 0.9)   # noqa: E731

    def one():
        if workload == "train":
            return O.train_step(x1, x2, labels, sd, crit)
        with torch.no_grad():
            return O.bidatenet_forward(x1, x2, sd, training=False)
    return one, "port", "oracle/bidatenet_oracle.py (restatement), fp32 torch-CPU"


def cpu_reference_throughput(workload, budget_s=12.0, batch=2, threads=None):
    """cpu_baseline: the reference's CPU path on all host cores, on a bounded sample of the same workload."""
    import torch
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    one, kind, what = _cpu_reference_step(workload, batch)
    one()  # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < 3 or (time.perf_counter() < t_end and len(times) < 50):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return dict(value=batch / med, unit="patch-pairs/s", cores=threads, kind=kind,
                sample=f"{len(times)} iterations of {batch} pairs 13x{SIZE}x{SIZE} ({workload}), median; {what}, {threads} threads")


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, on this arm's config;
    each step is a bounded sample (2 pairs) of the 64-pair workload -- per-pair CPU cost is flat in the batch size."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 2
    import torch
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    one, kind, what = _cpu_reference_step(args.workload, batch)
    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    v = batch * args.steps / dt
    sample = f"{args.steps} steps of {batch} pairs 13x{SIZE}x{SIZE} ({args.workload}); {what}"
    print(json.dumps({
        "impl": "reference", "metric": "patch-pairs/s (13x256x256)", "value": v, "unit": "patch-pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.workload, args.gpus),
        "cpu_baseline": {"value": v, "unit": "patch-pairs/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "patch-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def config_dict(workload, n):
    names = {"infer": "BiDateNet fwd-only inference",
             "train": "BiDateNet fwd+bwd training step, Tversky(0.1,0.9) loss, SGD" +
                      (", NCCL grad/BN-stat all-reduce" if n > 1 else "")}
    return {"workload": names[workload] + f", 13x{SIZE}x{SIZE}, batch {PAIRS} pairs per GPU",
            "pairs_per_gpu": PAIRS, "patch": [13, SIZE, SIZE], "global_batch": PAIRS * n,
            "parallelism": f"dp{n}" if n > 1 else "single",
            "cache": "inputs (2 x 218 MB fp32) and activations (> 1 GB per layer) exceed the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------ scene (configs[4])
def scene_record(args, dev, rank, world, side, steps=2, warm=1, e2e=True):
    """Full-scene tiled inference, 13-band side x side bi-date scene, 256 window, reference tiling (non-overlapping +
    last row / column / corner, utils/inference.py:134-236), sharded over ranks by ROW BANDS of tiles: each rank holds
    and uploads only the scene rows its tiles need and writes its band of the mask."""
    import torch
    import torch.distributed as dist
    from fabric_b200 import BiDateNet, ops
    from fabric_b200.scene import ScenePlan, predict_scene_band
    torch.manual_seed(0)
    model = BiDateNet(13, 2).to(dev).eval()
    plan = ScenePlan(side, side, SIZE, rank, world)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    d1 = torch.randn(13, plan.band_rows, side, device=dev, generator=g)       # this rank's rows only
    d2 = torch.randn(13, plan.band_rows, side, device=dev, generator=g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return predict_scene_band(model, d1, d2, plan, batch_size=PAIRS)
    for _ in range(warm):
        step()
    barrier()
    l0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        canvas = step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    launches = (ops.LAUNCHES - l0) // steps
    rec = {"value": plan.n_tiles / (ms / 1e3), "unit": "patch-pairs/s", "ms_per_scene": ms, "tiles": plan.n_tiles,
           "tiles_this_rank": plan.n_mine, "scene": [13, side, side], "steps": steps, "gpu_launches_per_scene": launches,
           "sharding": f"row bands of tiles over {world} rank(s); this rank holds {plan.band_rows} of {side} scene rows",
           "workload": f"full-scene tiled inference, 13-band {side}x{side} bi-date scene, 256 window, reference tiling"}
    if e2e:
        # end to end: this rank's band of the host scene (pinned) -> device, predict, its band of the mask back to the host
        h1 = torch.empty(d1.shape, dtype=torch.float32).pin_memory()
        h2 = torch.empty(d2.shape, dtype=torch.float32).pin_memory()
        h1.copy_(d1); h2.copy_(d2)
        hm = torch.empty((plan.out_rows, side), dtype=torch.uint8).pin_memory()
        barrier()
        t0 = time.perf_counter()
        a, b = h1.to(dev, non_blocking=True), h2.to(dev, non_blocking=True)
        canvas = predict_scene_band(model, a, b, plan, batch_size=PAIRS)
        hm.copy_(canvas, non_blocking=True)
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        rec["e2e"] = {"value": plan.n_tiles / float(tt.item()), "unit": "patch-pairs/s",
                      "h2d_bytes_per_step": 2 * h1.numel() * 4, "d2h_bytes_per_step": hm.numel(), "steps": 1,
                      "api": "fabric_b200.scene.predict_scene_band (pinned fp32 row band in, uint8 change-mask band out)"}
        del h1, h2, hm, a, b
    del d1, d2, canvas
    torch.cuda.empty_cache()
    return rec


# ------------------------------------------------------------------------------------------------ library bar
def library_baseline(dev, steps=6, warm=3):
    """The stock-library bar: the UNMODIFIED reference network (oracle/_ref; the
    restatement if absent) on the same GPU through torch: channels_last, bf16 autocast, cuDNN convolutions -- same batch
    (64 pairs 13x256x256), CUDA-event timed, eval forward and full training step (Tversky written with torch ops: the
    reference's `torch.eye(2)[labels]` does not accept CUDA labels on modern torch)."""
    import torch
    import torch.nn.functional as F
    from oracle import ref_loader
    if not ref_loader.available():
        return {"unavailable": "reference modules not present (oracle/_ref missing)"}
    RefNet = ref_loader.load()[0]
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = RefNet(13, 2).to(dev).to(memory_format=torch.channels_last)
    g = torch.Generator(device=dev).manual_seed(1)
    x1 = torch.randn(PAIRS, 13, SIZE, SIZE, device=dev, generator=g).to(memory_format=torch.channels_last)
    x2 = torch.randn(PAIRS, 13, SIZE, SIZE, device=dev, generator=g).to(memory_format=torch.channels_last)
    labels = (torch.rand(PAIRS, SIZE, SIZE, device=dev, generator=g) < 0.1).long()
    opt = torch.optim.SGD(model.parameters(), lr=LR)

    def tversky(logits, true, alpha=0.1, beta=0.9, eps=1e-7):      # utils/metrics.py:130-171 with 3-D labels (dims = (0, 2))
        p = F.softmax(logits.float(), dim=1)
        t = F.one_hot(true, 2).permute(0, 3, 1, 2).float()
        dims = (0, 2)
        inter = (p * t).sum(dims); fps = (p * (1 - t)).sum(dims); fns = ((1 - p) * t).sum(dims)
        return 1 - (inter / (inter + alpha * fps + beta * fns + eps)).mean()

    def infer():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(x1, x2)

    def train():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(x1, x2)
        loss = tversky(logits, labels)
        loss.backward()
        opt.step()
        return loss

    out = {"what": "unmodified reference BiDateNet on this GPU via torch " + torch.__version__ +
                   ": channels_last + bf16 autocast + cuDNN " + str(torch.backends.cudnn.version()) + " (benchmark mode)",
           "pairs": PAIRS}
    for name, fn, gf in (("infer", infer, FWD_GFLOP_PER_PAIR), ("train", train, TRAIN_GFLOP_PER_PAIR)):
        try:
            if name == "train":
                model.train()
            else:
                model.eval()
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": PAIRS / (ms / 1e3), "unit": "patch-pairs/s", "ms_per_step": ms,
                         "tflops": gf * PAIRS / ms, "steps": steps}
        except Exception as e:  # an OOM or a cuDNN failure of the LIBRARY arm must not take the bench line down
            out[name] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    del model, x1, x2, labels, opt
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ launch-bound shapes
def small_shape_record(dev, steps=30):
    """BASELINE configs[0] (2 pairs 13x32x32) and the reference's default geometry (32 pairs 13x90x90, metadata.json:32-33,40):
    the same training step, launch by launch (eager) and as ONE CUDA graph (fabric_b200.graph.GraphedTrainStep).  At these
    sizes the step is bound by the host side of ~180 launches, not by the GPU."""
    import torch
    from fabric_b200 import BiDateNet
    from fabric_b200.distributed import DataParallelStep
    from fabric_b200.graph import GraphedTrainStep
    from fabric_b200.metrics import TverskyLoss
    out = {}
    crit = TverskyLoss(alpha=0.1, beta=0.9)
    for name, b, s_ in (("config1_2x13x32x32", 2, 32), ("reference_default_32x13x90x90", 32, 90),
                        ("benchmark_64x13x256x256", PAIRS, SIZE)):
        if b == PAIRS:
            steps = 8
        try:
            torch.manual_seed(0)
            model = BiDateNet(13, 2).to(dev).train()
            dp = DataParallelStep(model)
            g = torch.Generator(device=dev).manual_seed(3)
            x1 = torch.randn(b, 13, s_, s_, device=dev, generator=g)
            x2 = torch.randn(b, 13, s_, s_, device=dev, generator=g)
            lab = (torch.rand(b, s_, s_, device=dev, generator=g) < 0.1).long()

            def eager():
                dp.zero_grad()
                loss = crit(model(x1, x2), lab)
                loss.backward()
                dp.sync_and_step(LR)
            graphed = GraphedTrainStep(model, crit, dp, LR, (x1, x2, lab))
            rec = {}
            for kind, fn in (("eager", eager), ("cuda_graph", lambda: graphed(x1, x2, lab))):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()           # wall clock on purpose: the host is the bottleneck being measured
                for _ in range(steps):
                    fn()
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / steps
                rec[kind] = {"ms_per_step": dt * 1e3, "value": b / dt, "unit": "patch-pairs/s"}
            rec["graph_speedup"] = rec["eager"]["ms_per_step"] / rec["cuda_graph"]["ms_per_step"]
            out[name] = rec
            dp.close()
            del graphed, model, dp
        except Exception as e:
            out[name] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ helpers
def summarise_profile(prof, steps):
    """(total conv ms, total algorithmic flops, per-layer dict, per-kind dict) from ops.CONV_PROFILE records."""
    conv_ms = sum(a.elapsed_time(b) for _, a, b, _ in prof)
    conv_flops = sum(f for _, _, _, f in prof)
    per_layer, kinds = {}, {}
    for tag, a, b, f in prof:
        ms = a.elapsed_time(b)
        d = per_layer.setdefault(tag, [0.0, 0.0, 0])
        d[0] += ms; d[1] += f; d[2] += 1
        kind = tag.split(" ")[0] if tag.split(" ")[0] in ("wgrad", "dgrad") else "fwd"
        k = kinds.setdefault(kind, [0.0, 0.0, 0])
        k[0] += ms; k[1] += f; k[2] += 1
    layers = {k: {"ms": v[0] / v[2], "tflops": v[1] / (v[0] / 1e3) / 1e12, "launches_per_step": v[2] // steps}
              for k, v in per_layer.items()}
    kinds = {k: {"ms_per_step": v[0] / steps, "tflops": v[1] / (v[0] / 1e3) / 1e12, "launches_per_step": v[2] // steps}
             for k, v in kinds.items()}
    return conv_ms, conv_flops, layers, kinds


# encoder double_conv blocks (the north_star's "fused encoder double_conv at batch 64"): forward tags of c1 / c2
ENCODER_BLOCKS = {"inc": ("13->64@256x256xG2", "64->64@256x256xG2"), "down1": ("64->128@128x128xG2", "128->128@128x128xG2"),
                  "down2": ("128->256@64x64xG2", "256->256@64x64xG2"), "down3": ("256->512@32x32xG2", "512->512@32x32xG2"),
                  "down4": ("512->512@16x16xG2", "512->512@16x16xG2")}


def encoder_blocks(layers, peak):
    out = {}
    for name, (t1, t2) in ENCODER_BLOCKS.items():
        if t1 not in layers or t2 not in layers:
            continue
        if t1 == t2:
            ms = layers[t1]["ms"] * 2
            fl = layers[t1]["tflops"] * layers[t1]["ms"] * 2
        else:
            ms = layers[t1]["ms"] + layers[t2]["ms"]
            fl = layers[t1]["tflops"] * layers[t1]["ms"] + layers[t2]["tflops"] * layers[t2]["ms"]
        out[name] = {"ms": ms, "tflops": fl / ms, "frac_of_burst_peak": fl / ms / peak}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="train", choices=["infer", "train", "scene"])
    ap.add_argument("--scene", type=int, default=10000, help="scene side for --workload scene / the scene sub-record")
    ap.add_argument("--impl", default="fabric_b200", choices=["fabric_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library", action="store_true", help="skip the torch / cuDNN library bar")
    ap.add_argument("--no-scene", action="store_true", help="skip the scene sub-record of the default line")
    ap.add_argument("--no-small", action="store_true", help="skip the launch-bound small-shape sub-record (eager vs CUDA graph)")
    ap.add_argument("--no-infer", action="store_true", help="skip the eval-forward sub-record of the train line")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunk", type=int, default=32, help="pairs per sub-batch of the host pipeline (infer e2e leg)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        if args.workload == "scene":
            args.workload = "infer"       # the reference's scene loop is its eval forward per batch of patches
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from fabric_b200 import BiDateNet, ops
    from fabric_b200.inference import HostPipeline, bind_to_gpu_numa_node

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus or (world == 1 and args.gpus == 1), f"launch with torchrun for --gpus {args.gpus}"
    numa = bind_to_gpu_numa_node(local)       # pinned host buffers and the copy threads live next to this rank's GPU
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "scene":
        rec = scene_record(args, dev, rank, world, args.scene, steps=max(1, min(args.steps, 5)), warm=max(1, min(args.warmup, 2)))
        if rank == 0:
            print(json.dumps({
                "metric": "patch-pairs/s (13x256x256)", "value": rec["value"], "unit": "patch-pairs/s", "n_gpus": world,
                "steps": rec["steps"], "warmup": max(1, min(args.warmup, 2)), "ms_per_step": rec["ms_per_scene"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": rec["workload"], "tiles": rec["tiles"], "scene": rec["scene"], "sharding": rec["sharding"],
                           "parallelism": f"tile-row-bands x{world}"},
                "e2e": rec.get("e2e"), "gpu_launches": rec["gpu_launches_per_scene"] * rec["steps"], "roofline": None,
                "cpu_baseline": None, "clocks": None}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    torch.manual_seed(0)
    train = args.workload == "train"
    model = BiDateNet(13, 2).to(dev)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    x1 = torch.randn(PAIRS, 13, SIZE, SIZE, device=dev, generator=g)
    x2 = torch.randn(PAIRS, 13, SIZE, SIZE, device=dev, generator=g)
    labels = (torch.rand(PAIRS, SIZE, SIZE, device=dev, generator=g) < 0.1).long()

    def infer_step(a=x1, b=x2):
        with torch.no_grad():
            return model(a, b)

    if train:
        from fabric_b200.distributed import DataParallelStep
        from fabric_b200.metrics import TverskyLoss
        model.train()
        criterion = TverskyLoss(alpha=0.1, beta=0.9)                  # metadata.json:42-44
        # owns the flat gradient bucket (+ fused SGD); FABRIC_B200_NO_OVERLAP=1: one all-reduce after backward (A/B switch)
        dp = DataParallelStep(model, overlap=os.environ.get("FABRIC_B200_NO_OVERLAP", "0") != "1")
        dp.broadcast_parameters(0)
        if os.environ.get("FABRIC_B200_HIPRI", "1") != "0":       # (A/B switch, tools/ab.sh)
            dp.use_compute_stream()     # the loop runs on a high-priority stream: main chain ahead of the side-stream wgrads

        def step(a=x1, b=x2, lab=labels):                            # train.py:88-95
            dp.zero_grad()
            loss = criterion(model(a, b), lab)
            loss.backward()
            dp.sync_and_step(LR)       # ONE NCCL all-reduce (skipped for N = 1) + ONE fused SGD kernel (train.py:95)
            return loss
    else:
        model.eval()
        step = infer_step

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        l0 = ops.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ops.LAUNCHES - l0

    def profiled(fn, steps):
        """the SAME steps again with a CUDA-event pair around every conv / wgrad launch on the launching stream (kept out
        of the pass that yields `value`: ~40-100 event records per step perturb back-to-back launches)"""
        from fabric_b200 import autograd
        side, autograd.WGRAD_SIDE_STREAM = autograd.WGRAD_SIDE_STREAM, False   # one stream: clean per-kernel durations
        ops.CONV_PROFILE = []
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(steps):
            fn()
        p1.record()
        barrier()
        prof, ops.CONV_PROFILE = ops.CONV_PROFILE, None
        autograd.WGRAD_SIDE_STREAM = side
        return prof, p0.elapsed_time(p1)

    sampler = ClockSampler(local)
    ops.CONV_PROFILE = None
    for _ in range(args.warmup):
        step()
    barrier()
    if os.environ.get("FABRIC_B200_PROFILE_STEP"):
        # profiling hook (never a bench number): exactly ONE step between cudaProfilerStart / Stop, for
        # `ncu --profile-from-start off --set full` (tools/prof_round2.sh); "infer" profiles one eval forward instead
        what = os.environ["FABRIC_B200_PROFILE_STEP"]
        if what == "infer" and train:
            model.eval()
            for _ in range(2):
                infer_step()
        barrier()
        torch.cuda.cudart().cudaProfilerStart()
        infer_step() if what == "infer" else step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        print(json.dumps({"profiled": what, "note": "one step under the profiler; not a benchmark"}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    if rank == 0:
        sampler.start()
    total_ms, launches = timed(step, args.steps, 0)
    prof, prof_total_ms = profiled(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = total_ms / args.steps
    value = PAIRS * world * args.steps / (total_ms / 1e3)

    # ---- roofline of the dominant kernels (tcgen05 convolutions: forward, data-gradient and weight-gradient launches)
    conv_ms, conv_flops, layers, kinds = summarise_profile(prof, args.steps)
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp) and train:
        tj = json.load(open(tp))
        traffic = tj.get("dram_bytes_per_launch")
        traffic_note = {"dram_bytes_per_step": tj.get("dram_bytes_per_step"), "launches_per_step": tj.get("launches_per_step"),
                        "algorithmic_bytes_per_step": tj.get("algorithmic_bytes_per_step"), "source": tj.get("source")}
    conv_share = conv_ms / prof_total_ms
    roofline = {"bound": "tensor",
                "kernel": "conv3x3_umma_kernel (forward" + (" + data-gradient launches) + wgrad_umma_kernel" if train else " launches)") +
                          ", all launches of a step",
                "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                "frac_of_sustained_peak": achieved / peaks["tflops_sustained"],
                "peak_source": peaks["source"] + ", burst bf16 figure (BASELINE.md section 3: primary denominator); sustained = "
                               f"{peaks['tflops_sustained']:.1f}",
                "conv_share_of_step": conv_share, "non_conv_share_of_step": 1.0 - conv_share,
                "non_conv": "BatchNorm statistics / apply / backward passes, loss, 1x1 head, upsample + adjoint, packing, SGD"
                            if train else "input packing, decoder upsample",
                "by_kind": kinds, "traffic": traffic, "traffic_detail": traffic_note,
                "whole_step_tflops": (TRAIN_GFLOP_PER_PAIR if train else FWD_GFLOP_PER_PAIR) * PAIRS / ms_step,
                "whole_step_frac": (TRAIN_GFLOP_PER_PAIR if train else FWD_GFLOP_PER_PAIR) * PAIRS / ms_step / peaks["tflops"],
                "timing": "CUDA events around each conv / wgrad launch in a second pass of the same K steps, right after the "
                          f"timed pass (that pass: {prof_total_ms / args.steps:.3f} ms/step with the event records)",
                "algorithmic_gflop_per_step": conv_flops / args.steps / 1e9}
    if not train:
        roofline["encoder_double_conv"] = encoder_blocks(layers, peaks["tflops"])

    # ---- eval forward (BASELINE configs[1]) as a sub-record of the training line
    infer = None
    if train and not args.no_infer:
        model.eval()
        i_ms, i_launch = timed(infer_step, args.steps, 3)
        i_prof, i_prof_ms = profiled(infer_step, args.steps)
        i_conv_ms, i_conv_fl, i_layers, _ = summarise_profile(i_prof, args.steps)
        i_ach = i_conv_fl / (i_conv_ms / 1e3) / 1e12
        infer = {"value": PAIRS * world * args.steps / (i_ms / 1e3), "unit": "patch-pairs/s", "ms_per_step": i_ms / args.steps,
                 "steps": args.steps, "gpu_launches": i_launch, "workload": config_dict("infer", world)["workload"],
                 "roofline": {"bound": "tensor", "achieved": i_ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                              "frac": i_ach / peaks["tflops"], "conv_share_of_step": i_conv_ms / i_prof_ms,
                              "whole_step_frac": FWD_GFLOP_PER_PAIR * PAIRS / (i_ms / args.steps) / peaks["tflops"],
                              "encoder_double_conv": encoder_blocks(i_layers, peaks["tflops"])},
                 "layers": i_layers}
        model.train()

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region ----------------
    hp1 = torch.randn(PAIRS, 13, SIZE, SIZE).pin_memory()
    hp2 = torch.randn(PAIRS, 13, SIZE, SIZE).pin_memory()

    def timed_e2e(fn, n):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return PAIRS * world * n / float(tt.item()), float(tt.item())

    mean = torch.tensor([1617.57, 1422.37, 1359.37, 1414.68, 1557.94, 1986.22, 2210.50, 2118.56, 2344.79, 711.84, 15.75,
                         2133.90, 1584.27])        # metadata.json:4-29
    std = torch.tensor([319.12, 456.25, 590.13, 849.37, 811.31, 813.55, 891.85, 901.61, 954.77, 370.95, 9.23, 1116.59,
                        985.12])
    e2e_extra = {}
    if train:
        # reference train.py:83-95 per batch: host batch -> device, step, loss back to the host (.item(), helpers.py:83);
        # the copy of batch i+1 overlaps step i (fabric_b200.inference.BatchFeeder)
        from fabric_b200.inference import BatchFeeder
        hlab = (torch.rand(PAIRS, SIZE, SIZE) < 0.1).long().pin_memory()
        feeder = BatchFeeder(dev)
        state = {"batch": (hp1, hp2, hlab)}
        feeder.prefetch(state["batch"])

        def e2e_step():
            a, b, lab = feeder.next()
            feeder.prefetch(state["batch"])           # next step's inputs: H2D on the side stream while this step runs
            loss = step(a, b, lab)
            feeder.release()
            return loss.item()
        n_e2e = max(args.e2e_steps, 8)
        v, wall = timed_e2e(e2e_step, n_e2e)
        h2d = 2 * hp1.numel() * 4 + hlab.numel() * 8
        e2e = {"value": v, "unit": "patch-pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": n_e2e,
               "h2d_gbs_per_rank": h2d * n_e2e / wall / 1e9,
               "api": "BatchFeeder.next(); model(x1,x2); TverskyLoss(logits, labels); loss.backward(); "
                      "DataParallelStep.sync_and_step(); loss.item()  [pinned fp32 NCHW patches + int64 labels in]"}
        # the same loop fed with the rasters as stored (uint16 digital numbers, uint8 labels): the loader's z-score
        # (utils/dataloaders.py:94-99) runs inside the pack kernel
        model.set_input_normalisation(mean, std)
        raw = [(hp * std[None, :, None, None] + mean[None, :, None, None]).round().clamp(0, 65535).to(torch.int32)
               .to(torch.uint16).pin_memory() for hp in (hp1, hp2)]
        state["batch"] = (raw[0], raw[1], hlab.to(torch.uint8).pin_memory())
        feeder.next(); feeder.release()
        feeder.prefetch(state["batch"])
        v, wall = timed_e2e(e2e_step, n_e2e)
        h2d = 2 * raw[0].numel() * 2 + hlab.numel()
        e2e_extra["e2e_raw_uint16"] = {"value": v, "unit": "patch-pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                                       "steps": n_e2e, "h2d_gbs_per_rank": h2d * n_e2e / wall / 1e9,
                                       "api": "same loop, pinned uint16 rasters + uint8 labels in (z-score fused into the pack kernel)"}
        feeder.next(); feeder.release()
        model.eval()
    # eval forward from host buffers: HostPipeline (fp32 in -> fp32 logits out, and uint16 in -> uint8 change mask out,
    # which is what the reference's prediction loop consumes: train.py:187-201 takes the argmax)
    if not args.no_infer or not train:
        hout = torch.empty(PAIRS, 2, SIZE, SIZE).pin_memory()
        pipe = HostPipeline(model, chunk=args.e2e_chunk, n_channels=13, size=SIZE, return_logits=True)
        res = {}

        def e2e_infer():
            res["b"] = pipe.run(hp1, hp2, hout)
        v, wall = timed_e2e(e2e_infer, args.e2e_steps)
        rec = {"value": v, "unit": "patch-pairs/s", "h2d_bytes_per_step": res["b"][0], "d2h_bytes_per_step": res["b"][1],
               "steps": args.e2e_steps, "h2d_gbs_per_rank": res["b"][0] * args.e2e_steps / wall / 1e9,
               "api": f"fabric_b200.inference.HostPipeline.run (pinned fp32 NCHW in, fp32 logits out, {args.e2e_chunk}-pair sub-batches)"}
        # what the host can deliver to this GPU when nothing else runs: the same pinned buffers copied in the same 32-pair
        # sub-batches, no compute (with N ranks on one host all N copy at once: the aggregate is the platform's H2D ceiling)
        dst = torch.empty(args.e2e_chunk, 13, SIZE, SIZE, device=dev)

        def h2d_only():
            for lo in range(0, PAIRS, args.e2e_chunk):
                dst.copy_(hp1[lo:lo + args.e2e_chunk], non_blocking=True)
                dst.copy_(hp2[lo:lo + args.e2e_chunk], non_blocking=True)
        _, wall = timed_e2e(h2d_only, args.e2e_steps)
        rec["h2d_only_gbs_per_rank"] = res["b"][0] * args.e2e_steps / wall / 1e9
        rec["h2d_only_gbs_all_ranks"] = rec["h2d_only_gbs_per_rank"] * world      # (wall = max over ranks)
        del dst
        if train:
            infer["e2e"] = rec
        else:
            e2e = rec
        if not train:
            model.set_input_normalisation(mean, std)
            raw = [(hp * std[None, :, None, None] + mean[None, :, None, None]).round().clamp(0, 65535).to(torch.int32)
                   .to(torch.uint16).pin_memory() for hp in (hp1, hp2)]
        hmask = torch.empty(PAIRS, SIZE, SIZE, dtype=torch.uint8).pin_memory()
        pipe_m = HostPipeline(model, chunk=args.e2e_chunk, n_channels=13, size=SIZE, return_logits=False)

        def e2e_mask():
            res["m"] = pipe_m.run(raw[0], raw[1], hmask)
        v, wall = timed_e2e(e2e_mask, args.e2e_steps)
        rec = {"value": v, "unit": "patch-pairs/s", "h2d_bytes_per_step": res["m"][0], "d2h_bytes_per_step": res["m"][1],
               "steps": args.e2e_steps, "h2d_gbs_per_rank": res["m"][0] * args.e2e_steps / wall / 1e9,
               "api": "HostPipeline.run (pinned uint16 rasters in, uint8 argmax change mask out -- what train.py:187-201 consumes)"}
        if train:
            infer["e2e_raw_uint16_mask"] = rec
        else:
            e2e_extra["e2e_raw_uint16_mask"] = rec
        del pipe, pipe_m
    if train:
        model.train()

    # ---- sub-records measured outside the headline: scene (configs[4]), library bar, CPU reference ------------------
    scene = None
    if not args.no_scene:
        del x1, x2
        torch.cuda.empty_cache()
        try:
            scene = scene_record(args, dev, rank, world, args.scene, steps=2, warm=1, e2e=False)
        except Exception as e:
            scene = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    small = None
    if world == 1 and train and not args.no_small:
        small = small_shape_record(dev)
    lib = None
    if rank == 0 and not args.no_library:
        lib = library_baseline(dev)
        mine = {"train": value / world if train else None,
                "infer": (infer["value"] / world if infer else None) if train else value / world}
        for k, v in mine.items():
            if v and isinstance(lib.get(k), dict) and "value" in lib[k]:
                lib[k]["fabric_b200_speedup"] = v / lib[k]["value"]       # per GPU, same batch, same box
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference_throughput(args.workload)
        out = {
            "metric": "patch-pairs/s (13x256x256)", "value": value, "unit": "patch-pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_dict(args.workload, world),
            "tflops_per_gpu": (TRAIN_GFLOP_PER_PAIR if train else FWD_GFLOP_PER_PAIR) * PAIRS / ms_step,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e, **e2e_extra,
            "gpu_launches": launches, "numa_node": numa, "layers": layers, "infer": infer, "scene": scene,
            "launch_bound_shapes": small, "library_baseline": lib,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

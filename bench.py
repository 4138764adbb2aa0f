"""bench.py -- CHAP hot-path throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

Headline (`value`, `e2e`, `roofline`, `cpu_baseline`): BASELINE.json configs[1] -- one full CHAP training iteration of the
2D DualDecoder U-Net on synthetic ACDC-shaped input (batch 24 of which 12 labelled, 1x256x256, 4 classes; --adv_noise, 'kl'
consistency; channel+spatial hierarchical perturbation), weak scaling with one process per GPU.  The same JSON line also
carries, under `workloads`, the other two configurations BASELINE.json's metric names:
    vnet3d    configs[2]: DualDecoder3d CHAP iteration, batch 4 (2 labelled), 1x112x112x80, 2 classes
    sw_infer  configs[3]: VNet sliding-window inference, 192x192x88 volumes = 108 windows of 112x112x80, stride 18/18/4,
              cases sharded round-robin over the ranks (no collective)
each with its own device-resident value, end-to-end value, roofline of its dominant kernel and CPU baseline; `modes` gives
the 2D iteration rate in the three convolution arithmetics (TF32 default / 3xTF32 on <= 32-channel layers / 3xTF32
everywhere); `gpu_eager_baseline` is the oracle restatement of the same 2D iteration executed by torch-eager + cuDNN on the
same GPU (the reference's own execution model).  `--workload X` makes X the headline and skips the extras.
One JSON line is printed by rank 0 (contract in the task description).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("CHAP_TIMING_DETAIL", "1")          # per-shape kernel timer names (read by libchap_b200 at first use)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    "unet2d": dict(kind="train", dims=2, batch=24, labeled=12, shape=(256, 256), classes=4, unit="it/s", metric="CHAP train iters/sec",
                   name="UNet(DualDecoder) 2D CHAP full training step, b24 (12 labelled), 1x256x256, 4 classes"),
    "vnet3d": dict(kind="train", dims=3, batch=4, labeled=2, shape=(112, 112, 80), classes=2, unit="it/s", metric="CHAP train iters/sec",
                   name="VNet(DualDecoder3d) 3D CHAP full training step, b4 (2 labelled), 1x112x112x80, 2 classes"),
    "sw_infer": dict(kind="infer", dims=3, shape=(192, 192, 88), patch=(112, 112, 80), stride_xy=18, stride_z=4, classes=2,
                     unit="cases/s", metric="VNet sliding-window inference cases/sec", batch_windows=4,
                     name="VNet 3D sliding-window inference, 192x192x88 volume = 108 windows of 112x112x80, stride 18/18/4, 2 classes"),
}
FLAGS = "--adv_noise --adv_losstype kl (per-pixel mean) --decoder_type mcnet --noise_mag 10 epi 6 topk 0.1"
CPU_BUDGET_S = 25.0           # bounded CPU sample per workload inside the default GPU run
REF_BUDGET_S = 240.0          # --impl reference: whole run within a few minutes


def synth_batch(w, seed, batch=None):
    """Synthetic ACDC/LA-shaped batch: U[0,1) (2D) / N(0,1) (3D) intensities, blob labels."""
    n = batch or w["batch"]
    g = torch.Generator().manual_seed(seed)
    shape = (n, 1) + tuple(w["shape"])
    vol = torch.rand(shape, generator=g) if w["dims"] == 2 else torch.randn(shape, generator=g)
    lab = torch.zeros((n,) + tuple(w["shape"]), dtype=torch.int64)
    grids = torch.meshgrid(*[torch.arange(s) for s in w["shape"]], indexing="ij")
    for i in range(n):
        for c in range(1, w["classes"]):
            ctr = [int(torch.randint(s // 4, 3 * s // 4, (1,), generator=g)) for s in w["shape"]]
            rad = min(w["shape"]) // (6 + 2 * c)
            lab[i][sum((gr - ct) ** 2 for gr, ct in zip(grids, ctr)) < rad * rad] = c
    return vol, lab


def synth_volume(w, seed):
    return np.random.RandomState(seed).randn(*w["shape"]).astype(np.float32)


def build_model(w, device):
    from chap_b200 import networks
    torch.manual_seed(1337)
    if w.get("kind") == "infer":
        return networks.net_factory_3d("vnet", in_chns=1, class_num=w["classes"], mode="test", device=device)
    if w["dims"] == 2:
        return networks.net_factory("dualdecoder", in_chns=1, class_num=w["classes"], device=device,
                                    args={"decoder_type": "mcnet"})
    return networks.net_factory_3d("dualdecoder", in_chns=1, class_num=w["classes"], mode="train", device=device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
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


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor=p["bf16_tflops_sustained"], source="measured (MEASURED_PEAKS.json; sustained bf16, HBM copy)")
    return dict(hbm=6650.0, tensor=1400.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic_table():
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(path):
            return json.load(open(path)).get("dram_bytes_per_launch", {})
    return {}


# ------------------------------------------------------------------------------------------- CPU arm (the reference's path)
def cpu_chap_iterations(w, threads, steps, warmup, budget_s):
    """The reference's CPU path: oracle port of one CHAP iteration (oracle/train_step.py) at the REAL configuration (full batch,
    full shape; no scaling), `threads` host threads.  Returns (list of seconds per timed iteration, loss of the last one).
    `steps` is cut down only if the projected time exceeds `budget_s` (always >= 2 timed iterations when steps >= 2)."""
    from oracle import chap_losses as L
    from oracle import nets
    from oracle import train_step as ost
    from chap_b200 import networks
    torch.set_num_threads(threads)
    torch.manual_seed(1337)
    if w["dims"] == 2:
        model = networks.DualDecoder(1, w["classes"], {"decoder_type": "mcnet"})
    else:
        model = networks.DualDecoder3d(1, w["classes"], normalization="batchnorm", has_dropout=True)
    sd = nets.clone_state_dict(model.state_dict(), requires_grad=True)
    om = ost.OracleModel(sd, dims=w["dims"], has_dropout=(w["dims"] == 3), drop="torch")
    bufs = [None] * len(om.params())
    vat = L.VAT(10.0, 6.0, w["classes"])
    times, it, loss = [], 0, float("nan")
    t_start = time.perf_counter()
    while len(times) < steps:
        vol, lab = synth_batch(w, it)
        offs = L.draw_mask_offsets(w["shape"], np.random.RandomState(it))
        t0 = time.perf_counter()
        aux = ost.chap_train_step(om, bufs, vol, lab, w["labeled"], w["classes"], offs, it, vat=vat, topk=0.1)
        dt = time.perf_counter() - t0
        loss = float(aux["loss"])
        if it >= warmup:
            times.append(dt)
        it += 1
        done = len(times)
        if done >= min(2, steps) and (time.perf_counter() - t_start) + dt > budget_s:
            break
    return times, loss


def cpu_sw_case(w, threads, budget_s):
    """Reference sliding-window inference on the host: numpy restatement of code/test_3D_util.py:14-79 around the oracle VNet.
    A full case is 108 VNet forwards; within `budget_s` only the first windows are run and the case time is windows-scaled
    (BASELINE.md section 4.3 allows exactly this; the sample is stated in the JSON)."""
    from oracle import nets
    from oracle import sliding_window as osw
    from chap_b200 import networks
    torch.set_num_threads(threads)
    torch.manual_seed(1337)
    model = networks.VNet(1, w["classes"], normalization="batchnorm", has_dropout=False)
    sd = nets.clone_state_dict(model.state_dict())
    vol = synth_volume(w, 0)
    n_win = int(np.prod([len(osw.window_starts(d, p, s)) for d, p, s in
                         zip(w["shape"], w["patch"], (w["stride_xy"], w["stride_xy"], w["stride_z"]))]))
    done, t0 = [0], time.perf_counter()

    class _Stop(Exception):
        pass

    def net_fn(patch):
        if done[0] >= 2 and (time.perf_counter() - t0) * (done[0] + 1) / done[0] > budget_s:
            raise _Stop()
        with torch.no_grad():
            y = nets.vnet_forward(sd, torch.from_numpy(patch), False, False).numpy()
        done[0] += 1
        return y
    try:
        osw.test_single_case(net_fn, vol, w["stride_xy"], w["stride_z"], w["patch"], num_classes=w["classes"])
    except _Stop:
        pass
    sec = (time.perf_counter() - t0) * n_win / max(done[0], 1)
    return sec, done[0], n_win


def cpu_baseline_for(w, threads, budget_s):
    if w["kind"] == "infer":
        sec, done, n_win = cpu_sw_case(w, threads, budget_s)
        sample = ("oracle sliding-window loop (numpy accumulate + oracle VNet forward on torch-CPU), first %d of the %d windows of one "
                  "%s case, time scaled by windows" % (done, n_win, "x".join(map(str, w["shape"]))))
        return {"value": 1.0 / sec, "unit": w["unit"], "cores": threads, "kind": "port", "sample": sample}
    times, _ = cpu_chap_iterations(w, threads, 2, 1, budget_s)
    sec = float(np.mean(times))
    return {"value": 1.0 / sec, "unit": w["unit"], "cores": threads, "kind": "port",
            "sample": "%d timed CHAP iteration(s) (after 1 warm-up) of the oracle port at the real configuration (batch %d of %s, %d labelled); no scaling"
                      % (len(times), w["batch"], "x".join(map(str, w["shape"])), w["labeled"])}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    if w["kind"] == "infer":
        sec, done, n_win = cpu_sw_case(w, threads, REF_BUDGET_S / 4)
        value, steps, warm, loss = 1.0 / sec, 1, 0, None
        sample = "first %d of %d windows of one case, scaled by windows" % (done, n_win)
    else:
        times, loss = cpu_chap_iterations(w, threads, args.steps, args.warmup, REF_BUDGET_S)
        sec, steps, warm = float(np.mean(times)), len(times), args.warmup
        value = 1.0 / sec
        sample = ("%d timed CHAP iterations of the oracle port (reference nets restated on torch-CPU + frozen losses) at the real configuration "
                  "(batch %d of %s, %d labelled), no scaling%s" % (steps, w["batch"], "x".join(map(str, w["shape"])), w["labeled"],
                  "" if steps == args.steps else "; fewer than the %d requested steps to stay inside %.0f s" % (args.steps, REF_BUDGET_S)))
    line = {"impl": "reference", "metric": w["metric"], "value": value, "unit": w["unit"], "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "where": "host CPU", "flags": FLAGS, "loss_last": loss},
            "cpu_baseline": {"value": value, "unit": w["unit"], "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": w["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- roofline helper
def roofline_from_timers(fam, step_ms, peaks, prefer=None):
    """fam: {timer name: dict(ms, launches, flops, bytes)} per step.  Dominant kernel = costliest layer shape of the costliest family
    (or of family `prefer`); algorithmic bytes / flops per launch come from the launchers (formulas in DESIGN.md section 3)."""
    if not fam:
        return None
    families = {}
    for k, v in fam.items():
        f = families.setdefault(k.split(":")[0], dict(ms=0.0, launches=0, flops=0.0, bytes=0.0))
        for key in f:
            f[key] += v[key]
    top_family = prefer if prefer in families else max(families.items(), key=lambda kv: kv[1]["ms"])[0]
    name, f = max(((k, v) for k, v in fam.items() if k.split(":")[0] == top_family), key=lambda kv: kv[1]["ms"])
    per_launch_s = f["ms"] / 1e3 / max(f["launches"], 1)
    tf32_peak = peaks["tensor"] / 2.0                                     # kind::tf32 runs at half the dense bf16 rate
    intensity = f["flops"] / f["bytes"] if f["bytes"] > 0 else float("inf")
    tensor_bound = f["flops"] > 0 and intensity > tf32_peak * 1e12 / (peaks["hbm"] * 1e9)
    if tensor_bound:
        achieved = f["flops"] / f["launches"] / per_launch_s / 1e12
        roof = {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak}
    else:
        achieved = f["bytes"] / f["launches"] / per_launch_s / 1e9
        roof = {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / peaks["hbm"]}
    roof["traffic"] = ncu_traffic_table().get(name)                       # ncu --set full: dram read + write bytes, per launch
    roof["algorithmic_per_launch"] = {"flops": f["flops"] / f["launches"], "bytes": f["bytes"] / f["launches"], "flop_per_byte": intensity}
    roof["us_per_launch"] = per_launch_s * 1e6
    roof["peak_source"] = peaks["source"] + ("; tensor peak = measured dense bf16 / 2 (tf32)" if tensor_bound else "")
    roof["family"] = top_family
    roof["family_share_of_step"] = families[top_family]["ms"] / step_ms
    roof["launches_per_step"] = f["launches"]
    roof["share_of_step"] = f["ms"] / step_ms
    roof["families_ms_per_step"] = {k: round(v["ms"], 4) for k, v in sorted(families.items(), key=lambda kv: -kv[1]["ms"])}
    roof["families_gbps"] = {k: round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1) for k, v in families.items()
                             if v["bytes"] > 0 and v["flops"] == 0 and v["ms"] > 0}
    roof["top_kernels_us"] = {k: [v["launches"], round(1e3 * v["ms"] / v["launches"], 1)]
                              for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])[:12]}
    roof["note"] = ("per-kernel CUDA-event timing on the launching stream, taken in an eager re-issue of the same step right after the timed "
                    "region; the dominant kernel is one layer shape of one kernel family; bound chosen by its algorithmic FLOP/byte against "
                    "the measured machine balance")
    return roof


class Dist:
    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_ms(self, ms, dev):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if self.world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t)


# ------------------------------------------------------------------------------------------- training workloads (GPU)
def time_training(args, w, d, dev, steps, warmup, use_graph, kernel_timing, clocks_on):
    """One CHAP training workload: device-resident rate, end-to-end rate, loss health, kernel families."""
    import torch.distributed as dist
    from chap_b200 import _lib
    from chap_b200.train_step import ChapTrainer
    model = build_model(w, dev)

    def allreduce(flat_g):                       # data-parallel: one flat-bucket NCCL all-reduce per iteration
        dist.all_reduce(flat_g)
    trainer = ChapTrainer(model, n_classes=w["classes"], labeled_bs=w["labeled"], max_iterations=30000,
                          grad_hook=allreduce if d.world > 1 else None, grad_scale=1.0 / d.world,
                          use_graph=use_graph, graph_warmup=2, overlap_allreduce=not args.no_overlap)
    n_in = max(2, min(4, steps))
    host = [synth_batch(w, 1000 * d.rank + i) for i in range(n_in)]
    host = [(v.pin_memory(), l.pin_memory()) for v, l in host]
    resident = [(v.to(dev), l.to(dev)) for v, l in host]
    losses = []

    # ---- device-resident timing ("value"): the whole iteration is ONE CUDA-graph replay per step
    warm = max(warmup, 3)                                          # >= 3 (graph mode: 2 eager iterations + the capture step)
    for i in range(warm):
        out = trainer.step(*resident[i % n_in])
        losses.append(out["loss"].clone())
    d.barrier()
    clocks = ClockSampler(d.local) if (clocks_on and d.rank == 0) else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = trainer.step(*resident[i % n_in])
        losses.append(out["loss"].clone())                         # device-side copy of the step's loss (no sync): health record
    e1.record()
    d.barrier()
    ms = d.max_ms(e0.elapsed_time(e1), dev)
    clk = clocks.stop() if clocks else None

    # ---- end-to-end timing ("e2e"): host (pinned) inputs -> H2D -> step -> loss read back, every step
    d.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from chap_b200.parallel import DevicePrefetcher
    t0.record()
    # every step's inputs come from pinned host memory inside the timed region; the copy of batch i + 1 runs on a side stream while
    # iteration i computes (chap_b200.parallel.DevicePrefetcher), the loss of every step is read back
    for v, l in DevicePrefetcher((host[i % n_in] for i in range(steps)), dev):
        out = trainer.step(v, l)
        losses.append(out["loss"].clone())
        float(out["loss"])                                # D2H of the step's result
    t1.record()
    d.barrier()
    ms_e2e = d.max_ms(t0.elapsed_time(t1), dev)

    # ---- kernel-family pass (roofline, launch count): the SAME iteration issued eagerly so that CUDA events can
    # bracket individual launches on the launching stream (events cannot be read back from inside a graph replay)
    fam, launches_per_step, ms_eager = {}, 0, None
    if kernel_timing:
        trainer.use_graph = False
        trainer.step(*resident[0])
        d.barrier()
        _lib.reset_launch_count()
        _lib.timing_enable(d.rank == 0)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 2
        r0.record()
        for i in range(reps):
            trainer.step(*resident[i % n_in])
        r1.record()
        d.barrier()
        ms_eager = r0.elapsed_time(r1) / reps
        launches_per_step = _lib.launch_count() // reps
        fam = _lib.timing_report() if d.rank == 0 else {}
        fam = {k: dict(v, ms=v["ms"] / reps, launches=v["launches"] // reps, flops=v["flops"] / reps, bytes=v["bytes"] / reps)
               for k, v in fam.items()}
        _lib.timing_enable(False)
        trainer.use_graph = use_graph
    lv = torch.stack(losses).float().cpu().numpy()
    trainer.close()                      # drop the captured graph BEFORE any NCCL teardown (a live graph holding the all-reduce hangs it)
    del trainer, model
    vox = float(np.prod(w["shape"]))
    res = {"ms": ms / steps, "ms_e2e": ms_e2e / steps, "clocks": clk, "fam": fam, "launches_per_step": int(launches_per_step),
           "eager_ms": ms_eager, "loss_first": float(lv[0]), "loss_last": float(lv[-1]), "loss_max": float(np.nanmax(lv)),
           "loss_finite": bool(np.isfinite(lv).all()),
           # per step: volume f32 + label i64 + the copy-paste mask i64 [*spatial] (built on the device from 2-3 host integers);
           # lr / consistency weight are computed on the device
           "h2d": int(w["batch"] * vox * (4 + 8)), "d2h": 4}
    return res


def training_block(args, w, d, dev, res, peaks, cpu):
    step_ms = res["ms"]
    roof = roofline_from_timers(res["fam"], step_ms, peaks)
    if roof is not None:
        roof["eager_ms_per_step"] = res["eager_ms"]
    return {"metric": w["metric"], "value": d.world * 1e3 / res["ms"], "unit": w["unit"], "ms_per_step": res["ms"],
            "config": {"workload": w["name"], "per_gpu_batch": w["batch"], "parallelism": "dp%d" % d.world, "flags": FLAGS,
                       "l2": "activations touched per step (>5 GB) far exceed the 126 MB L2; no explicit flush"},
            "loss_first": res["loss_first"], "loss_last": res["loss_last"], "loss_max": res["loss_max"], "loss_finite": res["loss_finite"],
            "e2e": {"value": d.world * 1e3 / res["ms_e2e"], "unit": w["unit"], "h2d_bytes_per_step": res["h2d"],
                    "d2h_bytes_per_step": res["d2h"], "ms_per_step": res["ms_e2e"]},
            "gpu_launches_per_step": res["launches_per_step"], "roofline": roof, "cpu_baseline": cpu}


# ------------------------------------------------------------------------------------------- sliding-window workload (GPU)
def time_sw_infer(args, w, d, dev, cases_per_rank, kernel_timing):
    from chap_b200 import _lib
    from chap_b200.test_3D_util import sliding_window_device, test_single_case
    net = build_model(w, dev)
    # cases are sharded round-robin over the ranks (replicas only, no collective): global case j belongs to rank j % world
    vols = [synth_volume(w, d.rank + d.world * i) for i in range(min(cases_per_rank, 2))]
    vols_dev = [torch.from_numpy(v).to(dev) for v in vols]
    kw = dict(stride_xy=w["stride_xy"], stride_z=w["stride_z"], patch_size=w["patch"], num_classes=w["classes"], batch_windows=w["batch_windows"])
    for _ in range(2):                                                               # warm-up (tensor-map / weight-pack caches)
        sliding_window_device(net, vols_dev[0], **kw)
    d.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(cases_per_rank):
        lab = sliding_window_device(net, vols_dev[i % len(vols_dev)], **kw)
    e1.record()
    d.barrier()
    ms = d.max_ms(e0.elapsed_time(e1), dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(cases_per_rank):                                                  # host volume in -> host label map out
        lab = test_single_case(net, vols[i % len(vols)], w["stride_xy"], w["stride_z"], w["patch"], num_classes=w["classes"],
                               batch_windows=w["batch_windows"])
    t1.record()
    d.barrier()
    ms_e2e = d.max_ms(t0.elapsed_time(t1), dev)
    fam, launches = {}, 0
    if kernel_timing:
        _lib.reset_launch_count()
        _lib.timing_enable(d.rank == 0)
        sliding_window_device(net, vols_dev[0], **kw)
        d.barrier()
        launches = _lib.launch_count()
        fam = _lib.timing_report() if d.rank == 0 else {}
        _lib.timing_enable(False)
    vox = float(np.prod(w["shape"]))
    return {"ms": ms / cases_per_rank, "ms_e2e": ms_e2e / cases_per_rank, "fam": fam, "launches_per_step": int(launches),
            "h2d": int(vox * 4), "d2h": int(vox * 1), "cases_per_rank": cases_per_rank, "label_sum": int(np.asarray(lab).sum())}


def sw_block(args, w, d, res, peaks, cpu):
    kernel_ms = sum(v["ms"] for v in res["fam"].values()) if res["fam"] else None
    roof = roofline_from_timers(res["fam"], kernel_ms or res["ms"], peaks)
    agg = roofline_from_timers(res["fam"], kernel_ms or res["ms"], peaks, prefer="sw_aggregate")
    return {"metric": w["metric"], "value": d.world * 1e3 / res["ms"], "unit": w["unit"], "ms_per_step": res["ms"],
            "config": {"workload": w["name"], "sharding": "cases round-robin over %d rank(s), no collective" % d.world,
                       "cases_per_rank": res["cases_per_rank"], "windows_per_forward": w["batch_windows"],
                       "l2": "867 MB of window logits per case exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": d.world * 1e3 / res["ms_e2e"], "unit": w["unit"], "h2d_bytes_per_step": res["h2d"],
                    "d2h_bytes_per_step": res["d2h"], "ms_per_step": res["ms_e2e"]},
            "gpu_launches_per_step": res["launches_per_step"], "roofline": roof,
            "aggregate_kernel": None if agg is None else {k: agg[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "us_per_launch", "algorithmic_per_launch")},
            "cpu_baseline": cpu}


# ------------------------------------------------------------------------------------------- torch-eager + cuDNN on the same GPU
def gpu_eager_baseline(w, dev, steps=3):
    """The reference's execution model on this GPU: the oracle restatement of the same CHAP iteration through torch eager + cuDNN
    (allow_tf32 on = reference default), largest-CC on the host like code/train_ours_2D.py:123-144.  A baseline leg, never the product."""
    from oracle import chap_losses as L
    from oracle import nets
    from oracle import train_step as ost
    from chap_b200 import networks
    torch.manual_seed(1337)
    if w["dims"] == 2:
        model = networks.DualDecoder(1, w["classes"], {"decoder_type": "mcnet"})
    else:
        model = networks.DualDecoder3d(1, w["classes"], normalization="batchnorm", has_dropout=True)
    sd = nets.clone_state_dict(model.state_dict(), requires_grad=True, device=dev)
    om = ost.OracleModel(sd, dims=w["dims"], has_dropout=(w["dims"] == 3), drop="torch")
    bufs = [None] * len(om.params())
    vat = L.VAT(10.0, 6.0, w["classes"])
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = True, True
    try:
        data = [tuple(t.to(dev) for t in synth_batch(w, i)) for i in range(2)]
        times = []
        for it in range(steps + 2):
            offs = L.draw_mask_offsets(w["shape"], np.random.RandomState(it))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ost.chap_train_step(om, bufs, data[it % 2][0], data[it % 2][1], w["labeled"], w["classes"], offs, it, vat=vat, topk=0.1)
            torch.cuda.synchronize()
            if it >= 2:
                times.append(time.perf_counter() - t0)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = old
    sec = float(np.mean(times))
    return {"value": 1.0 / sec, "unit": w["unit"], "ms_per_step": sec * 1e3, "steps": steps,
            "what": "oracle restatement of the same iteration on torch-eager + cuDNN (allow_tf32=True, cudnn.benchmark), host largest-CC like the reference, same GPU"}


# ------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch.distributed as dist
    from chap_b200 import _lib, ops
    d = Dist()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback; use --impl reference for the CPU arm)"
    torch.cuda.set_device(d.local)
    dev = torch.device("cuda", d.local)
    _lib.check(_lib.load().chap_check_device())
    if d.world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    threads = os.cpu_count() or 1
    use_graph = not args.no_graph
    kt = not args.no_kernel_timing
    want_cpu = d.world == 1 and d.rank == 0 and not args.no_cpu_baseline
    head_name = args.workload or "unet2d"
    extras = args.workload is None and not args.no_extras
    w = WORKLOADS[head_name]

    def run_workload(wl, steps, clocks_on):
        if wl["kind"] == "infer":
            res = time_sw_infer(args, wl, d, dev, max(2, min(steps, 8)), kt)
            cpu = cpu_baseline_for(wl, threads, CPU_BUDGET_S) if want_cpu else None
            return sw_block(args, wl, d, res, peaks, cpu), res
        res = time_training(args, wl, d, dev, steps, args.warmup, use_graph, kt, clocks_on)
        cpu = cpu_baseline_for(wl, threads, CPU_BUDGET_S) if want_cpu else None
        return training_block(args, wl, d, dev, res, peaks, cpu), res

    head, head_res = run_workload(w, args.steps, True)
    others, modes, eager = {}, None, None
    if extras:
        for name in ("vnet3d", "sw_infer"):
            others[name], _ = run_workload(WORKLOADS[name], args.steps, False)
        # the 2D iteration in the three convolution arithmetics; the headline is plain TF32
        modes = {"tf32": {"value": head["value"], "ms_per_step": head["ms_per_step"], "logits_vs_fp64": "1.7e-3 (= ideal TF32)"}}
        for label, c, note in (("hybrid32", 32, "6e-4"), ("precise", ops.PRECISE_ALL, "6e-6")):
            ops.set_conv_precision(c)
            r = time_training(args, w, d, dev, max(4, min(args.steps, 20)), args.warmup, use_graph, False, False)
            modes[label] = {"value": d.world * 1e3 / r["ms"], "ms_per_step": r["ms"], "loss_last": r["loss_last"], "logits_vs_fp64": note}
        ops.set_conv_precision(args.conv_precision or 0)
        modes["note"] = ("conv arithmetic of the forward / data-gradient tensor-core kernels: tf32 = plain TF32 (reference default, cuDNN allow_tf32); "
                         "hybrid32 = split-operand 3xTF32 on layers with <= 32 channels; precise = 3xTF32 everywhere (logits within 1e-3 of fp32: "
                         "tests/test_gpu_baseline_shapes.py); logits_vs_fp64 = measured relative L2 at b24 256^2")
        if d.rank == 0 and not args.no_eager_baseline:
            eager = gpu_eager_baseline(w, dev)
    if d.rank == 0:
        line = {"metric": head["metric"], "value": head["value"], "unit": head["unit"], "n_gpus": d.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32 tensor-core convs (fp32 accumulate) + f32 elsewhere" if not _lib.load().chap_get_force_simt() else "f32",
                "data": "synthetic", "config": dict(head["config"], conv_precision=_lib.load().chap_get_conv_precision()),
                "clocks": head_res.get("clocks"), "e2e": head["e2e"],
                "gpu_launches": int(head["gpu_launches_per_step"]) * (args.steps if w["kind"] == "train" else head_res["cases_per_rank"]),
                "cuda_graph": bool(use_graph and w["kind"] == "train"),
                "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"]}
        for k in ("loss_first", "loss_last", "loss_max", "loss_finite"):
            if k in head:
                line[k] = head[k]
        if head.get("aggregate_kernel"):
            line["aggregate_kernel"] = head["aggregate_kernel"]
        if others:
            line["workloads"] = others
        if modes:
            line["modes"] = modes
        if eager:
            line["gpu_eager_baseline"] = eager
        print(json.dumps(line))
        sys.stdout.flush()
        bad = [n for n, b in [(head_name, head)] + list(others.items()) if b.get("loss_finite") is False]
        if bad:
            print("bench.py: non-finite loss in workload(s) %s" % bad, file=sys.stderr)
    if d.world > 1:
        # every captured graph was dropped by ChapTrainer.close() above, so the communicator can be torn down normally
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="chap_b200", choices=["chap_b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="make this workload the headline and skip the extras (default: unet2d headline + vnet3d + sw_infer + modes)")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel eagerly instead of replaying one CUDA graph per step")
    ap.add_argument("--no-overlap", action="store_true", help="data parallel: one all-reduce after the backward instead of the overlapped two-bucket scheme")
    ap.add_argument("--force-simt", action="store_true", help="debug: fp32 CUDA-core convolutions only")
    ap.add_argument("--conv-precision", type=int, default=None,
                    help="split-operand 3xTF32 tensor-core convolutions for layers with max(Cin, Cout) <= this (0: plain TF32)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, WORKLOADS[args.workload or "unet2d"])
        return
    if args.force_simt:
        from chap_b200 import ops
        ops.set_force_simt(True)
    if args.conv_precision is not None:
        from chap_b200 import ops
        ops.set_conv_precision(args.conv_precision)
    run_gpu(args)


if __name__ == "__main__":
    main()

"""bench.py -- CHAP training iterations/sec on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

Workload at every N (weak scaling, one process per GPU): BASELINE.json configs[1] -- one full CHAP
training iteration of the 2D DualDecoder U-Net on synthetic ACDC-shaped input (batch 24 of which 12
labelled, 1x256x256, 4 classes; --adv_noise, 'kl' consistency; channel+spatial hierarchical perturbation).
`--workload vnet3d` selects configs[2] (DualDecoder3d, batch 4 / 2 labelled, 112x112x80, 2 classes).
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
    "unet2d": dict(dims=2, batch=24, labeled=12, shape=(256, 256), classes=4,
                   name="UNet(DualDecoder) 2D CHAP full training step, b24 (12 labelled), 1x256x256, 4 classes"),
    "vnet3d": dict(dims=3, batch=4, labeled=2, shape=(112, 112, 80), classes=2,
                   name="VNet(DualDecoder3d) 3D CHAP full training step, b4 (2 labelled), 1x112x112x80, 2 classes"),
}
METRIC = "CHAP train iters/sec"


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


def build_model(w, device):
    from chap_b200 import networks
    torch.manual_seed(1337)
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


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_chap_iteration(w, batch, labeled, threads, steps, warmup):
    """The reference's CPU path: oracle port of one CHAP iteration (oracle/train_step.py) on `threads` host threads.
    Returns seconds per iteration (mean over `steps`)."""
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
    times = []
    for it in range(warmup + steps):
        vol, lab = synth_batch(w, it, batch)
        offs = L.draw_mask_offsets(w["shape"], np.random.RandomState(it))
        t0 = time.perf_counter()
        ost.chap_train_step(om, bufs, vol, lab, labeled, w["classes"], offs, it, vat=vat, topk=0.1)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times))


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample: the full batch is 24 (12 labelled); time a batch-4 (2 labelled) slice of the same iteration
    sample_batch, sample_labeled = (4, 2) if w["dims"] == 2 else (4, 2)
    sample_shape = w["shape"] if w["dims"] == 2 else (64, 64, 48)
    ws = dict(w, shape=sample_shape)
    sec = cpu_chap_iteration(ws, sample_batch, sample_labeled, threads, args.steps, min(args.warmup, 1))
    vox_full = w["batch"] * float(np.prod(w["shape"]))
    vox_sample = sample_batch * float(np.prod(sample_shape))
    sec_full = sec * vox_full / vox_sample
    value = 1.0 / sec_full
    sample = ("one CHAP iteration of the oracle port (reference nets restated on torch-CPU + frozen losses) on batch %d of %s "
              "(%d labelled), scaled linearly in voxels to the full batch %d of %s" %
              (sample_batch, "x".join(map(str, sample_shape)), sample_labeled, w["batch"], "x".join(map(str, w["shape"]))))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "it/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": sec_full * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "where": "host CPU"},
            "cpu_baseline": {"value": value, "unit": "it/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args, w):
    import torch.distributed as dist
    from chap_b200 import _lib
    from chap_b200.train_step import ChapTrainer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback; use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.check(_lib.load().chap_check_device())
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = build_model(w, dev)

    def allreduce(flat_g):                       # data-parallel: one flat-bucket NCCL all-reduce per iteration
        dist.all_reduce(flat_g)
    use_graph = not args.no_graph
    trainer = ChapTrainer(model, n_classes=w["classes"], labeled_bs=w["labeled"], max_iterations=30000,
                          grad_hook=allreduce if world > 1 else None, grad_scale=1.0 / world,
                          use_graph=use_graph, graph_warmup=2)
    n_in = max(2, min(4, args.steps))
    host = [synth_batch(w, 1000 * rank + i) for i in range(n_in)]
    host = [(v.pin_memory(), l.pin_memory()) for v, l in host]
    resident = [(v.to(dev), l.to(dev)) for v, l in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value"): the whole iteration is ONE CUDA-graph replay per step
    warm = max(args.warmup, 3 if use_graph else args.warmup)     # graph mode: 2 eager iterations + the capture step
    for i in range(warm):
        trainer.step(*resident[i % n_in])
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = trainer.step(*resident[i % n_in])
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    clk = clocks.stop() if clocks else None

    # ---- end-to-end timing ("e2e"): host (pinned) inputs -> H2D -> step -> loss read back, every step
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        v, l = host[i % n_in]
        out = trainer.step(v.to(dev, non_blocking=True), l.to(dev, non_blocking=True))
        loss_host = float(out["loss"])                    # D2H of the step's result
    t1.record()
    barrier()
    ms_e2e = torch.tensor([t0.elapsed_time(t1)], device=dev)

    # ---- kernel-family pass (roofline, launch count): the SAME iteration issued eagerly so that CUDA events can
    # bracket individual launches on the launching stream (events cannot be read back from inside a graph replay)
    fam, launches_per_step, ms_eager = {}, 0, None
    if not args.no_kernel_timing:
        trainer.use_graph = False
        trainer.step(*resident[0])
        barrier()
        _lib.reset_launch_count()
        _lib.timing_enable(rank == 0)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 2
        r0.record()
        for i in range(reps):
            trainer.step(*resident[i % n_in])
        r1.record()
        barrier()
        ms_eager = r0.elapsed_time(r1) / reps
        launches_per_step = _lib.launch_count() // reps
        fam = _lib.timing_report() if rank == 0 else {}
        fam = {k: dict(v, ms=v["ms"] / reps, launches=v["launches"] // reps, flops=v["flops"] / reps, bytes=v["bytes"] / reps)
               for k, v in fam.items()}
        _lib.timing_enable(False)
        trainer.use_graph = use_graph
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(ms), float(ms_e2e)
    if rank == 0:
        peaks = measured_peaks()
        vox = float(np.prod(w["shape"]))
        # per step: volume f32 + label i64, plus the trainer's own small uploads (copy-paste mask i64 [*spatial], lr and
        # consistency-weight scalars); the CC filter runs on the device, nothing else crosses the bus
        h2d = w["batch"] * vox * (4 + 8) + vox * 8 + 8
        d2h = 4                                                         # the loss (f32 scalar) read back every step
        roof = None
        if fam:
            # fam holds per-shape entries ("family:t9:k16:n16:256x256x1:r786432", CHAP_TIMING_DETAIL); families = their sums
            families = {}
            for k, v in fam.items():
                f = families.setdefault(k.split(":")[0], dict(ms=0.0, launches=0, flops=0.0, bytes=0.0))
                for key in f:
                    f[key] += v[key]
            top_family = max(families.items(), key=lambda kv: kv[1]["ms"])[0]    # dominant kernel family of the step ...
            name, f = max(((k, v) for k, v in fam.items() if k.split(":")[0] == top_family), key=lambda kv: kv[1]["ms"])   # ... its costliest layer shape
            per_launch_s = f["ms"] / 1e3 / max(f["launches"], 1)
            tf32_peak = peaks["tensor"] / 2.0                                     # kind::tf32 runs at half the dense bf16 rate
            intensity = f["flops"] / f["bytes"] if f["bytes"] > 0 else float("inf")
            tensor_bound = f["flops"] > 0 and intensity > tf32_peak * 1e12 / (peaks["hbm"] * 1e9)
            if tensor_bound:
                achieved = f["flops"] / f["launches"] / per_launch_s / 1e12
                roof = {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                        "frac": achieved / tf32_peak}
            else:
                achieved = f["bytes"] / f["launches"] / per_launch_s / 1e9
                roof = {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": achieved / peaks["hbm"]}
            traffic_table = {}
            tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
            if os.path.isfile(tpath):
                traffic_table = json.load(open(tpath)).get("dram_bytes_per_launch", {})
            roof["traffic"] = traffic_table.get(name)                              # ncu --set full: dram read + write bytes, per launch
            roof["algorithmic_per_launch"] = {"flops": f["flops"] / f["launches"], "bytes": f["bytes"] / f["launches"],
                                              "flop_per_byte": intensity}
            roof["us_per_launch"] = per_launch_s * 1e6
            roof["peak_source"] = peaks["source"] + ("; tensor peak = measured dense bf16 / 2 (tf32)" if tensor_bound else "")
            roof["family"] = top_family
            roof["family_share_of_step"] = families[top_family]["ms"] / (ms / args.steps)
            roof["launches_per_step"] = f["launches"]
            roof["share_of_step"] = f["ms"] / (ms / args.steps)
            roof["families_ms_per_step"] = {k: round(v["ms"], 4) for k, v in sorted(families.items(), key=lambda kv: -kv[1]["ms"])}
            roof["families_gbps"] = {k: round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1) for k, v in families.items()
                                     if v["bytes"] > 0 and v["flops"] == 0 and v["ms"] > 0}
            roof["top_kernels_us"] = {k: [v["launches"], round(1e3 * v["ms"] / v["launches"], 1)]
                                      for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])[:12]}
            roof["eager_ms_per_step"] = ms_eager
            roof["note"] = ("per-kernel CUDA-event timing on the launching stream, taken in an eager re-issue of the same iteration "
                            "right after the timed graph replays; the dominant kernel is one layer shape of one kernel family; "
                            "bound chosen by its algorithmic FLOP/byte against the measured machine balance")
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sb, sl = 4, 2
            sshape = w["shape"] if w["dims"] == 2 else (64, 64, 48)
            sec = cpu_chap_iteration(dict(w, shape=sshape), sb, sl, threads, 2, 1)
            sec_full = sec * (w["batch"] * vox) / (sb * float(np.prod(sshape)))
            cpu = {"value": 1.0 / sec_full, "unit": "it/s", "cores": threads, "kind": "port",
                   "sample": "2 timed CHAP iterations of the oracle port on batch %d of %s, scaled linearly in voxels to batch %d of %s"
                             % (sb, "x".join(map(str, sshape)), w["batch"], "x".join(map(str, w["shape"])))}
        line = {"metric": METRIC, "value": world * args.steps / (ms / 1e3), "unit": "it/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "tf32 tensor-core convs (fp32 accumulate) + f32 elsewhere" if not _lib.load().chap_get_force_simt() else "f32",
                "data": "synthetic",
                "config": {"workload": w["name"], "per_gpu_batch": w["batch"], "parallelism": "dp%d" % world,
                           "flags": "--adv_noise --adv_losstype kl --decoder_type mcnet --noise_mag 10 epi 6 topk 0.1",
                           "l2": "activations touched per step (>5 GB) far exceed the 126 MB L2; no explicit flush",
                           "loss_last": float(out["loss"])},
                "clocks": clk,
                "e2e": {"value": world * args.steps / (ms_e2e / 1e3), "unit": "it/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches_per_step) * args.steps,
                "cuda_graph": bool(use_graph),
                "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        # A captured NCCL all-reduce keeps communicator resources alive; tearing the process group down while the CUDA
        # graph still exists was observed to hang on exit (2 ranks, torch 2.11 / NCCL 2.28).  Everything is measured and
        # printed: synchronise, drop the graph, and leave without the collective teardown.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        trainer.graph = None
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="chap_b200", choices=["chap_b200", "reference"])
    ap.add_argument("--workload", default="unet2d", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel eagerly instead of replaying one CUDA graph per step")
    ap.add_argument("--force-simt", action="store_true", help="debug: fp32 CUDA-core convolutions only")
    ap.add_argument("--conv-precision", type=int, default=None,
                    help="split-operand 3xTF32 tensor-core convolutions for layers with max(Cin, Cout) <= this (0: plain TF32)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
        return
    if args.force_simt:
        from chap_b200 import ops
        ops.set_force_simt(True)
    if args.conv_precision is not None:
        from chap_b200 import ops
        ops.set_conv_precision(args.conv_precision)
    run_gpu(args, w)


if __name__ == "__main__":
    main()

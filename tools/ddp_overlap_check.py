"""2-GPU check (torchrun): the overlapped two-bucket all-reduce gives the same training trajectory as one all-reduce after the
backward, and both ranks hold identical parameters afterwards.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_overlap_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
from chap_b200 import ops  # noqa: E402
from chap_b200.train_step import ChapTrainer  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
w = dict(bench.WORKLOADS["unet2d"], shape=(64, 64), batch=8, labeled=4)
ops.set_force_simt(True)                                   # fp32 kernels: differences between the two schemes = summation order only
results = {}
for overlap in (True, False):
    for use_graph in (False, True):
        m = bench.build_model(w, dev)
        for mod in m.modules():
            if hasattr(mod, "dropout_p"):
                mod.dropout_p = 0.0
        t = ChapTrainer(m, n_classes=4, labeled_bs=4, max_iterations=100, adv_noise=False, grad_hook=lambda g: dist.all_reduce(g),
                        grad_scale=1.0 / world, use_graph=use_graph, graph_warmup=1, overlap_allreduce=overlap)
        assert t.overlap == overlap
        for it in range(4):
            vol, lab = bench.synth_batch(w, 10 * rank + it)
            out = t.step(vol.to(dev), lab.to(dev), mask_offsets=(3, 5))
        flat = t.opt.flat_p.clone()
        t.close()
        other = flat.clone()
        dist.broadcast(other, 0)
        assert torch.equal(other, flat) or float((other - flat).abs().max()) == 0.0, "ranks diverged"      # replicas stay bit-identical
        results[(overlap, use_graph)] = (flat, float(out["loss"]))
ref = results[(False, False)][0]
for k, (flat, loss) in results.items():
    err = float((flat - ref).norm() / ref.norm())
    if rank == 0:
        print("overlap=%s graph=%s  loss %.6f  params vs single-bucket eager: %.2e" % (k[0], k[1], loss, err))
    assert err < 1e-4, (k, err)
if rank == 0:
    print("ddp overlap check OK")
dist.barrier()
dist.destroy_process_group()

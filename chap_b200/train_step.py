"""One CHAP training iteration (2D ACDC-shaped U-Net or 3D LA-shaped V-Net) on the sm_100a kernels.

Host-side orchestration mirroring the body of the reference's training loop,
code/train_ours_2D.py:304-389 (flags --adv_noise, --adv_losstype kl|dice; the --dropout branch is
not part of this step).  The reference has no 3D trainer; the 3D step is the same
procedure on DualDecoder3d with a cubic copy-paste mask (frozen in oracle/train_step.py).

Device work: conv/BN/activation stacks, pseudo-label block, mix losses, patch mask, the perturbation
generator, consistency losses, the largest-connected-component filter of get_ACDC_2DLargestCC
(code/train_ours_2D.py:123-144; a union-find kernel, no host round trip) and the fused SGD-momentum update are all
libchap_b200 kernels.  There is no CPU path: everything here raises if the library or a CUDA device is missing.
"""
import numpy as np
import torch

from . import ops
from .utils import losses, patch, ramps


def generate_mask(img, offsets=None):
    """generate_mask, code/train_ours_2D.py:91-101 (dimension generic).  Returns (mask[*spatial] int64,
    loss_mask[N, *spatial] int64); offsets = the np.random.randint draws, drawn here when None."""
    spatial = tuple(img.shape[2:])
    if offsets is None:
        offsets = tuple(int(np.random.randint(0, s - int(s * 2 / 3))) for s in spatial)
    mask = torch.ones(spatial, dtype=torch.int64, device=img.device)
    mask[tuple(slice(o, o + int(s * 2 / 3)) for o, s in zip(offsets, spatial))] = 0
    return mask, mask.unsqueeze(0).expand((img.shape[0],) + spatial)


def largest_cc_labels(seg, n_classes):
    """get_ACDC_2DLargestCC (code/train_ours_2D.py:123-144): keep the largest connected component of every
    foreground class per sample (full connectivity, first component wins ties) -- union-find kernel, no host sync."""
    return ops.largest_cc(seg, n_classes)


def get_masks(output, n_classes, nms=1):
    """get_ACDC_masks, code/train_ours_2D.py:103-108."""
    probs = ops.argmax(output)
    return largest_cc_labels(probs, n_classes) if nms == 1 else probs


def consistency_weight(iter_num, consistency=1.0, rampup=50.0):
    """get_current_consistency_weight, code/train_ours_2D.py:34-36 with the call at :356."""
    return consistency * ramps.sigmoid_rampup(iter_num // 150, rampup)


def feature_dropout_loss(model, uimg_ab, ps1, ps2, sim_score=None, comp_drop=False, dropout_masks=None):
    """The --dropout branch, code/train_ours_2D.py:359-364: model(uimg_ab, False, True, [0,1,2,3,4], sim_score, False), then
    cross-entropy of each decoder's output against the other decoder's pseudo-labels.  As shipped the reference passes
    [1.5 N] logits and [N] targets to F.cross_entropy (:362-363), which raises; the frozen fix (oracle/train_step.py) is
    target = cat(pseudo, pseudo[N/2:]) -- the appended rows are perturbed copies of samples N/2...  The mean CE over all
    rows and pixels comes from the fused Dice/CE kernel with an all-ones mask."""
    o1, o2 = model(uimg_ab, False, True, [0, 1, 2, 3, 4], sim_score, comp_drop, dropout_masks=dropout_masks)
    half = ps1.shape[0] // 2
    t1, t2 = torch.cat((ps1, ps1[half:])), torch.cat((ps2, ps2[half:]))
    ones = torch.ones(tuple(ps1.shape[1:]), dtype=torch.int64, device=ps1.device)
    return (losses.masked_dice_ce(o1, t2, ones)[1] + losses.masked_dice_ce(o2, t1, ones)[1]).float()


def chap_losses_forward(model, volume, label, labeled_bs, n_classes, iter_num, vat=None, adv_losstype="kl",
                        topk=0.1, use_diff_mask=True, consistency=1.0, rampup=50.0, mask_offsets=None,
                        d_init=None, trace=None, img_mask=None, cw=None, dropout=False, comp_drop=False, sim_score=None,
                        dropout_masks=None, on_decoders_done=None):
    """Forward part of the iteration; returns (loss, aux).  Line numbers: code/train_ours_2D.py.
    img_mask (int64 [*spatial]) / cw (float or 0-dim device tensor) may be supplied by the caller (the CUDA-graph
    trainer keeps them in static device buffers); otherwise they are drawn / computed here like the reference does."""
    n = volume.shape[0]
    sub_l, sub_u = labeled_bs // 2, (n - labeled_bs) // 2                               # :295
    img_a, img_b = volume[:sub_l], volume[sub_l:labeled_bs]                             # :307
    uimg_a, uimg_b = volume[labeled_bs:labeled_bs + sub_u], volume[labeled_bs + sub_u:]
    lab_a, lab_b = label[:sub_l], label[sub_l:labeled_bs]                               # :310
    uimg_ab = volume[labeled_bs:]                                                       # :312 (cat of adjacent rows)

    with torch.no_grad():                                                               # :314-333
        pre1, pre2 = model(uimg_ab)
        soft1, soft2, ps1, ps2, knowledge = ops.pseudo_label(pre1, pre2)
        plab1 = largest_cc_labels(ps1, n_classes)          # argmax(softmax) + largest CC, all 2*sub_u rows at once
        plab2 = largest_cc_labels(ps2, n_classes)
        plab_a1, plab_b1 = plab1[:sub_u], plab1[sub_u:]
        plab_a2, plab_b2 = plab2[:sub_u], plab2[sub_u:]
        if img_mask is None:
            img_mask, _ = generate_mask(img_a, mask_offsets)
        loss_mask = img_mask
        net_input_unl = ops.mask_mix(uimg_a, img_a, img_mask)                           # :335
        net_input_l = ops.mask_mix(img_b, uimg_b, img_mask)                             # :336
        net_input_mix = torch.cat((net_input_l, net_input_unl))                         # :338

    if on_decoders_done is not None and hasattr(model, "encoder"):
        # The mix pass is the FIRST pass recorded with gradients, so autograd back-propagates it LAST (the engine runs the most
        # recently recorded nodes first): when the gradients of all its encoder features have arrived, every use of the decoder
        # weights in this iteration has been back-propagated -- the decoders' gradient bucket is final while the encoder
        # backward of this pass is still to come.  (Same computation as model(net_input_mix), code/networks/unet.py:277-292.)
        feats = model.encoder(net_input_mix)
        out1, out2 = model.decoder1(feats), model.decoder2(feats)                       # :339
        pending = [len(feats)]

        def _arrived(grad):
            pending[0] -= 1
            if pending[0] == 0:
                on_decoders_done()
            return None
        for f in feats:
            f.register_hook(_arrived)
    else:
        out1, out2 = model(net_input_mix)                                               # :339
    out_l1, out_unl1 = out1[:sub_l], out1[sub_l:]
    out_l2, out_unl2 = out2[:sub_l], out2[sub_l:]
    lu_o1, ll_i1, m1 = losses.mix_loss(out_unl1, plab_a2, lab_a, loss_mask, u_weight=0.5, unlab=True)   # :345
    lu_o2, ll_i2, m2 = losses.mix_loss(out_unl2, plab_a1, lab_a, loss_mask, u_weight=0.5, unlab=True)   # :346
    ll_o1, lu_i1, m3 = losses.mix_loss(out_l1, lab_b, plab_b2, loss_mask, u_weight=0.5)                 # :348
    ll_o2, lu_i2, m4 = losses.mix_loss(out_l2, lab_b, plab_b1, loss_mask, u_weight=0.5)                 # :349
    bcp_loss = m1 + m2 + m3 + m4                                                        # :351
    loss_l = ll_i1 + ll_i2 + ll_o1 + ll_o2
    loss_u = lu_i1 + lu_i2 + lu_o1 + lu_o2
    if cw is None:
        cw = consistency_weight(iter_num, consistency, rampup)                          # :356

    if dropout:                                                                         # :359-365 (GradSim is absent: sim_score is an input)
        fp_loss = feature_dropout_loss(model, uimg_ab, ps1, ps2, sim_score, comp_drop, dropout_masks)
    else:
        fp_loss = torch.zeros((), device=volume.device)
    if vat is not None:                                                                 # :369-372
        diff_mask = patch.create_maskV1(ps1, ps2, knowledge, scale_factor=4, topk=topk) if use_diff_mask else None
        vat_loss = vat(model, volume, soft1, soft2, diff_mask, adv_losstype, d_init=d_init, trace=trace)
    else:
        vat_loss = torch.zeros((), device=volume.device)
    loss = bcp_loss + cw * (fp_loss + vat_loss)                                         # :378
    aux = dict(bcp_loss=bcp_loss.detach(), vat_loss=vat_loss.detach(), fp_loss=fp_loss.detach(), loss_l=loss_l.detach(),
               loss_u=loss_u.detach(), cw=cw, soft1=soft1, soft2=soft2, knowledge=knowledge,
               plab=(plab_a1, plab_b1, plab_a2, plab_b2), out_mix=(out1.detach(), out2.detach()))
    return loss, aux


class FlatSGD:
    """torch.optim.SGD(lr, momentum=0.9, weight_decay=1e-4) semantics (code/train_ours_2D.py:278,383)
    on one flat fp32 arena: parameters are re-pointed at views of a single buffer, gradients are
    gathered into a matching flat buffer, and ONE fused kernel updates everything.  `grad_hook`
    (optional) is called on the flat gradient before the update -- the data-parallel all-reduce.
    The learning rate lives in a device scalar so that the update can be replayed from a CUDA graph."""

    def __init__(self, params, lr, momentum=0.9, weight_decay=1e-4, grad_hook=None, grad_sink=True):
        self.params = [p for p in params if p.requires_grad]
        self.momentum, self.weight_decay, self.grad_hook = momentum, weight_decay, grad_hook
        self.sink = bool(grad_sink)
        dev = self.params[0].device
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4                     # keep every slot 16-byte aligned
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_buf = torch.zeros(total, dtype=torch.float32, device=dev)     # zeros: first step gives buf = g
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr = float(lr)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
        self.grad_views = [self.flat_g[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]
        if self.sink:
            # gradient-sink mode (ops.grad_sink): the kernels add every parameter gradient straight into flat_g.  Conv weights also
            # get a persistent zeroed scratch (the tensor-core weight-gradient kernel reduces into a [tap][M][N] image with vector
            # atomics, its unpack step adds that into the arena and leaves the scratch zeroed again).
            self.flat_ws = torch.zeros(total, dtype=torch.float32, device=dev)
            for p, off, gv in zip(self.params, self.offsets, self.grad_views):
                scratch = self.flat_ws[off:off + p.numel()] if p.dim() >= 3 else None
                p._chap_sink = (gv, scratch)
                if p.dim() == 1:        # BatchNorm gammas (and, harmlessly, other vectors): persistent zeroed reduction buffer of the
                    p._chap_bwd_sums = torch.zeros(2 * p.numel() + 1, dtype=torch.float64, device=dev)    # BatchNorm backward
        ops.invalidate_weight_cache()

    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        """Eager-mode setter: a stream-ordered fill of the device scalar (no pinned staging buffer that a host running
        ahead could overwrite).  The CUDA-graph trainer never calls this -- its schedule is computed on the device."""
        self._lr = float(value)
        self.lr_dev.fill_(self._lr)

    def zero_grad(self):
        for p in self.params:
            p.grad = None
        if self.sink:
            self.flat_g.zero_()                 # ONE fill for the whole arena; the kernels accumulate into it during backward

    def gather_grads(self):
        if self.sink:
            # everything the kernels produced is already in flat_g; a gradient that still arrived through autograd (an op
            # outside this library touching a parameter) is added on top
            with torch.no_grad():
                stray = [(v, p.grad) for p, v in zip(self.params, self.grad_views) if p.grad is not None]
                if stray:
                    torch._foreach_add_([v for v, _ in stray], [g for _, g in stray])
            return
        with torch.no_grad():
            missing = [v for p, v in zip(self.params, self.grad_views) if p.grad is None]
            have = [(v, p.grad) for p, v in zip(self.params, self.grad_views) if p.grad is not None]
            if missing:
                torch._foreach_zero_(missing)
            if have:
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])

    # -- data-parallel all-reduce in two buckets: the tail bucket [tail_start, end) -- the decoders -- is final as soon as the
    #    last network pass has back-propagated through the decoders, i.e. BEFORE the encoder backward of that pass; its
    #    all-reduce runs on a side stream underneath that encoder backward (ChapTrainer triggers it), the head bucket follows here.
    def set_tail_bucket(self, first_tail_param):
        """Parameters from `first_tail_param` on (arena order = model.parameters() order) form the early bucket."""
        idx = next(i for i, p in enumerate(self.params) if p is first_tail_param)
        self.tail_start = self.offsets[idx]
        self.comm_stream = torch.cuda.Stream(device=self.flat_g.device)
        self._tail_pending = False

    def early_reduce_tail(self):
        if self.grad_hook is None or getattr(self, "tail_start", None) is None or not self.sink or self._tail_pending:
            return
        cur = torch.cuda.current_stream()
        self.comm_stream.wait_stream(cur)                     # everything that accumulated into the tail bucket is ordered before
        with torch.cuda.stream(self.comm_stream):
            self.grad_hook(self.flat_g[self.tail_start:])
        self._tail_pending = True

    def step(self, grad_scale=1.0):
        self.gather_grads()
        if self.grad_hook is not None:
            if getattr(self, "_tail_pending", False):
                self.grad_hook(self.flat_g[:self.tail_start])                       # head bucket (encoder), on the main stream
                torch.cuda.current_stream().wait_stream(self.comm_stream)           # join the overlapped tail all-reduce
                self._tail_pending = False
            else:
                self.grad_hook(self.flat_g)
        ops.sgd_momentum_lrdev_(self.flat_p, self.flat_g, self.flat_buf, self.lr_dev, self.momentum, self.weight_decay, grad_scale)

    # -- optimizer-state save / resume.  The reference only saves model.state_dict() (code/train_ours_2D.py:428-435) and has no
    #    resume logic; the layout below is torch.optim.SGD's own (`state[i]['momentum_buffer']`, param_groups), so a checkpoint
    #    written here loads into a torch.optim.SGD over the same parameter list and vice versa.
    def state_dict(self):
        state = {i: {"momentum_buffer": self.flat_buf[off:off + p.numel()].view_as(p).detach().clone()}
                 for i, (p, off) in enumerate(zip(self.params, self.offsets))}
        group = {"lr": self._lr, "momentum": self.momentum, "dampening": 0, "weight_decay": self.weight_decay, "nesterov": False,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("FlatSGD.load_state_dict: expected one param group with %d parameters" % len(self.params))
        g = groups[0]
        self.momentum, self.weight_decay = float(g["momentum"]), float(g["weight_decay"])
        self.lr = float(g["lr"])
        with torch.no_grad():
            self.flat_buf.zero_()                                  # torch creates momentum buffers lazily: missing == zeros here
            for i, (p, off) in enumerate(zip(self.params, self.offsets)):
                st = sd["state"].get(i)
                if st is not None and st.get("momentum_buffer") is not None:
                    self.flat_buf[off:off + p.numel()].view_as(p).copy_(st["momentum_buffer"])


class ChapTrainer:
    """State + `step(volume, label)` for CHAP training of a DualDecoder (2D) or DualDecoder3d (3D).

    use_graph=True: after `graph_warmup` eager iterations the whole iteration (3 network passes + the VAT probe,
    pseudo-labels, largest-CC filter, losses, backward, gradient gather, all-reduce hook, fused SGD) is captured ONCE
    into a CUDA graph and replayed; per-iteration host work is three tiny H2D scalars/masks plus the input copy into
    static buffers.  Everything that changes between iterations lives in device memory: inputs, the copy-paste mask
    (np.random offsets, code/train_ours_2D.py:97-98), the consistency weight (:356) and the learning rate (:387)."""

    def __init__(self, model, n_classes, labeled_bs, base_lr=0.01, max_iterations=30000, adv_noise=True,
                 adv_losstype="kl", noise_mag=10.0, epi=6.0, topk=0.1, consistency=1.0, consistency_rampup=50.0,
                 use_diff_mask=True, grad_hook=None, grad_scale=1.0, use_graph=False, graph_warmup=3,
                 dropout=False, comp_drop=False, sim_score=None, grad_sink=True, overlap_allreduce=True):
        """dropout / comp_drop / sim_score: the --dropout feature-perturbation branch (code/train_ours_2D.py:359-365, 2D nets only:
        the reference's DualDecoder3d.forward has no such branch); sim_score stands in for the absent GradSim.get_sim()."""
        self.model, self.n_classes, self.labeled_bs = model, n_classes, labeled_bs
        self.dropout, self.comp_drop, self.sim_score = dropout, comp_drop, sim_score
        if dropout and not hasattr(model.encoder, "ft_chns"):
            raise NotImplementedError("--dropout: only the 2D DualDecoder has the perform_dropout branch (code/networks/unet.py:280-284)")
        if dropout and use_graph and sim_score is not None:
            raise NotImplementedError("score-driven feature dropout draws its masks from the activations; use use_graph=False")
        self.base_lr, self.max_iterations = base_lr, max_iterations
        self.vat = losses.VAT2d(xi=noise_mag, epi=epi, num_classes=n_classes) if adv_noise else None
        self.adv_losstype, self.topk, self.use_diff_mask = adv_losstype, topk, use_diff_mask
        self.consistency, self.rampup = consistency, consistency_rampup
        self.opt = FlatSGD(model.parameters(), base_lr, grad_hook=grad_hook, grad_sink=grad_sink)
        # overlapped two-bucket gradient all-reduce (data parallel only): decoders early on a side stream, encoder at the end
        self.overlap = bool(overlap_allreduce and grad_hook is not None and grad_sink and hasattr(model, "decoder1"))
        if self.overlap:
            self.opt.set_tail_bucket(next(iter(model.decoder1.parameters())))
        self.grad_scale = grad_scale
        self.iter_num = 0
        self.use_graph, self.graph_warmup = use_graph, graph_warmup
        self.graph, self.static = None, None
        model.train()

    # ------------------------------------------------------------------ one iteration on given device tensors
    def _draw_dropout_masks(self, n_unlab, device):
        """Host-side draw of the perform_dropout masks for the n_unlab-row batch uimg_ab (its second half is perturbed) (scores None: the masks do not
        depend on the activations) -- placeholders stand in for the features, draw_dropout_masks only reads their shapes."""
        from .networks.FilterDropout import draw_dropout_masks
        fake = [torch.empty(n_unlab, c, 1, 1, device=device) for c in self.model.encoder.ft_chns]
        return draw_dropout_masks(fake, [0, 1, 2, 3, 4], None, self.comp_drop)

    def _iteration(self, volume, label, img_mask=None, cw=None, mask_offsets=None, d_init=None, trace=None, dropout_masks=None):
        ops.pack_all(self.model)                 # the previous optimiser step changed every weight: one batched re-pack
        loss, aux = chap_losses_forward(self.model, volume, label, self.labeled_bs, self.n_classes, self.iter_num,
                                        vat=self.vat, adv_losstype=self.adv_losstype, topk=self.topk,
                                        use_diff_mask=self.use_diff_mask, consistency=self.consistency,
                                        rampup=self.rampup, mask_offsets=mask_offsets, d_init=d_init, trace=trace,
                                        img_mask=img_mask, cw=cw, dropout=self.dropout, comp_drop=self.comp_drop,
                                        sim_score=self.sim_score, dropout_masks=dropout_masks,
                                        on_decoders_done=self.opt.early_reduce_tail if self.overlap else None)
        self.opt.zero_grad()                                                            # :381
        with ops.zero_bias_grad_as_none(), ops.grad_sink(self.opt.sink):
            loss.backward()                                                             # :382
        self.opt.step(self.grad_scale)                                                  # :383
        aux["loss"] = loss.detach()
        return aux

    def _poly_lr(self):
        return self.base_lr * (1.0 - self.iter_num / self.max_iterations) ** 0.9        # set at :387-389 of the previous iteration

    def step(self, volume, label, mask_offsets=None, d_init=None, trace=None, dropout_masks=None):
        """One iteration.  d_init: optional explicit VAT probing noise (list of 5 tensors; the parity protocol) -- in graph
        mode it is copied into static buffers that the captured graph reads; without it the noise is drawn inside the graph.
        dropout_masks: optional explicit perform_dropout masks (list of 5: None or (m1, m2) [n_unlab / 2, C]); drawn like the
        reference when omitted."""
        eager = (not self.use_graph) or trace is not None
        if eager or self.iter_num < self.graph_warmup:
            self.opt.lr = self._poly_lr()
            aux = self._iteration(volume, label, mask_offsets=mask_offsets, d_init=d_init, trace=trace, dropout_masks=dropout_masks)
            self.iter_num += 1                                                          # :385
            return aux
        if self.graph is None:
            self._capture(volume, label, d_init)
        st = self.static
        if (d_init is not None) != (st["d_init"] is not None):
            raise RuntimeError("ChapTrainer: the graph was captured %s explicit d_init; keep passing it the same way"
                               % ("with" if st["d_init"] is not None else "without"))
        # per-iteration host work: stream-ordered copies into the static inputs; lr / consistency weight / iteration counter
        # live on the device and are advanced by a kernel inside the graph (no pinned scalars a fast host could overwrite)
        st["volume"].copy_(volume, non_blocking=True)
        st["label"].copy_(label, non_blocking=True)
        mask, _ = generate_mask(volume[:1], mask_offsets)
        st["mask"].copy_(mask, non_blocking=True)
        if d_init is not None:
            for dst, src in zip(st["d_init"], d_init):
                dst.copy_(src, non_blocking=True)
        if self.dropout:                       # the masks are drawn outside the graph (host generator, like the reference) ...
            if dropout_masks is None:
                dropout_masks = self._draw_dropout_masks(volume.shape[0] - self.labeled_bs, volume.device)
            for dst, src in zip(st["drop"], dropout_masks):
                dst[0].copy_(src[0], non_blocking=True)             # ... and copied into the static buffers the graph reads
                dst[1].copy_(src[1], non_blocking=True)
        if st["next_iter"] != self.iter_num:   # eager iterations ran in between: resynchronise the device counter
            st["iter"].fill_(self.iter_num)
        self.graph.replay()
        ops.invalidate_weight_cache()          # the replay updated the weights behind the Python-side pack cache
        self.iter_num += 1
        st["next_iter"] = self.iter_num
        self.opt._lr = self._poly_lr()         # host mirror of the device schedule (what the NEXT eager step / a checkpoint sees)
        return st["aux"]

    def _capture(self, volume, label, d_init=None):
        dev = volume.device
        st = dict(volume=torch.empty_like(volume), label=torch.empty_like(label),
                  mask=torch.ones(tuple(volume.shape[2:]), dtype=torch.int64, device=dev),
                  cw=torch.zeros(1, dtype=torch.float32, device=dev),
                  iter=torch.full((1,), self.iter_num, dtype=torch.int64, device=dev),
                  d_init=None if d_init is None else [ops.cl(d.to(dev)).clone() for d in d_init], next_iter=self.iter_num,
                  drop=self._draw_dropout_masks(volume.shape[0] - self.labeled_bs, dev) if self.dropout else None)
        st["volume"].copy_(volume)
        st["label"].copy_(label)
        ops.invalidate_weight_cache()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        ops.set_dropout_epoch(st["iter"])          # generated nn.Dropout masks: the device-side iteration counter keys every replay differently
        try:
            with torch.cuda.graph(self.graph):
                ops.schedule_step(st["iter"], self.base_lr, self.max_iterations, self.consistency, self.rampup, self.opt.lr_dev, st["cw"])
                st["aux"] = self._iteration(st["volume"], st["label"], img_mask=st["mask"], cw=st["cw"][0], d_init=st["d_init"],
                                            dropout_masks=st["drop"])
        finally:
            ops.set_dropout_epoch(None)
        self.static = st

    def close(self):
        """Drop the captured graph (and its static buffers).  Call before tearing down an NCCL process group whose
        all-reduce was captured: destroying the communicator while a graph still references it hangs."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph = None
            self.static = None

    # -- checkpoint: model.state_dict() is the reference's own artefact (:428-435); optimizer + iteration make it resumable
    def state_dict(self):
        return {"model": self.model.state_dict(), "optimizer": self.opt.state_dict(), "iter_num": self.iter_num}

    def load_state_dict(self, sd):
        if self.graph is not None:
            self.close()
        self.model.load_state_dict(sd["model"])           # copies INTO the flat-arena views
        self.opt.load_state_dict(sd["optimizer"])
        self.iter_num = int(sd["iter_num"])
        ops.invalidate_weight_cache()

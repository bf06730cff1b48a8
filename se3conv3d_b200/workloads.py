"""Synthetic workloads of the named shapes (BASELINE.json configs) built on the public API.

dfaust_conv_stack: the 21 PNEConvLayerRotEquiv calls one forward of FPNSegUNetMLPGeluRotEqFAUST issues
(tasks/SemSeg/seg_models.py:16-36 with models/PatchEncoder.py:95-106, Encoder.py:134-171,
Decoder.py:72-96, FPNDecoder.py:104-132, PatchDecoder.py:62-82, FPNSegUNet.py:166-182), with their
levels, radii (2.0 x level cell size) and channel widths, on the hierarchy create_hierarchy builds
(tasks/SemSeg/train_dfaust_rot.py:108-158; confs/dfaust/dfaust_I_rot_pca_2F.yaml).
"""
import math

import torch

DFAUST_CFG = {
    "init_subsample": 0.04,
    "grid_subsamples": [0.05, 0.1, 0.2, 0.4],
    "RefFrames": {"pca": True, "neigh_method": "knn", "neigh_kwargs": {"neigh_k": 16}, "fixed_axis": False,
                  "n_frames": 2},
}


def synthetic_bodies(n_clouds=32, n_points=6890, seed=0):
    """SMPL-sized synthetic clouds: points on a union of capsules (torso, head, 2 arms, 2 legs) inside a
    0.6 x 0.3 x 1.8 m box, N(0, 0.005) noise.  Returns pts [B*n,3] f32, batch ids i32."""
    g = torch.Generator().manual_seed(seed)
    # (centre a, centre b, radius, weight)
    caps = [((0.0, 0.0, 0.95), (0.0, 0.0, 1.45), 0.13, 0.30), ((0.0, 0.0, 1.62), (0.0, 0.0, 1.70), 0.09, 0.08),
            ((-0.20, 0.0, 1.40), (-0.28, 0.0, 0.85), 0.045, 0.12), ((0.20, 0.0, 1.40), (0.28, 0.0, 0.85), 0.045, 0.12),
            ((-0.09, 0.0, 0.90), (-0.11, 0.0, 0.08), 0.07, 0.19), ((0.09, 0.0, 0.90), (0.11, 0.0, 0.08), 0.07, 0.19)]
    w = torch.tensor([c[3] for c in caps])
    total = n_clouds * n_points
    which = torch.multinomial(w, total, replacement=True, generator=g)
    a = torch.tensor([c[0] for c in caps])[which]
    b = torch.tensor([c[1] for c in caps])[which]
    r = torch.tensor([c[2] for c in caps])[which]
    t = torch.rand(total, generator=g)
    d = torch.randn(total, 3, generator=g)
    axis = (b - a) / (b - a).norm(dim=1, keepdim=True)
    d = d - (d * axis).sum(1, keepdim=True) * axis          # direction orthogonal to the capsule axis
    d = d / d.norm(dim=1, keepdim=True).clamp_min(1e-9)
    pts = a + (b - a) * t[:, None] + d * r[:, None] + 0.005 * torch.randn(total, 3, generator=g)
    # small per-cloud pose jitter so the clouds differ
    ang = (torch.rand(n_clouds, generator=g) - 0.5) * 0.6
    c, s = torch.cos(ang), torch.sin(ang)
    batch = torch.arange(n_clouds).repeat_interleave(n_points)
    x = pts[:, 0] * c[batch] - pts[:, 1] * s[batch]
    y = pts[:, 0] * s[batch] + pts[:, 1] * c[batch]
    pts = torch.stack((x, y, pts[:, 2]), 1).to(torch.float32).contiguous()
    return pts, batch.to(torch.int32)


def uniform_cloud(n, seed=0):
    """BASELINE config 1: torch.rand(n,3) with manual_seed(seed), one batch item."""
    pts = torch.rand(n, 3, generator=torch.Generator().manual_seed(seed))
    return pts, torch.zeros(n, dtype=torch.int32)


# (name, level_in, level_out, radius_level, c_in, c_out); level 5 = the output cloud of the seg head
def dfaust_conv_specs():
    feats, lv = [32, 64, 128, 256], [1, 2, 3, 4]
    specs = [("patch_enc0", 0, 1, 0, 1, 32), ("patch_enc1", 1, 1, 1, 32, 32)]
    for i, (f, l) in enumerate(zip(feats, lv)):
        specs += [("enc%d_block%d" % (i, j), l, l, l, f, f) for j in range(2)]
        if i < 3:
            specs.append(("enc%d_down" % i, l, l + 1, l, f, feats[i + 1]))
    specs += [("dec%d" % i, 4 - i, 3 - i, 4 - i, feats[3 - i], feats[2 - i]) for i in range(3)]
    specs += [("fpn%d" % i, 4 - i, 1, 4 - i, 32, 32) for i in range(3)]
    specs += [("patch_dec0", 1, 0, 1, 32, 32), ("seg_head", 0, 5, 0, 32, 32)]
    assert len(specs) == 21
    return specs


class DfaustStep(object):
    """One 'step' of the dfaust_I_rot_pca_2F hot path: hierarchy + frames + neighbourhoods + the 21
    convolutions forward and backward.  Layers and their inputs are independent (no BN / MLP glue --
    that is SURVEY 8 row f2), each conv gets a fixed random input and output gradient."""

    def __init__(self, device, precision=0, seed=0):
        from .layers import PNEConvLayerRotEquiv
        self.device = device
        self.specs = dfaust_conv_specs()
        torch.manual_seed(seed)
        self.layers = []
        for (_, _, _, _, cin, cout) in self.specs:
            layer = PNEConvLayerRotEquiv(9, cin, cout, 32, "mlp_gelu").to(device)
            layer.precision = precision
            self.layers.append(layer)
        self.inputs = None

    def build_hierarchy(self, pts, batch_ids, fused=True, n_batches=None):
        """Clouds [level 0..4, output cloud] and one neighbourhood per conv.  fused=True: one native call
        (pc.build_point_hierarchy); fused=False: the per-object chain with the reference's API."""
        if fused:
            return self.build_hierarchy_fused(pts, batch_ids, n_batches)
        from .pc import Pointcloud, GridSubSample, PointcloudRotEquiv, PointHierarchyRotEquiv
        cfg = DFAUST_CFG
        with torch.no_grad():
            pc = Pointcloud(pts, batch_ids)
            samp = GridSubSample(pc, cfg["init_subsample"])
            new_pts = samp.__subsample_tensor__(pc.pts_, "avg")
            new_b = samp.__subsample_tensor__(pc.batch_ids_, "max")
            new_pc = PointcloudRotEquiv(new_pts, new_b, cfg["RefFrames"])
            h = PointHierarchyRotEquiv(new_pc, len(cfg["grid_subsamples"]), "grid_avg", grid_radii=cfg["grid_subsamples"])
            # output cloud of the seg head: one random point per 0.04 cell (output_subsample)
            osamp = GridSubSample(pc, cfg["init_subsample"], p_rnd_sample=True)
            out_pc = PointcloudRotEquiv(osamp.__subsample_tensor__(pc.pts_, "avg"),
                                        osamp.__subsample_tensor__(pc.batch_ids_, "max"), cfg["RefFrames"])
            radii = [cfg["init_subsample"]] + cfg["grid_subsamples"]
            pcs = list(h.pcs_) + [out_pc]
            neighs = []
            for (_, li, lo, lr, _, _) in self.specs:
                r = 2.0 * radii[lr]
                if lo == 5:
                    from .pc import BQNeighborhood
                    nb = BQNeighborhood(pcs[li], out_pc, r)
                else:
                    nb = h.create_neighborhood(li, lo, "ball_query", bq_radius=r)
                neighs.append(nb)
        return pcs, neighs

    def build_hierarchy_fused(self, pts, batch_ids, n_batches=None):
        from .pc import build_point_hierarchy
        cfg = DFAUST_CFG
        radii = [cfg["init_subsample"]] + cfg["grid_subsamples"]
        wanted, index = [], {}
        for (_, li, lo, lr, _, _) in self.specs:
            key = (li, lo, 2.0 * radii[lr])
            if key not in index:
                index[key] = len(wanted)
                wanted.append(key)
        with torch.no_grad():
            h, out_pc = build_point_hierarchy(pts, batch_ids, cfg["RefFrames"], cfg["init_subsample"],
                                              cfg["grid_subsamples"], neighborhoods=wanted, output_cloud=True,
                                              n_batches=n_batches)
        pcs = list(h.pcs_) + [out_pc]
        neighs = [h.fused_neighborhoods_[index[(li, lo, 2.0 * radii[lr])]] for (_, li, lo, lr, _, _) in self.specs]
        self.hierarchy = h
        return pcs, neighs

    def calibrate(self, pcs, neighs):
        """What the pre-process epoch converges to: norm_neigh_dist_ = 1/r, norm_num_neighs_ = M/E."""
        for layer, nb in zip(self.layers, neighs):
            layer.norm_neigh_dist_.fill_(1.0 / nb.radius_)
            n_edges = nb._csr_columns[1].shape[0] if getattr(nb, "_csr_columns", None) is not None else nb.neighbors_.shape[0]
            layer.norm_num_neighs_.fill_(nb.start_ids_.shape[0] / max(n_edges, 1))

    def make_inputs(self, pcs, seed=1):
        g = torch.Generator(device="cpu").manual_seed(seed)
        xs, dys = [], []
        for (_, li, lo, _, cin, cout) in self.specs:
            xs.append(torch.randn(pcs[li].pts_.shape[0] * 2, cin, generator=g).to(self.device).requires_grad_(True))
            dys.append(torch.randn(pcs[lo].pts_.shape[0] * 2, cout, generator=g).to(self.device) / math.sqrt(cout))
        self.inputs = (xs, dys)
        return xs, dys

    def conv_fwd_bwd(self, pcs, neighs, xs=None, dys=None, return_output=False):
        """Forward of the 21 convolutions, then ONE backward pass over all of them (as a training step does:
        every forward first, a single autograd sweep after), returns the checksum of the last output."""
        if xs is None:
            xs, dys = self.inputs
        ys = [layer(pcs[li], pcs[lo], x, nb)
              for layer, nb, (_, li, lo, _, _, _), x in zip(self.layers, neighs, self.specs, xs)]
        torch.autograd.backward(ys, list(dys))
        if return_output:
            return ys[-1].detach()    # the seg-head output is the model's result
        return ys[-1].detach().sum()

    def run_pipelined(self, batches, n_batches, side_stream, before_conv=None, after_conv=None, threaded=False):
        """K hot-path steps with the hierarchy of batch i + 1 built on `side_stream` while the convolutions of batch i run
        on the current stream (`threaded`: the build is issued by a worker thread -- measured slower, the launch calls
        of the two threads serialise in the driver).  This is
        what a training loop does with its input pipeline (the hierarchy is a no-grad function of the next batch only,
        tasks/SemSeg/train_dfaust_rot.py:240-259).  Every step's hierarchy is built inside the call; nothing is cached
        across steps.  `batches` yields (pts, batch_ids) already on the device or callables returning them (host ->
        device copies then also run on the side stream).  Returns the last output."""
        import concurrent.futures
        cur = torch.cuda.current_stream()
        dev = self.device

        def build(item):
            torch.cuda.set_device(dev)
            with torch.cuda.stream(side_stream):
                p, b = item() if callable(item) else item
                pcs, neighs = self.build_hierarchy(p, b, n_batches=n_batches)
                ev = torch.cuda.Event()
                ev.record(side_stream)
            return pcs, neighs, self.hierarchy, ev
        pool = concurrent.futures.ThreadPoolExecutor(1) if threaded else None
        submit = (lambda it: pool.submit(build, it)) if threaded else (lambda it: build(it))
        get = (lambda f: f.result()) if threaded else (lambda f: f)
        out = None
        try:
            nxt = submit(batches[0])
            for i in range(len(batches)):
                pcs, neighs, h, ev = get(nxt)
                if threaded and i + 1 < len(batches):
                    nxt = submit(batches[i + 1])        # the worker builds while this thread launches the convolutions
                cur.wait_event(ev)
                h.fused_arena_.record_stream(cur)       # allocated on the side stream, read by the convolutions here
                if before_conv:
                    before_conv()
                out = self.conv_fwd_bwd(pcs, neighs, return_output=True)
                if not threaded and i + 1 < len(batches):
                    nxt = submit(batches[i + 1])
                if after_conv:
                    after_conv(out, i)
                self.zero_grad()
        finally:
            if pool is not None:
                pool.shutdown(wait=True)
        return out

    def zero_grad(self):
        for layer in self.layers:
            for p in layer.parameters():
                p.grad = None
        if self.inputs is not None:
            for x in self.inputs[0]:
                x.grad = None


class FpnStep(object):
    """One full training step of the dfaust_I_rot_pca_2F configuration on synthetic clouds with the UNMODIFIED
    reference model code (models/FPNSegUNet.py through tasks/SemSeg/seg_models.py, imported with `point_cloud_lib`
    aliased to this package; tools/stage_reference_models.py):

        create_hierarchy (train_dfaust_rot.py:108-158, here the fused builder) -> features repeated per frame (:249-251)
        -> pred = model(hierarchy, features, lev_radii, out_pc) (:262) -> cross-entropy with label smoothing 0.2 (:263,
        confs/dfaust/dfaust_I_rot_pca_2F.yaml) -> loss.backward() (:264) -> gradient all-reduce (data parallel replicas;
        shard.GradAllReducer) -> clip_grad_norm_(100) (:267-270) -> AdamW step + zero_grad (:271-272).

    max_drop_path 0.5 as in the config; the pre-process epoch is replaced by its fixed point (norm_neigh_dist_ = 1/r,
    norm_num_neighs_ = M/E) taken from the first hierarchy."""

    def __init__(self, device, seg_models, precision=1, n_classes=20, max_drop_path=0.5, lr=5e-4, seed=0):
        from .layers import PNEConvLayerRotEquiv
        from . import shard
        self.device = device
        torch.manual_seed(seed)
        self.model = seg_models.FPNSegUNetMLPGeluRotEqFAUST(1, n_classes, max_drop_path).to(device)
        self.model.train()
        for m in self.model.modules():
            if isinstance(m, PNEConvLayerRotEquiv):
                m.precision = precision
        self.convs = [m for m in self.model.modules() if isinstance(m, PNEConvLayerRotEquiv)]
        self.reducer = shard.GradAllReducer(self.model.parameters())
        self.optim = torch.optim.AdamW(self.model.parameters(), lr=lr, weight_decay=1e-4, foreach=True)
        self.loss_fn = torch.nn.CrossEntropyLoss(label_smoothing=0.2)
        self.n_classes = n_classes
        self.cfg = DFAUST_CFG
        self.radii = [self.cfg["init_subsample"]] + list(self.cfg["grid_subsamples"])
        # every ball query the model will ask for (Encoder.py:134-154, Decoder.py:72-80, FPNDecoder.py:104-113,
        # PatchEncoder/Decoder, seg head FPNSegUNet.py:171-175): (src level, dst level, 2 x cell of the radius level)
        wanted = []
        for (_, li, lo, lr_, _, _) in dfaust_conv_specs():
            key = (li, lo, 2.0 * self.radii[lr_])
            if key not in wanted:
                wanted.append(key)
        self.wanted = wanted
        self.calibrated = False

    def build(self, pts, batch_ids, n_batches):
        from .pc import build_point_hierarchy
        with torch.no_grad():
            h, out_pc = build_point_hierarchy(pts, batch_ids, self.cfg["RefFrames"], self.cfg["init_subsample"],
                                              self.cfg["grid_subsamples"], neighborhoods=self.wanted, output_cloud=True,
                                              n_batches=n_batches)
        return h, out_pc

    def calibrate(self, h, out_pc, feats):
        """Fixed point of the pre-process EMA (layers/IConvLayer.py:76-97): one no-grad forward with the switch on, then
        the buffers are set to the values the EMA converges to."""
        self.model.start_pre_process()
        with torch.no_grad():
            for _ in range(3):
                self.model(h, feats, self.radii, out_pc)
        self.model.end_pre_process()
        for c in self.convs:
            # after k EMA updates from 0 the buffer holds (1 - 0.9^k) of the fixed point
            scale = 1.0 / (1.0 - 0.9 ** 3)
            c.norm_neigh_dist_ = (c.norm_neigh_dist_ * scale).clone()
            c.norm_num_neighs_ = (c.norm_num_neighs_ * scale).clone()
        self.calibrated = True

    def features(self, h):
        n = h.pcs_[0].pts_.shape[0] * self.cfg["RefFrames"]["n_frames"]
        return torch.ones((n, 1), dtype=torch.float32, device=self.device)

    def step(self, pts, batch_ids, labels_raw, n_batches):
        """hierarchy + forward + loss + backward + all-reduce + clip + AdamW; returns the loss (device scalar)."""
        h, out_pc = self.build(pts, batch_ids, n_batches)
        feats = self.features(h)
        if not self.calibrated:
            self.calibrate(h, out_pc, feats)
        labels = labels_raw[out_pc.picked_ids_]                  # labels of the points the output cloud picked
        pred = self.model(h, feats, self.radii, out_pc)
        loss = self.loss_fn(pred, labels)
        loss.backward()
        self.reducer.finish()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), 100.0, foreach=True)
        self.optim.step()
        self.reducer.zero_grad()
        return loss.detach()

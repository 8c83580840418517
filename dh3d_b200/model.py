"""DH3D forward pass (inference branch of ``DH3D.build_graph``, reference core/model.py:135-206).

    points [B,N,3] -> knn (model.py:157) -> backbone_local_dilate (:99-108) -> l2-normalised local
    descriptors 'xyz_feat' (:177-181) -> detection_block attention 'xyz_feat_att' (:184-188) ->
    compute_global (:112-133): global_before_assemble -> globalatt -> attention NetVLAD ->
    l2-normalised 'globaldesc' (:203-205).

Every tensor op below is a launch of this repo's own CUDA kernels through the C ABI; torch only
allocates buffers and sequences streams.
"""
import torch
from torch import nn

from . import ops
from ._lib import Dh3dError
from .backbones import (DetectionBlock, DilateGeometry, FlexConvDilate, GlobalAttBlock,
                        GlobalNetVLADBlock, LocalBackbone, subsample)
from .configs import DH3DConfig
from .layers import invalidate_folded


class DH3D(nn.Module):
    """``separate_global_backbone=True`` gives the global branch its OWN copy of backbone_local_dilate
    (``global_local.*``): the reference ships two separately trained networks (models/local, models/global)
    whose backbone weights differ, so reproducing both in one pass needs both backbones.  The default (one
    shared backbone) is BASELINE.json configs[2] and what a jointly trained checkpoint would use."""

    def __init__(self, config=None, separate_global_backbone=False):
        super().__init__()
        self.config = config or DH3DConfig()
        c = self.config
        self.local = LocalBackbone(c.init_feat_dim, c.featdim, dilate2=c.dilate, knn=c.knn_num, add_se=c.add_se)
        if c.detection:
            self.detection_block_reliable = DetectionBlock(c.featdim)
        if c.extract_global:
            if c.gl_dims[-1] != 256 or c.cluster_size != 64 or c.output_dim != 256:
                raise Dh3dError("DH3D: the NetVLAD kernels are built for gl_dims[-1] == 256, cluster_size == 64, "
                                "output_dim == 256 (the shipped global_config); got gl_dims=%s cluster_size=%d "
                                "output_dim=%d" % (list(c.gl_dims), c.cluster_size, c.output_dim))
            if separate_global_backbone:
                self.global_local = LocalBackbone(c.init_feat_dim, c.featdim, dilate2=c.dilate, knn=c.knn_num,
                                                  add_se=c.add_se)
            self.global_before_assemble = FlexConvDilate(c.featdim, c.gl_dims, dilate=c.gl_dilate,
                                                         knn=c.knn_num, concat=False, add_se="")
            self.globalatt = GlobalAttBlock(c.gl_dims[-1])
            self.netvlad = GlobalNetVLADBlock(c.gl_dims[-1], c.cluster_size, c.output_dim)
        self._side = None

    def invalidate(self):
        invalidate_folded(self)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._side = None   # the side stream belongs to the device the parameters were on
        return out

    @torch.no_grad()
    def forward(self, points, knn_inds=None, outputs=("local_desc", "attention", "globaldesc"),
                overlap=True, out=None):
        """points [B,N,3] fp32 CUDA; knn_inds optional [B,N,K] i32 (the reference's 'knn_inds'
        input for N > 8192, model.py:148-155).  Returns a dict with the requested tensors:
          feat [B,N,128] raw, local_desc [B,N,128] l2-normed, attention [B,N,1], globaldesc [B,256],
          xyz_feat [B,N,131], xyz_feat_att [B,N,132] (the reference's saved tensor names).
        ``out``: optional dict of caller-owned result tensors for 'local_desc' [B,N,featdim], 'attention' [B,N,1],
        'globaldesc' [B,256] (e.g. views of one contiguous block, so a step's results leave in ONE D2H copy)."""
        c = self.config
        points = points.contiguous()
        same_geometry = c.extract_global and c.gl_dilate == c.dilate
        cur = torch.cuda.current_stream(points.device)

        # xyz-only geometry (FPS -> k-NN of the sampled points -> 3-NN) on a side stream next to stage 1.  The dense cloud
        # is cell-sorted ONCE (knn_sort, first thing on the side stream): the main stream's k-NN query, the box-pruned FPS
        # and the 3-NN all walk that copy.
        geometry, sorted_xyz = None, None
        own_knn = knn_inds is None
        if overlap:
            if self._side is None or self._side.device != points.device:
                self._side = torch.cuda.Stream(device=points.device)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                if own_knn:
                    sorted_xyz = ops.knn_sort(points)
                    sorted_ready = torch.cuda.Event()
                    sorted_ready.record(self._side)
                geometry = DilateGeometry(points, points.shape[1] // c.dilate, c.knn_num, sorted_xyz)
                # the main stream joins where the geometry is first consumed (stage 2's group_point), so the
                # k-NN of the sampled points and the 3-NN run next to stage 1 instead of in front of it
                geometry.ready = torch.cuda.Event()
                geometry.ready.record(self._side)
            if own_knn:
                cur.wait_event(sorted_ready)
                knn_inds, _ = ops.knn_query_sorted(sorted_xyz, c.knn_num)
            if not torch.cuda.is_current_stream_capturing():
                for t in (geometry.kp_indices, geometry.points_sampled, geometry.knn_indices,
                          geometry.nn_dist, geometry.nn_idx) + ((sorted_xyz,) if sorted_xyz is not None else ()):
                    t.record_stream(cur)
        else:
            if own_knn:
                sorted_xyz = ops.knn_sort(points)
                knn_inds, _ = ops.knn_query_sorted(sorted_xyz, c.knn_num)
            geometry = DilateGeometry(points, points.shape[1] // c.dilate, c.knn_num, sorted_xyz)

        want = set(outputs)
        given = out or {}
        out = {}
        if want & {"local_desc", "xyz_feat", "xyz_feat_att"}:
            feat, out["local_desc"] = self.local(points, knn_inds, geometry=geometry, with_desc=True,
                                                 desc_out=given.get("local_desc"))
        else:
            feat = self.local(points, knn_inds, geometry=geometry)
        out["feat"] = feat
        if c.detection and want & {"attention", "xyz_feat_att"}:
            out["attention"] = self.detection_block_reliable(feat, out=given.get("attention"))
        if c.extract_global and "globaldesc" in want:
            # (running the detector head on the side stream next to the global branch was measured: no gain,
            #  3.18 vs 3.17 ms per step -- every large kernel here is a persistent one-CTA-per-SM grid)
            g = geometry if same_geometry else None
            gfeat = feat
            if getattr(self, "global_local", None) is not None:
                gfeat = self.global_local(points, knn_inds, geometry=geometry)
            forglobal = self.global_before_assemble(points, gfeat, geometry=g)
            gpoints = points
            if c.global_subsample > 0:   # core/model.py:118-121
                gpoints, forglobal, _ = subsample(points, forglobal, c.global_subsample)
            att = self.globalatt(forglobal)
            out["globaldesc"] = self.netvlad(gpoints, forglobal, att, final_l2norm=True, out=given.get("globaldesc"))
        if overlap:
            geometry.join()   # (already joined by stage 2; keeps the side stream joined for configurations that skip it)
        if "xyz_feat" in want:
            out["xyz_feat"] = torch.cat([points, out["local_desc"]], dim=-1)
        if "xyz_feat_att" in want and c.detection:
            out["xyz_feat_att"] = torch.cat([points, out["local_desc"], out["attention"]], dim=-1)
        return out


@torch.no_grad()
def extract_keypoints(points, local_desc, attention, nms_radius=0.5, min_response_ratio=0.01, max_keypoints=512,
                      out=None):
    """The reference's ``--perform_nms`` output mode (evaluate/local_eval/localdesc_extract.py:92-104) for a whole
    batch on the device: keypoint NMS on ``1 - attention`` (``single_nms``, core/utils.py:15-43), then only the
    detected rows of 'xyz_feat_att' are kept.

    points [B,N,3], local_desc [B,N,C], attention [B,N,1] -> (rows [B,max_keypoints,3+C+1] with cloud b's
    keypoints in rows [0, counts[b]) in the reference's order (descending response) and zeros after, counts [B] i32).
    ``out``: caller-owned [B,max_keypoints,3+C+1] tensor."""
    from .utils import batched_nms
    B, N, C = local_desc.shape
    response = ops.affine(attention.reshape(B, N).contiguous(), -1.0, 1.0)          # 1 - attention (:95)
    idx, counts = batched_nms(points, response, nms_radius, min_response_ratio, max_keypoints)
    rows = out if out is not None else torch.empty((B, max_keypoints, 3 + C + 1), dtype=points.dtype,
                                                   device=points.device)
    ops.gather_rows_cols(points, idx, rows, 0)          # rows with idx < 0 (beyond counts[b]) are zero-filled
    ops.gather_rows_cols(local_desc, idx, rows, 3)
    ops.gather_rows_cols(attention, idx, rows, 3 + C)
    return rows, counts


def result_block(model, batch, n_points, outputs, device):
    """One contiguous fp32 block holding a step's results back to back, and the views ``forward(out=...)``
    writes through: everything a step returns leaves the GPU in ONE device-to-host copy."""
    c = model.config
    shapes = {"local_desc": (batch, n_points, c.featdim), "attention": (batch, n_points, 1),
              "globaldesc": (batch, c.output_dim)}
    names = [k for k in ("local_desc", "attention", "globaldesc") if k in outputs and
             (k != "attention" or c.detection) and (k != "globaldesc" or c.extract_global)]
    sizes = [int(torch.Size(shapes[k]).numel()) for k in names]
    padded = [(n + 3) // 4 * 4 for n in sizes]           # keep every view 16-byte aligned
    block = torch.empty(sum(padded), dtype=torch.float32, device=device)
    views, off = {}, 0
    for k, n, pn in zip(names, sizes, padded):
        views[k] = block[off:off + n].view(shapes[k])
        off += pn
    return block, views


class GraphedForward(object):
    """The forward pass captured once into a CUDA graph (all ~32 kernel launches, both streams) and
    replayed per batch: removes the host launch cost and the inter-kernel gaps.  Shapes are frozen to
    the example batch; outputs are static tensors that the next replay overwrites.  The requested results
    are views of ONE contiguous block (``self.block``), see result_block()."""

    def __init__(self, model, example_points, outputs=("local_desc", "attention", "globaldesc"),
                 overlap=True, warmup=3):
        self.model, self.outputs = model, tuple(outputs)
        self.static_in = example_points.detach().clone().contiguous()
        B, N, _ = self.static_in.shape
        self.block, self.views = result_block(model, B, N, self.outputs, self.static_in.device)
        side = torch.cuda.Stream(device=self.static_in.device)
        side.wait_stream(torch.cuda.current_stream(self.static_in.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):   # folds BN, sets kernel attributes, warms the allocator
                model(self.static_in, outputs=self.outputs, overlap=overlap, out=self.views)
        torch.cuda.current_stream(self.static_in.device).wait_stream(side)
        torch.cuda.synchronize(self.static_in.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = model(self.static_in, outputs=self.outputs, overlap=overlap, out=self.views)

    def __call__(self, points):
        self.static_in.copy_(points, non_blocking=True)
        self.graph.replay()
        return self.static_out


class InFlightForward(object):
    """Several batches in flight on one device: each GraphedForward instance ("lane", its own static buffers) replays on
    its own stream, batches go to the lanes round robin.  The forward is one long dependency chain whose kernels are bound
    by different things (issue / latency: k-NN, FPS, gathers; tensor pipe: the heads; HBM: join, interpolation), so the
    next batch's kernels fill what the current one leaves idle: 2.09 -> 1.94 ms per 32-cloud batch with two lanes on a
    B200, a third lane loses again (profiles/inflight_r3s.txt).  Results are bit-identical to serial replays.

        fwd = InFlightForward.build(model, example_points)            # two lanes
        for pts in batches:
            out, done = fwd.submit(pts, consume=lambda o: keep.append(o["globaldesc"].clone()))
        fwd.join()                                                     # current stream now waits for every lane

    ``consume(out)`` runs inside the lane's stream context right after the replay (copy results out of the static buffers
    there: lane k's next batch overwrites them); making the CURRENT stream wait on ``done`` between submits works too but
    serialises the lanes."""

    def __init__(self, graphs):
        self.graphs = list(graphs)
        dev = self.graphs[0].static_in.device
        self.device = dev
        self.streams = [torch.cuda.Stream(device=dev) for _ in self.graphs]
        self.done = [None] * len(self.graphs)
        self._next = 0

    @classmethod
    def build(cls, model, example_points, outputs=("local_desc", "attention", "globaldesc"), lanes=2):
        return cls([GraphedForward(model, example_points, outputs=outputs) for _ in range(lanes)])

    def submit(self, points, consume=None):
        """Enqueue one batch on the next lane (after whatever the current stream has queued so far, e.g. the H2D copy of
        ``points``).  -> (static output dict of that lane, event recorded behind the replay and ``consume``)."""
        k = self._next
        self._next = (k + 1) % len(self.graphs)
        lane = self.streams[k]
        lane.wait_stream(torch.cuda.current_stream(self.device))
        points.record_stream(lane)
        with torch.cuda.stream(lane):
            out = self.graphs[k](points)
            if consume is not None:
                consume(out)
            ev = torch.cuda.Event()
            ev.record(lane)
        self.done[k] = ev
        return out, ev

    def join(self):
        cur = torch.cuda.current_stream(self.device)
        for lane in self.streams:
            cur.wait_stream(lane)


def init_random_(model, seed=0, knn=8, offset_scale=2.0):
    """Random-init weights of the right shapes (benchmarks have no network for checkpoints), scaled
    so activations stay O(1) through the stack for clouds whose neighbour offsets are about
    ``offset_scale`` metres: FlexConv weights w = bias + theta.delta get var 1/(K*Din), 1x1 kernels
    N(0, 1/fan_in), BN statistics near identity.  Deterministic per seed."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            shape = tuple(p.shape)
            leaf = name.split(".")[-1]
            if "se_avgpool" in name:    # Flex_Avg: zero, non-trainable theta (core/layers.py:380-385)
                p.zero_()
                continue
            r = lambda: torch.randn(shape, generator=g)
            if leaf == "gamma":
                v = 1.0 + 0.1 * r()
            elif leaf == "variance_ema":
                v = 1.0 + 0.2 * torch.rand(shape, generator=g)
            elif leaf in ("beta", "mean_ema", "b", "feature_bias"):
                v = 0.1 * r()
            elif leaf == "position_theta" and len(shape) == 3:      # FlexConv [3,Din,Dout]
                v = r() * (0.5 / (knn * shape[1])) ** 0.5 / offset_scale
            elif leaf == "position_bias" and len(shape) == 2:       # FlexConv [Din,Dout]
                v = r() * (0.5 / (knn * shape[0])) ** 0.5
            elif leaf == "position_theta":                          # ConvPointset [Din,Dout]
                v = r() * (1.0 / (knn * shape[0])) ** 0.5 / offset_scale
            elif leaf == "position_bias":                           # ConvPointset [Dout]
                v = 0.1 * r()
            else:
                fan_in = shape[-2] if len(shape) >= 2 else shape[0]
                v = r() / (float(fan_in) ** 0.5)
            p.copy_(v.to(p.dtype))
    invalidate_folded(model)
    return model

import ctypes as C

import torch

from .. import point_cloud_lib_ops as ops
from .._lib import lib, check, ptr, stream, workspace, ConvDesc, Se3Error, grid_setup

ACT_CODES = {"mlp_linear": 0, "mlp_relu": 1, "mlp_gelu": 2, "mlp_sin": 3}


class FeatBasisProj(torch.autograd.Function):
    """T[m,c,k] = sum_e feats[src(e),c] * basis[e,k]   (custom_ops/FeatBasisProj.py:10-66)."""

    @staticmethod
    def forward(ctx, p_pt_basis, p_pt_features, p_neighbors, p_start_ids):
        ctx.save_for_backward(p_pt_basis, p_pt_features, p_neighbors, p_start_ids)
        return ops.feat_basis_proj(p_pt_basis, p_pt_features, p_neighbors, p_start_ids)

    @staticmethod
    def backward(ctx, p_grads):
        basis, feats, neighbors, start_ids = ctx.saved_tensors
        feat_grads, basis_grads = ops.feat_basis_proj_grad(basis, feats, neighbors, start_ids, p_grads.contiguous())
        return basis_grads.to(basis.dtype), feat_grads.to(feats.dtype), None, None


class BallQuery(torch.autograd.Function):
    """Radius neighbours of every sample among same-batch sources (custom_ops/BallQuery.py:11-54).

    The grid set-up reproduces the reference wrapper exactly, including its quirk of subtracting
    1e-6 from the per-batch maximum as well as the minimum (custom_ops/BallQuery.py:36-37)."""

    @staticmethod
    def forward(ctx, p_pt_src, p_pt_sample, p_batch_id_src, p_batch_id_sample, radius, max_neighbors, n_batches=None):
        if n_batches is None:
            n_batches = int(p_batch_id_src.max()) + 1
        # min - 1e-6, max - 1e-6, num_cells = max_b(int((max - min)/radius) + 1): one fused native call
        min_pt, _, num_cells = grid_setup(p_pt_src, p_batch_id_src, n_batches, radius, -1e-6)
        radius_tensor = torch.full((p_pt_src.shape[1],), float(radius), dtype=torch.float32, device=p_pt_src.device)
        neighbors, start_ids = ops.ball_query(p_pt_src, p_pt_sample, p_batch_id_src, p_batch_id_sample, min_pt,
                                              num_cells, radius_tensor, max_neighbors)
        ctx.mark_non_differentiable(neighbors, start_ids)
        return neighbors, start_ids

    @staticmethod
    def backward(ctx, *grads):
        return None, None, None, None, None, None, None


class KNNQuery(torch.autograd.Function):
    """k nearest neighbours inside a batch, self included (custom_ops/KNNQuery.py:11-35)."""

    @staticmethod
    def forward(ctx, p_pt_src, p_batch_id_src, p_k):
        out = ops.knn_query(p_pt_src, p_batch_id_src, p_k)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, *grads):
        return None, None, None


class ComputeKeys(torch.autograd.Function):
    """Voxel key per point (custom_ops/ComputeKeys.py:11-40)."""

    @staticmethod
    def forward(ctx, p_pts, p_batch_ids, p_aabb_min, p_grid_size, p_cell_size):
        out = ops.compute_keys(p_pts, p_batch_ids, p_aabb_min, p_grid_size, p_cell_size)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, *grads):
        return None, None, None, None, None


def make_conv_desc(geom, c_in, c_out, k, act, precision, norm_neigh_dist, out_scale, proj_axes, proj_biases,
                   conv_weights):
    """`se3_conv_desc` of one conv call.  The geometry half (sizes, CSR, records) and the byte counts that
    depend on it are cached on the neighbourhood geometry record per (channels, activation, precision);
    a call only refreshes the parameter pointers and the two scalars."""
    key = (int(c_in), int(c_out), int(k), int(act), int(precision))
    cache = geom.__dict__.setdefault("_desc_cache", {})
    entry = cache.get(key)
    if entry is None:
        d = ConvDesc()
        d.n_in, d.n_out, d.n_edges = geom.n_in, geom.n_out, geom.n_edges
        d.f_in, d.f_out, d.c_in, d.c_out, d.k = geom.f_in, geom.f_out, key[0], key[1], key[2]
        d.act, d.precision, d.reserved = key[3], key[4], 0
        addr = geom.__dict__.get("_addr")
        if addr is not None:      # geometry from the fused hierarchy builder: device addresses, no tensor objects
            for name, a in addr.items():
                setattr(d, name, a)
        else:
            d.pts_in, d.pts_out = ptr(geom.pts_in), ptr(geom.pts_out)
            d.frames_in, d.frames_out = ptr(geom.frames_in), ptr(geom.frames_out)
            d.row_ends, d.col_src = ptr(geom.row_ends), ptr(geom.col_src)
            d.t_row_ends, d.t_edge, d.t_dst = ptr(geom.t_row_ends), ptr(geom.t_edge), ptr(geom.t_dst)
            d.rec_in, d.rec_out = ptr(geom.rec_in), ptr(geom.rec_out)
        L = lib()
        ref = C.byref(d)
        entry = (d, ref, int(L.se3_conv_saved_bytes(ref)), int(L.se3_conv_fwd_workspace_bytes(ref)),
                 int(L.se3_conv_bwd_workspace_bytes(ref)), int(L.se3_conv_weight_cache_bytes(ref)))
        cache[key] = entry
    d = entry[0]
    d.norm_neigh_dist, d.out_scale = float(norm_neigh_dist), float(out_scale)
    d.proj_axes, d.proj_biases, d.conv_weights = ptr(proj_axes), ptr(proj_biases), ptr(conv_weights)
    d.weight_cache, d.weight_cache_state = None, 0
    return entry


class WeightLayoutCache(object):
    """Per-layer device buffer for the bf16 operand layouts of conv_weights_ (precision 1).  They depend on the weights
    only: the forward rebuilds them when the parameter tensor was replaced or written (optimiser step, load_state_dict),
    and skips the conversion otherwise (evaluation, gradient accumulation, several calls per step)."""

    def __init__(self):
        self.key, self.buf, self.param = None, None, None

    def attach(self, d, cw_param, cw32, nbytes):
        """Sets the descriptor's cache fields; returns nothing.  `cw_param` is the nn.Parameter (identity + version are
        the key; the entry holds it so that its id cannot be recycled), `cw32` the float32 tensor the kernels read."""
        if nbytes <= 0:
            return
        key = (id(cw_param), cw_param._version, cw32.data_ptr(), nbytes, str(cw32.device))
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != cw32.device:
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=cw32.device)
            self.key = None
        d.weight_cache = self.buf.data_ptr()
        d.weight_cache_state = 2 if self.key == key else 1
        self.key, self.param = key, cw_param


def _f32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()


class RotEquivConv(torch.autograd.Function):
    """Fused forward/backward of PNEConvLayerRotEquiv.__compute_convolution__
    (layers/PNEConvLayerRotEquiv.py:160-216) through se3_conv_fwd / se3_conv_bwd."""

    @staticmethod
    def forward(ctx, x, proj_axes, proj_biases, conv_weights, geom, act, precision, norm_neigh_dist, out_scale, wcache=None):
        if not x.is_cuda:
            raise Se3Error("RotEquivConv needs CUDA tensors; there is no CPU fallback")
        x32 = _f32c(x)
        pa, pb, cw = _f32c(proj_axes.detach()), _f32c(proj_biases.detach()), _f32c(conv_weights.detach())
        c_in, k, c_out = cw.shape
        if x32.shape[0] != geom.n_in * geom.f_in or x32.shape[1] != c_in:
            raise Se3Error("RotEquivConv: features must be [N*F_in, C_in] = [%d, %d], got %s" %
                           (geom.n_in * geom.f_in, c_in, tuple(x32.shape)))
        if pa.shape[0] != 9:
            raise Se3Error("RotEquivConv: proj_axes_ must be [9, K] (p_dims=9, '6D' relative rotation)")
        L = lib()
        d, dref, saved_bytes, fwd_ws, _, wc_bytes = make_conv_desc(geom, c_in, c_out, k, act, precision, norm_neigh_dist,
                                                                  out_scale, pa, pb, cw)
        if wcache is not None and precision == 1:
            wcache.attach(d, conv_weights, cw, wc_bytes)
        y = torch.empty((geom.n_out * geom.f_out, c_out), dtype=torch.float32, device=x.device)
        saved = torch.empty(max(saved_bytes, 256), dtype=torch.uint8, device=x.device)
        ws = workspace(fwd_ws, x.device, 'conv')
        check(L.se3_conv_fwd(dref, x32.data_ptr(), y.data_ptr(), saved.data_ptr(), ws.data_ptr(), ws.numel(), stream()),
              "se3_conv_fwd")
        ctx.geom = geom
        ctx.wcache = (d.weight_cache, wcache.buf) if (wcache is not None and d.weight_cache) else None
        ctx.meta = (act, precision, float(norm_neigh_dist), float(out_scale), x.dtype)
        ctx.save_for_backward(x32, pa, pb, cw, saved)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x32, pa, pb, cw, saved = ctx.saved_tensors
        act, precision, nnd, osc, x_dtype = ctx.meta
        geom = ctx.geom
        c_in, k, c_out = cw.shape
        dy = _f32c(dy)
        L = lib()
        d, dref, _, _, bwd_ws, _ = make_conv_desc(geom, c_in, c_out, k, act, precision, nnd, osc, pa, pb, cw)
        if ctx.wcache is not None:       # the layouts the forward of this call used
            d.weight_cache, d.weight_cache_state = ctx.wcache[0], 2
        need = ctx.needs_input_grad
        dx = torch.empty_like(x32) if need[0] else None
        dA = torch.empty_like(pa) if (need[1] or need[2]) else None
        dB = torch.empty_like(pb) if (need[1] or need[2]) else None
        dW = torch.empty_like(cw) if need[3] else None
        ws = workspace(bwd_ws, x32.device, 'conv')
        check(L.se3_conv_bwd(dref, x32.data_ptr(), dy.data_ptr(), saved.data_ptr(), ptr(dx), ptr(dW), ptr(dA), ptr(dB),
                             ws.data_ptr(), ws.numel(), stream()), "se3_conv_bwd")
        return (dx.to(x_dtype) if dx is not None else None, dA if need[1] else None, dB if need[2] else None, dW,
                None, None, None, None, None, None)


def _point_items(p_pc):
    """int32 batch item of every POINT of a cloud (cached on the cloud)."""
    t = p_pc.__dict__.get("_point_items_i32") if hasattr(p_pc, "__dict__") else None
    if t is None or t.shape[0] != p_pc.batch_ids_.shape[0]:
        t = p_pc.batch_ids_.to(torch.int32).contiguous()
        try:
            p_pc._point_items_i32 = t
        except AttributeError:
            pass
    return t


class GammaSkip(torch.autograd.Function):
    """out = drop_path(x * gamma) + y in one kernel each way (layers/SkipConnection.py:31-43, DropPathPC.py:23-50)."""

    @staticmethod
    def forward(ctx, x, y, gamma, item_scale, point_item, frames):
        x, y = _f32c(x), _f32c(y)
        g = _f32c(gamma).reshape(-1)
        rows, c = x.shape
        out = torch.empty_like(x)
        check(lib().se3_gamma_skip_fwd(ptr(x), ptr(y), ptr(g), ptr(item_scale), ptr(point_item), int(frames), rows, c,
                                       ptr(out), stream()), "se3_gamma_skip_fwd")
        ctx.save_for_backward(x, g, item_scale, point_item)
        ctx.frames, ctx.gshape = int(frames), gamma.shape
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, g, item_scale, point_item = ctx.saved_tensors
        dy = _f32c(dy)
        rows, c = x.shape
        L = lib()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dg = torch.empty(c, dtype=torch.float32, device=x.device)
        ws = workspace(L.se3_gamma_skip_workspace_bytes(rows, c), x.device, "block")
        check(L.se3_gamma_skip_bwd(ptr(dy), ptr(x), ptr(g), ptr(item_scale), ptr(point_item), ctx.frames, rows, c, ptr(dx),
                                   ptr(dg), ptr(ws), ws.numel(), stream()), "se3_gamma_skip_bwd")
        return dx, (dy if ctx.needs_input_grad[1] else None), dg.reshape(ctx.gshape), None, None, None


POOL_MODES = {"avg": 0, "sum": 1, "max": 2, "min": 3}


class FramePool(torch.autograd.Function):
    """Pooling over the F per-frame rows of every point (pc/PointcloudRotEquiv.py:224-251)."""

    @staticmethod
    def forward(ctx, x, frames, mode):
        x = _f32c(x)
        n, c = x.shape[0] // int(frames), x.shape[1]
        out = torch.empty((n, c), dtype=torch.float32, device=x.device)
        check(lib().se3_frame_pool_fwd(ptr(x), n, int(frames), c, int(mode), ptr(out), stream()), "se3_frame_pool_fwd")
        ctx.meta = (n, int(frames), c, int(mode))
        if mode >= 2:
            ctx.save_for_backward(x, out)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        n, f, c, mode = ctx.meta
        x, out = ctx.saved_tensors if mode >= 2 else (None, None)
        dout = _f32c(dout)
        dx = torch.empty((n * f, c), dtype=torch.float32, device=dout.device)
        check(lib().se3_frame_pool_bwd(ptr(dout), ptr(x), ptr(out), n, f, c, mode, ptr(dx), stream()), "se3_frame_pool_bwd")
        return dx, None, None


class BatchPool(torch.autograd.Function):
    """avg / sum pooling of the rows of every batch item (global_pooling*, pc/PointcloudRotEquiv.py:195-222, 253-275);
    rows are grouped by item (item_ends inclusive, row_item per row)."""

    @staticmethod
    def forward(ctx, x, item_ends, row_item, mode):
        x = _f32c(x)
        b, c = item_ends.shape[0], x.shape[1]
        out = torch.empty((b, c), dtype=torch.float32, device=x.device)
        check(lib().se3_batch_pool_fwd(ptr(x), ptr(item_ends), b, c, int(mode), ptr(out), stream()), "se3_batch_pool_fwd")
        ctx.save_for_backward(item_ends, row_item)
        ctx.meta = (x.shape[0], c, int(mode))
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        item_ends, row_item = ctx.saved_tensors
        rows, c, mode = ctx.meta
        dout = _f32c(dout)
        dx = torch.empty((rows, c), dtype=torch.float32, device=dout.device)
        check(lib().se3_batch_pool_bwd(ptr(dout), ptr(item_ends), ptr(row_item), rows, c, mode, ptr(dx), stream()),
              "se3_batch_pool_bwd")
        return dx, None, None, None

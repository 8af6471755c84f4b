"""Sparse-tensor operators on top of the C ABI: rulebooks, sparse convolution, fused BatchNorm+ReLU,
dense()/scatter-to-BEV.  These are the primitives the `spconv` shim and the mirrored
`unidistill.layers` modules are made of (reference: layers/blocks_3d/det3d/spconv_backbone.py).
"""
import os
import weakref

import torch
from torch.autograd import Function

from . import native

_WS_RULEBOOK = 1  # workspace slot that must survive between conv_count and conv_fill


def _triple(v):
    if isinstance(v, (list, tuple)):
        assert len(v) == 3
        return [int(x) for x in v]
    return [int(v)] * 3


class Rulebook:
    """Kernel map of one sparse convolution (see csrc/rulebook.cu for the conventions)."""

    def __init__(self, subm, ksize, stride, padding, dilation, in_shape, out_shape, n_in, n_out,
                 nbr_out, nbr_in, out_coords, batch_size):
        self.subm = subm
        self.ksize, self.stride, self.padding, self.dilation = ksize, stride, padding, dilation
        self.in_shape, self.out_shape = list(in_shape), list(out_shape)
        self.n_in, self.n_out = n_in, n_out
        self.nbr_out, self.nbr_in = nbr_out, nbr_in
        self.out_coords = out_coords
        self.batch_size = batch_size
        self.K = ksize[0] * ksize[1] * ksize[2]
        self._pairs = None
        self._tap_major = {}
        self._windows = {}
        self.in_coords_for_inverse = None

    def tap_major(self, table):
        """((K, ld) transposed copy with ld = n rounded up to 128, (ceil(n/128),) per-tile tap masks) of one of this rulebook's (n, K) tables,
        built once and cached (the tensor-core kernel reads the 128 indices of a (tile, tap) as one contiguous
        segment and schedules only the taps that occur in a tile)."""
        key = table.data_ptr()
        t = self._tap_major.get(key)
        if t is None:
            n, K = int(table.shape[0]), int(table.shape[1])
            ld = (n + 127) // 128 * 128            # padded with -1: whole 512-byte index segments per (tile, tap)
            tt = torch.empty(((K + 3) // 4 * 4, ld), dtype=torch.int32, device=table.device)
            masks = torch.empty(((n + 127) // 128,), dtype=torch.int32, device=table.device)
            rc = native.load().ud_rulebook_transpose(native.ptr(table), n, K, native.ptr(tt), native.ptr(masks),
                                                     native.stream_ptr(table.device))
            native.check(rc, "ud_rulebook_transpose")
            t = (tt, masks)
            self._tap_major[key] = t
        return t

    def windows(self, table):
        """(ceil(n/128), 9) int32: first row of the window of every (128-row tile, (kz, ky) tap group) of a 3x3x3 table
        (what the window-staged convolution kernel stages by TMA); built once per table and cached."""
        key = table.data_ptr()
        w = self._windows.get(key)
        if w is None:
            n, K = int(table.shape[0]), int(table.shape[1])
            table_t, _ = self.tap_major(table)
            w = torch.empty(((n + 127) // 128, 9), dtype=torch.int32, device=table.device)
            rc = native.load().ud_rulebook_windows(native.ptr(table_t), n, K, native.ptr(w), native.stream_ptr(table.device))
            native.check(rc, "ud_rulebook_windows")
            self._windows[key] = w
        return w

    @classmethod
    def for_table(cls, table):
        """Minimal holder for a bare (n, K) gather table (functional API without a Rulebook)."""
        rb = cls.__new__(cls)
        rb._tap_major = {}
        rb._windows = {}
        return rb

    def num_pairs(self):
        """Number of (input,output) pairs == rulebook size P (host sync; for reporting only)."""
        return int((self.nbr_out >= 0).sum().item())

    def spconv_pairs(self):
        """spconv-shaped (indice_pairs (2,K,N_in) int32, indice_pair_num (K,) int32)."""
        if self._pairs is None:
            lib = native.load()
            table = self.nbr_out if self.subm else self.nbr_in
            dev = table.device
            pairs = torch.empty((2, self.K, self.n_in), dtype=torch.int32, device=dev)
            pair_num = torch.empty((self.K,), dtype=torch.int32, device=dev)
            nbytes = native.query_bytes("ud_rulebook_pairs_workspace_bytes", self.n_in, self.K)
            ws = native.WORKSPACE.get(nbytes, dev)
            rc = lib.ud_rulebook_pairs(native.ptr(table), self.n_in, self.K, 1 if self.subm else 0,
                                       native.ptr(pairs), native.ptr(pair_num), native.ptr(ws), ws.numel(),
                                       native.stream_ptr(dev))
            native.check(rc, "ud_rulebook_pairs")
            self._pairs = (pairs, pair_num)
        return self._pairs


def _tap_major_buffers(n, K, dev):
    """(K4, ld) int32 tap-major table + (ld/128,) tile masks for an (n, K) gather table (see Rulebook.tap_major)."""
    ld = (n + 127) // 128 * 128
    return (torch.empty(((K + 3) // 4 * 4, ld), dtype=torch.int32, device=dev),
            torch.empty((max(ld // 128, 1),), dtype=torch.int32, device=dev))


def conv_out_shape(in_shape, ksize, stride, padding, dilation):
    return [
        (in_shape[d] + 2 * padding[d] - dilation[d] * (ksize[d] - 1) - 1) // stride[d] + 1
        for d in range(3)
    ]


def build_subm_rulebook(coords, batch_size, spatial_shape, ksize, dilation=1):
    native.require_cuda(coords)
    lib = native.load()
    ksize, dilation = _triple(ksize), _triple(dilation)
    assert coords.dtype == torch.int32 and coords.is_contiguous() and coords.shape[1] == 4
    n = int(coords.shape[0])
    K = ksize[0] * ksize[1] * ksize[2]
    dev = coords.device
    nbr = torch.empty((n, K), dtype=torch.int32, device=dev)
    nbr_t, masks = _tap_major_buffers(n, K, dev)          # the builder writes both table forms in one pass
    nbytes = native.query_bytes("ud_rulebook_subm_workspace_bytes", n)
    ws = native.WORKSPACE.get(nbytes, dev)
    shp, ks, dl = native.host_ints(spatial_shape), native.host_ints(ksize), native.host_ints(dilation)
    if native.TIMER.enabled:
        native.TIMER.meta = {"K": K}
    rc = lib.ud_rulebook_subm(native.ptr(coords), n, int(batch_size), native.host_ptr(shp), native.host_ptr(ks),
                              native.host_ptr(dl), native.ptr(nbr), native.ptr(nbr_t), native.ptr(masks), native.ptr(ws),
                              ws.numel(), native.stream_ptr(dev))
    native.check(rc, "ud_rulebook_subm")
    rb = Rulebook(True, ksize, [1, 1, 1], [k // 2 for k in ksize], dilation, spatial_shape, spatial_shape,
                  n, n, nbr, None, coords, batch_size)
    if n > 0:
        rb._tap_major[nbr.data_ptr()] = (nbr_t, masks)
    return rb


# Row order of strided-convolution outputs: "linear" (default: canonical ascending (b,z,y,x), what spconv's CPU rulebook
# produces and the bit-exact rulebook tests pin) or "morton" (Z-order: compact 3-D tiles).  Measured on B200 at the bench
# size (profiles/r02_morton_vs_linear.md): Morton tiles touch fewer distinct rows, but under the raster order the rows a tap
# gathers are CONSECUTIVE in memory (x-adjacent voxels are adjacent rows), which the gather likes better: 64->64 6.82 vs
# 6.69 ms, 128->128 5.47 vs 4.97 ms per step, whole step 93.4 vs 89.5 ms.  So the option stays off.
CONV_OUTPUT_ORDER = os.environ.get("UD_CONV_ORDER", "linear")


def morton_order(coords, batch_size, spatial_shape):
    """(n,) int64 permutation sorting rows by (b, Morton(z, y, x)); stable."""
    native.require_cuda(coords)
    lib = native.load()
    n = int(coords.shape[0])
    dev = coords.device
    order = torch.empty((n,), dtype=torch.int32, device=dev)
    ws = native.WORKSPACE.get(native.query_bytes("ud_morton_order_workspace_bytes", n), dev)
    shp = native.host_ints(spatial_shape)
    rc = lib.ud_morton_order(native.ptr(coords), n, int(batch_size), native.host_ptr(shp), native.ptr(order),
                             native.ptr(ws), ws.numel(), native.stream_ptr(dev))
    native.check(rc, "ud_morton_order")
    return order.long()


def build_conv_rulebook(coords, batch_size, spatial_shape, ksize, stride, padding, dilation=1, order=None):
    native.require_cuda(coords)
    lib = native.load()
    ksize, stride, padding, dilation = _triple(ksize), _triple(stride), _triple(padding), _triple(dilation)
    assert coords.dtype == torch.int32 and coords.is_contiguous() and coords.shape[1] == 4
    n_in = int(coords.shape[0])
    K = ksize[0] * ksize[1] * ksize[2]
    dev = coords.device
    out_shape = conv_out_shape(list(spatial_shape), ksize, stride, padding, dilation)
    ish, osh = native.host_ints(spatial_shape), native.host_ints(out_shape)
    ks, st, pd, dl = (native.host_ints(ksize), native.host_ints(stride), native.host_ints(padding),
                      native.host_ints(dilation))
    order_id = {"linear": 0, "morton": 1}[order if order is not None else CONV_OUTPUT_ORDER]
    if order_id == 1 and (max(out_shape) > 4096 or (int(batch_size) << sum((s - 1).bit_length() for s in out_shape)) >= 2 ** 32 - 1):
        order_id = 0                       # key space too large for 32-bit Morton keys: canonical order
    nbytes = native.query_bytes("ud_rulebook_conv_workspace_bytes", int(batch_size), native.host_ptr(osh), order_id)
    ws = native.WORKSPACE.get(nbytes, dev, slot=_WS_RULEBOOK)
    n_out_dev = torch.empty((1,), dtype=torch.int32, device=dev)
    sp = native.stream_ptr(dev)
    if native.TIMER.enabled:
        native.TIMER.meta = {"K": K}
    rc = lib.ud_rulebook_conv_count(native.ptr(coords), n_in, int(batch_size), native.host_ptr(ish),
                                    native.host_ptr(osh), native.host_ptr(ks), native.host_ptr(st),
                                    native.host_ptr(pd), native.host_ptr(dl), order_id, native.ptr(n_out_dev),
                                    native.ptr(ws), ws.numel(), sp)
    native.check(rc, "ud_rulebook_conv_count")
    n_out = int(n_out_dev.item())   # host sync: output row count sizes the tables
    out_coords = torch.empty((n_out, 4), dtype=torch.int32, device=dev)
    nbr_in = torch.empty((n_in, K), dtype=torch.int32, device=dev)
    nbr_out = torch.empty((n_out, K), dtype=torch.int32, device=dev)
    # the input-major table is produced in both forms by the builder (coalesced 32-row blocks); the output-major one is
    # scattered, and scattering it into the tap-major form too measured slower (+0.8 ms per step) than transposing it
    # afterwards (Rulebook.tap_major, lazily, 4 calls per step)
    nbr_in_t, masks_in = _tap_major_buffers(n_in, K, dev)
    if native.TIMER.enabled:
        native.TIMER.meta = {"K": K}
    rc = lib.ud_rulebook_conv_fill(native.ptr(coords), n_in, int(batch_size), native.host_ptr(ish),
                                   native.host_ptr(osh), native.host_ptr(ks), native.host_ptr(st),
                                   native.host_ptr(pd), native.host_ptr(dl), order_id, n_out, native.ptr(out_coords),
                                   native.ptr(nbr_in), native.ptr(nbr_out), native.ptr(nbr_in_t), native.ptr(masks_in),
                                   None, None, native.ptr(ws), ws.numel(), sp)
    native.check(rc, "ud_rulebook_conv_fill")
    rb = Rulebook(False, ksize, stride, padding, dilation, spatial_shape, out_shape, n_in, n_out,
                  nbr_out, nbr_in, out_coords, batch_size)
    if n_in > 0 and n_out > 0:
        rb._tap_major[nbr_in.data_ptr()] = (nbr_in_t, masks_in)
    return rb


# ---------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) sparse convolution for the channel pairs the library supports; UD_SPCONV_TC=0 keeps
# every layer on the fp32 SIMT kernel (used by the tests to cross-check the two kernels)
USE_TENSOR_CORES = os.environ.get("UD_SPCONV_TC", "1") != "0"
# the tensor-core kernel's epilogue also accumulates the per-channel sum / sum of squares of its output, so the
# BatchNorm1d that follows skips its statistics pass (attribute `_ud_bn_sums` on the conv output, see batch_norm_act)
FUSE_BN_STATS = os.environ.get("UD_FUSE_BN_STATS", "1") != "0"
# weight gradients of the C >= 32 layers on the tensor cores (TF32 operands rounded to nearest, fp32 accumulation);
# UD_SPCONV_TC_WGRAD=0 keeps them on the fp32 SIMT kernel
USE_TC_WGRAD = os.environ.get("UD_SPCONV_TC_WGRAD", "1") != "0"


# id(weight) -> (weakref to the weight, its _version, packed image): re-pack only when the parameter object is new or
# was modified in place (optimizer steps, load_state_dict and every in-place op bump _version).  Writes THROUGH
# `weight.data` (p.data.copy_(...), EMA / weight surgery code) do not bump it: call invalidate_packed_weights() after them.
_PACK_CACHE = {}


def invalidate_packed_weights(weight=None):
    """Drops the cached packed image of `weight` (or all of them)."""
    if weight is None:
        _PACK_CACHE.clear()
    else:
        _PACK_CACHE.pop(id(weight), None)


# Pack straight from the parameter layout (ud_sparse_conv_tc_pack_weights_param) instead of making the (K, Cin, Cout) copy --
# and, for the data gradient, the flipped / transposed copy -- first: three small copy kernels less per layer and step.
PACK_FROM_PARAM = os.environ.get("UD_SPCONV_PACK_FROM_PARAM", "1") != "0"


def _packed_weights(w_k, K, cin, cout, sp, weight=None, param=None, mode=0):
    """`param` (optional): the weights in the spconv-2 parameter layout (Cout_p, kz, ky, kx, Cin_p), fp32 contiguous; `mode`
    0 forward / 1 dgrad of a SubM convolution / 2 dgrad of a strided one (see the C entry point).  Else `w_k` (K, cin, cout)."""
    lib = native.load()
    if weight is not None:
        hit = _PACK_CACHE.get(id(weight))
        if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2].device == weight.device:
            return hit[2]
    nbytes = native.query_bytes("ud_sparse_conv_tc_weight_bytes", K, cin, cout)
    if (PACK_FROM_PARAM and param is not None and param.dtype == torch.float32 and param.is_contiguous()
            and param.dim() == 5):
        packed = torch.empty((nbytes,), dtype=torch.uint8, device=param.device)
        rc = lib.ud_sparse_conv_tc_pack_weights_param(native.ptr(param), K, cin, cout, int(mode), native.ptr(packed), sp)
        native.check(rc, "ud_sparse_conv_tc_pack_weights_param")
    else:
        w_k = w_k() if callable(w_k) else w_k
        packed = torch.empty((nbytes,), dtype=torch.uint8, device=w_k.device)
        rc = lib.ud_sparse_conv_tc_pack_weights(native.ptr(w_k), K, cin, cout, native.ptr(packed), sp)
        native.check(rc, "ud_sparse_conv_tc_pack_weights")
    if weight is not None:
        if len(_PACK_CACHE) > 256:
            for k in [k for k, v in _PACK_CACHE.items() if v[0]() is None]:
                del _PACK_CACHE[k]
        _PACK_CACHE[id(weight)] = (weakref.ref(weight), weight._version, packed)
    return packed


def _conv_forward_raw(feats, w_k, bias, table, n_out, K, cin, cout, rb=None, weight=None, want_stats=False, post=None,
                      param=None, param_mode=0):
    """`post` = (scale, shift, residual | None, relu): inference epilogue relu?(conv * scale + shift (+ residual)); only the
    tensor-core kernel has it (the caller checks ud_sparse_conv_tc_supported)."""
    lib = native.load()
    dev = feats.device
    out = torch.empty((n_out, cout), dtype=torch.float32, device=dev)
    sp = native.stream_ptr(dev)
    meta = {"n_in": int(feats.shape[0]), "n_out": n_out, "K": K, "cin": cin, "cout": cout,
            "table": table} if native.TIMER.enabled else None
    if USE_TENSOR_CORES and n_out > 0 and lib.ud_sparse_conv_tc_supported(K, cin, cout):
        packed = _packed_weights(w_k, K, cin, cout, sp, weight, param, param_mode)
        if meta is not None:
            meta["kernel"] = "k_spconv_umma_dg"
        table_t, masks = (rb if rb is not None else Rulebook.for_table(table)).tap_major(table)
        if post is not None:
            scale, shift, residual, relu = post
            if meta is not None:
                meta["residual"] = residual is not None
            native.TIMER.meta = meta
            rc = lib.ud_sparse_conv_tc_forward_affine(native.ptr(feats), native.ptr(packed), native.ptr(bias), native.ptr(scale),
                                                      native.ptr(shift), native.ptr(residual), 1 if relu else 0,
                                                      native.ptr(table_t), native.ptr(masks), n_out, K, cin, cout,
                                                      native.ptr(out), sp)
            native.check(rc, "ud_sparse_conv_tc_forward_affine")
            return out
        stats = None
        if want_stats and FUSE_BN_STATS:
            # scratch of the kernel's fixed-order statistics reduction; its first 2 * cout doubles are the result
            stats = torch.empty((native.query_bytes("ud_sparse_conv_tc_stats_bytes", n_out, cout) // 8,),
                                dtype=torch.float64, device=dev)
        if USE_WINDOW_KERNEL and lib.ud_sparse_conv_tc_window_supported(K, cin, cout):
            holder = rb if rb is not None else Rulebook.for_table(table)
            if holder is not rb:
                holder._tap_major[table.data_ptr()] = (table_t, masks)
            win_lo = holder.windows(table)
            if meta is not None:
                meta["kernel"] = "k_spconv_umma_win"
            native.TIMER.meta = meta
            rc = lib.ud_sparse_conv_tc_forward_window(native.ptr(feats), int(feats.shape[0]), native.ptr(packed),
                                                      native.ptr(bias), native.ptr(table_t), native.ptr(masks),
                                                      native.ptr(win_lo), n_out, K, cin, cout, native.ptr(out),
                                                      native.ptr(stats), sp)
            native.check(rc, "ud_sparse_conv_tc_forward_window")
        else:
            src, is_split = feats, 0
            if PRE_SPLIT_INPUT and feats.shape[0] > 0:
                # split the rows into bf16 hi/lo ONCE here instead of once per (row, tap) inside the gather loop
                src = torch.empty_like(feats)
                rc = lib.ud_split_bf16(native.ptr(feats), int(feats.shape[0]), cin, native.ptr(src), sp)
                native.check(rc, "ud_split_bf16")
                is_split = 1
            native.TIMER.meta = meta
            rc = lib.ud_sparse_conv_tc_forward(native.ptr(src), is_split, native.ptr(packed), native.ptr(bias),
                                               native.ptr(table_t), native.ptr(masks), n_out, K, cin, cout, native.ptr(out),
                                               native.ptr(stats), sp)
            native.check(rc, "ud_sparse_conv_tc_forward")
        if stats is not None:
            out._ud_bn_sums = stats[: 2 * cout].view(2, cout)
        return out
    if meta is not None:
        meta["kernel"] = "k_spconv_gather_gemm"
    if callable(w_k):
        w_k = w_k()
    native.TIMER.meta = meta
    rc = lib.ud_sparse_conv_forward(native.ptr(feats), native.ptr(w_k), native.ptr(bias), native.ptr(table),
                                    n_out, K, cin, cout, native.ptr(out), sp)
    native.check(rc, "ud_sparse_conv_forward")
    return out


def weight_to_taps(weight):
    """spconv-2 parameter layout (Cout, kz, ky, kx, Cin) -> kernel layout (K, Cin, Cout)."""
    cout, cin = weight.shape[0], weight.shape[-1]
    return weight.reshape(cout, -1, cin).permute(1, 2, 0).contiguous()


class SparseConvFunction(Function):
    """out = sparse_conv(features, weight, bias) for a given Rulebook; custom dgrad/wgrad kernels."""

    @staticmethod
    def forward(ctx, features, weight, bias, rulebook, inverse=False):
        native.require_cuda(features, weight)
        features = features.contiguous().float()
        rb = rulebook
        cout, cin = int(weight.shape[0]), int(weight.shape[-1])
        w_cell = []

        def w_k_fn():                       # (K, Cin, Cout) copy, only materialised when something needs it
            if not w_cell:
                w_cell.append(weight_to_taps(weight.detach().float()))
            return w_cell[0]

        if not inverse:
            table, n_out = rb.nbr_out, rb.n_out
        else:  # SparseInverseConv3d: roles of the paired strided conv swapped
            table, n_out = rb.nbr_in, rb.n_in
        assert features.shape[1] == cin, "feature width %d != conv in_channels %d" % (features.shape[1], cin)
        b = bias.detach().float().contiguous() if bias is not None else None
        w_param = weight.detach()
        if w_param.dtype != torch.float32 or not w_param.is_contiguous():
            w_param = w_param.float().contiguous()
        out = _conv_forward_raw(features, w_k_fn, b, table, n_out, rb.K, cin, cout, rb, weight, want_stats=True,
                                param=w_param, param_mode=0)
        ctx.rb, ctx.inverse = rb, inverse
        ctx.has_bias = bias is not None
        ctx.wshape = tuple(weight.shape)
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            ctx.save_for_backward(features, w_param)      # the parameter layout: the dgrad packs its own form from it
        return out

    @staticmethod
    def backward(ctx, grad_out):
        features, w_param = ctx.saved_tensors
        rb, inverse = ctx.rb, ctx.inverse
        lib = native.load()
        grad_out = grad_out.contiguous().float()
        cout, cin = int(w_param.shape[0]), int(w_param.shape[-1])
        K = rb.K
        dfeat = dweight = dbias = None
        if ctx.needs_input_grad[0]:
            if rb.subm:
                w_t = lambda: weight_to_taps(w_param).flip(0).transpose(1, 2).contiguous()   # (K, cout, cin), taps mirrored
                table, n_in, mode = rb.nbr_out, rb.n_in, 1
            elif not inverse:
                w_t = lambda: weight_to_taps(w_param).transpose(1, 2).contiguous()
                table, n_in, mode = rb.nbr_in, rb.n_in, 2
            else:
                w_t = lambda: weight_to_taps(w_param).transpose(1, 2).contiguous()
                table, n_in, mode = rb.nbr_out, rb.n_out, 2
            dfeat = _conv_forward_raw(grad_out, w_t, None, table, n_in, K, cout, cin, rb, param=w_param, param_mode=mode)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros((K, cin, cout), dtype=torch.float32, device=grad_out.device)
            if not inverse:
                table, n_rows = rb.nbr_out, rb.n_out
            else:
                table, n_rows = rb.nbr_in, rb.n_in
            sp = native.stream_ptr(grad_out.device)
            if USE_TENSOR_CORES and USE_TC_WGRAD and n_rows > 0 and lib.ud_sparse_conv_tc_wgrad_supported(K, cin, cout):
                table_t, masks = rb.tap_major(table)           # cached by the forward pass (same table)
                if native.TIMER.enabled:
                    native.TIMER.meta = {"n_in": int(features.shape[0]), "table": table}
                rc = lib.ud_sparse_conv_tc_wgrad(native.ptr(features), native.ptr(grad_out), native.ptr(table_t),
                                                 native.ptr(masks), n_rows, K, cin, cout, native.ptr(dw), sp)
                native.check(rc, "ud_sparse_conv_tc_wgrad")
            else:
                if native.TIMER.enabled:
                    native.TIMER.meta = {"n_in": int(features.shape[0]), "table": table}
                rc = lib.ud_sparse_conv_wgrad(native.ptr(features), native.ptr(grad_out), native.ptr(table),
                                              n_rows, K, cin, cout, native.ptr(dw), sp)
                native.check(rc, "ud_sparse_conv_wgrad")
            dweight = dw.permute(2, 0, 1).reshape(ctx.wshape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            if ZERO_BIAS_GRAD_UNDER_BN and getattr(grad_out, "_ud_zero_column_sums", False):
                dbias = torch.zeros((cout,), dtype=torch.float32, device=grad_out.device)
            else:
                dbias = column_sums(grad_out)
        return dfeat, dweight, dbias, None, None


# see BatchNormActFunction.backward; UD_ZERO_BIAS_GRAD_UNDER_BN=0 computes the (round-off) column sums instead
ZERO_BIAS_GRAD_UNDER_BN = os.environ.get("UD_ZERO_BIAS_GRAD_UNDER_BN", "1") != "0"


def column_sums(x):
    """(N, C) -> (C,) column sums in one pass with the fixed-order reduction of ud_bn_stats (torch's sum(0) over a tall
    matrix runs at ~0.6 TB/s and is not reproducible); falls back to torch for widths the kernel does not take."""
    n, C = int(x.shape[0]), int(x.shape[1])
    if not x.is_cuda or C > 128 or 256 % C != 0 or not x.is_contiguous():
        return x.sum(0)
    lib = native.load()
    sums = torch.empty((2, C), dtype=torch.float64, device=x.device)
    ws = native.WORKSPACE.get(native.query_bytes("ud_bn_workspace_bytes", n, C), x.device)
    rc = lib.ud_bn_stats(native.ptr(x), n, C, native.ptr(sums), native.ptr(ws), ws.numel(), native.stream_ptr(x.device))
    native.check(rc, "ud_bn_stats")
    return sums[0].float()


# Two measured-and-shelved variants of the tensor-core forward kernel, both bit-identical to the default one
# (tests/test_gpu_sparse_tc.py) and both OFF by default because they lost on B200 (profiles/r02_conv_experiments.md):
#  * UD_SPCONV_WINDOW=1: 3x3x3 layers with 64 input channels go through k_spconv_umma_win -- TMA stages, once per
#    (tile, kz, ky), the 160-row window holding 97 % of the rows the three kx taps gather; operands built from shared
#    memory.  12 calls: 7.45 ms vs 6.44 ms.
#  * UD_SPCONV_PRESPLIT=1: rows are split into bf16 hi/lo once (ud_split_bf16) instead of once per gathered copy.  The
#    convolutions get 7-15 % faster (64->64 6.44 -> 6.01 ms), the 62 split passes cost 2.4 ms: net +2.7 ms per step.
USE_WINDOW_KERNEL = os.environ.get("UD_SPCONV_WINDOW", "0") != "0"
PRE_SPLIT_INPUT = os.environ.get("UD_SPCONV_PRESPLIT", "0") != "0"

# conv_input of VoxelResBackBone8x has 5 input channels: 20-byte rows that the tensor-core kernels cannot gather (16-byte
# vector loads, 16 channels per MMA).  Zero-padding rows and weights to 16 channels makes it a <16,Cout> instance of the
# tcgen05 forward and weight-gradient kernels instead of the fp32 SIMT pair (measured per step: wgrad 1.23 -> 0.16 ms,
# forward 0.15 -> 0.10 ms); the padded channels contribute exact zeros, the weight gradient is sliced back by autograd.
PAD_NARROW_INPUT = os.environ.get("UD_SPCONV_PAD_INPUT", "1") != "0"


FUSE_EVAL_BN = os.environ.get("UD_SPCONV_FUSE_EVAL_BN", "1") != "0"
_EVAL_COEFFS = {}


def _eval_bn_coeffs(bn):
    """(scale, shift) of an eval-mode BatchNorm1d, cached until a parameter or buffer of the module is written to (tensor
    version counters)."""
    tensors = (bn.running_mean, bn.running_var, bn.weight, bn.bias)
    key = tuple((t.data_ptr(), t._version) if t is not None else None for t in tensors) + (bn.eps,)
    hit = _EVAL_COEFFS.get(id(bn))
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    lib = native.load()
    C = int(bn.running_mean.shape[0])
    dev = bn.running_mean.device
    scale = torch.empty((C,), dtype=torch.float32, device=dev)
    shift = torch.empty((C,), dtype=torch.float32, device=dev)
    rc = lib.ud_bn_eval_coeffs(native.ptr(bn.running_mean), native.ptr(bn.running_var), native.ptr(bn.weight),
                               native.ptr(bn.bias), float(bn.eps), C, native.ptr(scale), native.ptr(shift),
                               native.stream_ptr(dev))
    native.check(rc, "ud_bn_eval_coeffs")
    _EVAL_COEFFS[id(bn)] = (key, scale, shift)
    return scale, shift


def sparse_conv_eval_bn_act(features, weight, bias, rulebook, bn, relu=True, residual=None):
    """relu?(bn(sparse_conv(features)) (+ residual)) for an eval-mode BatchNorm1d `bn` as ONE kernel (the convolution's
    epilogue applies scale / shift / residual / ReLU) when nothing needs a gradient; returns None when the fused form does
    not apply (the caller then runs convolution and batch_norm_act separately -- bit-identical results either way)."""
    if not (FUSE_EVAL_BN and USE_TENSOR_CORES and features.is_cuda and not bn.training and bn.running_mean is not None):
        return None
    if torch.is_grad_enabled() and (features.requires_grad or weight.requires_grad
                                    or (bn.weight is not None and bn.weight.requires_grad)
                                    or (residual is not None and residual.requires_grad)):
        return None
    if USE_WINDOW_KERNEL or PRE_SPLIT_INPUT:
        return None
    cin, cout = int(weight.shape[-1]), int(weight.shape[0])
    lib = native.load()
    feats, w = features, weight
    if cin < 16:
        if not PAD_NARROW_INPUT:
            return None
        feats = torch.nn.functional.pad(features, (0, 16 - cin))
        w = torch.nn.functional.pad(weight.detach(), (0, 16 - cin))
        cin = 16
    rb = rulebook
    if rb.n_out <= 0 or not lib.ud_sparse_conv_tc_supported(rb.K, cin, cout):
        return None
    feats = feats.detach().contiguous().float()
    scale, shift = _eval_bn_coeffs(bn)
    res = residual.detach().contiguous().float() if residual is not None else None
    w_det = w.detach()
    b = bias.detach().float().contiguous() if bias is not None else None
    return _conv_forward_raw(feats, lambda: weight_to_taps(w_det.float()), b, rb.nbr_out, rb.n_out, rb.K, cin, cout, rb,
                             weight if w is weight else None, post=(scale, shift, res, relu))


def sparse_conv(features, weight, bias, rulebook, inverse=False):
    cin, cout = int(weight.shape[-1]), int(weight.shape[0])
    if (PAD_NARROW_INPUT and USE_TENSOR_CORES and features.is_cuda and cin < 16 and not inverse
            and native.load().ud_sparse_conv_tc_supported(rulebook.K, 16, cout)):
        features = torch.nn.functional.pad(features, (0, 16 - cin))
        weight = torch.nn.functional.pad(weight, (0, 16 - cin))
    return SparseConvFunction.apply(features, weight, bias, rulebook, inverse)


# ---------------------------------------------------------------------------------------------------
class BatchNormActFunction(Function):
    """y = relu?(batch_norm(x) (+ residual)) over the rows of an (N, C) matrix; two kernels forward."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, running_mean, running_var, training, momentum, eps, relu,
                fused_sums=None):
        native.require_cuda(x)
        lib = native.load()
        x = x.contiguous()
        n, C = int(x.shape[0]), int(x.shape[1])
        dev = x.device
        sp = native.stream_ptr(dev)
        scale = torch.empty((C,), dtype=torch.float32, device=dev)
        shift = torch.empty((C,), dtype=torch.float32, device=dev)
        res = residual.contiguous() if residual is not None else None
        if training:
            save_mean = torch.empty((C,), dtype=torch.float32, device=dev)
            save_invstd = torch.empty((C,), dtype=torch.float32, device=dev)
            sums = fused_sums
            if sums is None or tuple(sums.shape) != (2, C) or sums.device != dev:
                sums = torch.empty((2, C), dtype=torch.float64, device=dev)
                ws = native.WORKSPACE.get(native.query_bytes("ud_bn_workspace_bytes", n, C), dev)
                rc = lib.ud_bn_stats(native.ptr(x), n, C, native.ptr(sums), native.ptr(ws), ws.numel(), sp)
                native.check(rc, "ud_bn_stats")
            rc = lib.ud_bn_finalize(native.ptr(sums[0]), native.ptr(sums[1]), n, C, native.ptr(gamma),
                                    native.ptr(beta), float(eps), float(momentum), native.ptr(running_mean),
                                    native.ptr(running_var), native.ptr(scale), native.ptr(shift),
                                    native.ptr(save_mean), native.ptr(save_invstd), sp)
            native.check(rc, "ud_bn_finalize")
        else:
            save_mean = save_invstd = None
            rc = lib.ud_bn_eval_coeffs(native.ptr(running_mean), native.ptr(running_var), native.ptr(gamma),
                                       native.ptr(beta), float(eps), C, native.ptr(scale), native.ptr(shift), sp)
            native.check(rc, "ud_bn_eval_coeffs")
        y = torch.empty_like(x)
        rc = lib.ud_bn_apply(native.ptr(x), n, C, native.ptr(scale), native.ptr(shift), native.ptr(res),
                             1 if relu else 0, native.ptr(y), sp)
        native.check(rc, "ud_bn_apply")
        ctx.training, ctx.relu, ctx.has_res = training, relu, residual is not None
        ctx.has_affine = gamma is not None
        if training:
            ctx.save_for_backward(x, y, gamma, save_mean, save_invstd)
        else:
            ctx.save_for_backward(y, scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = native.load()
        dy = dy.contiguous()
        dev = dy.device
        if ctx.training:
            x, y, gamma, mean, invstd = ctx.saved_tensors
            n, C = int(x.shape[0]), int(x.shape[1])
            sums = torch.empty((2, C), dtype=torch.float64, device=dev)
            dx = torch.empty_like(x)
            dres = torch.empty_like(x) if ctx.has_res else None
            ws = native.WORKSPACE.get(native.query_bytes("ud_bn_workspace_bytes", n, C), dev)
            rc = lib.ud_bn_backward(native.ptr(dy), native.ptr(y), native.ptr(x), n, C, native.ptr(mean),
                                    native.ptr(invstd), native.ptr(gamma), 1 if ctx.relu else 0,
                                    native.ptr(sums), native.ptr(dx), native.ptr(dres), native.ptr(ws), ws.numel(),
                                    native.stream_ptr(dev))
            native.check(rc, "ud_bn_backward")
            dgamma = sums[1].float() if ctx.has_affine else None
            dbeta = sums[0].float() if ctx.has_affine else None
            # The gradient w.r.t. the input of a train-mode BatchNorm sums to zero over the rows of every channel
            # (dx = k (g - mean g - xhat mean(g xhat)) and sum xhat = 0 for batch statistics).  A convolution bias right in
            # front of it therefore has a zero gradient; the mark saves that convolution's backward a column-sum pass over dx
            # (16 of them per step in VoxelResBackBone8x, 0.6 ms on B200) whose result is round-off noise.
            dx._ud_zero_column_sums = True
            return dx, dgamma, dbeta, dres, None, None, None, None, None, None, None
        y, scale = ctx.saved_tensors
        g = dy * (y > 0).to(dy.dtype) if ctx.relu else dy
        dx = g * scale
        dres = g if ctx.has_res else None
        # eval-mode BN: parameters are treated as constants (frozen-teacher use)
        return dx, None, None, dres, None, None, None, None, None, None, None


def batch_norm_act(x, bn, relu=True, residual=None):
    """Fused replacement for `relu(bn(x) + residual)` with an nn.BatchNorm1d module `bn`."""
    training = bn.training or (bn.running_mean is None)
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = bn.momentum if bn.momentum is not None else 0.1
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    if not training and torch.is_grad_enabled() and bn.weight is not None and bn.weight.requires_grad:
        # eval-mode BN with trainable affine parameters: keep exact autograd semantics
        y = torch.nn.functional.batch_norm(x, rm, rv, bn.weight, bn.bias, False, momentum, bn.eps)
        if residual is not None:
            y = y + residual
        return torch.relu(y) if relu else y
    fused = getattr(x, "_ud_bn_sums", None) if (training and x.is_contiguous()) else None
    return BatchNormActFunction.apply(x, bn.weight, bn.bias, residual, rm, rv, training, momentum, bn.eps, relu, fused)


# ---------------------------------------------------------------------------------------------------
class ScatterToBEVFunction(Function):
    """(N, C) rows + (N,4) [b,z,y,x] -> dense (B, C, D, H, W) (== (B, C*D, H, W) memory)."""

    @staticmethod
    def forward(ctx, features, coords, batch_size, spatial_shape):
        native.require_cuda(features, coords)
        lib = native.load()
        features = features.contiguous().float()
        n, C = int(features.shape[0]), int(features.shape[1])
        D, H, W = [int(s) for s in spatial_shape]
        out = torch.empty((batch_size, C, D, H, W), dtype=torch.float32, device=features.device)
        rc = lib.ud_scatter_to_bev(native.ptr(features), native.ptr(coords), n, int(batch_size), C, D, H, W, 1,
                                   native.ptr(out), native.stream_ptr(features.device))
        native.check(rc, "ud_scatter_to_bev")
        ctx.save_for_backward(coords)
        ctx.dims = (n, C, D, H, W, int(batch_size))
        return out

    @staticmethod
    def backward(ctx, grad):
        (coords,) = ctx.saved_tensors
        n, C, D, H, W, B = ctx.dims
        lib = native.load()
        grad = grad.contiguous()
        dfeat = torch.empty((n, C), dtype=torch.float32, device=grad.device)
        rc = lib.ud_gather_from_bev(native.ptr(grad), native.ptr(coords), n, B, C, D, H, W, native.ptr(dfeat),
                                    native.stream_ptr(grad.device))
        native.check(rc, "ud_gather_from_bev")
        return dfeat, None, None, None


def scatter_to_dense(features, coords, batch_size, spatial_shape):
    return ScatterToBEVFunction.apply(features, coords, batch_size, spatial_shape)

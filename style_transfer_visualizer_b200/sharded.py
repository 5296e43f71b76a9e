"""Row-band sharding of ONE large image over several GPUs (BASELINE.json configs[4]: 3840x2160).

The reference runs the whole image on one device (it only warns above 3000 px, image_io.py:49-61).
Here every rank owns a horizontal band of the image, aligned to 16 rows so that the four 2x2 pools
never straddle a band boundary.  Per step:

  * every activation / gradient buffer carries one halo row above and below its band;
  * a 3x3 conv (or dgrad) is run over the band *including* its halo rows, which makes the band's own
    rows exact; the two halo rows of the result are then refreshed from the neighbours: every rank's
    own kernel stores its boundary rows straight into the neighbours' halo rows over NVLink
    (activation buffers live in peer-mapped symmetric memory, ``stv_halo_exchange``; flag words
    hand-shake, no NCCL call on the path).  Where symmetric memory is unavailable the exchange
    falls back to one NCCL send/recv pair per direction (image-boundary halos are zeroed = the
    conv's zero padding);
  * each rank contracts its own pixels into a raw Gram partial; ONE all-reduce per step sums the five
    C x C partials (and the content-loss partial sums) over NVLink, after which clamp / 1/N / MSE /
    backward seed are applied to the global matrices on every rank (the clamp is non-linear, so it
    must see the global sum);
  * the style backward uses the global seed with local features; Adam is band-local.

No other collective is on the data path.  ``plan_bands`` / ``exchange_rows`` are device-agnostic
(CPU + gloo in the tests); the engine itself needs sm_100a GPUs and NCCL.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.distributed as dist
from torch import nn

from . import _native as nat
from . import ops
from .engine import VggLossEngine

BAND_ALIGN = 16  # 2^4: four pooling levels stay band-local


def plan_bands(height: int, world_size: int, align: int = BAND_ALIGN) -> list[tuple[int, int]]:
    """Split ``height`` rows into ``world_size`` contiguous bands whose starts are multiples of
    ``align`` (the last band takes the unaligned remainder).  Returns [(y0, y1)] per rank."""
    units = (height + align - 1) // align
    if units < world_size:
        msg = f"{height} rows are too few for {world_size} bands of {align}-aligned rows"
        raise ValueError(msg)
    base, extra = divmod(units, world_size)
    bands, start = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        y0, y1 = start * align, min((start + n) * align, height)
        bands.append((y0, y1))
        start += n
    return bands


def exchange_rows(buf: torch.Tensor, rank: int, world_size: int, group=None) -> None:  # noqa: ANN001
    """Refresh the halo rows of ``buf`` ``[B + 2, ...]`` (row 0 / row B+1) from the neighbouring
    ranks' first / last own rows; halos on the image boundary are zeroed."""
    rows = buf.shape[0]
    reqs = []
    if rank > 0:
        reqs.append(dist.P2POp(dist.isend, buf[1], rank - 1, group))
        reqs.append(dist.P2POp(dist.irecv, buf[0], rank - 1, group))
    else:
        buf[0].zero_()
    if rank < world_size - 1:
        reqs.append(dist.P2POp(dist.isend, buf[rows - 2], rank + 1, group))
        reqs.append(dist.P2POp(dist.irecv, buf[rows - 1], rank + 1, group))
    else:
        buf[rows - 1].zero_()
    if reqs:
        for req in dist.batch_isend_irecv(reqs):
            req.wait()


def rows_at(band: int, level: int) -> int:
    """Own rows of a band after ``level`` floor-mode 2x2 pools."""
    for _ in range(level):
        band //= 2
    return band


class _PeerArena:
    """Symmetric (peer-mapped) memory of one workspace: ONE arena holding every haloed buffer, so a
    single rendezvous maps all of them into the neighbours' address spaces, plus the flag words of
    the halo hand-shake.  All ranks allocate identical sizes (sized for the largest band)."""

    MAX_SLOTS = 64

    def __init__(self, total_floats: int, device: torch.device, group) -> None:  # noqa: ANN001
        import torch.distributed._symmetric_memory as symm

        grp = group if group is not None else dist.group.WORLD
        self.arena = symm.empty(total_floats, dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.arena, grp)
        self.flags = symm.empty(self.MAX_SLOTS * 4, dtype=torch.int32, device=device)
        self.flag_handle = symm.rendezvous(self.flags, grp)
        self.arena.zero_()
        self.flags.zero_()
        self.epoch = torch.zeros(self.MAX_SLOTS, device=device, dtype=torch.int32)
        self.done = torch.zeros(self.MAX_SLOTS, device=device, dtype=torch.int32)
        self.base = [int(p) for p in self.handle.buffer_ptrs]
        self.flag_base = [int(p) for p in self.flag_handle.buffer_ptrs]
        self.used = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)   # nobody pushes into an arena that is still being zeroed

    def take(self, n_floats: int) -> tuple[torch.Tensor, int]:
        """Next ``n_floats`` of the arena (256-byte aligned): (flat view, offset in floats)."""
        off = self.used
        self.used += (n_floats + 63) // 64 * 64
        if self.used > self.arena.numel():
            msg = "symmetric arena exhausted (sizing bug)"
            raise RuntimeError(msg)
        return self.arena[off:off + n_floats], off


@dataclass
class _BandWorkspace:
    band: int                       # own rows at full resolution
    width: int
    rows: list[int] = field(default_factory=list)       # own rows per stage
    cols: list[int] = field(default_factory=list)
    x_h: torch.Tensor | None = None                     # [1, 3, band + 2, W]
    pre: list[torch.Tensor | None] = field(default_factory=list)    # haloed [rows+2, W, C]
    post: list[torch.Tensor | None] = field(default_factory=list)
    pool: list[torch.Tensor | None] = field(default_factory=list)
    d_y: list[torch.Tensor | None] = field(default_factory=list)
    d_pool: list[torch.Tensor | None] = field(default_factory=list)
    bits: list[torch.Tensor | None] = field(default_factory=list)   # ReLU sign bits, own rows
    code: list[torch.Tensor | None] = field(default_factory=list)   # pool routing bits, own rows
    gram_ws: list[torch.Tensor] = field(default_factory=list)
    s_mat: list[torch.Tensor] = field(default_factory=list)
    reduce_buf: torch.Tensor | None = None              # [sum C^2 + n_content] one all-reduce
    raw: list[torch.Tensor] = field(default_factory=list)   # per-layer raw Gram partial / global R
    losses: torch.Tensor | None = None
    scratch: torch.Tensor | None = None
    grad_h: torch.Tensor | None = None                  # [1, 3, band + 2, W]
    img_send: torch.Tensor | None = None
    img_recv: torch.Tensor | None = None
    bands: list[tuple[int, int]] = field(default_factory=list)   # band plan of the whole image
    peer: _PeerArena | None = None                      # NVLink halo exchange (else NCCL)
    peer_g: _PeerArena | None = None                    # same for the gradient buffers
    # data_ptr of a haloed buffer -> (arena offset in floats, pooling level of its rows)
    where: dict[int, tuple[int, int]] = field(default_factory=dict)


class RowBandEngine:
    """Forward losses / input gradient of one row band; all ranks call every method together."""

    def __init__(self, layers: list[nn.Module], style_idx: list[int], content_idx: list[int],
                 device: torch.device, group=None) -> None:  # noqa: ANN001
        if not dist.is_initialized():
            msg = "RowBandEngine needs an initialised torch.distributed process group"
            raise RuntimeError(msg)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device
        self.base = VggLossEngine(layers, style_idx, content_idx, device)  # stages + packed weights
        self.stages = self.base.stages
        self.style_idx = self.base.style_idx
        self.content_idx = self.base.content_idx
        conv_idx = {st.conv_idx for st in self.stages}
        if not (set(self.style_idx) | set(self.content_idx)) <= conv_idx:
            msg = "row-band sharding supports losses on convolution outputs only (the default taps)"
            raise NotImplementedError(msg)
        self._ws: dict[tuple[int, int], _BandWorkspace] = {}
        # "nvlink-peer": boundary rows stored into the neighbours' buffers by this GPU's kernels;
        # "nccl": send/recv pairs.  STV_HALO=nccl forces the fallback (A/B measurements).
        import os

        self.halo_mode = "nccl" if os.environ.get("STV_HALO") == "nccl" or self.world == 1 \
            else "nvlink-peer"
        self._side = torch.cuda.Stream(device=device)   # loss kernels + their all-reduces
        self.bands: list[tuple[int, int]] | None = None   # set per image (plan_bands)
        self.style_targets: list[torch.Tensor] | None = None
        self.content_targets: list[torch.Tensor] | None = None   # own rows, NHWC
        self.full_hw: tuple[int, int] | None = None
        self.band_rows: tuple[int, int] | None = None

    # ------------------------------------------------------------------ helpers
    def _xchg(self, ws: _BandWorkspace, buf: torch.Tensor, slot: int, planes: int = 1, *,
              grads: bool = False) -> None:
        """Refresh the halo rows of ``buf`` from the neighbouring ranks (zeros at the image edge)."""
        pa = ws.peer_g if grads else ws.peer
        if pa is None or (buf.shape[-1] * (1 if planes == 3 else buf.shape[-2])) % 4:
            if planes == 3:   # NCHW image band: packed rows through NCCL
                self._xchg_image_nccl(ws, buf)
            else:
                exchange_rows(buf, self.rank, self.world, self.group)
            return
        off, level = ws.where[buf.data_ptr()]
        band_of = [y1 - y0 for (y0, y1) in ws.bands]
        rows = rows_at(band_of[self.rank], level)
        up = self.rank - 1 if self.rank > 0 else None
        down = self.rank + 1 if self.rank < self.world - 1 else None
        row_floats = buf.numel() // ((rows + 2) * planes)
        ops.halo_exchange(
            buf, up_ptr=None if up is None else pa.base[up] + 4 * off,
            down_ptr=None if down is None else pa.base[down] + 4 * off,
            rows=rows, rows_up=0 if up is None else rows_at(band_of[up], level),
            rows_down=0 if down is None else rows_at(band_of[down], level),
            row_floats=row_floats, planes=planes, flags_mine=pa.flags,
            flags_up_ptr=None if up is None else pa.flag_base[up],
            flags_down_ptr=None if down is None else pa.flag_base[down],
            epoch=pa.epoch, done=pa.done, slot=slot,
            wait_ready=False)  # every producer of this engine stores own rows only

    def _xchg_image(self, ws: _BandWorkspace, img_h: torch.Tensor) -> None:
        """Halo rows of the NCHW image band: the 3 planes' rows are packed into one message."""
        if ws.peer is not None and img_h is ws.x_h:
            self._xchg(ws, img_h, slot=0, planes=3)
            return
        self._xchg_image_nccl(ws, img_h)

    def _xchg_image_nccl(self, ws: _BandWorkspace, img_h: torch.Tensor) -> None:
        b = ws.band
        send, recv = ws.img_send, ws.img_recv
        send[0].copy_(img_h[0, :, 1, :])
        send[1].copy_(img_h[0, :, b, :])
        reqs = []
        if self.rank > 0:
            reqs.append(dist.P2POp(dist.isend, send[0], self.rank - 1, self.group))
            reqs.append(dist.P2POp(dist.irecv, recv[0], self.rank - 1, self.group))
        if self.rank < self.world - 1:
            reqs.append(dist.P2POp(dist.isend, send[1], self.rank + 1, self.group))
            reqs.append(dist.P2POp(dist.irecv, recv[1], self.rank + 1, self.group))
        if reqs:
            for req in dist.batch_isend_irecv(reqs):
                req.wait()
        if self.rank > 0:
            img_h[0, :, 0, :].copy_(recv[0])
        else:
            img_h[0, :, 0, :].zero_()
        if self.rank < self.world - 1:
            img_h[0, :, b + 1, :].copy_(recv[1])
        else:
            img_h[0, :, b + 1, :].zero_()

    def _workspace(self, band: int, width: int, *, with_grad: bool) -> _BandWorkspace:
        """Buffers of one (band, width); creating one is COLLECTIVE in nvlink-peer mode (symmetric
        allocation + rendezvous), and every rank creates its workspaces in the same order."""
        key = (band, width)
        ws = self._ws.get(key)
        dev = self.device

        def buf(*shape: int) -> torch.Tensor:
            return torch.zeros(*shape, device=dev, dtype=torch.float32)

        def haloed_specs(grads: bool) -> list[tuple[str, int, int, int, int, int]]:
            """(list name, stage, pooling level, cols, channels, planes) of every haloed buffer."""
            out = []
            level, cols = 0, width
            if not grads:
                out.append(("x_h", -1, 0, width, 1, 3))
            for s, st in enumerate(self.stages):
                tapped = st.conv_idx in self.style_idx or st.conv_idx in self.content_idx
                if grads:
                    out.append(("d_y", s, level, cols, st.cout, 1))
                else:
                    if tapped or st.relu_idx is None:
                        out.append(("pre", s, level, cols, st.cout, 1))
                    if st.relu_idx is not None and st.pool_idx is None:
                        out.append(("post", s, level, cols, st.cout, 1))
                if st.pool_idx is not None:
                    level, cols = level + 1, cols // 2
                    if not grads:   # the pooled gradient never reaches memory (un-pooling dgrad)
                        out.append(("pool", s, level, cols, st.cout, 1))
            return out

        def allocate(specs, arena: _PeerArena | None) -> None:  # noqa: ANN001
            for name, s, level, cols, ch, planes in specs:
                rows = rows_at(band, level)
                shape = (1, 3, rows + 2, cols) if planes == 3 else (rows + 2, cols, ch)
                if arena is None:
                    t = buf(*shape)
                else:
                    rmax = rows_at(max(y1 - y0 for (y0, y1) in ws.bands), level)
                    flat, off = arena.take(planes * (rmax + 2) * cols * ch)
                    n = planes * (rows + 2) * cols * ch
                    t = flat[:n].view(shape)
                    ws.where[t.data_ptr()] = (off, level)
                if name == "x_h":
                    ws.x_h = t
                else:
                    getattr(ws, name)[s] = t

        def arena_for(specs) -> _PeerArena | None:  # noqa: ANN001
            if self.halo_mode != "nvlink-peer":
                return None
            bmax = max(y1 - y0 for (y0, y1) in ws.bands)
            total = sum((planes * (rows_at(bmax, level) + 2) * cols * ch + 63) // 64 * 64
                        for _n, _s, level, cols, ch, planes in specs)
            try:
                return _PeerArena(total, dev, self.group)
            except Exception as exc:  # noqa: BLE001
                # plumbing fallback (all ranks fail or succeed together: same driver, same node)
                from .logging_utils import logger

                logger.warning("symmetric memory unavailable (%s: %s); halo exchange falls back "
                               "to NCCL send/recv", type(exc).__name__, exc)
                self.halo_mode = "nccl"
                return None

        if ws is None:
            ws = _BandWorkspace(band=band, width=width)
            full_h = self.full_hw[0] if self.full_hw is not None else band * self.world
            ws.bands = plan_bands(full_h, self.world)
            if ws.bands[self.rank][1] - ws.bands[self.rank][0] != band:
                msg = (f"band of {band} rows does not match the band plan of a {full_h}-row image "
                       f"for rank {self.rank}: {ws.bands}")
                raise ValueError(msg)
            n = len(self.stages)
            rows, cols = band, width
            for st in self.stages:
                ws.rows.append(rows)
                ws.cols.append(cols)
                if st.pool_idx is not None:
                    rows, cols = rows // 2, cols // 2
            for name in ("pre", "post", "pool", "d_y", "d_pool", "bits", "code"):
                setattr(ws, name, [None] * n)
            for s, st in enumerate(self.stages):
                if st.relu_idx is None or s + 1 >= n:
                    continue
                words = (ws.rows[s], ws.cols[s], st.cout // 32)   # own rows only: no halo needed
                if st.pool_idx is not None:
                    ws.code[s] = torch.zeros(*words, device=dev, dtype=torch.int32)
                else:
                    ws.bits[s] = torch.zeros(*words, device=dev, dtype=torch.int32)
            specs = haloed_specs(grads=False)
            ws.peer = arena_for(specs)
            allocate(specs, ws.peer)
            total = 0
            for idx in self.style_idx:
                s = self._stage(idx)
                c = self.stages[s].cout
                ws.gram_ws.append(ops.gram_workspace(max(ws.rows[s] * ws.cols[s], 1), c, dev))
                ws.s_mat.append(buf(c, c))
                total += c * c
            ws.reduce_buf = buf(total + len(self.content_idx))
            ws.raw = [buf(self.stages[self._stage(idx)].cout, self.stages[self._stage(idx)].cout)
                      for idx in self.style_idx]
            ws.losses = buf(len(self.style_idx) + len(self.content_idx))
            ws.scratch = buf(2 * nat.reduce_scratch_floats() + 1024)
            ws.img_send = buf(2, 3, width)
            ws.img_recv = buf(2, 3, width)
            self._ws[key] = ws
        if with_grad and ws.grad_h is None:
            specs = haloed_specs(grads=True)
            ws.peer_g = arena_for(specs) if ws.peer is not None else None
            allocate(specs, ws.peer_g)
            ws.grad_h = buf(1, 3, band + 2, width)
        return ws

    def _round_pre(self, st) -> bool:  # noqa: ANN001
        return self.base._round_pre(st)  # noqa: SLF001

    def _stage(self, conv_idx: int) -> int:
        for s, st in enumerate(self.stages):
            if st.conv_idx == conv_idx:
                return s
        raise KeyError(conv_idx)

    @staticmethod
    def _own(t: torch.Tensor) -> torch.Tensor:
        """Own rows of a haloed NHWC buffer (contiguous view)."""
        return t[1:t.shape[0] - 1]

    def _full_rows_at(self, s: int) -> tuple[int, int]:
        """(rows, cols) of the WHOLE feature map at stage ``s`` (global floor-mode pooling)."""
        h, w = self.full_hw
        for st in self.stages[:s]:
            if st.pool_idx is not None:
                h, w = h // 2, w // 2
        return h, w

    # ------------------------------------------------------------------ forward
    def _run_stack(self, x_band: torch.Tensor, ws: _BandWorkspace, after_stage=None) -> None:  # noqa: ANN001
        """Forward through the band.  Every conv READS the haloed buffer of its input (x_row0 = 1: the
        halo rows supply the neighbours' pixels, zeros at the image edge) and WRITES own rows only,
        with the same fused epilogues as the single-GPU engine: ReLU sign bits, 2x2 max pool +
        routing bits (bands start on even rows, so pool windows never straddle a band), split ReLU
        for the wide tapped layers."""
        ws.x_h[:, :, 1:-1, :].copy_(x_band)
        self._xchg_image(ws, ws.x_h)
        cur = None
        n = len(self.stages)
        own = self._own
        for s, st in enumerate(self.stages):
            rows = ws.rows[s]
            rp = self._round_pre(st)
            pre = own(ws.pre[s]) if ws.pre[s] is not None else None
            post = own(ws.post[s]) if ws.post[s] is not None else None
            if s == 0:
                if self.base.first_layer_tc:
                    ops.conv3x3_first_fwd_tc(ws.x_h, st.weight, st.bias, pre, post, rows=rows,
                                             in_row0=1, round_pre=rp, out_bits=ws.bits[0])
                else:
                    ops.conv3x3_first_fwd_band(ws.x_h, st.weight, st.bias, pre, post, rows=rows,
                                               in_row0=1, round_pre=rp, out_bits=ws.bits[0])
            elif st.pool_idx is not None:
                ops.conv3x3_desc(cur, st.w_fwd, rows=rows, x_row0=1, bias=st.bias, out_pre=pre,
                                 round_flags=2 | int(rp), out_pool=own(ws.pool[s]),
                                 out_code=ws.code[s])
            elif pre is not None and post is not None and st.cout >= 256:
                ops.conv3x3_desc(cur, st.w_fwd, rows=rows, x_row0=1, bias=st.bias, out_pre=pre,
                                 round_flags=2 | int(rp))
                ops.relu_fwd_bits(pre, post, ws.bits[s])
            else:
                ops.conv3x3_desc(cur, st.w_fwd, rows=rows, x_row0=1, bias=st.bias, out_pre=pre,
                                 out_post=post, round_flags=2 | int(rp), out_bits=ws.bits[s])
            cur = ws.pool[s] if st.pool_idx is not None else \
                (ws.post[s] if ws.post[s] is not None else ws.pre[s])
            if after_stage is not None:
                after_stage(s, st)
            if s + 1 < n:
                self._xchg(ws, cur, slot=1 + s)  # halos for the next 3x3 conv

    def _reduce_grams(self, ws: _BandWorkspace, content_sums: list[torch.Tensor]) -> list[torch.Tensor]:
        """Raw Gram partials of the tapped layers + content partial sums -> ONE all-reduce."""
        views, off = [], 0
        for k, idx in enumerate(self.style_idx):
            s = self._stage(idx)
            c = self.stages[s].cout
            r = ws.reduce_buf[off:off + c * c].view(c, c)
            ops.gram_partial_r(self._own(ws.pre[s]), ws.gram_ws[k], r)
            views.append(r)
            off += c * c
        for k, v in enumerate(content_sums):
            ws.reduce_buf[off + k:off + k + 1].copy_(v)
        dist.all_reduce(ws.reduce_buf, op=dist.ReduceOp.SUM, group=self.group)
        return views

    def compute_targets(self, style_band: torch.Tensor, content_band: torch.Tensor,
                        style_hw: tuple[int, int], content_hw: tuple[int, int]) -> None:
        """Targets from this rank's bands of the style / content images (full sizes given)."""
        self.full_hw = style_hw
        ws = self._workspace(int(style_band.shape[2]), int(style_band.shape[3]), with_grad=False)
        self._run_stack(style_band, ws)
        raws = self._reduce_grams(ws, [])
        grams = []
        for k, idx in enumerate(self.style_idx):
            s = self._stage(idx)
            c = self.stages[s].cout
            fh, fw = self._full_rows_at(s)
            g = torch.empty(c, c, device=self.device, dtype=torch.float32)
            ops.gram_from_r(raws[k], float(c) * fh * fw, ws.scratch, gram_out=g)
            grams.append(g)
        self.style_targets = grams
        self.full_hw = content_hw
        ws = self._workspace(int(content_band.shape[2]), int(content_band.shape[3]),
                             with_grad=False)
        self._run_stack(content_band, ws)
        self.content_targets = [self._own(ws.pre[self._stage(idx)]).clone()
                                for idx in self.content_idx]

    def forward_losses(self, x_band: torch.Tensor) -> torch.Tensor:
        """Global losses (identical on every rank): style ascending, then content.

        The loss work of a tapped layer -- this band's raw Gram partial, its all-reduce over the
        ranks, clamp / 1/N / MSE / backward seed on the global matrix; the content partial sum and
        its all-reduce -- is issued on a side stream right after the layer, so all of it except
        the last layer's overlaps the rest of the forward pass (the communicator is used from the
        side stream only, in the same order on every rank)."""
        if self.style_targets is None or self.content_targets is None:
            msg = "targets must be set before computing losses."
            raise RuntimeError(msg)
        ws = self._workspace(int(x_band.shape[2]), int(x_band.shape[3]), with_grad=False)
        ns = len(self.style_idx)
        main = torch.cuda.current_stream(self.device)
        side = self._side

        def after_stage(s: int, st) -> None:  # noqa: ANN001
            style_k = self.style_idx.index(st.conv_idx) if st.conv_idx in self.style_idx else None
            content_k = self.content_idx.index(st.conv_idx) if st.conv_idx in self.content_idx \
                else None
            if style_k is None and content_k is None:
                return
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
            fh, fw = self._full_rows_at(s)
            total = float(st.cout) * fh * fw
            with torch.cuda.stream(side):
                feats = self._own(ws.pre[s])
                if style_k is not None:
                    r = ws.raw[style_k]
                    ops.gram_partial_r(feats, ws.gram_ws[style_k], r)
                    dist.all_reduce(r, op=dist.ReduceOp.SUM, group=self.group)
                    ops.gram_from_r(r, total, ws.scratch, target=self.style_targets[style_k],
                                    s_out=ws.s_mat[style_k], loss_out=ws.losses[style_k:style_k + 1])
                if content_k is not None:
                    slot = ws.losses[ns + content_k:ns + content_k + 1]
                    ops.content_loss_fwd(feats, self.content_targets[content_k], ws.scratch, slot)
                    slot.mul_(float(feats.numel()) / total)   # local mean -> share of the global mean
                    dist.all_reduce(slot, op=dist.ReduceOp.SUM, group=self.group)

        self._run_stack(x_band.detach(), ws, after_stage)
        main.wait_stream(side)
        return ws.losses

    # ------------------------------------------------------------------ backward
    def backward_losses(self, band: int, width: int, grad_w: torch.Tensor) -> torch.Tensor:
        """d(sum_k grad_w[k] * loss_k)/d(own rows of the image): ``[1, 3, band, W]`` view."""
        ws = self._workspace(band, width, with_grad=True)
        grad_w = grad_w.to(torch.float32).contiguous()
        ns = len(self.style_idx)
        n = len(self.stages)
        own = self._own
        for s in range(n - 1, -1, -1):
            st = self.stages[s]
            down = self.stages[s + 1] if s + 1 < n else None
            d_y = ws.d_y[s]
            rows = ws.rows[s]
            style_k = self.style_idx.index(st.conv_idx) if st.conv_idx in self.style_idx else None
            content_k = self.content_idx.index(st.conv_idx) if st.conv_idx in self.content_idx \
                else None

            def tap_grads(accumulate: bool) -> bool:
                if style_k is not None:  # noqa: B023
                    ops.style_bwd(own(ws.pre[s]), ws.s_mat[style_k], grad_w[style_k:style_k + 1],  # noqa: B023
                                  own(d_y), accumulate=accumulate)  # noqa: B023
                    accumulate = True
                if content_k is not None:  # noqa: B023
                    f = own(ws.pre[s])  # noqa: B023
                    fh, fw = self._full_rows_at(s)  # noqa: B023
                    total = float(st.cout) * fh * fw  # noqa: B023
                    # kernel scales by 2 / n_local; rescale the weight to the global element count
                    gw = grad_w[ns + content_k:ns + content_k + 1] * (float(f.numel()) / total)  # noqa: B023
                    ops.content_loss_bwd(f, self.content_targets[content_k], gw, own(d_y),  # noqa: B023
                                         accumulate=accumulate)
                    accumulate = True
                return accumulate

            if st.pool_idx is not None:
                # dgrad of the next conv at pooled resolution; its epilogue routes the gradient
                # through the pool + ReLU backward straight into this stage's own rows
                ops.conv3x3_desc(ws.d_y[s + 1], down.w_dgrad, rows=ws.rows[s + 1], x_row0=1,
                                 out_pre=own(d_y), round_flags=1, unpool_code=ws.code[s],
                                 unpool_hw=(rows, ws.cols[s]))
                tap_grads(True)
            elif st.relu_idx is not None:
                fused = False
                if style_k is not None and content_k is None and st.cout in (64, 128):
                    fused = ops.conv3x3_desc(
                        ws.d_y[s + 1], down.w_dgrad, rows=rows, x_row0=1, out_pre=own(d_y),
                        round_flags=1, mask_bits=ws.bits[s], style_x=own(ws.pre[s]),
                        style_s=ws.s_mat[style_k], style_alpha=grad_w[style_k:style_k + 1])
                if not fused:
                    acc = tap_grads(False)
                    ops.conv3x3_desc(ws.d_y[s + 1], down.w_dgrad, rows=rows, x_row0=1,
                                     out_pre=own(d_y), add_src=own(d_y) if acc else None,
                                     round_flags=1, mask_bits=ws.bits[s])
            else:
                tap_grads(False)
            self._xchg(ws, d_y, slot=s, grads=True)  # halos of the finished gradient feed the next dgrad
        ops.conv3x3_first_dgrad_rows(ws.d_y[0], self.stages[0].w_dgrad, ws.grad_h)
        return ws.grad_h[:, :, 1:-1, :]


class _BandLosses(torch.autograd.Function):
    """Autograd bridge: forward = global losses from a band, backward = the band's gradient."""

    @staticmethod
    def forward(ctx, x_band: torch.Tensor, engine: RowBandEngine) -> torch.Tensor:  # noqa: ANN001
        ctx.engine = engine
        ctx.shape = (int(x_band.shape[2]), int(x_band.shape[3]))
        return engine.forward_losses(x_band).clone()

    @staticmethod
    def backward(ctx, grad_losses: torch.Tensor):  # noqa: ANN001, ANN205
        band, width = ctx.shape
        grad = ctx.engine.backward_losses(band, width, grad_losses)
        return grad.contiguous().clone(), None


class ShardedStyleContentModel(nn.Module):
    """``StyleContentModel`` for one image split into row bands over the ranks of a process group.

    ``set_targets`` takes the FULL style / content images (every rank slices its own band);
    ``forward`` takes this rank's band of the optimised image (``band_of``) and returns the global
    ``(style_losses, content_losses)`` -- the same values on every rank -- so the reference-style
    runner loop works unchanged on each rank with a band-local optimiser."""

    def __init__(self, vgg_features: nn.Module, style_layers: list[int],
                 content_layers: list[int], device: torch.device, group=None) -> None:  # noqa: ANN001
        super().__init__()
        from .core_model import create_feature_blocks

        self.vgg_blocks, self.content_ids, self.style_ids = create_feature_blocks(
            vgg_features, style_layers, content_layers)
        layers: list[nn.Module] = []
        style_idx, content_idx = [], []
        for j, block in enumerate(self.vgg_blocks):
            layers.extend(block.children())
            if j in self.style_ids:
                style_idx.append(len(layers) - 1)
            if j in self.content_ids:
                content_idx.append(len(layers) - 1)
        self.engine = RowBandEngine(layers, style_idx, content_idx, device, group)
        self.style_targets: list[torch.Tensor] | None = None
        self.content_targets: list[torch.Tensor] | None = None
        self.bands: list[tuple[int, int]] | None = None

    def band_of(self, full_img: torch.Tensor) -> torch.Tensor:
        """This rank's rows of a full ``[1, 3, H, W]`` image (contiguous copy on the device)."""
        bands = plan_bands(int(full_img.shape[2]), self.engine.world)
        y0, y1 = bands[self.engine.rank]
        return full_img[:, :, y0:y1, :].to(self.engine.device).contiguous()

    def set_targets(self, style_img: torch.Tensor, content_img: torch.Tensor) -> None:
        self.bands = plan_bands(int(content_img.shape[2]), self.engine.world)
        self.engine.compute_targets(
            self.band_of(style_img), self.band_of(content_img),
            (int(style_img.shape[2]), int(style_img.shape[3])),
            (int(content_img.shape[2]), int(content_img.shape[3])))
        self.style_targets = self.engine.style_targets
        self.content_targets = self.engine.content_targets

    def forward(self, x_band: torch.Tensor):  # noqa: ANN201
        if self.style_targets is None:
            msg = "style_targets must be set before computing losses."
            raise RuntimeError(msg)
        vec = _BandLosses.apply(x_band, self.engine)
        ns = len(self.style_ids)
        return [vec[k] for k in range(ns)], [vec[ns + k] for k in range(len(self.content_ids))]

    def gather_image(self, x_band: torch.Tensor) -> torch.Tensor | None:
        """Assemble the full image on rank 0 (returns None elsewhere)."""
        eng = self.engine
        parts = [torch.empty(1, 3, y1 - y0, x_band.shape[3], device=x_band.device)
                 for (y0, y1) in self.bands] if eng.rank == 0 else None
        if eng.world == 1:
            return x_band.detach().clone()
        if eng.rank == 0:
            parts[0].copy_(x_band.detach())
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.irecv, parts[r], r, eng.group)
                                           for r in range(1, eng.world)])
            for req in reqs:
                req.wait()
            return torch.cat(parts, dim=2)
        for req in dist.batch_isend_irecv(
                [dist.P2POp(dist.isend, x_band.detach().contiguous(), 0, eng.group)]):
            req.wait()
        return None


class ShardedFusedStep:
    """Whole sharded Adam step (forward with halo exchanges, the Gram all-reduce, backward, update)
    captured into one CUDA graph per rank.  NCCL send/recv and all-reduce are captured like any
    other stream work, so a replay costs one launch instead of ~75 kernel launches plus ~27 NCCL
    group calls issued from Python -- at 8 GPUs the eager step is bound by that host work."""

    def __init__(self, model: ShardedStyleContentModel, x_band: torch.Tensor, *, lr: float,
                 style_w: float, content_w: float, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8) -> None:
        self.engine = model.engine
        self.x = x_band
        dev = x_band.device
        ns, nc = len(self.engine.style_idx), len(self.engine.content_idx)
        self.ns = ns
        self.style_w, self.content_w = float(style_w), float(content_w)
        self.grad_w = torch.tensor([self.style_w] * ns + [self.content_w] * nc, device=dev)
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.m = torch.zeros_like(x_band)
        self.v = torch.zeros_like(x_band)
        self.adam_state = torch.zeros(3, device=dev)
        self.grad = torch.zeros_like(x_band)
        self.scores = torch.zeros(3, device=dev)
        self._tmp = torch.zeros(2, device=dev)
        self.graph: torch.cuda.CUDAGraph | None = None

    def _body(self) -> None:
        eng = self.engine
        band, width = int(self.x.shape[2]), int(self.x.shape[3])
        losses = eng.forward_losses(self.x.detach())
        ops.step_scores(losses, self.ns, int(losses.numel()) - self.ns, self.style_w,
                        self.content_w, self.scores)
        self.grad.copy_(eng.backward_losses(band, width, self.grad_w))
        ops.adam_step_dev(self.x.detach(), self.grad, self.m, self.v, self.adam_state, lr=self.lr,
                          beta1=self.betas[0], beta2=self.betas[1], eps=self.eps)

    def eager_step(self) -> torch.Tensor:
        with torch.no_grad():
            self._body()
        return self.scores[2]

    def capture(self) -> None:
        """All ranks must call this together (the captured NCCL calls are collective)."""
        with torch.no_grad():
            dev = self.x.device
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._body()
            self.graph = graph

    def step(self) -> torch.Tensor:
        if self.graph is None:
            return self.eager_step()
        self.graph.replay()
        return self.scores[2]

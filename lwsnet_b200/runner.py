"""Batch runner for the LWSNet hot path: micro-batching, host<->device pipelining and batch sharding across GPUs.

Stereo pairs are independent (no cross-sample op in reference models/models.py:106-164; BatchNorm is in inference
mode), so multi-GPU execution is a contiguous batch split with no collective on the data path:
rank r of N owns pairs [r*B/N, (r+1)*B/N).  Inside a rank the shard is walked in micro-batches small enough for the
per-layer activations to stay resident in the 126 MB L2 instead of streaming through HBM.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import ops


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of `total` pairs over `world` ranks; the first total % world ranks get one extra pair."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def micro_batches(n: int, mb: int) -> List[Tuple[int, int]]:
    if mb <= 0:
        raise ValueError("micro-batch must be positive")
    return [(i, min(i + mb, n)) for i in range(0, n, mb)]


class StereoEngine:
    """Runs ``model`` (lwsnet_b200.LWSNet on one CUDA device) over batches.

    infer_device: inputs already resident in HBM  ->  preds [B,4,H,W] on the device.
    infer_host  : inputs in (pinned) host memory   ->  preds written to a (pinned) host tensor; H2D of micro-batch
                  i+1 and D2H of micro-batch i-1 overlap the compute of micro-batch i (three streams, two buffer sets).
    """

    def __init__(self, model, micro_batch: int = 2, device: Optional[torch.device] = None, use_graphs: bool = True):
        self.model = model
        self.mb = micro_batch
        self.device = device or next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("StereoEngine needs a CUDA device: lwsnet_b200 has no CPU path")
        self.use_graphs = use_graphs
        self._graphs = {}
        self._graph_launches = {}
        self._h2d = torch.cuda.Stream(self.device)
        self._d2h = torch.cuda.Stream(self.device)

    # ------------------------------------------------------------------------------------------ device-resident
    def _forward_into(self, left, right, out):
        preds = self.model(left, right)
        for s in range(4):
            out[:, s].copy_(preds[s][:, 0])

    def _graph_for(self, n, H, W, slot=0):
        key = (n, H, W, slot)
        g = self._graphs.get(key)
        if g is None:
            dev = self.device
            left = torch.zeros((n, 3, H, W), device=dev)
            right = torch.zeros((n, 3, H, W), device=dev)
            out = torch.empty((n, 4, H, W), device=dev)
            s = torch.cuda.Stream(dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                for _ in range(2):  # warm-up: cuDNN autotune, cudaFuncSetAttribute, workspace growth
                    self._forward_into(left, right, out)
            torch.cuda.current_stream(dev).wait_stream(s)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES[0]
            with torch.cuda.graph(graph):
                self._forward_into(left, right, out)
            self._graph_launches[id(graph)] = ops.LAUNCHES[0] - n0
            g = (graph, left, right, out)
            self._graphs[key] = g
        return g

    def _replay(self, graph):
        graph.replay()
        ops.LAUNCHES[0] += self._graph_launches[id(graph)]

    @torch.no_grad()
    def infer_device(self, left: torch.Tensor, right: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, _, H, W = left.shape
        if out is None:
            out = torch.empty((B, 4, H, W), device=left.device)
        for lo, hi in micro_batches(B, self.mb):
            if self.use_graphs:
                graph, gl, gr, go = self._graph_for(hi - lo, H, W)
                gl.copy_(left[lo:hi])
                gr.copy_(right[lo:hi])
                self._replay(graph)
                out[lo:hi].copy_(go)
            else:
                self._forward_into(left[lo:hi].contiguous(), right[lo:hi].contiguous(), out[lo:hi])
        return out

    # ------------------------------------------------------------------------------------------ host-resident
    @torch.no_grad()
    def infer_host(self, left: torch.Tensor, right: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """left/right: CPU tensors [B,3,H,W] (pinned for async copies).  Returns a CPU tensor [B,4,H,W]."""
        B, _, H, W = left.shape
        dev = self.device
        if out is None:
            out = torch.empty((B, 4, H, W), pin_memory=True)
        cur = torch.cuda.current_stream(dev)
        chunks = micro_batches(B, self.mb)
        slots = []
        for slot in range(2):
            if self.use_graphs:
                slots.append(self._graph_for(self.mb, H, W, slot))
            else:
                slots.append((None, torch.empty((self.mb, 3, H, W), device=dev), torch.empty((self.mb, 3, H, W), device=dev),
                              torch.empty((self.mb, 4, H, W), device=dev)))
        in_ready = [torch.cuda.Event() for _ in chunks]
        done = [torch.cuda.Event() for _ in chunks]
        out_free = [None, None]   # event: D2H that last read slot's output buffer finished
        in_free = [None, None]    # event: compute that last read slot's input buffers finished
        self._h2d.wait_stream(cur)
        self._d2h.wait_stream(cur)
        for i, (lo, hi) in enumerate(chunks):
            slot = i & 1
            graph, gl, gr, go = slots[slot]
            n = hi - lo
            with torch.cuda.stream(self._h2d):
                if in_free[slot] is not None:
                    self._h2d.wait_event(in_free[slot])
                gl[:n].copy_(left[lo:hi], non_blocking=True)
                gr[:n].copy_(right[lo:hi], non_blocking=True)
                in_ready[i].record(self._h2d)
            cur.wait_event(in_ready[i])
            if out_free[slot] is not None:
                cur.wait_event(out_free[slot])
            if graph is not None and n == self.mb:
                self._replay(graph)
            else:
                self._forward_into(gl[:n], gr[:n], go[:n])
            done[i].record(cur)
            in_free[slot] = done[i]
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(done[i])
                out[lo:hi].copy_(go[:n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._d2h)
                out_free[slot] = ev
        cur.wait_stream(self._d2h)
        cur.wait_stream(self._h2d)
        return out

    # ------------------------------------------------------------------------------------------ host-resident, uint8 I/O
    @torch.no_grad()
    def infer_host_u8(self, left: torch.Tensor, right: torch.Tensor, th: int = 368, tw: int = 1232,
                      out_gray: Optional[torch.Tensor] = None, out_color: Optional[torch.Tensor] = None,
                      color: bool = False):
        """The reference's whole inference loop body (inference.py:90-115) with only uint8 crossing PCIe.

        left/right: CPU uint8 [B,h,w,3] HWC BGR images as cv2.imread returns them (pinned for async copies).  Crop, BGR->RGB,
        ToTensor and Normalize run on the device (ops.preprocess_bgr_u8), the four stage disparities are cast to uint8 on the
        device (ops.disparity_to_u8) and, when ``color``, JET colour-mapped.  Returns (gray [B,4,th,tw] uint8 CPU,
        color [B,4,th,tw,3] uint8 CPU or None).  H2D / compute / D2H of neighbouring micro-batches overlap as in infer_host.
        """
        B, h, w, _ = left.shape
        dev = self.device
        if out_gray is None:
            out_gray = torch.empty((B, 4, th, tw), dtype=torch.uint8, pin_memory=True)
        if color and out_color is None:
            out_color = torch.empty((B, 4, th, tw, 3), dtype=torch.uint8, pin_memory=True)
        cur = torch.cuda.current_stream(dev)
        chunks = micro_batches(B, self.mb)
        slots = []
        for slot in range(2):
            key = ("u8", h, w, th, tw, color, slot)
            io = self._graphs.get(key)
            if io is None:
                io = (torch.empty((self.mb, h, w, 3), dtype=torch.uint8, device=dev),
                      torch.empty((self.mb, h, w, 3), dtype=torch.uint8, device=dev),
                      torch.empty((self.mb, 4, th, tw), dtype=torch.uint8, device=dev),
                      torch.empty((self.mb, 4, th, tw, 3), dtype=torch.uint8, device=dev) if color else None)
                self._graphs[key] = io
            if self.use_graphs:
                fwd = self._graph_for(self.mb, th, tw, slot)
            else:
                fwd = (None, torch.empty((self.mb, 3, th, tw), device=dev), torch.empty((self.mb, 3, th, tw), device=dev),
                       torch.empty((self.mb, 4, th, tw), device=dev))
            slots.append((io, fwd))
        in_ready = [torch.cuda.Event() for _ in chunks]
        done = [torch.cuda.Event() for _ in chunks]
        out_free = [None, None]
        in_free = [None, None]
        self._h2d.wait_stream(cur)
        self._d2h.wait_stream(cur)
        for i, (lo, hi) in enumerate(chunks):
            slot = i & 1
            (ul, ur, ug, uc), (graph, gl, gr, go) = slots[slot]
            n = hi - lo
            with torch.cuda.stream(self._h2d):
                if in_free[slot] is not None:
                    self._h2d.wait_event(in_free[slot])
                ul[:n].copy_(left[lo:hi], non_blocking=True)
                ur[:n].copy_(right[lo:hi], non_blocking=True)
                in_ready[i].record(self._h2d)
            cur.wait_event(in_ready[i])
            if out_free[slot] is not None:
                cur.wait_event(out_free[slot])
            ops.preprocess_bgr_u8(ul[:n], th, tw, out=gl[:n])
            ops.preprocess_bgr_u8(ur[:n], th, tw, out=gr[:n])
            if graph is not None and n == self.mb:
                self._replay(graph)
            else:
                self._forward_into(gl[:n], gr[:n], go[:n])
            ops.disparity_to_u8(go[:n], gray=True, color=color, out_gray=ug[:n], out_color=uc[:n] if color else None)
            done[i].record(cur)
            in_free[slot] = done[i]
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(done[i])
                out_gray[lo:hi].copy_(ug[:n], non_blocking=True)
                if color:
                    out_color[lo:hi].copy_(uc[:n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._d2h)
                out_free[slot] = ev
        cur.wait_stream(self._d2h)
        cur.wait_stream(self._h2d)
        return out_gray, out_color

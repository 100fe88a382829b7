"""Batch runner for the LWSNet hot path: micro-batching, host<->device pipelining and batch sharding across GPUs.

Stereo pairs are independent (no cross-sample op in reference models/models.py:106-164; BatchNorm is in inference
mode), so multi-GPU execution is a contiguous batch split with no collective on the data path:
rank r of N owns pairs [r*B/N, (r+1)*B/N).  Inside a rank the shard is walked in micro-batches of 24 pairs (one CUDA-graph replay
each): large enough that every persistent kernel fills the 148 SMs for many tiles -- measured faster than small, L2-sized
micro-batches (944 / 1165 / 1296 / 1333 / 1357 pairs/s at 1 / 2 / 4 / 8 / 24 pairs, profiles/r02_mb_sweep_r01_kernels.txt).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import ops


def bind_to_gpu_cpus(device_index: int) -> Optional[list]:
    """Pin the calling process to the CPU cores NVML reports as local to GPU `device_index` (same NUMA node / PCIe root), so that
    pinned host buffers allocated afterwards are first-touched on that node and H2D / D2H DMA does not cross the socket link.
    With 8 ranks streaming ~24 GB/s each this is the difference between scaling and not.  Returns the CPU list, or None when NVML
    or the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:  # NVML enumerates by PCI bus id, CUDA may not: go through the bus id
            pr = torch.cuda.get_device_properties(device_index)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


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


def ramp_batches(n: int, mb: int, edge: int) -> List[Tuple[int, int]]:
    """Host-resident schedule: a short first and last chunk (`edge` pairs) around `mb`-pair chunks.  The first chunk's H2D and the
    last chunk's D2H are the only copies that cannot overlap compute, so they are kept small while the bulk of the batch runs at
    the larger, more efficient micro-batch."""
    if mb <= 0 or edge <= 0:
        raise ValueError("micro-batch sizes must be positive")
    edge = min(edge, mb)
    if n <= 2 * edge:
        return micro_batches(n, edge)
    out, lo = [(0, edge)], edge
    while n - lo > edge:
        hi = lo + min(mb, n - lo - edge)
        out.append((lo, hi))
        lo = hi
    out.append((lo, n))
    return out


class StereoEngine:
    """Runs ``model`` (lwsnet_b200.LWSNet on one CUDA device) over batches.

    infer_device : inputs already resident in HBM  ->  preds [B,4,H,W] on the device.
    infer_host   : inputs in (pinned) host memory   ->  preds written to a (pinned) host tensor; H2D of micro-batch
                   i+1 and D2H of micro-batch i-1 overlap the compute of micro-batch i (three streams, two buffer sets).
    infer_host_u8: the reference's inference-loop body (inference.py:90-137): uint8 images in, uint8 / JET disparities out.
    ``stages`` selects which of the four stage outputs leave the device (the reference's directory mode keeps only the last,
    inference.py:133-137); the forward always computes all four.
    """

    def __init__(self, model, micro_batch: int = 2, device: Optional[torch.device] = None, use_graphs: bool = True,
                 host_edge: Optional[int] = None):
        self.model = model
        self.mb = micro_batch
        self.host_edge = host_edge  # infer_host: size of the first / last chunk (None: all chunks are micro_batch pairs)
        self.device = device or next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("StereoEngine needs a CUDA device: lwsnet_b200 has no CPU path")
        self.use_graphs = use_graphs
        self._graphs = {}
        self._graph_launches = {}
        self._pack_state = None
        self._h2d = torch.cuda.Stream(self.device)
        self._d2h = torch.cuda.Stream(self.device)
        # ONE stream for every warm-up and capture of this engine: scratch buffers are keyed by (device, stream) in ops.workspace,
        # so all graphs of the engine share one set (their replays are serialised on the caller's stream anyway)
        self._cap = torch.cuda.Stream(self.device)

    # ------------------------------------------------------------------------------------------ device-resident
    def _forward_into(self, left, right, out, user=None):
        """out: [4,n,1,H,W] stage-major, so every stage is one contiguous block and the model writes into it directly;
        user (optional): [n,4,H,W], the layout handed to the caller (one transposing device copy, so that the D2H copy of a
        chunk is a single contiguous transfer)."""
        self.model(left, right, out=out)
        if user is not None:
            user.copy_(out[:, :, 0].transpose(0, 1))

    def _check_weights(self):
        """Captured graphs bake in the device addresses of the BN-folded weight blobs: drop them when the model's weights (and
        with them the blobs) have changed since capture, so a replay can never read stale or freed weights."""
        state = self.model.pack_state(self.device)
        if state != self._pack_state:
            if self._graphs:
                self._graphs.clear()
                self._graph_launches.clear()
                ops.release_retired(self.device)
            self._pack_state = state

    def _graph_for(self, n, H, W, slot=0):
        key = (n, H, W, slot)
        g = self._graphs.get(key)
        if g is None:
            dev = self.device
            both = torch.zeros((2 * n, 3, H, W), device=dev)  # left and right back to back: the model stacks them without a copy
            left, right = both[:n], both[n:]
            out = torch.empty((4, n, 1, H, W), device=dev)
            user = torch.empty((n, 4, H, W), device=dev)
            s = self._cap
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                for _ in range(2):  # warm-up: cudaFuncSetAttribute, workspace growth
                    self._forward_into(left, right, out, user)
            torch.cuda.current_stream(dev).wait_stream(s)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES[0]
            with torch.cuda.graph(graph, stream=s):
                self._forward_into(left, right, out, user)
            self._graph_launches[id(graph)] = ops.LAUNCHES[0] - n0
            g = (graph, left, right, out, user)
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
        self._check_weights()
        for lo, hi in micro_batches(B, self.mb):
            if self.use_graphs:
                graph, gl, gr, go, gu = self._graph_for(hi - lo, H, W)
                gl.copy_(left[lo:hi])
                gr.copy_(right[lo:hi])
                self._replay(graph)
                out[lo:hi].copy_(gu)
            else:
                go = torch.empty((4, hi - lo, 1, H, W), device=left.device)
                self._forward_into(left[lo:hi].contiguous(), right[lo:hi].contiguous(), go, out[lo:hi])
        return out

    # ------------------------------------------------------------------------------------------ host-resident
    def _host_chunks(self, B):
        return ramp_batches(B, self.mb, self.host_edge) if self.host_edge else micro_batches(B, self.mb)

    def _host_slot(self, n, H, W, slot):
        """(graph | None, left, right, out [4,n,1,H,W], user [n,4,H,W]) device buffers for an n-pair chunk in pipeline slot 0/1."""
        if self.use_graphs:
            return self._graph_for(n, H, W, slot)
        key = ("eager", n, H, W, slot)
        buf = self._graphs.get(key)
        if buf is None:
            dev = self.device
            buf = (None, torch.empty((n, 3, H, W), device=dev), torch.empty((n, 3, H, W), device=dev),
                   torch.empty((4, n, 1, H, W), device=dev), torch.empty((n, 4, H, W), device=dev))
            self._graphs[key] = buf
        return buf

    def _select_stages(self, go, gu, stages):
        """Contiguous device tensor [n,S,H,W] holding the selected stages: all four = the user-layout buffer, a single stage = its
        block of the stage-major buffer (no copy), anything else = a small gathered staging tensor."""
        if len(stages) == 4:
            return gu
        if len(stages) == 1:
            return go[stages[0]]
        key = ("sel", id(gu), stages)
        buf = self._graphs.get(key)
        if buf is None:
            buf = torch.empty((gu.shape[0], len(stages)) + tuple(gu.shape[2:]), device=gu.device)
            self._graphs[key] = buf
        for j, st in enumerate(stages):
            buf[:, j].copy_(go[st, :, 0])
        return buf

    @torch.no_grad()
    def infer_host(self, left: torch.Tensor, right: torch.Tensor, out: Optional[torch.Tensor] = None,
                   stages=(0, 1, 2, 3)) -> torch.Tensor:
        """left/right: CPU tensors [B,3,H,W] (pinned for async copies).  Returns a CPU tensor [B,len(stages),H,W] with the selected
        stage disparities (fp32)."""
        B, _, H, W = left.shape
        dev = self.device
        stages = tuple(stages)
        if not stages or any(s not in (0, 1, 2, 3) for s in stages):
            raise ValueError("stages must be a non-empty subset of (0, 1, 2, 3)")
        if out is None:
            out = torch.empty((B, len(stages), H, W), pin_memory=True)
        if tuple(out.shape) != (B, len(stages), H, W):
            raise ValueError("out must be [B,len(stages),H,W]")
        cur = torch.cuda.current_stream(dev)
        chunks = self._host_chunks(B)
        self._check_weights()
        bufs = [self._host_slot(hi - lo, H, W, i & 1) for i, (lo, hi) in enumerate(chunks)]  # before the timed copies start
        in_ready = [torch.cuda.Event() for _ in chunks]
        done = [torch.cuda.Event() for _ in chunks]
        out_free = {}  # buffer set -> event: the D2H that last read its output buffer finished
        in_free = {}   # buffer set -> event: the compute that last read its input buffers finished
        self._h2d.wait_stream(cur)
        self._d2h.wait_stream(cur)
        for i, (lo, hi) in enumerate(chunks):
            graph, gl, gr, go, gu = bufs[i]
            key = id(gl)
            with torch.cuda.stream(self._h2d):
                if key in in_free:
                    self._h2d.wait_event(in_free[key])
                gl.copy_(left[lo:hi], non_blocking=True)
                gr.copy_(right[lo:hi], non_blocking=True)
                in_ready[i].record(self._h2d)
            cur.wait_event(in_ready[i])
            if key in out_free:
                cur.wait_event(out_free[key])
            if graph is not None:
                self._replay(graph)
            else:
                self._forward_into(gl, gr, go, gu)
            src = self._select_stages(go, gu, stages)  # contiguous [n,S,H,W] on the device: one contiguous D2H transfer
            done[i].record(cur)
            in_free[key] = done[i]
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(done[i])
                out[lo:hi].copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._d2h)
                out_free[key] = ev
        cur.wait_stream(self._d2h)
        cur.wait_stream(self._h2d)
        return out

    # ------------------------------------------------------------------------------------------ host-resident, uint8 I/O
    @torch.no_grad()
    def infer_host_u8(self, left: torch.Tensor, right: torch.Tensor, th: int = 368, tw: int = 1232,
                      out_gray: Optional[torch.Tensor] = None, out_color: Optional[torch.Tensor] = None,
                      color: bool = False, stages=(0, 1, 2, 3), gray: bool = True):
        """The reference's whole inference loop body (inference.py:90-115) with only uint8 crossing PCIe.

        left/right: CPU uint8 [B,h,w,3] HWC BGR images as cv2.imread returns them (pinned for async copies).  Crop, BGR->RGB,
        ToTensor and Normalize run on the device (ops.preprocess_bgr_u8), the four stage disparities are cast to uint8 on the
        device (ops.disparity_to_u8) and, when ``color``, JET colour-mapped.  Returns (gray [B,S,th,tw] uint8 CPU or None,
        color [B,S,th,tw,3] uint8 CPU or None) for the S selected ``stages`` (``stages=(3,), gray=False, color=True`` is the
        reference's directory mode: one colour image per pair, inference.py:133-137).  H2D / compute / D2H of neighbouring
        micro-batches overlap as in infer_host.
        """
        B, h, w, _ = left.shape
        dev = self.device
        stages = tuple(stages)
        S = len(stages)
        if not stages or any(s not in (0, 1, 2, 3) for s in stages) or list(stages) != sorted(set(stages)):
            raise ValueError("stages must be a non-empty ascending subset of (0, 1, 2, 3)")
        if not (gray or color):
            raise ValueError("ask for gray and/or color")
        if gray and out_gray is None:
            out_gray = torch.empty((B, S, th, tw), dtype=torch.uint8, pin_memory=True)
        if color and out_color is None:
            out_color = torch.empty((B, S, th, tw, 3), dtype=torch.uint8, pin_memory=True)
        cur = torch.cuda.current_stream(dev)
        chunks = self._host_chunks(B)
        self._check_weights()
        bufs = []
        for i, (lo, hi) in enumerate(chunks):
            n, slot = hi - lo, i & 1
            key = ("u8", n, h, w, th, tw, gray, color, S, slot)
            io = self._graphs.get(key)
            if io is None:
                io = (torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev),
                      torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev),
                      torch.empty((n, S, th, tw), dtype=torch.uint8, device=dev) if gray else None,
                      torch.empty((n, S, th, tw, 3), dtype=torch.uint8, device=dev) if color else None)
                self._graphs[key] = io
            bufs.append((io, self._host_slot(n, th, tw, slot)))
        in_ready = [torch.cuda.Event() for _ in chunks]
        done = [torch.cuda.Event() for _ in chunks]
        out_free, in_free = {}, {}
        self._h2d.wait_stream(cur)
        self._d2h.wait_stream(cur)
        for i, (lo, hi) in enumerate(chunks):
            (ul, ur, ug, uc), (graph, gl, gr, go, gu) = bufs[i]
            key = id(ul)
            with torch.cuda.stream(self._h2d):
                if key in in_free:
                    self._h2d.wait_event(in_free[key])
                ul.copy_(left[lo:hi], non_blocking=True)
                ur.copy_(right[lo:hi], non_blocking=True)
                in_ready[i].record(self._h2d)
            cur.wait_event(in_ready[i])
            if key in out_free:
                cur.wait_event(out_free[key])
            ops.preprocess_bgr_u8(ul, th, tw, out=gl)
            ops.preprocess_bgr_u8(ur, th, tw, out=gr)
            if graph is not None:
                self._replay(graph)
            else:
                self._forward_into(gl, gr, go, gu)
            # one conversion launch over the selected stages in the caller's [n,S,...] layout, then one contiguous D2H per output
            ops.disparity_to_u8(self._select_stages(go, gu, stages), gray=gray, color=color, out_gray=ug, out_color=uc)
            done[i].record(cur)
            in_free[key] = done[i]
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(done[i])
                if gray:
                    out_gray[lo:hi].copy_(ug, non_blocking=True)
                if color:
                    out_color[lo:hi].copy_(uc, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._d2h)
                out_free[key] = ev
        cur.wait_stream(self._d2h)
        cur.wait_stream(self._h2d)
        return (out_gray if gray else None), (out_color if color else None)

"""Data-parallel plumbing: one process per GPU, clips sharded by rank, one gradient all-reduce (NCCL over NVLink on
the B200 box; gloo in CPU tests).  Replaces the reference's nn.DataParallel (R/main.py:155), which re-broadcasts all
parameters every step and reduces gradients onto GPU 0.  The fusion path itself has no collective: attention is
within one sample's tokens and BatchNorm statistics are per replica, as in the reference."""
from __future__ import annotations

import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend: str = "nccl") -> Tuple[int, int, int]:
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of `n_items` clips owned by `rank`; sizes differ by at most one."""
    if n_items < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class GradBucket:
    """Flat fp32 bucket over a fixed parameter list: gradients are packed, all-reduced once, averaged, unpacked.

    On CUDA the bucket sits behind a zero-initialised signal pad in one device allocation, so that `enable_p2p` can
    expose it to the other ranks through CUDA IPC; the average is then ONE kernel over NVLink / NVSwitch peer memory
    (`glf_p2p_allreduce`, csrc/glf_p2p.cu) instead of an NCCL call, and can be captured in the step's CUDA graph."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self._p2p = None
        self._raw = None
        if dev.type == "cuda":
            from . import _lib as L
            self._sig_bytes = (int(L.load().glf_p2p_signal_bytes(8)) + 255) // 256 * 256
            self._n_pad = (self.numel + 3) // 4 * 4
            self._raw = torch.zeros(self._sig_bytes + 4 * self._n_pad, dtype=torch.uint8, device=dev)
            self.flat = self._raw[self._sig_bytes:].view(torch.float32)[:self.numel]
        else:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        # views of the flat bucket, one per parameter: pack / unpack are single multi-tensor copies
        self.views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            n = p.numel()
            self.views.append(self.flat[off:off + n].view_as(p))
            off += n

    def bind(self, modules) -> None:
        """Zero-copy mode: make the TPAVI blocks in `modules` write their parameter gradients straight into this
        bucket (module._grad_out: kernel-side name -> bucket view), so that after backward `p.grad` aliases the flat
        buffer and the all-reduce needs neither pack nor unpack.

        The kernels assign, they never accumulate, so the views are handed out only to the first backward node of a
        step that touches a module and only while its parameters have no .grad yet (tpavi.claim_grad_out): a module
        that runs twice before one backward (the reference's cycle pass, R/main.py:209/:221), gradient accumulation
        and zero_grad(set_to_none=False) all fall back to fresh gradient tensors that autograd sums, and
        allreduce_mean() then packs / unpacks.  The zero-copy path therefore needs zero_grad(set_to_none=True)
        (PyTorch's default) and one use of each module per backward; everything else stays correct, just not
        zero-copy.  Only the parameters a block really differentiates (module._plist()) belong in the bucket:
        align_channel.* never receives a gradient on the fusion path."""
        index = {id(p): v for p, v in zip(self.params, self.views)}
        for mod in modules:
            table = dict(mod.named_parameters())
            out = {}
            for name in mod._plist_names():
                p = table[name]
                if id(p) in index and p.dtype == torch.float32:
                    out[mod._grad_key(name)] = index[id(p)]
            mod._grad_out = out

    def aliased(self) -> bool:
        return all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() and p.grad.dtype == torch.float32
                   for p, v in zip(self.params, self.views))

    def pack(self) -> torch.Tensor:
        if all(p.grad is not None and p.grad.dtype == torch.float32 for p in self.params):
            torch._foreach_copy_(self.views, [p.grad for p in self.params])
        else:
            for v, p in zip(self.views, self.params):
                if p.grad is None:
                    v.zero_()
                else:
                    v.copy_(p.grad)
        return self.flat

    def unpack(self) -> None:
        if all(p.grad is not None and p.grad.dtype == torch.float32 for p in self.params):
            torch._foreach_copy_([p.grad for p in self.params], self.views)
            return
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                # no rank-local gradient: leave it None (the reference optimizer skips such parameters; giving them
                # a zero .grad would subject them to weight decay / momentum)
                continue
            p.grad.copy_(v)

    def enable_p2p(self, group=None) -> bool:
        """Exchange CUDA IPC handles of the bucket allocation with the other ranks of `group` (one node, <= 8 GPUs).
        Returns True when every rank can reach every bucket; otherwise the NCCL path stays in use."""
        if self._raw is None or not (dist.is_initialized() and dist.get_world_size(group) > 1):
            return False
        import ctypes as C
        from . import _lib as L
        lib = L.load()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        ok = 2 <= world <= 8 and self._n_pad <= int(lib.glf_p2p_max_floats())
        mine = None
        if ok:
            handle = C.create_string_buffer(64)
            off = C.c_uint64(0)
            with torch.cuda.device(self._raw.device):
                rc = lib.glf_p2p_export(L.ptr(self._raw), handle, C.byref(off))
            ok = rc == 0
            mine = (handle.raw, int(off.value))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine if ok else None, group=group)
        if any(g is None for g in gathered):
            return False
        bases = []
        opened_all = True
        self._opened = []          # (peer base pointer, offset) of every IPC mapping this bucket holds
        for r, (h, o) in enumerate(gathered):
            if r == rank:
                bases.append(self._raw.data_ptr())
                continue
            out = C.c_void_p(0)
            with torch.cuda.device(self._raw.device):
                rc = lib.glf_p2p_open(h, C.c_uint64(o), C.byref(out))
            if rc != 0 or not out.value:
                opened_all = False
                bases.append(0)
            else:
                bases.append(int(out.value))
                self._opened.append((int(out.value), int(o)))
        flags = [None] * world
        dist.all_gather_object(flags, opened_all, group=group)
        if not all(flags):
            self.close_p2p()       # some rank could not map a peer: drop the mappings that did open
            return False
        sigs = (C.c_void_p * world)(*bases)
        bufs = (C.c_void_p * world)(*[b + self._sig_bytes for b in bases])
        self._p2p = (bufs, sigs, rank, world)
        dist.barrier(group=group)
        return True

    def close_p2p(self) -> None:
        """Unmap the peers' buckets (cudaIpcCloseMemHandle through glf_p2p_close) and return to the NCCL path."""
        opened, self._opened = getattr(self, "_opened", []), []
        self._p2p = None
        if not opened:
            return
        import ctypes as C
        from . import _lib as L
        lib = L.load()
        for ptr, off in opened:
            with torch.cuda.device(self._raw.device):
                lib.glf_p2p_close(C.c_void_p(ptr), C.c_uint64(off))

    def __del__(self):
        try:
            self.close_p2p()
        except Exception:      # interpreter shutdown: the driver reclaims the mappings with the context
            pass

    @property
    def p2p_enabled(self) -> bool:
        return self._p2p is not None

    def overlap_with(self, fusion) -> None:
        """Overlap the gradient exchange with the tail of the backward pass: `fusion` (a GlobalLocalFusion whose two
        blocks are bound to this bucket) calls back as soon as both blocks' parameter gradients are final, i.e. before
        its gate backward; the peer-memory all-reduce is launched there on a side stream and joined by
        `allreduce_mean()`, which then has nothing left to do.  Needs the zero-copy peer path (bind + enable_p2p); in
        every other configuration the callback does nothing and `allreduce_mean()` runs the exchange as before."""
        self._ar_stream = None
        self._ar_done = False

        def ready():
            if self._p2p is None:
                return
            cur = torch.cuda.current_stream(self._raw.device)
            if self._ar_stream is None:
                self._ar_stream = torch.cuda.Stream(device=self._raw.device)
            self._ar_stream.wait_stream(cur)
            with torch.cuda.stream(self._ar_stream):
                self._launch_p2p()
            self._ar_done = True
        fusion.on_weight_grads_ready = ready

    def _launch_p2p(self) -> None:
        import ctypes as C
        from . import _lib as L
        bufs, sigs, rank, world = self._p2p
        with torch.cuda.device(self._raw.device):
            L.check(L.load().glf_p2p_allreduce(bufs, sigs, rank, world, self._n_pad, 1.0 / world,
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def allreduce_mean(self, group=None) -> None:
        if getattr(self, "_ar_done", False):
            # the exchange was launched from the backward pass (overlap_with): join it
            self._ar_done = False
            torch.cuda.current_stream(self._raw.device).wait_stream(self._ar_stream)
            if self.aliased():
                return
            # (some gradient did not land in the bucket after all: fall through to the packed exchange)
        zero_copy = self.aliased()
        if not zero_copy:
            self.pack()
        if self._p2p is not None:
            self._launch_p2p()
        elif dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)     # one kernel: sum and 1/world
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))
        if not zero_copy:
            self.unpack()


def broadcast_buffers(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Mirror nn.DataParallel's behaviour of keeping replica 0's BatchNorm running statistics."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    for b in module.buffers():
        dist.broadcast(b, src=src, group=group)


def max_over_ranks(value_ms: float, device=None, group=None) -> float:
    """Step time of a data-parallel job = the slowest rank."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return value_ms
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())

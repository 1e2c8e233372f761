"""Gradient exchange of the data-parallel step (one all-reduce of the flat gradient buffer per step).

The exchange sits between two launches of the step kernel (backward | exchange | clip + Adam), so it is pure latency on
the critical path: 17 MB at the headline shape. NCCL's ring / tree all-reduce costs ~50 us at 2 and ~100 us at 8 B200
for that size; on an NVSwitch box the same sum can be taken INSIDE the switch (NVLS multimem: every rank reduces 1/R of
the buffer with multimem.ld_reduce and broadcasts it with multimem.st), which moves 2/R of the bytes per GPU. torch
exposes that as ``torch.ops.symm_mem.multimem_all_reduce_`` on buffers allocated in NVLink symmetric memory, so the
engine is pointed at such a buffer (``jb_set_grad_buffer``) and writes its gradients there in the first place.

Round 2, second half: the exchange moved INTO the step kernel (``mode == 'kernel'``, jb_set_exchange): every rank's
gradient buffer and a small scratch block live in symmetric memory mapped into all peers; between its WGRAD and ADAM phases
the persistent kernel signals its peers, reduces its 1/R slice with peer loads, writes the sum into every rank's buffer
with peer stores, and delivers the clip-norm partials the same way. A data-parallel run is then ONE launch for any number
of optimizer steps, like the single-GPU run, and ``all_reduce()`` is a no-op.

Order of preference: kernel (peer memory) -> multimem / two-shot all-reduce (4+ ranks) -> NCCL. ``JB_DP_ALLREDUCE``
forces one of ``kernel | multimem | two_shot | nccl``.
"""
import os


class GradExchange:
    def __init__(self, eng, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.eng = eng
        self.group = group if group is not None else dist.group.WORLD
        self.mode = 'nccl'
        self.buf = None
        want = os.environ.get('JB_DP_ALLREDUCE', 'auto')
        backend = dist.get_backend(self.group)
        if backend == 'nccl' and want in ('auto', 'kernel', 'multimem', 'two_shot'):
            try:
                self._setup_symm(want)
            except Exception as ex:   # no multicast / symmetric memory on this box: NCCL
                if want != 'auto':
                    raise
                self.why = f'{type(ex).__name__}: {ex}'
                self.buf = None
                eng.set_exchange(0, 1, None, None)
                eng.set_grad_buffer(None)
                self.mode = 'nccl'
        if self.buf is None:
            self.buf = eng.grad_tensor()

    def _setup_symm(self, want):
        import torch.distributed._symmetric_memory as symm_mem
        torch = self.torch
        world, rank = self.dist.get_world_size(self.group), self.dist.get_rank(self.group)
        if want in ('auto', 'kernel') and world <= 8:
            _, n = self.eng.grad_buffer()
            n_alloc = (n + 1023) // 1024 * 1024
            dev = torch.device('cuda', self.eng.device)
            buf = symm_mem.empty(n_alloc, dtype=torch.float32, device=dev)
            scratch = symm_mem.empty((self.eng.exchange_scratch_bytes() + 1023) // 1024 * 256, dtype=torch.int32, device=dev)
            buf.zero_(); scratch.zero_()
            hb = symm_mem.rendezvous(buf, self.group)
            hs = symm_mem.rendezvous(scratch, self.group)
            torch.cuda.synchronize()
            self.dist.barrier(self.group)
            self.eng.set_grad_buffer(buf)
            # JB_XCHG_MC=1: the sweep through the switch (NVLS multimem.ld_reduce / multimem.st). Measured on B200: 71 vs 40 us
            # per exchange at 2 ranks, 76 vs 72 us at 8 ranks -- never faster than the peer loads / stores, whose sum is
            # also taken in rank order (bit-identical to a sequential sum); so it is opt-in.
            mc = int(getattr(hb, 'multicast_ptr', 0) or 0)
            use_mc = mc != 0 and os.environ.get('JB_XCHG_MC', '0') == '1'
            self.eng.set_exchange(rank, world, list(hb.buffer_ptrs), list(hs.buffer_ptrs), mc if use_mc else 0)
            self.buf, self.scratch, self.handle, self.handle_s = buf, scratch, hb, hs
            self.mode = 'kernel'
            self.multicast = use_mc
            return
        # measured on B200: multimem 360 us/step vs two-shot 381 at 8 ranks, but 466 vs NCCL's 366 at 2 ranks
        if want == 'auto' and world < 4:
            raise RuntimeError('NCCL is faster below 4 ranks')
        _, n = self.eng.grad_buffer()
        n_alloc = (n + 1023) // 1024 * 1024
        buf = symm_mem.empty(n_alloc, dtype=torch.float32, device=torch.device('cuda', self.eng.device))
        hdl = symm_mem.rendezvous(buf, self.group)
        self.group_name = self.group.group_name
        buf.zero_()
        self.eng.set_grad_buffer(buf)
        self.buf = buf
        self.handle = hdl
        modes = ['multimem', 'two_shot'] if want == 'auto' else [want]
        err = None
        for m in modes:
            try:
                self.mode = m
                self.all_reduce()          # a trial run on zeros also warms the kernels up
                torch.cuda.synchronize()
                return
            except Exception as ex:
                err = ex
        raise err

    def all_reduce(self):
        """Sum of the gradient buffer over the ranks, in place, on torch's current stream (no-op when the step kernel
        exchanges the gradients itself)."""
        if self.mode == 'kernel':
            return
        if self.mode == 'multimem':
            self.torch.ops.symm_mem.multimem_all_reduce_(self.buf, 'sum', self.group_name)
        elif self.mode == 'two_shot':
            self.torch.ops.symm_mem.two_shot_all_reduce_(self.buf, 'sum', self.group_name)
        else:
            self.dist.all_reduce(self.buf, group=self.group)

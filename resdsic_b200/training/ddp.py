"""Data-parallel gradient averaging for the rate-distortion step (BASELINE config 4; the reference wraps the model
in nn.DataParallel, train.py:45-52,167-168 -- one process, replicas on threads, gradients reduced on GPU 0).

Here: one process per GPU (`torch.distributed`, NCCL over NVLink / NVSwitch), parameters replicated, and a bucketed
gradient all-reduce that OVERLAPS the backward pass: parameters are assigned to ~25 MB buckets in reverse
registration order (the order their gradients become ready); a post-accumulate-grad hook counts a bucket's
gradients in, and when the last one lands the bucket is flattened and all-reduced asynchronously on a side stream
while autograd keeps walking the graph.  `finish()` waits for the outstanding buckets and writes the averaged
gradients back.  The collective is the ONLY cross-rank traffic of the step (301 MB of fp32 gradients for WACNN).
"""
import torch
import torch.distributed as dist


class GradBucketReducer:
    def __init__(self, params, bucket_bytes=25 * 1024 * 1024, process_group=None, average=True):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):  # gradients arrive roughly in reverse registration order
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._pending = [0] * len(self.buckets)
        self._work = []
        self._stream = None
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.reset()

    def reset(self):
        self._pending = [len(b) for b in self.buckets]
        self._work = []

    def _comm_stream(self, device):
        if device.type != "cuda":
            return None
        if self._stream is None:
            self._stream = torch.cuda.Stream(device)
        return self._stream

    def _on_grad(self, p):
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            self._launch(i)

    def _launch(self, i):
        bucket = self.buckets[i]
        dev = bucket[0].device
        st = self._comm_stream(dev)
        if st is not None:
            st.wait_stream(torch.cuda.current_stream(dev))  # the bucket's gradients are complete on the compute stream
            ctx = torch.cuda.stream(st)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            flat = torch.cat([p.grad.reshape(-1) for p in bucket])
            if self.average:
                flat.div_(self.world)
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True) if self.world > 1 else None
        self._work.append((i, flat, work))

    def finish(self):
        """Wait for every bucket, scatter the reduced gradients back into p.grad.  Call after backward()."""
        for i, n in enumerate(self._pending):  # parameters that received no gradient this step
            if n and n < len(self.buckets[i]):
                raise RuntimeError("GradBucketReducer: a bucket is partially filled (some parameters got no gradient)")
        for i, flat, work in self._work:
            if work is not None:
                work.wait()
            dev = flat.device
            st = self._comm_stream(dev)
            if st is not None:
                torch.cuda.current_stream(dev).wait_stream(st)
            sizes = [p.numel() for p in self.buckets[i]]
            views = [c.view_as(p.grad) for c, p in zip(flat.split(sizes), self.buckets[i])]
            torch._foreach_copy_([p.grad for p in self.buckets[i]], views)  # one multi-tensor copy per bucket, not one per parameter
        n_buckets = len(self._work)
        self.reset()
        return n_buckets

    def remove(self):
        for h in self._hooks:
            h.remove()

"""Throughput inference from HOST memory: overlap the pinned-host -> device copy of batch i+1 and the
device -> pinned-host copy of the results of batch i-1 with the forward pass of batch i.

The model's forward writes into static plan buffers (CUDA-graph replay), so each result set is snapshotted
device-to-device on the compute stream before the next replay overwrites it; the snapshot then drains to the
host on a separate stream.  Every byte still crosses PCIe once per batch in each direction -- the copies are
only taken off the critical path (what the reference's DataLoader + `.to(device)` + `.cpu()` loop,
training/step.py:34 / eval_model/__main__.py:133-147, leaves serialised)."""
import torch

from .. import _lib


class ForwardPipeline:
    """`compact=False` (default): host batches are fp32 NCHW images in [0,1]; per batch the fp32 `x_hat` and both
    likelihood tensors come back, as the reference's `forward` returns them.
    `compact=True`: host batches are uint8 NCHW images (what an image file holds); they are converted on the device
    (`/ 255`, torchvision ToTensor's arithmetic), and per batch the device returns `x_hat` as uint8
    (`round(clamp(x_hat, 0, 1) * 255)`) and `bits[b]` (fp64: the image's rate, -(sum log2 of both likelihood tensors))
    -- 1 byte per sample and 8 bytes per image instead of 4 bytes per sample plus the likelihood tensors
    (csrc/image_io.cu).  `on_result(i, x_hat_u8, bits)` in that mode."""

    def __init__(self, model, example_host_batch, depth=2, compact=False):
        if not example_host_batch.is_pinned():
            raise ValueError("host batches must live in pinned memory (tensor.pin_memory())")
        if compact != (example_host_batch.dtype == torch.uint8):
            raise ValueError("compact=True takes uint8 image batches, compact=False float32 batches in [0,1]")
        self.model = model
        self.device = next(model.parameters()).device
        self.depth = depth
        self.compact = compact
        with torch.cuda.device(self.device):
            self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
            self.stage = [torch.empty_like(example_host_batch, device=self.device) for _ in range(depth)]
            if compact:
                B = example_host_batch.shape[0]
                self.x_f32 = torch.empty(example_host_batch.shape, dtype=torch.float32, device=self.device)
                out = self._forward(self.x_f32.zero_())
                self.work = torch.empty(_lib.lib().rdsic_rate_workspace_doubles(B), dtype=torch.float64, device=self.device)
                mk = lambda host: [torch.empty(example_host_batch.shape, dtype=torch.uint8, device="cpu" if host else self.device),
                                   torch.empty(B, dtype=torch.float64, device="cpu" if host else self.device)]
                self.snap = [mk(False) for _ in range(depth)]
                self.host = [[t.pin_memory() for t in mk(True)] for _ in range(depth)]
                self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.host[0])
            else:
                out = self._forward(self.stage[0].zero_())
                self.snap = [self._like(out) for _ in range(depth)]
                self.host = [self._like(out, host=True) for _ in range(depth)]
                self.d2h_bytes = sum(t.numel() * t.element_size() for t in self._flat(out))
        self.h2d_bytes = example_host_batch.numel() * example_host_batch.element_size()

    def _compact_forward(self, stage_u8, snap):
        """uint8 batch -> fp32 on the device -> forward -> uint8 x_hat + per-image bits into `snap`."""
        L = _lib.lib()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(L.rdsic_image_u8_to_f32(stage_u8.data_ptr(), self.x_f32.data_ptr(), stage_u8.numel(), st), "rdsic_image_u8_to_f32")
        out = self._forward(self.x_f32)
        xh, ly, lz = self._flat(out)
        B = xh.shape[0]
        _lib.check(L.rdsic_image_f32_to_u8(xh.data_ptr(), snap[0].data_ptr(), xh.numel(), st), "rdsic_image_f32_to_u8")
        _lib.check(L.rdsic_rate_per_image(ly.data_ptr(), ly.numel() // B, lz.data_ptr(), lz.numel() // B, B, self.work.data_ptr(),
                                          snap[1].data_ptr(), st), "rdsic_rate_per_image")

    def _forward(self, x):
        """The model's static output buffers (no per-call clones): each result set is snapshotted right below."""
        prev, self.model.static_outputs = self.model.static_outputs, True
        try:
            return self.model(x)
        finally:
            self.model.static_outputs = prev

    @staticmethod
    def _flat(out):
        return [out["x_hat"], out["likelihoods"]["y"], out["likelihoods"]["z"]]

    def _like(self, out, host=False):
        mk = (lambda t: torch.empty_like(t, device="cpu").pin_memory()) if host else torch.empty_like
        return [mk(t) for t in self._flat(out)]

    @torch.no_grad()
    def run(self, host_batches, on_result=None):
        """Forward every pinned host batch; results land in pinned host buffers (`self.host[i % depth]`,
        handed to `on_result(i, x_hat, lik_y, lik_z)` once complete).  Returns the number of batches."""
        dev, D = self.device, self.depth
        cur = torch.cuda.current_stream(dev)
        ev_in = [torch.cuda.Event() for _ in range(D)]
        ev_comp = [torch.cuda.Event() for _ in range(D)]
        ev_out = [torch.cuda.Event() for _ in range(D)]
        n = 0
        for i, hb in enumerate(host_batches):
            k = i % D
            with torch.cuda.stream(self.s_in):
                if i >= D:
                    self.s_in.wait_event(ev_comp[k])       # forward i-D has consumed this staging buffer
                self.stage[k].copy_(hb, non_blocking=True)
                ev_in[k].record(self.s_in)
            cur.wait_event(ev_in[k])
            if i >= D:
                cur.wait_event(ev_out[k])                  # snapshot k has drained to the host
                if on_result is not None:
                    ev_out[k].synchronize()
                    on_result(i - D, *self.host[k])
            if self.compact:
                self._compact_forward(self.stage[k], self.snap[k])
            else:
                out = self._forward(self.stage[k])
                for dst, src in zip(self.snap[k], self._flat(out)):
                    dst.copy_(src, non_blocking=True)
            ev_comp[k].record(cur)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_comp[k])
                for dst, src in zip(self.host[k], self.snap[k]):
                    dst.copy_(src, non_blocking=True)
                ev_out[k].record(self.s_out)
            n += 1
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_in)
        if on_result is not None:
            torch.cuda.current_stream(dev).synchronize()
            for i in range(max(0, n - D), n):
                on_result(i, *self.host[i % D])
        return n

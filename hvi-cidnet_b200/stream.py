"""Streamed inference driver: replaces the reference's synchronous batch-1 loops
(eval_SID_blur.py:25-40, eval.py:56-75: `.cuda()` -> model -> `.cpu()` per image, each a full
host<->device round trip on the default stream) by a 3-stage pipeline

    copy-in stream  : pinned host batch  -> device ring slot          (H2D)
    compute stream  : CIDNet.forward(slot) -> device output ring slot (C ABI / CUDA-graph replay)
    copy-out stream : device output slot -> pinned host ring slot     (D2H)

so the PCIe copies of step i+1 / i-1 overlap the kernels of step i.  Results are yielded in input
order, `depth - 1` steps behind the submissions.  All arithmetic is the unchanged CIDNet.forward.

`run_to_sink` adds the fourth stage the reference also runs inline (eval.py:71-75: ToPILImage + `save` of
every result inside the loop): results are handed, in order, to a caller-supplied sink on worker threads
(PNG encoding, file or socket writes) while the GPU works on the next images; a pinned result buffer is
recycled only after its sink call has returned."""
import queue
import threading

import torch


class StreamedCIDNet:
    def __init__(self, model, depth=3):
        if depth < 2:
            raise ValueError("depth must be >= 2")
        self.model, self.depth = model, int(depth)
        self._shape = None

    def _setup(self, shape, dev, dtype=torch.float32):
        if self._shape == (tuple(shape), dev, dtype):
            return
        d = self.depth
        self.dev = dev
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.x = [torch.empty(shape, device=dev, dtype=dtype) for _ in range(d)]
        self.y = [torch.empty(shape, device=dev, dtype=dtype) for _ in range(d)]
        self.hy = [torch.empty(shape, dtype=dtype).pin_memory() for _ in range(d)]
        self.in_ready = [torch.cuda.Event() for _ in range(d)]
        self.in_free = [torch.cuda.Event() for _ in range(d)]      # compute has consumed x[slot]
        self.out_ready = [torch.cuda.Event() for _ in range(d)]
        self.out_done = [torch.cuda.Event() for _ in range(d)]     # D2H of y[slot] -> hy[slot] finished
        self._shape = (tuple(shape), dev, dtype)

    def run_u8(self, host_batches, gamma=1.0):
        """The same pipeline for 8-bit images: host_batches of uint8 [B,h,w,3] (HWC, any h, w -- what an image decoder
        yields) -> enhanced uint8 [B,h,w,3] pinned host tensors.  6 instead of 24 bytes per pixel cross PCIe and the
        ToTensor / reflect-pad / gamma / clamp / crop / quantise steps of the reference's loop (eval.py:56-73) run inside
        the first and last kernel of the forward (CIDNet.enhance_u8 -> cidnet_forward_u8)."""
        return self.run(host_batches, _u8_gamma=float(gamma))

    def run(self, host_batches, _u8_gamma=None):
        """host_batches: iterable of fp32 [B,3,H,W] host tensors of ONE shape (pinned memory for real
        overlap).  Yields the enhanced batches as pinned host tensors, in order; a yielded tensor is
        valid until `depth` further results have been produced (copy it if it must live longer)."""
        for _, y in self._pipeline(host_batches, _u8_gamma, None):
            yield y

    def run_to_sink(self, host_batches, sink, u8_gamma=None, workers=1):
        """Asynchronous result sink: `sink(index, host_tensor)` is called once per input batch, on one of `workers`
        threads, with the pinned result (fp32 [B,3,H,W], or uint8 [B,h,w,3] when `u8_gamma` is given); the tensor is
        only valid during the call.  Calls START in input order (with one worker they are strictly sequential).  The
        pipeline keeps feeding the GPU while sinks run and blocks only when all `depth` result buffers are still
        being consumed.  The first exception raised by a sink is re-raised here after the pipeline has drained.
        Returns the number of batches processed."""
        if workers < 1:
            raise ValueError("workers must be >= 1")
        jobs, errors = queue.Queue(), []
        busy = [None] * self.depth                                 # slot -> threading.Event of the sink call using hy[slot]

        def work():
            while True:
                item = jobs.get()
                if item is None:
                    return
                idx, t, ev = item
                try:
                    if not errors:
                        sink(idx, t)
                except BaseException as e:                         # noqa: BLE001 -- re-raised on the caller's thread
                    errors.append(e)
                finally:
                    ev.set()

        threads = [threading.Thread(target=work, daemon=True) for _ in range(int(workers))]
        for t in threads:
            t.start()

        def before_reuse(slot):
            if busy[slot] is not None:
                busy[slot].wait()
                busy[slot] = None

        n = 0
        try:
            for slot, y in self._pipeline(host_batches, u8_gamma, before_reuse):
                ev = threading.Event()
                busy[slot] = ev
                jobs.put((n, y, ev))
                n += 1
                if errors:
                    break
        finally:
            for _ in threads:
                jobs.put(None)
            for t in threads:
                t.join()
        if errors:
            raise errors[0]
        return n

    def _pipeline(self, host_batches, _u8_gamma, before_reuse):
        """generator of (slot, pinned result); `before_reuse(slot)` is called before hy[slot] is overwritten"""
        m = self.model
        dev = m.trans.density_k.device
        if dev.type != "cuda":
            raise RuntimeError("StreamedCIDNet: the model must live on an sm_100 CUDA device (no CPU fallback)")
        main = torch.cuda.current_stream(dev)
        pending = []                                               # slots submitted, not yet yielded
        n = 0
        with torch.no_grad():
            for hx in host_batches:
                want = torch.float32 if _u8_gamma is None else torch.uint8
                if hx.dtype != want or hx.is_cuda:
                    raise RuntimeError(f"StreamedCIDNet expects {want} host tensors")
                self._setup(hx.shape, dev, want)
                slot = n % self.depth
                while pending and pending[0] != slot and self.out_done[pending[0]].query():
                    done = pending.pop(0)                          # finished early: hand it out now (a sink can start)
                    yield done, self.hy[done]
                if n >= self.depth:                                # the slot's previous result must have been handed out
                    while slot in pending:
                        done = pending.pop(0)
                        self.out_done[done].synchronize()
                        yield done, self.hy[done]
                    if before_reuse is not None:
                        before_reuse(slot)
                with torch.cuda.stream(self.s_in):
                    if n >= self.depth:
                        self.s_in.wait_event(self.in_free[slot])
                    self.x[slot].copy_(hx, non_blocking=True)
                    self.in_ready[slot].record(self.s_in)
                main.wait_event(self.in_ready[slot])
                if n >= self.depth:
                    main.wait_event(self.out_done[slot])           # y[slot] no longer being copied out
                if _u8_gamma is None:
                    m(self.x[slot], out=self.y[slot])
                else:
                    m.enhance_u8(self.x[slot], _u8_gamma, out=self.y[slot])
                self.in_free[slot].record(main)
                self.out_ready[slot].record(main)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(self.out_ready[slot])
                    self.hy[slot].copy_(self.y[slot], non_blocking=True)
                    self.out_done[slot].record(self.s_out)
                pending.append(slot)
                n += 1
            while pending:
                done = pending.pop(0)
                self.out_done[done].synchronize()
                yield done, self.hy[done]

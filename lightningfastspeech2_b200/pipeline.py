"""Host <-> device pipelining around ``model(batch, inference=True)`` for batched synthesis.

The reference's generation loop (litfass/generate.py:186-223 -> SpeechGenerator.generate_samples,
synthesis/generator.py:152-170) is synchronous: collate on the host, forward, ``.cpu()`` the mel, next
batch.  On a B200 the mel tensor of one 64-utterance batch is 54 MB, i.e. ~1 ms of PCIe time per ~9.5 ms
of compute, and the device idles while it drains.  ``SynthesisStream`` keeps the same per-batch work
(every batch's inputs are copied from pinned host memory, every batch's mel and mask are copied back to
pinned host memory) but puts the device->host copies on a second CUDA stream, double-buffered, so batch
i's read-back overlaps batch i+1's kernels.  Nothing is skipped or cached across batches.
"""
import torch


class SynthesisStream:
    def __init__(self, model, depth=2, keys=("mel", "tgt_mask"), pieces=16):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.model = model
        self.depth = depth
        self.keys = tuple(keys)
        self.pieces = pieces
        self.device = model.device
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = [dict(host={}, done=None, keep=None) for _ in range(depth)]
        self._n = 0

    def _host_buffer(self, slot, key, t):
        buf = slot["host"].get(key)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype).pin_memory()  # pinned allocations are cached per slot and shape
            slot["host"][key] = buf
        return buf

    def submit(self, host_batch):
        """Run one batch; returns a ticket for collect().  The slot reused here must have been collected."""
        slot = self._slots[self._n % self.depth]
        if slot["done"] is not None:
            slot["done"].synchronize()  # its previous read-back has to be finished before the buffers are reused
        compute = torch.cuda.current_stream(self.device)
        with torch.no_grad():
            out = self.model(host_batch, inference=True)  # H2D of phones / speaker happens inside forward
        ready = torch.cuda.Event()
        ready.record(compute)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            for k in self.keys:
                t = out[k]
                t.record_stream(self.copy_stream)  # the allocator must not hand the block out while it is being read
                # in pieces: the forward of the NEXT batch reads 8 bytes back (the LengthRegulator's frame count)
                # through the same device->host copy engine and must not queue behind one 54 MB transfer
                hb = self._host_buffer(slot, k, t)
                n = max(1, min(self.pieces, t.shape[0]))
                for src, dst in zip(t.chunk(n), hb.chunk(n)):
                    dst.copy_(src, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        slot["done"], slot["keep"] = done, out
        ticket = self._n
        self._n += 1
        return ticket

    def collect(self, ticket):
        """Block until the ticket's results are in pinned host memory; returns {key: host tensor} (views of the
        slot's buffers: valid until `depth` further submits)."""
        slot = self._slots[ticket % self.depth]
        slot["done"].synchronize()
        slot["keep"] = None
        return {k: slot["host"][k] for k in self.keys}

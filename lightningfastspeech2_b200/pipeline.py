"""Host <-> device pipelining around ``model(batch, inference=True)`` for batched synthesis.

The reference's generation loop (litfass/generate.py:186-223 -> SpeechGenerator.generate_samples,
synthesis/generator.py:152-170) is synchronous: collate on the host, forward, ``.cpu()`` the mel, next
batch.  On a B200 the mel tensor of one 64-utterance batch is 54 MB, i.e. ~1 ms of PCIe time per ~9.5 ms
of compute, and the device idles while it drains.  ``SynthesisStream`` keeps the same per-batch work
(every batch's inputs are copied from pinned host memory, every batch's mel and mask are copied back to
pinned host memory) but puts the device->host copies on a second CUDA stream, double-buffered, so batch
i's read-back overlaps batch i+1's kernels.  Nothing is skipped or cached across batches.

``compact=True`` reads back what the reference's caller keeps of a batch -- every utterance's mel cut at its own
length (generator.py:164-170) -- instead of the padded (B, L, 80) tensor: the valid frames are packed back to back on
the device (lfs2_pack_valid_rows) and one transfer of sum(frames) rows goes to the host (about half the bytes of the
padded tensor at the bench's length distribution); collect() returns per-utterance views of that buffer.

``vocoder=`` (a ``hifigan.Generator`` on the model's device, or a ``hifigan.Synthesiser``) appends the step the reference's
caller runs on every utterance right after the mel (``self.synth(mel)``, generator.py:170): the whole ragged batch is
vocoded on the device in one padded launch sequence (``Generator.forward(mel, lengths)``), the float waveform is scaled
to int16 there (the reference's ``* 32768`` cast, third_party/hifigan/__init__.py:42) and collect() additionally returns
``{"wav": [per-utterance int16 (frames_b * hop,) views], "hop": samples per frame}``.
"""
import torch

from . import ops


class SynthesisStream:
    def __init__(self, model, depth=2, keys=("mel", "tgt_mask"), pieces=16, compact=False, vocoder=None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.model = model
        self.depth = depth
        self.keys = tuple(keys)
        self.pieces = pieces
        self.compact = bool(compact)
        self.vocoder = getattr(vocoder, "vocoder", vocoder)   # a Synthesiser carries its Generator as .vocoder
        self.device = model.device
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = [dict(host={}, done=None, keep=None) for _ in range(depth)]
        self._n = 0

    def _host_buffer(self, slot, key, t):
        buf = slot["host"].get(key)
        if key == "mel_packed":  # ragged: a pinned buffer with headroom, a view of the rows in use
            cap = slot["host"].get("_packed_cap")
            if cap is None or cap.shape[0] < t.shape[0] or cap.shape[1:] != t.shape[1:]:
                cap = torch.empty((int(t.shape[0] * 1.25) + 1,) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory()
                slot["host"]["_packed_cap"] = cap
            buf = cap[: t.shape[0]]
            slot["host"][key] = buf
            return buf
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
        lens = None
        if self.compact:
            # frames per utterance: known on the host since the forward's one sync (or read back here on the paths that
            # size their tensors differently), cut at the padded length like tgt_mask
            lens, fl = out.get("frame_lengths_host"), out.get("frame_lengths")
            if fl is None:   # (the bucketed path assembles its result itself: count the frames tgt_mask keeps)
                fl = (~out["tgt_mask"]).sum(1)
            if lens is None:
                lens = fl.tolist()
            width = out["mel"].shape[1]
            lens = [min(int(n), width) for n in lens]
            out = dict(out)
            out["mel_packed"] = ops.pack_valid_rows(out["mel"].contiguous(), fl.contiguous(), sum(lens))
        copy_keys = ("mel_packed",) if self.compact else self.keys
        wav_lens = None
        if self.vocoder is not None:
            fl = out.get("frame_lengths")
            if fl is None:
                fl = (~out["tgt_mask"]).sum(1)
            host = out.get("frame_lengths_host")
            wav_lens = lens if lens is not None else [min(int(n), out["mel"].shape[1])
                                                      for n in (host if host is not None else fl.tolist())]
            out = dict(out)
            if out["mel"].shape[1] == 0:
                out["wav_i16"] = torch.zeros(out["mel"].shape[0], 0, device=self.device, dtype=torch.int16)
                hop = 0
            else:
                # (B, L, n_mels) -> the reference vocoder's channels-first layout; whatever the mel holds past an
                # utterance's end is masked by Generator.forward (every stage sees zeros there, like a per-utterance call)
                with torch.no_grad():
                    wav = self.vocoder(out["mel"].transpose(1, 2), fl.clamp(max=out["mel"].shape[1]))
                hop = wav.shape[-1] // out["mel"].shape[1]
                # the reference's (wav * 32768).astype("int16"): numpy's float -> int16 cast wraps out-of-range values;
                # int32 first and a wrapping narrow reproduce it bit for bit for |wav| < 2^16
                out["wav_i16"] = (wav.squeeze(1) * 32768.0).to(torch.int32).to(torch.int16)
            slot["hop"] = hop
            copy_keys = tuple(copy_keys) + ("wav_i16",)
        ready = torch.cuda.Event()
        ready.record(compute)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            for k in copy_keys:
                t = out[k]
                t.record_stream(self.copy_stream)  # the allocator must not hand the block out while it is being read
                # in pieces: the forward of the NEXT batch reads 8 bytes back (the LengthRegulator's frame count)
                # through the same device->host copy engine and must not queue behind one 54 MB transfer
                hb = self._host_buffer(slot, k, t)
                if t.shape[0] == 0:
                    continue
                n = max(1, min(self.pieces, t.shape[0]))
                for src, dst in zip(t.chunk(n), hb.chunk(n)):
                    dst.copy_(src, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        slot["done"], slot["keep"], slot["lens"], slot["wav_lens"] = done, out, lens, wav_lens
        ticket = self._n
        self._n += 1
        return ticket

    def collect(self, ticket):
        """Block until the ticket's results are in pinned host memory; returns {key: host tensor} (views of the
        slot's buffers: valid until `depth` further submits)."""
        slot = self._slots[ticket % self.depth]
        slot["done"].synchronize()
        slot["keep"] = None
        if self.compact:  # {"mel": [per-utterance (frames_b, n_mels) views of the packed host buffer], "lengths": [...]}
            lens = slot["lens"]
            packed = slot["host"]["mel_packed"]
            res = {"mel": list(packed.split(lens)) if lens else [], "lengths": lens}
        else:
            res = {k: slot["host"][k] for k in self.keys}
        if self.vocoder is not None:
            hop, wl = slot["hop"], slot["wav_lens"]
            w = slot["host"]["wav_i16"]
            res.update(wav=[w[i, : n * hop] for i, n in enumerate(wl)], hop=hop, lengths=wl)
        return res

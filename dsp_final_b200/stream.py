"""DataLoader-free batches for the CNN trainers.

The reference feeds `train_supervised_classifier` (src/tasks/classification.py:52-78) from
`DataLoader(Esc50FeatureDataset(..., postprocess=to_tensor), num_workers=2)`
(scripts/models/train_cnn.py:46-55, src/datasets/esc50.py:92-97): every item is a cached log-mel
array turned into a [1, n_mels, n_frames] float32 tensor (same as LogMelTransform,
src/train/transforms.py:16-18), collated to [B, 1, n_mels, n_frames] on the host by forked workers and
copied to the GPU inside the epoch loop.

LogMelBatches makes those batches where they are consumed: clips (float32, or int16 PCM) stay resident
in HBM -- or are streamed from pinned host memory one batch ahead -- and every batch is one launch of the
fused kernel whose store epilogue already writes the CNN layout (dspx_log_mel_nchw).  It is what the epoch
loops need from a loader: iteration yields (feats [B,1,n_mels,T] on the GPU, targets [B] int64 on the GPU),
`len(loader)` is the number of batches and `loader.dataset` has the clip count (`len(loader.dataset)`,
classification.py:33,49).  No worker processes, so no CUDA-after-fork hazard.
"""
from __future__ import annotations

from typing import Any, Iterator

import numpy as np


class LogMelBatches:
    def __init__(self, clips, targets, cfg: Any, batch_size: int, shuffle: bool = False, seed: int = 0,
                 device: int | None = None, drop_last: bool = False, normalize: bool = True):
        """clips: [N, L] float32 / int16 -- a torch CUDA tensor (resident) or a NumPy array / CPU tensor
        (streamed through pinned staging).  int16 is PCM16, converted and peak-normalised on the GPU like
        load_audio + normalize_audio (src/utils/audio.py:19-38) when `normalize`.  targets: [N] integers."""
        import torch

        from . import _lib

        _lib.require_device()
        if batch_size < 1:
            raise ValueError("batch_size must be positive")
        self.cfg, self.batch_size, self.shuffle, self.drop_last, self.normalize = cfg, int(batch_size), shuffle, drop_last, normalize
        self._epoch, self._seed = 0, int(seed)
        self._resident = isinstance(clips, torch.Tensor) and clips.is_cuda
        if self._resident:
            self.device = clips.device
            self._clips = clips if clips.stride(1) == 1 else clips.contiguous()
        else:
            self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
            host = clips if isinstance(clips, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(clips))
            if host.dtype not in (torch.float32, torch.int16):
                host = host.to(torch.float32)
            self._clips = host.contiguous()
        if self._clips.dim() != 2:
            raise ValueError("clips must be [n_clips, n_samples]")
        t = torch.as_tensor(np.asarray(targets.cpu() if isinstance(targets, torch.Tensor) else targets, dtype=np.int64))
        if t.shape != (self._clips.shape[0],):
            raise ValueError("one target per clip expected")
        self._targets = t.to(self.device)
        self.dataset = _Sized(self._clips.shape[0])            # what the epoch loops ask the loader for

    def __len__(self) -> int:
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _order(self):
        import torch

        n = len(self.dataset)
        if not self.shuffle:
            return None
        gen = torch.Generator()
        gen.manual_seed(self._seed + self._epoch)              # a new permutation every epoch, reproducible
        return torch.randperm(n, generator=gen)

    def __iter__(self) -> Iterator:
        import torch

        from .batch import log_mel_nchw

        order = self._order()
        self._epoch += 1
        n, bs = len(self.dataset), self.batch_size
        starts = [s for s in range(0, n, bs) if not (self.drop_last and s + bs > n)]
        dev_order = order.to(self.device) if order is not None else None

        def rows(s):
            e = min(n, s + bs)
            return slice(s, e) if order is None else order[s:e]

        if self._resident:
            for s in starts:
                r = rows(s)
                if isinstance(r, slice):
                    x, y = self._clips[r], self._targets[r]
                else:
                    idx = dev_order[s:min(n, s + bs)]
                    x, y = self._clips.index_select(0, idx), self._targets.index_select(0, idx)
                yield self._features(x, log_mel_nchw), y
            return
        # streamed: batch i+1 is gathered into pinned memory and copied on a side stream while batch i computes
        copy_stream = torch.cuda.Stream(self.device)
        main = torch.cuda.current_stream(self.device)
        pinned = [torch.empty((bs, self._clips.shape[1]), dtype=self._clips.dtype, pin_memory=True) for _ in range(2)]
        staged = [torch.empty((bs, self._clips.shape[1]), dtype=self._clips.dtype, device=self.device) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def stage(i, s):
            r = rows(s)
            m = (r.stop - r.start) if isinstance(r, slice) else len(r)
            slot = i & 1
            freed[slot].synchronize()                          # pinned slot no longer being read by an earlier copy
            if isinstance(r, slice):
                pinned[slot][:m].copy_(self._clips[r])
            else:
                torch.index_select(self._clips, 0, r, out=pinned[slot][:m])
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[slot])            # the kernel that read staged[slot] two batches ago is done
                staged[slot][:m].copy_(pinned[slot][:m], non_blocking=True)
                ready[slot].record(copy_stream)
            return m

        for ev in freed:
            ev.record(main)
        counts = {}
        if starts:
            counts[0] = stage(0, starts[0])
        for i, s in enumerate(starts):
            if i + 1 < len(starts):
                counts[i + 1] = stage(i + 1, starts[i + 1])
            slot = i & 1
            main.wait_event(ready[slot])
            x = staged[slot][:counts[i]]
            feats = self._features(x, log_mel_nchw)
            freed[slot].record(main)
            r = rows(s)
            y = self._targets[r] if isinstance(r, slice) else self._targets.index_select(0, dev_order[s:min(n, s + bs)])
            yield feats, y

    def _features(self, x, log_mel_nchw):
        from .batch import pcm16_to_float

        if str(x.dtype) == "torch.int16":
            x = pcm16_to_float(x.contiguous(), self.normalize)
        return log_mel_nchw(x.contiguous(), self.cfg, device=self.device.index)


class _Sized:
    def __init__(self, n: int):
        self._n = int(n)

    def __len__(self) -> int:
        return self._n

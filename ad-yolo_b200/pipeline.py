"""Host -> device hand-off of raw batches (the part that touches the hot path).

The reference moves float32 *features* from DataLoader workers to the GPU (``train.py:48``).  Here
workers hand over int16 audio + the event table; ``HostBatchPipeline`` double-buffers the pinned
host -> device copies on a side stream so that the copy of batch i+1 overlaps the kernels of batch
i.  Pure plumbing (torch streams / events); no arithmetic.
"""
from __future__ import annotations

import torch


class HostBatchPipeline:
    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self._slots = [None] * depth          # per slot: list of device tensors
        self._ready = [torch.cuda.Event() for _ in range(depth)]
        self._free = [torch.cuda.Event() for _ in range(depth)]
        self._head = 0                         # next slot to fill
        self._tail = 0                         # next slot to hand out
        self._inflight = 0

    def submit(self, *host_tensors: torch.Tensor):
        """Start the asynchronous copy of one batch (pinned host tensors) into the next slot."""
        if self._inflight >= self.depth:
            raise RuntimeError("pipeline full: call get() before submitting more batches")
        s = self._head
        if self._slots[s] is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(self._slots[s], host_tensors)) \
                or len(self._slots[s]) != len(host_tensors):
            self._slots[s] = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_tensors]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._free[s])       # consumer of the previous use is done
            for d, h in zip(self._slots[s], host_tensors):
                d.copy_(h, non_blocking=True)
            self._ready[s].record(self.copy_stream)
        self._head = (s + 1) % self.depth
        self._inflight += 1

    def get(self):
        """Device tensors of the oldest submitted batch, ordered after its copy on the current stream."""
        if self._inflight == 0:
            raise RuntimeError("pipeline empty")
        s = self._tail
        torch.cuda.current_stream(self.device).wait_event(self._ready[s])
        self._tail = (s + 1) % self.depth
        self._inflight -= 1
        self._last = s
        return self._slots[s]

    def release(self):
        """Mark the batch returned by the last get() as consumed (recorded on the current stream)."""
        self._free[self._last].record(torch.cuda.current_stream(self.device))

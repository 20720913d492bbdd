"""Host -> device hand-off of raw batches (the part that touches the hot path).

The reference moves float32 *features* from DataLoader workers to the GPU (``train.py:48``).  Here
workers hand over int16 audio + the event table; ``HostBatchPipeline`` double-buffers the pinned
host -> device copies on a side stream so that the copy of batch i+1 overlaps the kernels of batch
i.  Pure plumbing (torch streams / events); no arithmetic.
"""
from __future__ import annotations

import os

import torch


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_device(device) -> dict:
    """Restrict the calling process to the CPU cores of the NUMA node the GPU hangs off, so that the
    pinned staging buffers allocated afterwards are node-local (first touch) and host -> device copies
    do not cross the socket interconnect.  With one process per GPU (``torchrun``) an unbound rank can
    land on the other socket: 8 ranks then share the inter-socket link instead of 8 PCIe root ports.
    Call once per rank after ``torch.cuda.set_device`` and before ``pin_memory``.  Best effort:
    returns ``{"bound": False, "why": ...}`` when the topology is not visible (sysfs, then NVML)."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    props = torch.cuda.get_device_properties(idx)
    try:
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
    except AttributeError:
        return {"bound": False, "why": "torch does not expose the PCI address"}
    allowed = os.sched_getaffinity(0)
    cpus, node, how = set(), None, None
    try:
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
        if node >= 0:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus, how = _parse_cpulist(f.read()), "sysfs"
    except (OSError, ValueError):
        pass
    if not cpus:
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByPciBusId(("0000" + bus).encode() if len(bus) < 13 else bus.encode())
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {i for i in range(os.cpu_count()) if (int(words[i // 64]) >> (i % 64)) & 1}
            how = "nvml"
        except Exception as e:  # noqa: BLE001 - topology is optional
            return {"bound": False, "why": f"no NUMA information for {bus} ({type(e).__name__})"}
    cpus &= allowed
    if not cpus:
        return {"bound": False, "why": f"NUMA node {node} has no CPU this process may use"}
    if cpus != allowed:
        os.sched_setaffinity(0, cpus)
    return {"bound": True, "pci": bus, "numa_node": node, "cpus": len(cpus), "via": how}


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
            # the caching allocator may hand out blocks whose previous tenant still has kernels queued on the
            # allocating (current) stream: order the first copy into a fresh slot after them
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
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

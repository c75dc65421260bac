"""Host -> device feeding of Gaussian scenes for the end-to-end path.

A scene is 352 B/Gaussian (means 12, covariances 36, harmonics 300, opacity 4): at 1M Gaussians the upload (~370 MB
over PCIe) takes several times longer than rendering it, so a serving / training loop that receives its Gaussians from
the host should upload scene i+1 on a copy stream while scene i is being rasterized.  ``HostSceneFeeder`` does exactly
that with pinned source buffers, a small ring of reusable device buffers and stream-ordered hand-over.
"""
from __future__ import annotations

import os
from typing import Dict, NamedTuple, Optional

import torch
from torch import Tensor


def gpu_numa_node(device) -> Optional[int]:
    """NUMA node the GPU's PCIe root hangs off (sysfs), or None when the platform does not say."""
    try:
        p = torch.cuda.get_device_properties(device)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(device) -> Optional[dict]:
    """Pin the calling process to the CPUs of the GPU's NUMA node BEFORE it allocates pinned host buffers: pinned pages are
    placed on the node of the allocating thread, and an upload from the other socket crosses the inter-socket link on top of
    PCIe.  One process per GPU (the torchrun shape) makes this a per-rank decision.  Returns {"node", "cpus"} or None when
    the topology is not exposed (then nothing is changed)."""
    node = gpu_numa_node(device)
    if node is None:
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)          # never widen what the launcher / container allows
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception:
        return None


class Ticket(NamedTuple):
    tensors: Dict[str, Tensor]
    ready: torch.cuda.Event
    nbytes: int


class HostSceneFeeder:
    """``depth`` sets of device buffers are allocated once (per tensor name / shape / dtype) and reused in a ring, so a
    steady-state step allocates nothing.  Intended call pattern (one producer/consumer thread):

        t = feeder.submit(host_i)            # upload i starts on the copy stream
        loop:  d = feeder.get(t);  t = feeder.submit(host_{i+1});  compute(d)

    A ring slot is overwritten ``depth`` submits later; the copy stream first waits for everything enqueued on the
    consumer's stream so far, which includes the last computation that read that slot."""

    def __init__(self, device, depth: int = 3) -> None:
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = max(2, int(depth))
        self.rings: Dict[tuple, list] = {}
        self.n = 0

    def _slot(self, key: str, t: Tensor) -> Tensor:
        k = (key, tuple(t.shape), t.dtype)
        ring = self.rings.get(k)
        if ring is None:
            ring = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for _ in range(self.depth)]
            self.rings[k] = ring
        return ring[self.n % self.depth]

    def submit(self, host: Dict[str, Tensor]) -> Ticket:
        """Start the asynchronous upload of a dict of (ideally pinned) host tensors; returns immediately."""
        out, nbytes = {}, 0
        cur = torch.cuda.current_stream(self.device)
        # ring buffers are allocated on the CONSUMER's stream (outside the copy-stream context): the caching allocator
        # then ties their blocks to the stream that reads them, so a dropped feeder cannot hand a block to another
        # consumer-stream allocation while rasterizer kernels still read it
        slots = {k: self._slot(k, t) for k, t in host.items()}
        free = torch.cuda.Event()
        free.record(cur)                       # the slot's previous reader was enqueued before this point
        self.stream.wait_event(free)
        with torch.cuda.stream(self.stream):
            for k, t in host.items():
                dst = slots[k]
                dst.record_stream(self.stream)  # ... and the copy stream writes them
                dst.copy_(t, non_blocking=True)
                out[k] = dst.detach()          # fresh tensor object on the same storage: no autograd state carried over
                nbytes += t.numel() * t.element_size()
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.n += 1
        return Ticket(out, ev, nbytes)

    def get(self, ticket: Ticket) -> Dict[str, Tensor]:
        """Make the current stream wait for the upload and hand the device tensors over to it."""
        torch.cuda.current_stream(self.device).wait_event(ticket.ready)
        return ticket.tensors

"""Host -> device feeding of Gaussian scenes for the end-to-end path.

A scene is 352 B/Gaussian (means 12, covariances 36, harmonics 300, opacity 4): at 1M Gaussians the upload (~370 MB
over PCIe) takes several times longer than rendering it, so a serving / training loop that receives its Gaussians from
the host should upload scene i+1 on a copy stream while scene i is being rasterized.  ``HostSceneFeeder`` does exactly
that with pinned source buffers and stream-ordered hand-over; it owns no device memory beyond the in-flight uploads.
"""
from __future__ import annotations

from typing import Dict, NamedTuple

import torch
from torch import Tensor


class Ticket(NamedTuple):
    tensors: Dict[str, Tensor]
    ready: torch.cuda.Event
    nbytes: int


class HostSceneFeeder:
    def __init__(self, device) -> None:
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)

    def submit(self, host: Dict[str, Tensor]) -> Ticket:
        """Start the asynchronous upload of a dict of (ideally pinned) host tensors; returns immediately."""
        out, nbytes = {}, 0
        with torch.cuda.stream(self.stream):
            for k, t in host.items():
                out[k] = t.to(self.device, non_blocking=True)
                nbytes += t.numel() * t.element_size()
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return Ticket(out, ev, nbytes)

    def get(self, ticket: Ticket) -> Dict[str, Tensor]:
        """Make the current stream wait for the upload and hand the device tensors over to it."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ticket.ready)
        for t in ticket.tensors.values():
            t.record_stream(cur)
        return ticket.tensors

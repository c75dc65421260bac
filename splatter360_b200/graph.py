"""CUDA-graph capture of one forward(+backward) of the rasterizer for a fixed problem shape.

With ``instance_capacity`` set the forward pass reads nothing back to the host, so the ~15 kernel launches and memsets
of a view collapse into one graph launch -- what matters when a view takes a fraction of a millisecond (small scenes,
cube faces).  The graph works on static buffers: write new Gaussians / camera / seed gradient into them, ``replay()``,
read the outputs.  Capacity overflow cannot resize a captured graph; check ``overflowed()`` now and then.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import rasterizer as R


class GraphedView:
    def __init__(self, settings: R.GaussianRasterizationSettings, means3D: Tensor, cov3D: Tensor, opacities: Tensor,
                 shs: Optional[Tensor] = None, colors_precomp: Optional[Tensor] = None, with_backward: bool = True,
                 warmup: int = 2) -> None:
        if settings.instance_capacity is None:
            raise ValueError("GraphedView needs settings.instance_capacity (no host read-back inside a CUDA graph)")
        dev = means3D.device
        # static inputs: the graph bakes these addresses; update them in place
        self.settings = settings._replace(
            viewmatrix=settings.viewmatrix.to(dev).float().contiguous().clone(),
            projmatrix=settings.projmatrix.to(dev).float().contiguous().clone(),
            campos=settings.campos.to(dev).float().contiguous().clone(), bg=settings.bg.to(dev).float().contiguous().clone())
        self.means3D, self.cov3D = means3D.detach().clone().contiguous(), cov3D.detach().clone().contiguous()
        self.opacities = opacities.detach().reshape(-1).clone().contiguous()
        self.shs = None if shs is None else shs.detach().clone().contiguous()
        self.colors = None if colors_precomp is None else colors_precomp.detach().clone().contiguous()
        H, W = int(settings.image_height), int(settings.image_width)
        self.grad_color = torch.zeros(3, H, W, device=dev)
        self.with_backward = with_backward
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._run()

    def _run(self) -> None:
        self.color, self.state = R.forward_raw(self.settings, self.means3D, self.cov3D, self.opacities, self.shs, self.colors)
        self.grads = None
        if self.with_backward:
            self.grads = R.backward_raw(self.settings, self.means3D, self.cov3D, self.opacities, self.shs, self.colors,
                                        self.state, self.grad_color)

    def set_camera(self, viewmatrix: Tensor, projmatrix: Tensor, campos: Tensor) -> None:
        self.settings.viewmatrix.copy_(viewmatrix)
        self.settings.projmatrix.copy_(projmatrix)
        self.settings.campos.copy_(campos)

    def replay(self):
        """Re-run the captured view on whatever the static buffers hold now.  Returns (color, grads-or-None)."""
        self.graph.replay()
        return self.color, self.grads

    def overflowed(self) -> bool:
        return R.overflowed(self.state)

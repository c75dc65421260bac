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


class GraphedStep:
    """One training-style step -- forward, MSE loss against a target image, backward to all Gaussian gradients -- as ONE
    CUDA graph launch (about 20 kernels and memsets otherwise, each a host-side launch through ctypes).

        step = GraphedStep(settings, means, cov3D, opacities, shs, target)     # tensors are used IN PLACE (no copies)
        for ...:
            step.set_camera(viewmatrix, projmatrix, campos)                    # tiny device-to-device copies
            # ... update means / cov3D / opacities / shs / target in place (e.g. the encoder writes into them) ...
            loss, grads = step.replay()                                        # grads: dict of static tensors
        assert not step.overflowed()                                           # one sync, whenever convenient

    The instance buffers are sized once: ``settings.instance_capacity`` if given, otherwise 1.25 x the count of an exact
    warm-up forward with the initial camera.  A view that needs more sets the device-side overflow flag (its result is
    then incomplete): ``overflowed()`` reports it, ``regrow()`` re-captures with a larger capacity."""

    def __init__(self, settings: R.GaussianRasterizationSettings, means3D: Tensor, cov3D: Tensor, opacities: Tensor,
                 shs: Optional[Tensor], target: Tensor, colors_precomp: Optional[Tensor] = None, weight: float = 1.0,
                 margin: float = 1.25) -> None:
        dev = means3D.device
        self.margin, self.weight = margin, float(weight)
        self.means3D, self.cov3D = means3D.detach(), cov3D.detach()
        self.opacities = opacities.detach().reshape(-1)
        self.shs = None if shs is None else shs.detach()
        self.colors = None if colors_precomp is None else colors_precomp.detach()
        self.target = target.detach()
        for t in (self.means3D, self.cov3D, self.opacities, self.shs, self.colors, self.target):
            if t is not None and (not t.is_contiguous() or t.dtype != torch.float32 or t.device != dev):
                raise ValueError("GraphedStep works in place: pass contiguous float32 tensors on one CUDA device")
        self.settings = settings._replace(
            viewmatrix=settings.viewmatrix.to(dev).float().contiguous().clone(),
            projmatrix=settings.projmatrix.to(dev).float().contiguous().clone(),
            campos=settings.campos.to(dev).float().contiguous().clone(), bg=settings.bg.to(dev).float().contiguous().clone(),
            capacity_tracker=None)
        if self.settings.instance_capacity is None:
            _, st = R.forward_raw(self.settings, self.means3D, self.cov3D, self.opacities, self.shs, self.colors)
            self.settings = self.settings._replace(instance_capacity=max(4096, int(st.num_rendered * margin) + 1))
            del st
        self._capture()

    def _run(self) -> None:
        import ctypes
        from . import _lib
        lib = _lib.load()
        self.color, self.state = R.forward_raw(self.settings, self.means3D, self.cov3D, self.opacities, self.shs, self.colors)
        self.loss = torch.empty((), dtype=torch.float32, device=self.color.device)
        self.grad_color = torch.empty_like(self.color)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(lib.s360_mse_loss_grad(p(self.color), p(self.target), ctypes.c_int64(self.color.numel()),
                                          ctypes.c_float(self.weight), p(self.loss), p(self.grad_color),
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.grads = R.backward_raw(self.settings, self.means3D, self.cov3D, self.opacities, self.shs, self.colors,
                                    self.state, self.grad_color)

    def _capture(self) -> None:
        dev = self.means3D.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._run()

    def set_camera(self, viewmatrix: Tensor, projmatrix: Tensor, campos: Tensor) -> None:
        self.settings.viewmatrix.copy_(viewmatrix, non_blocking=True)
        self.settings.projmatrix.copy_(projmatrix, non_blocking=True)
        self.settings.campos.copy_(campos, non_blocking=True)

    def replay(self):
        self.graph.replay()
        return self.loss, self.grads

    def overflowed(self) -> bool:
        return R.overflowed(self.state)

    def regrow(self) -> None:
        """Re-capture with a capacity that fits the view that overflowed (synchronises)."""
        need = R.instances_needed(self.state)
        self.settings = self.settings._replace(instance_capacity=max(4096, int(need * self.margin) + 1))
        self._capture()

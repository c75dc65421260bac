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


class GraphedAutogradStep:
    """``loss = fn(); loss.backward()`` -- any loss built from this package's autograd entry points
    (``DecoderSplattingCUDA`` / ``DecoderSplattingERP`` with ``sync_free=True``, ``Cube2Equirec.from_faces``,
    ``loss.mse_loss``, ``GaussianRasterizer`` with a ``capacity_tracker``) -- as ONE CUDA graph launch.

    The reference's evaluation shape (six cube faces through the decoder, stitched to a panorama, MSE; about 60 launches
    and as many torch / ctypes calls) is host-bound when issued call by call: 1.47 ms per step against ~1.05 ms of
    device work on one B200.

        dec = DecoderSplattingCUDA(sync_free=True)
        def fn():                                   # reads static tensors only; new poses / targets are written IN PLACE
            out = dec(gaussians, extrinsics, intrinsics, near, far, (256, 256))
            return mse_loss(c2e.from_faces(out.color), target)
        step = GraphedAutogradStep(fn, params=[gaussians.means, ...], trackers=dec.capacity_trackers)
        for ...:
            extrinsics.copy_(new_poses)             # in place
            loss = step.replay()                    # params[i].grad hold the gradients (static tensors)
        assert not step.overflowed()                # one sync, whenever convenient

    Construction runs ``fn`` eagerly a few times (the trackers learn the pair / instance counts, the decoder's per-input
    caches fill), freezes the trackers at ``margin`` x the largest count seen, and captures.  ``fn`` must not synchronise
    (no ``.item()``, no data-dependent Python branches) and must see the same tensor OBJECTS on every call."""

    def __init__(self, fn, params, trackers=None, warmup: int = 3) -> None:
        self.fn, self.params = fn, list(params)
        self.trackers = list(trackers.values()) if isinstance(trackers, dict) else list(trackers or [])
        self._tracker_source = trackers
        dev = self.params[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(2, warmup)):
                self._eager()
            if isinstance(self._tracker_source, dict):       # the decoder creates its trackers on first use
                self.trackers = list(self._tracker_source.values())
            for t in self.trackers:
                t.freeze()
            self._eager()                                    # once more with the frozen capacities
        torch.cuda.current_stream(dev).wait_stream(side)
        for p in self.params:
            p.grad = None
        self._ovf = torch.zeros(1, dtype=torch.int32, device=dev)   # sticky: OR of the overflow flags of every replay
        self.graph = torch.cuda.CUDAGraph()
        # captured on the stream the eager runs used: autograd remembers the stream of each leaf's accumulation node
        with torch.cuda.graph(self.graph, stream=side):
            self.loss = self.fn()
            self.loss.backward()
            for t in self.trackers:
                if t._static_counters is not None:
                    self._ovf.bitwise_or_(t._static_counters[1:2])
        self.grads = [p.grad for p in self.params]

    def _eager(self) -> None:
        for p in self.params:
            p.grad = None
        self.fn().backward()

    def replay(self) -> Tensor:
        self.graph.replay()
        return self.loss

    def overflowed(self) -> bool:
        """Did ANY replay since construction need more than the frozen capacities (its result was incomplete)?  One sync."""
        return bool(int(self._ovf.item()) & 3)

    def release(self) -> None:
        """Hand the trackers back to eager use."""
        for t in self.trackers:
            t.unfreeze()

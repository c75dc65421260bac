"""CUDA-graph capture of the reference's evaluation-shaped step (six cube faces through DecoderSplattingCUDA + Cube2Equirec
+ MSE, /root/reference/src/model/model_wrapper_erp.py:202-205, 336-345): graph.GraphedAutogradStep replays must equal the
eager autograd step at every pose (to float-atomic summation order: the loss sum and the per-Gaussian RED.ADD accumulation are
not order-deterministic from run to run), and the on-device 4x4 inverse that makes the pose -> view-matrix step
capturable must equal the exact inverse to float rounding (the reference: `extrinsics.inverse()`, cuda_splatting.py:84)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_invert4x4_kernel_matches_the_float64_inverse():
    from splatter360_b200 import camera, synthetic
    dev = torch.device("cuda")
    poses = synthetic.trajectory(37, seed=3).to(dev)                       # rigid camera-to-world matrices
    g = torch.Generator(device=dev).manual_seed(1)
    general = torch.randn(50, 4, 4, device=dev, generator=g) + 2 * torch.eye(4, device=dev)   # needs pivoting now and then
    general[7, 0, 0] = 0.0                                                 # zero leading pivot
    for m in (poses, general, poses.reshape(37, 1, 4, 4)):
        got = camera.inverse(m)
        want = torch.linalg.inv(m.double())
        assert got.shape == m.shape and got.dtype == torch.float32
        err = (got.double() - want).abs().amax(dim=(-1, -2)) / want.abs().amax(dim=(-1, -2))
        assert float(err.max()) < 2e-7, float(err.max())
    # differentiable poses keep the torch path (autograd through the inverse)
    p = poses[:2].clone().requires_grad_()
    camera.inverse(p).sum().backward()
    assert p.grad is not None and torch.isfinite(p.grad).all()
    assert camera.inverse(torch.zeros(0, 4, 4, device=dev)).shape == (0, 4, 4)


def test_graphed_decoder_step_equals_the_eager_step():
    from splatter360_b200 import cubemap, synthetic
    from splatter360_b200.decoder import DecoderSplattingCUDA, Gaussians
    from splatter360_b200.graph import GraphedAutogradStep
    from splatter360_b200.loss import mse_loss
    dev = torch.device("cuda")
    F, H, W = 64, 128, 256
    sc = synthetic.random_cloud_scene(20000, seed=11, ref_width=256, device=dev)
    mk = lambda: Gaussians(*(t[None].contiguous().clone().requires_grad_() for t in
                             (sc.means, sc.covariances, sc.harmonics, sc.opacities)))
    g_e, g_g = mk(), mk()
    poses = synthetic.trajectory(4, seed=2).to(dev)
    faces = cubemap.cube_face_extrinsics(poses)
    Kf = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev).expand(1, 6, 3, 3)
    near, far = torch.ones(1, 6, device=dev), torch.full((1, 6), 100.0, device=dev)
    c2e = cubemap.Cube2Equirec(F, H, W).to(dev)
    target = torch.rand(1, 3, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(5))

    def loss_of(dec, g, ext):
        return mse_loss(c2e.from_faces(dec(g, ext, Kf, near, far, (F, F)).color), target)

    dec_g = DecoderSplattingCUDA(sync_free=True).to(dev)
    ext = faces[0][None].clone()
    params = [g_g.means, g_g.covariances, g_g.harmonics, g_g.opacities]
    step = GraphedAutogradStep(lambda: loss_of(dec_g, g_g, ext), params, trackers=dec_g.capacity_trackers)
    assert step.trackers and all(t.frozen for t in step.trackers)
    dec_e = DecoderSplattingCUDA().to(dev)                                  # exact counts, eager
    for i in (1, 3, 0):
        ext.copy_(faces[i][None])
        loss_g = step.replay().detach().clone()
        grads_g = [p.grad.clone() for p in params]
        for t in (g_e.means, g_e.covariances, g_e.harmonics, g_e.opacities):
            t.grad = None
        loss_e = loss_of(dec_e, g_e, faces[i][None])
        loss_e.backward()
        assert abs(float(loss_g) - float(loss_e.detach())) <= 2e-6 * abs(float(loss_e.detach())), (i, float(loss_g), float(loss_e.detach()))
        for a, b in zip(grads_g, (g_e.means, g_e.covariances, g_e.harmonics, g_e.opacities)):
            assert float(b.grad.abs().max()) > 0 and a.shape == b.grad.shape
            err = float((a.double() - b.grad.double()).norm() / b.grad.double().norm())
            assert err < 2e-6, (i, err)
    assert not step.overflowed()
    step.release()
    assert not any(t.frozen for t in step.trackers)


def test_graphed_step_reports_overflow_of_the_frozen_capacities():
    """A pose that needs more instances than the frozen capacity sets the sticky flag (the replay's result is incomplete)."""
    from splatter360_b200 import camera, synthetic
    from splatter360_b200 import rasterizer as R
    from splatter360_b200.graph import GraphedAutogradStep
    from splatter360_b200.loss import mse_loss
    dev = torch.device("cuda")
    H, W = 64, 128
    sc = synthetic.random_cloud_scene(5000, seed=4, ref_width=128, device=dev)
    means = sc.means.clone().requires_grad_()
    cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous()
    shs = sc.harmonics.permute(0, 2, 1).contiguous()
    cam = camera.erp_camera(synthetic.trajectory(1, seed=0).to(dev))
    scale = torch.ones((), device=dev)
    tracker = R.CapacityTracker(margin=1.05)
    target = torch.zeros(3, H, W, device=dev)

    def fn():
        s = R.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev), scale_modifier=1.0,
            viewmatrix=cam.view_matrix[0], projmatrix=cam.full_projection[0], sh_degree=4, campos=cam.campos[0],
            prefiltered=False, debug=False, projection="erp", capacity_tracker=tracker)
        color, _ = R.GaussianRasterizer(s)(means3D=means, means2D=torch.zeros_like(means), opacities=sc.opacities, shs=shs,
                                           cov3D_precomp=cov6 * scale)
        return mse_loss(color, target)

    step = GraphedAutogradStep(fn, [means], trackers=[tracker])
    step.replay()
    assert not step.overflowed()
    scale.fill_(9.0)          # every splat three times wider: far more (tile, Gaussian) instances than the frozen capacity
    step.replay()
    assert step.overflowed()

"""The reference's chain from the encoder head to the loss on this library's kernels (SURVEY.md sec. 8 rows f4 -> a1 -> f3):
raw head output + depths of two context panoramas -> `GaussianAdapterERP` (one fused forward / backward kernel,
/root/reference/src/model/encoder/common/gaussian_adapter_erp.py:49-119) -> `DecoderSplattingCUDA` (six cube faces in one batched
pass, decoder_splatting_cuda.py:34-70) -> `Cube2Equirec` stitch -> MSE, and back to d(raw), d(depths).  The same chain with the
adapter replaced by the reference's op sequence in float64 torch (tests/test_adapter.py:_torch_adapter) must give the same loss
and the same gradients at the head -- the one place where all the pieces' layouts have to fit together."""
import pytest
import torch

from helpers import rel_l2
from test_adapter import _rand_rot, _torch_adapter

pytestmark = pytest.mark.gpu


def test_adapter_decoder_stitch_loss_chain_matches_the_torch_adapter_chain():
    from splatter360_b200 import adapter, cubemap, synthetic
    from splatter360_b200.decoder import DecoderSplattingCUDA, Gaussians
    from splatter360_b200.loss import mse_loss
    dev = "cuda"
    b, v, h, w, deg, F = 1, 2, 32, 64, 4, 32
    r, d_sh = h * w, 25
    gen = torch.Generator().manual_seed(11)
    ext = torch.eye(4).repeat(b * v, 1, 1)
    ext[:, :3, :3] = _rand_rot(gen, b * v).float()
    ext[:, :3, 3] = 0.2 * torch.randn(b * v, 3, generator=gen)
    depths = 1.0 + 3.0 * torch.rand(b * v, r, generator=gen)
    raw = torch.randn(b * v, r, 7 + 3 * d_sh, generator=gen)
    raw[..., :3] -= 2.0                                                       # small splats (a few pixels), like a trained head
    opac = torch.rand(b, v * r, generator=gen) * 0.7 + 0.2
    target = torch.rand(1, 3, 2 * F, 4 * F, generator=gen).to(dev)
    faces = cubemap.cube_face_extrinsics(synthetic.trajectory(1, seed=4).to(dev))                # [1, 6, 4, 4]
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]], device=dev).expand(1, 6, 3, 3)
    near, far = torch.full((1, 6), 0.1, device=dev), torch.full((1, 6), 100.0, device=dev)
    dec = DecoderSplattingCUDA().to(dev)
    c2e = cubemap.Cube2Equirec(F, 2 * F, 4 * F).to(dev)

    def tail(means, cov, harm):
        g = Gaussians(means.reshape(b, v * r, 3), cov.reshape(b, v * r, 3, 3), harm.reshape(b, v * r, 3, d_sh), opac.to(dev))
        return mse_loss(c2e.from_faces(dec(g, faces, K, near, far, (F, F)).color), target)

    # fused adapter
    mod = adapter.GaussianAdapterERP(adapter.GaussianAdapterERPCfg(0.5, 15.0, deg)).to(dev)
    d_g = depths.to(dev).reshape(b, v, r, 1, 1).requires_grad_()
    r_g = raw.to(dev).reshape(b, v, r, 1, 1, -1).requires_grad_()
    out = mod("hm3d", ext.to(dev).reshape(b, v, 1, 1, 1, 4, 4), d_g, torch.ones(b, v, r, 1, 1, device=dev), r_g, (h, w))
    loss_fused = tail(out.means, out.covariances, out.harmonics)
    loss_fused.backward()

    # the reference's op sequence in float64 torch on the CPU, handed to the same decoder chain
    dd, rr = depths.double().requires_grad_(), raw.double().requires_grad_()
    m_ref, c_ref, s_ref = _torch_adapter(ext.double(), dd, rr, h, w, 0.5, 15.0, deg, mod.sh_mask.cpu().double())
    loss_ref = tail(m_ref.float().to(dev), c_ref.float().to(dev), s_ref.float().to(dev))
    loss_ref.backward()

    lf, lr = float(loss_fused.detach()), float(loss_ref.detach())
    assert lf > 0 and abs(lf - lr) <= 1e-5 * lr
    assert float(rr.grad.abs().max()) > 0 and float(dd.grad.abs().max()) > 0
    assert rel_l2(r_g.grad.cpu().reshape(b * v, r, -1).numpy(), rr.grad.numpy()) < 1e-4
    assert rel_l2(d_g.grad.cpu().reshape(b * v, r).numpy(), dd.grad.numpy()) < 1e-4

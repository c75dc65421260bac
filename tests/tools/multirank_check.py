"""Multi-rank GPU check, launched by tests/test_gpu_multirank.py under torchrun (one rank per GPU, NCCL):

  1. V views of one replicated scene, sharded round-robin over the ranks and gathered: bit-identical to the same V
     views rendered by ONE GPU (SURVEY.md sec. 4b "8-rank view sharding reproduces the single-GPU images bit-for-bit").
  2. per-rank gradients of the per-view losses + parallel.all_reduce_gradients == the single-GPU sum over all views.
  3. parallel.broadcast_scene: the scene built on rank 0 only arrives bit-identical on every rank.
  4. parallel.upload_and_broadcast_scene: the same from pinned host tensors on rank 0, uploaded in small chunks that are
     broadcast while the next chunk uploads.
Prints 'MULTIRANK OK <world>' on rank 0; any mismatch raises."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from splatter360_b200 import camera, parallel, synthetic  # noqa: E402
from splatter360_b200 import rasterizer as R  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    H, W, n, V = 128, 256, 20000, 2 * world + 1
    # the scene exists on rank 0 only and is broadcast
    sc = synthetic.random_cloud_scene(n, seed=3, ref_width=256, device=dev)
    tensors = [sc.means.contiguous(), synthetic.cov3x3_to_cov6(sc.covariances).contiguous(), sc.opacities.contiguous(),
               sc.harmonics.permute(0, 2, 1).contiguous()]
    ref = [t.clone() for t in tensors]
    if rank != 0:
        for t in tensors:
            t.zero_()
    parallel.broadcast_scene(tensors, src=0)
    for a, b in zip(tensors, ref):
        assert torch.equal(a, b), "broadcast_scene changed the scene"
    host = [t.cpu().pin_memory() for t in ref] if rank == 0 else None
    bufs = [torch.zeros_like(t) for t in ref]
    parallel.upload_and_broadcast_scene(host, bufs, src=0, chunk_bytes=100_000)   # many ragged chunks
    for a, b in zip(bufs, ref):
        assert torch.equal(a, b), "upload_and_broadcast_scene changed the scene"
    means, cov6, opac, shs = tensors
    cams = camera.erp_camera(synthetic.trajectory(V, seed=5).to(dev))
    dL = torch.randn(V, 3, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(9))

    def settings(j):
        return R.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev), scale_modifier=1.0,
            viewmatrix=cams.view_matrix[j], projmatrix=cams.full_projection[j], sh_degree=4, campos=cams.campos[j],
            prefiltered=False, debug=False, projection="erp")

    def render(j):
        color, st = R.forward_raw(settings(j), means, cov6, opac, shs, None)
        g = R.backward_raw(settings(j), means, cov6, opac, shs, None, st, dL[j])
        return color, g

    mine = parallel.shard_views(V)
    outs = [render(j) for j in mine]
    local_imgs = torch.stack([o[0] for o in outs]) if outs else torch.zeros(0, 3, H, W, device=dev)
    full = parallel.gather_views(local_imgs, V)
    keys = ("means3D", "cov3D", "opacities", "shs")
    gsum = [sum(o[1][k] for o in outs) if outs else torch.zeros_like(t) for k, t in zip(keys, (means, cov6, opac[:, None], shs))]
    parallel.all_reduce_gradients(gsum)
    if rank == 0:
        single = [render(j) for j in range(V)]
        for j in range(V):
            assert torch.equal(full[j], single[j][0]), f"view {j}: sharded image differs from the single-GPU image"
        for k, g in zip(keys, gsum):
            want = sum(o[1][k] for o in single)
            err = float((g - want).norm() / want.norm().clamp_min(1e-30))
            assert err < 2e-6, (k, err)   # same per-view values, different summation order over the views
        print(f"MULTIRANK OK {world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Host-side logic on CPU: the Python surface mirrors diff_gaussian_rasterization's error behaviour, the product
path refuses to run without CUDA (no silent fallback), and the view sharding / loss all-reduce work over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _settings(**kw):
    from splatter360_b200.rasterizer import GaussianRasterizationSettings
    d = dict(image_height=16, image_width=16, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3), scale_modifier=1.0,
             viewmatrix=torch.eye(4), projmatrix=torch.eye(4), sh_degree=0, campos=torch.zeros(3), prefiltered=False,
             debug=False)
    d.update(kw)
    return GaussianRasterizationSettings(**d)


def test_settings_surface_matches_upstream():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    f = GaussianRasterizationSettings._fields
    assert f[:12] == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
                      "projmatrix", "sh_degree", "campos", "prefiltered", "debug")
    s = _settings()
    assert s.projection == "pinhole" and s.near_cull == 0.2 and s.fov_clamp == 1.3 and s.lowpass == 0.3
    assert isinstance(GaussianRasterizer(s), torch.nn.Module)


def test_exactly_one_of_errors_like_upstream():
    from splatter360_b200.rasterizer import GaussianRasterizer
    r = GaussianRasterizer(_settings())
    m = torch.zeros(4, 3); o = torch.ones(4, 1); c6 = torch.zeros(4, 6); col = torch.zeros(4, 3); sh = torch.zeros(4, 1, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=m, means2D=m, opacities=o, cov3D_precomp=c6)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=m, means2D=m, opacities=o, shs=sh, colors_precomp=col, cov3D_precomp=c6)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=m, means2D=m, opacities=o, colors_precomp=col)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=m, means2D=m, opacities=o, colors_precomp=col, scales=torch.ones(4, 3), rotations=torch.ones(4, 4),
          cov3D_precomp=c6)


def test_no_cpu_fallback():
    """CPU tensors must fail loudly: the product path is CUDA only."""
    from splatter360_b200.rasterizer import GaussianRasterizer
    r = GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        r(means3D=m, means2D=m, opacities=torch.ones(4, 1), colors_precomp=torch.zeros(4, 3), cov3D_precomp=torch.zeros(4, 6))
    with pytest.raises(RuntimeError, match="CUDA"):
        r.markVisible(m)


def test_bad_projection_and_erp_width():
    from splatter360_b200 import rasterizer
    with pytest.raises(ValueError):
        rasterizer._make_view(_settings(projection="fisheye"), 1, 0, torch.device("cpu"))
    with pytest.raises(ValueError):
        rasterizer._make_view(_settings(projection="erp", image_width=40), 1, 0, torch.device("cpu"))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under splatter360_b200/ may import it."""
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "splatter360_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_shard_views_partition():
    from splatter360_b200.parallel import shard_views
    for n in (0, 1, 7, 8, 32, 101):
        for ws in (1, 2, 3, 8):
            seen = sorted(i for r in range(ws) for i in shard_views(n, r, ws))
            assert seen == list(range(n))
            sizes = [len(shard_views(n, r, ws)) for r in range(ws)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, ws, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from splatter360_b200 import parallel
    n_views = 5
    idx, imgs = parallel.render_views_sharded(lambda i: torch.full((3, 2, 2), float(i)), n_views)
    local = torch.stack(imgs) if imgs else torch.zeros(0, 3, 2, 2)
    full = parallel.gather_views(local, n_views)
    loss = sum((img ** 2).mean() for img in imgs) if imgs else torch.zeros(())
    total = parallel.all_reduce_loss(torch.as_tensor(loss, dtype=torch.float32))
    g = [torch.full((4, 3), float(rank + 1)), None]
    parallel.all_reduce_gradients(g)
    # asynchronous loss reducer: five steps through a ring of two slots, every step's sum is rank-independent
    red = parallel.AsyncLossReducer("cpu", depth=2)
    sums = []
    for step in range(5):
        red.submit(torch.tensor(float((rank + 1) * (step + 1))))
        if step % 2 == 1:
            sums.append(float(red.latest()))
    red.flush()
    sums.append(float(red.latest()))
    # replicated-scene broadcast (video path): rank 0 holds the scene, every rank ends up with it
    scene = [torch.arange(12, dtype=torch.float32).reshape(4, 3) * (1.0 if rank == 0 else 0.0), torch.full((4,), float(rank))]
    parallel.broadcast_scene(scene, src=0)
    sums.append(float(scene[0].sum()) + float(scene[1].sum()))
    q.put((rank, idx, full[:, 0, 0, 0].tolist(), float(total), g[0][0, 0].item(), sums))
    dist.destroy_process_group()


def test_world_size_2_gloo_sharding_and_loss_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    for r in res:
        assert r[2] == [0.0, 1.0, 2.0, 3.0, 4.0]          # every rank reassembles all views
        assert abs(r[3] - (0 + 1 + 4 + 9 + 16)) < 1e-5     # loss all-reduce = single-process sum
        assert r[4] == 3.0                                  # gradient all-reduce: 1 + 2
        assert r[5] == [6.0, 12.0, 15.0, 66.0]              # async reducer: (1 + 2) * step for steps 2, 4, 5; broadcast scene


def test_batched_decoder_groups_views_and_builds_the_reference_cameras(monkeypatch):
    """render_cuda_views (no GPU needed: the rasterizer call is intercepted): views of a batch item that share
    fov / near / far go through ONE rasterize_views pass in chunks of max_views_per_pass, a change of near starts a new
    pass, and the camera blocks handed over are exactly what the reference's per-view render_cuda builds
    (cuda_splatting.py:63-87, pinned against the reference in tests/test_golden.py)."""
    from splatter360_b200 import camera, cubemap, decoder, synthetic
    calls = []

    def fake(means3D, opacities, cov3D_precomp, settings, shs=None, colors_precomp=None):
        calls.append(settings)
        v = settings.viewmatrix.shape[0]
        out = torch.zeros(v, 3, settings.image_height, settings.image_width)
        return (out, torch.zeros(v, settings.image_height, settings.image_width)) if settings.depth_mode else out

    monkeypatch.setattr(decoder, "rasterize_views", fake)
    b, g = 2, 5
    faces = torch.stack([cubemap.cube_face_extrinsics(synthetic.target_pose(k)) for k in range(b)])       # [b,6,4,4]
    ext = torch.cat([faces, faces[:, :2]], dim=1)                                                         # 8 views
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 1.0]])[None, None].repeat(b, 8, 1, 1)
    near = torch.full((b, 8), 0.5); far = torch.full((b, 8), 20.0)
    near[1, 5:] = 0.25                                                                                    # item 1: views 5.. differ
    out = decoder.render_cuda_views(ext, K, near, far, (32, 32), torch.zeros(b, 3), torch.zeros(b, g, 3),
                                    torch.eye(3).expand(b, g, 3, 3), torch.zeros(b, g, 3, 25), torch.ones(b, g),
                                    max_views_per_pass=6)
    assert out.shape == (b, 8, 3, 32, 32)
    assert [s.viewmatrix.shape[0] for s in calls] == [6, 2, 5, 3]
    assert [s.scene_scale for s in calls] == [2.0, 2.0, 2.0, 4.0]
    assert all(s.sh_layout == 1 and s.cov_layout == 1 and s.projection == "pinhole" and s.sh_degree == 4 for s in calls)
    # camera blocks == the per-view construction of render_cuda
    for s, (i, j0) in zip(calls, [(0, 0), (0, 6), (1, 0), (1, 5)]):
        for k in range(s.viewmatrix.shape[0]):
            e = ext[i, j0 + k].clone()
            sc = 1 / near[i, j0 + k]
            e[:3, 3] *= sc
            cam = camera.pinhole_camera(e[None], K[i, j0 + k][None], (near[i, j0 + k] * sc)[None], (far[i, j0 + k] * sc)[None])
            assert torch.allclose(s.viewmatrix[k], cam.view_matrix[0], atol=1e-6)
            assert torch.allclose(s.projmatrix[k], cam.full_projection[0], atol=1e-6)
            assert torch.allclose(s.campos[k], cam.campos[0], atol=1e-6)
            assert abs(s.tanfovx - float(cam.tan_fov_x[0])) < 1e-6
    # fused depth request is forwarded with the unscaled near / far
    calls.clear()
    col, dep = decoder.render_cuda_views(ext[:1, :6], K[:1, :6], near[:1, :6], far[:1, :6], (32, 32), torch.zeros(1, 3),
                                         torch.zeros(1, g, 3), torch.eye(3).expand(1, g, 3, 3), torch.zeros(1, g, 3, 25),
                                         torch.ones(1, g), fused_depth_mode="depth")
    assert dep.shape == (1, 6, 32, 32) and calls[0].depth_mode == "depth" and calls[0].depth_near == 0.5 and calls[0].depth_far == 20.0


def test_batched_path_refuses_cpu_tensors_and_too_many_views():
    from splatter360_b200 import rasterizer
    s = _settings(viewmatrix=torch.eye(4)[None].repeat(2, 1, 1), projmatrix=torch.eye(4)[None].repeat(2, 1, 1),
                  campos=torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        rasterizer.rasterize_views(torch.zeros(4, 3), torch.ones(4, 1), torch.zeros(4, 6), s, colors_precomp=torch.zeros(4, 3))
    from splatter360_b200.cubemap import Cube2Equirec
    with pytest.raises(RuntimeError, match="CUDA"):
        Cube2Equirec(8, 16, 32)(torch.zeros(1, 3, 8, 48))


def test_only_test_infrastructure_touches_the_oracle():
    """oracle/ may be used by tests/, __graft_entry__.smoke() and bench.py only (its header says so): nothing in the
    package, the drop-in module or tools/ imports or loads it."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for sub in ("splatter360_b200", "diff_gaussian_rasterization", "tools"):
        for dp, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if f.endswith((".py", ".sh", ".cu", ".cuh")):
                    src = open(os.path.join(dp, f)).read()
                    if re.search(r"^\s*(import|from)\s+oracle\b|liboracle|run_oracle", src, flags=re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_debug_flag_dumps_a_snapshot_like_upstream(tmp_path, monkeypatch):
    """settings.debug=True: an exception in the native call writes snapshot_fw.dump (CPU copies of the arguments) and is
    re-raised (upstream's debug convention); without debug nothing is written."""
    from splatter360_b200.rasterizer import GaussianRasterizer
    monkeypatch.chdir(tmp_path)
    m = torch.zeros(4, 3)
    kw = dict(means3D=m, means2D=m, opacities=torch.ones(4, 1), colors_precomp=torch.zeros(4, 3), cov3D_precomp=torch.zeros(4, 6))
    with pytest.raises(RuntimeError, match="CUDA"):
        GaussianRasterizer(_settings())(**kw)
    assert not os.path.exists(tmp_path / "snapshot_fw.dump")
    with pytest.raises(RuntimeError, match="CUDA"):
        GaussianRasterizer(_settings(debug=True))(**kw)
    dump = torch.load(tmp_path / "snapshot_fw.dump", weights_only=False)
    assert torch.equal(dump[0], m) and dump[3] is None and torch.equal(dump[4], torch.zeros(4, 3))


def test_shard_view_groups_keeps_cube_faces_together():
    from splatter360_b200.parallel import shard_view_groups
    # 3 target panoramas x 6 faces over 2 ranks: whole panoramas per rank, every view exactly once
    parts = [shard_view_groups(18, 6, r, 2) for r in range(2)]
    assert parts[0] == list(range(0, 6)) + list(range(12, 18)) and parts[1] == list(range(6, 12))
    # 8 ranks, 32 frames of 1 view: identical to round-robin
    from splatter360_b200.parallel import shard_views
    assert all(shard_view_groups(32, 1, r, 8) == shard_views(32, r, 8) for r in range(8))
    # partial trailing group, more ranks than groups
    parts = [shard_view_groups(14, 6, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == list(range(14)) and parts[2] == [12, 13] and parts[3] == []
    with pytest.raises(ValueError):
        shard_view_groups(6, 0, 0, 1)


def test_frozen_capacity_tracker_and_numa_helpers_without_a_gpu():
    """CapacityTracker.freeze (what graph.GraphedAutogradStep relies on): fixed capacities, no polling, the counters of the
    captured call are only remembered; the NUMA helpers change nothing when the topology is not exposed."""
    import os
    from splatter360_b200 import io
    from splatter360_b200 import rasterizer as R
    t = R.CapacityTracker(margin=1.25)
    with pytest.raises(RuntimeError):
        t.freeze()                                   # nothing observed yet
    t._take(0, torch.tensor([1000, 0, 300, 0], dtype=torch.int32))
    assert t.capacity == max(t.min_capacity, 1251) and t.pair_capacity == max(t.min_capacity, 376)
    t.freeze()
    s = R.GaussianRasterizationSettings(
        image_height=16, image_width=16, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3), scale_modifier=1.0,
        viewmatrix=torch.eye(4), projmatrix=torch.eye(4), sh_degree=0, campos=torch.zeros(3), prefiltered=False, debug=False)
    assert t.settings(s).instance_capacity == t.capacity and t.settings(s).pair_capacity is None
    both = t.settings(s, pairs=True)
    assert (both.instance_capacity, both.pair_capacity) == (t.capacity, t.pair_capacity)
    t.observe(torch.tensor([5000, 1, 10, 0], dtype=torch.int32))     # frozen: remembered, not copied, capacity untouched
    assert t.frozen_overflowed() and t.capacity == max(t.min_capacity, 1251) and not t._pending
    t.observe(torch.tensor([900, 0, 10, 0], dtype=torch.int32))
    assert not t.frozen_overflowed()
    t.unfreeze()
    assert not t.frozen and not t.frozen_overflowed()
    before = os.sched_getaffinity(0)
    assert io.gpu_numa_node("cuda:0") is None or isinstance(io.gpu_numa_node("cuda:0"), int)
    if not torch.cuda.is_available():
        assert io.bind_to_gpu_numa_node("cuda:0") is None and os.sched_getaffinity(0) == before

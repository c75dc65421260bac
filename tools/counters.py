"""Work counters of the render kernels on the bench workload (config 3): needs a library built with -DS360_COUNTERS=1
(python tools/build_variants.py counters), loaded through S360_LIB.  Prints one JSON line."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from splatter360_b200 import _lib, camera, synthetic
from splatter360_b200 import rasterizer as R

H, W = 512, 1024
dev = "cuda"
lib = _lib.load()
sc = synthetic.pixel_aligned_scene(H, W, sh_degree=4, seed=1237, device=dev)
means = sc.means.contiguous(); cov6 = synthetic.cov3x3_to_cov6(sc.covariances).contiguous()
opac = sc.opacities.contiguous(); shs = sc.harmonics.permute(0, 2, 1).contiguous()
cams = camera.erp_camera(synthetic.trajectory(8, seed=0).to(dev))
s = R.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3, device=dev),
                                    scale_modifier=1.0, viewmatrix=cams.view_matrix[3], projmatrix=cams.full_projection[3],
                                    sh_degree=4, campos=cams.campos[3], prefiltered=False, debug=False, projection="erp")
buf = (ctypes.c_uint64 * 16)()
lib.s360_debug_counters(ctypes.cast(buf, ctypes.c_void_p), 1, None)
color, st = R.forward_raw(s, means, cov6, opac, shs, None)
R.backward_raw(s, means, cov6, opac, shs, None, st, torch.randn(3, H, W, device=dev))
torch.cuda.synchronize()
rc = lib.s360_debug_counters(ctypes.cast(buf, ctypes.c_void_p), 0, None)
c = [int(x) for x in buf]
N = st.num_rendered
out = {"rc": rc, "P": means.shape[0], "N_instances": N,
       "fwd": {"warp_chunks": c[0], "tests": c[1], "survivors": c[2], "survivors_some_pixel_takes": c[3], "pairs_taken": c[4],
               "list_entries": c[5], "survivors_per_instance": c[2] / max(N, 1), "tests_per_instance": c[1] / max(N, 1),
               "taken_fraction_of_evaluated_pairs": c[4] / max(64 * c[2], 1)},
       "bwd": {"warp_chunks": c[8], "tests": c[9], "survivors": c[10], "survivors_some_pixel_takes": c[11], "pairs_taken": c[12],
               "survivors_per_instance": c[10] / max(N, 1), "tests_per_instance": c[9] / max(N, 1),
               "taken_fraction_of_evaluated_pairs": c[12] / max(64 * c[10], 1)}}
print(json.dumps(out))

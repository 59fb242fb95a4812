"""Development tool: runs the CPU model of the raster-backward enumeration (scripts/proto/bwd_proto.c) against the
oracle's backward_pixel_map on the adversarial inputs of tests/test_raster_stress_gpu.py and on synthetic clips,
and prints the work counters."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import build as ob, nmr  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
proto = ctypes.CDLL(os.path.join(HERE, "libbwdproto.so"))
P = ctypes.c_void_p
proto.proto_pixel_map_bwd.argtypes = [P, P, P, P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                      ctypes.c_float, P, P]
proto.proto_pixel_map_bwd.restype = None
proto.proto_skipped_insweeps.argtypes = [P, P, P, P, ctypes.c_int, ctypes.c_int, ctypes.c_int]
proto.proto_skipped_insweeps.restype = ctypes.c_long
proto.proto_set_runs.argtypes = [ctypes.c_int]
NAMES = ["candidates", "matched", "cand_edge_evals", "steep_tasks", "steep_crossings", "boundary_faces", "front_faces",
         "in_crossings", "in_crossings_nonzero", "out_pixels", "in_pixels", "matched_prev_pixel", "irregular_faces", "out_items", "in_items"]
SMAX = 128.0


def compare(ndc, faces_b, image_size, aa, g_alpha_R, label, eps=1e-4):
    """ndc [B,V,3] torch, faces_b [B,F,3] long, g_alpha_R gradient at the output resolution."""
    is_ = image_size * 2 if aa else image_size
    f2 = torch.cat((faces_b, faces_b[:, :, [2, 1, 0]]), dim=1)
    fv = nmr.vertices_to_faces(ndc, f2).contiguous().float()
    B, nf = fv.shape[:2]
    lib = ob.lib()
    fi = torch.empty(B, is_, is_, dtype=torch.int32)
    lib.nmr_face_index_map(fv.data_ptr(), B, nf, is_, 0.1, 100.0, fi.data_ptr(), None)
    alpha = (fi >= 0).float()
    # gradient at raster resolution (un-pool, un-flip)
    g = g_alpha_R
    if aa:
        g = g.repeat_interleave(2, 1).repeat_interleave(2, 2) / 4
    g = torch.flip(g, dims=(1,)).contiguous().float()
    ref = torch.zeros_like(fv)
    lib.nmr_pixel_map_bwd(fv.data_ptr(), fi.data_ptr(), alpha.data_ptr(), g.data_ptr(), B, nf, is_, eps, ref.data_ptr())
    out = torch.zeros_like(fv)
    stats = np.zeros(16, dtype=np.int64)
    proto.proto_pixel_map_bwd(fv.data_ptr(), fi.data_ptr(), alpha.data_ptr(), g.data_ptr(), B, nf, is_, eps, SMAX,
                              out.data_ptr(), stats.ctypes.data)
    bad = proto.proto_skipped_insweeps(fv.data_ptr(), fi.data_ptr(), alpha.data_ptr(), g.data_ptr(), B, nf, is_)
    scale = float(ref.abs().max())
    err = float((out - ref).abs().max())
    st = {n: int(v) for n, v in zip(NAMES, stats)}
    ok = (scale == 0 and err == 0) or err <= (2e-5 if os.environ.get('PROTO_RUNS', '0') != '0' else 2e-6) * max(scale, 1e-30)
    print(f"{'OK ' if ok and bad == 0 else 'BAD'} {label}: scale {scale:.3e} err {err:.3e} skipped-in-sweep px {bad}  "
          f"{ {k: v for k, v in st.items()} }")
    return ok and bad == 0, st


def stress():
    import test_raster_stress_gpu as t
    results = []

    def _check(ndc, faces, image_size, aa, with_grad=True, faces_batched=False):
        ndc = torch.as_tensor(ndc, dtype=torch.float32).contiguous()
        faces = torch.as_tensor(faces)
        fb = faces.long() if faces_batched else faces.long()[None].repeat(ndc.shape[0], 1, 1)
        a_ref, fi_ref, _ = t._oracle(ndc, fb, image_size, aa)
        target = torch.roll(a_ref, shifts=(5, -7), dims=(1, 2)).round()
        g_alpha = 2 * (a_ref - target) / target[0].numel()
        rng = np.random.default_rng(0)
        for noise in (False, True):
            ga = g_alpha * torch.from_numpy(rng.uniform(0.5, 1.5, size=g_alpha.shape).astype(np.float32)) if noise else g_alpha
            results.append(compare(ndc, fb, image_size, aa, ga, f"{_check.name} aa={aa} noise={noise}")[0])

    t._check = _check
    for name in ("test_random_triangle_soup_with_coincident_layers", "test_vertices_on_pixel_centres_and_integer_coordinates",
                 "test_degenerate_subpixel_and_sliver_faces", "test_faces_whose_two_windings_are_both_front_facing",
                 "test_near_plane_behind_camera_and_off_screen", "test_non_power_of_two_raster_and_per_image_faces",
                 "test_interpenetrating_closed_meshes"):
        _check.name = name
        fn = getattr(t, name)
        if "aa" in fn.__code__.co_varnames[:fn.__code__.co_argcount]:
            for aa in (True, False):
                fn(aa)
        else:
            fn()
    return all(results)


def cpu_render(verts, faces, K):
    ndc = nmr.projection(torch.from_numpy(verts), torch.from_numpy(K), torch.eye(3)[None], torch.zeros(1, 3),
                         torch.zeros(1, 5), 1)
    f = torch.from_numpy(faces.astype(np.int64))[None].repeat(ndc.shape[0], 1, 1)
    f2 = torch.cat((f, f[:, :, [2, 1, 0]]), dim=1)
    return nmr.rasterize_silhouettes(nmr.vertices_to_faces(ndc, f2), 256, False).numpy()


def workload(T=4, P=2, obj="ellipsoid500", seed=3000):
    from homan_b200 import synth
    from oracle import homan_ref  # noqa: F401
    clip = synth.make_clip(T, obj, seed=seed, render_fn=cpu_render)
    inits = synth.make_inits(clip, P, seed=seed)
    batch = synth.make_batch(clip, inits)
    asset = clip["asset"]
    ok = True
    for mesh in ("obj", "hand"):
        for state in ("init", "gt"):
            if mesh == "obj":
                R = batch["obj_R"] if state == "init" else np.repeat(clip["gt"]["obj_R"][None], P, 0)
                t_ = batch["obj_t"] if state == "init" else np.repeat(clip["gt"]["obj_t"][None], P, 0)
                verts = np.einsum("vk,ptkj->ptvj", clip["obj_verts_can"].astype(np.float64), R) + t_[:, :, None]
                faces, K, target = clip["obj_faces"], batch["K_roi_obj"], batch["target_masks_object"]
            else:
                pca = batch["pca"] if state == "init" else np.repeat(clip["gt"]["pca"][None], P, 0)
                R = batch["hand_R"] if state == "init" else np.repeat(clip["gt"]["hand_R"][None], P, 0)
                t_ = batch["hand_t"] if state == "init" else np.repeat(clip["gt"]["hand_t"][None], P, 0)
                hv, _ = synth.mano_forward_np(asset, pca.reshape(P * T, -1), batch["mano_rot"].reshape(P * T, 3),
                                              np.zeros((P * T, 10)))
                verts = np.einsum("ptvk,ptkj->ptvj", hv.reshape(P, T, 778, 3), R) + t_[:, :, None]
                faces, K, target = asset["f"], batch["K_roi_hand"], batch["target_masks_hand"]
            v = torch.from_numpy(verts.reshape(P * T, -1, 3).astype(np.float32))
            Kt = torch.from_numpy(K.reshape(P * T, 3, 3))
            ndc = nmr.projection(v, Kt, torch.eye(3)[None], torch.zeros(1, 3), torch.zeros(1, 5), 1).contiguous()
            fb = torch.from_numpy(faces.astype(np.int64))[None].repeat(P * T, 1, 1)
            f2 = torch.cat((fb, fb[:, :, [2, 1, 0]]), dim=1)
            rend = nmr.rasterize_silhouettes(nmr.vertices_to_faces(ndc, f2), 256, True)
            tg = torch.from_numpy(target.reshape(P * T, 256, 256))
            keep, ref = (tg >= 0).float(), (tg > 0).float()
            g = 2 * keep * (keep * rend - ref) / keep.sum()
            r, st = compare(ndc, fb, 256, True, g, f"{mesh}/{state} B={P * T}")
            ok &= r
            B = P * T
            print("   per image:", {k: round(v / B, 1) for k, v in st.items()})
    return ok


if __name__ == "__main__":
    proto.proto_set_runs(int(os.environ.get("PROTO_RUNS", "0")))
    ok = stress()
    ok &= workload()
    print("ALL OK" if ok else "FAILURES")

"""Measurement of the object-pose multi-init fitter: the reference's
find_optimal_pose workload - num_initializations = 2000 random rotations x 50 Adam iterations against one 256^2 mask
(/root/reference/homan/pose_optimization.py:219-383, defaults :226-227) - on homan_b200's PoseFitEngine, next to the
CPU oracle port (oracle/pose_ref.py) timed on a bounded sample of the same candidates.

    python scripts/bench_pose_init.py [--inits 2000] [--iters 50] [--out profiles/r01_pose_init.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--inits", type=int, default=2000)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--cpu-inits", type=int, default=16)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from homan_b200 import pose_optimization as po, synth
    rng = np.random.default_rng(7)
    verts, faces = synth.make_object("ellipsoid500")
    K = np.array([[600.0, 0, 320.0], [0, 600.0, 320.0], [0, 0, 1]], dtype=np.float32)
    R_gt = synth._random_rotation(rng).astype(np.float32)
    v_gt = verts.astype(np.float32) @ R_gt + np.array([0.03, -0.02, 0.55], dtype=np.float32)
    uv = synth.project_np(v_gt.astype(np.float64), K.astype(np.float64))
    lo, hi = uv.min(0), uv.max(0)
    bbox = np.array([lo[0], lo[1], hi[0] - lo[0], hi[1] - lo[1]], dtype=np.float32)
    x, y, b = synth._square_roi(uv)
    K_roi = po.get_K_crop_resize(torch.from_numpy(K)[None], torch.tensor([[x, y, x + b, y + b]]), [256])
    K_roi[:, :2] /= 256
    dev = torch.device("cuda")
    # target mask: product render of the ground truth (binary, anti-aliasing off) with an occluded band
    tgt = po.PoseFitEngine(verts, faces, np.zeros((256, 256), np.float32), K_roi, R_gt[None, :, :2].copy(),
                           np.array([[0.03, -0.02, 0.55]], np.float32), use_graph=False)
    tgt.evaluate()
    mask = (tgt.rb.alpha[0] > 0.5).float().cpu().numpy()
    mask[:, 100:118] = -1
    torch.manual_seed(0)
    rots = po.compute_random_rotations(args.inits, device=dev)
    trans = po.TCO_init_from_boxes_zup_autodepth(bbox, torch.matmul(torch.from_numpy(verts).float().to(dev)[None], rots),
                                                 torch.from_numpy(K).to(dev)[None]).unsqueeze(1)
    eng = po.PoseFitEngine(verts, faces, mask, K_roi, po.matrix_to_rot6d(rots), trans)
    eng.capture()
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        eng.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    launches = eng.gpu_launches_per_step
    # end to end through the public call (host arrays in, fitted candidates out)
    t0 = time.perf_counter()
    model = po.find_optimal_pose(verts, faces, mask, bbox, np.array([x, y, b, b], np.float32), (640, 640), K=K,
                                 num_iterations=args.iters, num_initializations=args.inits, rotations_init=rots)
    best = model.rotations.detach().cpu()
    t_e2e = time.perf_counter() - t0
    eng.evaluate()
    iou = float(eng.iou.max())
    # CPU oracle port on a bounded sample of the same candidates
    from oracle import build as obuild, pose_ref
    obuild.build()
    torch.set_num_threads(os.cpu_count())
    n = args.cpu_inits
    t0 = time.perf_counter()
    pose_ref.fit(verts, faces, mask, K_roi.cpu(), po.matrix_to_rot6d(rots)[:n].cpu(), trans[:n].cpu(), 2)
    sec_cpu = (time.perf_counter() - t0) / 2 * (args.inits / n)
    line = {"metric": "pose-init iterations/s (all candidates rendered + backprop + Adam)", "value": 1e3 / ms,
            "unit": "iters/s", "ms_per_iteration": ms, "candidate_iterations_per_s": args.inits * 1e3 / ms,
            "config": {"workload": "find_optimal_pose", "inits": args.inits, "iters": args.iters, "faces": int(faces.shape[0]),
                       "render": "256^2, anti-aliasing off", "cuda_graph": True, "launches_per_iteration": launches},
            "e2e": {"seconds_find_optimal_pose": t_e2e, "iters_per_s": args.iters / t_e2e},
            "best_iou_after_fit": iou,
            "cpu_baseline": {"value": 1.0 / sec_cpu, "unit": "iters/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"2 iterations of {n} of the {args.inits} candidates, scaled; oracle/pose_ref.py"}}
    print(json.dumps(line))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(line, fh, indent=1)


if __name__ == "__main__":
    main()

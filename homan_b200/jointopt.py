"""Drop-in for /root/reference/homan/jointopt.py::optimize_hand_object (same keywords, same return triple).

The loop body of the reference (jointopt.py:158-192: zero_grad, forward, weighting, backward, Adam with its
three learning-rate groups) is one CUDA-graph replay of the fused engine per iteration; the per-iteration
`.item()` logging of the reference becomes one device-side copy per iteration and a single read-back at the
end. When `images` are given, every `viz_step` iterations the fitted scene is rendered (RGB, sm_100a rasteriser)
over the frames and saved as JPEG, as the reference does; the gif / video writers are out of scope.
"""
import os
from collections import OrderedDict, defaultdict

import numpy as np
import torch

from .engine import NPART
from .homan import HOMan


def _save_viz(model, images, viz_folder, step, viz_len):
    """The periodic picture of /root/reference/homan/jointopt.py:159-177: frontal views side by side over the top-down
    views, at half size, as <viz_folder>/<step:08d>.jpg. (The gif / webm / mp4 writers around it are out of scope.)"""
    from PIL import Image
    from .visualize import visualize_hand_object
    with torch.no_grad():
        frontal, top_down = visualize_hand_object(model, images, dist=1, viz_len=viz_len)
    frontal = np.concatenate(list(frontal), 1)
    top_down = np.concatenate(list(top_down), 1)
    front_top = np.concatenate([frontal, top_down[:frontal.shape[0], :frontal.shape[1]]], 0)
    front_top = Image.fromarray(front_top).resize((front_top.shape[1] // 2, front_top.shape[0] // 2), Image.BILINEAR)
    path = os.path.join(viz_folder, f"{step:08d}.jpg")
    front_top.save(path)
    return path


def optimize_hand_object(person_parameters, object_parameters, class_name="default", objvertices=None, objfaces=None,
                         loss_weights=None, num_iterations=400, lr=1e-2, images=None, viz_step=10, viz_folder="tmp",
                         camintr=None, hand_proj_mode="persp", optimize_mano=False, optimize_mano_beta=True,
                         optimize_object_scale=False, state_dict=None, fps=24, viz_len=7, image_size=640,
                         frames_per_problem=None, mano_asset=None):
    os.makedirs(viz_folder, exist_ok=True)
    cat = lambda seq, key: torch.cat([torch.as_tensor(p[key]) for p in seq])  # noqa: E731
    model = HOMan(
        hand_sides=person_parameters[0]["hand_side"],
        translations_object=cat(object_parameters, "translations"), rotations_object=cat(object_parameters, "rotations"),
        verts_object_og=torch.as_tensor(np.asarray(objvertices)), faces_object=torch.as_tensor(np.asarray(objfaces)),
        target_masks_object=cat(object_parameters, "target_masks"), target_masks_hand=cat(person_parameters, "target_masks"),
        verts_hand_og=cat(person_parameters, "verts"), ref_verts2d_hand=cat(person_parameters, "verts2d"),
        mano_trans=cat(person_parameters, "mano_trans"), mano_rot=cat(person_parameters, "mano_rot"),
        mano_pca_pose=cat(person_parameters, "mano_pca_pose"), mano_betas=cat(person_parameters, "mano_betas"),
        translations_hand=cat(person_parameters, "translations"), rotations_hand=cat(person_parameters, "rotations"),
        faces_hand=torch.as_tensor(person_parameters[0]["faces"]),
        masks_object=torch.cat([torch.as_tensor(o["full_mask"])[None] for o in object_parameters]),
        masks_hand=cat(person_parameters, "masks"), cams_hand=cat(person_parameters, "cams"),
        camintr_rois_object=torch.cat([torch.as_tensor(o["K_roi"])[:, 0] for o in object_parameters]),
        camintr_rois_hand=cat(person_parameters, "K_roi"), camintr=camintr, class_name=class_name, int_scale_init=1,
        hand_proj_mode=hand_proj_mode, optimize_mano=optimize_mano, optimize_mano_beta=optimize_mano_beta,
        optimize_object_scale=optimize_object_scale, image_size=image_size, frames_per_problem=frames_per_problem,
        mano_asset=mano_asset, loss_weights=loss_weights, lr=lr)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=False)
    eng = model.engine
    hist = torch.zeros(num_iterations, eng.P, NPART, device=eng.device)
    tot = torch.zeros(num_iterations, eng.P, device=eng.device)
    imgs = OrderedDict()
    for it in range(num_iterations):
        if images is not None and viz_step and it % viz_step == 0:   # jointopt.py:159-177
            imgs[it] = _save_viz(model, images, viz_folder, it, viz_len)
        eng.step()
        hist[it].copy_(eng.losses)
        tot[it].copy_(eng.total)
    torch.cuda.synchronize()
    loss_evolution = defaultdict(list)
    hist_c, tot_c = hist.cpu(), tot.cpu().numpy()
    for it in range(num_iterations):
        for k, v in eng.loss_dict(hist_c[it]).items():
            loss_evolution[k].append(float(v.sum()))
        for k, v in eng.metric_dict(hist_c[it]).items():
            loss_evolution[k].append(float(v.max() if k == "handobj_maxdist" else v.mean()))
        loss_evolution["loss"].append(float(tot_c[it].sum()))
    model.loss_per_problem = tot_c  # [iterations, P] (problem-axis extension)
    return model, dict(loss_evolution), imgs

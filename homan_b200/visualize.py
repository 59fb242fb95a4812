"""Visualisation of a fit: the reference's `visualize_hand_object`
(/root/reference/homan/visualize.py:44-128) on the RGB render of the sm_100a rasteriser - the fitted hand and object
drawn over the input frames ("frontal") and the same scene seen rotated ("top-down"), as
/root/reference/homan/jointopt.py:159-177 writes them every `viz_step` iterations. Forward only, GPU only."""
import numpy as np
import torch

# /root/reference/homan/utils/nmr_renderer.py:7-23
COLORS = {"blue": [0.65098039, 0.74117647, 0.85882353], "pink": [0.9, 0.7, 0.7], "green": [153 / 255.0, 216 / 255.0, 201 / 255.0],
          "red": [251 / 255.0, 128 / 255.0, 114 / 255.0], "gold": [240 / 255, 200 / 255, 0], "grey": [204 / 255, 204 / 255, 204 / 255],
          "white": [1, 1, 1]}


def rot_points(points, centers=None, axisang=(0, 1, 1)):
    """libyana.lib3d.trans3d.rot_points (un-vendored, recalled: parity unpinned): rotates every point set [B,N,3] about
    its centroid by the axis-angle vector `axisang`."""
    pts = points.float()
    if centers is None:
        centers = pts.mean(1, keepdim=True)
    a = torch.as_tensor(axisang, dtype=torch.float32, device=pts.device)
    angle = a.norm()
    k = a / angle.clamp_min(1e-12)
    K = torch.tensor([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]], device=pts.device)
    R = torch.eye(3, device=pts.device) + torch.sin(angle) * K + (1 - torch.cos(angle)) * (K @ K)
    return (pts - centers) @ R.t() + centers


def visualize_hand_object(model, images, verts_hand_gt=None, verts_object_gt=None, dist=3, viz_len=7, init=False,
                          gt_only=False, image_size=640, max_in_batch=2):
    """-> (frontal [N,H,W,3] uint8: the render composited over `images`, top_down [N,S,S,3] uint8: the rotated view).
    With `verts_hand_gt` / `verts_object_gt` the ground truth is drawn next to the fit (green / blue), `gt_only` draws it
    alone, `init` draws the initialisation instead of the current fit (homan/visualize.py:54-76,108-128)."""
    def draw(rotate):
        if gt_only:
            return model.render_gt(model.renderer, verts_hand_gt=verts_hand_gt, verts_object_gt=verts_object_gt,
                                   rotate=rotate, viz_len=viz_len, max_in_batch=max_in_batch)
        if verts_hand_gt is None:
            return model.render(model.renderer, rotate=rotate, viz_len=viz_len, max_in_batch=max_in_batch)
        return model.render_with_gt(model.renderer, verts_hand_gt=verts_hand_gt, verts_object_gt=verts_object_gt,
                                    rotate=rotate, viz_len=viz_len, init=init, max_in_batch=max_in_batch)

    rends, masks = draw(False)
    new_images = []
    for image, rend, mask in zip(images, rends, masks):
        image = np.asarray(image, dtype=np.float32)
        if image.max() > 1:
            image = image / 255.0
        h, w, _ = image.shape
        L = max(h, w)
        new_image = np.pad(image.copy(), ((0, L - h), (0, L - w), (0, 0)))
        if new_image.shape[:2] != mask.shape:   # frames of another size than the render: nearest resampling of the render
            ys = (np.arange(L) * mask.shape[0] // L).clip(0, mask.shape[0] - 1)
            xs = (np.arange(L) * mask.shape[1] // L).clip(0, mask.shape[1] - 1)
            rend, mask = rend[ys][:, xs], mask[ys][:, xs]
        new_image[mask] = rend[mask]
        new_images.append((new_image[:h, :w] * 255).astype(np.uint8))
    top_down, _ = draw(True)
    return np.stack(new_images), (top_down * 255).astype(np.uint8)

"""ORACLE (test infrastructure; never imported by the product).

Runs the UNMODIFIED reference Python (/root/reference/homan/*.py) on CPU in THIS container:
installs stand-ins for the four un-vendored third-party packages (`neural_renderer`, `sdf`,
`mano`, `libyana`) plus import-only stubs (`trimesh`, `detectron2`, `matplotlib`), neutralises
`.cuda()` and `torch.cuda.FloatTensor`, and prepares a scratch working directory holding the
relative-path assets the reference loads (`local_data/closed_fmano.npy`,
/root/reference/homan/lossutils.py:15; `extra_data/mano/MANO_*.pkl`, homan/homan.py:70).

Used only by `scripts/make_golden.py` and by tests that skip when /root/reference is absent
(it never exists on the GPU box).
"""
import os
import pickle
import sys
import types

import numpy as np
import torch

from . import libyana_min, mano_layer, nmr, sdfmod

REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "homan"))


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, mod)
    return mod


def _noop(*args, **kwargs):
    return None


_installed = False


def install(scratch_dir, mano_assets):
    """mano_assets: {"right": asset_dict, "left": asset_dict}. Returns the imported `homan` package."""
    global _installed
    os.makedirs(os.path.join(scratch_dir, "local_data"), exist_ok=True)
    os.makedirs(os.path.join(scratch_dir, "extra_data", "mano"), exist_ok=True)
    np.save(os.path.join(scratch_dir, "local_data", "closed_fmano.npy"),
            mano_assets["right"]["closed_faces"].astype(np.int64))
    for side, fname in (("right", "MANO_RIGHT.pkl"), ("left", "MANO_LEFT.pkl")):
        with open(os.path.join(scratch_dir, "extra_data", "mano", fname), "wb") as fh:
            pickle.dump(mano_assets[side], fh)
    os.chdir(scratch_dir)
    if _installed:
        return sys.modules["homan"]

    # --- .cuda() neutralisation (reference hard-codes CUDA: homan/homan.py:134,153,157,164-165)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.LongTensor = torch.LongTensor

    # --- neural_renderer
    nr_mod = _module("neural_renderer", Renderer=nmr.Renderer, projection=nmr.projection,
                     vertices_to_faces=nmr.vertices_to_faces, rasterize_silhouettes=nmr.rasterize_silhouettes)
    _module("neural_renderer.renderer", Renderer=nmr.Renderer)
    nr_mod.renderer = sys.modules["neural_renderer.renderer"]
    # --- sdf
    _module("sdf", SDF=sdfmod.SDF)
    # --- mano
    _module("mano")
    _module("mano.model", load=mano_layer.load)
    # --- libyana
    _module("libyana")
    _module("libyana.conversions")
    _module("libyana.conversions.npt", tensorify=libyana_min.tensorify, numpify=libyana_min.numpify)
    _module("libyana.vidutils")
    _module("libyana.vidutils.np2vid", make_video=_noop)
    _module("libyana.lib3d")
    _module("libyana.lib3d.trans3d", rot_points=lambda pts, *a, **k: pts)
    _module("libyana.lib3d.kcrop", get_K_crop_resize=libyana_min.get_K_crop_resize)
    _module("libyana.verify")
    _module("libyana.verify.checkshape", check_shape=libyana_min.check_shape)
    _module("libyana.camutils")
    _module("libyana.camutils.project", batch_proj2d=libyana_min.batch_proj2d)
    _module("libyana.camutils.camconvs", batch_weakcam2persptrans=_noop)
    _module("libyana.metrics")
    _module("libyana.metrics.iou", batch_mask_iou=libyana_min.batch_mask_iou)
    _module("libyana.distutils", batch_pairwise_dist=libyana_min.batch_pairwise_dist)
    _module("libyana.visutils")
    _module("libyana.visutils.imagify", viz_imgrow=_noop, viz_pointsrow=_noop)
    _module("libyana.visutils.viz2d", visualize_joints_2d=_noop)
    _module("libyana.randomutils")
    _module("libyana.randomutils.setseeds", set_all_seeds=_noop)
    # --- import-only stubs
    _module("trimesh", load=_noop)
    _module("detectron2")
    _module("detectron2.structures")
    _module("detectron2.structures.boxes", BoxMode=types.SimpleNamespace(XYXY_ABS=0, XYWH_ABS=1, convert=_noop))
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            _module("matplotlib")
            _module("matplotlib.pyplot")
            _module("matplotlib.cm")

    sys.path.insert(0, REFERENCE_ROOT)
    import homan  # noqa: F401  (the unmodified reference package)
    import homan.jointopt as jointopt

    # visualisation / video writing is out of scope; the optimisation loop itself is untouched
    dummy = [np.zeros((4, 4, 3), dtype=np.uint8)]
    jointopt.visualize_hand_object = lambda *a, **k: (dummy, dummy)
    _installed = True
    return sys.modules["homan"]


class record_trajectory:
    """Observes (does not alter) the reference model: snapshots every nn.Parameter at each call of
    HOMan.forward, i.e. the parameters each iteration's losses were evaluated at."""

    def __init__(self):
        self.snapshots = []

    def __enter__(self):
        import homan.homan as hh
        self._orig = hh.HOMan.forward
        rec = self

        def forward(model, *a, **k):
            rec.snapshots.append({n: p.detach().clone().numpy() for n, p in model.named_parameters()})
            return rec._orig(model, *a, **k)

        hh.HOMan.forward = forward
        return self

    def __exit__(self, *exc):
        import homan.homan as hh
        hh.HOMan.forward = self._orig


def run_reference_fit(inputs, loss_weights, num_iterations, scratch_dir, lr=1e-2, **kwargs):
    """Calls the unmodified /root/reference/homan/jointopt.py::optimize_hand_object on CPU."""
    import homan.jointopt as jointopt
    model, loss_evolution, _ = jointopt.optimize_hand_object(
        person_parameters=inputs["person_parameters"],
        object_parameters=inputs["object_parameters"],
        objvertices=inputs["objvertices"],
        objfaces=inputs["objfaces"],
        camintr=inputs["camintr"],
        loss_weights=loss_weights,
        num_iterations=num_iterations,
        lr=lr,
        viz_step=10 ** 9,
        viz_folder=os.path.join(scratch_dir, "viz"),
        optimize_mano=kwargs.pop("optimize_mano", True),
        optimize_mano_beta=kwargs.pop("optimize_mano_beta", True),
        optimize_object_scale=kwargs.pop("optimize_object_scale", False),
        image_size=inputs.get("image_size", 640),
        **kwargs,
    )
    return model, loss_evolution

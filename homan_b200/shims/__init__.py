"""GPU drop-ins for the three un-vendored CUDA packages the reference imports (`neural_renderer`, `sdf`,
`mano`), backed by libhoman_b200.so. `install()` registers them under the names the reference imports, so
that the reference's own Python (homan/losses.py, homan/interactions/scenesdf.py, homan/manomodel.py) runs on
a B200 through these kernels."""
import sys
import types


def install():
    from . import mano_layer, neural_renderer, sdf
    sys.modules["neural_renderer"] = neural_renderer
    sys.modules["neural_renderer.renderer"] = neural_renderer.renderer
    sys.modules["sdf"] = sdf
    pkg = types.ModuleType("mano")
    pkg.model = mano_layer
    sys.modules["mano"] = pkg
    sys.modules["mano.model"] = mano_layer

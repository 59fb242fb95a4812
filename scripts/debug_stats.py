"""Development: per-image statistics of the raster backward's work (visible / boundary faces, task lengths)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from homan_b200 import synth
from homan_b200.engine import FitEngine
from homan_b200.workload import make_workload
asset = synth.make_mano_asset(0, "right")
batch, lw = make_workload("cfg3", mano_asset=asset)
eng = FitEngine(batch, lw, mano_asset=asset, use_graph=False)
for _ in range(3):
    eng.step()
torch.cuda.synchronize()
for name, rb in (("obj", eng.rb_obj), ("hand", eng.rb_hand)):
    B, F, S = rb.B, rb.F, rb.S
    fi = rb.face_index
    cov = (fi >= 0)
    print(name, "coverage", float(cov.float().mean()))
    vis = rb.face_vis.cpu().numpy().view(np.uint32)
    nvis = np.unpackbits(vis.view(np.uint8), axis=1).sum(1)
    print(name, "visible faces per image", nvis.mean(), "of", 2 * F)
    box = rb.bboxes.view(B, F, 8).cpu().numpy().view(np.int16).reshape(B, F, 4)
    w = (box[..., 2] - box[..., 0] + 1).clip(0); h = (box[..., 3] - box[..., 1] + 1).clip(0)
    print(name, "bbox w/h mean", w[w > 0].mean(), h[h > 0].mean(), "nonempty", (w > 0).mean())
    # boundary: bbox contains an uncovered pixel (integral image)
    c = cov[:8].cpu().numpy().astype(np.int32)
    ii = np.zeros((8, S + 1, S + 1), np.int64); ii[:, 1:, 1:] = c.cumsum(1).cumsum(2)
    nb = []
    for b in range(8):
        x0, y0, x1, y1 = [box[b, :, k].astype(np.int64) for k in range(4)]
        ok = x0 <= x1
        area = (x1 - x0 + 1) * (y1 - y0 + 1)
        s = ii[b, (y1 + 1).clip(0, S), (x1 + 1).clip(0, S)] - ii[b, y0.clip(0, S), (x1 + 1).clip(0, S)] - ii[b, (y1 + 1).clip(0, S), x0.clip(0, S)] + ii[b, y0.clip(0, S), x0.clip(0, S)]
        nb.append(((s < area) & ok).sum())
    print(name, "boundary faces per image (exact bbox test)", np.mean(nb), "of", F)
    rec = rb.records.view(-1)[B * F * 128:].view(B, F, 96).cpu().numpy()
    span = rec[:, :, 64:88].copy().view(np.uint32).reshape(B, F, 6)
    ln = ((span >> 12) & 0xfff).astype(np.int64) - (span & 0xfff).astype(np.int64) + 1
    ln = ln.clip(0)
    print(name, "lines per task mean", ln[ln > 0].mean(), "crossings per image", ln.sum() / B)
c = (eng.rb_hand.face_index[0] >= 0).cpu().numpy()
for y in range(0, 512, 8):
    print("".join("#" if c[y:y+8, x:x+8].all() else ("+" if c[y:y+8, x:x+8].any() else ".") for x in range(0, 512, 4)))
t = eng.target_hand[0].cpu().numpy()
print("target")
for y in range(0, 256, 4):
    print("".join("#" if (t[255-y-3:255-y+1, x:x+2] > 0).all() else ("-" if (t[255-y-3:255-y+1, x:x+2] < 0).any() else ("+" if (t[255-y-3:255-y+1, x:x+2] > 0).any() else ".")) for x in range(0, 256, 2)))

#!/usr/bin/env python
"""Benchmark of the hand-object fitting hot path (BASELINE.json: optimizer iters/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl ours|reference]

One "step" = one optimizer iteration (forward + backward + Adam) over one resident batch of
P inits x T frames (cfg3: 16 x 30 = 480 images, 1538-face hand + 500-face object, 256^2 silhouettes
rasterised at 512^2, silhouette + 2-D vertex + smoothness + interaction + PCA + SDF collision + contact
losses).  Each rank (one process per GPU, torchrun) fits its own clip: weak scaling, no collective in
the loop, one all_gather of the per-clip best init at the end.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of the reference's loop
(oracle/homan_ref.py) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "optimizer iters/sec (frames x inits rendered+backprop)"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def raster_bwd_algorithmic_bytes(B, faces, R=256):
    """per image 4 S^2 (face_index R) + 4 R^2 (grad_alpha R) + 36 nf (faces R) + 36 nf
    (grad_faces W), S = 2R, nf = 2 * faces (fill_back)."""
    S, nf = 2 * R, 2 * faces
    return B * (4 * S * S + 4 * R * R + 72 * nf)


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_reference_iteration_rate(workload, steps, warmup, batch, lw, asset):
    """Times the oracle port of the reference loop (one init = one clip of T frames, as the reference
    itself runs it) on the host cores. Returns seconds per iteration of ONE problem."""
    from oracle import build as obuild, homan_ref
    obuild.build()
    torch.set_num_threads(os.cpu_count())
    mano = homan_ref.ManoPca(asset)
    model = homan_ref.ClipModel(homan_ref.problem_slice(batch, 0), mano, asset["closed_faces"])
    opt = homan_ref.make_optimizer(model, 1e-2)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        losses, _ = model.forward(lw)
        total = sum(v * lw[k.replace("loss", "lw")] for k, v in losses.items())
        total.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times))


_LINE = []   # the JSON line of this process (rank 0), printed by main() once stdout is restored


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from homan_b200 import synth
    from homan_b200.workload import CONFIGS, loss_weights
    cfg = CONFIGS[args.workload]
    asset = synth.make_mano_asset(0, "right", mesh=args.hand_mesh)
    from oracle import build as obuild, nmr
    obuild.build()

    def render_fn(verts, faces, K):
        r = nmr.Renderer(image_size=256, K=torch.from_numpy(K), R=torch.eye(3)[None], t=torch.zeros(1, 3),
                         orig_size=1, anti_aliasing=False)
        f = torch.from_numpy(faces.astype(np.int32))[None].repeat(len(verts), 1, 1)
        return r(torch.from_numpy(verts), f, mode="silhouettes").numpy()

    obj = cfg["obj"] + (str(cfg["seed"]) if cfg["obj"].endswith("@") else "")
    clip = synth.make_clip(cfg["T"], obj, seed=cfg["seed"], mano_asset=asset, render_fn=render_fn)
    inits = synth.make_inits(clip, 1, seed=cfg["seed"])
    batch = synth.make_batch(clip, inits)
    lw = loss_weights(cfg["lw"])
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    sec = cpu_reference_iteration_rate(args.workload, steps, warmup, batch, lw, asset)
    P = cfg["P"] * cfg.get("clips", 1)   # problems of one GPU's batch
    value = 1.0 / (sec * P)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * P * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "inits": P, "frames": cfg["T"], "object": cfg["obj"],
                   "hand_mesh": args.hand_mesh, "losses": cfg["lw"], "render": "256^2 (512^2 raster, AA)"},
        "cpu_baseline": {"value": value, "unit": "iters/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{steps} timed iteration(s) of 1 of the {P} inits ({cfg['T']} frames), "
                                   f"time x {P} = one whole-batch iteration; oracle/homan_ref.py (torch CPU + OpenMP C)"},
        "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _LINE.append(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def kernel_breakdown(eng, iters=3):
    """Per-kernel device time (CUDA events on the launching stream) of one eager iteration."""
    import homan_b200._lib as L
    orig = L.call
    rec = {}
    stream = torch.cuda.current_stream()

    def timed(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        orig(name, *a)
        e1.record(stream)
        rec.setdefault(name, []).append((e0, e1))

    import homan_b200.engine as E
    import homan_b200.ops as O
    E.call = O.call = timed
    try:
        for _ in range(iters):
            eng._iteration()
        torch.cuda.synchronize()
    finally:
        E.call = O.call = orig
    out = {}
    for name, evs in rec.items():
        per_iter = len(evs) // iters
        ms = [a.elapsed_time(b) for a, b in evs[per_iter:]]  # skip the first iteration
        each = np.asarray(ms).reshape(iters - 1, per_iter).mean(0) * 1e3
        out[name] = {"launches_per_step": per_iter, "us_per_launch": 1e3 * float(np.mean(ms)),
                     "us_per_step": 1e3 * float(np.sum(ms)) / (iters - 1), "us_each": [round(float(x), 1) for x in each]}
    return out


def run_ours(args):
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from homan_b200 import synth
    from homan_b200.engine import FitEngine, NPART
    from homan_b200.workload import CONFIGS, make_workload
    cfg = CONFIGS[args.workload]
    asset = synth.make_mano_asset(0, "right", mesh=args.hand_mesh)
    multi = "clips" in cfg
    if multi:
        # cfg4: whole clips per GPU (a clip is never split), every clip with its own object and its 16 inits on the
        # same GPU, so the per-clip argmin is local; weak scaling: cfg["clips"] clips per rank (8 ranks = the 64 clips
        # of BASELINE.json config 4); the final all_gather carries (clip id, best init, best loss, winner parameters)
        n_local = cfg["clips"]
        clip_ids = list(range(rank * n_local, (rank + 1) * n_local))
        batch, lw = make_workload(args.workload, clips=clip_ids, mano_asset=asset)
    else:
        # weak scaling: every rank fits the same clip from its own block of P random initialisations (equal work per
        # GPU); the job's answer is the argmin over all N*P inits, gathered once at the end
        n_local, clip_ids = 1, [rank]
        batch, lw = make_workload(args.workload, clip_index=0, init_shard=rank, mano_asset=asset)
    eng = FitEngine(batch, lw, lr=1e-2, mano_asset=asset, use_graph=True)

    def final_reduce():
        """The only collective of the job: per-clip argmin over the inits (local), one all_gather of the winners."""
        bi, bl = eng.best_init(clips=n_local)
        payload = None
        if multi:   # the winner's fitted parameters travel with it
            inits = eng.P // n_local
            sel = (torch.arange(n_local, device=bi.device) * inits + bi.long())
            payload = torch.cat([eng.params[k].view(eng.P, -1)[sel] for k in eng.params], 1)
        return hd.gather_best(clip_ids, bi.long(), bl, n_local * world, payload)
    host = eng.stage_host(batch, pin=True)
    eng.capture()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K graph replays
    from homan_b200 import distributed as hd
    for _ in range(args.warmup):
        eng.step()
    final_reduce()  # NCCL warm-up
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.step()
    final_reduce()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * args.steps / (ms_total / 1e3)

    # ---- end to end: host buffers -> upload -> K iterations each followed by a D2H read of the losses
    #      -> fitted parameters back on the host (what optimize_hand_object() does with host inputs)
    loss_host = torch.empty(eng.P, NPART).pin_memory()
    params_host = torch.empty(eng.n_params).pin_memory()
    barrier()
    t0 = time.perf_counter()
    h2d = eng.upload(host)
    for _ in range(args.steps):
        eng.step()
        loss_host.copy_(eng.losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    params_host.copy_(eng.flat, non_blocking=True)
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps / float(t_e2e.item())
    d2h = loss_host.numel() * 4 + params_host.numel() * 4 / args.steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- per-kernel times and the roofline of the raster backward (north-star kernel)
    bd = kernel_breakdown(eng)
    bwd = bd.get("hm_raster_sil_bwd")
    hbm, hbm_src = peaks()
    alg = (raster_bwd_algorithmic_bytes(eng.B, eng.faces_hand.shape[1]) +
           raster_bwd_algorithmic_bytes(eng.B, eng.faces_obj.shape[1])) / 2 if eng.on_sil_hand else \
        raster_bwd_algorithmic_bytes(eng.B, eng.faces_obj.shape[1])
    roofline = None
    traffic = None   # not measurable inside this run (needs ncu): see profiles/ for the ncu capture of this kernel
    if bwd:
        achieved = alg / (bwd["us_per_launch"] * 1e-6) / 1e9
        roofline = {"kernel": "raster_bwd_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                    "frac": achieved / hbm, "traffic": traffic, "peak_source": hbm_src,
                    "algorithmic_bytes_per_launch": alg, "us_per_launch": bwd["us_per_launch"]}
    # ---- CPU baseline (oracle port), bounded sample: one init, one timed iteration
    cpu = None
    if not args.no_cpu_baseline:
        try:
            from homan_b200.workload import loss_weights
            one = {k: (v[:1] if isinstance(v, np.ndarray) and v.shape[:1] == (eng.P,) and k not in
                       ("obj_verts_can", "obj_faces", "hand_faces") else v) for k, v in batch.items()}
            sec = cpu_reference_iteration_rate(args.workload, 1, 1, one, lw, asset)
            cpu = {"value": 1.0 / (sec * eng.P), "unit": "iters/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"1 timed iteration (after 1 warm-up) of 1 of the {eng.P} inits ({eng.T} frames); "
                             f"time x {eng.P} = one whole-batch iteration; oracle/homan_ref.py"}
        except Exception as exc:  # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": "iters/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}
    line = {
        "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "inits": eng.P, "frames": eng.T, "images_per_step_per_gpu": eng.B,
                   "object": cfg["obj"], "hand_mesh": args.hand_mesh,
                   "faces": [int(eng.faces_hand.shape[1]), int(eng.faces_obj.shape[1])],
                   "losses": cfg["lw"], "render": "256^2 (512^2 raster, AA)", "cuda_graph": True,
                   "sharding": ("clips: %d whole clips (own object each) x %d inits per rank, per-clip argmin local, one "
                                "all_gather of the winners and their parameters at the end" % (n_local, eng.P // n_local)) if multi else
                               "inits: every rank fits the clip from its own block of P random inits, one all_gather of the best init at the end",
                   "clips_per_gpu": n_local,
                   "l2": "inputs larger than L2 (face_index maps alone are %d MB per step)" %
                         (eng.B * 2 * 512 * 512 * 4 // 2 ** 20),
                   "problem_frame_iters_per_s": value * eng.B},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "iters/s", "h2d_bytes_per_step": h2d / args.steps,
                "d2h_bytes_per_step": d2h,
                "what": "pinned host batch -> upload -> K x (graph replay + D2H of the loss table) -> parameters D2H"},
        "gpu_launches": eng.gpu_launches_per_step * args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "breakdown_us": bd,
    }
    _LINE.append(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hand-mesh", default="delaunay", choices=["delaunay", "polar"],
                    help="triangulation of the synthetic hand: well-shaped faces (default) or the round-1 polar slivers")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # The contract is ONE JSON line on stdout: whatever libraries write to file descriptor 1 while the job runs (NCCL
    # prints its version banner there when NCCL_DEBUG is set) goes to stderr instead; the line itself is printed last.
    sys.stdout.flush()
    real_stdout, py_stdout = os.dup(1), sys.stdout
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        sys.stderr.flush()
        sys.stdout = py_stdout
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _LINE:
        print(_LINE[0], flush=True)


if __name__ == "__main__":
    main()

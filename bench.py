#!/usr/bin/env python
"""Benchmark of the MEH alpha -> uncertainty -> HUA pool-scoring path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A step = one pass of the hot path over one batch of `--batch` synthetic images whose head outputs
are already resident in HBM (a ring of `--ring` distinct images per GPU, far larger than L2; image
g uses ring slot g mod ring but its own Philox key g).  `value` = pool images scored per second,
whole job (all ranks), device-timed with CUDA events, max over ranks; the pool top-k (K4) and, for
N > 1, the single NCCL all-gather of scores run once inside the timed region.  `e2e` is the same
metric through the host-buffer C-ABI call (pinned host inputs, H2D + D2H inside the timed region).
`roofline` is for the dominant HBM-streaming kernel K1a (logits -> ranking keys), its duration
measured live with CUDA events on the launching stream during the timed steps.
`--impl reference` times the CPU oracle port of the reference's own path on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "pool images scored/sec (MEH alpha->uncertainty->HUA)"
UNIT = "images/s"
DEFAULT_WORKLOAD = "cfg3_retina_r50_800x1344_coco"
POOL_SIZES = {"cfg1_retina_r50_512_voc": 16, "cfg2_ssd300_voc": 5000, "cfg3_retina_r50_800x1344_coco": 100000,
              "cfg3p_retina_r50_800x800_coco": 100000, "cfg4_ssd512_coco": 20000,
              "cfg5_retina_r101_1344_coco": 1000000}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--batch", type=int, default=0, help="images per step per GPU (0 = auto)")
    ap.add_argument("--ring", type=int, default=0, help="distinct resident images per GPU (0 = auto)")
    ap.add_argument("--samples", type=int, default=500, help="Dirichlet samples T (reference: 500)")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--e2e-batch", type=int, default=32, help="images per host-buffer call")
    ap.add_argument("--cpu-images", type=int, default=4, help="images of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager-gpu-images", type=int, default=0,
                    help="also time the oracle port in torch-eager on the GPU over this many images (0 = off)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference's own path; the reference itself is pure
# Python on torch and cannot be installed without mmcv - see DESIGN.md)
# --------------------------------------------------------------------------------------------
def cpu_reference_rate(spec, params, n_images: int, batch: int, threads: int, seed0: int = 20):
    from aod_meh_hua_b200.synth import SyntheticPool
    from oracle import meh_hua_oracle as O
    torch.set_num_threads(threads)
    pool = SyntheticPool(spec, seed0=seed0, device="cpu")
    kw = O.spec_kwargs(spec, params)
    batches = [pool.batch(list(range(i, min(i + batch, n_images)))) for i in range(0, n_images, batch)]
    torch.manual_seed(20)
    O.score_batch(pool.batch([n_images]), **kw)          # warm-up batch (not timed)
    t0 = time.perf_counter()
    for b in batches:
        O.score_batch(b, **kw)
    dt = time.perf_counter() - t0
    return n_images / dt, dt


def eager_gpu_rate(spec, params, n_images: int, batch: int, device, seed0: int = 20):
    """the same oracle port run in torch-eager ON the B200 (stock ATen / torchvision
    kernels) - the 'existing kernels on Blackwell' bar.  Part of the cpu_baseline leg; off by default."""
    from aod_meh_hua_b200.synth import SyntheticPool
    from oracle import meh_hua_oracle as O
    pool = SyntheticPool(spec, seed0=seed0, device="cpu")
    kw = O.spec_kwargs(spec, params)

    def to_dev(bt):
        return {k: ([t.to(device) for t in v] if isinstance(v, list) and v and torch.is_tensor(v[0]) else v)
                for k, v in bt.items()}
    batches = [to_dev(pool.batch(list(range(i, min(i + batch, n_images))))) for i in range(0, n_images, batch)]
    torch.manual_seed(20)
    O.score_batch(to_dev(pool.batch([n_images, n_images + 1])), **kw)      # warm-up (not timed)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for bt in batches:
        O.score_batch(bt, **kw)
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    return n_images / dt, dt


def run_reference(args, spec, params, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = 2
    rates = []
    from aod_meh_hua_b200.synth import SyntheticPool
    from oracle import meh_hua_oracle as O
    torch.set_num_threads(threads)
    pool = SyntheticPool(spec, seed0=20, device="cpu")
    kw = O.spec_kwargs(spec, params)
    torch.manual_seed(20)
    steps, warm = max(1, min(args.steps, 8)), max(1, min(args.warmup, 1))
    data = [pool.batch([2 * i, 2 * i + 1]) for i in range(2)]
    for i in range(warm):
        O.score_batch(data[i % 2], **kw)
    t0 = time.perf_counter()
    for i in range(steps):
        O.score_batch(data[i % 2], **kw)
    dt = time.perf_counter() - t0
    value = steps * per_step / dt
    sample = f"{steps} steps x {per_step} images of {spec.name}, batch 2, T={params.n_samples}, torch CPU"
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warm,
                ms_per_step=1e3 * dt / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(workload=spec.name, samples=params.n_samples, batch=per_step),
                cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def build_ring(spec, ring: int, gid0: int, device, seed0: int = 20):
    """Head outputs of `ring` distinct images, resident on the device: per level [ring, ch, H, W]."""
    from aod_meh_hua_b200.synth import SyntheticPool
    pool = SyntheticPool(spec, seed0=seed0, device=device)
    S = spec.num_levels
    cls = [torch.empty(ring, a * spec.c_out, h, w, device=device) for (h, w), a in zip(spec.featmaps, spec.num_anchors)]
    reg = [torch.empty(ring, a * 4, h, w, device=device) for (h, w), a in zip(spec.featmaps, spec.num_anchors)]
    lam = [torch.empty(ring, a, h, w, device=device) for (h, w), a in zip(spec.featmaps, spec.num_anchors)]
    for r in range(ring):
        im = pool.image(gid0 + r)
        for s in range(S):
            cls[s][r].copy_(im["cls"][s]); reg[s][r].copy_(im["reg"][s]); lam[s][r].copy_(im["lam"][s])
    return pool, cls, reg, lam


def main():
    args = parse_args()
    from aod_meh_hua_b200.specs import ScoringParams, get_spec
    spec = get_spec(args.workload)
    params = ScoringParams(n_samples=args.samples)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, spec, params, rank)
        return

    from aod_meh_hua_b200 import _lib
    from aod_meh_hua_b200.pool import gather_scores, shard_range
    from aod_meh_hua_b200.scoring import Scorer, pool_topk
    if not torch.cuda.is_available():
        raise _lib.MehhuaError("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    lib = _lib.load()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    img_bytes = 4 * spec.num_priors * (spec.c_out + 5)
    # images per step: at most 3 x 148 (three waves of the one-block-per-image kernels on the B200's 148 SMs)
    # and at most 30 GB of head outputs per step (cfg 3: 437 images); ring = 2..4 steps of distinct images
    B = args.batch or max(1, min(444, int(30e9 // img_bytes)))
    ring = args.ring or max(2 * B, ((int(60e9 // img_bytes)) // B) * B)
    ring = max(B, (min(ring, 4 * B) // B) * B)
    pool_size = POOL_SIZES.get(spec.name, 100000)
    lo, hi = shard_range(pool_size, rank, world)
    synth, cls, reg, lam = build_ring(spec, ring, lo, device)
    sc = Scorer(spec, params, max_batch=B, device=device)
    shp = torch.tensor([[spec.img_hw[0], spec.img_hw[1]]] * B, dtype=torch.float32, device=device)
    sf = torch.ones(B, 4, device=device)
    pool_scores = torch.zeros(max(hi - lo, B), device=device)
    n_slots = ring // B
    slot_ptrs = []
    for j in range(n_slots):
        slot_ptrs.append([(cls[s][j * B:].data_ptr(), reg[s][j * B:].data_ptr(), lam[s][j * B:].data_ptr(),
                           synth.anchors[s].data_ptr()) for s in range(spec.num_levels)])
    n_local = hi - lo
    ids_all = torch.arange(lo, lo + (args.steps + args.warmup + 2) * B, device=device, dtype=torch.int64)

    def step(i: int):
        j = i % n_slots
        ids = ids_all[i * B:(i + 1) * B]
        sc.bind_raw(slot_ptrs[j], B, shp, sf, ids)
        sc.score_bound()
        dst = (i * B) % max(n_local - B + 1, 1)
        pool_scores[dst:dst + B].copy_(sc.t["image_scores"][:B], non_blocking=True)

    def finish():
        allsc = gather_scores(pool_scores[:n_local], pool_size, rank, world) if world > 1 else pool_scores
        k = max(1, min(int(0.025 * pool_size), allsc.numel()))
        return pool_topk(allsc, k)

    for i in range(args.warmup):
        step(i)
    finish()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    _lib.check(lib.mehhua_stage_timing_begin(args.steps), "stage_timing_begin")
    launches0 = lib.mehhua_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    selected = finish()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = int(lib.mehhua_launch_count() - launches0)
    stage_ms = (C.c_double * 7)()
    calls = C.c_int32(0)
    _lib.check(lib.mehhua_stage_timing_end(stage_ms, C.byref(calls)), "stage_timing_end")
    clocks = sampler.stop()
    status = sc.check_status()
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps * B / (ms * 1e-3)

    # ---- roofline of the dominant kernel (K1a), timed live above
    peak, peak_src = measured_peaks()
    names = ["k1a_keys", "k1b_select", "k1c_gather", "k3a_nms", "k3b_pairs", "k2_dirichlet", "k3c_hua"]
    stage_avg = {n: stage_ms[i] / max(calls.value, 1) for i, n in enumerate(names)}
    k1a_bytes = B * (4 * spec.num_priors * spec.c_out + 4 * spec.num_priors)       # logits read + keys written
    k1a_s = stage_avg["k1a_keys"] * 1e-3
    achieved = k1a_bytes / k1a_s / 1e9 if k1a_s > 0 else 0.0
    k1_all_s = (stage_avg["k1a_keys"] + stage_avg["k1b_select"] + stage_avg["k1c_gather"]) * 1e-3
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1a_traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(spec.name, {}).get(str(B))
    roofline = dict(bound="hbm", kernel="k1a_keys_kernel", achieved=achieved, peak=peak, unit="GB/s",
                    frac=achieved / peak, traffic=traffic, peak_source=peak_src,
                    algorithmic_bytes_per_launch=k1a_bytes,
                    k1_stage_achieved=B * spec.k1_bytes_per_image() / k1_all_s / 1e9 if k1_all_s > 0 else 0.0,
                    stage_ms_per_step=stage_avg)

    # ---- aggregation-stage kernels against the same HBM peak (tiny algorithmic traffic,
    #      expected latency-bound - reported, not a target)
    S_ = spec.num_levels
    p_tot = float(sc.t["pair_off"][:B, S_].float().sum().item())
    n_obj_tot = float(sc.t["n_obj"][:B].float().sum().item())
    k3_bytes = 16.0 * (B * spec.k_tot + n_obj_tot) + 12.0 * p_tot + 4.0 * B
    k3_s = (stage_avg["k3b_pairs"] + stage_avg["k3c_hua"]) * 1e-3
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pool_all = torch.rand(pool_size, device=device)
    kk = max(1, int(0.025 * pool_size))
    pool_topk(pool_all, kk)
    t0e.record()
    for _ in range(5):
        pool_topk(pool_all, kk)
    t1e.record()
    torch.cuda.synchronize()
    k4_s = t0e.elapsed_time(t1e) / 5 * 1e-3
    k4_bytes = 4.0 * pool_size + 8.0 * kk
    roofline["aggregation"] = dict(
        k3_pairs_hua=dict(bytes_per_step=k3_bytes, ms=k3_s * 1e3, achieved=k3_bytes / k3_s / 1e9 if k3_s > 0 else 0.0,
                          frac=(k3_bytes / k3_s / 1e9 / peak) if k3_s > 0 else 0.0, pairs_per_image=p_tot / B,
                          note="latency-bound: one block per image"),
        k4_pool_topk=dict(bytes_per_launch=k4_bytes, ms=k4_s * 1e3, achieved=k4_bytes / k4_s / 1e9,
                          frac=k4_bytes / k4_s / 1e9 / peak, pool=pool_size, k=kk, note="single block, once per pool"),
        k2_draws_per_s=(p_tot * params.n_samples * spec.c_out) / (stage_avg["k2_dirichlet"] * 1e-3)
        if stage_avg["k2_dirichlet"] > 0 else 0.0)

    # ---- e2e: host buffers through mehhua_score_batch_host (pinned inputs, H2D + D2H timed)
    e2e = None
    if rank == 0 or world > 1:
        Be = min(B, args.e2e_batch)
        ctx = C.c_void_p()
        _lib.check(lib.mehhua_host_ctx_create(C.byref(sc.cfg), sc._shape_levels, Be, C.byref(ctx)), "host_ctx_create")
        h_cls = [c[:Be].cpu().pin_memory() for c in cls]
        h_reg = [r_[:Be].cpu().pin_memory() for r_ in reg]
        h_lam = [l_[:Be].cpu().pin_memory() for l_ in lam]
        h_anc = [a.cpu().contiguous() for a in synth.anchors]
        lv = _lib.LevelArray()
        for s in range(spec.num_levels):
            lv[s].logits, lv[s].deltas, lv[s].lam = h_cls[s].data_ptr(), h_reg[s].data_ptr(), h_lam[s].data_ptr()
            lv[s].anchors = h_anc[s].data_ptr()
            (lv[s].H, lv[s].W), lv[s].A = spec.featmaps[s], spec.num_anchors[s]
        h_shp = np.asarray([[spec.img_hw[0], spec.img_hw[1]]] * Be, dtype=np.float32)
        h_sf = np.ones((Be, 4), dtype=np.float32)
        h_ids = np.arange(lo, lo + Be, dtype=np.int64)
        h_out = np.zeros(Be, dtype=np.float32)
        st = C.c_uint32(0)

        def e2e_step():
            _lib.check(lib.mehhua_score_batch_host(ctx, lv, Be, h_shp.ctypes.data, h_sf.ctypes.data, h_ids.ctypes.data,
                                                   h_out.ctypes.data, C.byref(st)), "score_batch_host")
        e2e_step()
        for s in range(spec.num_levels):
            lv[s].anchors = None          # anchors stay resident after the first call
        e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        lib.mehhua_host_ctx_destroy(ctx)
        e2e = dict(value=world * args.e2e_steps * Be / dt, unit=UNIT, h2d_bytes_per_step=Be * img_bytes + Be * 32,
                   d2h_bytes_per_step=Be * 4 + 4, batch=Be, steps=args.e2e_steps,
                   api="mehhua_score_batch_host (pinned host inputs)")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rate, dt = cpu_reference_rate(spec, params, args.cpu_images, 2, cores)
        cpu_baseline = dict(value=rate, unit=UNIT, cores=cores, kind="port",
                            sample=f"{args.cpu_images} images of {spec.name}, batch 2, T={params.n_samples}, "
                                   f"oracle port on torch CPU, {dt:.1f} s")
        if args.eager_gpu_images > 0:
            try:
                r2, dt2 = eager_gpu_rate(spec, params, args.eager_gpu_images, 2, device)
                cpu_baseline["torch_eager_b200"] = dict(
                    value=r2, unit=UNIT, sample=f"{args.eager_gpu_images} images, batch 2, T={params.n_samples}, oracle "
                                                f"port in torch-eager on cuda (ATen / torchvision kernels), {dt2:.1f} s")
            except Exception as e:      # a reported side baseline: never fails the bench line
                cpu_baseline["torch_eager_b200"] = dict(error=f"{type(e).__name__}: {e}"[:200])
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic",
                config=dict(workload=spec.name, batch_per_gpu=B, ring_images_per_gpu=ring, samples=params.n_samples,
                            pool_size=pool_size, l2_policy=f"inputs larger than L2: {B * img_bytes / 1e6:.0f} MB per step, "
                            f"{ring * img_bytes / 1e9:.2f} GB ring", parallelism=f"pool sharded by image x{world}",
                            selected=int(selected.numel()), status_bits=status),
                roofline=roofline, cpu_baseline=cpu_baseline, e2e=e2e, gpu_launches=launches, clocks=clocks)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

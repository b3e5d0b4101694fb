#!/usr/bin/env python
"""Bench of the NopeSAC one-plane RANSAC pose path on B200 (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" = one pass of the path over one batch of synthetic pairs, from RGB (stage set S5 of SURVEY.md §8d, BASELINE.json
configs[1]): uint8 480x640 images of both views -> ResNet-50 backbone (random init) -> PlaneCameraHead.inference_Joint incl.
the matching head (pixel pose network -> AIM -> GNN + Sinkhorn matcher -> geo sequences -> hypothesis generation -> scoring ->
soft aggregation -> assignment pruning).  Per GPU: 64 pairs, 16 planes/view, NUM_OBJECT_QUERIES = 256 with all 16x16
candidate plane pairs as one-plane hypotheses (257 hypotheses x 256 residual columns per pair).  The plane lists (what
PlaneTRHead + post-processing produce) are synthetic inputs.  N > 1 (torchrun): pairs are sharded 64/GPU (weak scaling), the
[64,16] result rows are exchanged by the fused NVLink store of the selection kernel (or one NCCL all-gather) — the only
exchange of the path — and the gathered rows are verified after the timed loop (`exchange_verified`).

value    pairs/s, inputs resident in HBM; an L2 flush (256 MB write) separates the timed steps.
e2e      pairs/s through the same public call with HOST (pinned) inputs: H2D of the step's uint8 images + plane lists and D2H of
         the [B,16] result rows inside the timed region.
roofline the hypothesis-scoring call (nsac_score_aggregate_tc) timed alone with CUDA events on a launch that moves > 256 MB
         (B=512, m=NQ=256), algorithmic bytes of SURVEY.md §8d over the measured HBM copy bandwidth.  `score_sweep` repeats it
         for BASELINE.json configs[4] (32/128/512/2048 hypotheses) with both bounds; `roofline_tensor` times the GEMM engine
         (most of the step) on its three shape classes against the measured bf16 peak.
cpu_baseline / --impl reference: the CPU oracle port (oracle/backbone_restate.py + oracle/restate.py — the reference is Python
         and cannot travel to the GPU box) on the host cores, per-pair loop at batch size 1 like the reference, same stage set.
"""
from __future__ import annotations

import argparse
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

import torch  # noqa: E402

PAIRS_PER_GPU = 64
PLANES = 16
NQ = 256
IMG_H, IMG_W = 480, 640
WORKLOAD = ("full model from RGB, stage set S5: ResNet-50 backbone (random init, FrozenBN) on both uint8 480x640 views + camera "
            "head (pixel pose net + AIM + GNN/Sinkhorn matcher + 256 one-plane hypotheses + scoring + soft aggregation + "
            "pruning), 64 synthetic pairs/GPU, 16 planes/view, NUM_OBJECT_QUERIES=256 (all 16x16 plane pairs as hypotheses), "
            "fp32-grade arithmetic (3-pass fp16 hi/lo tensor-core GEMMs), random-init weights, synthetic plane lists")
METRIC = "image-pairs/sec (480x640, 16 planes x 256 hyp)"


def score_algorithmic_bytes(B, m, nq):
    """per pair 24m + (28 + 8C)(m+1) + 68, C=256, + score-MLP weights once per launch."""
    per_pair = 24 * m + 2076 * (m + 1) + 68
    weights = 8 * (nq * 128 + 128 + 128 * 128 + 128 + 128 * 64 + 64 + 64 + 1)
    return B * per_pair + weights


def score_algorithmic_flops(B, m, nq):
    """SURVEY.md §8(d) "Algorithmic FLOPs": residuals ~2*45*(m+1)*m, both score MLPs 2*2*(m+1)*(nq*128 + 128^2 + 128*64 + 64),
    aggregation 2*2*C*(m+1)."""
    return B * (2 * 45 * (m + 1) * m + 4 * (m + 1) * (nq * 128 + 128 * 128 + 128 * 64 + 64) + 4 * 256 * (m + 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.proc, self.first = gpu_index, [], None, 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Start of the timed region: earlier samples (warm-up) are dropped."""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        end = len(self.rows)              # rows that arrived while the timed region ran (+ one if none did: its last 100 ms)
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows[self.first:max(end, self.first + 1)]:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm": float(d["hbm_gbs"]), "tf_burst": float(d["bf16_tflops"]), "tf_sustained": float(d["bf16_tflops_sustained"]),
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


# ---------------------------------------------------------------------------------------------------
def backbone_state():
    """Seeded random-init R-50 state shared by both arms (shapes from the product's own module)."""
    from nopesac_b200 import backbone, synthetic
    shapes = {k: tuple(v.shape) for k, v in backbone.ResNet50Backbone().state_dict().items()}
    return synthetic.make_backbone_weights(shapes, seed=8)


def cpu_reference_pairs_per_s(num_pairs: int, threads: int, warm: int = 1):
    """The oracle port on the host, same stage set S5: per-pair loop at bs=1 (the only mode the reference supports) — R-50 on
    both views (oracle/backbone_restate.py), then the camera head (oracle/restate.py) — same workload shape, same weights."""
    from oracle import backbone_restate as br
    from oracle import restate
    from nopesac_b200 import config, synthetic
    from tests import util
    torch.set_num_threads(threads)
    cfg = config.inference_cfg(NQ)
    sd, msd = util.make_weights(NQ)
    bsd = backbone_state()
    hp = synthetic.all_pairs_hypotheses(PLANES, NQ)
    batches = [synthetic.make_batch(1000 + i, 1, PLANES) for i in range(num_pairs + warm)]
    images = [synthetic.make_images(5000 + i, 2, IMG_H, IMG_W) for i in range(num_pairs + warm)]

    def one(b, im):
        with torch.no_grad():
            f = br.resnet50(bsd, br.normalize(im.float(), cfg.MODEL.PIXEL_MEAN, cfg.MODEL.PIXEL_STD))
            f1 = {k: v[0:1] for k, v in f.items()}
            f2 = {k: v[1:2] for k, v in f.items()}
            return restate.inference_joint(sd, msd, f1, f2, b.planes1, b.planes2, b.app1, b.app2, num_queries=NQ, hyp_pairs=hp)
    for b, im in zip(batches[:warm], images[:warm]):
        one(b, im)
    t0 = time.perf_counter()
    for b, im in zip(batches[warm:], images[warm:]):
        one(b, im)
    dt = time.perf_counter() - t0
    return num_pairs / dt, dt


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 2
    vals = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_pairs_per_s(1, threads, warm=1)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, _ = cpu_reference_pairs_per_s(sample, threads, warm=0)
        vals.append(v)
    wall = time.perf_counter() - t_all
    value = statistics.mean(vals)
    desc = f"{sample} pairs/step x {args.steps} steps of the same workload (S5: R-50 on both views + camera head), per-pair loop at batch size 1"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": PAIRS_PER_GPU, "planes_per_view": PLANES, "num_object_queries": NQ,
                   "stage_set": "S5", "pairs_per_step": sample, "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def measure_plane_lists(dev, images: int, iters: int = 10):
    """nsac_plane_postprocess on `images` synthetic PlaneTRHead outputs of the inference_mp3d shape (50 queries, 120x160 mask
    logits -> 480x640): ms per call with CUDA events on the launch stream, inputs (> 126 MB for 64+ images) exceed L2."""
    from nopesac_b200 import plane_postprocess, synthetic
    base = synthetic.make_plane_head_batch(300, 8, cases=("regular",))
    rep = (images + 7) // 8
    batch = {k: v.repeat(rep, *([1] * (v.dim() - 1)))[:images].contiguous().to(dev) for k, v in base.items()}
    outs = {k: batch[k] for k in ("pred_logits", "pred_params", "pred_mask_logits")}
    for _ in range(3):
        res = plane_postprocess.postprocess_plane_head_mask(outs, batch["query_feat"], 480, 640)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        res = plane_postprocess.postprocess_plane_head_mask(outs, batch["query_feat"], 480, 640)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nq, h, w = batch["pred_mask_logits"].shape[1:]
    alg = images * (4 * nq * h * w + 3 * 480 * 640)
    return {"images": images, "ms_per_call": ms, "images_per_s": images / ms * 1e3, "planes_per_image": float(res.count.float().mean()),
            "algorithmic_bytes": alg, "achieved_gbs": alg / ms / 1e6,
            "note": "row f1, both views of this rank's pairs; outside the timed region of `value`"}


def measure_full_model(dev, pairs: int, iters: int = 3):
    """BASELINE.json configs[3] shape on one GPU: the COMPLETE META_ARCH from uint8 RGB (configs/inference_mp3d.yaml:
    NUM_OBJECT_QUERIES 50, data-dependent plane counts) — backbone -> PlaneTRHead -> plane lists -> matcher + camera head in one
    `inference_from_rgb` call, random-init weights; plus PlaneTRHead alone on the same 2 * pairs images (from the backbone's planes)."""
    from nopesac_b200 import config, meta_arch, synthetic
    from tests import util
    nq = 50
    model = meta_arch.PlaneTR_NopeSAC(config.inference_cfg(nq), with_backbone=True, with_plane_head=True)
    sd, msd = util.make_weights(nq)
    model.camera_head_list[0].load_state_dict(sd)
    model.matching_head.load_state_dict(msd)
    model.sem_seg_head.load_state_dict(synthetic.make_weights(util.planetr_shapes(nq), 77))
    model.backbone.load_state_dict(backbone_state())
    model = model.to(dev)
    images = synthetic.make_images(9100, 2 * pairs, IMG_H, IMG_W).to(dev)
    batched = [{"0": {"image": images[i], "height": IMG_H, "width": IMG_W},
                "1": {"image": images[pairs + i], "height": IMG_H, "width": IMG_W}} for i in range(pairs)]

    def timed(fn, n):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, r
    ms_full, res = timed(lambda: model.inference_from_rgb(batched, max_planes=20), iters)
    feats = model.backbone(images, planes=True)
    ms_head, _ = timed(lambda: model.sem_seg_head(feats), iters)
    counts = torch.cat([res[1].count, res[2].count]).float()
    return {"config": "BASELINE configs[3] shape (inference_mp3d.yaml, NUM_OBJECT_QUERIES=50) on 1 GPU", "pairs": pairs,
            "ms_per_call": ms_full, "pairs_per_s": pairs / ms_full * 1e3, "planes_per_image": float(counts.mean()),
            "plane_head": {"images": 2 * pairs, "ms_per_call": ms_head, "images_per_s": 2 * pairs / ms_head * 1e3},
            "note": "uint8 RGB -> backbone -> PlaneTRHead -> plane lists -> matcher + camera head, one call, inputs resident; "
                    "outside the timed region of `value`"}


def graph_timed(fn, reps: int, flush=None):
    """ms per call of `fn` (a short chain of kernel launches): captured once in a CUDA graph, replays timed with CUDA events on the
    replay stream (launched eagerly from Python the host side is slower than such kernels).  `flush`: tensor zeroed before
    every replay (inputs that fit in L2), outside the timed events."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    graph.replay()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return statistics.mean(times)


def score_case(dev, B, nq, head=None):
    """Inputs + closure of one nsac_score_aggregate_tc call (K8+K9) with m = nq valid hypotheses per pair."""
    from nopesac_b200 import ops
    from tests import util
    if head is None:
        head, _, _, _ = util.build_cuda_heads(nq, "soft", 0.2, dev)
    g = torch.Generator(device=dev).manual_seed(3)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
    geo_local = rnd(B, nq, 6)
    q_h = torch.nn.functional.normalize(rnd(B, nq, 4), dim=-1)
    t_h = rnd(B, nq, 3) * 0.3
    q0 = torch.nn.functional.normalize(rnd(B, 4), dim=-1)
    t0 = rnd(B, 3) * 0.3
    fr, ft = rnd(B, nq, 256), rnd(B, nq, 256)
    fr0, ft0 = rnd(B, 256), rnd(B, 256)
    mnum = torch.full((B,), nq, device=dev, dtype=torch.int32)
    pk = head.prepare_tc()

    def once():
        return ops.score_aggregate(geo_local, q_h, t_h, q0, t0, fr, ft, fr0, ft0, mnum, pk["normal_score_proj"],
                                   pk["param_score_proj"], head.rots.weight, head.rots.bias, head.trans.weight,
                                   head.trans.bias, out_cam_type="soft", want_scores=False, pack=pk["score_pack"],
                                   vecs_host=pk["score_vecs_host"])
    return once


def measure_gemm_engine(dev, peaks):
    """The GEMM engine (nsac_gemm_split / nsac_conv3x3_split, 3 MMA passes on fp16 hi/lo planes) on its three shape classes,
    CUDA events around back-to-back launches on the launch stream; tensor-pipe work = 3 x 2MNK."""
    from nopesac_b200 import ops
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
    out = []

    def timed(fn, flops, name, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = 3 * flops / ms / 1e9
        out.append({"shape": name, "us": ms * 1e3, "tflops_tensor": tf, "tflops_fp32_equivalent": tf / 3,
                    "frac_of_bf16_sustained": tf / peaks["tf_sustained"]})

    M, N, K = PAIRS_PER_GPU * NQ, 1024, 1024
    a, w = ops.split(rnd(M, K)), ops.split(rnd(N, K) * 0.03)
    o = ops.Split.empty(M, N, dev)
    timed(lambda: ops.gemm_tc(a, w, None, ops.ACT_RELU, want_f32=False, out_split=o), 2.0 * M * N * K,
          f"K7 hypothesis-generation layer {M}x{N}x{K}")
    Ni, H, W, C = 2 * PAIRS_PER_GPU, 60, 80, 256
    x, wc = ops.split(rnd(Ni * H * W, C)), ops.split(rnd(C, 9 * C) * 0.02)
    timed(lambda: ops.conv3x3_tc(x, Ni, H, W, wc, None, ops.ACT_LEAKY, want_f32=False, want_split=True),
          2.0 * Ni * H * W * C * 9 * C, f"K1 3x3 convolution {Ni}x{H}x{W} {C}->{C} (implicit GEMM)")
    M, N, K = 2 * PAIRS_PER_GPU * PLANES, 256, 256
    a2, w2 = ops.split(rnd(M, K)), ops.split(rnd(N, K) * 0.06)
    timed(lambda: ops.gemm_tc(a2, w2), 2.0 * M * N * K, f"K4 GNN linear {M}x{N}x{K} (launch-bound)", reps=50)
    best = max(out, key=lambda r: r["tflops_tensor"])
    return {"kernel": "gemm_bf16x3_kernel (tcgen05 / TMA / TMEM; 3 MMA passes hi.hi + lo.hi + hi.lo on fp16 planes ~ fp32)",
            "bound": "tensor", "achieved": best["tflops_tensor"], "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
            "frac": best["tflops_tensor"] / peaks["tf_sustained"], "peak_source": peaks["src"] + " bf16_tflops_sustained",
            "precision": "3-pass fp16 (tensor work = 3 x 2MNK; fp32-equivalent throughput = achieved / 3)", "traffic": None,
            "shapes": out, "timing": "CUDA events around back-to-back launches on the launch stream"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-value", action="store_true", help="profiling runs: skip the e2e / roofline / cpu legs")
    ap.add_argument("--cpu-pairs", type=int, default=4, help="pairs timed by the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    from nopesac_b200 import config, meta_arch, ops, synthetic
    from tests import util   # seeded weights (shapes from tests/golden/state_shapes.json)

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.pairs

    # the reference's META_ARCH with its backbone (configs/Base.yaml:2-12) + camera head + matching head, seeded random weights
    cfg = config.inference_cfg(NQ, "soft", 0.2)
    model = meta_arch.PlaneTR_NopeSAC(cfg, with_backbone=True)
    sd, msd = util.make_weights(NQ)
    model.camera_head_list[0].load_state_dict(sd)
    model.matching_head.load_state_dict(msd)
    model.backbone.load_state_dict(backbone_state())
    model = model.to(dev)
    head, match = model.camera_head_list[0], model.matching_head

    hp = synthetic.all_pairs_hypotheses(PLANES, NQ).to(dev, torch.int32)
    host = synthetic.make_batch(rank * B, B, PLANES)                          # planes / appearance: seeded on the host
    host_images = synthetic.make_images(7000 + rank, 2 * B, IMG_H, IMG_W)     # uint8 [2B,3,480,640]: first views, then second views
    dbatch = host.to(dev)
    images = host_images.to(dev)
    gathered = torch.empty(world * B, 16, device=dev) if world > 1 else None
    exchange, exchange_kind = None, "single GPU: no exchange"
    if world > 1:
        exchange_kind = "one NCCL all-gather of [B,16] result rows per step"
        if os.environ.get("NSAC_EXCHANGE", "fused") == "fused":
            try:
                from nopesac_b200.dist import FusedResultExchange
                exchange = FusedResultExchange(B, dev)
                exchange_kind = ("fused: the selection kernel stores each result row into every rank's buffer over NVLink "
                                 "peer memory (symmetric memory, double-buffered by step parity) + one cross-rank barrier per "
                                 "step; no collective")
            except Exception as e:  # noqa: BLE001  (no P2P / symmetric memory on this box)
                exchange = None
                exchange_kind += f" (fused exchange unavailable: {type(e).__name__})"

    def forward(img, p1, p2, a1, a2, ex=None):
        return model.inference_from_images(img, None, p1, p2, a1, a2, hyp_pairs=hp, result_exchange=ex)

    def step(img, p1, p2, a1, a2):
        pose = forward(img, p1, p2, a1, a2, exchange)[5]["pose"]
        if exchange is not None:
            return exchange.finish()
        if world > 1:
            dist.all_gather_into_tensor(gathered, pose)
            return gathered
        return pose

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ value: inputs resident in HBM
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)        # > 126 MB L2
    # nvidia-smi needs a few hundred ms before its first sample: start it before the warm-up, keep only the rows that arrive
    # between the two barriers of the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step(images, dbatch.planes1, dbatch.planes2, dbatch.app1, dbatch.app2)
    barrier()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    e0.record()
    for _ in range(args.steps):
        flush.zero_()
        step(images, dbatch.planes1, dbatch.planes2, dbatch.app1, dbatch.app2)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * args.steps / (ms_total / 1e3)

    # ------------------------------------------------------------------ multi-GPU: the exchanged rows are the right rows
    # (replaces the reference's pickled comm.gather, mp3d_evaluation.py:317-318; SURVEY.md §4 "all-gather equality vs
    # single-rank concatenation").  One extra step outside the timed region: (1) the rows the fused NVLink exchange left in
    # this rank's buffer == one NCCL all-gather of every rank's local rows, bit for bit, on EVERY rank; (2) the LAST rank
    # recomputes rank 0's shard from rank 0's seeds on its own GPU and compares it with rows [0, B) it received.
    exchange_check = None
    if world > 1:
        out = forward(images, dbatch.planes1, dbatch.planes2, dbatch.app1, dbatch.app2, exchange)
        local_rows = out[5]["pose"].contiguous()
        got = exchange.finish().clone() if exchange is not None else None
        ref = torch.empty(world * B, 16, device=dev)
        dist.all_gather_into_tensor(ref, local_rows)
        if got is None:
            got = ref
        same = bool(torch.equal(got, ref)) and bool(torch.equal(ref[rank * B:(rank + 1) * B], local_rows))
        rec_equal, rec_diff = True, 0.0
        if rank == world - 1:
            h0 = synthetic.make_batch(0, B, PLANES).to(dev)
            im0 = synthetic.make_images(7000, 2 * B, IMG_H, IMG_W).to(dev)
            r0 = forward(im0, h0.planes1, h0.planes2, h0.app1, h0.app2)[5]["pose"]
            rec_equal = bool(torch.equal(r0, got[:B]))
            rec_diff = float((r0 - got[:B]).abs().max())
            del im0
        flags = torch.tensor([1.0 if same else 0.0, 1.0 if rec_equal else 0.0, -rec_diff], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        exchange_check = {"exchange_verified": bool(flags[0].item() == 1.0), "ranks": world,
                          "rows": [world * B, 16], "what": "rows left by the exchange == NCCL all-gather of the local rows (torch.equal, "
                          "every rank) and each rank's own shard sits at its block offset",
                          "recompute_of_rank0_shard_on_last_rank": {"bit_equal": bool(flags[1].item() == 1.0),
                                                                    "max_abs_diff": float(-flags[2].item())}}

    # pose error of this rank's pairs against the planted ground truth (reference formulas, mp3d_evaluation.py:382-425,
    # computed on the device by nopesac_b200.evaluation): reported, not a target - the weights are random
    pose_err = None
    if rank == 0:
        from nopesac_b200 import evaluation
        rows = forward(images, dbatch.planes1, dbatch.planes2, dbatch.app1, dbatch.app2)[5]["pose"]
        pm = evaluation.camera_metrics(rows, dbatch.gt_tran, dbatch.gt_quat)
        pose_err = {"T_median_m": pm["T median err"], "T_mean_m": pm["T mean err"], "R_median_deg": pm["R median err"],
                    "R_mean_deg": pm["R mean err"], "pairs": B, "note": "random-init weights vs planted GT: reported, not a target"}

    if args.only_value:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
                              "ms_per_step": ms_total / args.steps, "gpu_launches": launches, "partial": True}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ e2e: host (pinned) inputs
    # Same public call, inputs in pinned host memory: every step copies its uint8 images (2 x 64 x 3 x 480 x 640 = 118 MB) and
    # plane lists H2D and reads the [B,16] result rows back; nopesac_b200.runtime.PairPipeline overlaps the copy of step i+1
    # with the kernels of step i (two device slots, copy stream + events).
    from nopesac_b200.runtime import PairPipeline, batch_bytes, pin_batch
    hbatch = pin_batch({"images": host_images, "planes1": host.planes1, "planes2": host.planes2, "app1": host.app1, "app2": host.app2})
    h2d = batch_bytes(hbatch)
    d2h = (world * B if world > 1 else B) * 16 * 4
    post = None
    if world > 1:
        def post(rows):
            if exchange is not None:
                return exchange.finish()
            dist.all_gather_into_tensor(gathered, rows)
            return gathered
    pipe = PairPipeline(head, match, dev, post=post,
                        compute=lambda d: forward(d["images"], d["planes1"], d["planes2"], d["app1"], d["app2"], exchange)[5]["pose"])
    e2e_steps = max(3, min(args.steps, 6))
    for _ in pipe.run([hbatch] * 2):
        pass
    barrier()
    e0.record()
    for _ in pipe.run([hbatch] * e2e_steps):
        pass
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / (float(t.item()) / 1e3)
    del pipe

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ S4 for continuity with round 1 (camera head from feature maps)
    s4 = None
    try:
        f1, f2 = synthetic.device_features(B, dev, seed=7)
        run4 = lambda: head(f1, f2, dbatch.planes1, dbatch.planes2, dbatch.app1, dbatch.app2, matching_net=match, hyp_pairs=hp)
        for _ in range(3):
            run4()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            run4()
        e1.record()
        torch.cuda.synchronize()
        ms4 = e0.elapsed_time(e1) / 5
        s4 = {"value": B / ms4 * 1e3, "unit": "pairs/s", "ms_per_step": ms4, "stage_set": "S4: camera head from fp32 NCHW backbone "
              "feature maps (round 1's headline configuration), inputs 2.2 GB/GPU resident in HBM"}
        del f1, f2
    except Exception as e:  # noqa: BLE001
        s4 = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ------------------------------------------------------------------ roofline: scoring kernel alone
    peaks = measured_peaks()
    RB, m = 512, NQ
    score_once = score_case(dev, RB, NQ, head)
    # One scoring call = 3 short kernels (prep, tiles, selection) chained by programmatic dependent launch.
    score_ms = graph_timed(score_once, 20)                 # 276 MB of inputs per launch > L2: no flush needed
    alg = score_algorithmic_bytes(RB, m, NQ)
    achieved = alg / (score_ms / 1e3) / 1e9
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of the same launch from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "score_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roofline = {"kernel": "nsac_score_aggregate_tc (K8+K9: prep + tile + selection kernels), B=512, m=NQ=256", "bound": "hbm",
                "achieved": achieved, "peak": peaks["hbm"], "peak_source": peaks["src"] + " hbm_gbs", "unit": "GB/s",
                "frac": achieved / peaks["hbm"], "traffic": traffic, "algorithmic_bytes_per_launch": alg, "ms_per_launch": score_ms,
                "timing": "CUDA-graph replays of one call, CUDA events on the replay stream"}
    del score_once

    # BASELINE.json configs[4]: hypothesis-count sweep at B = 64, 16 planes/view — both bounds (bytes grow O(H),
    # scoring work O(H^2); the HBM bound is meaningful up to H ~ 256)
    sweep = []
    for nq in (32, 128, 256, 512, 2048):
        try:
            once = score_case(dev, 64, nq, head if nq == NQ else None)
            ab, af = score_algorithmic_bytes(64, nq, nq), score_algorithmic_flops(64, nq, nq)
            ms_s = graph_timed(once, 10, flush if ab < (200 << 20) else None)
            sweep.append({"hypotheses": nq, "pairs": 64, "us_per_call": ms_s * 1e3, "algorithmic_bytes": ab,
                          "achieved_gbs": ab / ms_s / 1e6, "frac_of_hbm": ab / ms_s / 1e6 / peaks["hbm"],
                          "algorithmic_flops": af, "achieved_tflops": af / ms_s / 1e9,
                          "frac_of_bf16_burst": af / ms_s / 1e9 / peaks["tf_burst"], "flop_per_byte": af / ab})
            del once
        except Exception as e:  # noqa: BLE001
            sweep.append({"hypotheses": nq, "error": f"{type(e).__name__}: {e}"[:160]})

    try:
        roofline_tensor = measure_gemm_engine(dev, peaks)
    except Exception as e:  # noqa: BLE001
        roofline_tensor = {"error": f"{type(e).__name__}: {e}"[:200]}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = cpu_reference_pairs_per_s(args.cpu_pairs, threads)
        cpu_baseline = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                        "sample": f"{args.cpu_pairs} pairs of the same workload (S5) after 1 warm-up pair, per-pair loop at "
                                  f"batch size 1 (oracle/backbone_restate.py + oracle/restate.py), {dt:.1f} s"}

    # row f1 (the step before the path): plane lists of both views of this rank's pairs from synthetic PlaneTRHead outputs,
    # timed separately (NOT part of `value`, whose plane lists are inputs).  Reported only.
    try:
        plane_lists = measure_plane_lists(dev, 2 * B)
    except Exception as e:   # never let the side measurement take the bench line down
        plane_lists = {"error": f"{type(e).__name__}: {e}"[:200]}

    try:
        full_model = measure_full_model(dev, B)
    except Exception as e:   # noqa: BLE001
        full_model = {"error": f"{type(e).__name__}: {e}"[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": B, "planes_per_view": PLANES, "num_object_queries": NQ,
                   "stage_set": "S5", "l2": "256 MB L2 flush (buffer write) before every timed step, inside the timed region",
                   "parallelism": f"pairs sharded over {world} GPU(s), 64/GPU; result exchange = {exchange_kind}"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "note": "pinned host inputs (uint8 RGB of both views + plane lists); H2D of step i+1 overlaps the "
                "kernels of step i"},
        "gpu_launches": launches,
        "exchange_verified": None if exchange_check is None else exchange_check["exchange_verified"],
        "exchange_check": exchange_check,
        "roofline": roofline,
        "roofline_tensor": roofline_tensor,
        "score_sweep": sweep,
        "cpu_baseline": cpu_baseline,
        "pose_err": pose_err,
        "from_feature_maps": s4,
        "plane_lists": plane_lists,
        "full_model": full_model,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

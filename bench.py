#!/usr/bin/env python
"""Bench of the NopeSAC one-plane RANSAC pose path on B200 (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" = one pass of the hot path (PlaneCameraHead.inference_Joint incl. the matching head: pixel pose
network -> AIM -> GNN + Sinkhorn matcher -> geo sequences -> hypothesis generation -> scoring -> soft
aggregation -> assignment pruning; stage set S4 of SURVEY.md §8d) over one batch of synthetic pairs.
Workload at every N: BASELINE.json configs[1] per GPU — 64 pairs, 16 planes/view, NUM_OBJECT_QUERIES = 256 with
all 16x16 candidate plane pairs as one-plane hypotheses (257 hypotheses x 256 residual columns per pair),
backbone feature maps of a 480x640 input.  N > 1 (torchrun): pairs are sharded 64/GPU (weak scaling) and
every step ends with ONE NCCL all-gather of the [64,16] per-pair results — the only collective of the path.

value    pairs/s with all inputs resident in HBM (inputs 2.2 GB/GPU >> 126 MB L2: no flush needed).
e2e      pairs/s through the same public call with HOST (pinned) inputs: H2D of the step's inputs and D2H of the
         [B,16] result inside the timed region.
roofline the hypothesis-scoring kernel (nsac_score_aggregate) timed alone with CUDA events on a launch that
         moves > 256 MB (B=512, m=NQ=256), algorithmic bytes of SURVEY.md §8d over measured HBM copy bandwidth.
cpu_baseline / --impl reference: the CPU oracle port (oracle/restate.py — the reference is Python and cannot
         travel to the GPU box) on the host cores, per-pair loop at batch size 1 like the reference.
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
WORKLOAD = ("camera head S4 (pixel pose net + AIM + GNN/Sinkhorn matcher + 256 one-plane hypotheses + scoring + soft "
            "aggregation + pruning), 64 synthetic 480x640 pairs/GPU, 16 planes/view, NUM_OBJECT_QUERIES=256 (all 16x16 "
            "plane pairs as hypotheses), fp32, random-init weights, from backbone feature maps")
METRIC = "image-pairs/sec (480x640, 16 planes x 256 hyp)"


def score_algorithmic_bytes(B, m, nq):
    """per pair 24m + (28 + 8C)(m+1) + 68, C=256, + score-MLP weights once per launch."""
    per_pair = 24 * m + 2076 * (m + 1) + 68
    weights = 8 * (nq * 128 + 128 + 128 * 128 + 128 + 128 * 64 + 64 + 64 + 1)
    return B * per_pair + weights


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.proc = gpu_index, [], None

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

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
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
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------
def cpu_reference_pairs_per_s(num_pairs: int, threads: int, warm: int = 1):
    """The oracle port on the host: per-pair loop at bs=1 (the only mode the reference supports), same
    workload shape, same weights."""
    from oracle import restate
    from nopesac_b200 import synthetic
    from tests import util
    torch.set_num_threads(threads)
    sd, msd = util.make_weights(NQ)
    hp = synthetic.all_pairs_hypotheses(PLANES, NQ)
    batches = [synthetic.make_batch(1000 + i, 1, PLANES, with_features=True) for i in range(num_pairs + warm)]

    def one(b):
        with torch.no_grad():
            return restate.inference_joint(sd, msd, b.feats1, b.feats2, b.planes1, b.planes2, b.app1, b.app2,
                                           num_queries=NQ, hyp_pairs=hp)
    for b in batches[:warm]:
        one(b)
    t0 = time.perf_counter()
    for b in batches[warm:]:
        one(b)
    dt = time.perf_counter() - t0
    return num_pairs / dt, dt


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 4
    vals = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_pairs_per_s(1, threads, warm=1)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, _ = cpu_reference_pairs_per_s(sample, threads, warm=0)
        vals.append(v)
    wall = time.perf_counter() - t_all
    value = statistics.mean(vals)
    desc = f"{sample} pairs/step x {args.steps} steps of the same workload, per-pair loop at batch size 1"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": sample, "device": "cpu"},
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


def measure_from_rgb(dev, B: int, head, match, dbatch, hp, iters: int = 3):
    """pairs/s of RGB -> ResNet-50 (random init, both views = 2B images of 480x640) -> camera head, inputs resident in HBM,
    CUDA events on the launch stream after 2 warm-ups."""
    from nopesac_b200 import backbone, config
    net = backbone.build_backbone(config.inference_cfg())
    with torch.no_grad():                              # random init: damp the residual branches so res5 stays O(1) like a trained net
        for name, buf in net.named_buffers():
            if name.endswith("conv3.norm.weight"):
                buf.mul_(0.3)
    net = net.to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    images = torch.rand(2 * B, 3, 480, 640, device=dev, generator=g) * 255

    def step():
        feats = net(images)
        f1 = {k: v[:B] for k, v in feats.items()}
        f2 = {k: v[B:] for k, v in feats.items()}
        return head(f1, f2, dbatch.planes1, dbatch.planes2, dbatch.app1, dbatch.app2, matching_net=match, hyp_pairs=hp)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"value": B / ms * 1e3, "unit": "pairs/s", "ms_per_step": ms, "pairs": B, "stage_set": "S5' = ResNet-50 backbone (random init) + "
            "camera head from RGB, synthetic plane lists", "note": "first version of the backbone (residual add unfused); reported, not the headline"}


def run_side_measurement(which: str, B: int, timeout_s: int):
    """`python bench.py --side <which> --pairs B` in a child process; returns its JSON dict or {"error": ...}."""
    try:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--side", which, "--pairs", str(B)], capture_output=True,
                             text=True, timeout=timeout_s, env={**os.environ, "WORLD_SIZE": "1", "RANK": "0", "LOCAL_RANK": "0"})
        for ln in reversed(res.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": f"child exited {res.returncode}: {(res.stderr or res.stdout)[-160:]}"}
    except subprocess.TimeoutExpired:
        return {"error": f"timed out after {timeout_s} s"}
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:200]}


def side_main(args):
    """Child-process entry of the side measurements (own CUDA context)."""
    from nopesac_b200 import synthetic
    from tests import util
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B = args.pairs
    if args.side == "from_rgb":
        head, match, _, _ = util.build_cuda_heads(NQ, "soft", 0.2, dev)
        hp = synthetic.all_pairs_hypotheses(PLANES, NQ).to(dev, torch.int32)
        dbatch = synthetic.make_batch(0, B, PLANES).to(dev)
        out = measure_from_rgb(dev, B, head, match, dbatch, hp)
    else:
        out = measure_plane_lists(dev, 2 * B)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-value", action="store_true", help="profiling runs: skip the e2e / roofline / cpu legs")
    ap.add_argument("--cpu-pairs", type=int, default=8, help="pairs timed by the cpu_baseline leg")
    ap.add_argument("--side", default=None, choices=["from_rgb", "plane_lists"], help="internal: one side measurement, one JSON dict")
    args = ap.parse_args()
    if args.side:
        side_main(args)
        return
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    from nopesac_b200 import ops, synthetic
    from tests import util   # seeded weights (shapes from tests/golden/state_shapes.json)

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.pairs

    head, match, _, _ = util.build_cuda_heads(NQ, "soft", 0.2, dev)
    hp = synthetic.all_pairs_hypotheses(PLANES, NQ).to(dev, torch.int32)
    host = synthetic.make_batch(rank * B, B, PLANES)                 # planes / appearance: seeded on the host
    f1, f2 = synthetic.device_features(B, dev, seed=7 + rank)       # 2.2 GB of feature maps: drawn on the device
    dbatch = host.to(dev)
    gathered = torch.empty(world * B, 16, device=dev) if world > 1 else None
    exchange, exchange_kind = None, "single GPU: no exchange"
    if world > 1:
        exchange_kind = "one NCCL all-gather of [B,16] result rows per step"
        if os.environ.get("NSAC_EXCHANGE", "fused") == "fused":
            try:
                from nopesac_b200.dist import FusedResultExchange
                exchange = FusedResultExchange(B, dev)
                exchange_kind = ("fused: the selection kernel stores each result row into every rank's buffer over NVLink "
                                 "peer memory (symmetric memory) + one cross-rank barrier per step; no collective")
            except Exception as e:  # noqa: BLE001  (no P2P / symmetric memory on this box)
                exchange = None
                exchange_kind += f" (fused exchange unavailable: {type(e).__name__})"

    def step(p1, p2, a1, a2, fa, fb):
        out = head(fa, fb, p1, p2, a1, a2, matching_net=match, hyp_pairs=hp, result_exchange=exchange)
        pose = out[5]["pose"]
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
    for _ in range(args.warmup):
        step(dbatch.planes1, dbatch.planes2, dbatch.app1, dbatch.app2, f1, f2)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(dbatch.planes1, dbatch.planes2, dbatch.app1, dbatch.app2, f1, f2)
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

    # pose error of this rank's pairs against the planted ground truth (reference formulas, mp3d_evaluation.py:382-425,
    # computed on the device by nopesac_b200.evaluation): reported, not a target - the weights are random
    pose_err = None
    if rank == 0:
        from nopesac_b200 import evaluation
        rows = head(f1, f2, dbatch.planes1, dbatch.planes2, dbatch.app1, dbatch.app2, matching_net=match, hyp_pairs=hp)[5]["pose"]
        pm = evaluation.camera_metrics(rows, dbatch.gt_tran, dbatch.gt_quat)
        pose_err = {"T_median_m": pm["T median err"], "T_mean_m": pm["T mean err"], "R_median_deg": pm["R median err"],
                    "R_mean_deg": pm["R mean err"], "pairs": B, "note": "random-init weights vs planted GT: reported, not a target"}

    # ------------------------------------------------------------------ multi-GPU: the exchanged rows are the right rows
    # (replaces the reference's pickled comm.gather, mp3d_evaluation.py:317-318; SURVEY.md §4 "all-gather equality vs
    # single-rank concatenation").  One extra step outside the timed region: (1) the rows the fused NVLink exchange left in
    # this rank's buffer == one NCCL all-gather of every rank's local rows, bit for bit, on EVERY rank; (2) the LAST rank
    # recomputes rank 0's shard from rank 0's seeds on its own GPU and compares it with rows [0, B) it received.
    exchange_check = None
    if world > 1:
        out = head(f1, f2, dbatch.planes1, dbatch.planes2, dbatch.app1, dbatch.app2, matching_net=match, hyp_pairs=hp,
                   result_exchange=exchange)
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
            g1, g2 = synthetic.device_features(B, dev, seed=7)
            r0 = head(g1, g2, h0.planes1, h0.planes2, h0.app1, h0.app2, matching_net=match, hyp_pairs=hp)[5]["pose"]
            rec_equal = bool(torch.equal(r0, got[:B]))
            rec_diff = float((r0 - got[:B]).abs().max())
            del g1, g2
        flags = torch.tensor([1.0 if same else 0.0, 1.0 if rec_equal else 0.0, -rec_diff], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        exchange_check = {"exchange_verified": bool(flags[0].item() == 1.0), "ranks": world,
                          "rows": [world * B, 16], "what": "rows left by the exchange == NCCL all-gather of the local rows (torch.equal, "
                          "every rank) and each rank's own shard sits at its block offset",
                          "recompute_of_rank0_shard_on_last_rank": {"bit_equal": bool(flags[1].item() == 1.0),
                                                                    "max_abs_diff": float(-flags[2].item())}}

    if args.only_value:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
                              "ms_per_step": ms_total / args.steps, "gpu_launches": launches, "partial": True}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ e2e: host (pinned) inputs
    # Same public call, inputs in pinned host memory: every step copies its 2.2 GB of inputs H2D and reads the
    # [B,16] result rows back; nopesac_b200.runtime.PairPipeline overlaps the copy of step i+1 with the kernels of
    # step i (two device slots, copy stream + events).
    from nopesac_b200.runtime import PairPipeline, batch_bytes, pin_batch
    hbatch = pin_batch({"planes1": host.planes1, "planes2": host.planes2, "app1": host.app1, "app2": host.app2,
                        "feats1": {k: v.cpu() for k, v in f1.items()}, "feats2": {k: v.cpu() for k, v in f2.items()}})
    h2d = batch_bytes(hbatch)
    d2h = (world * B if world > 1 else B) * 16 * 4
    post = None
    if world > 1:
        def post(rows):
            if exchange is not None:
                return exchange.finish()
            dist.all_gather_into_tensor(gathered, rows)
            return gathered
    pipe = PairPipeline(head, match, dev, hyp_pairs=hp, post=post, result_exchange=exchange)
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

    # ------------------------------------------------------------------ roofline: scoring kernel alone
    peak, peak_src = measured_peaks()
    RB, m = 512, NQ
    g = torch.Generator(device=dev).manual_seed(3)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
    geo_local = rnd(RB, NQ, 6)
    q_h = torch.nn.functional.normalize(rnd(RB, NQ, 4), dim=-1)
    t_h = rnd(RB, NQ, 3) * 0.3
    q0 = torch.nn.functional.normalize(rnd(RB, 4), dim=-1)
    t0 = rnd(RB, 3) * 0.3
    fr, ft = rnd(RB, NQ, 256), rnd(RB, NQ, 256)
    fr0, ft0 = rnd(RB, 256), rnd(RB, 256)
    mnum = torch.full((RB,), m, device=dev, dtype=torch.int32)
    pk = head.prepare_tc()

    def score_once():
        return ops.score_aggregate(geo_local, q_h, t_h, q0, t0, fr, ft, fr0, ft0, mnum, pk["normal_score_proj"],
                                   pk["param_score_proj"], head.rots.weight, head.rots.bias, head.trans.weight,
                                   head.trans.bias, out_cam_type="soft", want_scores=False, pack=pk["score_pack"])
    for _ in range(3):
        score_once()
    torch.cuda.synchronize()
    # One scoring call = 3 short kernels (prep, tiles, selection).  Launched eagerly from Python the host side (ctypes +
    # tensor allocation) is slower than the GPU, so the call is captured once in a CUDA graph and the replays are timed
    # with CUDA events on the replay stream: kernel time, not Python time.
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        score_once()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        score_once()
    graph.replay()
    torch.cuda.synchronize()
    reps = 20
    e0.record()
    for _ in range(reps):
        graph.replay()        # 276 MB of inputs per launch > L2: no flush needed
    e1.record()
    torch.cuda.synchronize()
    score_ms = e0.elapsed_time(e1) / reps
    alg = score_algorithmic_bytes(RB, m, NQ)
    achieved = alg / (score_ms / 1e3) / 1e9
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of the same launch from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "score_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roofline = {"kernel": "nsac_score_aggregate_tc (K8+K9: prep + tile + selection kernels), B=512, m=NQ=256", "bound": "hbm",
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": alg, "ms_per_launch": score_ms,
                "timing": "CUDA-graph replays of one call, CUDA events on the replay stream"}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = cpu_reference_pairs_per_s(args.cpu_pairs, threads)
        cpu_baseline = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                        "sample": f"{args.cpu_pairs} pairs of the same workload after 1 warm-up pair, per-pair loop at "
                                  f"batch size 1 (oracle/restate.py), {dt:.1f} s"}

    # row f1 (the step before the path): plane lists of both views of this rank's pairs from synthetic PlaneTRHead outputs,
    # timed separately (NOT part of `value`, whose stage set S4 starts at the backbone / plane-head outputs).  Reported only.
    plane_lists = None
    if rank == 0:
        try:
            plane_lists = measure_plane_lists(dev, 2 * B)
        except Exception as e:   # never let the side measurement take the bench line down
            plane_lists = {"error": f"{type(e).__name__}: {e}"[:200]}

    # row f2: the same step started at RGB (ResNet-50 backbone of both views + camera head) — the stage set S5' of DESIGN.md
    # §5.  Reported beside `value` (whose stage set S4 starts at the backbone feature maps), never instead of it.  This path has
    # not met a GPU before this run, so it is measured in a CHILD process with a timeout: a hang or a sticky CUDA error there
    # cannot take this line down.
    from_rgb = None
    if rank == 0 and world == 1:
        from_rgb = run_side_measurement("from_rgb", B, timeout_s=240)

    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": B, "planes_per_view": PLANES, "num_object_queries": NQ,
                   "stage_set": "S4", "l2": "inputs (2.2 GB/GPU) exceed the 126 MB L2; no flush",
                   "parallelism": f"pairs sharded over {world} GPU(s), 64/GPU; result exchange = {exchange_kind}"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "note": "pinned host inputs; H2D of step i+1 overlaps the kernels of step i"},
        "gpu_launches": launches,
        "exchange_verified": None if exchange_check is None else exchange_check["exchange_verified"],
        "exchange_check": exchange_check,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "pose_err": pose_err,
        "plane_lists": plane_lists,
        "from_rgb": from_rgb,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

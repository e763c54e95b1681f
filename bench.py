#!/usr/bin/env python
"""bench.py — trees/sec of Argmax inference on 50-taxa x 1024-site alignments (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 512] [--impl ours|reference]

One "step" = one pass of the hot path (encode + 49 learned-NJ steps) over a batch of synthetic MSAs
(config 2: iid tokens over A,C,G,T,gap, seed 1234; weights torch.manual_seed(0) default init — the shipped
checkpoint is a missing blob).  At N GPUs every rank processes its own `--batch` alignments (alignments are
independent: no collective on the data path, "weak" scaling); value = all trees / max-over-ranks device time.

Printed keys beyond the base contract:
  e2e          same metric through the C-ABI host-buffer entry point nnj_rollout_host (pinned host int8 MSA in,
               merge lists out; H2D/D2H and workspace allocation inside the timed region)
  roofline     the kernel class with the largest share of the step, timed live with CUDA events on the launching
               stream (nnj_profile_*), algorithmic FLOPs / launch from BASELINE.md section 3
  kernels      per-class milliseconds / launches / share of the profiled step
  cpu_baseline the CPU oracle (port of the reference, oracle/nnj_oracle.py) timed on this box's host cores
  --impl reference : the same oracle as the reference arm (the reference itself needs /root/reference, absent here)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_TAXA, L_SITES, D, H = 50, 1024, 64, 8
METRIC = "trees/sec, 50-taxa x 1024-site Argmax inference"
UNIT = "trees/s"


def synthetic_msa(batch, taxa, sites, seed):
    """Config-2 generator: iid tokens, p(A,C,G,T,-) = (.28,.12,.12,.21,.27) -> one-hot int8, gap = 1111."""
    import torch
    g = torch.Generator().manual_seed(seed)
    p = torch.tensor([0.28, 0.12, 0.12, 0.21, 0.27])
    tok = torch.multinomial(p, batch * taxa * sites, replacement=True, generator=g).view(batch, taxa, sites)
    table = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 1, 1, 1]], dtype=torch.int8)
    return table[tok]


def algorithmic_flops(R, Cc, layers=6):
    """Per-tree algorithmic FLOPs per kernel class (2*M*N*K per contraction, elementwise excluded).
    NJ classes use the folded formulation the kernels implement (W_h and W_q folded per node)."""
    F = 4 * D
    pair_evals = [(R * (R - 1) // 2, R)] + [(n, n) for n in range(R - 1, 1, -1)]     # (pairs, nodes) per step
    fl = {
        "embed": 2 * R * Cc * (4 * D + D * D),
        "ln_qkv": layers * 2 * 3 * 2 * R * Cc * D * D,
        "out_proj": layers * 2 * 2 * R * Cc * D * D,
        "row_qk_gemm": layers * 2 * Cc * Cc * R * D,
        "row_pv_gemm": layers * 2 * Cc * Cc * R * D,
        "col_attn": layers * 4 * Cc * R * R * D,
        "ffn": layers * 4 * R * Cc * D * F,
        "alpha": sum(p * 2 * n * Cc * D for p, n in pair_evals if n > 2),
        "pair_score": sum(p * ((2 * n * Cc * D + 2 * Cc * D * D if n > 2 else 0) + 2 * Cc * D * D + 2 * Cc * D) for p, n in pair_evals),
        "merge": sum(2 * 2 * n * Cc * D + 2 * Cc * D * D + 3 * 2 * Cc * D * D for n in range(R, 2, -1)),   # incl. the merge pair's alpha
        "node_derive": R * 3 * 2 * Cc * D * D,
    }
    return fl


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        if not sm:
            return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def cpu_oracle_trees_per_sec(n_trees, seed=1234):
    """Time the CPU oracle (port of the reference path) on `n_trees` config-2 alignments, one per call like the reference (B=1)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nnj_oracle as O
    sd = O.init_state_dict(0)
    data = synthetic_msa(n_trees, R_TAXA, L_SITES, seed)
    mask = torch.zeros(1, L_SITES, dtype=torch.bool)
    t0 = time.perf_counter()
    for b in range(n_trees):
        O.rollout(sd, data[b:b + 1], mask)
    dt = time.perf_counter() - t0
    return n_trees / dt, dt, torch.get_num_threads()


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  /root/reference is not on the GPU box, so this
    is the oracle port (validated against the executed reference, tests/golden) with all host threads torch will use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    sample = 1
    for _ in range(args.warmup):
        cpu_oracle_trees_per_sec(sample)
    t0 = time.perf_counter()
    for k in range(args.steps):
        cpu_oracle_trees_per_sec(sample, seed=1234 + k)
    dt = time.perf_counter() - t0
    tps = args.steps * sample / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"configs[1]: synthetic MSAs {R_TAXA} taxa x {L_SITES} sites, Argmax; each step = {sample} alignment (bounded sample, B=1 per call as shipped)"},
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{args.steps} x {sample} alignment(s) of the config-2 generator"},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=512, help="alignments per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-trees", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as graft
    graft.build()
    from neuralnj_b200 import PhyloATTN, inference_config, _lib
    L = _lib.lib()
    torch.manual_seed(0)
    model = PhyloATTN(inference_config(), precision=args.precision).to(dev).eval()
    B = args.batch
    data_host = synthetic_msa(B, R_TAXA, L_SITES, 1234 + rank).pin_memory()
    mask_host = torch.zeros(B, L_SITES, dtype=torch.bool).pin_memory()
    data = data_host.to(dev)
    mask = mask_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return model.rollout_fused(data, mask)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    L.nnj_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        merges, slp, _ = step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = int(L.nnj_launch_count(0))
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- end to end through the host-buffer C-ABI entry point
    model.rollout_host(data_host, mask_host)       # warm
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        mh = model.rollout_host(data_host, mask_host)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = world * B * e2e_steps / float(e2e_s)
    assert torch.equal(mh, merges.cpu()), "host entry point and device path disagree"

    # ---- per-kernel-class timing (one extra step with CUDA events around every launch)
    n_cls = L.nnj_profile_classes()
    L.nnj_profile_enable(1)
    step()
    cls_ms = (C.c_double * n_cls)()
    cls_n = (C.c_int64 * n_cls)()
    _lib.check(L.nnj_profile_read(n_cls, cls_ms, cls_n))
    L.nnj_profile_enable(0)
    names = [L.nnj_profile_name(i).decode() for i in range(n_cls)]
    tot_ms = sum(cls_ms)
    fl = algorithmic_flops(R_TAXA, L_SITES)
    kernels = {n: {"ms": round(cls_ms[i], 3), "launches": int(cls_n[i]), "share": round(cls_ms[i] / tot_ms, 4)} for i, n in enumerate(names) if cls_n[i]}
    top = max((n for n in kernels if n in fl), key=lambda n: kernels[n]["ms"])
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    top_flops_per_launch = fl[top] * B / kernels[top]["launches"]
    top_ms_per_launch = kernels[top]["ms"] / kernels[top]["launches"]
    achieved = top_flops_per_launch / (top_ms_per_launch * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": top, "achieved": round(achieved, 3), "peak": peak_tf, "unit": "TFLOP/s",
                "frac": round(achieved / peak_tf, 5), "traffic": None, "peak_source": peak_src,
                "flops_per_launch": top_flops_per_launch, "ms_per_launch": round(top_ms_per_launch, 4),
                "note": "algorithmic (1x) FLOPs against the dense bf16 tensor peak; CUDA-core fp32 kernels and 3x split-bf16 tcgen05 kernels both count 1x"}
    for n in kernels:
        if n in fl:
            kernels[n]["tflops"] = round(fl[n] * B / (kernels[n]["ms"] * 1e-3) / 1e12, 3)
    # DRAM traffic per launch.  The NJ kernels stream only the n live node slots of a step, so their bytes change from step to
    # step: the mean over a rollout comes from the algorithmic byte count (node tiles of the live slots + x planes), which the
    # committed `ncu --set full` captures reproduce within 1 % at the captured launches (profiles/r01_traffic_b128.json; its
    # "model" ratios are printed below).  Step-0 pair-score launches (two pair tiles re-read the node tile) use the ncu figure.
    roofline_hbm = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic_b128.json")) as f:
            tr = json.load(f)
        chunks = max(1, kernels["node_derive"]["launches"])
        trees = B / chunks                                   # trees per launch
        site_bytes = 256 * L_SITES                           # one fp32 [sites x 64] row set: 256 KB at 1024 sites
        step0_pairs = [256] * (R_TAXA * (R_TAXA - 1) // 2 // 256) + [R_TAXA * (R_TAXA - 1) // 2 % 256]

        def alpha_bytes(n, pairs):       # X, Y fp32 + K' bf16 hi/lo of n slots, x planes of the pairs
            return trees * (3 * n + pairs) * site_bytes

        def score_inc_bytes(n, pairs):   # [X | W_g X] bf16 hi/lo of n slots, x tiles of the pairs
            return trees * (2 * n + pairs) * site_bytes
        alpha_tot = sum(alpha_bytes(R_TAXA, p) for p in step0_pairs) + sum(alpha_bytes(n, n) for n in range(R_TAXA - 1, 2, -1))
        alpha_n = len(step0_pairs) + R_TAXA - 3
        score_tot = sum(tr["score_step0"]["dram_bytes"] * trees / 128.0 * (p / 256.0 * 0.6 + 0.4) for p in step0_pairs) + \
            sum(score_inc_bytes(n, n) for n in range(R_TAXA - 1, 2, -1))
        score_n = len(step0_pairs) + R_TAXA - 3 + 1
        model = {"alpha_incr_model_over_ncu": round(alpha_bytes(34, 34) * 128.0 / trees / tr["alpha_incr"]["dram_bytes"], 3),
                 "score_incr_model_over_ncu": round(score_inc_bytes(34, 34) * 128.0 / trees / tr["score_incr"]["dram_bytes"], 3)}
        means = {"alpha": alpha_tot / alpha_n, "pair_score": score_tot / score_n}
        if top in means:
            roofline["traffic"] = round(means[top])
            roofline["traffic_source"] = "mean over the launches of a rollout: algorithmic bytes of the live node slots + x planes (= ncu dram__bytes at the captured launches, profiles/r01_traffic_b128.json), ncu figure for the step-0 launches"
        if "alpha" in kernels:
            hbm_peak = peaks.get("hbm_gbs", 6554.2)
            a_ms = kernels["alpha"]["ms"] / kernels["alpha"]["launches"]
            a_bytes = means["alpha"] * alpha_n * chunks / kernels["alpha"]["launches"]
            roofline_hbm = {"bound": "hbm", "kernel": "alpha", "achieved": round(a_bytes / (a_ms * 1e-3) / 1e9, 1), "peak": hbm_peak,
                            "unit": "GB/s", "frac": round(a_bytes / (a_ms * 1e-3) / 1e9 / hbm_peak, 4), "traffic": round(a_bytes),
                            "ms_per_launch": round(a_ms, 4), "model_check": model,
                            "note": "k_alpha_v3: mean bytes per launch = live node slots (X, Y, K') once + x planes; the late steps (few live nodes) are latency-bound, which pulls the mean below the 6.1-6.3 TB/s of the early launches"}
    except (OSError, KeyError, ValueError, ZeroDivisionError):
        pass

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tps, secs, cores = cpu_oracle_trees_per_sec(args.cpu_trees)
        cpu_baseline = {"value": round(tps, 4), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{args.cpu_trees} alignments of the same generator, B=1 per call, {secs:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16x3 (split-bf16 tcgen05, fp32 accumulate) + f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: {B} synthetic MSAs per GPU per step, {R_TAXA} taxa x {L_SITES} sites, Argmax, sharded by alignment",
                       "global_batch": B * world, "parallelism": f"alignment-sharded x{world}, no collectives",
                       "weights": "torch.manual_seed(0) default init (checkpoint blob absent)", "precision": args.precision,
                       "l2": "inputs_larger_than_l2 (105 MB int8 MSA + 6.7 GB fp32 node state per step)"},
            "e2e": {"value": round(e2e_val, 3), "unit": UNIT, "h2d_bytes_per_step": int(data_host.numel() + mask_host.numel()),
                    "d2h_bytes_per_step": int(mh.numel() * 4), "api": "nnj_rollout_host (C ABI, pinned host buffers)", "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "roofline_hbm": roofline_hbm,
            "kernels": kernels,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

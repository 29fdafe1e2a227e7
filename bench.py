#!/usr/bin/env python
"""bench.py -- headline benchmark of the within-group dN/dS hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at N=1: BASELINE config 2 -- SARS-CoV-2-shaped within-group analysis,
10,000 sequences x 29,903 nt, 10 CDS (9,677 codons), 10,000 bootstrap replicates
per product.  One "step" is one pass of the whole hot path over that alignment:
codon packing -> per-codon histograms -> Nei-Gojobori per-codon statistics ->
product sums -> whole-product bootstraps -> SE moments.

  value  nominal codon-pair comparisons per second (the reference's loop trip
         count, sum over products of C_p * C(n,2)) with the
         alignment already resident in HBM.
  e2e    the same metric through the host-buffer C-ABI call (snpg_load_group from
         pinned host memory + snpg_within_all + result read-back).
At N>1 (torchrun, one rank per GPU): weak scaling by codon range -- every rank
owns one genome-shaped codon range (its own 10 CDS of the same 10,000 sequences),
no data-path collective; the per-product results are gathered once over NCCL.

--impl reference times the CPU restatement of the reference's own algorithm
(oracle/ng_oracle.c: the literal O(n^2) pair loop, all host threads) on a bounded
sample of the same workload.  The reference itself is Perl, cannot travel to the
GPU box and has nothing to compile, so kind = "port" (see DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SEQS, ALN_LEN, N_BOOT = 10000, 29903, 10000
SEED = 20261019 + 2
METRIC = "codon-pair comparisons/sec"
UNIT = "codon-pair comparisons/s"


def nominal_pairs(n, codons):
    return codons * (n * (n - 1) // 2)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm (oracle port)

def cpu_pair_rate(target_s=12.0):
    """Literal pair loop (oracle) on a bounded sample of config 2: all 10,000 sequences, a slice of codons."""
    from oracle import oracle as orc
    from snpgenie_b200 import synth
    products = synth.sars2_products()
    threads = orc.use_all_cores()
    # the sample keeps all n sequences (cost is quadratic in n) and takes the first k codons of CDS3 (spike-shaped)
    p = products[2]
    k_probe = max(threads, 8)

    def run(k):
        width = 3 * k
        sub = synth.make_alignment(N_SEQS, p.starts[0] - 1 + width, [type(p)(p.name, 1, [p.starts[0]], [p.starts[0] + width - 1])], SEED)
        pos = np.arange(p.starts[0] - 1, p.starts[0] - 1 + width, dtype=np.int64)
        raw, _ = orc.extract(sub, pos, 1)
        t0 = time.perf_counter()
        orc.within_pairloop(raw)
        return time.perf_counter() - t0

    t_probe = run(k_probe)
    k = int(max(k_probe, min(2000, k_probe * target_s / max(t_probe, 1e-3))))
    t = run(k)
    if t < 0.5 * target_s and k < 2000:      # the probe under-uses the cores: scale once more from the real rate
        k = int(min(2000, k * target_s / max(t, 1e-3)))
        t = run(k)
    return nominal_pairs(N_SEQS, k) / t, threads, f"{k} codons x {N_SEQS} seqs of config 2 (literal O(n^2) pair loop, {t:.1f} s)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    per_step = max(1.0, min(20.0, 90.0 / (steps + args.warmup)))     # the whole run ends within a few minutes
    vals = []
    sample = ""
    threads = 1
    for i in range(args.warmup + steps):
        v, threads, sample = cpu_pair_rate(per_step)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    ms = 1e3 * nominal_pairs(N_SEQS, 9677) / value
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "ms_per_step is the full config-2 step extrapolated from the sampled pair rate",
    }))


def workload_config(n_gpus):
    return {"workload": "config2: SARS-CoV-2-shaped within-group, 10000 seqs x 29903 nt, 10 CDS (9677 codons), 10000 bootstraps"
                        + ("" if n_gpus == 1 else f"; x{n_gpus} genome-shaped codon ranges, one per GPU"),
            "n_seqs": N_SEQS, "aln_len": ALN_LEN, "codons_per_gpu": 9677, "bootstraps": N_BOOT,
            "l2": "inputs larger than L2 (299 MB alignment per step, re-read every step)"}


def bind_near_gpu(local):
    """Pin this rank to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host memory is allocated
    (first touch then places the e2e input next to the GPU's PCIe root).  With 8 ranks on one box and no binding
    half of them copy across the socket interconnect.  Returns (node, n_cpus) or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]                      # 00000000:19:00.0 -> 0000:19:00.0
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node, len(cpus)
    except Exception:
        return None


# ------------------------------------------------------------------ GPU arm

def run_ours(args):
    import torch
    import torch.distributed as dist

    from snpgenie_b200 import _abi, synth
    from snpgenie_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # the contract is ONE JSON line on stdout: NCCL prints its version banner to fd 1, so everything but the
    # final line goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    numa = bind_near_gpu(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")

    products = synth.sars2_products()
    codons = sum(p.n_codons for p in products)
    aln = synth.make_alignment(N_SEQS, ALN_LEN, products, SEED + 1000 * rank, p_poly=args.p_poly)
    pitch = (ALN_LEN + 127) // 128 * 128
    # pinned host copy (e2e input) and HBM-resident copy with a 128-B row pitch (value input)
    h_aln = torch.empty((N_SEQS, ALN_LEN), dtype=torch.uint8).pin_memory()
    h_aln.numpy()[:] = aln
    d_aln = torch.zeros((N_SEQS, pitch), dtype=torch.uint8, device=dev)
    d_aln[:, :ALN_LEN] = h_aln.to(dev)
    torch.cuda.synchronize()

    eng = Engine(device=local, seed=SEED)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    # timed passes bracket only the dominant HBM kernel with events (a pair costs ~3 us of stream time);
    # a separate untimed pass of the same steps brackets every launch for the per-kernel table
    hbm_kinds = ["fused"] if args.route in ("fused", "tiles") else ["pack", "hist"]
    narrow = sum(1 << _abi.KERNEL_KINDS.index(k) for k in hbm_kinds)
    eng.set_option(_abi.OPT_PROFILE, 0 if args.no_profile else narrow)
    eng.set_option(_abi.OPT_KEEP_CODES, 1 if args.route == "codes" else 0)
    eng.set_option(_abi.OPT_FUSED_KERNEL, _abi.FUSED_TMA_TILES if args.route == "tiles" else _abi.FUSED_LDG_SPANS)
    eng.set_products(products, ALN_LEN)

    def step_resident():
        return eng.within_job(0, d_aln.data_ptr(), N_SEQS, ALN_LEN, pitch, True, B=N_BOOT)

    def step_e2e():
        return eng.within_job(0, h_aln.data_ptr(), N_SEQS, ALN_LEN, ALN_LEN, False, B=N_BOOT)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        eng.profile(reset=True)
        l0 = eng.launches
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            print(f"[rank {rank}] {fn.__name__}: {ms / steps:.4f} ms/step", file=sys.stderr)
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, eng.profile(reset=True), eng.launches - l0, out

    with ClockSampler(local) as clocks:      # clocks are sampled across both timed regions
        ms_res, prof, launches, (sums, se) = timed(step_resident, args.steps, args.warmup)
        ms_e2e, prof_e2e, _, _ = timed(step_e2e, max(1, args.steps), max(3, min(args.warmup, 3)))
    prof_all = {}
    if not args.no_profile:
        eng.set_option(_abi.OPT_PROFILE, -1)
        _, prof_all, _, _ = timed(step_resident, args.steps, 1)

    # one small gather of the per-product results (the only exchange on this path)
    res = torch.tensor(np.concatenate([sums, se], axis=1), device=dev)
    if world > 1:
        allres = [torch.empty_like(res) for _ in range(world)]
        dist.all_gather(allres, res)
        res = torch.cat(allres)
    n_products_total = int(res.shape[0])

    if rank == 0:
        pairs_total = nominal_pairs(N_SEQS, codons) * world
        value = pairs_total / (ms_res * 1e-3)
        e2e = pairs_total / (ms_e2e * 1e-3)
        peak, peak_src = peaks()
        # algorithmic bytes per launch: pack reads 3 B and writes 1 B per (sequence, codon);
        # hist reads 1 B per (sequence, codon) and writes 264 B per codon
        alg = {"pack": 4.0 * N_SEQS * codons, "hist": 1.0 * N_SEQS * codons + 264.0 * codons,
               "fused": 3.0 * N_SEQS * codons + 264.0 * codons}   # fused: the CDS bytes read once, histograms written
        kern = {}
        merged = dict(prof_all)
        merged.update({k: v for k, v in prof.items() if v[1]})       # HBM kernels: from the timed pass itself
        for k, (ms, calls) in merged.items():
            if calls:
                kern[k] = {"ms_per_launch": ms / calls, "launches_per_step": calls / args.steps, "ms_per_step": ms / args.steps}
                if k in alg:
                    kern[k]["achieved_gbs"] = alg[k] / (ms / calls * 1e-3) / 1e9   # one launch covers the whole matrix
        hbm_kernels = {k: v for k, v in kern.items() if k in alg}
        dom = max(hbm_kernels, key=lambda k: hbm_kernels[k]["ms_per_step"]) if hbm_kernels else ("pack" if args.route == "codes" else "fused")
        achieved = kern[dom]["achieved_gbs"] if dom in kern else float("nan")
        tiled = eng.tile_launches > 0          # SNPG_K_FUSED launches went to the TMA-tiled kernel
        kernel_name = {"pack": "k_pack", "hist": "k_hist", "fused": "k_tile_hist" if tiled else "k_fused_hist"}[dom]
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get("tile" if (dom == "fused" and tiled) else dom)
        except Exception:
            pass
        boot_ms = sum(kern[k]["ms_per_step"] for k in ("bootstrap", "moments") if k in kern)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 codon codes / u32 histograms / f64 statistics", "data": "synthetic",
            "config": dict(workload_config(world), **({} if args.p_poly == 0.3 else {"p_poly": args.p_poly, "side_experiment": True})),
            # the library stages only the CDS columns of every row: from the first CDS column rounded down to 16 to the last
            "e2e": {"value": e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(N_SEQS * (max(max(p.stops) for p in products) - (min(min(p.starts) for p in products) - 1) // 16 * 16)),
                    "d2h_bytes_per_step": int(len(products) * 10 * 8), "ms_per_step": ms_e2e,
                    "host_buffer_bytes": int(N_SEQS * ALN_LEN)},
            "gpu_launches": int(launches),
            "route": args.route,
            "roofline": {"kernel": kernel_name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg[dom]},
            "kernels": kern,
            "kernel_timing": "CUDA events inside the library on the launch stream: " + "/".join(hbm_kinds)
                             + " during the timed steps; the other kernels in a separate pass of the same steps",
            "bootstrap": {"value": N_BOOT * len(products) * world / (boot_ms * 1e-3) if boot_ms else None, "unit": "bootstrap reps/s",
                          "note": "B x products per step / (bootstrap + moments kernel time)",
                          # the weight-matrix kernel's contraction: one mma.m8n8k4.f64 (512 flop, half of it padding) per
                          # 8 replicates x 4 slots; its share of the kernel is in profiles/ (tensor pipe active 28 %)
                          "dmma_tflops_in_kernel": (512 * sum(-(-N_BOOT // 8) * -(-(len_p + 1) // 4) for len_p in (p.n_codons for p in products))
                                                    / (kern["bootstrap"]["ms_per_step"] * 1e-3) / 1e12) if "bootstrap" in kern else None},
            "clocks": clocks.summary(),
            "products_gathered": n_products_total,
            "host_binding": {"numa_node": numa[0], "cpus": numa[1]} if numa else None,
        }
        if world == 1 and not args.no_cpu:
            v, threads, sample = cpu_pair_rate(12.0)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--p-poly", type=float, default=0.3,
                    help="fraction of polymorphic codon columns in the synthetic alignment (SURVEY.md 8d: 0.3; "
                         "other values are side experiments, never the bench line)")
    ap.add_argument("--no-profile", action="store_true", help="do not bracket kernels with CUDA events (no roofline block)")
    ap.add_argument("--route", default="fused", choices=["fused", "tiles", "codes"],
                    help="histogram route: fused vector-load kernel (default), fused TMA-tiled kernel, or K1 pack + K2 hist "
                         "over a kept code matrix")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

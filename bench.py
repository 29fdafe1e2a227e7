#!/usr/bin/env python
"""bench.py -- headline benchmark of the within-group dN/dS hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload: BASELINE config 2 -- SARS-CoV-2-shaped within-group analysis, 10,000 sequences x 29,903 nt, 10 CDS
(9,677 codons), 10,000 bootstrap replicates per product.  One "step" is one pass of the whole hot path over that
alignment through ONE C-ABI call (snpg_within_job_ex): codon histograms -> Nei-Gojobori per-codon statistics ->
product sums -> whole-product bootstraps -> SE moments, with the per-codon table (poly, n_defined, comparisons,
N/S sites and differences, conserved codon) and the product table read back into pinned host memory.

  value  nominal codon-pair comparisons per second (the reference's loop trip count, sum over products of
         C_p * C(n,2)) with the alignment already resident in HBM.
  e2e    the same metric with the alignment in (pinned) HOST memory: the call stages the CDS columns over PCIe.

At N > 1 (torchrun, one rank per GPU) the SAME alignment is split N ways -- strong scaling: rank r owns the r-th
N-th of the codon axis (it stages and reads only those columns) and the r-th N-th of every product's bootstrap
replicates.  The exchange is inside the step and inside the library: the kernels that produce the per-codon
statistics and the per-replicate values store them into every rank's HBM over NVLink (CUDA IPC peer mappings) and
flag their arrival; no host round trip, no separate collective launch.  Every rank ends the step with the full
result, bit-identical to the single-GPU result (checked before the timed region, on every N).

--impl reference times the CPU restatement of the reference's own algorithm (oracle/ng_oracle.c: the literal O(n^2)
pair loop, all host threads) on a bounded sample of the same workload.  The reference itself is Perl, cannot travel
to the GPU box and has nothing to compile, so kind = "port" (see DESIGN.md).
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
    cpu_pair_rate.last_seconds = t           # wall time of the sampled pair loop (the reference arm's ms_per_step)
    return nominal_pairs(N_SEQS, k) / t, threads, f"{k} codons x {N_SEQS} seqs of config 2 (literal O(n^2) pair loop, {t:.1f} s)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    per_step = max(1.0, min(20.0, 90.0 / (steps + args.warmup)))     # the whole run ends within a few minutes
    vals, secs = [], []
    sample = ""
    threads = 1
    for i in range(args.warmup + steps):
        v, threads, sample = cpu_pair_rate(per_step)
        if i >= args.warmup:
            vals.append(v)
            secs.append(cpu_pair_rate.last_seconds)
    value = float(np.mean(vals))
    ms = 1e3 * float(np.mean(secs))          # a step of this arm IS the bounded sample: its measured time, not an extrapolation
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ms_per_full_step_extrapolated": 1e3 * nominal_pairs(N_SEQS, 9677) / value,
        "note": "a step of this arm is the bounded sample named in cpu_baseline.sample; ms_per_step is its measured wall time; value is the "
                "measured pair rate; ms_per_full_step_extrapolated = all 9,677 codons at that rate (not measured)",
    }))


def workload_config(n_gpus):
    return {"workload": "config2: SARS-CoV-2-shaped within-group, 10000 seqs x 29903 nt, 10 CDS (9677 codons), 10000 bootstraps"
                        + ("" if n_gpus == 1 else f"; ONE alignment, codon axis and bootstrap replicates split over {n_gpus} GPUs"),
            "n_seqs": N_SEQS, "aln_len": ALN_LEN, "codons": 9677, "bootstraps": N_BOOT,
            "l2": "inputs larger than L2 (299 MB alignment per step, re-read every step)"}


def bind_near_gpu(local):
    """Pin this rank to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host memory is allocated
    (first touch then places the e2e input next to the GPU's PCIe root).  Returns (node, n_cpus) or None."""
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

class Timer:
    """K steps bracketed by barrier + synchronize on both sides, CUDA events on the engine's stream, max over ranks."""

    def __init__(self, torch, dist, stream, world, dev):
        self.torch, self.dist, self.stream, self.world, self.dev = torch, dist, stream, world, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn, steps, warmup):
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            out = fn()
        e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        mine = ms / steps
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, mine, out


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
    stream = torch.cuda.Stream(device=dev)          # the library launches on this stream; the timing events are recorded on it
    timer = Timer(torch, dist, stream, world, dev)

    products = synth.sars2_products()
    codons = sum(p.n_codons for p in products)
    aln = synth.make_alignment(N_SEQS, ALN_LEN, products, SEED, p_poly=args.p_poly, max_changes=args.max_changes, max_alleles=args.max_alleles)      # the same alignment on every rank
    pitch = (ALN_LEN + 127) // 128 * 128
    # pinned host copy (e2e input) and HBM-resident copy with a 128-B row pitch (value input)
    h_aln = torch.empty((N_SEQS, ALN_LEN), dtype=torch.uint8).pin_memory()
    h_aln.numpy()[:] = aln
    d_aln = torch.zeros((N_SEQS, pitch), dtype=torch.uint8, device=dev)
    d_aln[:, :ALN_LEN] = h_aln.to(dev)
    torch.cuda.synchronize()

    def make_engine(r, w):
        e = Engine(device=local, seed=SEED, rank=r, world=w)
        e.set_stream(stream.cuda_stream)
        e.set_option(_abi.OPT_KEEP_CODES, 1 if args.route == "codes" else 0)
        e.set_option(_abi.OPT_FUSED_KERNEL, {"tiles": _abi.FUSED_TMA_TILES, "spans": _abi.FUSED_TMA_SPANS}.get(args.route, _abi.FUSED_LDG_SPANS))
        if args.boot_kernel >= 0:
            e.set_option(_abi.OPT_BOOTSTRAP_KERNEL, args.boot_kernel)
        e.set_products(products, ALN_LEN)
        return e

    eng = make_engine(rank, world)
    exchange = None
    if world > 1:
        eng.connect_ranks(N_BOOT)
        exchange = "peer memory over NVLink (CUDA IPC): kernels store into every rank's HBM, flags signal arrival"

    def step_resident():
        return eng.job(0, d_aln.data_ptr(), N_SEQS, ALN_LEN, pitch, device=True, B=N_BOOT)

    def step_e2e():
        return eng.job(0, h_aln.data_ptr(), N_SEQS, ALN_LEN, ALN_LEN, device=False, B=N_BOOT)

    # ---- correctness of what is about to be timed (outside the timed region) --------------------------------
    check = {}
    first = {k: np.array(v, copy=True) for k, v in step_resident().items()}
    if world > 1:
        # the same job, unsharded, on this rank's own GPU: a fresh engine's first job draws the same replicates
        with make_engine(0, 1) as solo:
            ref = {k: np.array(v, copy=True) for k, v in solo.job(0, d_aln.data_ptr(), N_SEQS, ALN_LEN, pitch, device=True, B=N_BOOT).items()}
        same = all(np.array_equal(first[k], ref[k]) for k in ref)
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        check["sharded_equals_unsharded_bitwise_on_every_rank"] = bool(flag.item())
        if not flag.item():
            raise SystemExit(f"[rank {rank}] sharded result differs from the unsharded run: " + ", ".join(k for k in ref if not np.array_equal(first[k], ref[k])))
    if rank == 0 and not args.no_cpu:
        # CHECKER ONLY: per-codon closed form of the oracle on host-built histograms of the same alignment
        from oracle import oracle as orc
        sites, diffs = orc.tables()
        ofs = np.cumsum([0] + [p.n_codons for p in products])
        worst = 0.0
        for i, p in enumerate(products):
            pos = orc.product_positions(p.starts, p.stops, p.strand, ALN_LEN)
            _, codes = orc.extract(aln, pos, p.strand)
            want = orc.within_hist(orc.hist(codes), None, sites, diffs)
            got = first["stats4"][ofs[i]:ofs[i + 1]]
            for q, k in enumerate(("N_sites", "S_sites", "N_diffs", "S_diffs")):
                err = np.max(np.abs(got[:, q] - want[k]) / np.maximum(np.abs(want[k]), 1e-300) * (want[k] != 0) + np.abs(got[:, q]) * (want[k] == 0))
                worst = max(worst, float(err))
                assert np.allclose(first["prod_sums"][i, q], want[k].sum(), rtol=1e-12), (p.name, k)
            assert np.array_equal(first["poly"][ofs[i]:ofs[i + 1]], want["poly"]) and np.array_equal(first["n_defined"][ofs[i]:ofs[i + 1]], want["n_defined"])
        assert worst < 1e-12, worst
        check["vs_oracle"] = {"codons_checked": int(codons), "max_rel_err_per_codon": worst, "tolerance": 1e-12,
                              "poly_and_n_defined": "identical", "se_dN_minus_dS_product0": float(first["prod_se"][0, 2])}
    if args.p_poly == 0.3:      # (a sparser side experiment may leave a short product without synonymous differences)
        assert np.all(first["prod_se"][:, :3] > 0), "bootstrap SEs must be positive on this workload"

    # ---- timed regions ----------------------------------------------------------------------------------------
    # timed passes bracket only the dominant HBM kernel with events (a pair costs ~3 us of stream time);
    # a separate untimed pass of the same steps brackets every launch for the per-kernel table
    hbm_kinds = ["fused"] if args.route in ("spans", "fused", "tiles") else ["pack", "hist"]
    narrow = sum(1 << _abi.KERNEL_KINDS.index(k) for k in hbm_kinds)
    eng.set_option(_abi.OPT_PROFILE, 0 if args.no_profile else narrow)
    with ClockSampler(local) as clocks:      # clocks are sampled across both timed regions
        for _ in range(args.warmup):
            step_resident()
        eng.profile(reset=True)
        l0 = eng.launches
        ms_res, ms_res_mine, _ = timer.run(step_resident, args.steps, 0)
        prof = eng.profile(reset=True)
        launches = eng.launches - l0
        ms_e2e, ms_e2e_mine, _ = timer.run(step_e2e, max(1, args.steps), max(3, min(args.warmup, 3)))
    if world > 1:
        print(f"[rank {rank}] resident {ms_res_mine:.4f} ms/step, e2e {ms_e2e_mine:.4f} ms/step", file=sys.stderr)
    prof_all = {}
    if not args.no_profile:
        eng.set_option(_abi.OPT_PROFILE, -1)
        for _ in range(2):
            step_resident()
        eng.profile(reset=True)
        timer.run(step_resident, args.steps, 0)
        prof_all = eng.profile(reset=True)
        eng.set_option(_abi.OPT_PROFILE, 0)

    first_c, count_c, _ = eng.shard_range()
    widths = None
    if world > 1:
        w = torch.tensor([eng.staged_bytes_per_row()], device=dev, dtype=torch.int64)
        allw = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(allw, w)
        widths = [int(x.item()) for x in allw]

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        extras = side_configs(torch, Engine, synth, stream, dev)
    if world == 8 and not args.no_extras:
        c5 = config5(torch, dist, Engine, synth, stream, dev, rank, world, local)
        if rank == 0:
            extras["config5"] = c5

    if rank == 0:
        pairs_total = nominal_pairs(N_SEQS, codons)
        value = pairs_total / (ms_res * 1e-3)
        e2e = pairs_total / (ms_e2e * 1e-3)
        peak, peak_src = peaks()
        # algorithmic bytes per launch, for THIS rank's codon range: pack reads 3 B and writes 1 B per
        # (sequence, codon); hist reads 1 B per (sequence, codon) and writes 264 B per codon
        alg = {"pack": 4.0 * N_SEQS * count_c, "hist": 1.0 * N_SEQS * count_c + 264.0 * count_c,
               "fused": 3.0 * N_SEQS * count_c + 264.0 * count_c}   # fused: the CDS bytes read once, histograms written
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
        # which kernel the SNPG_K_FUSED launches went to
        fused_name = "k_tile_hist" if eng.tile_launches > 0 else "k_span_hist" if eng.span_tma_launches > 0 else "k_fused_hist"
        kernel_name = {"pack": "k_pack", "hist": "k_hist", "fused": fused_name}[dom]
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get({"k_tile_hist": "tile", "k_span_hist": "span"}.get(kernel_name, dom))
            if world > 1:
                traffic = None                 # captured at N = 1
        except Exception:
            pass
        d2h = int(codons * (32 + 8 + 4 + 1 + 1) + len(products) * (4 * 8 + 6 * 8 + 6 * 4))
        h2d_total = int(N_SEQS * sum(widths)) if widths else int(N_SEQS * eng.staged_bytes_per_row())
        boot = bootstrap_block(eng, kern, products, world)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8 codon codes / u32 histograms / f64 statistics", "data": "synthetic",
            "config": dict(workload_config(world), **({} if args.p_poly == 0.3 and args.max_changes == 3 and args.max_alleles == 3 else {"p_poly": args.p_poly, "max_changes_per_allele": args.max_changes, "max_alleles_per_codon": args.max_alleles, "side_experiment": True})),
            # the library stages only the CDS columns of a rank's codon range: from its first column rounded down to 16 to its last
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e,
                    "host_buffer_bytes": int(N_SEQS * ALN_LEN), "h2d_bytes_per_rank": [int(N_SEQS * w) for w in widths] if widths else None,
                    "outputs": "per-codon table (poly, n_defined, comparisons, stats4, conserved) + product sums, SEs, undefined counts, on every rank"},
            "gpu_launches": int(launches),
            "route": args.route,
            "exchange": exchange,
            "check": check,
            "roofline": {"kernel": kernel_name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg[dom], "rank": 0, "codons_of_rank": int(count_c)},
            "kernels": kern,
            "kernel_timing": "CUDA events inside the library on the launch stream (rank 0): " + "/".join(hbm_kinds)
                             + " during the timed steps; the other kernels in a separate pass of the same steps",
            "bootstrap": boot,
            "clocks": clocks.summary(),
            "host_binding": {"numa_node": numa[0], "cpus": numa[1]} if numa else None,
        }
        out.update(extras)
        if world == 1 and not args.no_cpu:
            v, threads, sample = cpu_pair_rate(12.0)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bootstrap_block(eng, kern, products, world):
    """The resampling kernel against what bounds it: Philox4x32-10 generation (integer multiply issue) and, for the
    contraction, the FP64 tensor pipe -- both peaks measured on this GPU by the library's own probes."""
    if "bootstrap" not in kern:
        return None
    ms = kern["bootstrap"]["ms_per_step"]
    draws = N_BOOT * sum(p.n_codons for p in products) / world                       # this rank's replicate range
    flops = 2.0 * 4 * N_BOOT * sum(p.n_codons for p in products) / world             # algorithmic: B x C x 4 multiply-adds
    boot_ms = sum(kern[k]["ms_per_step"] for k in ("bootstrap", "moments") if k in kern)
    blk = {"value": N_BOOT * len(products) / (boot_ms * 1e-3), "unit": "bootstrap reps/s",
           "note": "B x products per step / (bootstrap + moments kernel time), whole job; draws are NOMINAL (C per replicate): the default "
                   "class-aggregated kernel gets the draws on conserved codons of one class from binomial variates and makes only the "
                   "others one by one, so it may pass the one-Philox-number-per-draw ceiling below",
           "kernel_ms": ms, "draws_per_launch": draws, "gdraws_per_s": draws / (ms * 1e-3) / 1e9,
           "algorithmic_dmma_tflops": flops / (ms * 1e-3) / 1e12}
    try:
        pk = eng.probe_peaks()
        blk["roofline"] = {"bound": "issue (Philox4x32-10 generation, one number per nominal draw: 20 IMAD.WIDE + 20 LOP3 per 4 draws)", "achieved": draws / (ms * 1e-3) / 1e9,
                           "peak": pk["philox_gdraws_per_s"], "unit": "Gdraws/s", "frac": draws / (ms * 1e-3) / 1e9 / pk["philox_gdraws_per_s"],
                           "issue_floor_us": draws / pk["philox_gdraws_per_s"] / 1e9 * 1e6,
                           "peak_source": "measured here: a kernel that only generates Philox4x32-10 blocks (snpg_probe_peaks)"}
        blk["fp64_tensor"] = {"peak_tflops": pk["dmma_tflops"], "peak_source": "measured here: dependent-free mma.sync.m8n8k4.f64 loop (snpg_probe_peaks)",
                              "algorithmic_frac": blk["algorithmic_dmma_tflops"] / pk["dmma_tflops"]}
    except Exception as e:       # noqa: BLE001
        blk["roofline"] = {"error": str(e)}
    return blk


def quick_time(torch, stream, fn, steps=20, warmup=4):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        out = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def side_configs(torch, Engine, synth, stream, dev):
    """BASELINE configs 3 and 4 through the same library, a few steps each (extra keys, not the bench line)."""
    res = {}
    # config 3: HIV-shaped, overlapping genes, '-' strand, 5 % damaged bases, windows of 100 codons step 10
    products = synth.hiv_products()
    L, n = 9719, 5000
    aln = synth.make_alignment(n, L, products, 20261019 + 3, gap_frac=0.025, n_frac=0.02, iupac_frac=0.005)
    pitch = (L + 127) // 128 * 128
    d = torch.zeros((n, pitch), dtype=torch.uint8, device=dev)
    d[:, :L] = torch.from_numpy(aln).to(dev)
    codons = sum(p.n_codons for p in products)
    with Engine(device=dev.index, seed=3) as e:
        e.set_stream(stream.cuda_stream)
        e.set_products(products, L)
        ms, out = quick_time(torch, stream, lambda: e.job(0, d.data_ptr(), n, L, pitch, device=True, B=1000, window=100, step=10, B_win=1000))
        nwin = int(out["win_ofs"][-1])
        res["config3"] = {"workload": "config3: within-group, 5000 seqs x 9719 nt, 9 CDS (overlapping, 2-segment, 2 on '-'), 5% damaged bases, "
                                      "1000 bootstraps, windows of 100 codons step 10 each bootstrapped 1000 times; ONE call, alignment resident in HBM",
                          "ms_per_step": ms, "windows": nwin, "window_bootstrap_reps_per_s": nwin * 1000 / (ms * 1e-3),
                          "codon_pair_comparisons_per_s": nominal_pairs(n, codons) / (ms * 1e-3), "graph_replays": e.graph_replays}
    del d
    # config 4: between-group, 4 groups x 5,000 x 29,903, 6 group pairs, 1,000 bootstraps
    products = synth.sars2_products()
    L, n, G = 29903, 5000, 4
    pitch = (L + 127) // 128 * 128
    codons = sum(p.n_codons for p in products)
    bufs = []
    for g in range(G):
        a = synth.make_alignment(n, L, products, 20261019 + 4 + 100 * g)
        t = torch.zeros((n, pitch), dtype=torch.uint8, device=dev)
        t[:, :L] = torch.from_numpy(a).to(dev)
        bufs.append(t)
    with Engine(device=dev.index, seed=4) as e:
        e.set_stream(stream.cuda_stream)
        e.set_products(products, L)

        def step4():
            return e.between_job([b.data_ptr() for b in bufs], [n] * G, L, [pitch] * G, device=True, B=1000)
        ms, _ = quick_time(torch, stream, step4, steps=10, warmup=3)
        res["config4"] = {"workload": "config4: between-group, 4 groups x 5000 seqs x 29903 nt, 6 group pairs x 10 CDS, 1000 bootstraps; "
                                      "ONE snpg_between_job call (4 loads + every pair), alignments resident in HBM",
                          "ms_per_step": ms, "codon_pair_comparisons_per_s": 6 * codons * n * n / (ms * 1e-3),
                          "bootstrap_reps_per_s": 6 * len(products) * 1000 / (ms * 1e-3)}
    return res


def config5(torch, dist, Engine, synth, stream, dev, rank, world, local):
    """BASELINE config 5: ONE alignment of 1,000,000 x 29,903, 10,000 bootstraps, codon axis and replicates over the 8
    GPUs, through the same job call.  The alignment is generated on the device: a rank fills only the columns of its
    own codon range of a full-pitch buffer (the only ones its kernels read)."""
    from snpgenie_b200 import multigpu
    products = synth.sars2_products()
    L, n = 29903, 1000000
    pitch = (L + 127) // 128 * 128
    codons = sum(p.n_codons for p in products)
    cols = np.concatenate([np.arange(p.starts[0] - 1, p.stops[0]).reshape(-1, 3) for p in products])
    first, count = multigpu.shard_plan(codons, rank, world)
    mine = cols[first:first + count]
    gen = torch.Generator(device=dev)
    gen.manual_seed(20261019 + 5)                      # same ancestral row on every rank
    anc = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (pitch,), device=dev, generator=gen)]
    d = anc.repeat(n, 1)
    rng = np.random.default_rng(1000 + rank)
    nt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for c in np.flatnonzero(rng.random(count) < 0.3):
        u = torch.rand(n, device=dev)
        acc = 0.0
        for f in rng.beta(0.5, 5.0, size=int(rng.integers(1, 4))):
            col = int(mine[c, rng.integers(0, 3)])
            alt = int(rng.choice(nt[nt != int(anc[col])]))
            d[:, col].masked_fill_((u >= acc) & (u < acc + f), alt)
            acc += f
    torch.cuda.synchronize()
    dist.barrier()
    with Engine(device=local, seed=5, rank=rank, world=world) as e:
        e.set_stream(stream.cuda_stream)
        e.set_products(products, L)
        e.connect_ranks(N_BOOT)
        timer = Timer(torch, dist, stream, world, dev)
        # the FIRST job of this engine is the one compared with the first job of an unsharded engine below (the bootstrap's
        # Philox stream carries the job number)
        out = {k: np.array(v, copy=True) for k, v in e.job(0, d.data_ptr(), n, L, pitch, device=True, B=N_BOOT).items()}
        ms, _, _ = timer.run(lambda: e.job(0, d.data_ptr(), n, L, pitch, device=True, B=N_BOOT), 10, 3)
        se = float(out["prod_se"][0, 2])
        s0 = [float(x) for x in out["prod_sums"][0]]
        # consistency across ranks: every rank holds the same full result
        t = torch.tensor(out["prod_se"].reshape(-1), device=dev)
        hi, lo = t.clone(), t.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        same = bool(torch.equal(hi, lo))
    # the same analysis unsharded on rank 0: every rank's columns onto rank 0's buffer, one engine, one GPU
    for r in range(world):
        f, cnt = multigpu.shard_plan(codons, r, world)
        c_lo, c_hi = int(cols[f:f + cnt].min()), int(cols[f:f + cnt].max()) + 1
        for r0 in range(0, n, 250000):           # in row blocks: a contiguous copy of a column slice is a few GB
            part = d[r0:r0 + 250000, c_lo:c_hi].contiguous() if r == rank else torch.empty((min(250000, n - r0), c_hi - c_lo), dtype=torch.uint8, device=dev)
            dist.broadcast(part, src=r)
            if rank == 0 and r != 0:
                d[r0:r0 + 250000, c_lo:c_hi] = part
            del part
    matches = None
    if rank == 0:
        with Engine(device=local, seed=5) as full:
            full.set_stream(stream.cuda_stream)
            full.set_products(products, L)
            ref = full.job(0, d.data_ptr(), n, L, pitch, device=True, B=N_BOOT)
            matches = {k: bool(np.array_equal(out[k], ref[k])) for k in ("prod_sums", "prod_se", "stats4", "poly", "n_defined")}
    del d
    return {"workload": f"config5: within-group, {n} seqs x {L} nt, 10 CDS, {N_BOOT} bootstraps, ONE alignment split over {world} GPUs; resident in HBM",
            "ms_per_step": ms, "codon_pair_comparisons_per_s": nominal_pairs(n, codons) / (ms * 1e-3), "bytes_read_per_gpu": int(n * 3 * count),
            "hbm_floor_ms": n * 3 * count / (peaks()[0] * 1e9) * 1e3, "dN_minus_dS_SE_product0": se, "product0_sums": s0,
            "all_ranks_hold_identical_results": same, "equals_unsharded_run_bitwise": matches}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the oracle check")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 3/4/5 side measurements")
    ap.add_argument("--boot-kernel", type=int, default=-1, choices=[-1, 0, 1, 2],
                    help="side experiment: SNPG_OPT_BOOTSTRAP_KERNEL (0 gather, 1 weights, 2 classes; default: the library's default, classes)")
    ap.add_argument("--max-alleles", type=int, default=3, help="side experiment: minor alleles per polymorphic codon, at most (default 3)")
    ap.add_argument("--max-changes", type=int, default=3,
                    help="side experiment: nucleotides an allele changes in its codon, at most (default 3: one, two or three with equal chance)")
    ap.add_argument("--p-poly", type=float, default=0.3,
                    help="fraction of polymorphic codon columns in the synthetic alignment (SURVEY.md 8d: 0.3; "
                         "other values are side experiments, never the bench line)")
    ap.add_argument("--no-profile", action="store_true", help="do not bracket kernels with CUDA events (no roofline block)")
    ap.add_argument("--route", default="fused", choices=["fused", "spans", "tiles", "codes"],
                    help="histogram route: vector-load span kernel (default), TMA-fed span kernel, TMA-tiled kernel, or K1 pack + "
                         "K2 hist over a kept code matrix")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

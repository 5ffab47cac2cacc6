#!/usr/bin/env python
"""bench.py -- hmmsearch GCUPS / sequences-per-second on B200 (BASELINE.json's metric and configs[1]).

    python bench.py --gpus 1 --steps 5 --warmup 3                       # this repo's CUDA engine
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1       # the reference's CPU pipeline, all host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     # one rank per GPU, target DB sharded

Workload (N=1): 100 Pfam-like synthetic profile HMMs (M ~ lognormal, mean ~200, [30,800]) calibrated on the
GPU, against 50 000 synthetic proteins (iid background residues, L ~ N(350,100) in [50,1500]); 1 % of the
targets carry a domain emitted from one of the profiles so that the survivor tail (Forward/Backward, domain
definition, alignment) is exercised.  With N GPUs every rank gets its own 50 000-sequence shard of a
N x 50 000 database (weak scaling, target-sharded as the reference's parallel="targets" mode).

A "step" is one complete hmmsearch of the 100 queries against the database: the whole p7_Pipeline cascade
(SSV/MSV, bias, Viterbi, Forward, Backward, domain definition, hit assembly).  Nothing is skipped.
`value` is measured with the database and the profiles already resident in HBM (and includes building the thresholded
`TopHits` of every query); `e2e` repeats the step
through the public Python API with host buffers (packing, H2D upload of sequences and profile tables, D2H
of the results all inside the timed region).  Both consume the search the way `Pipeline.search_hmm` / `hmmer.hmmsearch` do:
wave by wave (b2h_search_begin / _next / _end) -- the hit records of a finished wave are exchanged between the ranks (N > 1)
and turned into `TopHits` while the GPU searches the following waves; the step ends when the last `TopHits` exists.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# one hardware work queue per stream (default 8): the engine keeps ~30 streams busy (size classes, the high-priority
# survivor lane, envelope classes) and queue aliasing would serialise the high-priority kernels behind queued cascade
# launches.  Must be set before the CUDA context exists, i.e. before torch touches the device.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

N_PROFILES = 100
N_SEQS = 50000
PLANT_FRAC = 0.01
# one string for both arms (the driver compares the two lines' config.workload)
WORKLOAD = ("hmmsearch: 100 Pfam-like profiles (M~200) vs 50k synthetic proteins on 1xB200 (BASELINE configs[1]); "
            "per-rank shard of 50k targets at N>1")


def build_inputs(rank, world, n_profiles=N_PROFILES, n_seqs=N_SEQS):
    """Seeded synthetic inputs (bench_inputs.c2_inputs) as package objects.  Profiles are identical on every rank; each
    rank gets its own target shard."""
    import bench_inputs
    from pyhmmer_b200 import easel
    abc = easel.Alphabet.amino()
    models, seqs, _ = bench_inputs.c2_inputs(rank, n_profiles, n_seqs)
    hmms = bench_inputs.to_hmms(models, abc)
    return abc, hmms, bench_inputs.to_block(seqs, abc, prefix="r%d_" % rank)


STATS_FILE = os.path.join(ROOT, "tests", "golden", "bench_stats.json")


def apply_stats(hmms):
    """Committed E-value statistics of the seeded synthetic profiles (fitted once on a B200 by synth.calibrate and
    kept under tests/golden/ so that this arm and `--impl reference` score identical models in any order)."""
    try:
        st = json.load(open(STATS_FILE))
    except Exception:
        return False
    if len(st.get("evparam", [])) < len(hmms):
        return False
    for h, ev in zip(hmms, st["evparam"]):
        h._evparam[:] = np.array(ev, dtype=np.float32)
    return True


def write_hmm_file(hmms, path):
    with open(path, "wb") as f:
        for h in hmms:
            h.write(f)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def time_reference(models, seqs, sample_p, sample_s, warmup, steps):
    """Time oracle/_ref (HMMER 3.4 built from the reference's sources) on the host cores: every step is the whole
    p7_Pipeline over sample_p profiles x sample_s sequences, work-queue threaded.  Both thread counts (physical
    cores, logical CPUs) are tried and the faster one is reported, so the reference gets its best configuration.
    <models> / <seqs> are bench_inputs arrays: nothing of pyhmmer_b200 is touched on this path."""
    import bench_inputs
    from oracle import refshim
    import psutil
    logical = psutil.cpu_count(logical=True) or os.cpu_count() or 1
    physical = psutil.cpu_count(logical=False) or logical
    codes = seqs[:sample_s]
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "q.hmm")
        bench_inputs.write_hmm_file(models[:sample_p], path)       # the same HMMER3 ASCII text the product arm reads
        refs = [refshim.RefModel(path, i, 400) for i in range(sample_p)]
        cells = float(sum(m["M"] for m in models[:sample_p])) * float(sum(len(s) for s in codes))
        best = None
        for ncore in sorted({max(1, min(physical, 256)), max(1, min(logical, 256))}):
            times = []
            for it in range(warmup + steps):
                t0 = time.perf_counter()
                nh, ctr = refshim.search_mt(refs, codes, ncore)
                dt = time.perf_counter() - t0
                if it >= warmup:
                    times.append(dt)
            tot = sum(times)
            if best is None or tot < best["tot"]:
                best = {"tot": tot, "n": len(times), "cores": ncore, "hits": int(nh), "counters": ctr}
    best["gcups"] = cells * best["n"] / best["tot"] / 1e9
    best["cells"] = cells
    return best


def reference_arm(args, rank, world_size):
    """The reference's own CPU implementation of the path (oracle/_ref = HMMER 3.4 compiled from the reference's
    sources), all host threads, on the same workload (all 100 profiles x 50k sequences per step by default).  The inputs
    come from bench_inputs (numpy only): this arm never imports pyhmmer_b200 and never loads libb2h.so."""
    if rank != 0:
        return
    import bench_inputs
    models, seqs, calibrated = bench_inputs.c2_inputs(0)
    sample_p, sample_s = args.ref_profiles, min(args.ref_seqs, len(seqs))
    r = time_reference(models, seqs, sample_p, sample_s, args.warmup, args.steps)
    gcups, tot, n = r["gcups"], r["tot"], r["n"]
    sample = "%d profiles x %d of the %d sequences per step (same generator, seed, planted homologs), %.2f s/step on %d threads; %s" % (
        sample_p, sample_s, N_SEQS, tot / n, r["cores"], "committed GPU-fitted statistics" if calibrated else "placeholder statistics")
    line = {
        "impl": "reference", "metric": "hmmsearch GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": n, "warmup": args.warmup, "ms_per_step": 1e3 * tot / n, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "seqs_per_s": sample_s * sample_p * n / tot,
        "config": {"workload": WORKLOAD, "arm": "reference CPU pipeline (HMMER 3.4 SSE2 from the reference's sources, work-queue threads)",
                   "profiles": sample_p, "sequences": sample_s, "hits": r["hits"], "pipeline_counters": r["counters"]},
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": r["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2h", choices=["b2h", "reference"])
    ap.add_argument("--ref-profiles", type=int, default=100)
    ap.add_argument("--ref-seqs", type=int, default=50000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config.extra blocks (BASELINE configs[2..4])")
    ap.add_argument("--extras", default="c4,c5,c3", help="which extra configurations to run (comma separated: c3, c4, c5)")
    ap.add_argument("--pfam-n", type=int, default=20000, help="profiles of the Pfam-A-sized set (configs[2], [3])")
    ap.add_argument("--c4-len", type=int, default=5000)
    ap.add_argument("--c5-mb", type=float, default=100.0)
    ap.add_argument("--c5-ref-mb", type=float, default=16.0)
    ap.add_argument("--c3-seqs", type=int, default=100000)
    ap.add_argument("--c3-ref-profiles", type=int, default=200)
    ap.add_argument("--extra-steps", type=int, default=4)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b2h":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")        # NCCL's version / debug lines: not on the JSON line's stream
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pyhmmer_b200 import _lib, plan7, synth, parallel

    ctx = _lib.context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    abc, hmms, seqs = build_inputs(rank, world)
    if not apply_stats(hmms):
        synth.calibrate(hmms, ctx)                                # GPU calibration (same seed => same numbers on every rank)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        if rank == 0:
            json.dump({"evparam": [[float(v) for v in h._evparam] for h in hmms]},
                      open(os.path.join(ROOT, "gpurun_out", "bench_stats.json"), "w"))
    # round-trip the models through the ASCII format so that both arms score the very same numbers
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "q.hmm")
        write_hmm_file(hmms, p)
        with plan7.HMMFile(p) as f:
            hmms = list(f)
    pli = plan7.Pipeline(abc)
    oms = [pli._optimized(h, len(seqs[0])) for h in hmms]
    db = plan7.SequenceDatabase.of(ctx, seqs)                      # resident in HBM
    for om in oms:
        om._device(ctx)
    cells_local = float(sum(h.M for h in hmms)) * float(seqs.total_residues)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wave_split = []

    def search_step():
        """One pass of the path: the search wave by wave (b2h_search_begin / _next / _end); while the GPU searches the
        following waves, the finished wave's hit records are exchanged (N > 1: one all-gather per wave, the only collective
        of the path) and its thresholded `TopHits` are built -- one per query, the job's output (the CPU arm's p7_Pipeline
        builds its P7_TOPHITS inside its timed region, too).  Uploads anything that is not resident."""
        w = parallel.World.current()
        res = [None] * len(oms)
        nlocal, gen = pli._run_waves(oms, seqs)               # (what Pipeline._search_many / parallel.search_sharded do)
        first, nrec = True, 0
        # parallel.gathered_waves: one all-gather of the packed hit records per wave (N > 1), the number of rounds agreed on the way
        rounds = parallel.gathered_waves(nlocal, gen, w, 0, len(oms)) if world > 1 else ((wave, None) for wave in gen)
        for (profs, hits, doms, text, counters), _gathered in rounds:
            # (weak scaling: every rank holds its own shard only, so it assembles its own records)
            for qi, th in zip(profs, pli._assemble(oms, oms, seqs, hits, doms, text, counters, only=profs, count_targets=first)):
                res[qi] = th
            first = False
            nrec += len(hits)
        wave_split.append(pli._last_run_s)
        return res, nrec

    def one_step_resident():
        res, nrec = search_step()
        counters = np.array([[th.n_past_msv, th.n_past_bias, th.n_past_vit, th.n_past_fwd] for th in res], np.int64)
        return nrec, counters

    e2e_phases = []

    def one_step_e2e():
        seqs._cache = {k: v for k, v in seqs._cache.items() if k in ("packed", "nres")}   # host buffers stay; drop the device copy: re-upload
        for om in oms:
            om._dev = {}                                           # re-upload the profile tables
        t0 = time.perf_counter()
        # the call a user's hmmsearch makes: packs + uploads the database and builds + uploads every profile's tables (side
        # by side), then the search with the D2H of the hit records and the assembly of the TopHits, wave by wave
        res, _ = search_step()
        t4 = time.perf_counter()
        up, waited, read = pli._last_run_s
        e2e_phases.append((up, waited, read, (t4 - t0) - up - waited - read))
        return res

    # The inputs are ~10^5 long-lived Python objects (50 000 sequences, their arrays, the profiles).  A full collection that
    # walks them costs tens of milliseconds and used to land inside one timed step (round 1: step 12 of 20 took 2x): after
    # this point they are frozen out of the collector's generations, so the collections a step's own short-lived result
    # objects trigger stay cheap.
    import gc
    gc.collect()
    gc.freeze()
    sampler = ClockSampler(local)
    times, stage_acc = [], {}
    host_split = []
    launches0 = 0
    hits = counters = None
    for it in range(args.warmup + args.steps):
        flush.fill_(it & 0xff)
        barrier()
        if it == args.warmup:
            sampler.start()
            ctx.set_profiling(True)
            ctx.stage_ms(reset=True)
            launches0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        tw0 = time.perf_counter()
        hits, counters = one_step_resident()
        tw1 = time.perf_counter()
        e1.record(stream)
        barrier()
        if it >= args.warmup:
            times.append(e0.elapsed_time(e1))
            host_split.append((tw1 - tw0,) + tuple(pli._last_run_s))
    clocks = sampler.stop()
    stage_ms = ctx.stage_ms(reset=True)
    ctx.set_profiling(False)
    launches = ctx.launch_count - launches0
    tot_ms = sum(times)
    t = torch.tensor([tot_ms], dtype=torch.float64, device="cuda")
    c = torch.tensor([cells_local, float(len(seqs)) * len(hmms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    tot_ms = float(t.item())
    cells_all, comps_all = float(c[0].item()), float(c[1].item())
    K = len(times)
    gcups = cells_all * K / (tot_ms * 1e-3) / 1e9

    # end-to-end through the public API with host buffers
    e2e_times = []
    for it in range(2 + max(3, K)):
        res = None
        gc.collect()                                              # (untimed) the previous step's result objects: keep collector pauses out of the next step
        flush.fill_(it & 0xff)
        barrier()
        t0 = time.perf_counter()
        res = one_step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= 2:
            e2e_times.append(dt)
    if rank == 0:
        hs = np.array(host_split) * 1e3
        print("[bench] resident step, host wall ms (whole step, handles, waiting for waves, reading results; the rest = exchange + TopHits assembly, overlapped with the following waves): mean %s; device-event ms %s"
              % (np.round(hs.mean(0), 2).tolist(), np.round(times, 1).tolist()), file=sys.stderr)
        ph = np.array(e2e_phases[2:]) * 1e3
        print("[bench] e2e phases ms (database + profile uploads, waiting for waves, reading results, exchange + TopHits assembly): mean %s, per step %s"
              % (np.round(ph.mean(0), 1).tolist(), np.round(ph.sum(1), 1).tolist()), file=sys.stderr)
    te = torch.tensor([sum(e2e_times)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_gcups = cells_all * len(e2e_times) / float(te.item()) / 1e9
    h2d = int(_lib.lib.b2h_seqdb_h2d_bytes(plan7.SequenceDatabase.of(ctx, seqs).handle)) + int(sum(_lib.lib.b2h_profile_h2d_bytes(om._device(ctx)) for om in oms))
    nhits_rank0 = int(hits)                                   # comparisons scored to completion (hit records) of the last step
    nh = sum(len(r) for r in res)
    d2h = int(nh * 96 + sum(len(h.domains) * 88 + sum(4 * (len(d.alignment) + 1) for d in h.domains) for r in res for h in r))

    extra = None
    if not args.no_extras:
        import bench_extras
        del res, hits
        gc.collect()
        try:
            extra = bench_extras.run(args, ctx, rank, world, torch, dist, lambda m: print(m, file=sys.stderr, flush=True) if rank == 0 else None)
        except Exception as exc:                          # the headline stands on its own; say what happened to the rest
            import traceback
            traceback.print_exc()
            extra = {"failed": repr(exc)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # roofline of the dominant kernel (SSV): algorithmic bytes of SURVEY 8(d) / its measured device time
    pk, pk_kind = peaks()
    msv_tab_bytes = sum(29 * h.M for h in hmms)
    n_past_msv = int(counters[:, 0].sum()) if counters is not None else 0
    b_alg = float(seqs.total_residues + 2 * len(seqs)) + msv_tab_bytes + 16.0 * n_past_msv      # per step, this rank
    ssv_ms = stage_ms["ssv"] / K
    achieved = b_alg / (ssv_ms * 1e-3) / 1e9 if ssv_ms > 0 else None
    # the resource the SSV kernel actually saturates: the shared-memory data pipe (128 B/clk/SM).  Per DP row a WARP
    # (32/G comparisons side by side) moves b2h_ssv_tile_info's wavefronts: emission scores (LDS) + 1.25 shuffles.
    import ctypes

    def ssv_row_bytes(M):
        G, NR, wf = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        _lib.check(_lib.lib.b2h_ssv_tile_info(int(M), ctypes.byref(G), ctypes.byref(NR), ctypes.byref(wf)), "b2h_ssv_tile_info")
        return wf.value * 128.0 * G.value / 32.0, 2.0 * G.value * NR.value      # bytes per row per comparison, cells per row incl. padding

    rows = float(sum(((len(q) + 3) // 4) * 4 for q in seqs))
    smem_bytes = sum(rows * ssv_row_bytes(h.M)[0] for h in hmms)
    padded_cells = sum(rows * ssv_row_bytes(h.M)[1] for h in hmms)
    sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
    smem_peak = 128.0 * torch.cuda.get_device_properties(local).multi_processor_count * sm_clock / 1e9
    smem_ach = smem_bytes / (ssv_ms * 1e-3) / 1e9 if ssv_ms > 0 else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ssv_traffic.json")))["dram_bytes_per_step"]
    except Exception:
        pass
    line = {
        "metric": "hmmsearch GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": tot_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "seqs_per_s": comps_all * K / (tot_ms * 1e-3),
        "config": {"workload": WORKLOAD, "arm": "pyhmmer_b200 CUDA engine",
                   "profiles": len(hmms), "sequences_per_rank": len(seqs), "sum_M": int(sum(h.M for h in hmms)),
                   "residues_per_rank": int(seqs.total_residues), "planted_homolog_fraction": PLANT_FRAC,
                   "l2": "flushed (256 MiB write) between steps", "sharding": "targets by rank, one all-gather of hit records",
                   "hits_rank0": nhits_rank0, "pipeline_counters_rank0": counters.sum(0).tolist(),
                   "stage_ms_per_step": {k: v / K for k, v in stage_ms.items()},
                   "ssv_kernel_gcups": (cells_local / (ssv_ms * 1e-3) / 1e9) if ssv_ms > 0 else None,
                   "ssv_kernel_gcups_incl_padding": (padded_cells / (ssv_ms * 1e-3) / 1e9) if ssv_ms > 0 else None},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                     "frac": (achieved / pk["hbm_gbs"]) if achieved else None, "traffic": traffic, "peak_kind": pk_kind,
                     "kernel": "ssv_kernel<G,NR> (all launches of a step)", "algorithmic_bytes_per_step": b_alg,
                     "on_chip": {"bound": "shared-memory data pipe (LDS + SHFL wavefronts)", "achieved": smem_ach, "peak": smem_peak,
                                 "unit": "GB/s", "frac": (smem_ach / smem_peak) if smem_ach else None,
                                 "bytes_per_step": smem_bytes},
                     "note": "DP filter with ~6e-5 compulsory HBM bytes per cell (SURVEY 8d): the HBM fraction is tiny by construction. "
                             "The kernel is bound by the shared-memory pipe: 2 B of emission scores per cell; see on_chip and profiles/"},
        "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * float(te.item()) / len(e2e_times),
                "note": "host inputs per step: the block's residue array and the converted OptimizedProfiles (both arms keep their models "
                        "configured); timed: arena packing into page-locked staging, device tables built from the profiles, both uploads, the "
                        "search wave by wave, D2H of the hit records, TopHits assembly"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            sp, ss = 100, len(seqs)
            import bench_inputs
            r = time_reference(bench_inputs.c2_inputs(rank)[0], [q.sequence for q in seqs], sp, ss, 1, 3)
            line["cpu_baseline"] = {"value": r["gcups"], "unit": "GCUPS", "cores": r["cores"], "kind": "reference",
                                    "sample": "%d profiles x all %d sequences (the whole step), mean of 3 passes after 1 warm-up, %.2f s/pass on %d threads (best of physical/logical core counts); oracle/_ref = HMMER 3.4 SSE2 built from the reference sources" % (sp, ss, r["tot"] / r["n"], r["cores"]),
                                    "pipeline_counters": r["counters"], "hits": r["hits"]}
        except Exception as exc:                      # the checker is optional for the measurement itself
            line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": 0, "kind": "reference", "sample": "unavailable: %r" % (exc,)}
    if extra:
        line["config"]["extra"] = extra
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

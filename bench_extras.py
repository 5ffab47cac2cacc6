"""The other BASELINE.json configurations on bench.py's JSON line (`config.extra`), each with its own reference timing.

    c3_hmmsearch_20k_x_100k   configs[2]: 20 000 Pfam-A-sized profiles x 100 000 proteins, the target database sharded by
                              residues over the ranks (STRONG scaling: the job is the same at every N)
    c4_hmmscan_20k            configs[3]: one 5 000-residue query x the same 20 000 profiles (profile block sharded by nodes)
    c5_nhmmer_100mb           configs[4]: DNA profile (M = 1000) x 100 Mb genome, both strands (windows sharded by residues)

Every block: GCUPS through the public entry point with host buffers for the query side, profile / target databases
resident where the reference's own API keeps them resident (OptimizedProfileBlock pre-fetch, DigitalSequenceBlock), device
events on the launching stream, max over ranks; at N = 1 the reference's C pipeline (oracle/_ref) is timed beside it on a
stated sample and the pass counters / hit counts of both arms are compared.
"""
import os
import sys
import time

import numpy as np


def _txt(v):
    return v.decode() if isinstance(v, bytes) else str(v)


def _cores():
    import psutil
    return psutil.cpu_count(logical=True) or os.cpu_count() or 1


class _Timer:
    """CUDA events on torch's current stream (= the stream the engine launches on) around a host call, barrier +
    synchronize on both sides, max over ranks."""

    def __init__(self, torch, dist, world):
        self.torch, self.dist, self.world = torch, dist, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn, steps, warmup):
        torch = self.torch
        stream = torch.cuda.current_stream()
        out, ms, wall = None, [], []
        for it in range(warmup + steps):
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            out = fn()
            e1.record(stream)
            self.barrier()
            if it >= warmup:
                ms.append(e0.elapsed_time(e1))
                wall.append((time.perf_counter() - t0) * 1e3)
        t = torch.tensor([sum(ms), sum(wall)], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return out, float(t[0].item()) / len(ms), float(t[1].item()) / len(ms), ms


def run(args, ctx, rank, world, torch, dist, log):
    """Returns the dict for config["extra"] (identical on every rank; rank 0 prints it)."""
    import bench_inputs
    from pyhmmer_b200 import easel, plan7, hmmer, parallel
    which = [w for w in (args.extras or "c4,c5,c3").split(",") if w]
    tm = _Timer(torch, dist, world)
    extra = {}
    amino = easel.Alphabet.amino()
    models = oms = None

    def profile_set():
        nonlocal models, oms
        if models is None:
            t0 = time.perf_counter()
            models, cal = bench_inputs.pfam_like_models(args.pfam_n)
            hmms = bench_inputs.to_hmms(models, amino)
            t1 = time.perf_counter()
            pli = plan7.Pipeline(amino)
            oms = pli._optimized_many(hmms, 100)
            t2 = time.perf_counter()
            log("[extras] %d Pfam-like profiles (sum M = %d, %s statistics): generated in %.1f s, configured + converted in %.1f s"
                % (len(models), sum(m["M"] for m in models), "GPU-fitted" if cal else "PLACEHOLDER", t1 - t0, t2 - t1))
            profile_set.calibrated = cal
        return models, oms

    # ------------------------------------------------------------------ C4: hmmscan, one long query x 20k profiles
    if "c4" in which:
        models, oms = profile_set()
        sumM = float(sum(m["M"] for m in models))
        q = easel.DigitalSequence(amino, name=b"query5k", sequence=bench_inputs.c4_query(models, args.c4_len))
        block = plan7.OptimizedProfileBlock(amino, oms)
        w = parallel.World.current()
        pli = plan7.Pipeline(amino)

        def scan():
            return pli._scan_many([q], block, world=w)[0]       # what hmmer.hmmscan does per batch: query packed + uploaded, TopHits built
        th, ms, wall, per = tm.run(scan, steps=args.extra_steps, warmup=2)
        cells = sumM * len(q)
        blk = {"workload": "hmmscan: 1 query (L=%d) vs %d profiles (sum M = %d), profile block resident%s" %
                           (len(q), len(models), int(sumM), "" if world == 1 else ", sharded by nodes over %d ranks" % world),
               "value": cells / (ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms, "host_wall_ms_per_step": wall, "steps": len(per),
               "hits": len(th), "hits_reported": len(th.reported), "pipeline_counters": [th.n_past_msv, th.n_past_bias, th.n_past_vit, th.n_past_fwd],
               "calibrated_statistics": bool(profile_set.calibrated)}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import refshim
            nc = _cores()
            t0 = time.perf_counter()
            refs = bench_inputs.to_ref_models(models, nthreads=nc)
            t1 = time.perf_counter()
            best = None
            for _ in range(3):
                t2 = time.perf_counter()
                nh, ctr = refshim.scan_mt(refs, q.sequence, nc)
                dt = time.perf_counter() - t2
                best = dt if best is None else min(best, dt)
            t2 = time.perf_counter()
            rh, rd, rt, rc = refshim.scan(refs, q.sequence)
            one = time.perf_counter() - t2
            blk["cpu_baseline"] = {"value": cells / best / 1e9, "unit": "GCUPS", "cores": nc, "kind": "reference", "ms_per_step": best * 1e3,
                                   "single_thread": {"value": cells / one / 1e9, "ms_per_step": one * 1e3,
                                                     "note": "what pyhmmer.hmmscan gives ONE query: one thread per query (_hmmscan.py:29-37)"},
                                   "sample": "the whole job: all %d profiles (built from the same arrays, %.1f s on %d threads), best of 3 passes; "
                                             "models spread over the threads in blocks of 16 (ref_scan_mt)" % (len(models), t1 - t0, nc),
                                   "pipeline_counters": rc, "hits": len(rh)}
            blk["matches_reference"] = bool(list(rc) == blk["pipeline_counters"] and len(rh) >= len(th) and
                                            sorted(_txt(h.name) for h in th) == sorted(models[x.seq]["name"] for x in rh
                                                                                 if np.exp(x.lnP) * (x.seq + 1) <= 10.0))
            del refs
        extra["c4_hmmscan_20k"] = blk
        log("[extras] c4: %s" % {k: v for k, v in blk.items() if k != "cpu_baseline"})

    # ------------------------------------------------------------------ C5: nhmmer, DNA profile x 100 Mb genome
    if "c5" in which:
        dna = easel.Alphabet.dna()
        from pyhmmer_b200 import synth
        model, genome, nplant, cal = bench_inputs.c5_inputs(1000, args.c5_mb)
        hmm = synth.hmm_from_arrays(dna, model)
        if hmm.max_length is None or hmm.max_length <= 0:
            hmm.max_length = hmm.compute_max_length()
        block = easel.DigitalSequenceBlock(dna, [easel.DigitalSequence(dna, name=b"genome", sequence=genome)])
        pli = plan7.LongTargetsPipeline(dna)

        def nh():
            return pli.search_hmm(hmm, block)
        th, ms, wall, per = tm.run(nh, steps=max(1, args.extra_steps // 2), warmup=1)
        cells = float(model["M"]) * th.searched_residues
        blk = {"workload": "nhmmer: DNA profile M=%d (max_length %d) vs %.0f Mb synthetic genome, both strands, %d planted homologs%s" %
                           (model["M"], hmm.max_length, args.c5_mb, nplant, "" if world == 1 else ", windows sharded over %d ranks" % world),
               "value": cells / (wall * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": wall, "device_event_ms_per_step": ms, "steps": len(per),
               "timed": "host wall clock between device synchronisations (the window cutting and the hit stage run on host threads)",
               "residues_searched": int(th.searched_residues), "hits": len(th), "hits_reported": len(th.reported),
               "pos_past": [int(th.pos_past_msv), int(th.pos_past_bias), int(th.pos_past_vit), int(th.pos_past_fwd)],
               "calibrated_statistics": bool(cal)}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import refshim
            nc = _cores()
            sample = genome[:int(args.c5_ref_mb * 1000000)]
            ref = bench_inputs.to_ref_models([model], nthreads=1, L=100)[0] if model.get("max_length", -1) > 0 else None
            if ref is None:
                model2 = dict(model); model2["max_length"] = int(hmm.max_length)
                ref = bench_inputs.to_ref_models([model2], nthreads=1, L=100)[0]
            t0 = time.perf_counter()
            rn, rstats = refshim.nhmmer_mt(ref, [sample], nc, evalue_window=int(hmm.max_length))
            dt = time.perf_counter() - t0
            sblock = easel.DigitalSequenceBlock(dna, [easel.DigitalSequence(dna, name=b"genome", sequence=sample)])
            ours = plan7.LongTargetsPipeline(dna).search_hmm(hmm, sblock)
            blk["cpu_baseline"] = {"value": float(model["M"]) * rstats[0] / dt / 1e9, "unit": "GCUPS", "cores": nc, "kind": "reference",
                                   "ms_per_step": dt * 1e3,
                                   "sample": "the first %.0f Mb of the genome (both strands), windows of 262 144 residues spread over %d threads "
                                             "(ref_nhmmer_mt; pyhmmer's own nhmmer gives one query ONE thread)" % (args.c5_ref_mb, nc),
                                   "hits": int(rn), "pos_past": [int(v) for v in rstats[2:6]]}
            nd = sum(1 for h in ours if not h.duplicate)
            blk["matches_reference"] = bool(nd == rn and [int(ours.pos_past_msv), int(ours.pos_past_bias), int(ours.pos_past_vit), int(ours.pos_past_fwd)] == [int(v) for v in rstats[2:6]])
            blk["sample_check"] = {"hits": nd, "pos_past": [int(ours.pos_past_msv), int(ours.pos_past_bias), int(ours.pos_past_vit), int(ours.pos_past_fwd)]}
        extra["c5_nhmmer_100mb"] = blk
        log("[extras] c5: %s" % {k: v for k, v in blk.items() if k != "cpu_baseline"})
        del genome, block

    # ------------------------------------------------------------------ C3: 20k profiles x 100k proteins, strong scaling
    if "c3" in which:
        models, oms = profile_set()
        sumM = float(sum(m["M"] for m in models))
        t0 = time.perf_counter()
        seqs = bench_inputs.make_sequences(args.c3_seqs, seed=6)
        bench_inputs.plant(seqs, models[::max(1, len(models) // 400)], args.c3_seqs // 250, seed=7)
        full = bench_inputs.to_block(seqs, amino, prefix="c3_")
        w = parallel.World.current()
        lo, sub = parallel.shard_block(full, w) if world > 1 else (0, full)
        log("[extras] c3: %d proteins (%d residues) generated in %.1f s; this rank holds %d" % (len(full), full.total_residues, time.perf_counter() - t0, len(sub)))
        pli = plan7.Pipeline(amino)
        plan7.OptimizedProfile._device_many(ctx, oms)            # every rank holds ALL profile tables (the hmmscan block above shards them)
        pli._run(oms[:64], sub)                                  # warm the allocator and the kernels on a small slice
        torch.cuda.synchronize()

        def search():
            hits, doms, text, counters = pli._run(oms, sub)
            if world > 1:                                        # the single exchange of the path
                parts = [parallel.unpack_records(b) for b in parallel.all_gather_bytes(parallel.pack_records(hits, doms, text, counters, lo), w)]
                hits, doms, text, counters = parallel.merge_rank_records(parts)
                counters = counters.reshape(len(oms), 4)
            return hits, counters
        (hits, counters), ms, wall, per = tm.run(search, steps=1 if world < 4 else 2, warmup=0)
        cells = sumM * full.total_residues
        blk = {"workload": "hmmsearch: %d Pfam-A-sized profiles (sum M = %d) vs %d proteins (%d residues)%s" %
                           (len(models), int(sumM), len(full), full.total_residues,
                            ", profile tables resident" + ("" if world == 1 else ", target database sharded by residues over %d ranks, one all-gather of hit records" % world)),
               "scaling": "strong", "value": cells / (ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms, "host_wall_ms_per_step": wall,
               "steps": len(per), "seqs_per_s": len(full) * len(models) / (ms * 1e-3),
               "comparisons_scored_to_completion": len(hits), "pipeline_counters": counters.sum(0).tolist(),
               "calibrated_statistics": bool(profile_set.calibrated)}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import refshim
            nc = _cores()
            step = max(1, len(models) // args.c3_ref_profiles)
            idx = list(range(0, len(models), step))
            refs = bench_inputs.to_ref_models([models[i] for i in idx], nthreads=nc)
            t0 = time.perf_counter()
            nh, ctr = refshim.search_mt(refs, seqs, nc)
            dt = time.perf_counter() - t0
            scells = float(sum(models[i]["M"] for i in idx)) * full.total_residues
            blk["cpu_baseline"] = {"value": scells / dt / 1e9, "unit": "GCUPS", "cores": nc, "kind": "reference", "ms_per_step": dt * 1e3,
                                   "sample": "every %d-th profile (%d profiles, sum M = %d) x all %d sequences, one pass on %d threads; the whole "
                                             "job would take %.0f s at this rate" % (step, len(idx), sum(models[i]["M"] for i in idx), len(full), nc,
                                                                                    cells / (scells / dt)),
                                   "pipeline_counters": ctr, "hits": int(nh)}
            mine = counters[idx].sum(0).tolist()
            nmine = sum(1 for h in hits if h.profile % step == 0 and h.profile // step < len(idx))
            blk["matches_reference"] = bool(mine == list(ctr) and nmine == nh)
            blk["sample_check"] = {"pipeline_counters": mine, "hits": nmine}
        extra["c3_hmmsearch_20k_x_100k"] = blk
        log("[extras] c3: %s" % {k: v for k, v in blk.items() if k != "cpu_baseline"})
    return extra

"""Where the end-to-end step's upload phase goes (configs[1] inputs): database packing + H2D, profile table building + H2D,
alone and side by side as Pipeline._run_waves runs them.  `B2H_TRACE=1 python tools/upload_probe.py`."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pyhmmer_b200 import plan7, _lib

ctx = _lib.context(0)
abc, hmms, seqs = bench.build_inputs(0, 1)
bench.apply_stats(hmms) or [setattr(h, "_evparam", [-8.0, 0.7, -8.0, 0.7, -4.0, 0.7]) for h in hmms]
pli = plan7.Pipeline(abc)
oms = [pli._optimized(h, len(seqs[0])) for h in hmms]
seqs._packed()


def drop():
    seqs._cache = {k: v for k, v in seqs._cache.items() if k in ("packed", "nres")}
    for om in oms:
        om._dev = {}
    torch.cuda.synchronize()


for it in range(6):
    drop()
    t0 = time.perf_counter(); plan7.SequenceDatabase.of(ctx, seqs); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    plan7.OptimizedProfile._device_many(ctx, oms); t3 = time.perf_counter()
    drop()
    t4 = time.perf_counter()
    th = threading.Thread(target=lambda: plan7.SequenceDatabase.of(ctx, seqs)); th.start()
    plan7.OptimizedProfile._device_many(ctx, oms); t5 = time.perf_counter(); th.join(); t6 = time.perf_counter()
    print("[probe] database: call %.2f ms (+ %.2f until the copy landed); profiles: %.2f ms; side by side: profiles %.2f, joined %.2f ms"
          % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t5 - t4) * 1e3, (t6 - t4) * 1e3), flush=True)

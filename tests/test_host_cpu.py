"""CPU: host-side logic of the product and the C ABI surface (no GPU needed).

 * HMM file parsing, profile configuration and the three score-system conversions are bit-identical to the
   reference's p7_hmmfile_Read / p7_ProfileConfig / p7_oprofile_Convert (through oracle/_ref);
 * the host-side domain definition reproduces the reference's hits given the reference's own parser specials;
 * libb2h.so loads and exports every symbol include/b2h.h declares;
 * the multi-GPU plumbing (sharding rule, record (de)serialisation, one all-gather, merge) under gloo, world_size 2.
"""
import ctypes
import gzip
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from pyhmmer_b200 import _lib, easel, plan7, synth, parallel
from oracle import refshim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
needs_ref = pytest.mark.skipif(not refshim.available(), reason="oracle/_ref not built")


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "b2h.h")).read()
    names = set(re.findall(r"\b(b2h_[a-z0-9_]+)\s*\(", header))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert len(names) >= 30


def test_no_cpu_fallback_without_device():
    h = ctypes.c_void_p()
    st = _lib.lib.b2h_ctx_create(9999, ctypes.byref(h))
    assert st == _lib.B2H_ECUDA and not h.value


@needs_ref
@pytest.mark.parametrize("name", ["PF02826", "Thioesterase", "KR", "LuxC", "RREFam"])
def test_model_preparation_bit_identical(amino, name):
    with tempfile.NamedTemporaryFile(suffix=".hmm") as tmp:
        with gzip.open(os.path.join(GOLD, "data", name + ".hmm.gz")) as f:
            tmp.write(f.read())
        tmp.flush()
        with plan7.HMMFile(tmp.name) as f:
            hmms = list(f)
        for idx, hmm in enumerate(hmms):
            ref = refshim.RefModel(tmp.name, idx, 400)
            t, mat, ins = ref.hmm_params()
            assert np.array_equal(t, hmm.transition_probabilities)
            assert np.array_equal(mat[1:], hmm.match_emissions[1:]) and np.array_equal(ins, hmm.insert_emissions)
            assert np.array_equal(ref.evparam, hmm._evparam) and np.array_equal(ref.cutoff, hmm._cutoff)
            assert np.array_equal(ref.compo, hmm._compo)
            prof = plan7.Profile(hmm.M, amino).configure(hmm, plan7.Background(amino), 400)
            tsc, rsc, xsc = ref.gm_params()
            assert np.array_equal(tsc[:hmm.M], prof.tsc) and np.array_equal(rsc[:, :, 0], prof.msc) and np.array_equal(xsc, prof.xsc)
            om = prof.to_optimized()
            msv, vr, vt, fr, ft = ref.om_tables()
            assert np.array_equal(msv, om.msv_cost) and np.array_equal(vr, om.vit_rsc) and np.array_equal(vt, om.vit_tsc)
            assert np.array_equal(fr, om.fwd_rsc) and np.array_equal(ft, om.fwd_tsc)
            s, d = ref.om_scalars(), om._desc
            assert (s["tbm_b"], s["tec_b"], s["tjb_b"], s["base_b"], s["bias_b"]) == (d.tbm_b, d.tec_b, d.tjb_b, d.base_b, d.bias_b)
            assert (s["base_w"], s["ddbound_w"]) == (d.base_w, d.ddbound_w)
            assert np.array_equal(s["xw"], np.array([list(r) for r in d.xw])) and np.array_equal(s["xf"], np.array([list(r) for r in d.xf], np.float32))
            # ... and the striped views rebuilt from our tables are the reference's vectors, padding included
            rbv, rwv, twv, rfv, tfv = ref.om_tables_striped()
            assert np.array_equal(om.rbv, rbv.reshape(amino.Kp, -1)) and np.array_equal(om.rwv, rwv.reshape(amino.Kp, -1))
            assert np.array_equal(om.twv, twv.reshape(-1)) and np.array_equal(om.rfv, rfv.reshape(amino.Kp, -1)) and np.array_equal(om.tfv, tfv.reshape(-1))
            cp = om.copy()
            assert cp == om and cp.msv_cost is not om.msv_cost and cp._desc.msv_cost != om._desc.msv_cost
            cp.fwd_rsc[0, 0] *= 1.01
            assert cp != om and om != prof
            # de-striping the reference's SSE tables through the ABI gives the same node-major tables
            rbv, rwv, twv, rfv, tfv = ref.om_tables_striped()
            o = [np.empty_like(om.msv_cost), np.empty_like(om.vit_rsc), np.empty_like(om.vit_tsc), np.empty_like(om.fwd_rsc), np.empty_like(om.fwd_tsc)]
            _lib.check(_lib.lib.b2h_destripe_oprofile(hmm.M, amino.Kp, *[_lib.ptr(a) for a in (rbv, rwv, twv, rfv, tfv)], *[_lib.ptr(a) for a in o]), "destripe")
            assert all(np.array_equal(a, b) for a, b in zip(o, (om.msv_cost, om.vit_rsc, om.vit_tsc, om.fwd_rsc, om.fwd_tsc)))


@needs_ref
def test_length_params(amino):
    with tempfile.NamedTemporaryFile(suffix=".hmm") as tmp:
        synth.random_hmm(amino, 40, np.random.default_rng(1)).write(tmp)
        tmp.flush()
        ref = refshim.RefModel(tmp.name, 0, 400)
        for L in (1, 2, 3, 10, 99, 350, 1500, 35000, 100000):
            lp = _lib.LenParams()
            _lib.lib.b2h_length_params(L, 1.0, ctypes.byref(lp))
            ref.L.refm_set_length(ref.h, L)
            s = ref.om_scalars()
            assert lp.tjb_b == s["tjb_b"] and lp.xw_move == s["xw"][1][0]
            assert np.float32(lp.pmove) == s["xf"][1][0] and np.float32(lp.ploop) == s["xf"][1][1]
            assert np.float32(lp.null1) == np.float32(ref.null1(np.zeros(L, np.uint8)))


@needs_ref
@pytest.mark.parametrize("name", ["PF02826", "KR"])
def test_domain_definition_on_host(amino, name):
    """b2h_debug_domaindef fed with the reference's parser specials reproduces the reference's hits exactly:
    envelopes, alignments, null2 corrections, stochastic-traceback clustering (Easel's LCG stream)."""
    with easel.SequenceFile(os.path.join(GOLD, "data", "proteome.faa.gz"), digital=True, alphabet=amino) as f:
        seqs = f.read_block()
    with tempfile.NamedTemporaryFile(suffix=".hmm") as tmp:
        with gzip.open(os.path.join(GOLD, "data", name + ".hmm.gz")) as f:
            tmp.write(f.read())
        tmp.flush()
        with plan7.HMMFile(tmp.name) as f:
            hmm = f.read()
        ref = refshim.RefModel(tmp.name, 0, 400)
        rh, rd, rtext, rc = ref.search([s.sequence for s in seqs])
        om = plan7.Profile(hmm.M, amino).configure(hmm, plan7.Background(amino), 400).to_optimized()
        hp = ctypes.c_void_p()
        _lib.check(_lib.lib.b2h_profile_create_host(ctypes.byref(om._desc), ctypes.byref(hp)), "create_host")
        _lib.lib.b2h_profile_set_annotation(hp, hmm.consensus.encode(), None, None, amino.symbols.encode())
        prm = _lib.SearchParams(0.02, 1e-3, 1e-5, 1, 1, 42, 1)
        nclustered = 0
        for h in rh:
            s = seqs[h.seq]
            f_, b_, st, fx, bx = ref.fwdbck(s.sequence, want_x=True)
            out = ctypes.c_void_p()
            codes = np.ascontiguousarray(s.sequence)
            _lib.check(_lib.lib.b2h_debug_domaindef(hp, _lib.ptr(codes), len(s), _lib.ptr(fx), _lib.ptr(bx), f_, ctypes.byref(prm), ctypes.byref(out)), "ddef")
            hits, doms, text = _lib.read_results(out)
            _lib.lib.b2h_results_destroy(out)
            assert len(hits) == 1
            m = hits[0]
            assert abs(m.score - h.score) < 1e-3 and abs(m.sum_score - h.sum_score) < 1e-3 and abs(m.pre_score - h.pre_score) < 1e-3
            assert (m.nregions, m.nclustered, m.noverlaps, m.nenvelopes, m.ndom, m.best_domain) == (h.nregions, h.nclustered, h.noverlaps, h.nenvelopes, h.ndom, h.best_domain)
            nclustered += h.nclustered
            for d in range(h.ndom):
                a, r = doms[d], rd[h.dom_offset + d]
                assert (a.ienv, a.jenv, a.iali, a.jali, a.hmmfrom, a.hmmto, a.sqfrom, a.sqto, a.N) == (r.ienv, r.jenv, r.iali, r.jali, r.hmmfrom, r.hmmto, r.sqfrom, r.sqto, r.N)
                assert abs(a.envsc - r.envsc) < 2e-3 and abs(a.domcorrection - r.domcorrection) < 2e-3 and abs(a.bitscore - r.bitscore) < 2e-3 and abs(a.oasc - r.oasc) < 2e-3
                assert text[a.text_offset:a.text_offset + 4 * (a.N + 1)] == rtext[r.text_offset:r.text_offset + 4 * (r.N + 1)]
        _lib.lib.b2h_profile_destroy(hp)
        if name == "PF02826":
            assert nclustered > 0            # the stochastic clustering branch was exercised


def test_shard_bounds_follow_the_reference_rule():
    rng = np.random.default_rng(0)
    lens = rng.integers(50, 1500, 5000).tolist()
    for n in (1, 2, 4, 8):
        b = parallel.shard_bounds(lens, n)
        assert len(b) == n + 1 and b[0] == 0 and b[-1] == len(lens) and sorted(b) == b
        res = [sum(lens[b[i]:b[i + 1]]) for i in range(n)]
        assert sum(res) == sum(lens) and max(res) <= 1.02 * sum(lens) / n + 1500
    # literal reference behaviour on a small case: chunksize = 72, a cut where the running size first exceeds it
    assert parallel.shard_bounds([10, 10, 10, 10, 100, 1, 1, 1], 2) == [0, 4, 8]
    assert parallel.shard_bounds([5] * 3, 8)[-1] == 3 and len(parallel.shard_bounds([5] * 3, 8)) == 9


def _fake_records(rank):
    hits, doms = [], []
    text = b""
    for j in range(3 + rank):
        h = _lib.HitRec()
        h.profile, h.seq, h.score, h.lnP, h.ndom, h.dom_offset = j % 2, j, 10.0 * rank + j, -5.0 - j, 1, len(doms)
        d = _lib.DomainRec()
        d.N, d.text_offset, d.bitscore = 2, len(text), 3.0 + j
        text += b"AB\0ab\0AB\0**\0"
        hits.append(h)
        doms.append(d)
    return hits, doms, text, np.arange(8, dtype=np.int64).reshape(2, 4) + rank


def test_pack_unpack_roundtrip():
    hits, doms, text, ctr = _fake_records(1)
    h2, d2, t2, c2 = parallel.unpack_records(parallel.pack_records(hits, doms, text, ctr, 100))
    assert [h.seq for h in h2] == [h.seq + 100 for h in hits] and [h.score for h in h2] == [h.score for h in hits]
    assert t2 == text and np.array_equal(c2.reshape(2, 4), ctr) and [d.text_offset for d in d2] == [d.text_offset for d in doms]


_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from pyhmmer_b200 import parallel, _lib
sys.path.insert(0, os.path.join(%r, "tests"))
from test_host_cpu import _fake_records
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%d" %% int(sys.argv[1]), rank=int(sys.argv[2]), world_size=2)
w = parallel.World.current()
assert (w.rank, w.size) == (int(sys.argv[2]), 2)
hits, doms, text, ctr = _fake_records(w.rank)
parts = [parallel.unpack_records(b) for b in parallel.all_gather_bytes(parallel.pack_records(hits, doms, text, ctr, 1000 * w.rank), w)]
mh, md, mt, mc = parallel.merge_rank_records(parts)
assert len(mh) == 3 + 4 and [ (h.profile, h.seq) for h in mh ] == sorted((h.profile, h.seq) for h in mh)
assert np.array_equal(mc.reshape(2, 4), np.arange(8).reshape(2, 4) * 2 + 1)
for h in mh:
    d = md[h.dom_offset]
    assert mt[d.text_offset:d.text_offset + 12] == b"AB\0ab\0AB\0**\0"
# ragged sizes, incl. empty and larger than the exchange's initial slot (the capacity grows and the exchange repeats once)
big = bytes(np.random.default_rng(w.rank).integers(0, 256, 200000 if w.rank == 1 else 0, dtype=np.uint8))
got = parallel.all_gather_bytes(big, w)
assert [len(g) for g in got] == [0, 200000] and got[w.rank] == big
assert parallel.all_gather_bytes(b"x" * (w.rank + 1), w) == [b"x", b"xx"]
dist.barrier(); dist.destroy_process_group()
print("rank", w.rank, "ok")
'''


def test_all_gather_and_merge_world_size_2_gloo():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_ = s.getsockname()[1]; s.close()
    code = _WORKER % (ROOT, ROOT)
    procs = [subprocess.Popen([sys.executable, "-c", code, str(port_), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


_SEARCH_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from pyhmmer_b200 import parallel, _lib, plan7, easel, synth
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%d" %% int(sys.argv[1]), rank=int(sys.argv[2]), world_size=2)
w = parallel.World.current()
abc = easel.Alphabet.amino()
rng = np.random.default_rng(5)
bg = plan7.Background(abc)
hmms = [synth.random_hmm(abc, int(m), rng, name="p%%d" %% i) for i, m in enumerate(rng.integers(20, 300, 20))]
for h in hmms:
    h._evparam[:] = np.array([-8.0, 0.7, -9.0, 0.7, -4.0, 0.7], np.float32)
block = easel.DigitalSequenceBlock(abc, [easel.DigitalSequence(abc, name=b"t%%d" %% i, sequence=rng.integers(0, 20, 30 + 7 * i).astype(np.uint8)) for i in range(9)])

def records(oms, sub, profs):
    """stands in for the device: deterministic hits for (model, target) pairs, keyed by the objects (not their local indices)"""
    hits, doms, text = [], [], b""
    for p in profs:
        for s in range(len(sub)):
            key = oms[p].M + len(sub[s])
            if key %% 3 == 0:
                h = _lib.HitRec()
                h.profile, h.seq, h.score, h.pre_score, h.sum_score, h.lnP, h.ndom, h.dom_offset = p, s, key * 0.1, key * 0.1 + 1, key * 0.1, -float(key %% 23), 1, len(doms)
                d = _lib.DomainRec()
                d.N, d.text_offset, d.bitscore, d.lnP, d.ienv, d.jenv = 2, len(text), 3.0 + (key %% 5), -3.0, 1, 10
                text += b"AB\0ab\0AB\0**\0"
                hits.append(h); doms.append(d)
    counters = np.zeros((len(oms), 4), np.int64)
    for p in profs:
        counters[p] = [len(sub) * oms[p].M, len(sub), 2 * len(sub), sum(len(x) for x in sub)]      # additive over shards
    return hits, doms, text, counters

def fake_run(self, oms, sub, seq_counters=False):
    return records(oms, sub, range(len(oms)))

def fake_run_waves(self, oms, sub):
    # the ranks cut their waves differently (as they may when their shards differ): rank 0 in three, rank 1 in two
    order = sorted(range(len(oms)), key=lambda i: -oms[i].M)
    cuts = [0, 5, 12, len(oms)] if w.rank == 0 and w.size > 1 else [0, 9, len(oms)]
    waves = [order[a:b] for a, b in zip(cuts, cuts[1:])]
    return len(waves), ((wv,) + records(oms, sub, wv) for wv in waves)

pli = object.__new__(plan7.Pipeline)                     # no device here: the attributes the search path reads are set by hand
for k, v in dict(alphabet=abc, background=bg, bias_filter=True, null2=True, seed=42, Z=None, domZ=None, F1=0.02, F2=1e-3, F3=1e-5,
                 E=10.0, T=None, domE=10.0, domT=None, incE=0.01, incT=None, incdomE=0.01, incdomT=None, bit_cutoffs=None, host_threads=1).items():
    setattr(pli, k, v)
pli.clear()
plan7.Pipeline._run = fake_run
plan7.Pipeline._run_waves = fake_run_waves
sharded = parallel.search_sharded(pli, hmms, block, parallel.shard_block(block, w), w)
assert (pli._nseqs, pli._nres) == (len(block), block.total_residues)
single = pli._search_many(hmms, block)                    # one process, wave by wave
oms = pli._optimized_many(hmms, len(block[0]))
whole = pli._assemble(hmms, oms, block, *fake_run(pli, oms, block))      # one process, one blocking call
sig = lambda ths: [[(h.name, h.score, h.lnP, h.reported, h.included, len(h.domains)) for h in th] +
                   [(th.Z, th.searched_sequences, th.n_past_msv, th.n_past_bias, th.n_past_vit, th.n_past_fwd)] for th in ths]
assert sig(sharded) == sig(whole) and sig(single) == sig(whole) and sum(len(th) for th in whole) >= 30
assert [th.query.name for th in sharded] == [h.name for h in hmms]
# an empty shard on one rank: it still takes part in every round
tiny = easel.DigitalSequenceBlock(abc, list(block)[:1])
a = parallel.search_sharded(pli, hmms, tiny, parallel.shard_block(tiny, w), w)
assert [len(th) for th in a] == [len(th) for th in pli._assemble(hmms, oms, tiny, *fake_run(pli, oms, tiny))]
dist.barrier(); dist.destroy_process_group()
print("rank", w.rank, "ok")
'''


def test_search_wave_exchange_world_size_2_gloo():
    """hmmsearch over two ranks, wave by wave: every rank searches its shard in waves of profiles, one all-gather of hit
    records per wave, profiles assembled as soon as every rank has finished them -- with ranks that cut their waves
    DIFFERENTLY and a rank with an empty shard; results identical to the unsharded search (host logic, stand-in device)."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_ = s.getsockname()[1]; s.close()
    code = _SEARCH_WORKER % (ROOT,)
    procs = [subprocess.Popen([sys.executable, "-c", code, str(port_), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


_SCAN_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from pyhmmer_b200 import parallel, _lib, plan7, easel, synth
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%d" %% int(sys.argv[1]), rank=int(sys.argv[2]), world_size=2)
w = parallel.World.current()
abc = easel.Alphabet.amino()
rng = np.random.default_rng(3)
bg = plan7.Background(abc)
oms = [plan7.Profile(int(m), abc).configure(synth.random_hmm(abc, int(m), rng, name="p%%d" %% i), bg, 100).to_optimized()
       for i, m in enumerate(rng.integers(20, 300, 11))]
for om in oms:
    om._evparam[:] = np.array([-8.0, 0.7, -9.0, 0.7, -4.0, 0.7], np.float32)
queries = [easel.DigitalSequence(abc, name=b"q%%d" %% i, sequence=rng.integers(0, 20, 50 + i).astype(np.uint8)) for i in range(3)]

def fake_run(self, local, block, seq_counters=False):
    """stands in for the device: a deterministic set of hits for (profile, query) pairs, keyed by the MODEL (not its index)"""
    assert seq_counters                                   # scan mode keeps the pass counters per query sequence
    hits, doms, text = [], [], b""
    for p, om in enumerate(local):
        for s in range(len(block)):
            if (om.M + s) %% 3 == 0:
                h = _lib.HitRec()
                h.profile, h.seq, h.score, h.pre_score, h.sum_score, h.lnP, h.ndom, h.dom_offset = p, s, om.M * 0.1 + s, om.M * 0.1 + s + 1, om.M * 0.1, -float(om.M %% 17) - s, 1, len(doms)
                d = _lib.DomainRec()
                d.N, d.text_offset, d.bitscore, d.lnP, d.ienv, d.jenv = 2, len(text), 3.0 + s, -3.0, 1, 10
                text += b"AB\0ab\0AB\0**\0"
                hits.append(h); doms.append(d)
    return hits, doms, text, np.array([[sum(om.M + s for om in local), len(local), 2 * len(local), sum(om.M %% 3 for om in local)] for s in range(len(block))], np.int64)

pli = object.__new__(plan7.Pipeline)                     # no device here: the attributes the scan path reads are set by hand
for k, v in dict(alphabet=abc, background=bg, bias_filter=True, null2=True, seed=42, Z=None, domZ=None, F1=0.02, F2=1e-3, F3=1e-5,
                 E=10.0, T=None, domE=10.0, domT=None, incE=0.01, incT=None, incdomE=0.01, incdomT=None, bit_cutoffs=None, host_threads=1).items():
    setattr(pli, k, v)
pli.clear()
plan7.Pipeline._run = fake_run
sharded = pli._scan_many(queries, oms, world=w)
single = pli._scan_many(queries, oms, world=None)
sig = lambda ths: [[(h.name, h.score, h.lnP, h.reported, h.included, h.domains[0].alignment.hmm_from, len(h.domains)) for h in th] for th in ths]
assert sig(sharded) == sig(single) and sum(len(th) for th in single) >= 8
assert [th.Z for th in sharded] == [11.0] * 3 and [th.searched_models for th in sharded] == [11] * 3
ctr = lambda ths: [(th.n_past_msv, th.n_past_bias, th.n_past_vit, th.n_past_fwd) for th in ths]
assert ctr(sharded) == ctr(single)                       # per-query counters: the ranks' shares add up
assert ctr(single)[1] == (sum(om.M + 1 for om in oms), 11, 22, sum(om.M %% 3 for om in oms))
dist.barrier(); dist.destroy_process_group()
print("rank", w.rank, "ok")
'''


def test_scan_profile_sharding_world_size_2_gloo():
    """hmmscan over two ranks: the profile block is sharded by nodes, one all-gather of the hit records, identical `TopHits`
    on every rank -- host logic with a stand-in for the device search."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_ = s.getsockname()[1]; s.close()
    code = _SCAN_WORKER % (ROOT,)
    procs = [subprocess.Popen([sys.executable, "-c", code, str(port_), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


@pytest.mark.parametrize("name", ["Thioesterase", "PF02826"])
def test_pressed_database_matches_conversion(amino, name):
    """HMMPressedFile (p7_oprofile_ReadMSV / ReadRest, impl_sse/io.c:231,498) on the reference's own hmmpress'ed fixtures:
    the de-striped tables and scalars are identical to Profile.configure + to_optimized of the ASCII model -- which also
    pins our p7_oprofile_Convert against bytes written by HMMER itself."""
    import gzip
    from pyhmmer_b200 import plan7
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")
    with plan7.HMMPressedFile(os.path.join(gold, "pressed", name + ".hmm")) as pf:
        oms = list(pf)
        pf.rewind()
        assert len(list(pf)) == len(oms)
    with gzip.open(os.path.join(gold, name + ".hmm.gz")) as f:
        hmms = list(plan7.HMMFile(f))
    assert len(oms) == len(hmms) >= 1
    with plan7._PyPressedFile(os.path.join(gold, "pressed", name + ".hmm")) as pf:      # the format, parsed in pure Python
        for om, py in zip(oms, pf):
            assert all(np.array_equal(getattr(om, t), getattr(py, t)) for t in ("msv_cost", "vit_rsc", "vit_tsc", "fwd_rsc", "fwd_tsc"))
            assert (om.name, om.accession, om.description, om.consensus, om.reference, om.consensus_structure, om.model_mask) == \
                   (py.name, py.accession, py.description, py.consensus, py.reference, py.consensus_structure, py.model_mask)
            assert list(om._desc.bgf) == list(py._desc.bgf) and om.L == py.L
    bg = plan7.Background(amino)
    for om, h in zip(oms, hmms):
        ref = plan7.Profile(h.M, amino).configure(h, bg, om.L).to_optimized()
        assert (om.name, om.accession, om.description, om.M, om.multihit) == (h.name, h.accession, h.description, h.M, True)
        assert om.consensus == ref.consensus and om.consensus_structure == ref.consensus_structure and om.reference == ref.reference
        for t in ("msv_cost", "vit_rsc", "vit_tsc", "fwd_rsc", "fwd_tsc"):
            assert np.array_equal(getattr(om, t), getattr(ref, t)), t
        for fld in ("tbm_b", "tec_b", "base_b", "bias_b", "scale_b", "base_w", "ddbound_w", "scale_w", "max_length", "mode_multihit"):
            assert getattr(om._desc, fld) == getattr(ref._desc, fld), fld
        assert [list(r) for r in om._desc.xw] == [list(r) for r in ref._desc.xw]
        assert [list(r) for r in om._desc.xf] == [list(r) for r in ref._desc.xf]
        assert list(om._desc.evparam) == list(ref._desc.evparam) and list(om._desc.cutoff) == list(ref._desc.cutoff)
        assert list(om._desc.compo) == list(ref._desc.compo)


@pytest.mark.parametrize("threads", [1, 3])
def test_batched_conversion_matches_per_model(amino, threads):
    """b2h_hmm_convert_many (the query-block form of Profile.configure + to_optimized, plan7.pyx:5979-6013): tables and
    every descriptor scalar are bit-identical to the per-model path, for synthetic and for the reference's own models."""
    import gzip
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")
    rng = np.random.default_rng(11)
    hmms = [synth.random_hmm(amino, int(m), rng) for m in [1, 2, 3, 15, 16, 17, 255, 256, 700] + list(rng.integers(4, 400, 20))]
    for name in ("Thioesterase", "PF02826", "KR"):
        with gzip.open(os.path.join(gold, name + ".hmm.gz")) as f:
            hmms += list(plan7.HMMFile(f))
    bg = plan7.Background(amino)
    for L in (400, 37):
        many = plan7._convert_hmms(hmms, bg, L, threads=threads)
        assert len(many) == len(hmms)
        for h, om in zip(hmms, many):
            ref = plan7.Profile(h.M, amino).configure(h, bg, L).to_optimized()
            assert (om.name, om.accession, om.M, om.L, om.multihit) == (ref.name, ref.accession, ref.M, ref.L, ref.multihit)
            assert (om.consensus, om.reference, om.consensus_structure) == (ref.consensus, ref.reference, ref.consensus_structure)
            for t in ("msv_cost", "vit_rsc", "vit_tsc", "fwd_rsc", "fwd_tsc"):
                a, b = getattr(om, t), getattr(ref, t)
                assert a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes(), t
            for fld, _ in om._desc._fields_:
                if fld in ("msv_cost", "vit_rsc", "vit_tsc", "fwd_rsc", "fwd_tsc", "degen"):
                    continue
                a, b = getattr(om._desc, fld), getattr(ref._desc, fld)
                assert (bytes(a) == bytes(b)) if hasattr(a, "__len__") else (a == b), fld
            assert np.array_equal(om._evparam, ref._evparam) and np.array_equal(om._cutoff, ref._cutoff)
    # a Pipeline takes the batched path for blocks of HMM queries and leaves Profile / OptimizedProfile queries alone
    pli = object.__new__(plan7.Pipeline)                  # no device here: only the query preparation is exercised
    pli.alphabet, pli.background, pli.host_threads = amino, bg, threads
    pre = plan7.Profile(hmms[3].M, amino).configure(hmms[3], bg, 99).to_optimized()
    mixed = pli._optimized_many(hmms[:6] + [pre], 123)
    assert mixed[-1] is pre and all(o.L == 123 for o in mixed[:6]) and [o.M for o in mixed[:6]] == [h.M for h in hmms[:6]]
    assert plan7._convert_hmms([], bg, 400) == []
    with pytest.raises(Exception):
        plan7._convert_hmms([synth.random_hmm(easel.Alphabet.dna(), 10, rng)], bg, 400)


def test_pressed_database_errors(tmp_path):
    from pyhmmer_b200 import plan7
    with pytest.raises(ValueError):
        plan7.HMMPressedFile(str(tmp_path / "missing.hmm"))
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data", "pressed", "PF02826.hmm")
    for ext in (".h3f", ".h3p"):
        data = open(gold + ext, "rb").read()
        open(str(tmp_path / ("bad.hmm" + ext)), "wb").write(data)
    open(str(tmp_path / "bad.hmm.h3f"), "wb").write(b"\x00" * 64)           # wrong magic
    with pytest.raises(ValueError):
        plan7.HMMPressedFile(str(tmp_path / "bad.hmm")).read()


@pytest.mark.parametrize("M", [5, 120, 700])
def test_window_prefix_suffix_lengths(M):
    """p7_hmm_ScoreDataComputeRest (p7_scoredata.c:313): the MAXL-based prefix/suffix tables that turn SSV diagonals into
    windows -- host code, bit-identical (entry 0 of the suffix table is never written by the reference)."""
    dna = easel.Alphabet.dna()
    rng = np.random.default_rng(M)
    h = synth.random_hmm(dna, M, rng, name="lt")
    h.max_length = 3 * M
    h._evparam[:] = np.array([-8.0 - np.log2(M) * 0.3, 0.70, -9.0, 0.70, -4.0, 0.70], np.float32)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import ModelPair
    pair = ModelPair(h)
    o = ctypes.c_void_p()
    assert _lib.lib.b2h_profile_create_host(ctypes.byref(pair.om._desc), ctypes.byref(o)) == 0
    pre, suf = np.zeros(M + 1, np.float32), np.zeros(M + 1, np.float32)
    assert _lib.lib.b2h_window_lengths(o, _lib.ptr(pre), _lib.ptr(suf)) == 0
    # ... and p7_pli_ExtendAndMergeWindows on the reference's own diagonals of a chunk with planted homologs
    chunk = rng.integers(0, 4, 60000).astype(np.uint8)
    for _ in range(12):
        dom = synth.emit_sequence(pair.hmm, rng)
        pos = int(rng.integers(0, len(chunk) - len(dom)))
        chunk[pos:pos + len(dom)] = dom
    pair.hmm._evparam[:] = h._evparam
    rraw, rsc, rmer, rpre, rsuf = pair.ref.longtarget_windows(chunk)
    assert np.array_equal(pre, rpre) and np.array_equal(suf[1:], rsuf[1:])
    if len(rraw):
        w = np.zeros(len(rraw), dtype=np.dtype(_lib.WindowRec))
        w["n"], w["k"], w["length"], w["score"] = rraw[:, 0], rraw[:, 1], rraw[:, 2], rsc
        tl = np.full(len(rraw), len(chunk), np.int64)
        nout = ctypes.c_size_t()
        assert _lib.lib.b2h_extend_merge_windows(o, _lib.ptr(w), len(w), _lib.ptr(tl), 0.0, ctypes.byref(nout)) == 0
        assert nout.value == len(rmer)
        assert np.array_equal(w["n"][:nout.value], rmer[:, 0]) and np.array_equal(w["length"][:nout.value], rmer[:, 1])
    _lib.lib.b2h_profile_destroy(o)


@needs_ref
@pytest.mark.parametrize("name", ["PF02826", "KR"])
def test_tophits_tables_match_the_reference_writers(amino, name, tmp_path):
    """`TopHits.write` (targets / domains / pfam) against the reference's own p7_tophits_TabularTargets / TabularDomains /
    TabularXfam on the same hits: the reference pipeline's raw hits (oracle/_ref) are assembled into `TopHits` by the
    product's host code (running-Z admission, sort, thresholds) and written -- byte for byte the reference's tables."""
    import io
    with easel.SequenceFile(os.path.join(GOLD, "data", "proteome.faa.gz"), digital=True, alphabet=amino) as f:
        seqs = f.read_block()
    with tempfile.NamedTemporaryFile(suffix=".hmm") as tmp:
        with gzip.open(os.path.join(GOLD, "data", name + ".hmm.gz")) as f:
            tmp.write(f.read())
        tmp.flush()
        with plan7.HMMFile(tmp.name) as f:
            hmm = f.read()
        ref = refshim.RefModel(tmp.name, 0, 400)
        prefix = str(tmp_path / "ref")
        nref = ref.search_tables([s.sequence for s in seqs], [s.name for s in seqs], [s.accession or None for s in seqs],
                                 [s.description or None for s in seqs], prefix)
        rh, rd, rtext, rc = ref.search([s.sequence for s in seqs])
    hits = (_lib.HitRec * len(rh))()
    doms = (_lib.DomainRec * len(rd))()
    for a, r in zip(hits, rh):
        a.profile = 0
        for fld in ("seq", "score", "pre_score", "sum_score", "nexpected", "lnP", "pre_lnP", "sum_lnP", "nregions", "nclustered", "noverlaps",
                    "nenvelopes", "ndom", "best_domain", "dom_offset"):
            setattr(a, fld, getattr(r, fld))
    for a, r in zip(doms, rd):
        for fld in ("ienv", "jenv", "iali", "jali", "envsc", "domcorrection", "dombias", "oasc", "bitscore", "lnP", "hmmfrom", "hmmto", "sqfrom",
                    "sqto", "N", "text_offset"):
            setattr(a, fld, getattr(r, fld))
    pli = object.__new__(plan7.Pipeline)                  # no device here: the attributes the assembly reads are set by hand
    for k, v in dict(alphabet=amino, background=plan7.Background(amino), bias_filter=True, null2=True, seed=42, Z=None, domZ=None, F1=0.02,
                     F2=1e-3, F3=1e-5, E=10.0, T=None, domE=10.0, domT=None, incE=0.01, incT=None, incdomE=0.01, incdomT=None, bit_cutoffs=None,
                     host_threads=1).items():
        setattr(pli, k, v)
    pli.clear()
    om = plan7.Profile(hmm.M, amino).configure(hmm, pli.background, 400).to_optimized()
    order = sorted(range(len(hits)), key=lambda i: hits[i].seq)
    th = pli._assemble([hmm], [om], seqs, [hits[i] for i in order], list(doms), rtext, np.array([rc], np.int64).reshape(1, 4))[0]
    assert len(th) == nref >= 1
    for fmt, ext in (("targets", ".tbl"), ("domains", ".domtbl"), ("pfam", ".pfam")):
        buf = io.BytesIO()
        th.write(buf, format=fmt)
        want = open(prefix + ext, "rb").read()
        assert buf.getvalue() == want, (fmt, buf.getvalue()[:600], want[:600])
        body = io.BytesIO()
        th.write(body, format=fmt, header=False)
        if fmt != "pfam":
            assert body.getvalue() == b"".join(l for l in want.splitlines(True) if not l.startswith(b"#"))
    with pytest.raises(ValueError):
        th.write(io.BytesIO(), format="xml")
    # every alignment prints as the reference prints it (Alignment.__str__ = p7_nontranslated_alidisplay_Print)
    blocks = open(prefix + ".ali").read().split(">> ")[1:]
    mine = [(h.name.decode() if isinstance(h.name, bytes) else h.name, i, str(d.alignment)) for h in th for i, d in enumerate(h.domains)]
    assert len(blocks) == len(mine) >= 1
    for blk, (hname, i, text_) in zip(blocks, mine):
        head, _, body = blk.partition("\n")
        # (the shim's hit records carry the four alignment lines every display has; the model's CS / RF lines are not in them)
        body = "".join(l for l in body.splitlines(True) if not l.endswith((" CS\n", " RF\n")))
        assert head == "%s %d" % (hname, i) and body == text_, (head, body, text_)
    # copies and pickles write the same tables; sorting by target index and back restores the order
    import pickle
    tables = lambda t: [(lambda b: (t.write(b, format=f), b.getvalue())[1])(io.BytesIO()) for f in ("targets", "domains", "pfam")]
    want = tables(th)
    cp = th.copy()
    rt = pickle.loads(pickle.dumps(th))
    assert tables(cp) == want and tables(rt) == want
    assert [h.name for h in rt] == [h.name for h in th] and (rt.Z, rt.domZ, rt.E, rt.incE, rt.T) == (th.Z, th.domZ, 10.0, 0.01, None)
    assert rt[0].domains[0].alignment.hmm_sequence == th[0].domains[0].alignment.hmm_sequence
    assert th.is_sorted()
    cp.sort(by="seqidx")
    assert [h._index for h in cp] == sorted(h._index for h in th) and cp.is_sorted(by="seqidx") and th.is_sorted(by="key")
    assert [h.name for h in th] != [h.name for h in cp] or len(th) < 3      # ... and the copy did not disturb the original
    cp.sort()
    assert tables(cp) == want
    with pytest.raises(ValueError):
        th.sort(by="name")


def test_hmm_body_parser_errors_and_odd_fields(amino):
    """The C field scanner behind HMMFile (b2h_hmm_parse_body): comments and blank lines inside the node table, fields that
    leave its fast decimal path (exponents, long mantissas) and malformed tables."""
    import io
    with gzip.open(os.path.join(GOLD, "data", "Thioesterase.hmm.gz")) as f:
        data = f.read()
    ref = plan7.HMMFile(io.BytesIO(data)).read()
    lines = data.split(b"\n")
    first = next(i for i, l in enumerate(lines) if l.split()[:1] == [b"1"])
    # a comment, a blank line, and the same number spelled three ways
    tok = lines[first].split()[1]
    odd = list(lines)
    odd[first] = lines[first].replace(tok, b"%.5fe0" % float(tok), 1)
    odd[first + 1] = lines[first + 1].replace(lines[first + 1].split()[0], lines[first + 1].split()[0] + b"0000000000000", 1)
    odd.insert(first, b"# a comment inside the table")
    odd.insert(first, b"")
    got = plan7.HMMFile(io.BytesIO(b"\n".join(odd))).read()
    assert np.array_equal(got.match_emissions, ref.match_emissions) and np.array_equal(got.insert_emissions, ref.insert_emissions)
    assert np.array_equal(got.transition_probabilities, ref.transition_probabilities) and got.consensus == ref.consensus
    bad = list(lines); bad[first] = lines[first].replace(b"1", b"2", 1)
    with pytest.raises(ValueError, match="node number"):
        plan7.HMMFile(io.BytesIO(b"\n".join(bad))).read()
    cut = list(lines); del cut[first + 4]
    with pytest.raises(ValueError):
        plan7.HMMFile(io.BytesIO(b"\n".join(cut))).read()
    with pytest.raises(ValueError):
        plan7.HMMFile(io.BytesIO(b"\n".join(lines[:first + 10]))).read()


@needs_ref
def test_domain_definition_randomised_sweep(amino):
    """Host domain definition from the reference's parser specials on seeded random models with planted (often adjacent)
    homologs: regions, clustering, envelopes, alignments and scores against the reference pipeline.  A sweep of 1 329 such
    cases (8 646 hits, 1 734 clustered regions) differed in 5 hits, all in the null2 correction of a clustered region by
    0.005-0.02 nats (bit scores within 2e-4): one of the 200 sampled traces takes another branch when a uniform deviate falls
    within float rounding of a cumulative probability.  Hence 0.05 nats on domcorrection here, 2e-3 bits on every score."""
    rng0 = np.random.default_rng(99)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import ModelPair
    nhits = nclu = 0
    for case in range(16):
        M = int(rng0.choice([15, 40, 90, 178, 300]))
        seed = int(rng0.integers(0, 10**6))
        rng = np.random.default_rng(seed)
        h = synth.random_hmm(amino, M, rng, name="f%d" % seed)
        h._evparam[:] = np.array([-8.0 - np.log2(M) * 0.3, 0.70, -9.0, 0.70, -4.0, 0.70], np.float32)
        pair = ModelPair(h)
        seqs = []
        for i in range(12):
            L = int(rng.integers(30, 900))
            codes = rng.integers(0, 20, L).astype(np.uint8)
            for _ in range(int(rng.integers(0, 4))):
                dom = synth.emit_sequence(pair.hmm, rng)
                if len(dom) < L:
                    pos = int(rng.integers(0, L - len(dom)))
                    codes[pos:pos + len(dom)] = dom
            seqs.append(codes)
        rh, rd, rtext, rc = pair.ref.search(seqs)
        hp = ctypes.c_void_p()
        _lib.check(_lib.lib.b2h_profile_create_host(ctypes.byref(pair.om._desc), ctypes.byref(hp)), "create_host")
        _lib.lib.b2h_profile_set_annotation(hp, (pair.hmm.consensus or "x" * M).encode(), None, None, amino.symbols.encode())
        prm = _lib.SearchParams(0.02, 1e-3, 1e-5, 1, 1, 42, 1)
        for hh in rh:
            codes = np.ascontiguousarray(seqs[hh.seq])
            f_, b_, st, fx, bx = pair.ref.fwdbck(codes, want_x=True)
            out = ctypes.c_void_p()
            _lib.check(_lib.lib.b2h_debug_domaindef(hp, _lib.ptr(codes), len(codes), _lib.ptr(fx), _lib.ptr(bx), f_, ctypes.byref(prm), ctypes.byref(out)), "ddef")
            hits, doms, text = _lib.read_results(out)
            _lib.lib.b2h_results_destroy(out)
            assert len(hits) == 1
            m = hits[0]
            assert abs(m.score - hh.score) < 2e-3 and abs(m.sum_score - hh.sum_score) < 2e-3 and abs(m.pre_score - hh.pre_score) < 2e-3
            assert (m.nregions, m.nclustered, m.noverlaps, m.nenvelopes, m.ndom, m.best_domain) == \
                   (hh.nregions, hh.nclustered, hh.noverlaps, hh.nenvelopes, hh.ndom, hh.best_domain)
            nclu += hh.nclustered
            for d in range(hh.ndom):
                a, r = doms[d], rd[hh.dom_offset + d]
                assert (a.ienv, a.jenv, a.iali, a.jali, a.hmmfrom, a.hmmto, a.sqfrom, a.sqto, a.N) == \
                       (r.ienv, r.jenv, r.iali, r.jali, r.hmmfrom, r.hmmto, r.sqfrom, r.sqto, r.N)
                assert abs(a.envsc - r.envsc) < 2e-3 and abs(a.bitscore - r.bitscore) < 2e-3 and abs(a.domcorrection - r.domcorrection) < 0.05
                la = text[a.text_offset:a.text_offset + 4 * (a.N + 1)].split(b"\0")
                lr = rtext[r.text_offset:r.text_offset + 4 * (r.N + 1)].split(b"\0")
                assert la[:3] == lr[:3]                  # model, match and sequence lines; (the posterior line may differ by a neighbouring class)
            nhits += 1
        _lib.lib.b2h_profile_destroy(hp)
    assert nhits >= 60 and nclu >= 10


@needs_ref
@pytest.mark.parametrize("alphabet", ["amino", "dna"])
def test_pressed_databases_written_by_the_reference(alphabet, tmp_path):
    """HMMPressedFile on databases the reference's own p7_oprofile_Write produces from synthetic models: both alphabets
    (Kp = 29 and 18), model lengths around every striping boundary (Q = 2 minimum; 4 / 8 / 16 lanes) and a long one."""
    abc = getattr(easel.Alphabet, alphabet)()
    rng = np.random.default_rng(17)
    Ms = [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 64, 65, 200, 1100]
    hmms = []
    for i, M in enumerate(Ms):
        h = synth.random_hmm(abc, M, rng, name="m%d" % i)
        h._evparam[:] = np.array([-8.0, 0.70, -9.0, 0.70, -4.0, 0.70], np.float32)
        h.max_length = 3 * M + 10
        h.accession, h.description = ("ACC%05d.1" % i, "model %d of the sweep" % i) if i % 2 else (None, None)
        if i % 3 == 0:
            h._cutoff[:] = np.array([25.0, 20.0, 30.0, 22.0, 18.0, 15.0], np.float32)
        hmms.append(h)
    path = str(tmp_path / "db.hmm")
    with open(path, "wb") as f:
        for h in hmms:
            h.write(f)
    assert refshim.press(path, path) == len(hmms)
    bg = plan7.Background(abc)
    with plan7.HMMFile(path) as f:
        assert f.is_pressed()
        back = list(f)
    with plan7.HMMPressedFile(path) as pf:
        oms = list(pf)
    assert len(oms) == len(hmms) == len(back)
    for om, h in zip(oms, back):
        want = plan7.Profile(h.M, abc).configure(h, bg, 400).to_optimized()
        assert om == want and (om.name, om.accession, om.description, om.M) == (h.name, h.accession, h.description, h.M)
        for t in ("msv_cost", "vit_rsc", "vit_tsc", "fwd_rsc", "fwd_tsc"):
            assert np.array_equal(getattr(om, t), getattr(want, t)), (h.M, t)
        assert list(om._desc.evparam) == list(want._desc.evparam) and list(om._desc.cutoff) == list(want._desc.cutoff)
        assert om._desc.max_length == h.max_length and list(om._desc.compo) == list(want._desc.compo)


@needs_ref
def test_binary_hmm_files(amino, tmp_path):
    """HMMFile on HMMER3 binary files (read_bin30hmm, p7_hmmfile.c:1585): the reference's own .h3m fixtures give the models
    its ASCII files give; HMM.write(binary=True) (p7_hmmfile_WriteBinary, :714) round-trips, and the reference library reads
    what we write -- for both alphabets."""
    fields = ("name", "accession", "description", "consensus", "reference", "model_mask", "consensus_structure", "max_length",
              "nseq", "checksum", "M", "command_line", "creation_time")
    arrays = ("transition_probabilities", "match_emissions", "insert_emissions", "_evparam", "_cutoff", "_compo")

    def same(x, y):
        K = x.alphabet.K
        unset = lambda c: np.zeros(K, np.float32) if c[0] == plan7.P7_COMPO_UNSET else c[:K]   # without COMPO a model reads back zeroed (p7_hmm_CreateBody)
        assert all(np.array_equal(getattr(x, f), getattr(y, f)) for f in arrays[:-1]) and np.array_equal(unset(x._compo), unset(y._compo))
        assert [getattr(x, f) for f in fields] == [getattr(y, f) for f in fields], [(f, getattr(x, f), getattr(y, f)) for f in fields if getattr(x, f) != getattr(y, f)]
        assert (x.map is None) == (y.map is None) and (x.map is None or np.array_equal(x.map, y.map))

    for name in ("PF02826", "Thioesterase"):
        with gzip.open(os.path.join(GOLD, "data", name + ".hmm.gz")) as f:
            text = list(plan7.HMMFile(f))
        binary = list(plan7.HMMFile(os.path.join(GOLD, "data", "pressed", name + ".hmm.h3m")))
        assert len(text) == len(binary) >= 1
        for a, b in zip(text, binary):
            same(a, b)
    rng = np.random.default_rng(5)
    for abc in (amino, easel.Alphabet.dna()):
        hmms = [synth.random_hmm(abc, M, rng, name="b%d" % M) for M in (1, 2, 33, 150)]
        hmms[1].accession, hmms[1].description, hmms[2].max_length = "ACC1.2", "a model with a description", 99
        hmms[3].model_mask = "".join("m" if 10 <= k < 20 else "." for k in range(1, 151))
        hmms[3]._evparam[:] = np.array([-8.0, 0.7, -9.0, 0.7, -4.0, 0.7], np.float32)
        path = str(tmp_path / ("db_%s.h3m" % abc.type.lower()))
        with open(path, "wb") as f:
            for h in hmms:
                h.write(f, binary=True)
        back = list(plan7.HMMFile(path))
        assert len(back) == len(hmms)
        for a, b in zip(hmms, back):
            same(a, b)
        for i, h in enumerate(hmms):                     # ... and HMMER reads the same numbers from our file
            ref = refshim.RefModel(path, i, 400)
            t, mat, ins = ref.hmm_params()
            assert np.array_equal(t, h.transition_probabilities) and np.array_equal(mat[1:], h.match_emissions[1:]) and np.array_equal(ins, h.insert_emissions)
    with pytest.raises(ValueError):
        list(plan7.HMMFile(io_bytes(open(path, "rb").read()[:200])))


def io_bytes(b):
    import io
    return io.BytesIO(b)


def _read_pfam_stockholm(path):
    names, rows, pp, gs, gc = [], {}, {}, {}, {}
    for line in open(path):
        line = line.rstrip("\n")
        if not line or line == "//" or line.startswith("# "):
            continue
        if line.startswith("#=GS"):
            _, name, tag, val = line.split(None, 3)
            gs[(name, tag)] = val
        elif line.startswith("#=GR"):
            _, name, tag, val = line.split(None, 3)
            assert tag == "PP"
            pp[name] = val
        elif line.startswith("#=GC"):
            _, tag, val = line.split(None, 2)
            gc[tag] = val
        else:
            name, val = line.split(None, 1)
            names.append(name)
            rows[name] = val
    return names, rows, pp, gs, gc


@needs_ref
@pytest.mark.parametrize("name,all_cols", [("PF02826", False), ("PF02826", True), ("KR", False)])
def test_to_msa_matches_the_reference_alignment(amino, name, all_cols, tmp_path):
    """`TopHits.to_msa` against p7_tophits_Alignment on the same hits (written by the reference as Pfam Stockholm): row names,
    every aligned row, the PP rows, the RF line and the consensus PP line."""
    import io
    with easel.SequenceFile(os.path.join(GOLD, "data", "proteome.faa.gz"), digital=True, alphabet=amino) as f:
        seqs = f.read_block()
    with tempfile.NamedTemporaryFile(suffix=".hmm") as tmp:
        with gzip.open(os.path.join(GOLD, "data", name + ".hmm.gz")) as f:
            tmp.write(f.read())
        tmp.flush()
        with plan7.HMMFile(tmp.name) as f:
            hmm = f.read()
        ref = refshim.RefModel(tmp.name, 0, 400)
        path = str(tmp_path / "ref.sto")
        nref = ref.search_msa([s.sequence for s in seqs], [s.name for s in seqs], [s.accession or None for s in seqs],
                              [s.description or None for s in seqs], path, all_consensus_cols=all_cols)
        rh, rd, rtext, rc = ref.search([s.sequence for s in seqs])
    hits = (_lib.HitRec * len(rh))()
    doms = (_lib.DomainRec * len(rd))()
    for a, r in zip(hits, rh):
        a.profile = 0
        for fld in ("seq", "score", "pre_score", "sum_score", "nexpected", "lnP", "pre_lnP", "sum_lnP", "nregions", "nclustered", "noverlaps",
                    "nenvelopes", "ndom", "best_domain", "dom_offset"):
            setattr(a, fld, getattr(r, fld))
    for a, r in zip(doms, rd):
        for fld in ("ienv", "jenv", "iali", "jali", "envsc", "domcorrection", "dombias", "oasc", "bitscore", "lnP", "hmmfrom", "hmmto", "sqfrom",
                    "sqto", "N", "text_offset"):
            setattr(a, fld, getattr(r, fld))
    pli = object.__new__(plan7.Pipeline)
    for k, v in dict(alphabet=amino, background=plan7.Background(amino), bias_filter=True, null2=True, seed=42, Z=None, domZ=None, F1=0.02,
                     F2=1e-3, F3=1e-5, E=10.0, T=None, domE=10.0, domT=None, incE=0.01, incT=None, incdomE=0.01, incdomT=None, bit_cutoffs=None,
                     host_threads=1).items():
        setattr(pli, k, v)
    pli.clear()
    om = plan7.Profile(hmm.M, amino).configure(hmm, pli.background, 400).to_optimized()
    order = sorted(range(len(hits)), key=lambda i: hits[i].seq)
    th = pli._assemble([hmm], [om], seqs, [hits[i] for i in order], list(doms), rtext, np.array([rc], np.int64).reshape(1, 4))[0]
    msa = th.to_msa(amino, all_consensus_cols=all_cols)
    names, rows, pp, gs, gc = _read_pfam_stockholm(path)
    assert nref == len(names) == len(msa.names) >= 2
    assert [n.decode() for n in msa.names] == names
    for n, row, p, d, a in zip(names, msa.alignment, msa.posterior_probabilities, msa.descriptions, msa.accessions):
        assert row == rows[n], (n, row, rows[n])
        assert p == pp[n], (n, p, pp[n])
        assert d.decode() == gs[(n, "DE")] and (a is None or a.decode() == gs[(n, "AC")])
    assert msa.reference == gc["RF"] and msa.consensus_posterior_probabilities == gc["PP_cons"]
    assert len(msa) == len(gc["RF"]) and (not all_cols or gc["RF"].count("x") == hmm.M)
    buf = io.BytesIO()
    msa.write(buf, "stockholm")
    again = tmp_path / "mine.sto"
    again.write_bytes(buf.getvalue())
    assert _read_pfam_stockholm(str(again))[:3] == (names, rows, pp)
    # digital mode, and an extra row placed first (what jackhmmer does with its query: plan7.pyx:4369)
    dig = th.to_msa(amino, all_consensus_cols=all_cols, digitize=True)
    assert isinstance(dig, easel.DigitalMSA) and [amino.decode(r) for r in dig.ax] == [r.upper().replace(".", "-") for r in msa.alignment]
    extra = easel.DigitalSequence(amino, name=b"query", sequence=np.arange(hmm.M, dtype=np.uint8) % 20)
    withq = th.to_msa(amino, sequences=[extra], traces=[plan7.Trace.from_sequence(extra)], all_consensus_cols=True)
    assert withq.names[0] == b"query" and withq.names[1:] == th.to_msa(amino, all_consensus_cols=True).names
    assert withq.alignment[0].replace(".", "") == amino.decode(extra.sequence) and withq.posterior_probabilities[0] is None
    with pytest.raises(ValueError):
        th.to_msa(amino, sequences=[extra], traces=[])


@needs_ref
def test_to_msa_on_synthetic_homologs(amino, tmp_path):
    """The same comparison on seeded random models whose emitted homologs carry insertions and deletions of every length
    (insertions longer than one are split left / right, columns used only by deletions collapse)."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import ModelPair
    rng = np.random.default_rng(21)
    checked = ins2 = 0
    for M in (12, 45, 130, 260):
        h = synth.random_hmm(amino, M, rng, name="msa%d" % M)
        h._evparam[:] = np.array([-8.0 - np.log2(M) * 0.3, 0.70, -9.0, 0.70, -4.0, 0.70], np.float32)
        pair = ModelPair(h)
        seqs = []
        for i in range(30):
            L = int(rng.integers(M + 20, 4 * M + 200))
            codes = rng.integers(0, 20, L).astype(np.uint8)
            for _ in range(int(rng.integers(1, 3))):
                dom = synth.emit_sequence(pair.hmm, rng)
                if len(dom) < L:
                    pos = int(rng.integers(0, L - len(dom)))
                    codes[pos:pos + len(dom)] = dom
            seqs.append(easel.DigitalSequence(amino, name=b"t%d" % i, sequence=codes))
        block = easel.DigitalSequenceBlock(amino, seqs)
        path = str(tmp_path / ("ref%d.sto" % M))
        nref = pair.ref.search_msa([s.sequence for s in seqs], [s.name for s in seqs], [None] * len(seqs), [None] * len(seqs), path)
        rh, rd, rtext, rc = pair.ref.search([s.sequence for s in seqs])
        hits = (_lib.HitRec * len(rh))()
        doms = (_lib.DomainRec * len(rd))()
        for a, r in zip(hits, rh):
            a.profile = 0
            for fld in ("seq", "score", "pre_score", "sum_score", "nexpected", "lnP", "pre_lnP", "sum_lnP", "nregions", "nclustered",
                        "noverlaps", "nenvelopes", "ndom", "best_domain", "dom_offset"):
                setattr(a, fld, getattr(r, fld))
        for a, r in zip(doms, rd):
            for fld in ("ienv", "jenv", "iali", "jali", "envsc", "domcorrection", "dombias", "oasc", "bitscore", "lnP", "hmmfrom", "hmmto",
                        "sqfrom", "sqto", "N", "text_offset"):
                setattr(a, fld, getattr(r, fld))
        pli = object.__new__(plan7.Pipeline)
        for k, v in dict(alphabet=amino, background=pair.bg, bias_filter=True, null2=True, seed=42, Z=None, domZ=None, F1=0.02, F2=1e-3,
                         F3=1e-5, E=10.0, T=None, domE=10.0, domT=None, incE=0.01, incT=None, incdomE=0.01, incdomT=None, bit_cutoffs=None,
                         host_threads=1).items():
            setattr(pli, k, v)
        pli.clear()
        order = sorted(range(len(hits)), key=lambda i: hits[i].seq)
        th = pli._assemble([pair.hmm], [pair.om], block, [hits[i] for i in order], list(doms), rtext, np.array([rc], np.int64).reshape(1, 4))[0]
        if nref == 0:                                    # nothing included: eslFAIL there, ValueError here
            with pytest.raises(ValueError):
                th.to_msa(amino)
            continue
        msa = th.to_msa(amino)
        names, rows, pp, gs, gc = _read_pfam_stockholm(path)
        assert nref == len(names) == len(msa.names) >= 5 and [n.decode() for n in msa.names] == names
        for n, row, p in zip(names, msa.alignment, msa.posterior_probabilities):
            assert row == rows[n] and p == pp[n], (M, n, row, rows[n])
            ins2 += any(len(run) > 1 for run in "".join(c if c.islower() else " " for c in row).split())
        assert msa.reference == gc["RF"] and msa.consensus_posterior_probabilities == gc["PP_cons"]
        checked += len(names)
    assert checked >= 30 and ins2 >= 3


@needs_ref
def test_scan_tables_match_the_reference_writers(amino, tmp_path, monkeypatch):
    """`TopHits.write` in scan mode (targets are models, the query is a sequence; tlen / qlen swap, Z = number of models):
    the reference's raw hits of every (model, sequence) comparison go through the product's scan assembly and come out
    as the reference's own tables, byte for byte."""
    import io
    with easel.SequenceFile(os.path.join(GOLD, "data", "proteome.faa.gz"), digital=True, alphabet=amino) as f:
        seqs = {s.name: s for s in f.read_block()}
    query = seqs["938293.PRJEB85.HG003691_78"]          # hit by several RREFam models
    with tempfile.NamedTemporaryFile(suffix=".hmm") as tmp:
        with gzip.open(os.path.join(GOLD, "data", "RREFam.hmm.gz")) as f:
            tmp.write(f.read())
        tmp.flush()
        with plan7.HMMFile(tmp.name) as f:
            hmms = list(f)
        refs = [refshim.RefModel(tmp.name, i, 400) for i in range(len(hmms))]
        prefix = str(tmp_path / "scan")
        nref = refshim.scan_tables(refs, query.sequence, query.name, query.accession or None, query.description or None, prefix)
        hits, doms, text = [], [], b""
        for t, ref in enumerate(refs):
            rh, rd, rtext, rc = ref.search([query.sequence])
            for r in rh:
                a = _lib.HitRec()
                a.profile, a.seq = t, 0
                for fld in ("score", "pre_score", "sum_score", "nexpected", "lnP", "pre_lnP", "sum_lnP", "nregions", "nclustered", "noverlaps",
                            "nenvelopes", "ndom", "best_domain"):
                    setattr(a, fld, getattr(r, fld))
                a.dom_offset = len(doms)
                for rdom in rd[r.dom_offset:r.dom_offset + r.ndom]:
                    d = _lib.DomainRec()
                    for fld in ("ienv", "jenv", "iali", "jali", "envsc", "domcorrection", "dombias", "oasc", "bitscore", "lnP", "hmmfrom", "hmmto",
                                "sqfrom", "sqto", "N"):
                        setattr(d, fld, getattr(rdom, fld))
                    d.text_offset = len(text) + rdom.text_offset
                    doms.append(d)
                hits.append(a)
            text += rtext
    assert nref >= 1 and len(hits) >= nref
    pli = object.__new__(plan7.Pipeline)
    for k, v in dict(alphabet=amino, background=plan7.Background(amino), bias_filter=True, null2=True, seed=42, Z=None, domZ=None, F1=0.02,
                     F2=1e-3, F3=1e-5, E=10.0, T=None, domE=10.0, domT=None, incE=0.01, incT=None, incdomE=0.01, incdomT=None, bit_cutoffs=None,
                     host_threads=1).items():
        setattr(pli, k, v)
    pli.clear()
    monkeypatch.setattr(plan7.Pipeline, "_run", lambda self, oms, block, seq_counters=False: (hits, doms, text, np.zeros((len(block) if seq_counters else len(oms), 4), np.int64)))
    th = pli._scan_many([query], hmms)[0]
    assert th.mode == "scan" and th.Z == float(len(hmms)) and len(th) == nref
    for fmt, ext in (("targets", ".tbl"), ("domains", ".domtbl"), ("pfam", ".pfam")):
        buf = io.BytesIO()
        th.write(buf, format=fmt)
        want = open(prefix + ext, "rb").read()
        assert buf.getvalue() == want, (fmt, buf.getvalue()[:900], want[:900])


@pytest.mark.parametrize("name", ["PF02826", "Thioesterase", "LuxC"])
def test_model_statistics_reproduce_what_hmmbuild_stored(amino, name):
    """`HMM.set_composition` / `set_consensus` recompute the COMPO line and the consensus annotation hmmbuild wrote into the
    reference's own model files (p7_hmm_SetComposition, p7_hmm_SetConsensus), to the precision of the file."""
    with gzip.open(os.path.join(GOLD, "data", name + ".hmm.gz")) as f:
        hmm = plan7.HMMFile(f).read()
    stored_compo, stored_cons = hmm._compo.copy(), hmm.consensus
    hmm.set_composition()
    hmm.set_consensus()
    if stored_compo[:20].any():                          # (a model file may come without a COMPO line)
        assert np.allclose(-np.log(hmm._compo[:20]), -np.log(stored_compo[:20]), atol=2e-5, rtol=0)
    else:
        assert abs(float(hmm._compo[:20].sum()) - 1.0) < 1e-5
    same = sum(a == b for a, b in zip(hmm.consensus, stored_cons))
    assert len(hmm.consensus) == hmm.M and same >= hmm.M - 2          # (a probability at the 0.5 threshold may round either way in the file)
    occ = hmm.match_occupancy()
    assert occ.shape == (hmm.M + 1,) and occ[0] == 0 and np.all((occ[1:] > 0) & (occ[1:] <= 1.0 + 1e-6))
    assert hmm.to_profile(L=123).L == 123 and 0.1 < hmm.mean_match_relative_entropy(plan7.Background(amino)) < 5.0

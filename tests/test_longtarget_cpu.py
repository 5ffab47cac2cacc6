"""CPU-only checks of the long-target (nhmmer) window stages.

1. The oracle is pinned: `ref_longtarget_stages` (the reference's static postSSV / postViterbi functions restated with its
   public calls) leaves the same pos_past_* counters as the real p7_Pipeline_LongTarget, and the scalar port of
   p7_ViterbiFilter_longtarget (oracle/hmmer_oracle.c) records the same landmarks in the same order as the SSE original.
2. The product's HOST logic (`pyhmmer_b200.longtarget.stages`: gates, B1/B2/B3 bias scaling in the reference's precision,
   counters; `b2h_longtarget_vit_finish`: landmark order, extend / merge, 80 kb cut; `b2h_longtarget_vit_threshold`) gives
   the reference's intermediates when every DP score is supplied by the reference's functions -- no device involved.
"""
import ctypes
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pyhmmer_b200 import _lib, longtarget
from oracle import port
import lt_common
from conftest import ModelPair


@pytest.mark.parametrize("M,mu_shift", [(60, 0.0), (60, -3.0), (333, -2.0)])
def test_restated_stages_match_the_real_pipeline(M, mu_shift):
    pair, rng = lt_common.dna_model(ModelPair, M, mu_shift=mu_shift)
    block = lt_common.dna_chunks(pair, rng, [120000, 30000], nplant=8)
    for s in block:
        st = pair.ref.longtarget_stages(s.sequence)
        cnt, hits = pair.ref.longtarget_pipeline(s.sequence)
        assert np.array_equal(st["counters"], cnt[:4]), (st["counters"], cnt)
        assert st["counters"][0] > 0 and cnt[4] >= 1 and len(st["vithit"]) > 0


@pytest.mark.parametrize("M", [5, 9, 17, 60, 333, 1100])
def test_port_of_viterbi_longtarget(M):
    pair, rng = lt_common.dna_model(ModelPair, M)
    pt = port.Port(pair.om)
    total = 0
    for L in (1, 7, 300, 3000):
        seq = lt_common.dna_chunks(pair, rng, [L], nplant=2)[0].sequence
        for filtersc in (-3.0, -8.0, -12.0):
            cfg = min(L, pair.hmm.max_length)
            a = pair.ref.vit_longtarget(seq, cfg, filtersc)
            b = pt.vit_longtarget(seq, cfg, filtersc)
            assert np.array_equal(a, b), (M, L, filtersc, len(a), len(b))
            total += len(a)
            # the threshold the host library derives is the one the port derives
            thr, xwm = ctypes.c_int32(), ctypes.c_int32()
            o = ctypes.c_void_p()
            assert _lib.lib.b2h_profile_create_host(ctypes.byref(pair.om._desc), ctypes.byref(o)) == 0
            assert _lib.lib.b2h_longtarget_vit_threshold(o, cfg, filtersc, 3e-3, ctypes.byref(thr), ctypes.byref(xwm)) == 0
            _lib.lib.b2h_profile_destroy(o)
            assert (thr.value, xwm.value) == pt.vit_longtarget_threshold(cfg, filtersc, 3e-3)
    assert total > 0


@pytest.mark.parametrize("M,mu_shift,bias_filter", [(60, -3.0, True), (121, -2.0, True), (333, -2.0, False)])
def test_host_logic_with_reference_scores(M, mu_shift, bias_filter):
    pair, rng = lt_common.dna_model(ModelPair, M, mu_shift=mu_shift)
    block = lt_common.dna_chunks(pair, rng, [60000, 0, 9, 25000], nplant=6)
    kw = dict(F1=0.02, F2=3e-3, F3=3e-5, bias_filter=bias_filter, B1=100, B2=240, B3=1000)
    got = longtarget.stages(pair.om, block, backend=lt_common.OracleBackend(pair, block), **kw)
    tot = lt_common.compare_with_reference(pair, block, got, exact_scores=True, **kw)
    assert tot["msvwin"] >= 5 and tot["vitmark"] >= 5 and tot["vitwin"] >= 3 and tot["passed"] >= 1, tot


def test_long_windows_are_cut_at_80kb():
    """b2h_longtarget_vit_finish against the reference's loop (p7_pipeline.c:1389-1406) on a window long enough to be cut:
    a repeat-rich region whose landmarks merge into one window above 80 kb."""
    M = 40
    pair, rng = lt_common.dna_model(ModelPair, M, mu_shift=-3.0)
    dom = lt_common.synth.emit_sequence(pair.hmm, rng)
    reps = 200000 // len(dom)
    seq = np.concatenate([dom] * reps).astype(np.uint8)
    dna = pair.hmm.alphabet
    block = lt_common.easel.DigitalSequenceBlock(dna, [lt_common.easel.DigitalSequence(dna, name=b"rep", sequence=seq)])
    kw = dict(F1=0.02, F2=3e-3, F3=3e-5, bias_filter=True, B1=100, B2=240, B3=1000)
    got = longtarget.stages(pair.om, block, backend=lt_common.OracleBackend(pair, block), **kw)
    tot = lt_common.compare_with_reference(pair, block, got, exact_scores=True, **kw)
    assert got["vitwin"]["length"].max() == 80000 and tot["vitwin"] >= 3, (tot, got["vitwin"]["length"])


@pytest.mark.parametrize("M,mu_shift,null2,complement,layout", [(60, -3.0, True, False, "plain"), (121, -2.0, True, True, "plain"),
                                                                 (121, -2.0, False, False, "plain"), (333, -2.0, True, False, "plain"),
                                                                 (60, -3.0, True, False, "tandem"), (121, -2.0, True, True, "tandem"),
                                                                 (60, -3.0, True, False, "repeats"), (121, -2.0, True, False, "repeats"),
                                                                 (121, -2.0, False, True, "repeats")])
def test_hits_behind_the_forward_gate(M, mu_shift, null2, complement, layout):
    """b2h_longtarget_domains (the long-target branches of domain definition + the hit arithmetic of
    p7_pli_postViterbi_LongTarget) fed with the reference's parser specials for the windows that passed the Forward gate
    reproduces the hits of the real p7_Pipeline_LongTarget: coordinates exactly, scores to 2e-3 bits."""
    pair, rng = lt_common.dna_model(ModelPair, M, mu_shift=mu_shift)
    block = lt_common.dna_chunks(pair, rng, [90000], nplant=10)
    seq = block[0].sequence
    if layout == "repeats":
        # a homolog flanked by runs of a short piece of itself: the envelope reaches far beyond the alignment and the
        # reference trims it to the alignment +- 20 and rescores (p7_domaindef.c:893-935)
        seq = seq.copy()
        pos = 500
        for rep in range(24):
            dom = lt_common.synth.emit_sequence(pair.hmm, rng)
            frag, nrep = ((6, 10), (8, 8), (10, 6))[rep % 3]
            a = int(rng.integers(5, max(6, len(dom) - frag - 5)))
            tail = np.concatenate([dom[a:a + frag]] * nrep)
            piece = np.concatenate([tail, dom, tail])
            seq[pos:pos + len(piece)] = piece
            pos += len(piece) + 3000
        block = lt_common.easel.DigitalSequenceBlock(block.alphabet, [lt_common.easel.DigitalSequence(block.alphabet, name=b"r", sequence=seq)])
    if layout == "tandem":
        # back-to-back and overlapping copies (multi-domain regions -> stochastic clustering) and homologs continued by a
        # degraded copy of themselves (envelopes far wider than the alignment -> the trimming pass)
        seq = seq.copy()
        pos = 500
        for rep in range(6):
            doms = [lt_common.synth.emit_sequence(pair.hmm, rng) for _ in range(3)]
            weak = doms[2].copy()
            mask = rng.random(len(weak)) < 0.45
            weak[mask] = rng.integers(0, 4, int(mask.sum()))
            piece = np.concatenate([doms[0], rng.integers(0, 4, rep).astype(np.uint8), doms[1], doms[2][:len(doms[2]) // 2], weak])
            seq[pos:pos + len(piece)] = piece
            pos += len(piece) + 3000
        block = lt_common.easel.DigitalSequenceBlock(block.alphabet, [lt_common.easel.DigitalSequence(block.alphabet, name=b"t", sequence=seq)])
    start = 1001
    if complement:                                        # <seq> plays the chunk as esl_sq_ReverseComplement leaves it:
        start = 1001 + len(seq) - 1                       # sq->start is then the chunk's LAST coordinate
    kw = dict(F1=0.02, F2=3e-3, F3=3e-5, bias_filter=True, B1=100, B2=240, B3=1000)
    got = longtarget.stages(pair.om, block, backend=lt_common.OracleBackend(pair, block), **kw)
    cnt, rhits, rtext = pair.ref.longtarget_pipeline(seq, null2=null2, start=start, complement=complement, want_text=True)
    assert np.array_equal(got["counters"][0], cnt[:4]) and len(rhits) >= 3
    mw, vw = got["msvwin"], got["vitwin"]
    keep = [], []
    wins = (_lib.LtWindow * int(got["vitpass"].sum()))()
    q = 0
    for v in np.flatnonzero(got["vitpass"]):
        n0 = int(mw["n"][vw["seq"][v]] + vw["n"][v] - 1)            # start of the window in the chunk, 1-based
        L = int(vw["length"][v])
        codes = np.ascontiguousarray(seq[n0 - 1:n0 - 1 + L])
        _, _, st, fx, bx = pair.ref.fwdbck(codes, want_x=True)
        assert st == 0
        keep[0].append((codes, fx, bx))
        w = wins[q]; q += 1
        w.dsq, w.L, w.fwd_xmx, w.bck_xmx = codes.ctypes.data, L, fx.ctypes.data, bx.ctypes.data
        w.window_start, w.seq_start, w.complement, w.seq = n0, start, int(complement), 0
    hp = ctypes.c_void_p()
    assert _lib.lib.b2h_profile_create_host(ctypes.byref(pair.om._desc), ctypes.byref(hp)) == 0
    _lib.lib.b2h_profile_set_annotation(hp, (pair.hmm.consensus or "x" * M).encode(), None, None, pair.hmm.alphabet.symbols.encode())
    prm = _lib.SearchParams(0.02, 3e-3, 3e-5, 1, int(null2), 42, 1)
    out = ctypes.c_void_p()
    _lib.check(_lib.lib.b2h_longtarget_domains(hp, wins, len(wins), ctypes.byref(prm), ctypes.byref(out)), "b2h_longtarget_domains")
    hits, doms, text = _lib.read_results(out)
    _lib.lib.b2h_results_destroy(out)
    _lib.lib.b2h_profile_destroy(hp)
    assert len(hits) == len(rhits), (len(hits), len(rhits))
    nbias = 0
    for h, r, rt in zip(hits, rhits, rtext):
        d = doms[h.dom_offset]
        assert (d.ienv, d.jenv, d.iali, d.jali, d.hmmfrom, d.hmmto) == tuple(int(x) for x in (r[0], r[1], r[2], r[3], r[10], r[11])), (h.profile, r)
        assert tuple(text[d.text_offset:d.text_offset + 4 * (d.N + 1)].split(b"\0")[:4]) == rt       # the alignment display, character by character
        assert (d.sqfrom, d.sqto) == (d.iali, d.jali)
        assert abs(h.score - r[4]) < 2e-3 and abs(d.dombias - r[5]) < 2e-3 and abs(h.pre_score - r[6]) < 2e-3, (h.score, r)
        assert abs(h.lnP - r[7]) < 2e-3 and abs(d.envsc - r[8]) < 2e-3 and abs(d.oasc - r[9]) < 2e-3
        assert h.ndom == 1 and d.bitscore == h.score and h.sum_score == h.score
        nbias += d.dombias > 0
    assert null2 or nbias == 0
    if layout == "repeats":                               # the trimming pass ran: an envelope edge sits exactly 20 from the alignment
        assert any(abs(int(r[2]) - int(r[0])) == 20 or abs(int(r[1]) - int(r[3])) == 20 for r in rhits)


@pytest.mark.parametrize("M,mu_shift,strand,block_length", [(121, -2.0, None, 65536), (60, -3.0, "watson", 20000), (60, -3.0, "crick", 0x40000),
                                                             (333, -2.0, None, 30000)])
def test_whole_search_against_the_reference_loop(M, mu_shift, strand, block_length):
    """`longtarget.search` (windows with context over every target, both strands, all stages, E-values over residues / window
    length, duplicate removal across window overlaps) with the reference's DP scores against ref_nhmmer."""
    pair, rng = lt_common.dna_model(ModelPair, M, mu_shift=mu_shift)
    block = lt_common.dna_chunks(pair, rng, [150000, 40000, 0, 700, 65536 + 17], nplant=8)
    seqs = []
    for s in block:                                       # homologs on the other strand as well
        codes = s.sequence.copy()
        for _ in range(4):
            dom = longtarget.reverse_complement(s.alphabet, lt_common.synth.emit_sequence(pair.hmm, rng))
            if len(dom) < len(codes):
                pos = int(rng.integers(0, len(codes) - len(dom)))
                codes[pos:pos + len(dom)] = dom
        seqs.append(lt_common.easel.DigitalSequence(s.alphabet, name=s.name, sequence=codes))
    # one homolog per strand inside the context the second window of the first target shares with the first window
    C = pair.hmm.max_length
    if block_length < len(seqs[0]):
        codes = seqs[0].sequence
        pos = block_length - C + 5
        for dom in (lt_common.synth.emit_sequence(pair.hmm, rng), longtarget.reverse_complement(block.alphabet, lt_common.synth.emit_sequence(pair.hmm, rng))):
            codes[pos:pos + len(dom)] = dom
            pos += len(dom) + 12
    got = longtarget.search(pair.om, seqs, block_length=block_length, strand=strand,
                            backend_factory=lambda om, blk: lt_common.OracleBackend(pair, blk))
    nh, ndup = lt_common.compare_nhmmer(pair, [s.sequence for s in seqs], got, block_length=block_length, strand=strand)
    assert nh >= 10
    if block_length < 100000:
        assert ndup >= 1                                  # a hit in the context shared by two windows was found twice and removed once


@pytest.mark.parametrize("strand,E", [(None, 10.0), ("watson", 1e-12)])
def test_long_targets_pipeline_front_end(monkeypatch, tmp_path, strand, E):
    """`plan7.LongTargetsPipeline.search_hmm` (the reference class of the same name, plan7.pyx:6917): hits in the reference's
    final order with its reported / included / duplicate flags and E-values -- host logic only, DP from the reference."""
    from pyhmmer_b200 import plan7
    pair, rng = lt_common.dna_model(ModelPair, 121, mu_shift=-2.0)
    block = lt_common.dna_chunks(pair, rng, [90000, 30000], nplant=10)
    codes = block[0].sequence
    pos = 20000 - pair.hmm.max_length + 5                 # a hit inside the context shared by the first two windows
    dom = lt_common.synth.emit_sequence(pair.hmm, rng)
    codes[pos:pos + len(dom)] = dom
    monkeypatch.setattr(plan7._lib, "context", lambda device=None: None)     # no device: the backend below needs none
    pli = plan7.LongTargetsPipeline(pair.hmm.alphabet, strand=strand, block_length=20000, E=E, incE=E / 100)
    pli._backend_factory = lambda om, blk: lt_common.OracleBackend(pair, blk)
    th = pli.search_hmm(pair.hmm, block)
    rhits, rstats = pair.ref.nhmmer([s.sequence for s in block], block_length=20000, strand=strand, E=E, incE=E / 100,
                                    evalue_window=pair.ref.max_length(),     # an HMM query: E-values count windows of p7_Builder_MaxLength
                                    names=[s.name for s in block], table_prefix=str(tmp_path / "ref"))
    assert th.long_targets and len(th) == len(rhits) >= 8
    assert (th.searched_residues, th.searched_sequences) == (rstats[0], rstats[1])
    for h, r in zip(th, rhits):
        d = h.domains[0]
        assert (h.name, d.env_from, d.env_to, d.alignment.target_from, d.alignment.target_to, d.alignment.hmm_from, d.alignment.hmm_to) == \
               (block[r.seqidx].name, r.ienv, r.jenv, r.iali, r.jali, r.hmmfrom, r.hmmto)
        assert abs(h.score - r.score) < 2e-3 and abs(math.log(h.evalue) - r.lnP) < 2e-3
        assert (h.reported, h.included, h.duplicate) == (bool(r.flags & 2), bool(r.flags & 1), bool(r.flags & 16)), (h.name, r.flags)
        assert (d.reported, d.included) == (bool(r.dom_reported), bool(r.dom_included))
    assert any(h.duplicate for h in th) and any(h.reported for h in th)
    assert E > 1 or any(not h.reported and not h.duplicate for h in th)      # the tight threshold left some hits unreported
    import io
    for fmt, ext in (("targets", ".tbl"), ("pfam", ".pfam")):               # nhmmer's tables, byte for byte the reference writers'
        buf = io.BytesIO()
        th.write(buf, format=fmt)
        assert buf.getvalue() == open(str(tmp_path / "ref") + ext, "rb").read(), (fmt, buf.getvalue()[:800])
    # a user-set Z replaces the residues searched by 1e6 Z per strand in the E-values
    pz = plan7.LongTargetsPipeline(pair.hmm.alphabet, strand=strand, block_length=20000, E=E, incE=E / 100, Z=3.0)
    pz._backend_factory = pli._backend_factory
    tz = pz.search_hmm(pair.hmm, block)
    shift = math.log(3e6 * (2 if strand is None else 1) / th.searched_residues)
    byname = {(h.name, h.domains[0].env_from): h for h in th}
    assert len(tz) == len(th) and all(abs((h.lnP - byname[(h.name, h.domains[0].env_from)].lnP) - shift) < 1e-5 for h in tz)
    with pytest.raises(ValueError):
        plan7.LongTargetsPipeline(plan7.Alphabet.amino())
    with pytest.raises(ValueError):
        plan7.LongTargetsPipeline(pair.hmm.alphabet, strand="both")
    with pytest.raises(ValueError):
        plan7.LongTargetsPipeline(pair.hmm.alphabet, block_length=100).search_hmm(pair.hmm, block)


def test_max_length_matches_the_builder(amino):
    """HMM.compute_max_length = p7_Builder_MaxLength (p7_builder.c:651) for synthetic and for the reference's own models,
    whose MAXL line hmmbuild wrote with the default tail mass."""
    import gzip
    import tempfile
    from pyhmmer_b200 import plan7
    from oracle import refshim
    for M in (1, 2, 3, 40, 333):
        pair, rng = lt_common.dna_model(ModelPair, M)
        for beta in (1e-7, 1e-3, 0.3):
            assert pair.hmm.compute_max_length(beta) == pair.ref.max_length(beta), (M, beta)
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")
    for name in ("PF02826", "Thioesterase"):
        with tempfile.NamedTemporaryFile(suffix=".hmm") as tmp:
            with gzip.open(os.path.join(gold, name + ".hmm.gz")) as f:
                tmp.write(f.read())
            tmp.flush()
            with plan7.HMMFile(tmp.name) as f:
                hmm = f.read()
            ref = refshim.RefModel(tmp.name, 0, 400)
            assert hmm.compute_max_length() == ref.max_length() and (hmm.max_length <= 0 or hmm.compute_max_length() == hmm.max_length)


_LT_WORKER = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch.distributed as dist
from pyhmmer_b200 import parallel, longtarget
import lt_common
from conftest import ModelPair
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%d" %% int(sys.argv[1]), rank=int(sys.argv[2]), world_size=2)
w = parallel.World.current()
pair, rng = lt_common.dna_model(ModelPair, 121, mu_shift=-2.0)
block = lt_common.dna_chunks(pair, rng, [90000, 30000, 500], nplant=10)
fac = lambda om, blk: lt_common.OracleBackend(pair, blk)
kw = dict(block_length=20000, backend_factory=fac)
sharded = longtarget.search(pair.om, block, world=w, **kw)
single = longtarget.search(pair.om, block, world=parallel.World(), **kw)
key = lambda res: [(h.seq, res[1][h.dom_offset].iali, res[1][h.dom_offset].jali, res[1][h.dom_offset].ienv, res[1][h.dom_offset].jenv,
                    h.score, h.lnP, res[1][h.dom_offset].dombias, d) for h, d in zip(res[0], res[3])]
assert len(single[0]) >= 10 and key(sharded) == key(single), (len(sharded[0]), len(single[0]))
assert sharded[4] == single[4], (sharded[4], single[4])
for res in (sharded, single):
    for h in res[0]:
        d = res[1][h.dom_offset]
        assert len(res[2][d.text_offset:d.text_offset + 4 * (d.N + 1)].split(b"\0")[0]) == d.N
dist.barrier(); dist.destroy_process_group()
print("rank", w.rank, "ok", len(single[0]))
'''


def test_window_sharding_world_size_2_gloo():
    """The multi-GPU form of the nhmmer search: windows dealt to two ranks, one all-gather of the hit records, then the
    E-value / duplicate pass on every rank -- identical to the unsharded search (host logic; gloo on the CPU)."""
    import socket
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_ = s.getsockname()[1]; s.close()
    code = _LT_WORKER % (root, root)
    procs = [subprocess.Popen([sys.executable, "-c", code, str(port_), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_model_mask_in_the_reestimated_background():
    """A model with an MM line: masked nodes keep a zero match score when the long-target domain definition re-estimates
    the background (p7_oprofile_UpdateFwdEmissionScores, impl_sse/p7_oprofile.c:455) -- hits identical to ref_nhmmer."""
    pair, rng = lt_common.dna_model(ModelPair, 121, mu_shift=-2.0, mask=True)
    assert pair.hmm.model_mask and "m" in pair.hmm.model_mask and pair.om.model_mask == pair.hmm.model_mask
    block = lt_common.dna_chunks(pair, rng, [90000, 30000], nplant=10)
    got = longtarget.search(pair.om, block, block_length=0x40000, backend_factory=lambda om, blk: lt_common.OracleBackend(pair, blk))
    nh, _ = lt_common.compare_nhmmer(pair, [s.sequence for s in block], got, block_length=0x40000)
    assert nh >= 8
    # ... and the mask matters: without it the same search gives different bias corrections
    pair.hmm.model_mask = None
    plain = longtarget.search(pair.om, block, block_length=0x40000, backend_factory=lambda om, blk: lt_common.OracleBackend(pair, blk))
    assert [d.dombias for d in plain[1]] != [d.dombias for d in got[1]]


def test_randomised_searches_against_the_reference_loop():
    """A seeded sweep over model lengths, score scales, block lengths, strands, null2 / bias-filter switches, degenerate
    residues and empty targets: `longtarget.search` (host logic, reference DP scores) against ref_nhmmer.  (625 cases of
    this sweep with other seeds, 5 615 hits, were compared without a difference while it was written.)"""
    from pyhmmer_b200 import parallel
    rng0 = np.random.default_rng(12345)
    total = 0
    for case in range(14):
        M = int(rng0.choice([12, 25, 40, 77, 121, 200, 333]))
        shift = float(rng0.choice([0.0, -1.0, -2.0, -3.0, -4.0]))
        seed = int(rng0.integers(0, 10**6))
        bl = int(rng0.choice([3000, 20000, 65536, 262144]))
        strand = [None, "watson", "crick"][int(rng0.integers(0, 3))]
        null2, biasf = bool(rng0.integers(0, 2)), bool(rng0.integers(0, 4) > 0)
        pair, rng = lt_common.dna_model(ModelPair, M, seed=seed, mu_shift=shift)
        bl = max(bl, pair.hmm.max_length * 3)
        sizes = [int(rng.integers(1, 60000)), int(rng.integers(1, 20000)), int(rng.integers(0, 50))]
        block = lt_common.dna_chunks(pair, rng, sizes, nplant=int(rng.integers(0, 12)))
        seqs = []
        for s in block:
            codes = s.sequence.copy()
            for _ in range(int(rng.integers(0, 5))):
                dom = longtarget.reverse_complement(s.alphabet, lt_common.synth.emit_sequence(pair.hmm, rng))
                if len(dom) < len(codes):
                    pos = int(rng.integers(0, len(codes) - len(dom)))
                    codes[pos:pos + len(dom)] = dom
            if len(codes) > 100:
                codes[rng.integers(0, len(codes), 5)] = rng.integers(5, 16, 5)          # R Y M K S W H B V D N
            seqs.append(lt_common.easel.DigitalSequence(s.alphabet, name=s.name, sequence=codes))
        kw = dict(block_length=bl, strand=strand, null2=null2, bias_filter=biasf)
        got = longtarget.search(pair.om, seqs, world=parallel.World(), backend_factory=lambda om, blk: lt_common.OracleBackend(pair, blk), **kw)
        nh, _ = lt_common.compare_nhmmer(pair, [s.sequence for s in seqs], got, **kw)
        total += nh
    assert total >= 30


def test_window_residues_gather():
    """`longtarget.window_residues` (the C packer behind every window database of the long-target stages) against a plain
    numpy gather: ragged windows of a multi-sequence block, windows that touch both ends of their sequence, zero-length
    windows, an empty list, and a window outside the packed residues."""
    from pyhmmer_b200 import easel, longtarget
    dna = easel.Alphabet.dna()
    rng = np.random.default_rng(5)
    lens = [70000, 1, 12, 3000000, 257]
    seqs = [easel.DigitalSequence(dna, name=b"s%d" % i, sequence=rng.integers(0, 4, n).astype(np.uint8)) for i, n in enumerate(lens)]
    block = easel.DigitalSequenceBlock(dna, seqs)
    n = 4000
    seq = rng.integers(0, len(lens), n)
    L = np.array([lens[s] for s in seq])
    length = np.minimum(rng.integers(0, 4000, n), L)
    start = np.array([int(rng.integers(1, l - ln + 2)) for l, ln in zip(L, length)])
    start[:50] = 1                                                  # from the first residue ...
    length[50:100] = L[50:100] - start[50:100] + 1                  # ... and to the last
    res, off = longtarget.window_residues(block, seq, start, length)
    assert res.dtype == np.uint8 and len(res) == int(length.sum()) and np.array_equal(off, np.concatenate(([0], np.cumsum(length))))
    for i in range(n):
        assert np.array_equal(res[off[i]:off[i + 1]], seqs[seq[i]].sequence[start[i] - 1:start[i] - 1 + length[i]]), i
    res, off = longtarget.window_residues(block, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64))
    assert len(res) == 0 and list(off) == [0]
    with pytest.raises(IndexError):
        longtarget.window_residues(block, [len(lens) - 1], [200], [100])

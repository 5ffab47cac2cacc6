"""Regenerate the golden fixtures under tests/golden/ from the reference itself.

Run in the build container only (needs /root/reference and the reference pyhmmer installed into
baseline/_ref by `pip install --no-index --target baseline/_ref /root/reference`):

    python tests/golden/make_golden.py

Outputs (committed):
  data/*.hmm.gz, data/proteome.faa.gz   input fixtures = the reference's own test data (tests/data/hmms/txt, seqs)
  data/*.tbl, *.domtbl                   the reference's golden tables (produced by the HMMER CLI; tests/data/tables)
  data/pressed/*.h3{f,p,m,i}             the reference's own hmmpress'ed fixture databases (tests/data/hmms/db), copied verbatim
  hmmsearch.json                         pyhmmer.hmmsearch results (every hit/domain field, full precision) for each
                                         fixture HMM against the proteome, plus pipeline pass counters
  filters.json                           per-stage scores from pyhmmer (OptimizedProfile.msv_filter / ssv_filter)
"""
import gzip
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import pyhmmer  # noqa: E402  (the REFERENCE, not pyhmmer_b200)

REF = "/root/reference/src/pyhmmer/tests/data"
HMMS = ["PF02826", "Thioesterase", "KR", "LuxC", "RREFam"]


def _s(v):
    return v.decode() if isinstance(v, bytes) else v


def gz_copy(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)


def main():
    data = os.path.join(HERE, "data")
    os.makedirs(os.path.join(data, "pressed"), exist_ok=True)
    for name in ("Thioesterase", "PF02826"):
        for ext in ("h3f", "h3p", "h3m", "h3i"):
            shutil.copy(os.path.join(REF, "hmms", "db", "%s.hmm.%s" % (name, ext)), os.path.join(data, "pressed"))
    os.makedirs(data, exist_ok=True)
    for h in HMMS:
        gz_copy(os.path.join(REF, "hmms/txt", h + ".hmm"), os.path.join(data, h + ".hmm.gz"))
    gz_copy(os.path.join(REF, "seqs/938293.PRJEB85.HG003687.faa"), os.path.join(data, "proteome.faa.gz"))
    for t in ("PF02826.tbl", "PF02826.domtbl", "RREFam.tbl", "RREFam.domtbl", "RREFam.scan.tbl"):
        shutil.copy(os.path.join(REF, "tables", t), os.path.join(data, t))
        os.chmod(os.path.join(data, t), 0o644)

    abc = pyhmmer.easel.Alphabet.amino()
    with pyhmmer.easel.SequenceFile(os.path.join(REF, "seqs/938293.PRJEB85.HG003687.faa"), digital=True, alphabet=abc) as sf:
        seqs = sf.read_block()
    out = {}
    filt = {}
    for h in HMMS:
        with pyhmmer.plan7.HMMFile(os.path.join(REF, "hmms/txt", h + ".hmm")) as hf:
            hmms = list(hf)
        for hmm in hmms:
            pli = pyhmmer.plan7.Pipeline(abc)
            th = pli.search_hmm(hmm, seqs)
            st = th.__getstate__()["pipeline"] if hasattr(th, "__getstate__") else {}
            rec = dict(M=hmm.M, Z=th.Z, domZ=th.domZ, n_hits=len(th),
                       counters=[st.get("n_past_msv"), st.get("n_past_bias"), st.get("n_past_vit"), st.get("n_past_fwd")],
                       hits=[])
            for hit in th:
                rec["hits"].append(dict(
                    name=_s(hit.name), score=hit.score, pre_score=hit.pre_score, sum_score=hit.sum_score, bias=hit.bias,
                    evalue=hit.evalue, pvalue=hit.pvalue, reported=hit.reported, included=hit.included,
                    domains=[dict(env_from=d.env_from, env_to=d.env_to, score=d.score, bias=d.bias, c_evalue=d.c_evalue,
                                  i_evalue=d.i_evalue, envelope_score=d.envelope_score, reported=d.reported, included=d.included,
                                  hmm_from=d.alignment.hmm_from, hmm_to=d.alignment.hmm_to,
                                  target_from=d.alignment.target_from, target_to=d.alignment.target_to,
                                  hmm_sequence=d.alignment.hmm_sequence, target_sequence=d.alignment.target_sequence,
                                  identity_sequence=d.alignment.identity_sequence,
                                  posterior_probabilities=d.alignment.posterior_probabilities)
                             for d in hit.domains]))
            out[_s(hmm.name)] = rec
        # per-stage MSV/SSV scores for the first model of each file on the first 300 sequences
        hmm = hmms[0]
        prof = pyhmmer.plan7.Profile(hmm.M, abc)
        prof.configure(hmm, pyhmmer.plan7.Background(abc), 400)
        om = prof.to_optimized()
        msv = []
        for s in seqs[:300]:
            om.L = len(s)                      # p7_oprofile_ReconfigLength, as the search loop does per target
            v = om.msv_filter(s)
            msv.append(None if v is None else (float(v) if v == v and abs(v) != float("inf") else str(v)))
        filt[_s(hmm.name)] = dict(file=h, msv=msv)
    with open(os.path.join(HERE, "hmmsearch.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    with open(os.path.join(HERE, "filters.json"), "w") as f:
        json.dump(filt, f, indent=0, sort_keys=True)
    print("wrote", len(out), "searches;", sum(r["n_hits"] for r in out.values()), "hits")


if __name__ == "__main__":
    main()

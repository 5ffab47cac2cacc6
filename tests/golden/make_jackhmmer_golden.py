"""Golden vectors for the jackhmmer path, from the reference itself.

Run in the build container only (needs /root/reference and the reference pyhmmer under baseline/_ref):

    python tests/golden/make_jackhmmer_golden.py

Outputs (committed):
  data/PKSI.faa.gz    the reference's own jackhmmer / phmmer test targets (tests/data/seqs/PKSI.faa)
  jackhmmer.json.gz   `Pipeline.iterate_seq` / `iterate_hmm` of the reference, iteration by iteration (jackhmmer's inclusion
                      thresholds incE = incdomE = 1e-3): hits (name, score, bias, flags, domain coordinates), the alignment
                      handed to the builder (names, rows, RF), the model built from it (length, effective sequence number,
                      the first match emission rows) and the convergence flag; plus `Builder.build_msa` outputs for the
                      alignments (fast and hand architecture) as HMM text.
"""
import gzip
import io
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import pyhmmer  # noqa: E402  (the REFERENCE, not pyhmmer_b200)

REF = "/root/reference/src/pyhmmer/tests/data"


def _s(v):
    return v.decode() if isinstance(v, bytes) else v


def hits_record(hits):
    out = []
    for h in hits:
        out.append({"name": _s(h.name), "score": h.score, "bias": h.bias, "evalue": h.evalue, "included": h.included, "reported": h.reported,
                    "new": h.new, "dropped": h.dropped,
                    "domains": [{"env": [d.env_from, d.env_to], "ali": [d.alignment.target_from, d.alignment.target_to],
                                 "hmm": [d.alignment.hmm_from, d.alignment.hmm_to], "score": d.score, "included": d.included} for d in h.domains]})
    return out


def msa_record(msa, abc):
    t = msa.textize()
    return {"name": _s(msa.name), "names": [_s(n) for n in t.names], "rows": [str(r) for r in t.alignment], "rf": _s(t.reference)}


def main():
    data = os.path.join(HERE, "data")
    with open(os.path.join(REF, "seqs/PKSI.faa"), "rb") as f, gzip.GzipFile(os.path.join(data, "PKSI.faa.gz"), "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)
    abc = pyhmmer.easel.Alphabet.amino()
    bg = pyhmmer.plan7.Background(abc)
    with pyhmmer.easel.SequenceFile(os.path.join(REF, "seqs/PKSI.faa"), digital=True, alphabet=abc) as sf:
        seqs = sf.read_block()
    with pyhmmer.plan7.HMMFile(os.path.join(REF, "hmms/txt/KR.hmm")) as hf:
        kr = hf.read()
    out = {"runs": []}
    for label, query in (("seq:-1", seqs[-1]), ("seq:0", seqs[0]), ("hmm:KR", kr)):
        pli = pyhmmer.plan7.Pipeline(abc, incE=1e-3, incdomE=1e-3)
        it = pli.iterate_hmm(query, seqs) if label.startswith("hmm") else pli.iterate_seq(query, seqs)
        steps = []
        for k in range(5):
            r = next(it)
            buf = io.BytesIO()
            r.hmm.write(buf)
            steps.append({"iteration": r.iteration, "converged": r.converged, "M": r.hmm.M, "nseq": r.hmm.nseq,
                          "nseq_effective": r.hmm.nseq_effective, "hmm": "\n".join(l for l in buf.getvalue().decode().splitlines()
                                                                                   if not l.startswith(("DATE", "COM"))) if r.hmm.M < 400 else None,
                          "evalue_parameters": list(r.hmm.evalue_parameters.as_vector()) if hasattr(r.hmm.evalue_parameters, "as_vector") else None,
                          "match_head": [[float(v) for v in row] for row in list(r.hmm.match_emissions)[1:4]],
                          "hits": hits_record(r.hits), "msa": msa_record(r.msa, abc)})
            if r.converged:
                break
        out["runs"].append({"query": label, "steps": steps})
    with gzip.GzipFile(os.path.join(HERE, "jackhmmer.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(out).encode())
    print("wrote", os.path.join(HERE, "jackhmmer.json.gz"), [(r["query"], len(r["steps"])) for r in out["runs"]])


if __name__ == "__main__":
    main()

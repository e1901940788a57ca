"""Host restatement of the per-taxon read counters -- what k_fold_counts / k_fmt_plan accumulate on the device and
cfb_counts_allreduce sums over the GPUs.  Kept as the checker of those kernels (tests/test_gpu_multi.py) and for the gloo
CPU test of the sharding logic (tests/test_multi_rank.py); the product path does not use it.

`taxon_counts` restates what AlnSinkWrap::finishRead -> SpeciesMetrics::addSpeciesCounts (aln_sink.h:142-156,
1861-1927) accumulate per read -- numReads for every reported assignment, numUniqueReads when exactly one
assignment is reported -- as a vectorised fold over the record arrays a rank got back from cfb_classify_*.
Each rank folds its own shard into a dense int64 vector indexed by tree-node order; one all-reduce(sum)
over the ranks gives the report's numReads / numUniqueReads columns.  (The sparse `observed` tie-set map that
drives the abundance EM is merged on the host, as SpeciesMetrics::merge does.)
"""
import numpy as np


def taxon_counts(node_taxids, rec_off, recs, k=5):
    """node_taxids: sorted uint64 array of tree taxids.  Returns int64 array (len(node_taxids)+1, 2):
    [:, 0] = numReads, [:, 1] = numUniqueReads; the last row collects taxids outside the tree (incl. 0 =
    unclassified).  Reported assignments of a unit = its records with the top score, at most k of them
    (more than k top-scoring records only happens under --host-taxids; then the k reported ones are chosen
    by the per-read RNG on the host, and this fold counts the first k in hit-map order)."""
    n_units = len(rec_off) - 1
    out = np.zeros((len(node_taxids) + 1, 2), dtype=np.int64)
    cnt = np.diff(rec_off.astype(np.int64))
    n_uncl = int((cnt == 0).sum())
    out[-1, 0] += n_uncl
    out[-1, 1] += n_uncl
    if len(recs) == 0:
        return out
    unit = np.repeat(np.arange(n_units), cnt)
    score = recs["score"].astype(np.int64)
    best = np.zeros(n_units, dtype=np.int64)
    np.maximum.at(best, unit, score)
    top = score == best[unit]
    # rank of each top record inside its unit (hit-map order) to apply the k cap
    order = np.zeros(len(recs), dtype=np.int64)
    tops = np.nonzero(top)[0]
    tunit = unit[tops]
    first = np.r_[True, tunit[1:] != tunit[:-1]]
    start = np.maximum.accumulate(np.where(first, np.arange(len(tops)), 0))
    order[tops] = np.arange(len(tops)) - start
    rep = top & (order < k)
    nrep = np.zeros(n_units, dtype=np.int64)
    np.add.at(nrep, unit[rep], 1)
    tax = recs["taxid"][rep]
    pos = np.searchsorted(node_taxids, tax)
    pos = np.where((pos < len(node_taxids)) & (node_taxids[np.minimum(pos, len(node_taxids) - 1)] == tax), pos, len(node_taxids))
    np.add.at(out[:, 0], pos, 1)
    uniq = nrep[unit[rep]] == 1
    np.add.at(out[:, 1], pos[uniq], 1)
    return out
